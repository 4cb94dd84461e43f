"""Callbacks (src/callbacks.jl).  Host-side; they define the device->host sync points:
the block planner (run.__run) ends a block wherever some callback's `check_if_execute`
fires, mirrors the histories on the host, and only then calls `execute_`."""
import os

import numpy as np

from .types import PostMCMCStep, PreMCMCStep
from . import workspaces as W


class Callback:                                           # callbacks.jl:15-85
    def init_(self, ws):                                   # init!
        pass

    def check_if_execute(self, step, flag):
        return False

    def execute_(self, global_ws, local_wss, step, flag):  # execute!
        pass

    def cleanup_(self, ws, local_wss, step):               # cleanup!
        pass


def _find_available_name(path, filename, disambig="", ext=".csv"):   # callbacks.jl:173-180
    while True:
        name = os.path.join(path, f"{filename}{disambig}{ext}")
        if not os.path.isfile(name):
            return name
        disambig = "1" if disambig == "" else str(int(disambig) + 1)


class SavingCallback(Callback):
    """callbacks.jl:125-256.  Row format of the reference (data_to_csv, :246-256), one row
    per (iteration i, update j):
        "i, j, !, th_1, ..., th_p, !, th°_1, ..., th°_p, !, ll, !, ll°, !,accepted, \\n"
    With n_chains > 1 one file per chain is written (`<name>_chain<c>.csv`), each in exactly
    that format.  Quirks kept: the end-of-run save stops at iteration M - 1 (:223 with
    run.jl:51); the ll° column is 0.0 (the reference never fills sub_ws°.ll_history,
    workspaces.jl:337) unless write_ll_prop=True."""

    def __init__(self, save_at_the_end=True, save_at_iters=(), overwrite_at_save=False,
                 filename="mcmc_results", add_datestamp=False, path=".", write_ll_prop=False,
                 chains=None):
        stamp = ""
        if add_datestamp:
            import datetime
            stamp = "_" + datetime.datetime.now().strftime("%Y-%m-%d_%H:%M:%S")
        filename = filename + stamp
        self.save_at_the_end = bool(save_at_the_end)
        self.save_at_iters = sorted(int(i) for i in save_at_iters)
        self.save_intermediate = len(self.save_at_iters) > 0
        self.filename = (_find_available_name(path, filename) if not overwrite_at_save
                         else os.path.join(path, filename + ".csv"))
        self.write_ll_prop = write_ll_prop
        self.chains = chains

    def _files(self, ws):
        chains = range(ws.C) if self.chains is None else self.chains
        if ws.C == 1:
            return [(0, self.filename)]
        stem, ext = os.path.splitext(self.filename)
        return [(c, f"{stem}_chain{c}{ext}") for c in chains]

    def init_(self, ws):                                    # callbacks.jl:188-190
        for _, f in self._files(ws):
            open(f, "w").close()

    def check_if_execute(self, step, flag):                 # callbacks.jl:198-202
        if not isinstance(flag, PreMCMCStep) or not self.save_intermediate:
            return False
        return step.mcmciter in self.save_at_iters and step.pidx == 1

    def cleanup_(self, ws, local_wss, step):                # callbacks.jl:209-211
        if self.save_at_the_end:
            self.execute_(ws, local_wss, step, None)

    def _find_starting_idx(self, step):                     # callbacks.jl:234-239
        if not self.save_intermediate:
            return 1
        import bisect
        idx = bisect.bisect_left(self.save_at_iters, step.mcmciter) + 1
        if idx == 1:
            return 1
        return self.save_at_iters[idx - 2]

    def execute_(self, ws, local_wss, step, flag):          # callbacks.jl:220-227
        start = self._find_starting_idx(step)
        sh, sph = ws.sub_ws.state_history, ws.sub_ws.state_proposal_history
        if sh is None:
            raise RuntimeError("SavingCallback needs backend history='full'")
        for c, fname in self._files(ws):
            with open(fname, "a") as f:
                for i in range(start, step.mcmciter):
                    for j in range(1, ws.NU + 1):
                        th = "".join(f"{_jl(v)}, " for v in sh[i - 1, j - 1, :, c])
                        thp = "".join(f"{_jl(v)}, " for v in sph[i - 1, j - 1, :, c])
                        lw = local_wss[j - 1]
                        l = "".join(f"{_jl(v)}, " for v in lw.sub_ws.ll_history[i - 1, :, c])
                        lp_src = lw.sub_ws_prop.ll_history[i - 1, :, c] if self.write_ll_prop else np.zeros(1)
                        lp = "".join(f"{_jl(v)}, " for v in lp_src)
                        ar = f"{'true' if lw.acceptance_history[i - 1, c] else 'false'}, "
                        f.write(f"{i}, {j}, !, {th}!, {thp}!, {l}!, {lp}!,{ar}\n")


def _jl(v):
    """Julia-style Float64 printing: shortest round-trip repr, always with a decimal point."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Inf" if v > 0 else "-Inf"
    r = repr(v)
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    return r


class REPLCallback(Callback):
    """callbacks.jl:279-325: progress printing every k iterations."""

    def __init__(self, print_every_k_iter=100, show_all_updates=True, basic_info_only=True, file=None):
        self.print_every_k_iter = int(print_every_k_iter)
        self.show_all_updates = bool(show_all_updates)
        self.basic_info_only = bool(basic_info_only)
        self.file = file

    def _p(self, *a):
        print(*a, file=self.file)

    def init_(self, ws):                                    # callbacks.jl:293-298
        self._p("*" * 40)
        self._p("Initializing an MCMC chain")
        W.summary(ws, init=True, file=self.file)
        self._p("* * *")

    def check_if_execute(self, step, flag):                 # callbacks.jl:300-304
        if not isinstance(flag, PostMCMCStep):
            return False
        if step.mcmciter % self.print_every_k_iter != 0:
            return False
        return self.show_all_updates or step.pidx == 1

    def execute_(self, ws, local_wss, step, flag):          # callbacks.jl:306-319
        lws, M = local_wss[step.pidx - 1], step.mcmciter
        self._p("- - - - - - - - - - -")
        self._p(f"{M}.{step.pidx} {W.name_of_update(lws)}")
        r4 = lambda x: float(f"{x:.4g}")
        l, lp = W.ll(lws, M)[0], W.ll_prop(lws, M)[0]
        acc = W.accepted(lws, M)
        n = l.shape[0]
        if n == 1:
            a_r = "✔" if acc[0] else "✗"
            self._p(f"\tll: {r4(l[0])}, ll°: {r4(lp[0])}, llr: {r4(lp[0] - l[0])}, a/r: {a_r}")
        else:
            self._p(f"\tll: {r4(l.mean())}, ll°: {r4(lp.mean())}, llr: {r4((lp - l).mean())} "
                    f"(means over {n} chains), accepted: {int(acc.sum())}/{n}")
        if not self.basic_info_only:
            self._p(f"\t\tθ : {np.round(W.state(lws)[:, 0], 4)}")
            self._p(f"\t    θ°: {np.round(W.state_prop(lws)[:, 0], 4)}")

    def cleanup_(self, ws, local_wss, step):                # callbacks.jl:321-325
        self._p("\n\nMCMC sampling has been successful!")
        self._p("Doing some clean-up and finishing...")
        self._p("\n⋆ ⋆ ⋆\n")
