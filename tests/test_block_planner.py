"""The host block planner that replaces __run! (src/run.jl:64-83), without a GPU: the device
calls are replaced by recorders.  A block must end exactly where a callback's check_if_execute
fires (the reference queries it before and after every step, src/callbacks.jl:33-45), callbacks
must run only after everything launched so far is host-visible, and reschedule! from a callback
must change the elements that follow."""
import types

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import run as run_mod


class _Recorder:
    def __init__(self, block_len):
        self.ws = types.SimpleNamespace(block_len=block_len, keep_history=False, executed=None)
        self.events = []

    def install(self, monkeypatch):
        monkeypatch.setattr(run_mod, "_flush", lambda ws, lws, block: block and self.events.append(
            ("flush", [(s.mcmciter, s.pidx) for s in block])))
        monkeypatch.setattr(run_mod, "_drain", lambda ws, lws: self.events.append(("drain",)))

    def blocks(self):
        return [e[1] for e in self.events if e[0] == "flush"]


def test_blocks_without_callbacks_are_block_len_long(monkeypatch):
    r = _Recorder(block_len=5)
    r.install(monkeypatch)
    run_mod.__run_(r.ws, [], [None, None], em.MCMCSchedule(6, 2), [])
    blocks = r.blocks()
    assert [len(b) for b in blocks] == [5, 5, 2]
    assert sum(blocks, []) == [(i, j) for i in range(1, 7) for j in (1, 2)]
    assert r.events[-1] == ("drain",)


def test_callbacks_cut_blocks_and_see_drained_state(monkeypatch):
    r = _Recorder(block_len=100)
    r.install(monkeypatch)
    seen = []

    class Pre(em.Callback):
        def check_if_execute(self, step, flag):
            return isinstance(flag, em.PreMCMCStep) and (step.mcmciter, step.pidx) == (3, 2)

        def execute_(self, ws, lws, step, flag):
            seen.append(("pre", step.mcmciter, step.pidx, list(r.events)))

    class Post(em.Callback):
        def check_if_execute(self, step, flag):
            return isinstance(flag, em.PostMCMCStep) and (step.mcmciter, step.pidx) == (4, 1)

        def execute_(self, ws, lws, step, flag):
            seen.append(("post", step.mcmciter, step.pidx, list(r.events)))

    run_mod.__run_(r.ws, [], [None, None], em.MCMCSchedule(5, 2), [Pre(), Post()])
    blocks = r.blocks()
    # (3, 2) starts a new block (pre-step callback), (4, 1) ends one (post-step callback)
    assert blocks == [[(1, 1), (1, 2), (2, 1), (2, 2), (3, 1)], [(3, 2), (4, 1)], [(4, 2), (5, 1), (5, 2)]]
    pre, post = seen
    assert pre[:3] == ("pre", 3, 2) and pre[3][-1] == ("drain",) and pre[3][-2][0] == "flush"
    assert post[:3] == ("post", 4, 1) and post[3][-1] == ("drain",) and post[3][-2] == ("flush", [(3, 2), (4, 1)])


def test_reschedule_from_a_callback_changes_what_follows(monkeypatch):
    r = _Recorder(block_len=4)
    r.install(monkeypatch)
    sched = em.MCMCSchedule(4, 3)

    class Drop(em.Callback):
        def check_if_execute(self, step, flag):
            return isinstance(flag, em.PostMCMCStep) and (step.mcmciter, step.pidx) == (2, 1)

        def execute_(self, ws, lws, step, flag):
            em.reschedule_(sched, 0, [2])                      # update 2 is off from now on

    run_mod.__run_(r.ws, [], [None] * 3, sched, [Drop()])
    flat = sum(r.blocks(), [])
    # the iterator computed (2, 2) before handing out (2, 1) (schedule.jl:56-66), so the change
    # first shows one element later -- exactly as in the reference
    assert flat == [(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1), (3, 3), (4, 1), (4, 3)]
