"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

1. chains shard: the concatenation of the ranks' trajectories equals the single-GPU run
   bit-for-bit (Philox key space is global).
2. observations shard + NCCL all-reduce: every rank holds identical chains; log-likelihoods
   agree with the single-GPU sweep within 1e-10 relative, decisions agree (a different
   summation association can only flip a near-tie; flips are counted and reported).
3. logistic law, rows sharded, MALA: ll and gradients all-reduced.
4. hierarchical law, observations sharded, MALA + walks: per-group sums all-reduced.
5. general-d Gaussian law, rows sharded.
Prints one JSON line on rank 0 and exits non-zero on failure.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import extensiblemcmc_jl_b200 as em  # noqa: E402
from extensiblemcmc_jl_b200 import parallel as par  # noqa: E402
from tests.parity import cfg2_updates, theta_init_for  # noqa: E402


def run(backend, x, th0, M):
    mcmc = em.MCMC(cfg2_updates(eps0=0.05, scale=5e-3, k=10, offset=2.0), backend=backend)
    ws, lws = em.run_(mcmc, M, dict(P=em.GsnTargetLaw([0.0]), obs=x), th0)
    out = dict(theta=ws.sub_ws.state_history.copy(), ll=ws.ll_all.copy(), acc=ws.acc_all.copy(),
               eps=ws.eps(1).copy())
    ws.close()
    return out


def main():
    rank, world, local = par.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok, report = True, {}
    Cn, M = 192, 40
    x = 1.5 + 2.0 * np.random.default_rng(2).standard_normal(20001)
    th0 = theta_init_for(x, Cn)
    single = run(em.CUDAMCMCBackend(n_chains=Cn, device=local, seed=9, block_len=16), x, th0, M)

    # 1. chains shard
    b = par.backend_for_rank(rank, world, local, Cn, shard="chains", seed=9, block_len=16)
    mine = run(b, x, th0[:, b.chain_offset:b.chain_offset + b.n_chains], M)
    full = par.gather_chain_axis(dist, mine["theta"])
    eps = par.gather_chain_axis(dist, mine["eps"])
    if rank == 0:
        report["chains_bitexact"] = bool(np.array_equal(full, single["theta"]) and np.array_equal(eps, single["eps"]))
        ok &= report["chains_bitexact"]

    # 2. observations shard
    cid = par.exchange_comm_id(dist)
    first, cnt = par.shard_obs(len(x), rank, world)
    bo = par.backend_for_rank(rank, world, local, Cn, shard="obs", comm_id=cid, seed=9, block_len=16)
    sh = run(bo, x[first:first + cnt], th0, M)
    fin = np.isfinite(single["ll"])
    same_dec = np.array_equal(sh["acc"], single["acc"])
    rel = np.abs(sh["ll"][fin] - single["ll"][fin]) / np.abs(single["ll"][fin])
    flips = int((sh["acc"] != single["acc"]).any(axis=(0, 1)).sum())
    # 2b. the same with the library's own NVLink peer exchange instead of ncclAllReduce
    bp = par.backend_for_rank(rank, world, local, Cn, shard="obs", comm_id=par.exchange_comm_id(dist), seed=9,
                              block_len=16, p2p_allgather=par.p2p_allgather_fn(dist))
    shp = run(bp, x[first:first + cnt], th0, M)
    p2p_same = bool(np.array_equal(shp["acc"], sh["acc"]) and np.array_equal(shp["theta"], sh["theta"]))
    fin_p = np.isfinite(sh["ll"])
    p2p_rel = float((np.abs(shp["ll"][fin_p] - sh["ll"][fin_p]) / np.abs(sh["ll"][fin_p])).max())
    # 2c. few chains (observation-mapped sweep whose fused tail reduces and pushes the sums itself)
    C6 = 6
    th6 = th0[:, :C6]
    s6 = run(em.CUDAMCMCBackend(n_chains=C6, device=local, seed=9, block_len=16), x, th6, M)
    n6 = run(par.backend_for_rank(rank, world, local, C6, shard="obs", comm_id=par.exchange_comm_id(dist), seed=9,
                                  block_len=16), x[first:first + cnt], th6, M)
    p6 = run(par.backend_for_rank(rank, world, local, C6, shard="obs", comm_id=par.exchange_comm_id(dist), seed=9,
                                  block_len=16, p2p_allgather=par.p2p_allgather_fn(dist)), x[first:first + cnt], th6, M)
    fin6 = np.isfinite(s6["ll"])
    few_ok = bool(np.array_equal(s6["acc"], n6["acc"]) and np.array_equal(s6["acc"], p6["acc"])
                  and np.array_equal(s6["theta"], n6["theta"]) and np.array_equal(s6["theta"], p6["theta"])
                  and np.allclose(s6["ll"][fin6], n6["ll"][fin6], rtol=1e-12, atol=0)
                  and np.allclose(s6["ll"][fin6], p6["ll"][fin6], rtol=1e-12, atol=0))

    # 3. logistic regression (FP64 tensor-core kernel), rows of X sharded over the ranks, MALA:
    #    per-chain ll and the [d][C] gradients are all-reduced; vs the single-GPU run the sums
    #    differ in the last bits, so proposals agree to ~1e-12 and decisions agree
    rng = np.random.default_rng(3)
    dl, nl, Cl, Ml = 8, 4000, 96, 25
    Xl = rng.standard_normal((nl, dl)) / np.sqrt(dl)
    yl = (rng.random(nl) < 1.0 / (1.0 + np.exp(-Xl @ rng.standard_normal(dl)))).astype(np.float64)
    mala = lambda: [em.MALAUpdate(0.2, list(range(1, dl + 1)), prior=em.StandardPrior(em.Normal(0.0, 10.0)),
                                  adpt=em.AdaptationMALA(adapt_every_k_steps=5, scale=0.01, offset=1.0))]

    def run_logi(backend, Xs, ys):
        mcmc = em.MCMC(mala(), backend=backend)
        ws, lws = em.run_(mcmc, Ml, dict(P=em.LogisticLaw(dl), obs=Xs, y=ys), np.zeros(dl))
        out = (ws.sub_ws.state_history.copy(), ws.acc_all.copy(), ws.ll_all.copy())
        ws.close()
        return out
    l1 = run_logi(em.CUDAMCMCBackend(n_chains=Cl, device=local, seed=4, block_len=10), Xl, yl)
    f3, c3 = par.shard_obs(nl, rank, world)
    lb = par.backend_for_rank(rank, world, local, Cl, shard="obs", comm_id=par.exchange_comm_id(dist), seed=4,
                              block_len=10)
    l2 = run_logi(lb, Xl[f3:f3 + c3], yl[f3:f3 + c3])
    logi_ok = bool(np.array_equal(l1[1], l2[1]) and np.allclose(l1[0], l2[0], rtol=1e-9, atol=1e-11)
                   and np.allclose(l1[2][np.isfinite(l1[2])], l2[2][np.isfinite(l1[2])], rtol=1e-10, atol=0))
    # 4. gradient sweeps of the Gaussian laws under observation sharding: the per-group sums of both
    #    orders are all-reduced before the MALA kernels finish ll and gradient (cfg 4 schedule: MALA on
    #    theta_1..G, walks on mu and tau).  Contiguous slices of group-sorted data: ranks see different
    #    (possibly no) observations of a group.
    G, ng, Ch, Mh = 4, 300, 64, 30
    rng = np.random.default_rng(5)
    tg = rng.standard_normal(G)
    yh = np.concatenate([tg[g] + rng.standard_normal(ng) for g in range(G)])
    gh = np.repeat(np.arange(G), ng).astype(np.float64)
    hier_ups = lambda: [em.MALAUpdate(0.1, list(range(1, G + 1)),
                                      adpt=em.AdaptationMALA(adapt_every_k_steps=6, scale=0.004, offset=1.0)),
                        em.RandomWalkUpdate(em.UniformRandomWalk([0.4]), [G + 1]),
                        em.RandomWalkUpdate(em.UniformRandomWalk([0.4], [True]), [G + 2], prior=em.ImproperPosPrior())]
    thh = np.concatenate([np.zeros(G), [0.0, 1.0]])

    def run_hier(backend, ys, gs):
        mcmc = em.MCMC(hier_ups(), backend=backend)
        ws, lws = em.run_(mcmc, Mh, dict(P=em.HierNormalLaw(G), obs=ys, groups=gs), thh)
        out = (ws.sub_ws.state_history.copy(), ws.acc_all.copy(), ws.ll_all.copy())
        ws.close()
        return out
    h1 = run_hier(em.CUDAMCMCBackend(n_chains=Ch, device=local, seed=8, block_len=9), yh, gh)
    f4, c4 = par.shard_obs(len(yh), rank, world)
    h2 = run_hier(par.backend_for_rank(rank, world, local, Ch, shard="obs", comm_id=par.exchange_comm_id(dist),
                                       seed=8, block_len=9), yh[f4:f4 + c4], gh[f4:f4 + c4])
    finh = np.isfinite(h1[2])
    hier_ok = bool(np.array_equal(h1[1], h2[1]) and np.allclose(h1[0], h2[0], rtol=1e-9, atol=1e-11)
                   and np.allclose(h1[2][finh], h2[2][finh], rtol=1e-10, atol=0))
    # 5. general-d Gaussian law (theta = [mu; vec Sigma]) with the rows of the data sharded
    dm, nm, Cm, Mm = 3, 1500, 80, 20
    rng = np.random.default_rng(7)
    Am = rng.standard_normal((dm, dm))
    Sm = Am @ Am.T / dm + np.eye(dm)
    Xm = rng.multivariate_normal(np.zeros(dm), Sm, size=nm)
    mv_ups = lambda: [em.RandomWalkUpdate(em.UniformRandomWalk([0.1]), [k + 1]) for k in range(dm)] + \
                     [em.RandomWalkUpdate(em.UniformRandomWalk([0.1], [True]), [dm + 1], prior=em.ImproperPosPrior())]
    mv_law = em.GsnTargetLaw(np.zeros(dm), Sm)

    def run_mv(backend, Xs):
        mcmc = em.MCMC(mv_ups(), backend=backend)
        ws, lws = em.run_(mcmc, Mm, dict(P=mv_law, obs=Xs), mv_law.theta)
        out = (ws.sub_ws.state_history.copy(), ws.acc_all.copy(), ws.ll_all.copy())
        ws.close()
        return out
    m1 = run_mv(em.CUDAMCMCBackend(n_chains=Cm, device=local, seed=11, block_len=8), Xm)
    f5, c5 = par.shard_obs(nm, rank, world)
    m2 = run_mv(par.backend_for_rank(rank, world, local, Cm, shard="obs", comm_id=par.exchange_comm_id(dist),
                                     seed=11, block_len=8), Xm[f5:f5 + c5])
    finm = np.isfinite(m1[2])
    mv_ok = bool(np.array_equal(m1[1], m2[1]) and np.array_equal(m1[0], m2[0])
                 and np.allclose(m1[2][finm], m2[2][finm], rtol=1e-10, atol=0))
    gathered = par.gather_chain_axis(dist, sh["theta"][..., :4])
    if rank == 0:
        report.update(obs_ll_rel_err=float(rel.max()) if same_dec else None, obs_decisions_equal=bool(same_dec),
                      obs_chains_with_flips=flips,
                      obs_ranks_identical=bool(all(np.array_equal(gathered[..., :4], gathered[..., 4 * r:4 * r + 4])
                                                   for r in range(world))))
        report.update(p2p_matches_nccl=p2p_same, p2p_ll_rel_vs_nccl=p2p_rel, logistic_obs_sharded_ok=logi_ok,
                      few_chains_fused_tail_ok=few_ok, hier_mala_obs_sharded_ok=hier_ok,
                      gsnmv_obs_sharded_ok=mv_ok)
        ok &= logi_ok and few_ok and hier_ok and mv_ok
        ok &= same_dec and rel.max() < 1e-10 and report["obs_ranks_identical"] and p2p_same and p2p_rel < 1e-12
        report["world"] = world
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
