"""The N > 1 path on CPU: two processes (gloo, 127.0.0.1), each one oracle rank owning the determinants
DetermineDetNode assigns to it, spawns exchanged through the process group, statistics reduced by
driver.reduce_stats exactly as the GPU job does with NCCL.  Because the random streams are keyed by determinant,
the 2-rank job must reproduce the 1-rank job: same walker set, same shift trajectory."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from neci_stable_b200 import capi, host, driver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ITER = 300


def _system(kind="hub_k"):
    if kind == "pchb_hphf":                    # HPHF functions: spawns are routed by their representative determinant
        return host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3), 0.004, dict(hphf=True)
    return host.hubbard_k_system(3, 2, nel=4, U=4.0), 0.03, {}


def _run(engine, system, hii, tau, nranks, rank, owner_of_ref):
    run = driver.FciMC(system, engine, hii, tau=tau, init_walkers=150, steps_sft=5, sft_damp=0.2, nranks=nranks)
    run.init_walkers = 150 / nranks            # the shift trigger is InitWalkers * nNodes (fcimc_iter_utilities.F90:1029)
    run.seed_reference(20, rank_of_ref=owner_of_ref, my_rank=rank)
    run.run(N_ITER)
    return run


def _worker(rank, world, port, out_dir, kind):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    system, tau, extra = _system(kind)
    hii = driver.diag_energy(system, system.ref_orbs)
    o, params = helpers.make_pair(system, hii, max_walkers=50000, max_spawned=50000, nranks=world, rank=rank, seed=5,
                                  blocks_per_rank=4, **extra)
    _, node = o.probe_det_node(system.ilut(system.ref_orbs).reshape(1, -1))
    eng = helpers.DistOracle(o, dist)
    run = _run(eng, system, hii, tau, world, rank, int(node[0]))
    d, gd, go = o.download_walkers()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), dets=d, gd=gd, go=go,
             shift=np.array([h["shift"] for h in run.history]), parts=np.array([h["tot_parts"] for h in run.history]),
             enum=np.array([h["enum_cyc"] for h in run.history]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["hub_k", "pchb_hphf"])
def test_two_rank_job_reproduces_single_rank(tmp_path, kind):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), kind), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, "rank%d.npz" % k)) for k in range(2)]
    # both ranks saw the same reduced statistics
    assert np.array_equal(r[0]["shift"], r[1]["shift"]) and np.array_equal(r[0]["parts"], r[1]["parts"])
    # single-rank run of the same job
    system, tau, extra = _system(kind)
    hii = driver.diag_energy(system, system.ref_orbs)
    o, params = helpers.make_pair(system, hii, max_walkers=50000, max_spawned=50000, nranks=1, rank=0, seed=5, **extra)
    run = _run(o, system, hii, tau, 1, 0, 0)
    assert np.array_equal(r[0]["parts"], np.array([h["tot_parts"] for h in run.history]))
    assert np.allclose(r[0]["shift"], np.array([h["shift"] for h in run.history]), rtol=1e-12, atol=1e-12)
    assert np.allclose(r[0]["enum"], np.array([h["enum_cyc"] for h in run.history]), rtol=1e-10, atol=1e-10)
    assert r[0]["parts"][-1] > 100 and np.any(r[0]["shift"] != 0.0)
    both = helpers.canon(np.concatenate([r[0]["dets"], r[1]["dets"]]), np.concatenate([r[0]["gd"], r[1]["gd"]]),
                         np.concatenate([r[0]["go"], r[1]["go"]]), nw=system.nw)
    one = helpers.canon(*o.download_walkers(), nw=system.nw)
    for a, b in zip(both[:3], one[:3]):
        assert np.array_equal(a, b)
    # ownership
    o2, _ = helpers.make_pair(system, hii, max_walkers=100, max_spawned=100, nranks=2, rank=0, seed=5, blocks_per_rank=4, **extra)
    for k in range(2):
        c = helpers.canon(r[k]["dets"], nw=system.nw)
        _, nd = o2.probe_det_node(c[0])
        assert np.all(nd == k)


def _tau_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    system = host.random_fcidump_system(8, 6, sparse=0.8, sparse_t=0.8, seed=5, p_singles=0.2)
    hii = driver.diag_energy(system, system.ref_orbs)
    o, params = helpers.make_pair(system, hii, max_walkers=100000, max_spawned=100000, nranks=world, rank=rank, seed=9,
                                  blocks_per_rank=4, tau_search=True, initiator=False)
    _, node = o.probe_det_node(system.ilut(system.ref_orbs).reshape(1, -1))
    eng = helpers.DistOracle(o, dist)
    t = system.tables["pchb"]
    ts = driver.TauSearch(0.02, t["p_singles"], t["p_doubles"], t["p_parallel"], consider_par_bias=True,
                          reduce_or=driver.dist_reduce_or, reduce_max=driver.dist_reduce_max)
    rec = host.record(system, system.ref_orbs, 400.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1)
    o.upload_walkers(rec if rank == int(node[0]) else np.zeros((0, system.W), dtype=np.int64))
    tau = ts.tau
    local_gamma = np.zeros(4); trace = []
    for it in range(1, 61):
        st = eng.iterate(tau, 0.0, it)                     # this rank's statistics vector
        ts.log(st)
        local_gamma = np.maximum(local_gamma, st[capi.ST["TAU_GAMMA_SING"]:capi.ST["TAU_GAMMA_SING"] + 4])
        if it % 10 == 0:
            tau, ps, pd, pp = ts.update()                  # collective: every rank calls it
            o.set_excit_probs(ps, pd, pp)
            trace.append((tau, ps, pd, pp))
    np.savez(os.path.join(out_dir, "tau%d.npz" % rank), trace=np.array(trace), local_gamma=local_gamma, gamma=ts.gamma,
             cnt=ts.cnt)
    dist.barrier()
    dist.destroy_process_group()


def test_tau_search_switches_and_maxima_are_reduced_over_the_ranks(tmp_path):
    """update_tau on two ranks (src/tau/tau_search_conventional.F90:295-312): gamma_* and max_death_cpt are reduced with
    MPI_MAX and the enough_* switches with a logical OR, so every rank assigns the same tau and biases although their
    own counters differ."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_tau_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, "tau%d.npz" % k)) for k in range(2)]
    assert np.array_equal(r[0]["trace"], r[1]["trace"])                    # same tau, pSingles, pDoubles, pParallel on both ranks
    assert not np.array_equal(r[0]["cnt"], r[1]["cnt"])                     # from different local counters
    assert np.array_equal(r[0]["gamma"], r[1]["gamma"])
    assert np.array_equal(r[0]["gamma"], np.maximum(r[0]["local_gamma"], r[1]["local_gamma"]))
    assert r[0]["trace"][-1, 0] < 0.02 and 0.0 < r[0]["trace"][-1, 1] < 1.0
