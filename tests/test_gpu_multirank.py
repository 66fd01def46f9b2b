"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): one process per GPU, determinants partitioned by
DetermineDetNode, spawns exchanged with NCCL grouped send/recv inside neci_gpu_iterate.  The union of the rank-local
lists must be bit-identical to the CPU oracle world of the same number of ranks, iteration by iteration."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ITER = 30


def _worker(rank, world, port, out_dir, name, semi, balance=False, p2p=False, devbuild=False):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import helpers
    import test_gpu_parity as T
    from neci_stable_b200 import capi, host, driver
    from neci_stable_b200.capi import ST
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    system, tau = T.get_system(name)
    hii = driver.diag_energy(system, system.ref_orbs)
    mapping = None
    if balance:
        # a deliberately lopsided LoadBalanceMapping (two thirds of the blocks on rank 0) so that the greedy
        # balancer has blocks to move whatever the world size
        nb = 10 * world
        mapping = np.array([0 if b < (2 * nb) // 3 else 1 + (b % (world - 1)) for b in range(nb)], dtype=np.int32)
    params = host.make_params(system, hii, max_walkers=400000, max_spawned=400000, nranks=world, rank=rank, device=rank,
                              seed=11, blocks_per_rank=10, semi_stochastic=semi, mapping=mapping)
    gpu = capi.Engine(params)
    system.apply(gpu)
    uid = [gpu.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    gpu.nccl_init(uid[0])
    if p2p:
        # spawn exchange over NVLink peer memory: all-gather of the CUDA IPC handles of the inboxes
        hs = [None] * world
        dist.all_gather_object(hs, gpu.p2p_handle())
        gpu.p2p_open(hs)
    # the oracle world lives on rank 0
    orcs = []
    if rank == 0:
        for r in range(world):
            p = dict(params); p["rank"] = r
            o = helpers.Oracle(p); system.apply(o); orcs.append(o)
    rng = np.random.default_rng(3)
    dets = helpers.random_dets(system, 300, rng)
    ref = [int(x) for x in system.ref_orbs]
    if ref not in dets:
        dets = [ref] + dets[:-1]
    il = np.array([system.ilut(d) for d in dets]).reshape(len(dets), system.nw)
    _, node = gpu.probe_det_node(il)
    signs = rng.integers(1, 9, len(dets)) * rng.choice([-1, 1], len(dets))
    n_core = 100 if semi else 0
    flags = [((1 << capi.FLAG_DETERMINISTIC) if i < n_core else 0) | (1 << capi.FLAG_INITIATOR) for i in range(len(dets))]
    recs = np.array([host.record(system, d, float(s), f) for d, s, f in zip(dets, signs, flags)])
    if semi:
        tmp = helpers.Oracle(params); system.apply(tmp)
        core, sizes, displs, per_rank, H = helpers.build_core_space(tmp, system, dets[:n_core], hii, nranks=world)
        # core determinants first, in core-space order, on their owner
        core_il = np.array([system.ilut(d) for d in core]).reshape(len(core), system.nw)
        lookup = {tuple(d): r_ for d, r_ in zip(dets, recs)}
        mine = [lookup[tuple(d)] for d in core[displs[rank]:displs[rank] + sizes[rank]]]
        rest = [recs[i] for i in range(n_core, len(dets)) if node[i] == rank]
        my = np.array(mine + rest).reshape(-1, system.W)
        gpu.upload_walkers(my)
        c = per_rank[rank]
        if devbuild:       # this rank's rows built on its own device (neci_gpu_build_core_space)
            nnz = gpu.build_core_space(sizes, displs, c["iluts"])
            assert nnz == c["row_ptr"][-1], (nnz, c["row_ptr"][-1])
        else:
            gpu.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
        if rank == 0:
            for r in range(world):
                mine_r = [lookup[tuple(d)] for d in core[displs[r]:displs[r] + sizes[r]]]
                rest_r = [recs[i] for i in range(n_core, len(dets)) if node[i] == r]
                orcs[r].upload_walkers(np.array(mine_r + rest_r).reshape(-1, system.W))
                cr = per_rank[r]
                orcs[r].set_core_space(cr["row_ptr"], cr["col"], cr["val"], sizes, displs, cr["iluts"])
    else:
        gpu.upload_walkers(recs[node == rank])
        if rank == 0:
            for r in range(world):
                orcs[r].upload_walkers(recs[node == r])
    ok = True
    msg = ""
    n_moves = 0
    for it in range(1, N_ITER + 1):
        sg = gpu.iterate(tau, -0.1, it)
        allg = [None] * world
        dist.all_gather_object(allg, sg)
        if rank == 0:
            so = helpers.world_iterate(orcs, tau, -0.1, it, nthreads=1)
            for r in range(world):
                for n in ("NVALIDEXCITS", "NINVALIDEXCITS", "NSPAWNED_SENT", "NSPAWNED_RECV", "NSPAWNED_MERGED", "NINSERTED",
                          "NODIED", "NOBORN", "NOABORTED"):
                    if allg[r][ST[n]] != so[r][ST[n]]:
                        ok = False; msg = "stat %s rank %d it %d: gpu %r oracle %r" % (n, r, it, allg[r][ST[n]], so[r][ST[n]])
                for n in ("TOTPARTS", "ANNIHILATED", "ENUMCYC"):
                    if not np.isclose(allg[r][ST[n]], so[r][ST[n]], rtol=1e-11, atol=1e-9):
                        ok = False; msg = "stat %s rank %d it %d: gpu %r oracle %r" % (n, r, it, allg[r][ST[n]], so[r][ST[n]])
        if not ok:
            break
        if balance and it % 10 == 0:
            # adjust_load_balance: block populations summed over ranks, greedy plan (identical on every rank), block moves
            allp = [None] * world
            dist.all_gather_object(allp, gpu.block_populations())
            parts = np.sum(allp, axis=0)
            new, moves = driver.plan_load_balance(parts, gpu.params["load_balance_mapping"], world)
            if rank == 0:
                op = np.sum([o.block_populations() for o in orcs], axis=0)
                if not np.array_equal(op, parts):
                    ok = False; msg = "block populations differ at it %d" % it
                helpers.world_rebalance(orcs, new)
            gpu.rebalance(new)
            n_moves += len(moves)
    if balance and n_moves == 0:
        ok = False; msg = "no block was moved"
    lists = [None] * world
    dist.all_gather_object(lists, gpu.download_walkers())
    if rank == 0 and ok:
        for r in range(world):
            a = helpers.canon(*lists[r], nw=system.nw)
            b = helpers.canon(*orcs[r].download_walkers(), nw=system.nw)
            if not (a[0].shape == b[0].shape and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
                    and np.allclose(a[1], b[1], rtol=1e-12, atol=1e-12)):
                ok = False; msg = "walker lists differ on rank %d (%d vs %d dets)" % (r, a[0].shape[0], b[0].shape[0])
            if a[0].shape[0]:
                _, nd = gpu.probe_det_node(a[0])
                if not np.all(nd == r):
                    ok = False; msg = "determinant on the wrong rank %d" % r
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write("OK %d" % sum(helpers.canon(*l, nw=system.nw)[0].shape[0] for l in lists) if ok else "FAIL " + msg)
    elif rank == 0:
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write("FAIL " + msg)
    gpu.close()
    dist.barrier()
    dist.destroy_process_group()




@pytest.mark.parametrize("name,semi,balance,p2p,devbuild", [
    ("pchb_14e28o", False, False, False, False), ("hub_k_6x6_2words", False, False, True, False),
    ("hub_rs_4x4", False, True, False, False), ("pchb_14e28o", False, True, True, False),
    ("pchb_6e6o", True, False, True, False), ("pchb_14e28o", False, False, True, False),
    pytest.param("pchb_6e6o", True, False, True, True)])
def test_multi_gpu_matches_oracle_world(tmp_path, name, semi, balance, p2p, devbuild):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), name, semi, balance, p2p, devbuild), nprocs=world, join=True)
    res = open(os.path.join(tmp_path, "result.txt")).read()
    assert res.startswith("OK"), res
    assert int(res.split()[1]) > 100
