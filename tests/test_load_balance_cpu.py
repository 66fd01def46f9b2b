"""adjust_load_balance (src/load_balancer.fpp:178-351): the greedy planner, the trigger, and block moves on the oracle
world -- moving blocks between ranks must not change the physics (same walker set as an unbalanced run)."""
import numpy as np

import helpers
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST


def test_planner_moves_smallest_block_from_fullest_to_emptiest():
    # 2 ranks, 6 blocks; rank 0 owns blocks 0,2,4 (100, 30, 10), rank 1 owns 1,3,5 (5, 5, 0)
    parts = np.array([100, 5, 30, 5, 10, 0])
    mapping = np.array([0, 1, 0, 1, 0, 1])
    new, moves = driver.plan_load_balance(parts, mapping, 2)
    # avg = 75: first move block 4 (10): 140->130, 10->20 ok; then block 2 (30): 130->100, 20->50 ok; then block 0 (100)
    # would take rank 0 to 0 and rank 1 to 150: not an improvement -> stop
    assert moves == [(4, 0, 1), (2, 0, 1)]
    assert list(new) == [0, 1, 1, 1, 1, 1]
    # already balanced: nothing to do
    new2, moves2 = driver.plan_load_balance(np.array([10, 10, 10, 10]), np.array([0, 1, 0, 1]), 2)
    assert moves2 == [] and list(new2) == [0, 1, 0, 1]


def test_planner_never_increases_imbalance():
    rng = np.random.default_rng(1)
    for trial in range(30):
        nr = int(rng.integers(2, 9)); nb = nr * int(rng.integers(2, 40))
        parts = (rng.exponential(100.0, nb) * (rng.random(nb) < 0.8)).astype(np.int64)
        mapping = np.arange(nb) % nr
        before = np.bincount(mapping, weights=parts, minlength=nr)
        new, moves = driver.plan_load_balance(parts, mapping, nr)
        after = np.bincount(new, weights=parts, minlength=nr)
        assert after.max() <= before.max() and after.sum() == before.sum()
        assert len(moves) <= nb * 4


def test_trigger_state_machine():
    t = driver.LoadBalanceTrigger()
    assert not t.need(0.05)                 # below the absolute threshold
    assert t.need(0.3)                      # balance
    assert not t.need(0.12)                 # the cycle after balancing only logs the measure
    assert not t.need(0.2)                  # 0.2 < 2 x 0.12
    assert t.need(0.25)
    imb = driver.LoadBalanceTrigger.imbalance(np.array([[1.0, 1.0], [2.0, 1.0]]))
    assert abs(imb - 0.5 / 5.0) < 1e-15


def test_block_moves_do_not_change_the_walker_set():
    s, tau = host.hubbard_k_system(4, 4, U=4.0), 0.006
    hii = driver.diag_energy(s, s.ref_orbs)
    nr = 4
    results = []
    for balance in (False, True):
        orcs = []
        for r in range(nr):
            o, params = helpers.make_pair(s, hii, max_walkers=100000, max_spawned=100000, nranks=nr, rank=r, seed=7,
                                          blocks_per_rank=25)
            orcs.append(o)
        mapping = np.array(params["load_balance_mapping"]).copy()
        rec = host.record(s, s.ref_orbs, 100.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1)
        _, node = orcs[0].probe_det_node(rec[:, :s.nw])
        for r in range(nr):
            orcs[r].upload_walkers(rec if node[0] == r else np.zeros((0, s.W), dtype=np.int64))
        n_moves = 0
        for it in range(1, 41):
            helpers.world_iterate(orcs, tau, 0.0, it, nthreads=1)
            if balance and it % 10 == 0:
                parts = np.zeros(nr * 25)
                for o in orcs:
                    parts += o.block_populations()
                new, moves = driver.plan_load_balance(parts, mapping, nr)
                n_moves += len(moves)
                helpers.world_rebalance(orcs, new)
                mapping = new
                # every determinant now sits on its new owner, populations as planned
                for r, o in enumerate(orcs):
                    c = helpers.canon(o.download_walkers()[0], nw=s.nw)
                    if c[0].shape[0]:
                        _, nd = o.probe_det_node(c[0])
                        assert np.all(nd == r)
                after = np.array([o.block_populations().sum() for o in orcs])
                assert np.array_equal(after, np.bincount(new, weights=parts, minlength=nr))
        parts_ = [o.download_walkers() for o in orcs]
        d = np.concatenate([p[0] for p in parts_]); gd = np.concatenate([p[1] for p in parts_]); go = np.concatenate([p[2] for p in parts_])
        results.append(helpers.canon(d, gd, go, nw=s.nw))
        if balance:
            assert n_moves > 0
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(a, b)
