"""CPU-side checks: the C-ABI library loads and exports every symbol include/neci_gpu.h declares, the engine
refuses to run without a GPU (no CPU fallback), and the host-side mirror (tables, hashing, shift update,
blocking analysis) behaves as the reference's setup code prescribes."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import helpers
from neci_stable_b200 import capi, host, driver, _build
from neci_stable_b200.capi import ST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_lib():
    if not os.path.exists(capi.GPU_LIB):
        _build.build_gpu()
    return C.CDLL(capi.GPU_LIB)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "neci_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(neci_gpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20, names
    lib = _gpu_lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_stat_enum_matches_python_table():
    hdr = open(os.path.join(ROOT, "include", "neci_gpu.h")).read()
    body = hdr[hdr.index("enum neci_stat_index"):hdr.index("NECI_ST_COUNT\n")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"NECI_ST_([A-Z0-9_]+)", body)
    assert names == capi.ST_NAMES


def test_config_struct_layout_matches_header():
    """ctypes mirror of neci_gpu_config: same field order as the header."""
    hdr = open(os.path.join(ROOT, "include", "neci_gpu.h")).read()
    body = hdr[hdr.index("typedef struct neci_gpu_config {"):hdr.index("} neci_gpu_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = body.replace("typedef struct neci_gpu_config {", "")
    fields = []
    for decl in body.split(";"):
        decl = re.sub(r"\b(const|int32_t|int64_t|uint64_t|double)\b", " ", decl).strip()
        if decl:
            fields += [x.strip().lstrip("*").strip() for x in decl.split(",")]
    assert fields == [n for n, _ in capi.Config._fields_]


def test_engine_refuses_to_run_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    s = host.hubbard_rs_system(2, 2, nel=4)
    params = host.make_params(s, 0.0, max_walkers=1000, max_spawned=1000)
    with pytest.raises(capi.EngineError) as ei:
        capi.Engine(params)
    assert "no CUDA device" in str(ei.value) or "failed" in str(ei.value)


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "neci_stable_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "capi.py", (dirpath, f)
    txt = open(os.path.join(ROOT, "neci_stable_b200", "capi.py")).read()
    assert "liboracle" not in txt and "_build/liboracle" not in txt


# ---- host-side tables ------------------------------------------------------------------------------
def test_alias_tables_are_consistent():
    rng = np.random.default_rng(2)
    L = host.lib()
    for n in (1, 2, 7, 64, 406):
        w = rng.random(n) ** 3
        w[rng.random(n) < 0.2] = 0.0
        if w.sum() == 0:
            w[0] = 1.0
        probs = np.zeros(n); bias = np.zeros(n); alias = np.zeros(n, dtype=np.int32)
        L.neci_host_alias_build(C.c_int32(n), w.ctypes.data_as(C.c_void_p), probs.ctypes.data_as(C.c_void_p),
                                bias.ctypes.data_as(C.c_void_p), alias.ctypes.data_as(C.c_void_p))
        assert np.allclose(probs, w / w.sum(), rtol=1e-14, atol=0)
        assert np.all(bias >= -1e-12) and np.all(bias <= 1 + 1e-12)
        assert np.all(alias >= 1) and np.all(alias <= n)
        # exact reconstruction of the distribution from (bias, alias): p_i = (bias_i + sum_{j: alias_j = i} (1 - bias_j)) / n
        p = bias.copy()
        np.add.at(p, alias - 1, 1.0 - bias)
        assert np.allclose(p / n, probs, atol=1e-12)


def test_pchb_tables_shape_and_normalisation():
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    t = s.tables["pchb"]
    ij, ab = t["ij_max"], t["ab_max"]
    assert ij == ab == 6 * 7 // 2
    probs = t["probs"].reshape(ij, 3, ab)
    sums = probs.sum(axis=2)
    assert np.all((np.abs(sums - 1.0) < 1e-12) | (sums == 0.0))
    assert np.all(t["p_exch"] >= 0) and np.all(t["p_exch"] <= 1)
    # same-spin sampler of a diagonal pair (i == i) cannot exist
    for i in range(1, 7):
        ii = i + i * (i - 1) // 2
        assert sums[ii - 1, 0] == 0.0
    tg = t["tgt_orbs"].reshape(ab, 2)
    assert np.all(tg[:, 0] <= tg[:, 1]) and tg.min() == 1 and tg.max() == 6


def test_random_hash_tables_distinct_and_in_range():
    for nb in (8, 56, 72, 128):
        roi, rh2 = host.random_hash_tables(nb, seed=7)
        for t in (roi, rh2):
            assert len(set(t.tolist())) == nb and t.min() >= 1 and t.max() <= nb * 1000


def test_update_shift_formula():
    # S <- S - SftDamp * ln(N_new / N_old) / (tau * StepsSft)   (src/fcimc_iter_utilities.F90:1156-1158)
    s = host.update_shift(0.3, 0.1, 0.01, 10, 1100.0, 1000.0)
    assert abs(s - (0.3 - 0.1 * math.log(1.1) / (0.01 * 10))) < 1e-15
    assert host.update_shift(0.3, 0.1, 0.01, 10, 1000.0, 1000.0) == 0.3


def test_blocking_analysis_on_correlated_series():
    rng = np.random.default_rng(4)
    n = 1 << 14
    x = np.zeros(n)
    for i in range(1, n):
        x[i] = 0.9 * x[i - 1] + rng.normal()
    mean, err = driver.blocking(x + 5.0)
    naive = x.std(ddof=1) / math.sqrt(n)
    true = naive * math.sqrt((1 + 0.9) / (1 - 0.9))
    assert abs(mean - 5.0) < 5 * true
    assert 0.6 * true < err < 1.6 * true and err > 2 * naive


# ---- partitioning: DetermineDetNode / FindWalkerHash on the oracle ---------------------------------------
def test_det_node_and_walker_hash_properties():
    s = host.random_fcidump_system(28, 14, seed=25)
    o, params = helpers.make_pair(s, 0.0, max_walkers=1000, max_spawned=1000, nranks=8, blocks_per_rank=100)
    rng = np.random.default_rng(0)
    dets = helpers.random_dets(s, 20000, rng)
    il = np.array([s.ilut(d) for d in dets]).reshape(len(dets), s.nw)
    blk, node = o.probe_det_node(il)
    assert blk.min() >= 1 and blk.max() <= 800
    assert np.array_equal(node, (blk - 1) // 100)                   # init_load_balance: LoadBalanceMapping(i) = int((i-1)/oversample_factor)
    counts = np.bincount(node, minlength=8)
    assert counts.min() > 0.8 * len(dets) / 8 and counts.max() < 1.2 * len(dets) / 8
    # hand evaluation of get_det_block for one determinant (load_balance_calcnodes.F90:92-115)
    roi = params["random_orb_index"]
    acc = 0
    for i, orb in enumerate(dets[0], start=1):
        acc = (1099511628211 * acc + int(roi[orb - 1]) * i) & ((1 << 64) - 1)
    sacc = acc - (1 << 64) if acc >= (1 << 63) else acc
    m = abs(sacc) % 800            # |Fortran mod(a, n)| == |a| mod n (truncated division)
    assert blk[0] == m + 1
    wh = o.probe_walker_hash(il, 7001)
    assert wh.min() >= 1 and wh.max() <= 7001 and len(np.unique(wh)) > 5000


# ---- partition independence of the whole algorithm (oracle world of 1 vs 4 ranks) -------------------------------
@pytest.mark.parametrize("kind", ["pchb", "hub_k"])
def test_oracle_world_is_partition_independent(kind):
    """Random streams are keyed by determinant, so the union of the rank-local walker lists after n iterations
    must not depend on the number of ranks (integer walkers: bit-exact)."""
    if kind == "pchb":
        s, tau = host.random_fcidump_system(8, 6, sparse=0.9, sparse_t=0.9, seed=3), 2e-3
    else:
        s, tau = host.hubbard_k_system(4, 4, U=4.0), 0.006
    hii = driver.diag_energy(s, s.ref_orbs)
    lists = {}
    for nr in (1, 4):
        orcs = []
        for r in range(nr):
            o, params = helpers.make_pair(s, hii, max_walkers=200000, max_spawned=200000, nranks=nr, rank=r, seed=11,
                                          blocks_per_rank=10)
            orcs.append(o)
        rec = host.record(s, s.ref_orbs, 100.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1)
        _, node = orcs[0].probe_det_node(rec[:, :s.nw])
        for r in range(nr):
            orcs[r].upload_walkers(rec if node[0] == r else np.zeros((0, s.W), dtype=np.int64))
        tot = []
        for it in range(1, 41):
            st = helpers.world_iterate(orcs, tau, 0.0, it, nthreads=nr)
            tot.append((st[:, ST["TOTPARTS"]].sum(), st[:, ST["NOBORN"]].sum(), st[:, ST["ANNIHILATED"]].sum(),
                        st[:, ST["NINSERTED"]].sum()))
            assert st[:, ST["ERR_FLAGS"]].max() == 0
        parts = [o.download_walkers() for o in orcs]
        d = np.concatenate([p[0] for p in parts]); gd = np.concatenate([p[1] for p in parts]); go = np.concatenate([p[2] for p in parts])
        lists[nr] = (helpers.canon(d, gd, go, nw=s.nw), tot)
        if nr > 1:   # every determinant sits on its owner
            for r, p in enumerate(parts):
                c = helpers.canon(p[0], nw=s.nw)
                if c[0].shape[0]:
                    _, nd = orcs[0].probe_det_node(c[0])
                    assert np.all(nd == r)
    a, b = lists[1], lists[4]
    assert a[1] == b[1]
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y)
    assert a[0][0].shape[0] > 50
