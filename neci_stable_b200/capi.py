"""ctypes binding of the engine's C ABI (include/neci_gpu.h).

`Engine` is a thin, typed wrapper: one method per exported symbol, numpy arrays
in and out.  The same wrapper class can bind any library exporting the same
entry points under another prefix (the parity tests bind the CPU oracle with
prefix ``orc_``); the product only ever loads ``libneci_gpu.so`` and raises if
it is missing -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(HERE, "libneci_gpu.so")

# statistic indices -- keep in sync with enum neci_stat_index
ST_NAMES = [
    "NOBORN", "NODIED", "ANNIHILATED", "NOABORTED", "NOREMOVED", "SPAWNFROMSING", "ACCEPTANCES", "HFCYC",
    "NOATDOUBS", "ENUMCYC", "ENUMCYCABS", "INITSENUMCYC", "NOINITDETS", "NONONINITDETS", "NOINITWALK",
    "NONONINITWALK", "NOADDEDINITIATORS", "NVALIDEXCITS", "NINVALIDEXCITS", "BLOOM_COUNT_1", "BLOOM_COUNT_2",
    "MAX_CYC_SPAWN", "BLOOM_SIZE_1", "BLOOM_SIZE_2", "TAU_GAMMA_SING", "TAU_GAMMA_DOUB", "TAU_GAMMA_PAR",
    "TAU_GAMMA_OPP", "TAU_MAX_DEATH_CPT", "TOTPARTS", "NORM_PSI_SQ", "NORM_SEMISTOCH_SQ",
    "INSTNOATHF", "TOTWALKERS", "HOLESINLIST", "NSPAWNED_SENT", "NSPAWNED_RECV", "NSPAWNED_MERGED",
    "NINSERTED", "HIGHEST_POP", "TRIAL_NUMERATOR", "TRIAL_DENOM", "INIT_TRIAL_NUMERATOR", "INIT_TRIAL_DENOM",
    "TAU_CNT_SING", "TAU_CNT_DOUB", "TAU_CNT_PAR", "TAU_CNT_OPP", "ERR_FLAGS", "TIME_SPAWN_MS", "TIME_COMM_MS", "TIME_ANNIHIL_MS",
    "TIME_DETERM_MS",
]
ST = {n: i for i, n in enumerate(ST_NAMES)}
ST_COUNT = len(ST_NAMES)
ST_MAX_REDUCED = ("MAX_CYC_SPAWN", "BLOOM_SIZE_1", "BLOOM_SIZE_2", "TAU_GAMMA_SING", "TAU_GAMMA_DOUB", "TAU_GAMMA_PAR",
                  "TAU_GAMMA_OPP", "TAU_MAX_DEATH_CPT", "HIGHEST_POP")

SYS_FCIDUMP_PCHB, SYS_HUBBARD_RS, SYS_HUBBARD_K = 1, 2, 3
FLAG_REMOVED, FLAG_DETERM_PARENT, FLAG_INITIATOR, FLAG_DETERMINISTIC = 0, 1, 13, 19
FLAG_TRIAL, FLAG_CONNECTED = 2, 3


class Config(C.Structure):
    _fields_ = [
        ("nel", C.c_int32), ("nbasis", C.c_int32), ("nifd", C.c_int32), ("niftot", C.c_int32),
        ("nocc_alpha", C.c_int32), ("nocc_beta", C.c_int32), ("nranks", C.c_int32), ("rank", C.c_int32),
        ("device", C.c_int32), ("balance_blocks", C.c_int32), ("max_walkers", C.c_int64),
        ("max_spawned", C.c_int64), ("system_type", C.c_int32), ("t_trunc_initiator", C.c_int32),
        ("t_all_real_coeff", C.c_int32), ("t_real_spawn_cutoff", C.c_int32), ("t_death_before_comms", C.c_int32),
        ("t_init_coherent_rule", C.c_int32), ("t_no_brillouin", C.c_int32), ("t_exch", C.c_int32),
        ("t_semi_stochastic", C.c_int32), ("t_core_inits", C.c_int32), ("t_tau_search", C.c_int32),
        ("t_consider_par_bias", C.c_int32), ("t_hphf", C.c_int32), ("reserved0", C.c_int32),
        ("initiator_walk_no", C.c_double),
        ("real_spawn_cutoff", C.c_double), ("occupied_thresh", C.c_double), ("av_mc_excits", C.c_double),
        ("hii", C.c_double), ("ecore", C.c_double), ("seed", C.c_uint64),
        ("random_orb_index", C.POINTER(C.c_int32)), ("random_hash2", C.POINTER(C.c_int32)),
        ("load_balance_mapping", C.POINTER(C.c_int32)), ("ilut_ref", C.POINTER(C.c_int64)),
    ]


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class EngineError(RuntimeError):
    pass


class Engine:
    """One rank of the FCIQMC engine (one GPU)."""

    def __init__(self, params, lib_path=None, prefix="neci_gpu_"):
        """params: dict with the Config fields (arrays as numpy)."""
        # NECI_GPU_LIB selects a tuning variant of the CUDA library (_build.build_gpu_variant); never a CPU library
        lib_path = lib_path or os.environ.get("NECI_GPU_LIB") or GPU_LIB
        if not os.path.exists(lib_path):
            raise EngineError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the engine has no CPU fallback)" % lib_path)
        self.lib = C.CDLL(lib_path, mode=C.RTLD_GLOBAL)
        self.prefix = prefix
        self.params = dict(params)
        self._keep = {}
        cfg = Config()
        for name, ctype in Config._fields_:
            v = params[name]
            if name in ("random_orb_index", "random_hash2", "load_balance_mapping"):
                arr = _i32(v); self._keep[name] = arr; v = _p(arr, C.c_int32)
            elif name == "ilut_ref":
                arr = _i64(v); self._keep[name] = arr; v = _p(arr, C.c_int64)
            setattr(cfg, name, v)
        self.cfg = cfg
        self.nw = cfg.nifd + 1
        self.W = cfg.niftot + 1
        self.h = C.c_void_p()
        f = self._fn("init"); f.restype = C.c_int
        rc = f(C.byref(cfg), C.byref(self.h))
        self._check(rc, "init")

    # -- plumbing --------------------------------------------------------------
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _check(self, rc, what):
        if rc != 0:
            msg = ""
            if self.h and hasattr(self.lib, self.prefix + "last_error"):
                g = getattr(self.lib, self.prefix + "last_error"); g.restype = C.c_char_p
                msg = (g(self.h) or b"").decode()
            raise EngineError("%s%s failed (rc=%d): %s" % (self.prefix, what, rc, msg))

    def close(self):
        if self.h:
            self._fn("finalize")(self.h)
            self.h = C.c_void_p()
        self.free_host()

    def free_host(self):
        """Releases every page-locked array alloc_host handed out (arrays must not be used afterwards)."""
        pinned = self._keep.pop("_pinned", []) if hasattr(self, "_keep") else []
        f = getattr(self.lib, self.prefix + "free_host", None)
        for p, _ in pinned:
            if f is not None and p.value:
                f(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- system tables -----------------------------------------------------------
    def set_system_fcidump(self, umat, tmat):
        umat, tmat = _f64(umat), _f64(tmat)
        self._check(self._fn("set_system_fcidump")(self.h, _p(umat, C.c_double), C.c_int64(umat.size), _p(tmat, C.c_double)),
                    "set_system_fcidump")

    def set_pchb(self, t):
        a = dict(probs=_f64(t["probs"]), bias=_f64(t["bias"]), alias=_i32(t["alias"]), p_exch=_f64(t["p_exch"]),
                 tgt=_i32(t["tgt_orbs"]), cls=_i32(t["class_of_spinorb"]))
        self._check(self._fn("set_pchb")(
            self.h, C.c_int32(t["n_spat"]), C.c_int32(t["ij_max"]), C.c_int32(t["ab_max"]), _p(a["probs"], C.c_double),
            _p(a["bias"], C.c_double), _p(a["alias"], C.c_int32), _p(a["p_exch"], C.c_double), _p(a["tgt"], C.c_int32),
            C.c_double(t["p_singles"]), C.c_double(t["p_doubles"]), C.c_double(t["p_parallel"]),
            C.c_int32(t["n_classes"]), _p(a["cls"], C.c_int32)), "set_pchb")

    def set_pchb_particles(self, mode, p_first, p_second):
        """PCHB particle selection: 0 UNIF-UNIF (default), 1 FULL-FULL with the probability tables of the particle selector."""
        a, b = _f64(p_first), _f64(p_second)
        self._check(self._fn("set_pchb_particles")(self.h, C.c_int32(mode), _p(a, C.c_double), _p(b, C.c_double)), "set_pchb_particles")

    def set_excit_probs(self, p_singles, p_doubles, p_parallel):
        self._check(self._fn("set_excit_probs")(self.h, C.c_double(p_singles), C.c_double(p_doubles), C.c_double(p_parallel)),
                    "set_excit_probs")

    def set_system_hubbard_rs(self, max_neigh, neighbours, tmat, uhub):
        nb, tm = _i32(neighbours), _f64(tmat)
        self._check(self._fn("set_system_hubbard_rs")(self.h, C.c_int32(max_neigh), _p(nb, C.c_int32), _p(tm, C.c_double),
                                                       C.c_double(uhub)), "set_system_hubbard_rs")

    def set_system_hubbard_k(self, n_k, ksum, kdiff, eps_k, u_over_n):
        ks, kd, ek = _i32(ksum), _i32(kdiff), _f64(eps_k)
        self._check(self._fn("set_system_hubbard_k")(self.h, C.c_int32(n_k), _p(ks, C.c_int32), _p(kd, C.c_int32),
                                                      _p(ek, C.c_double), C.c_double(u_over_n)), "set_system_hubbard_k")

    def set_core_space(self, row_ptr, col, val, sizes, displs, core_iluts):
        rp, cl, vl, sz, dp, ci = _i64(row_ptr), _i32(col), _f64(val), _i32(sizes), _i32(displs), _i64(core_iluts)
        n_local = rp.size - 1                      # core_iluts: ALL core determinants (replicated), rank-major
        self._check(self._fn("set_core_space")(self.h, C.c_int64(n_local), _p(rp, C.c_int64), _p(cl, C.c_int32),
                                                _p(vl, C.c_double), _p(sz, C.c_int32), _p(dp, C.c_int32),
                                                _p(ci, C.c_int64)), "set_core_space")

    def build_core_space(self, sizes, displs, core_iluts):
        """Core space with the sparse core Hamiltonian built on the device; returns this rank's nnz."""
        sz, dp, ci = _i32(sizes), _i32(displs), _i64(core_iluts)
        nnz = C.c_int64(0)
        self._check(self._fn("build_core_space")(self.h, _p(sz, C.c_int32), _p(dp, C.c_int32), _p(ci, C.c_int64),
                                                  C.byref(nnz)), "build_core_space")
        return int(nnz.value)

    def get_core_hamiltonian(self, n_local):
        rp = np.zeros(n_local + 1, dtype=np.int64)
        self._check(self._fn("get_core_hamiltonian")(self.h, _p(rp, C.c_int64), None, None), "get_core_hamiltonian")
        col = np.zeros(int(rp[-1]), dtype=np.int32); val = np.zeros(int(rp[-1]))
        self._check(self._fn("get_core_hamiltonian")(self.h, _p(rp, C.c_int64), _p(col, C.c_int32), _p(val, C.c_double)),
                    "get_core_hamiltonian")
        return dict(row_ptr=rp, col=col, val=val)

    def set_trial_space(self, trial_iluts, trial_amps, con_iluts, con_amps):
        ti, ta = _i64(trial_iluts).reshape(-1, self.nw), _f64(trial_amps)
        ci, ca = _i64(con_iluts).reshape(-1, self.nw), _f64(con_amps)
        self._check(self._fn("set_trial_space")(self.h, C.c_int64(ti.shape[0]), _p(ti, C.c_int64), _p(ta, C.c_double),
                                                 C.c_int64(ci.shape[0]), _p(ci, C.c_int64), _p(ca, C.c_double)),
                    "set_trial_space")

    # -- walkers -------------------------------------------------------------------
    def upload_walkers(self, dets, gdata_diag=None, gdata_offdiag=None):
        dets = _i64(dets).reshape(-1, self.W)
        gd = _f64(gdata_diag) if gdata_diag is not None else None
        go = _f64(gdata_offdiag) if gdata_offdiag is not None else None
        self._check(self._fn("upload_walkers")(self.h, _p(dets, C.c_int64), C.c_int64(dets.shape[0]),
                                                _p(gd, C.c_double), _p(go, C.c_double)), "upload_walkers")

    def download_walkers(self, with_gdata=True):
        n = C.c_int64(0)
        self._check(self._fn("download_walkers")(self.h, None, C.byref(n), None, None), "download_walkers")
        nn = n.value
        dets = np.zeros((max(nn, 1), self.W), dtype=np.int64)
        gd = np.zeros(max(nn, 1)); go = np.zeros(max(nn, 1))
        self._check(self._fn("download_walkers")(self.h, _p(dets, C.c_int64), C.byref(n),
                                                  _p(gd, C.c_double) if with_gdata else None,
                                                  _p(go, C.c_double) if with_gdata else None), "download_walkers")
        return dets[:nn], gd[:nn], go[:nn]

    def download_occupied(self, min_weight=0.0):
        """Occupied determinants (|sign| > min_weight) compacted in list order + their gdata rows: the POPSFILE body."""
        n = C.c_int64(0)
        self._check(self._fn("download_occupied")(self.h, C.c_double(min_weight), None, C.byref(n), None, None), "download_occupied")
        nn = n.value
        dets = np.zeros((max(nn, 1), self.W), dtype=np.int64); gd = np.zeros(max(nn, 1)); go = np.zeros(max(nn, 1))
        if nn:
            self._check(self._fn("download_occupied")(self.h, C.c_double(min_weight), _p(dets, C.c_int64), C.byref(n),
                                                       _p(gd, C.c_double), _p(go, C.c_double)), "download_occupied")
        return dets[:nn], gd[:nn], go[:nn]

    # -- hot path ---------------------------------------------------------------------
    def iterate(self, tau, diag_sft, it):
        st = np.zeros(ST_COUNT)
        self._check(self._fn("iterate")(self.h, C.c_double(tau), C.c_double(diag_sft), C.c_int64(it), _p(st, C.c_double)),
                    "iterate")
        return st

    def iterate_into(self, tau, diag_sft, it, out):
        """iterate() writing the statistics vector into a caller-owned float64 array of ST_COUNT entries: no allocation
        and no attribute look-ups per call (the per-iteration host overhead counts in a 0.8 ms iteration)."""
        fast = self.__dict__.get("_iter_fast")
        if fast is None:
            f = getattr(self.lib, self.prefix + "iterate")
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p]
            fast = self.__dict__["_iter_fast"] = f
        rc = fast(self.h, tau, diag_sft, it, out.ctypes.data)
        if rc != 0:
            self._check(rc, "iterate")

    def iterate_host(self, dets_buf, n, gd_buf, go_buf, tau, diag_sft, it):
        """dets_buf: int64 [max_walkers, W] host buffer holding n records; updated in place."""
        st = np.zeros(ST_COUNT)
        nn = C.c_int64(n)
        null = C.POINTER(C.c_double)()
        self._check(self._fn("iterate_host")(self.h, _p(dets_buf, C.c_int64), C.byref(nn),
                                              _p(gd_buf, C.c_double) if gd_buf is not None else null,
                                              _p(go_buf, C.c_double) if go_buf is not None else null, C.c_double(tau),
                                              C.c_double(diag_sft), C.c_int64(it), _p(st, C.c_double)), "iterate_host")
        return st, nn.value

    def annihilate(self, spawned, it):
        sp = _i64(spawned).reshape(-1, self.W)
        st = np.zeros(ST_COUNT)
        self._check(self._fn("annihilate")(self.h, _p(sp, C.c_int64), C.c_int64(sp.shape[0]), C.c_int64(it),
                                            _p(st, C.c_double)), "annihilate")
        return st

    # -- measurement -----------------------------------------------------------------------
    def timer_start(self):
        self._check(self._fn("timer_start")(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_double(0.0)
        self._check(self._fn("timer_stop")(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self):
        f = getattr(self.lib, self.prefix + "launch_count"); f.restype = C.c_int64
        return int(f(self.h))

    def alloc_host(self, shape, dtype):
        """Page-locked numpy array (neci_gpu_alloc_host); freed with free_host."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        f = getattr(self.lib, self.prefix + "alloc_host"); f.restype = C.c_int
        if f(C.c_int64(n), C.byref(p)) != 0:
            raise EngineError("alloc_host(%d bytes) failed" % n)
        buf = (C.c_uint8 * max(n, 8)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._keep.setdefault("_pinned", []).append((p, buf))
        return arr

    # -- multi-rank ---------------------------------------------------------------------
    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        f = self._fn("nccl_unique_id")
        if f(buf) != 0:
            raise EngineError("nccl_unique_id failed")
        return bytes(buf)

    def nccl_init(self, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self._fn("nccl_init")(self.h, buf), "nccl_init")

    def p2p_handle(self):
        buf = (C.c_uint8 * 64)()
        self._check(self._fn("p2p_handle")(self.h, buf), "p2p_handle")
        return bytes(buf)

    def p2p_open(self, handles):
        """handles: list of the 64-byte IPC handles of all ranks, in rank order."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._fn("p2p_open")(self.h, buf), "p2p_open")

    def block_populations(self):
        out = np.zeros(self.cfg.balance_blocks)
        self._check(self._fn("block_populations")(self.h, _p(out, C.c_double)), "block_populations")
        return out

    def rebalance(self, new_mapping):
        m = _i32(new_mapping)
        self._check(self._fn("rebalance")(self.h, _p(m, C.c_int32)), "rebalance")
        self.params["load_balance_mapping"] = m.copy()

    def synthetic_list(self, n_dets_total, seed):
        """Benchmark set-up: this rank's share of a frozen synthetic list of n_dets_total random determinants,
        generated on the device (include/neci_gpu.h).  Returns the number of records taken in on this rank."""
        n = C.c_int64(0)
        self._check(self._fn("synthetic_list")(self.h, C.c_int64(int(n_dets_total)), C.c_uint64(int(seed)), C.byref(n)), "synthetic_list")
        return n.value

    # -- probes ---------------------------------------------------------------------------
    def probe_det_node(self, iluts):
        il = _i64(iluts).reshape(-1, self.nw)
        n = il.shape[0]
        b = np.zeros(n, dtype=np.int32); nd = np.zeros(n, dtype=np.int32)
        self._check(self._fn("probe_det_node")(self.h, C.c_int64(n), _p(il, C.c_int64), _p(b, C.c_int32), _p(nd, C.c_int32)),
                    "probe_det_node")
        return b, nd

    def probe_helement(self, iluts_i, iluts_j):
        a = _i64(iluts_i).reshape(-1, self.nw); b = _i64(iluts_j).reshape(-1, self.nw)
        out = np.zeros(a.shape[0])
        self._check(self._fn("probe_helement")(self.h, C.c_int64(a.shape[0]), _p(a, C.c_int64), _p(b, C.c_int64),
                                                _p(out, C.c_double)), "probe_helement")
        return out

    def probe_gen_excit(self, iluts, attempts, it):
        il = _i64(iluts).reshape(-1, self.nw); at = _i32(attempts)
        n = il.shape[0]
        out = dict(ilut_j=np.zeros((n, self.nw), dtype=np.int64), ic=np.zeros(n, dtype=np.int32),
                   ex=np.zeros((n, 4), dtype=np.int32), parity=np.zeros(n, dtype=np.int32),
                   pgen=np.zeros(n), hel=np.zeros(n))
        self._check(self._fn("probe_gen_excit")(self.h, C.c_int64(n), _p(il, C.c_int64), _p(at, C.c_int32), C.c_int64(it),
                                                 _p(out["ilut_j"], C.c_int64), _p(out["ic"], C.c_int32),
                                                 _p(out["ex"], C.c_int32), _p(out["parity"], C.c_int32),
                                                 _p(out["pgen"], C.c_double), _p(out["hel"], C.c_double)),
                    "probe_gen_excit")
        return out
