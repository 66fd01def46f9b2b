#!/usr/bin/env python
"""bench.py -- spawn attempts/s and iteration time of the FCIQMC hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the CUDA engine (one rank per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference arm: CPU restatement of the reference's
                                                           algorithm (oracle/) on all host cores, bounded sample

A "step" is one FCIQMC iteration (one PerformFCIMCycPar, src/FciMCPar.F90:1177) over the resident walker
list.  Workload at N = 1: BASELINE.json configs[1] -- N2 cc-pVDZ-sized synthetic FCIDUMP (14 electrons in 28
spatial orbitals), i-FCIQMC with the PCHB generator, 1e7 walkers on one B200 (frozen synthetic start list per
SURVEY.md section 8d: distinct uniformly random determinants, |sign| = round(1 + Exp(1))).  For N > 1 the same
per-GPU population is hashed over the ranks with DetermineDetNode and spawns are exchanged over NCCL (weak scaling).

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spawn_attempts_per_sec"
UNIT = "attempts/s"


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (n_spat, nel, tau, description)
    "n2_14e28o_pchb": (28, 14, 2.0e-5, "N2 cc-pVDZ-sized synthetic FCIDUMP 14e/28o, i-FCIQMC, PCHB (BASELINE configs[1])"),
    "cr2_24e30o_pchb": (30, 24, 1.0e-5, "Cr2-sized synthetic FCIDUMP 24e/30o, i-FCIQMC, PCHB (BASELINE configs[4])"),
    "semistoch_20e40o_pchb": (40, 20, 4.0e-6, "semi-stochastic i-FCIQMC on a synthetic FCIDUMP 20e/40o, PCHB, real "
                              "coefficients, core space = mp1-core <--core-size> (largest first-order amplitudes among "
                              "the reference's singles and doubles), trial wavefunction = mp1-trial <--trial> "
                              "(BASELINE configs[3])"),
}
SEMISTOCH = ("semistoch_20e40o_pchb",)


def build_system(workload, particle_selection="UNIF-UNIF"):
    from neci_stable_b200 import host
    if workload in WORKLOADS:
        n_spat, nel, tau, _ = WORKLOADS[workload]
        return host.random_fcidump_system(n_spat, nel, sparse=1.0, sparse_t=1.0, seed=25, particle_selection=particle_selection), tau
    if workload == "hubk_6x6":
        return host.hubbard_k_system(6, 6, U=4.0), 5.0e-4
    if workload == "hubrs_4x4":
        return host.hubbard_rs_system(4, 4, U=4.0), 5.0e-3
    raise SystemExit("unknown workload %s" % workload)


_COMBO_CACHE = {}


def _combination_masks(n, k):
    """All C(n, k) k-subsets of n spatial orbitals as bit masks over spatial orbitals (uint64), built by the
    Pascal recursion with numpy concatenations."""
    key = (n, k)
    if key in _COMBO_CACHE:
        return _COMBO_CACHE[key]
    if k == 0:
        out = np.zeros(1, dtype=np.uint64)
    elif k == n:
        out = np.array([(1 << n) - 1], dtype=np.uint64)
    else:
        out = np.concatenate([_combination_masks(n - 1, k), _combination_masks(n - 1, k - 1) | np.uint64(1 << (n - 1))])
    _COMBO_CACHE[key] = out
    return out


def _spread_spin(masks, n_spat, alpha, nw):
    """Spatial-orbital masks -> spin-orbital occupation words: spatial i (1-based) -> spin orbital 2i (alpha) or 2i-1 (beta)."""
    words = np.zeros((masks.shape[0], nw), dtype=np.uint64)
    for i in range(n_spat):
        orb = 2 * (i + 1) - (0 if alpha else 1)
        bit = orb - 1
        has = (masks >> np.uint64(i)) & np.uint64(1)
        words[:, bit // 64] |= has << np.uint64(bit % 64)
    return words


def random_walker_records(system, n_dets, seed, keep=None, keep_frac=1.0, chunk=1 << 22):
    """Distinct uniformly random determinants (n_alpha of n_spat, n_beta of n_spat), signs +-round(1 + Exp(1)).
    keep(iluts) -> bool mask selects the determinants this rank owns."""
    import math
    rng = np.random.default_rng(seed)
    ns = system.nbasis // 2
    nw = system.nw
    use_table = max(math.comb(ns, system.nocc_alpha), math.comb(ns, system.nocc_beta)) <= 30_000_000
    if use_table:
        ta = _spread_spin(_combination_masks(ns, system.nocc_alpha), ns, True, nw)
        tb = _spread_spin(_combination_masks(ns, system.nocc_beta), ns, False, nw)
    out = []
    have = 0
    while have < n_dets:
        m = int(min(chunk, max(4096, 1.05 * (n_dets - have) / keep_frac + 1024)))
        if use_table:
            words = ta[rng.integers(0, ta.shape[0], m)] | tb[rng.integers(0, tb.shape[0], m)]
        else:
            words = np.zeros((m, nw), dtype=np.uint64)
            for nocc, alpha in ((system.nocc_alpha, True), (system.nocc_beta, False)):
                pick = np.argpartition(rng.random((m, ns)), nocc - 1, axis=1)[:, :nocc]  # nocc distinct spatial orbitals
                orb = 2 * (pick + 1) - (0 if alpha else 1)                                # 1-based spin orbital
                bit = (orb - 1).astype(np.uint64)
                for w in range(nw):
                    sel = (bit // 64) == w
                    contrib = np.where(sel, np.uint64(1) << (bit % np.uint64(64)), np.uint64(0))
                    words[:, w] |= np.bitwise_or.reduce(contrib, axis=1)
        if keep is not None:
            words = words[keep(words.view(np.int64))]
        out.append(words)
        have += words.shape[0]
    words = np.concatenate(out)
    # distinct
    if nw == 1:
        _, idx = np.unique(words[:, 0], return_index=True)
    else:
        _, idx = np.unique(words, axis=0, return_index=True)
    words = words[np.sort(idx)][:n_dets]
    n = words.shape[0]
    mag = np.round(1.0 + rng.exponential(1.0, n))
    sgn = mag * rng.choice(np.array([-1.0, 1.0]), n)
    rec = np.zeros((n, nw + 2), dtype=np.int64)
    rec[:, :nw] = words.view(np.int64)
    rec[:, nw] = sgn.view(np.int64)
    return rec


# ---------------------------------------------------------------------------------------------
# semi-stochastic set-up (BASELINE configs[3]): what init_semi_stochastic / init_trial_wf hand to the engine
# ---------------------------------------------------------------------------------------------
def semistoch_space(system, hii, params, nranks, core_size, n_trial):
    """Core space = `mp1-core <core_size>`: the reference and the singles and doubles with the largest first-order
    amplitudes ("all singles+doubles truncated"), laid out rank-major as store_whole_core_space does; trial
    space = `mp1-trial <n_trial>`, its largest members.  Returns dict(iluts, sizes, displs, weights (signed first-order amplitudes), trial)."""
    from neci_stable_b200 import host
    sd = host.sing_doub_space(system)
    ref = np.repeat(sd[:1], sd.shape[0], 0)
    h0 = host.get_helement(system, ref, sd)
    # `mp1-core`: first-order amplitude H_0j / (F_00 - F_jj) with the Fock orbital energies of the reference
    # (return_mp1_amp_and_mp2_energy, src/semi_stoch_procs.F90:2143-2218)
    eps = fock_orbital_energies(system)
    occ = ((sd.view(np.uint64)[:, :, None] >> np.arange(64, dtype=np.uint64)[None, None, :]) & np.uint64(1)).astype(np.float64)
    occ = occ.reshape(sd.shape[0], -1)[:, :system.nbasis]
    f = occ @ np.repeat(eps, 2)
    amp = h0 / np.where(np.abs(f[0] - f) > 1e-9, f[0] - f, -1e-9)
    amp[0] = 1.0
    a = np.abs(amp)
    a[0] = np.inf                                           # the reference first
    pick = np.argsort(-a, kind="stable")[:min(core_size, sd.shape[0])]
    core, camp = sd[pick], amp[pick]
    if nranks > 1:
        _, nodes = host.det_node(params, core, system.nw)
    else:
        nodes = np.zeros(core.shape[0], dtype=np.int32)
    il, sizes, displs = host.layout_core_space(core, nodes, nranks)
    # amplitudes in the new order
    key = {tuple(r): a for r, a in zip(core.tolist(), camp)}
    w = np.array([key[tuple(r)] for r in il.tolist()])
    trial = core[:n_trial].copy() if n_trial > 0 else None
    return dict(iluts=il, sizes=sizes, displs=displs, weights=w, trial=trial)


def fock_orbital_energies(system):
    """Diagonal Fock elements over spatial orbitals for the closed-shell reference of a synthetic FCIDUMP system:
    eps_p = h_pp + sum_{j occupied} [2 (pp|jj) - (pj|jp)] -- what the reference takes as Arr when the FCIDUMP carries no
    orbital energies."""
    ns = system.nbasis // 2
    umat, tmat, nb = system.tables["umat"], system.tables["tmat"], system.nbasis
    tri = lambda a, b: a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a
    um = lambda i, j, k, l: umat[tri(tri(i, k), tri(j, l)) - 1]          # <ij|kl>
    nocc = system.nocc_alpha
    assert system.nocc_alpha == system.nocc_beta
    eps = np.zeros(ns)
    for p in range(1, ns + 1):
        e = tmat[(2 * p - 2) + nb * (2 * p - 2)]
        for j in range(1, nocc + 1):
            e += 2.0 * um(p, j, p, j) - um(p, j, j, p)
        eps[p - 1] = e
    return eps


def semistoch_records(system, space, rank, l1_total):
    """CurrentDets records of this rank's core determinants: deterministic + initiator flags, signs = first-order
    amplitudes scaled so that the whole core space carries l1_total walkers (sum |sign| over all ranks)."""
    from neci_stable_b200 import capi
    lo, n = int(space["displs"][rank]), int(space["sizes"][rank])
    nw = system.nw
    rec = np.zeros((n, nw + 2), dtype=np.int64)
    rec[:, :nw] = space["iluts"][lo:lo + n]
    scale = l1_total / float(np.abs(space["weights"]).sum())
    rec[:, nw] = (scale * space["weights"][lo:lo + n]).view(np.int64)
    rec[:, nw + 1] = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    return rec


def semistoch_apply(engine, system, hii, space, rank, build="host"):
    """Hands core and trial space over through the C ABI.  build = "host": this rank's rows of the sparse core
    Hamiltonian come from the host library's threads (neci_gpu_set_core_space); "device": the engine builds them
    itself (neci_gpu_build_core_space).  Returns (nnz of this rank, seconds spent building)."""
    from neci_stable_b200 import host
    t0 = time.perf_counter()
    if build == "device":
        nnz = engine.build_core_space(space["sizes"], space["displs"], space["iluts"])
        dt = time.perf_counter() - t0
        c = {"row_ptr": [nnz]}
    else:
        c = host.core_hamiltonian(system, space["iluts"], hii, displ=int(space["displs"][rank]), n_local=int(space["sizes"][rank]))
        dt = time.perf_counter() - t0
        engine.set_core_space(c["row_ptr"], c["col"], c["val"], space["sizes"], space["displs"], space["iluts"])
    if space["trial"] is not None:
        ti, ta, ci, ca, _ = host.trial_space(system, space["trial"])
        engine.set_trial_space(ti, ta, ci, ca)
    return int(c["row_ptr"][-1]), dt


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        """Starts the poller.  nvidia-smi's own start-up (NVML initialisation over all GPUs of the box) takes a second
        on an 8-GPU node and stalls kernel launches of every process while it lasts -- inside a 20 ms timed region that
        was a factor of ten on ms_per_step -- so the poller is started during set-up, wait_ready() makes sure it is
        past that point before the timed region begins, and only the samples taken inside the window count."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout=10.0):
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def window(self, t_begin, t_end):
        """Restricts the report to the samples that arrived between the two perf_counter stamps (one polling period of
        slack at the end: a sample describes the interval before it)."""
        self.win = (t_begin, t_end + 0.06)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        win = getattr(self, "win", None)
        lines = [ln for (t, ln) in self.lines if win is None or win[0] <= t <= win[1]]
        if not lines and self.lines and win is not None:        # region shorter than the polling period: the nearest sample
            lines = [min(self.lines, key=lambda tl: abs(tl[0] - 0.5 * (win[0] + win[1])))[1]]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle (restatement of the reference algorithm), ranks played by host threads
# ---------------------------------------------------------------------------------------------
def cpu_run(workload, n_dets_total, steps, warmup, cores, seed=5, core_size=0, n_trial=0):
    """Times `steps` iterations of the CPU restatement on `cores` threads (one NECI 'rank' per thread,
    determinants partitioned by DetermineDetNode, in-memory all-to-all).  Returns attempts/s etc."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build_cpu_side()
    import helpers
    from neci_stable_b200 import capi, host, driver
    from neci_stable_b200.capi import ST
    system, tau = build_system(workload)
    hii = driver.diag_energy(system, system.ref_orbs)
    nr = max(1, cores)
    per = int(n_dets_total * 3 // nr + 1000)
    oracles = []
    semi = workload in SEMISTOCH
    per += core_size if semi else 0
    for r in range(nr):
        params = host.make_params(system, hii, max_walkers=per, max_spawned=max(per, 200000), nranks=nr, rank=r, seed=11,
                                  semi_stochastic=semi, all_real_coeff=semi)
        o = helpers.Oracle(params)
        system.apply(o)
        oracles.append(o)
    rec = random_walker_records(system, n_dets_total, seed)
    space = None
    if semi:
        space = semistoch_space(system, hii, params, nr, core_size, n_trial)
        rec = rec[~host.rows_in(rec[:, :system.nw], space["iluts"])]
    _, node = oracles[0].probe_det_node(rec[:, :system.nw])
    tot = tot_rand = float(np.abs(rec[:, system.nw].view(np.float64)).sum())
    for r in range(nr):
        mine = rec[node == r]
        if semi:
            crec = semistoch_records(system, space, r, l1_total=tot_rand)
            tot += float(np.abs(crec[:, system.nw].view(np.float64)).sum())
            mine = np.concatenate([crec, mine])
        oracles[r].upload_walkers(mine)
        if semi:
            semistoch_apply(oracles[r], system, hii, space, r)
    sft = 0.0
    attempts = 0.0
    t_used = 0.0
    for it in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        st = helpers.world_iterate(oracles, tau, sft, it, nthreads=nr)
        dt = time.perf_counter() - t0
        new = st[:, ST["TOTPARTS"]].sum()
        if it <= warmup:
            if new > 0 and tot > 0:
                sft -= 0.5 * np.log(new / tot) / tau
        else:
            attempts += st[:, ST["NVALIDEXCITS"]].sum() + st[:, ST["NINVALIDEXCITS"]].sum()
            t_used += dt
        tot = new
    for o in oracles:
        o.close()
    return dict(value=attempts / t_used, ms_per_step=1e3 * t_used / steps, attempts_per_step=attempts / steps,
                walkers=tot, cores=nr)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="n2_14e28o_pchb")
    ap.add_argument("--walkers", type=float, default=1.0e7, help="walkers (sum |sign|) per GPU")
    ap.add_argument("--cpu-sample-walkers", type=float, default=2.0e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="default workload only: skip the short runs of the other BASELINE configurations")
    ap.add_argument("--core-size", type=int, default=100000, help="semi-stochastic workloads: determinants in the core space")
    ap.add_argument("--core-build", default="host", choices=["host", "device"],
                    help="semi-stochastic workloads: who builds the sparse core Hamiltonian")
    ap.add_argument("--trial", type=int, default=10, help="semi-stochastic workloads: determinants in the trial space (0 = none)")
    ap.add_argument("--particle-selection", default="UNIF-UNIF", choices=["UNIF-UNIF", "FULL-FULL", "UNIF-FULL"],
                    help="PCHB workloads: PCHB_ParticleSelection of the doubles generator (BASELINE configs quote UNIF-UNIF; "
                         "FULL-FULL is what the reference's own PCHB regression input selects)")
    ap.add_argument("--list", default="auto", choices=["auto", "host", "device"],
                    help="where the frozen start list is generated: numpy on the host (uploaded), or on the device "
                         "(neci_gpu_synthetic_list); auto = device on several GPUs and from 5e7 walkers per GPU")
    ap.add_argument("--ref-fraction", type=float, default=None,
                    help="fraction of all walkers placed on the reference determinant (a converged FCIQMC wavefunction holds a "
                         "few per cent there; it is what unbalances the hash partition).  Default: 0.05 with --load-balance, else 0")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --walkers per GPU (default, the driver's scaling run); strong: --walkers in total, "
                         "split over the GPUs (BASELINE configs[4])")
    ap.add_argument("--load-balance", action="store_true",
                    help="100 balancing blocks per rank (load-balance-blocks) and one adjust_load_balance pass "
                         "(block populations -> greedy plan -> device-to-device block moves) during warm-up")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="spawn exchange for N > 1: pushes over NVLink peer memory issued by the spawning kernels (default) or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    desc = WORKLOADS.get(args.workload, (0, 0, 0, args.workload))[3]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_dets = int(args.cpu_sample_walkers / 2.0)
        r = cpu_run(args.workload, n_dets, args.steps, args.warmup, cores, core_size=args.core_size, n_trial=args.trial)
        sample = "%d iterations of a %.3g-walker (%d determinants) list of the same system and distribution" % (
            args.steps, r["walkers"], n_dets)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "walkers": r["walkers"],
                       "note": "CPU restatement of the reference algorithm (the NECI Fortran/MPI build needs gfortran+MPI, "
                               "absent from this image); bounded sample of the GPU workload"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    import __graft_entry__ as g
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "neci_stable_b200", "libneci_gpu.so")):
        g.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    ctx = dict(rank=rank, local_rank=local_rank, world=world, cores=cores, dist=dist, torch=torch)
    line = run_workload(args, ctx, primary=True)

    # ---- short runs of the other BASELINE configurations, so that the driver's one default command leaves evidence
    #      for all of them (value, ms per step, phase split, roofline fractions); no e2e / CPU legs there
    if args.secondary and args.workload == "n2_14e28o_pchb":
        sec = {}
        for wl, extra in (("hubk_6x6", {}), ("cr2_24e30o_pchb", {}), ("semistoch_20e40o_pchb", {"core_build": "device"})):
            a2 = argparse.Namespace(**vars(args))
            a2.workload = wl; a2.steps = max(3, min(args.steps, 8)); a2.warmup = 4
            a2.no_e2e = True; a2.no_cpu_baseline = True; a2.walkers = min(args.walkers, 1.0e7)
            a2.scaling = "weak"; a2.load_balance = False
            for k, v in extra.items():
                setattr(a2, k, v)
            try:
                r = run_workload(a2, ctx, primary=False)
            except Exception as exc:                         # a failed secondary run must not cost the headline line
                r = {"error": "%s: %s" % (type(exc).__name__, exc)}
            if rank == 0:
                if "error" in r:
                    sec[wl] = r
                else:
                    sec[wl] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "gpu_launches", "selfcheck") if k in r}
                    sec[wl]["config"] = {k: r["config"][k] for k in ("description", "walkers_per_gpu", "walkers_total_end",
                                                                       "determinants_total_end", "attempts_per_step",
                                                                       "spawned_per_step", "semi_stochastic") if k in r["config"]}
                    rf = r["roofline"]
                    sec[wl]["phase_ms_per_step"] = rf["phase_ms_per_step"]
                    sec[wl]["roofline"] = {kk: {x: kv[x] for x in ("achieved", "peak", "frac", "ms_per_launch", "algorithmic_bytes_per_launch")}
                                           for kk, kv in rf.get("kernels", {"K1": rf}).items()}
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_workload(args, ctx, primary=True):
    """One workload on the CUDA engine: set-up, warm-up, K timed iterations with the list resident, then (primary
    only) the host-buffer e2e leg and the CPU leg.  Returns the JSON line as a dict on rank 0."""
    rank, local_rank, world, cores, dist, torch = (ctx[k] for k in ("rank", "local_rank", "world", "cores", "dist", "torch"))
    desc = WORKLOADS.get(args.workload, (0, 0, 0, args.workload))[3]
    args = argparse.Namespace(**vars(args))
    from neci_stable_b200 import capi, host, driver
    from neci_stable_b200.capi import ST

    system, tau = build_system(args.workload, args.particle_selection)
    hii = driver.diag_energy(system, system.ref_orbs)
    if args.scaling == "strong":
        args.walkers = args.walkers / world                # total fixed: each GPU holds its 1/N share
    n_dets = int(args.walkers / 2.0)                       # mean |sign| of round(1 + Exp(1)) is ~2.0
    max_walkers = int(3 * n_dets + 100000)
    max_spawned = int(max(2 * args.walkers, 400000))
    semi = args.workload in SEMISTOCH
    # multi-GPU runs and lists of 5e7 walkers and more (the 1e8 - 1e9 walker configurations): the list is generated on
    # the device (neci_gpu_synthetic_list: every rank keeps its share of ONE global list; numpy needs minutes per rank for
    # the large ones and draws N times the candidates on N ranks), and MemoryFacPart / MemoryFacSpawn are sized for a
    # stationary list (the frozen start list shrinks; a spawning pass sends ~0.16 spawns per walker)
    device_list = (args.list == "device") or (args.list == "auto" and not semi and (world > 1 or args.walkers >= 5.0e7))
    if device_list and args.walkers >= 5.0e7:
        max_walkers = int(2.2 * n_dets + 100000)
        max_spawned = int(0.6 * args.walkers)
    if semi:
        # real coefficients (readinput.F90:569 makes them mandatory with a core space): a third of the attempts leave a
        # spawn of RealSpawnCutoff walkers on a new determinant, so the list of this far-from-equilibrium start grows
        # by ~0.2 n_dets per iteration
        max_walkers = int(16 * n_dets + args.core_size + 100000)
        max_spawned *= 2
    params = host.make_params(system, hii, max_walkers=max_walkers, max_spawned=max_spawned, nranks=world, rank=rank,
                              device=local_rank, seed=11, blocks_per_rank=100 if args.load_balance else 1,
                              semi_stochastic=semi, all_real_coeff=semi)
    eng = capi.Engine(params)
    system.apply(eng)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                    # polls through set-up and warm-up; see ClockSampler.start
    if world > 1:
        uid = [eng.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.nccl_init(uid[0])
        if args.exchange == "p2p":
            hs = [None] * world
            dist.all_gather_object(hs, eng.p2p_handle())
            eng.p2p_open(hs)

    keep = None
    if world > 1:
        def keep(il):
            _, node = eng.probe_det_node(il)
            return node == rank
    rec = None
    if not device_list:
        rec = random_walker_records(system, n_dets, seed=1000 + rank, keep=keep, keep_frac=1.0 / world)
    core_info = None
    if semi:
        # the core determinants join the list with their flags; the sparse core Hamiltonian (this rank's rows) and
        # the trial / connected spaces are built by the host library, as the Fortran host would, before the loop
        space = semistoch_space(system, hii, params, world, args.core_size, args.trial)
        rec = rec[~host.rows_in(rec[:, :system.nw], space["iluts"])]
        # half of the population sits in the core space, as in a converged semi-stochastic run (the nominal figure,
        # so that every rank scales the core amplitudes alike)
        rec = np.concatenate([semistoch_records(system, space, rank, l1_total=args.walkers * world), rec])
    if device_list:
        eng.synthetic_list(n_dets * world, 1000)
    else:
        eng.upload_walkers(rec)
    ref_fraction = args.ref_fraction if args.ref_fraction is not None else (0.05 if args.load_balance else 0.0)
    if ref_fraction > 0.0 and not semi:
        # the reference determinant with its share of the population, on the rank that owns it (merged into the list by
        # the annihilation step: AnnihilateSpawnedParts / AddNewHashDet)
        ref_rec = host.record(system, system.ref_orbs, float(round(ref_fraction * args.walkers * world)),
                              1 << capi.FLAG_INITIATOR).reshape(1, -1)
        _, node = eng.probe_det_node(ref_rec[:, :system.nw])
        if world == 1 or int(node[0]) == rank:
            eng.annihilate(ref_rec, 0)
    if semi:
        nnz, t_build = semistoch_apply(eng, system, hii, space, rank, build=args.core_build)
        core_info = {"core_build": args.core_build,"core_size": int(space["iluts"].shape[0]), "core_local": int(space["sizes"][rank]), "nnz_local": nnz,
                     "trial_size": 0 if space["trial"] is None else int(space["trial"].shape[0]),
                     "build_s": t_build}
    tot0 = float(np.abs(rec[:, system.nw].view(np.float64)).sum()) if rec is not None else 0.0
    del rec

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: also settles the shift so that the population is stationary (update_shift with damping 0.5/step)
    sft = 0.0
    tot = allsum(tot0)
    it = 0
    lb_moves = None
    for _ in range(args.warmup):
        it += 1
        st = eng.iterate(tau, sft, it)
        new = allsum(st[ST["TOTPARTS"]])
        if new > 0 and tot > 0:
            sft -= 0.5 * np.log(new / tot) / tau
        tot = new
        if args.load_balance and world > 1 and not semi and it == 2:
            # adjust_load_balance (src/load_balancer.fpp:178-351): block populations summed over the ranks, the greedy
            # plan (identical on every rank), then the block moves through the spawn exchange.  Not with a core
            # space: the reference switches balancing off then (load_balancer.fpp:198-201)
            allp = [None] * world
            dist.all_gather_object(allp, eng.block_populations())
            blocks = np.sum(allp, axis=0)
            old_map = np.asarray(params["load_balance_mapping"], dtype=np.int32)
            new_map, moves = driver.plan_load_balance(blocks, old_map, world)
            eng.rebalance(new_map)
            params["load_balance_mapping"] = np.asarray(new_map, dtype=np.int32)
            before = np.bincount(old_map, weights=blocks, minlength=world)
            after = np.bincount(np.asarray(new_map, dtype=np.int32), weights=blocks, minlength=world)
            lb_moves = {"blocks_moved_in_warmup": len(moves),
                        "rank_walkers_before": [float(x) for x in before], "rank_walkers_after": [float(x) for x in after],
                        "imbalance_before": float(before.max() / before.mean() - 1.0),
                        "imbalance_after": float(after.max() / after.mean() - 1.0)}

    tot_before_timed = tot                                 # global TotParts entering the timed region
    # ---- timed region: K iterations, list resident in HBM
    if rank == 0:
        sampler.wait_ready()
    acc = np.zeros(capi.ST_COUNT)
    t_spawn = t_ann = t_comm = t_det = 0.0
    cons = []                                              # per iteration: (TotParts after, net change the counters claim)
    tot_prev_local = None
    bytes_spawn = 0.0
    launches0 = eng.launch_count()
    # the timed loop does nothing but the calls: every statistics vector goes into a preallocated row, the sums are
    # taken afterwards (per-iteration Python work sits between two iterations on the device time line)
    sts = np.zeros((args.steps, capi.ST_COUNT))
    barrier()
    eng.timer_start()
    w0 = time.perf_counter()
    for k in range(args.steps):
        it += 1
        eng.iterate_into(tau, sft, it, sts[k])
    ms_dev = eng.timer_stop()
    barrier()
    w1 = time.perf_counter()
    for st in sts:
        acc += st
        cons.append((st[ST["TOTPARTS"]], st[ST["NOBORN"]] - st[ST["NODIED"]] - st[ST["ANNIHILATED"]] - st[ST["NOABORTED"]]
                     - st[ST["NOREMOVED"]], st[ST["NSPAWNED_SENT"]], st[ST["NSPAWNED_RECV"]]))
        t_spawn += st[ST["TIME_SPAWN_MS"]]; t_ann += st[ST["TIME_ANNIHIL_MS"]]; t_comm += st[ST["TIME_COMM_MS"]]
        t_det += st[ST["TIME_DETERM_MS"]]
        # algorithmic bytes of the spawn/death kernels (DESIGN.md "K1"): SoA record (8*nw + 8 sign + 4 flags) and diagH per
        # slot, offdiagH per occupied slot, sign+flag write-back per occupied slot, one AoS record per spawn
        n_slot = st[ST["TOTWALKERS"]]; n_occ = n_slot - st[ST["HOLESINLIST"]]
        bytes_spawn += n_slot * (8 * system.nw + 12 + 8) + n_occ * 8 + n_occ * 12 + st[ST["NSPAWNED_SENT"]] * 8 * system.W
    wall = w1 - w0
    if rank == 0:
        sampler.window(w0, w1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    ms = allmax(ms_dev)
    attempts = allsum(acc[ST["NVALIDEXCITS"]] + acc[ST["NINVALIDEXCITS"]])
    value = attempts / (ms * 1e-3)
    walkers_end = allsum(st[ST["TOTPARTS"]])
    dets_end = allsum(st[ST["TOTWALKERS"]] - st[ST["HOLESINLIST"]])
    spawned = allsum(acc[ST["NSPAWNED_SENT"]])

    # ---- self-check (the device-side analogue of the reference's per-iteration consistency checks,
    #      src/FciMCPar.F90:1764,1851): over the timed iterations and summed over the ranks, every spawn record sent was
    #      received; the population changed by exactly born - died - annihilated - aborted - removed (exact for integer
    #      walkers without a core space; determ_projection moves amplitude without counters); a sample of every rank's
    #      list hashes to that rank (DetermineDetNode)
    selfcheck = None
    if cons:
        c = np.array(cons)
        if world > 1:
            t = torch.tensor(c, dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            c = t.cpu().numpy()
        prev = np.concatenate([[tot_before_timed], c[:-1, 0]])
        resid = np.abs((c[:, 0] - prev) - c[:, 1])
        owner_ok = None
        if world > 1:
            d, _, _ = eng.download_walkers()
            occ = d[np.abs(d[:, system.nw].view(np.float64)) > 0][:200000]
            _, node = eng.probe_det_node(occ[:, :system.nw])
            owner_ok = allsum(float(np.count_nonzero(node != rank))) == 0.0
        selfcheck = {"iterations": int(c.shape[0]), "spawns_sent": float(c[:, 2].sum()), "spawns_received": float(c[:, 3].sum()),
                     "sent_equals_received": bool(np.array_equal(c[:, 2], c[:, 3])),
                     "population_residual_max": float(resid.max()),
                     "population_conserved": (bool(resid.max() <= 1e-9 * max(1.0, c[:, 0].max())) if not semi else None),
                     "sampled_owner_is_rank": owner_ok}

    # ---- roofline of the dominant phase (K1: the four kernels of the loop over determinants), measured live with CUDA
    #      events on the engine's stream
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = bytes_spawn / (t_spawn * 1e-3) / 1e9 if t_spawn > 0 else 0.0
    traffic = None
    ncu_extra = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        # the capture was taken on the default workload at 1e7 walkers: only that run may quote it
        if args.workload != "n2_14e28o_pchb" or abs(args.walkers - 1.0e7) > 1.0:
            raise KeyError("no ncu capture for this configuration")
        traffic = tj.get("k1_dram_bytes_per_launch")
        # K1 is bound by instruction issue and L2 latency, not by HBM: the committed ncu capture's
        # issue-slot utilisation and warp-instruction counts of its four kernels are repeated here beside the HBM fraction
        ncu_extra = {"ncu_issue_active_pct": tj.get("k1_issue_active_pct"),
                     "ncu_warp_instructions_per_launch": tj.get("k1_warp_instructions_per_launch"),
                     "ncu_kernel_us": tj.get("k1_kernel_us"),
                     "ncu_source": tj.get("source")}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "K1 = k_walk + k_generate + k_evaluate + k_singles", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": bytes_spawn / args.steps, "ms_per_launch": t_spawn / args.steps,
                "phase_ms_per_step": {"spawn_death": t_spawn / args.steps, "exchange": t_comm / args.steps,
                                      "annihilation": t_ann / args.steps}}
    roofline.update(ncu_extra)
    if semi:
        # K3 determ_projection.  The engine keeps the core Hamiltonian column-blocked with 16-bit
        # columns (kernels.cuh: k_determ_spmv_blocked): 10 bytes per non-zero, per (block, row) chunk its pointer (8)
        # and its partial sum written and read again (16), the local slices of the in/out vectors and one read of the
        # gathered vector.  `achieved` uses these bytes (what the kernels must move); SURVEY 8(d)'s figure for the
        # reference's CSR layout (12 bytes per non-zero) is repeated as csr12_*.  Time: the device span from the start
        # of the iteration to the start of the spawning kernel (core gather + SpMV + finish).
        nb3 = max(1, -(-int(core_info["core_size"]) // 27648))
        b3 = (10.0 * core_info["nnz_local"] + 24.0 * nb3 * core_info["core_local"] + 16.0 * core_info["core_local"]
              + 8.0 * core_info["core_size"])
        b3_csr = 12.0 * core_info["nnz_local"] + 16.0 * core_info["core_local"] + 8.0 * core_info["core_size"]
        ach3 = b3 / (t_det / args.steps * 1e-3) / 1e9 if t_det > 0 else 0.0
        k3 = {"bound": "hbm", "kernel": "k_determ_spmv_blocked", "achieved": ach3, "peak": peak, "unit": "GB/s", "frac": ach3 / peak,
              "traffic": None, "algorithmic_bytes_per_launch": b3, "ms_per_launch": t_det / args.steps,
              "csr12_bytes_per_launch": b3_csr, "csr12_equivalent_gbs": b3_csr / (t_det / args.steps * 1e-3) / 1e9 if t_det > 0 else 0.0,
              "column_blocks": nb3}
        roofline["phase_ms_per_step"]["determ_projection"] = t_det / args.steps
        k1 = {k: roofline[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic",
                                       "algorithmic_bytes_per_launch", "ms_per_launch")}
        roofline["kernels"] = {"K1": k1, "k_determ_spmv_blocked": k3}
        if t_det > t_spawn:                                   # the dominant kernel heads the object
            roofline.update(k3)

    # ---- end to end through the C ABI with HOST buffers: CurrentDets + gdata in pinned host memory, uploaded, iterated
    #      and downloaded inside the timed region (neci_gpu_iterate_host)
    e2e = None
    if semi:
        # the host-buffer path re-uploads the list every iteration, which would also re-locate the core space; for
        # this workload only the resident API (the deployment path) is timed end to end
        e2e = {"value": attempts / wall, "unit": UNIT, "h2d_bytes_per_step": 24, "d2h_bytes_per_step": 8 * capi.ST_COUNT + 128,
               "api": "neci_gpu_iterate: list, core Hamiltonian and trial tables resident in HBM; host passes tau/shift/iter "
                      "and reads the statistics vector every iteration (wall clock over the same K steps)"}
    elif not args.no_e2e and primary:
        # Host-authoritative CurrentDets in page-locked host memory: every step uploads the whole list (H2D inside the
        # timed region) and ends with the host list current.  The engine writes the records an iteration changes
        # through to the host arrays (host mirror, neci_gpu_iterate_host) instead of copying the whole list back.
        # Two variants are timed: the host also passes / receives global_determinant_data (H_ii - Hii, H_0i per slot),
        # or leaves that derived data to the engine (the headline: it is a function of the determinant alone).
        W = system.W
        dets_h = eng.alloc_host((max_walkers, W), np.int64)
        gd_h = eng.alloc_host((max_walkers,), np.float64)
        go_h = eng.alloc_host((max_walkers,), np.float64)
        variants = {}
        for tag, with_gdata in (("with_gdata", True), ("dets_only", False)):
            d, gd, go = eng.download_walkers()
            n = d.shape[0]
            dets_h[:n] = d; gd_h[:n] = gd; go_h[:n] = go
            h2d = d2h = 0
            att = 0.0
            ga, gb = (gd_h, go_h) if with_gdata else (None, None)
            for k in range(2):                                   # warm-up of the host path
                it += 1
                st, n = eng.iterate_host(dets_h, n, ga, gb, tau, sft, it)
            e_steps = max(3, min(args.steps, 10))
            barrier()
            w0 = time.perf_counter()
            for k in range(e_steps):
                it += 1
                h2d += n * (8 * W + (16 if with_gdata else 0))
                st, n = eng.iterate_host(dets_h, n, ga, gb, tau, sft, it)
                # written through to the host arrays: the sign word of every merged target and of every determinant that
                # died, a whole record (+ gdata) per new determinant, the flag word of every removed one; + the statistics
                changed = st[ST["NSPAWNED_MERGED"]] + st[ST["NODIED"]]
                d2h += 8 * changed + st[ST["NINSERTED"]] * (8 * W + (16 if with_gdata else 0)) + 8 * capi.ST_COUNT
                att += st[ST["NVALIDEXCITS"]] + st[ST["NINVALIDEXCITS"]]
            barrier()
            e_wall = allmax(time.perf_counter() - w0)
            variants[tag] = {"value": allsum(att) / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d / e_steps),
                             "d2h_bytes_per_step": int(d2h / e_steps), "ms_per_step": 1e3 * e_wall / e_steps, "steps": e_steps}
        best = max(variants, key=lambda k: variants[k]["value"])
        e2e = dict(variants[best])
        e2e.update({"variant": best, "variants": variants,
                    "api": "neci_gpu_iterate_host: CurrentDets%s in page-locked host memory, uploaded every iteration; the "
                           "iteration's changes are written through to the host arrays by the kernels (d2h bytes estimated "
                           "from the iteration's counters)" % (" + global_determinant_data" if best == "with_gdata" else ""),
                    "resident_api": {"value": attempts / wall, "unit": UNIT, "h2d_bytes_per_step": 24,
                                     "d2h_bytes_per_step": 8 * capi.ST_COUNT + 128,
                                     "note": "neci_gpu_iterate: list stays in HBM, host passes tau/shift/iter and reads the "
                                             "statistics vector every iteration (wall clock, same K steps as `value`)"}})

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and primary:
        nd = int(args.cpu_sample_walkers / 2.0)
        r = cpu_run(args.workload, nd, 6, 3, cores, core_size=args.core_size, n_trial=args.trial)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "6 iterations of a %.3g-walker (%d determinants) list of the same system and distribution, "
                         "ranks = host threads" % (r["walkers"], nd), "ms_per_step": r["ms_per_step"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "walkers_per_gpu": args.walkers,
                       "start_list": "device (neci_gpu_synthetic_list)" if device_list else "host (numpy, uploaded)",
                       **({"particle_selection": args.particle_selection} if args.particle_selection != "UNIF-UNIF" else {}),
                       "walkers_total_end": walkers_end, "determinants_total_end": dets_end, "tau": tau, "shift": sft,
                       "initiator": True, "attempts_per_step": attempts / args.steps,
                       "spawned_per_step": spawned / args.steps, "partition": "DetermineDetNode hash" if world > 1 else "single rank",
                       "exchange": ("routing + pushes over NVLink peer memory inside the spawning kernels" if args.exchange == "p2p" else "NCCL send/recv") if world > 1 else "none",
                       "l2": "inputs larger than L2 (walker list %.0f MB per GPU > 126 MB)" % (dets_end / world * (8 * system.nw + 28) / 1e6),
                       "wall_ms_per_step": 1e3 * wall / args.steps, **({"semi_stochastic": core_info} if semi else {}),
                       **({"load_balance": {"blocks_per_rank": 100, **(lb_moves or {})}} if args.load_balance else {}),
                       **({"reference_fraction": ref_fraction} if ref_fraction > 0.0 else {})},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "selfcheck": selfcheck,
        }
    eng.close()
    return line if rank == 0 else None


if __name__ == "__main__":
    sys.exit(main())
