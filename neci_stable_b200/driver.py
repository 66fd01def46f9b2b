"""Host driver mirroring the outer loop of FciMCPar (src/FciMCPar.F90:394-854):
one engine call per iteration where the reference calls PerformFCIMCycPar
(:507), statistics reduced over ranks every StepsSft iterations as in
communicate_estimates (src/fcimc_iter_utilities.F90:390-791), shift updated as
in update_shift (:883-1260), and the FCIMCStats-style history kept for the
blocking analysis (src/ErrorAnalysis.F90).

The per-iteration work is entirely inside the engine; nothing here touches
walker data.
"""
import math

import numpy as np

from . import capi, host
from .capi import ST


def diag_energy(system, orbs):
    """<D|H|D> (sltcnd_0_base src/sltcnd.fpp:585-622 / lattice diagonal elements)."""
    t = system.tables
    orbs = list(orbs)
    if system.kind == capi.SYS_HUBBARD_RS:
        s = set(orbs)
        nd = sum(1 for o in orbs if o % 2 == 1 and (o + 1) in s)
        return t["uhub"] * nd
    if system.kind == capi.SYS_HUBBARD_K:
        na = sum(1 for o in orbs if o % 2 == 0)
        nb = len(orbs) - na
        return sum(t["eps_k"][(o - 1) // 2] for o in orbs) + t["u_over_n"] * na * nb + system.ecore
    nbas = system.nbasis
    umat, tmat = t["umat"], t["tmat"]

    def tri(a, b):
        return a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a

    def um(i, j, k, l):
        return umat[tri(tri(i, k), tri(j, l)) - 1]

    e = sum(tmat[(o - 1) + nbas * (o - 1)] for o in orbs)
    for a in range(len(orbs)):
        for b in range(a + 1, len(orbs)):
            i, j = (orbs[a] + 1) // 2, (orbs[b] + 1) // 2
            e += um(i, j, i, j)
            if (orbs[a] - orbs[b]) % 2 == 0:
                e -= um(i, j, j, i)
    return e + system.ecore


def _world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def reduce_stats(st):
    """MPISumAll / MPIAllReduce(MAX) of the per-rank accumulators (fcimc_iter_utilities.F90:616,733,783)."""
    dist = _world()
    if dist is None or dist.get_world_size() == 1:
        return st
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    s = torch.tensor(st, dtype=torch.float64, device=dev)
    m = s.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out = s.cpu().numpy()
    mx = m.cpu().numpy()
    for name in capi.ST_MAX_REDUCED + ("TIME_SPAWN_MS", "TIME_COMM_MS", "TIME_ANNIHIL_MS", "TIME_DETERM_MS"):
        out[ST[name]] = mx[ST[name]]
    out[ST["ERR_FLAGS"]] = mx[ST["ERR_FLAGS"]]
    return out


def dist_reduce_or(flags):
    """MPIAllLORLogical over the ranks of the initialised process group (identity on one rank)."""
    dist = _world()
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(flags, dtype=bool)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(np.asarray(flags, dtype=np.int32), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy().astype(bool)


def dist_reduce_max(values):
    """MPIAllReduce(MPI_MAX) over the ranks of the initialised process group (identity on one rank)."""
    dist = _world()
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(values, dtype=np.float64)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(np.asarray(values, dtype=np.float64), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def plan_load_balance(block_parts_all, mapping, nranks):
    """The greedy balancer of adjust_load_balance (src/load_balancer.fpp:239-304): repeatedly move the smallest
    non-empty block of the fullest rank to the emptiest rank while that brings both closer to the average.
    block_parts_all: summed block populations (sum(ceiling(abs(sgn)))); mapping: block -> rank (0-based).
    Returns (new_mapping, movelist[(block, from, to)])."""
    parts = np.asarray(block_parts_all, dtype=np.int64)
    mapping = np.array(mapping, dtype=np.int32).copy()
    moves = []
    while True:
        proc_parts = np.bincount(mapping, weights=parts, minlength=nranks).astype(np.int64)
        avg = proc_parts.sum() / float(nranks)
        min_proc, max_proc = int(np.argmin(proc_parts)), int(np.argmax(proc_parts))      # first occurrence, as min/maxloc
        min_parts, max_parts = int(proc_parts[min_proc]), int(proc_parts[max_proc])
        smallest_block, smallest_size = -1, -1
        for b in np.nonzero(mapping == max_proc)[0]:
            if parts[b] > 0 and (parts[b] < smallest_size or smallest_size == -1):
                smallest_block, smallest_size = int(b), int(parts[b])
        unbalanced = (smallest_block != -1 and abs(min_parts + smallest_size - avg) < abs(min_parts - avg)
                      and abs(max_parts - smallest_size - avg) < abs(max_parts - avg))
        if not unbalanced:
            break
        moves.append((smallest_block, max_proc, min_proc))
        mapping[smallest_block] = min_proc
    return mapping, moves


class LoadBalanceTrigger:
    """need_load_balancing (src/load_balancer.fpp:836-855) with the imbalance measure of the main loop
    (src/FciMCPar.F90:564-578, 842-849): over a measuring cycle of 100 iterations, lt_imb = sum_iter (max_rank t - mean_rank t)
    / sum_iter sum_rank t, i.e. the fraction of loop time lost to imbalance.  Balance when it exceeds
    max(0.1, 2 x the value logged right after the previous balancing step); never in two consecutive cycles."""

    def __init__(self):
        self.last_imb = 0.0
        self.last_t_lb = False

    @staticmethod
    def imbalance(loop_times):
        """loop_times: array [n_iter, n_ranks] of per-iteration loop times."""
        t = np.asarray(loop_times, dtype=np.float64)
        tot = t.sum()
        return float((t.max(axis=1) - t.mean(axis=1)).sum() / tot) if tot > 0 else 0.0

    def need(self, lt_imb):
        if self.last_t_lb:
            self.last_imb = lt_imb
            self.last_t_lb = False
            return False
        t_lb = lt_imb > max(0.1, 2.0 * self.last_imb)
        self.last_t_lb = t_lb
        return t_lb


class TauSearch:
    """The conventional tau search of the host (update_tau, src/tau/tau_search_conventional.F90:274-499), fed by the
    per-iteration maxima and counts the engine returns (log_spawn_magnitude, :138-260).  Keeps the running maxima
    gamma_* and the enough_* switches (cnt_threshold = 50), proposes tau = MaxWalkerBloom * p_class / gamma_class and,
    once every class has been seen often enough, the biases pParallel = gamma_par / (gamma_opp + gamma_par) and
    pSingles = gamma_sing pParallel / (gamma_par + gamma_sing pParallel); tau never exceeds 1 / max_death_cpt
    (the largest K_ii - S seen at a death attempt, :419-430)."""
    CNT_THRESHOLD = 50

    def __init__(self, tau, p_singles, p_doubles, p_parallel, consider_par_bias, max_walker_bloom=1.0,
                 min_tau=1e-7, max_tau=1.0, t_hub=False, t_k_space_hubbard=False, reduce_or=None, reduce_max=None):
        """t_hub / t_k_space_hubbard: the reference's tHub and t_k_space_hubbard (lattice models: tau is re-assigned at
        every update).  reduce_or / reduce_max: the MPIAllLORLogical / MPIAllReduce(MAX) of update_tau for runs on several
        ranks (callables on a bool / float array; None = one rank)."""
        self.tau, self.p_singles, self.p_doubles, self.p_parallel = tau, p_singles, p_doubles, p_parallel
        self.par_bias, self.bloom, self.min_tau, self.max_tau = consider_par_bias, max_walker_bloom, min_tau, max_tau
        self.t_hub, self.t_k_hub = bool(t_hub), bool(t_k_space_hubbard)
        self.reduce_or, self.reduce_max = reduce_or, reduce_max
        self.gamma = np.zeros(4)                 # sing, doub, par, opp
        self.cnt = np.zeros(4)
        self.max_death_cpt = 0.0

    def log(self, stats):
        """Accumulate THIS RANK's statistics vector of one iteration: the reference keeps gamma_*, cnt_* and the
        enough_* switches per rank (log_spawn_magnitude) and reduces them inside update_tau (maxima with MPI_MAX, the
        switches with a logical OR: tau_search_conventional.F90:295-312)."""
        g0 = ST["TAU_GAMMA_SING"]; c0 = ST["TAU_CNT_SING"]
        self.gamma = np.maximum(self.gamma, stats[g0:g0 + 4])
        self.cnt += stats[c0:c0 + 4]
        self.max_death_cpt = max(self.max_death_cpt, stats[ST["TAU_MAX_DEATH_CPT"]])

    @property
    def enough(self):
        e = self.cnt > self.CNT_THRESHOLD
        if self.reduce_or is not None:
            e = np.asarray(self.reduce_or(e), dtype=bool)
        sing, doub, par, opp = bool(e[0]), bool(e[1]), bool(e[2]), bool(e[3])
        if self.par_bias:
            doub = par and opp
        return sing, doub, par, opp

    def update(self):
        """update_tau: returns (tau, p_singles, p_doubles, p_parallel) to use from now on."""
        eps = 1e-13
        if self.reduce_max is not None:
            red = np.asarray(self.reduce_max(np.concatenate([self.gamma, [self.max_death_cpt]])), dtype=np.float64)
            self.gamma = red[:4].copy(); self.max_death_cpt = float(red[4])
        g_sing, g_doub, g_par, g_opp = self.gamma
        e_sing, e_doub, e_par, e_opp = self.enough
        ps_new, pp_new = self.p_singles, self.p_parallel
        if self.par_bias:
            if e_sing and e_doub:
                pp_new = g_par / (g_opp + g_par)
                ps_new = g_sing * pp_new / (g_par + g_sing * pp_new)
                tau_new = ps_new * self.bloom / g_sing
            elif g_sing > eps and g_par > eps and g_opp > eps:
                tau_new = self.bloom * min(self.p_singles / g_sing, self.p_doubles * self.p_parallel / g_par,
                                           self.p_doubles * (1.0 - self.p_parallel) / g_opp)
            else:
                tau_new = self.tau
            if e_opp and e_par:
                self.p_parallel = pp_new
        else:
            gsum = g_sing + g_doub
            if e_sing and e_doub:
                ps_new = max(g_sing / gsum, 1e-8)
                tau_new = self.bloom / gsum
            elif abs(g_doub) > eps and abs(g_sing) > eps:
                tau_new = self.bloom * min(self.p_singles / g_sing, self.p_doubles / g_doub)
            elif abs(g_doub) > eps:
                tau_new = self.bloom * self.p_doubles / g_doub
            elif abs(g_sing) > eps:
                tau_new = self.bloom * self.p_singles / g_sing
            else:
                tau_new = self.tau
        if abs(self.max_death_cpt) > eps:
            tau_death = 1.0 / self.max_death_cpt
            if tau_death < tau_new:
                self.min_tau = min(self.min_tau, tau_death)
                tau_new = tau_death
        tau_new = min(max(tau_new, self.min_tau), self.max_tau)
        # the reference's condition as Fortran parses it (.and. binds tighter than .or., :445-450; UEG and the
        # transcorrelated branches are outside this engine): enough_sing alone re-assigns tau, and so does tHub
        if (tau_new < self.tau or (e_sing and e_doub) or self.t_hub or e_sing or (self.t_k_hub and e_doub)):
            self.tau = tau_new * 0.99999
        if e_sing and e_doub and 1e-5 < ps_new < 1.0 - 1e-5:
            self.p_singles = ps_new
            self.p_doubles = 1.0 - ps_new
        return self.tau, self.p_singles, self.p_doubles, self.p_parallel


class FciMC:
    """Drives one engine (rank) through the FCIQMC iteration loop."""

    def __init__(self, system, engine, hii, tau, init_walkers, steps_sft=10, sft_damp=0.1, diag_sft=0.0,
                 nranks=1, jump_shift=False):
        self.system, self.engine, self.hii = system, engine, hii
        self.tau, self.init_walkers = tau, init_walkers
        self.steps_sft, self.sft_damp = steps_sft, sft_damp
        self.diag_sft = diag_sft
        self.nranks = nranks
        self.jump_shift = jump_shift           # tJumpShift (src/fcimc_iter_utilities.F90:1044-1054)
        self.iter = 0
        self.single_part_phase = True          # tSinglePartPhase
        self.tot_parts = 0.0                   # AllTotParts of the previous iteration
        self.old_av_walkers = 0.0              # OldAllAvWalkersCyc
        self.sum_walkers_cyc = 0.0             # AllSumWalkersCyc
        self.cyc = np.zeros(capi.ST_COUNT)     # accumulators over the update cycle
        self.history = []                      # one row per update cycle (FCIMCStats)
        self.attempts = 0.0

    def seed_reference(self, n_walkers, rank_of_ref=0, my_rank=0):
        """InitFCIMC_HF: start from n walkers on the reference determinant."""
        s = self.system
        if my_rank == rank_of_ref:
            rec = host.record(s, s.ref_orbs, float(n_walkers), (1 << capi.FLAG_INITIATOR))
            self.engine.upload_walkers(rec.reshape(1, -1))
            self.tot_parts = float(n_walkers)
        else:
            self.engine.upload_walkers(np.zeros((0, s.W), dtype=np.int64))
        self.old_av_walkers = float(n_walkers)

    def step(self):
        """One iteration == one PerformFCIMCycPar."""
        self.iter += 1
        self.sum_walkers_cyc += self.tot_parts          # end_iter_stats: SumWalkersCyc += TotParts
        st = self.engine.iterate(self.tau, self.diag_sft, self.iter)
        self.tot_parts = st[ST["TOTPARTS"]]
        self.cyc += st
        for name in capi.ST_MAX_REDUCED:
            self.cyc[ST[name]] = max(self.cyc[ST[name]] - st[ST[name]], st[ST[name]])
        self.attempts += st[ST["NVALIDEXCITS"]] + st[ST["NINVALIDEXCITS"]]
        if self.iter % self.steps_sft == 0:
            self._update_cycle(st)
        return st

    def _update_cycle(self, last):
        # communicate_estimates
        loc = self.cyc.copy()
        loc[ST["TOTPARTS"]] = last[ST["TOTPARTS"]]
        loc[ST["TOTWALKERS"]] = last[ST["TOTWALKERS"]] - last[ST["HOLESINLIST"]]
        extra = np.array([self.sum_walkers_cyc])
        allst = reduce_stats(np.concatenate([loc, extra]))
        all_sum_walkers_cyc = allst[-1]
        allst = allst[:-1]
        all_tot_parts = allst[ST["TOTPARTS"]]
        av_walkers = all_sum_walkers_cyc / self.steps_sft
        # update_shift
        hf_now = allst[ST["HFCYC"]]
        defer_update = False
        if self.single_part_phase and all_tot_parts > self.init_walkers * self.nranks:
            self.single_part_phase = False
            if self.jump_shift and abs(hf_now) > 1e-13:
                # jump the shift to the value the projected energy predicts and defer the update by one cycle
                self.diag_sft = allst[ST["ENUMCYC"]] / hf_now
                defer_update = True
        if not self.single_part_phase and not defer_update and self.old_av_walkers > 0 and av_walkers > 0:
            self.diag_sft = host.update_shift(self.diag_sft, self.sft_damp, self.tau, self.steps_sft,
                                              av_walkers, self.old_av_walkers)
        hf = allst[ST["HFCYC"]]
        proje = allst[ST["ENUMCYC"]] / hf if abs(hf) > 1e-13 else float("nan")
        self.history.append(dict(iter=self.iter, shift=self.diag_sft, tot_parts=all_tot_parts,
                                 n_dets=allst[ST["TOTWALKERS"]], enum_cyc=allst[ST["ENUMCYC"]], hf_cyc=hf,
                                 proje_corr=proje, proje=proje + self.hii, varying=not self.single_part_phase,
                                 noathf=hf / self.steps_sft, trial_num=allst[ST["TRIAL_NUMERATOR"]],
                                 trial_den=allst[ST["TRIAL_DENOM"]]))
        self.old_av_walkers = av_walkers
        self.sum_walkers_cyc = 0.0
        self.cyc[:] = 0.0

    def run(self, n_iter):
        for _ in range(n_iter):
            self.step()
        return self.history


# ---------------------------------------------------------------------------------------
def blocking(x):
    """Flyvbjerg-Petersen reblocking (src/ErrorAnalysis.F90:374-470): returns
    (mean, error) with the error taken at the first plateau of the blocked
    standard error (largest value that is still well determined)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 2:
        return (float(x.mean()) if n else float("nan")), float("inf")
    mean = x.mean()
    errs = []
    y = x.copy()
    while y.size >= 8:
        m = y.size
        se = y.std(ddof=1) / math.sqrt(m)
        ee = se / math.sqrt(2.0 * (m - 1))
        errs.append((se, ee))
        if m % 2:
            y = y[:-1]
        y = 0.5 * (y[0::2] + y[1::2])
    if not errs:
        return float(mean), float(x.std(ddof=1) / math.sqrt(n))
    best = errs[0][0]
    for se, ee in errs:
        if se > best and ee < 0.5 * se:
            best = se
    return float(mean), float(best)


def ratio_estimate(num, den):
    """Projected-energy estimate <num>/<den> with blocking errors and the
    numerator/denominator covariance (src/ErrorAnalysis.F90:1116-1180)."""
    num, den = np.asarray(num, float), np.asarray(den, float)
    mn, en = blocking(num)
    md, ed = blocking(den)
    if abs(md) < 1e-300:
        return float("nan"), float("inf")
    cov = np.cov(num, den)[0, 1] / max(len(num), 1)
    r = mn / md
    var = (en / mn) ** 2 + (ed / md) ** 2 - 2.0 * cov / (mn * md) if mn != 0 else float("inf")
    return r, abs(r) * math.sqrt(max(var, 0.0))
