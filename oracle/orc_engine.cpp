// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see orc_common.hpp).
//
// CPU restatement of one FCIQMC iteration, PerformFCIMCycPar
// (src/FciMCPar.F90:1177-1920), for the `neci` build (lenof_sign = 1,
// inum_runs = 1), exposing the same C entry points as include/neci_gpu.h with
// the prefix orc_ instead of neci_gpu_.
#include "orc_system.hpp"
#include <thread>
#include <unordered_set>
#include <cstdio>

using namespace orc;

namespace {

struct DetKey {
    uint64_t w[2];
    bool operator==(const DetKey &o) const { return w[0] == o.w[0] && w[1] == o.w[1]; }
};
struct DetKeyHash {
    size_t operator()(const DetKey &k) const { return (size_t)mix64(k.w[0] ^ mix64(k.w[1] + 0x9E3779B97F4A7C15ull)); }
};

}  // namespace

struct orc_engine {
    neci_gpu_config cfg;
    std::vector<int32_t> random_orb_index, random_hash2, lb_mapping;
    std::vector<uint64_t> ilut_ref;
    System S;
    int W = 3, nwords = 1;
    // main list, NECI AoS layout
    std::vector<int64_t> dets;          // W * max_walkers
    std::vector<double> diagH, offdiagH;
    int64_t TotWalkers = 0;
    std::unordered_map<DetKey, int64_t, DetKeyHash> hash;   // HashIndex (src/hash.F90): det -> slot
    std::vector<int64_t> FreeSlot;
    int64_t iStartFreeSlot = 0, iEndFreeSlot = 0;           // 0-based [start, end)
    int64_t HolesInList = 0;
    // spawned lists per destination rank (SpawnedParts segments)
    std::vector<std::vector<int64_t>> spawned;
    // semi-stochastic
    int64_t n_core_local = 0, n_core_total = 0;
    std::vector<int64_t> row_ptr; std::vector<int32_t> col; std::vector<double> val;
    std::vector<int32_t> core_sizes, core_displs;
    std::vector<int64_t> indices_of_determ_states;
    struct KH { size_t operator()(const std::array<uint64_t, 2> &k) const { return (size_t)mix64(k[0] ^ mix64(k[1] + 0x9E3779B97F4A7C15ull)); } };
    std::unordered_set<std::array<uint64_t, 2>, KH> core_set;      // core_space hash (is_core_state)
    std::vector<double> partial_determ_vecs, full_determ_vecs;
    std::vector<double> core_ham_diag;                             // fast_determ_hamil.F90:1494-1507
    // trial wavefunction: trial_ht / con_ht (src/searching.F90:182-223) and current_trial_amps
    bool t_trial = false;
    std::unordered_map<std::array<uint64_t, 2>, double, KH> trial_ht, con_ht;
    std::vector<double> current_trial_amps;
    double stats[NECI_ST_COUNT];
    std::string err;

    const uint64_t *orb(int64_t slot) const { return (const uint64_t *)&dets[(size_t)slot * W]; }
    double sign(int64_t slot) const { return sign_to_double(dets[(size_t)slot * W + nwords]); }
    void set_sign(int64_t slot, double s) { dets[(size_t)slot * W + nwords] = double_to_sign(s); }
    int64_t &flags(int64_t slot) { return dets[(size_t)slot * W + nwords + 1]; }
    DetKey key(const uint64_t *o) const { DetKey k; k.w[0] = o[0]; k.w[1] = (nwords > 1) ? o[1] : 0; return k; }
    bool test_flag(int64_t slot, int f) { return (flags(slot) >> f) & 1; }
    void set_flag(int64_t slot, int f, bool v) { if (v) flags(slot) |= (1ll << f); else flags(slot) &= ~(1ll << f); }
};

namespace {

// get_det_block / DetermineDetNode, src/load_balance_calcnodes.F90:25-117
// (hash_iter = 0, tUniqueHFNode = .false.).  int64 wrap-around as in Fortran.
int det_block(const orc_engine &e, const int *nI) {
    uint64_t acc = 0;
    for (int i = 1; i <= e.cfg.nel; ++i) {
        const int o = (nI[i - 1] - 1) % e.cfg.nbasis + 1;
        acc = 1099511628211ull * acc + (uint64_t)(int64_t)(e.random_orb_index[o - 1] * i);
    }
    const int64_t sacc = (int64_t)acc;
    int64_t m = sacc % (int64_t)e.cfg.balance_blocks;     // C remainder == Fortran mod (sign of dividend)
    if (m < 0) m = -m;
    return (int)m + 1;
}
int det_node(const orc_engine &e, const int *nI) { return e.lb_mapping[det_block(e, nI) - 1]; }

// FindWalkerHash, src/hash.F90:23-37
int find_walker_hash(const orc_engine &e, const int *nI, int len) {
    uint64_t h = 0;
    for (int i = 1; i <= e.cfg.nel; ++i) h = 1099511628211ull * h + (uint64_t)(int64_t)(e.random_hash2[nI[i - 1] - 1] * i);
    int64_t m = (int64_t)h % (int64_t)len;
    if (m < 0) m = -m;
    return (int)m + 1;
}

// stochastic_round, src/lib/util_mod.fpp:182-204
inline int stochastic_round(double r, Stream &rng) {
    int i = (int)r;
    const double res = r - (double)i;
    if (std::fabs(res) >= 1.0e-12) {
        if (std::fabs(res) > rng.draw()) i += (int)std::lround(dsign(1.0, r));
    }
    return i;
}

void zero_stats(orc_engine &e) { for (int i = 0; i < NECI_ST_COUNT; ++i) e.stats[i] = 0.0; }

// RemoveHashDet, src/load_balancer.fpp:631-644
void RemoveHashDet(orc_engine &e, int64_t slot) {
    e.hash.erase(e.key(e.orb(slot)));
    e.FreeSlot[e.iEndFreeSlot++] = slot;
    e.set_flag(slot, NECI_FLAG_REMOVED, true);
}

// ---------------------------------------------------------------------------
// Loop over determinants: src/FciMCPar.F90:1294-1758
// ---------------------------------------------------------------------------
void spawn_phase(orc_engine &e, double tau, double DiagSft, int64_t iter) {
    const neci_gpu_config &c = e.cfg;
    const System &S = e.S;
    zero_stats(e);
    for (auto &v : e.spawned) v.clear();
    // ValidSpawnedList = InitialSpawnedSlots; reset FreeSlot   (:1237-1240)
    e.iStartFreeSlot = 0; e.iEndFreeSlot = 0;
    const int64_t seg_cap = c.max_spawned / c.nranks;
    int determ_index = 0;
    int nI[128];

    for (int64_t j = 0; j < e.TotWalkers; ++j) {
        const uint64_t *ilut = e.orb(j);
        const bool tCoreDet = e.test_flag(j, NECI_FLAG_DETERMINISTIC);       // check_determ_flag :1318
        const double SignCurr = e.sign(j);
        decode(ilut, c.nbasis, nI);
        const int walkExcitLevel = excit_level_hphf(S, e.ilut_ref.data(), ilut);     // :1354 (t_hphf_ic = .true.)

        if (c.t_semi_stochastic && tCoreDet) {                               // :1387-1411
            e.indices_of_determ_states[determ_index] = j;
            e.partial_determ_vecs[determ_index] = SignCurr;
            ++determ_index;
        }
        if (unocc(SignCurr)) {                                               // IsUnoccDet :1415-1427
            if (tCoreDet) continue;
            e.FreeSlot[e.iEndFreeSlot++] = j;
            continue;
        }
        const double HDiagCurr = e.diagH[j];                                 // :1437
        const double HOffDiagCurr = e.offdiagH[j];

        // CalcParentFlag -> TestInitiator_explicit, src/fcimc_helper.F90:1036-1243
        if (c.t_trunc_initiator) {
            bool parent_init = e.test_flag(j, NECI_FLAG_INITIATOR);
            const bool popInit = std::fabs(SignCurr) > c.initiator_walk_no;  // initiator_criterium :1274
            bool initiator = parent_init;
            if (!initiator) {
                if (popInit) { initiator = true; e.stats[NECI_ST_NOADDEDINITIATORS] += 1; }
            } else {
                bool staticInit = (walkExcitLevel == 0);                     // DetBitEQ(ilut, ilutRef) :1216
                if (!staticInit && !(tCoreDet && c.t_core_inits) && !popInit) {
                    initiator = false; e.stats[NECI_ST_NOADDEDINITIATORS] -= 1;
                }
            }
            if (initiator) { e.stats[NECI_ST_NOINITDETS] += 1; e.stats[NECI_ST_NOINITWALK] += std::fabs(SignCurr); }
            else { e.stats[NECI_ST_NONONINITDETS] += 1; e.stats[NECI_ST_NONONINITWALK] += std::fabs(SignCurr); }
            e.set_flag(j, NECI_FLAG_INITIATOR, initiator);
        }

        // SumEContrib, src/fcimc_helper.F90:518-802
        if (walkExcitLevel == 0) e.stats[NECI_ST_HFCYC] += SignCurr;
        if (walkExcitLevel == 2) e.stats[NECI_ST_NOATDOUBS] += std::fabs(SignCurr);
        {
            const double dE = HOffDiagCurr * SignCurr;
            e.stats[NECI_ST_ENUMCYC] += dE;
            e.stats[NECI_ST_ENUMCYCABS] += std::fabs(dE);
            if (e.test_flag(j, NECI_FLAG_INITIATOR)) e.stats[NECI_ST_INITSENUMCYC] += dE;
        }

        if (e.t_trial) {                                                     // :586-648 (ntrial_excits = 1)
            const double amp = e.current_trial_amps[j];
            if (e.test_flag(j, NECI_FLAG_TRIAL)) {
                e.stats[NECI_ST_TRIAL_DENOM] += amp * SignCurr;
                if (e.test_flag(j, NECI_FLAG_INITIATOR)) e.stats[NECI_ST_INIT_TRIAL_DENOM] += amp * SignCurr;
            } else if (e.test_flag(j, NECI_FLAG_CONNECTED)) {
                e.stats[NECI_ST_TRIAL_NUMERATOR] += amp * SignCurr;
                if (e.test_flag(j, NECI_FLAG_INITIATOR)) e.stats[NECI_ST_INIT_TRIAL_NUMERATOR] += amp * SignCurr;
            }
        }

        const uint64_t h = det_hash64(ilut, e.nwords);
        // decide_num_to_spawn, src/fcimc_helper.F90:2160-2174
        int WalkersToSpawn;
        {
            const double x = SignCurr * c.av_mc_excits;
            WalkersToSpawn = std::abs((int)x);
            if (std::fabs(std::fabs(x) - (double)WalkersToSpawn) > 1.e-12) {
                const double prob_extra = std::fabs(x) - (double)WalkersToSpawn;
                Stream rng(c.seed, iter, h, 0, RNG_NSPAWN);
                if (prob_extra > rng.draw()) ++WalkersToSpawn;
            }
        }
        const bool parent_is_init = e.test_flag(j, NECI_FLAG_INITIATOR);

        for (int p = 0; p < WalkersToSpawn; ++p) {                           // loop_over_walkers :1622
            Stream rng(c.seed, iter, h, (uint32_t)p, RNG_ATTEMPT);
            Excitation E;
            double HElGen = 0.0;
            if (S.t_hphf) { E = Excitation(); gen_hphf_excit(S, nI, ilut, rng, E, HElGen); }   // fcimc_initialisation.fpp:2162-2165
            else generate_excitation(S, nI, ilut, rng, E);
            if (E.err) e.stats[NECI_ST_ERR_FLAGS] = (double)((int)e.stats[NECI_ST_ERR_FLAGS] | 16);
            if (!E.valid) { e.stats[NECI_ST_NINVALIDEXCITS] += 1; continue; }
            e.stats[NECI_ST_NVALIDEXCITS] += 1;
            int64_t child_flags = 0;
            if (c.t_semi_stochastic && tCoreDet) {                           // :1651-1670
                // is_core_state (src/semi_stoch_procs.F90:547-587): lookup in the replicated core space
                if (e.core_set.count({E.ilutJ[0], (e.nwords > 1) ? E.ilutJ[1] : 0ull})) continue;
                child_flags |= (1ll << NECI_FLAG_DETERM_PARENT);
            }
            // attempt_create_normal, src/fcimc_pointed_fns.F90:178-491
            const double prob = E.pgen * c.av_mc_excits;
            const double rh = S.t_hphf ? HElGen : get_spawn_helement(S, nI, E);          // hphf_spawn_sign
            const double walkerweight = dsign(1.0, SignCurr);
            if (c.t_tau_search) {
                // log_spawn_magnitude, src/tau/tau_search_conventional.F90:138-260 (per-iteration maxima and counts; the
                // host keeps the running maxima and the `enough_*` switches)
                double tp; int cls;
                if (E.ic == 1) { tp = prob / S.p_singles; cls = 0; }
                else {
                    tp = prob / S.p_doubles; cls = 1;
                    if (c.t_consider_par_bias) {
                        if (((E.ex[0] ^ E.ex[1]) & 1) == 0) { tp = tp / S.p_parallel; cls = 2; }
                        else { tp = tp / (1.0 - S.p_parallel); cls = 3; }
                    }
                }
                const double g = std::fabs(rh) / tp;
                if (cls >= 2 || g > 0.0) {
                    e.stats[NECI_ST_TAU_GAMMA_SING + cls] = std::max(e.stats[NECI_ST_TAU_GAMMA_SING + cls], g);
                    e.stats[NECI_ST_TAU_CNT_SING + cls] += 1;
                }
            }
            double nSpawn = -tau * rh * walkerweight / prob;
            e.stats[NECI_ST_MAX_CYC_SPAWN] = std::max(e.stats[NECI_ST_MAX_CYC_SPAWN], std::fabs(nSpawn));
            // the rounding number is the first of the attempt's own RNG_ATT_ROUND stream (the engine evaluates the
            // spawn in a later stage than it draws the excitation)
            Stream rng_round(c.seed, iter, h, (uint32_t)p, RNG_ATT_ROUND);
            if (c.t_all_real_coeff) {
                if (c.t_real_spawn_cutoff && std::fabs(nSpawn) < c.real_spawn_cutoff)
                    nSpawn = c.real_spawn_cutoff * stochastic_round(nSpawn / c.real_spawn_cutoff, rng_round);
            } else {
                nSpawn = (double)stochastic_round(nSpawn, rng_round);
            }
            const double child = nSpawn;
            if (near_zero(child)) continue;                                  // is_child_created :1706
            // new_child_stats_normal, :506-571
            e.stats[NECI_ST_NOBORN] += std::fabs(child);
            if (E.ic == 1) e.stats[NECI_ST_SPAWNFROMSING] += std::fabs(child);
            if (std::fabs(child) > c.initiator_walk_no) {
                const int bc = (E.ic == 1) ? NECI_ST_BLOOM_COUNT_1 : NECI_ST_BLOOM_COUNT_2;
                const int bs = (E.ic == 1) ? NECI_ST_BLOOM_SIZE_1 : NECI_ST_BLOOM_SIZE_2;
                e.stats[bc] += 1; e.stats[bs] = std::max(e.stats[bs], std::fabs(child));
            }
            // create_particle, src/fcimc_helper.F90:152-308
            const int proc = det_node(e, E.nJ);
            if ((int64_t)(e.spawned[proc].size() / e.W) >= seg_cap) {
                e.stats[NECI_ST_ERR_FLAGS] = (double)((int)e.stats[NECI_ST_ERR_FLAGS] | 1);
                continue;
            }
            if (c.t_trunc_initiator && parent_is_init) child_flags |= (1ll << NECI_FLAG_INITIATOR);
            for (int w = 0; w < e.nwords; ++w) e.spawned[proc].push_back((int64_t)E.ilutJ[w]);
            e.spawned[proc].push_back(double_to_sign(child));
            e.spawned[proc].push_back(child_flags);
            e.stats[NECI_ST_ACCEPTANCES] += std::fabs(child);
        }

        // walker_death, src/fcimc_helper.F90:2279-2407.  tDeathBeforeComms: here, with t_core_die_ = .false.
        // (FciMCPar.F90:1752-1756).  Otherwise perform_death_all_walkers (fcimc_helper.F90:2253-2277) runs it after
        // the loop with t_core_die = .true.; death of determinant j reads and writes nothing but j's own sign, which
        // the loop does not touch afterwards, so running it here gives the same list.
        {
            const bool t_core_die = !c.t_death_before_comms;
            double iDie;
            // attempt_die_normal, src/fcimc_pointed_fns.F90:573-705
            const double fac = tau * (HDiagCurr - DiagSft);
            if (fac > 2.0) e.stats[NECI_ST_ERR_FLAGS] = (double)((int)e.stats[NECI_ST_ERR_FLAGS] | 4);
            // log_death_magnitude(Kii - shift), src/fcimc_pointed_fns.F90:640, src/tau/tau_main.F90:198-207
            if (c.t_tau_search)                          // attempt_die runs (and logs) for core determinants too
                e.stats[NECI_ST_TAU_MAX_DEATH_CPT] = std::max(e.stats[NECI_ST_TAU_MAX_DEATH_CPT], HDiagCurr - DiagSft);
            if (c.t_all_real_coeff) iDie = fac * std::fabs(SignCurr);
            else {
                double rat = fac * std::fabs(SignCurr);
                iDie = (double)(int64_t)rat;
                rat = rat - iDie;
                Stream rng(c.seed, iter, h, 0, RNG_DEATH);
                const double r = rng.draw();
                if (std::fabs(rat) > r) iDie += (double)std::lround(dsign(1.0, rat));
            }
            if (tCoreDet && !t_core_die) iDie = 0.0;
            e.stats[NECI_ST_NODIED] += std::min(iDie, std::fabs(SignCurr));
            e.stats[NECI_ST_NOBORN] += std::max(iDie - std::fabs(SignCurr), 0.0);
            double CopySign = SignCurr - (iDie * dsign(1.0, SignCurr));
            if (c.t_trunc_initiator && std::fabs(CopySign) > 1.0e-12) {
                if ((CopySign > 0.0) != (SignCurr > 0.0)) {
                    e.stats[NECI_ST_NOABORTED] += std::fabs(CopySign);
                    if (e.test_flag(j, NECI_FLAG_INITIATOR)) e.stats[NECI_ST_NOADDEDINITIATORS] -= 1;
                    CopySign = 0.0;
                }
            }
            if (std::fabs(CopySign) > 1.0e-12 || tCoreDet) e.set_sign(j, CopySign);
            else {
                if (c.t_trunc_initiator && e.test_flag(j, NECI_FLAG_INITIATOR)) e.stats[NECI_ST_NOADDEDINITIATORS] -= 1;
                RemoveHashDet(e, j);
                e.set_sign(j, 0.0);
            }
        }
    }
    int64_t ns = 0;
    for (auto &v : e.spawned) ns += (int64_t)(v.size() / e.W);
    e.stats[NECI_ST_NSPAWNED_SENT] = (double)ns;
}

// determ_projection, src/semi_stoch_procs.F90:105-241 (after the host-side gather)
void determ_projection(orc_engine &e, double tau, double DiagSft) {
    if (e.n_core_total == 0) return;
    const int64_t displ = e.core_displs[e.cfg.rank];
    for (int64_t i = 0; i < e.n_core_local; ++i) {
        double acc = 0.0;
        for (int64_t k = e.row_ptr[i]; k < e.row_ptr[i + 1]; ++k) acc = acc - e.val[k] * e.full_determ_vecs[e.col[k]];
        // determ_projection adds the shift (:171-202); determ_projection_no_death (:285-374, used when death is not
        // before comms, FciMCPar.F90:1778-1782) adds the diagonal element back instead: death does both later
        if (e.cfg.t_death_before_comms) acc = acc + DiagSft * e.full_determ_vecs[i + displ];
        else acc = acc + e.core_ham_diag[i] * e.full_determ_vecs[i + displ];
        e.partial_determ_vecs[i] = acc * tau;
    }
}

// ilut_lt, src/DetBitOps.F90:431-473: signed word compare, word 0 first
struct SpawnLess {
    int nwords;
    bool operator()(const int64_t *a, const int64_t *b) const {
        for (int w = 0; w < nwords; ++w) { if (a[w] < b[w]) return true; if (a[w] > b[w]) return false; }
        return false;
    }
};

// Merged amplitude of a block of real-coefficient spawns.  The reference adds them in the order its (unstable)
// quicksort leaves them, so its last bits depend on the arrival order.  The engine fixes a canonical value that no
// order can change: every contribution is split exactly into a multiple of 2^-24 and a remainder resolved to 2^-64,
// both parts are summed as 64-bit integers, and the two sums are joined once.  For amplitudes of at least 2^-11 the
// split is exact, so the result is the exact sum rounded twice (within one ulp of the sequential sum).
static inline void fixed_split(double s, int64_t &hi, int64_t &lo) {
    const double x = s * 16777216.0;                   // 2^24, exact
    hi = (int64_t)x;                                   // truncation
    lo = (int64_t)((x - (double)hi) * 1099511627776.0);  // 2^40; x - trunc(x) is exact
}
static inline double fixed_join(int64_t hi, int64_t lo) {
    return ((double)hi + (double)lo * (1.0 / 1099511627776.0)) * (1.0 / 16777216.0);
}

// CompressSpawnedList + FindResidualParticle, src/Annihilation.F90:249-515,551-634
void compress_spawned(orc_engine &e, std::vector<int64_t> &sp) {
    const int W = e.W, nw = e.nwords;
    const int64_t n = (int64_t)(sp.size() / W);
    std::vector<const int64_t *> idx(n);
    for (int64_t i = 0; i < n; ++i) idx[i] = &sp[(size_t)i * W];
    std::stable_sort(idx.begin(), idx.end(), SpawnLess{nw});
    std::vector<int64_t> out; out.reserve(sp.size());
    int64_t b = 0;
    while (b < n) {
        int64_t c = b + 1;
        while (c < n && std::memcmp(idx[b], idx[c], 8 * nw) == 0) ++c;
        if (c - b == 1) {
            const double s = sign_to_double(idx[b][nw]);
            if (std::fabs(s) >= 1.e-12) out.insert(out.end(), idx[b], idx[b] + W);
        } else {
            int64_t cum_flags = 0; double cum_sgn = 0.0;
            int64_t fx_hi = 0, fx_lo = 0;
            for (int64_t i = b; i < c; ++i) {
                const double new_sgn = sign_to_double(idx[i][nw]);
                { int64_t h, l; fixed_split(new_sgn, h, l); fx_hi += h; fx_lo += l; }
                const bool new_init = (idx[i][nw + 1] >> NECI_FLAG_INITIATOR) & 1;
                if (e.cfg.t_trunc_initiator) {
                    if (e.cfg.t_init_coherent_rule) {
                        if ((std::fabs(cum_sgn) > 1.e-12 && std::fabs(new_sgn) > 1.e-12) || new_init)
                            cum_flags |= (1ll << NECI_FLAG_INITIATOR);
                    } else if (new_init) cum_flags |= (1ll << NECI_FLAG_INITIATOR);
                }
                if (cum_sgn * new_sgn < 0.0)
                    e.stats[NECI_ST_ANNIHILATED] += 2 * std::min(std::fabs(cum_sgn), std::fabs(new_sgn));
                cum_sgn = cum_sgn + new_sgn;
            }
            // integer amplitudes add exactly in any order; real ones take the order-independent sum
            if (e.cfg.t_all_real_coeff) cum_sgn = fixed_join(fx_hi, fx_lo);
            if (std::fabs(cum_sgn) > 1.e-12) {
                for (int w = 0; w < nw; ++w) out.push_back(idx[b][w]);
                out.push_back(double_to_sign(cum_sgn));
                out.push_back(cum_flags);
            }
        }
        b = c;
    }
    sp.swap(out);
}

// hash_search_trial (src/searching.F90:182-223) + the flag / amplitude update of load_balancer.fpp:596-611
void trial_search_and_flag(orc_engine &e, int64_t pos) {
    const std::array<uint64_t, 2> k = {e.orb(pos)[0], (e.nwords > 1) ? e.orb(pos)[1] : 0ull};
    bool tTrial = false, tCon = false; double amp = 0.0;
    auto it = e.trial_ht.find(k);
    if (it != e.trial_ht.end()) { tTrial = true; amp = it->second; }
    else { auto ic = e.con_ht.find(k); if (ic != e.con_ht.end()) { tCon = true; amp = ic->second; } }
    e.set_flag(pos, NECI_FLAG_TRIAL, tTrial);
    e.set_flag(pos, NECI_FLAG_CONNECTED, tCon);
    e.current_trial_amps[pos] = amp;
}

// AddNewHashDet, src/load_balancer.fpp:514-629
bool AddNewHashDet(orc_engine &e, const int64_t *rec, double sgn, double HDiag, double HOffDiag) {
    int64_t pos;
    if (e.iStartFreeSlot < e.iEndFreeSlot) pos = e.FreeSlot[e.iStartFreeSlot++];
    else {
        if (e.TotWalkers + 1 >= e.cfg.max_walkers) return false;
        pos = e.TotWalkers++;
    }
    for (int w = 0; w < e.nwords; ++w) e.dets[(size_t)pos * e.W + w] = rec[w];
    e.set_sign(pos, sgn);
    e.flags(pos) = rec[e.nwords + 1];
    e.diagH[pos] = HDiag - e.cfg.hii;
    e.offdiagH[pos] = HOffDiag;
    e.set_flag(pos, NECI_FLAG_REMOVED, false);
    e.hash[e.key((const uint64_t *)rec)] = pos;
    if (e.t_trial) trial_search_and_flag(e, pos);                            // :586-611
    return true;
}

// AnnihilateSpawnedParts, src/Annihilation.F90:965-1352 (+ deterministic_annihilation :930-963)
void annihilate_phase(orc_engine &e, std::vector<int64_t> &sp, int64_t iter) {
    const neci_gpu_config &c = e.cfg;
    const int W = e.W, nw = e.nwords;
    e.stats[NECI_ST_NSPAWNED_RECV] = (double)(sp.size() / W);
    compress_spawned(e, sp);
    const int64_t n = (int64_t)(sp.size() / W);
    e.stats[NECI_ST_NSPAWNED_MERGED] = (double)n;

    if (c.t_semi_stochastic) {
        for (int64_t i = 0; i < e.n_core_local; ++i) {
            const int64_t slot = e.indices_of_determ_states[i];
            const double cur = e.sign(slot), spn = e.partial_determ_vecs[i];
            e.set_sign(slot, spn + cur);
            // iter_data%nborn / nannihil only (not the NoBorn/Annihilated globals)
        }
    }

    for (int64_t i = 0; i < n; ++i) {
        int64_t *rec = &sp[(size_t)i * W];
        auto it = e.hash.find(e.key((const uint64_t *)rec));
        double SpawnedSign = sign_to_double(rec[nw]);
        const bool spawn_init = (rec[nw + 1] >> NECI_FLAG_INITIATOR) & 1;
        if (it != e.hash.end()) {
            const int64_t PartInd = it->second;
            const double CurrentSign = e.sign(PartInd);
            const double SignProd = CurrentSign * SpawnedSign;
            const bool tDetermState = e.test_flag(PartInd, NECI_FLAG_DETERMINISTIC);
            if (std::fabs(CurrentSign) >= 1.e-12 || tDetermState) {
                if (unocc(CurrentSign)) {     // is_run_unnocc: only reachable for core dets
                    if (c.t_trunc_initiator && !spawn_init && !tDetermState) {
                        e.stats[NECI_ST_NOABORTED] += std::fabs(SpawnedSign); SpawnedSign = 0.0;
                    }
                }
                if (SignProd < 0)
                    e.stats[NECI_ST_ANNIHILATED] += 2 * std::min(std::fabs(CurrentSign), std::fabs(SpawnedSign));
                e.set_sign(PartInd, SpawnedSign + CurrentSign);
                if (!tDetermState && unocc(e.sign(PartInd))) RemoveHashDet(e, PartInd);
            }
        } else {
            if (c.t_trunc_initiator && !spawn_init) {                        // test_abort_spawn :1462
                e.stats[NECI_ST_NOABORTED] += std::fabs(SpawnedSign);
                SpawnedSign = 0.0;
            }
            if (!unocc(SpawnedSign)) {
                // stochRoundSpawn :1354-1418 (scFVal = 1)
                const double thr = c.occupied_thresh;
                if (std::fabs(SpawnedSign) > 1.e-12 && std::fabs(SpawnedSign) < thr) {
                    const double pRemove = 1.0 - std::fabs(SpawnedSign) / thr;
                    Stream rng(c.seed, iter, det_hash64((const uint64_t *)rec, nw), 0, RNG_ROUND_SPAWN);
                    if (pRemove > rng.draw()) { e.stats[NECI_ST_NOREMOVED] += std::fabs(SpawnedSign); SpawnedSign = 0.0; }
                    else { e.stats[NECI_ST_NOBORN] += thr - std::fabs(SpawnedSign); SpawnedSign = dsign(thr, SpawnedSign); }
                }
                if (!unocc(SpawnedSign)) {
                    const double diagH = get_diagonal_matel(e.S, (const uint64_t *)rec);
                    const double offdiagH = get_off_diagonal_matel(e.S, (const uint64_t *)rec, e.ilut_ref.data());
                    if (!AddNewHashDet(e, rec, SpawnedSign, diagH, offdiagH)) {
                        e.stats[NECI_ST_ERR_FLAGS] = (double)((int)e.stats[NECI_ST_ERR_FLAGS] | 2);
                        break;
                    }
                    e.stats[NECI_ST_NINSERTED] += 1;
                }
            }
        }
    }
    e.HolesInList = e.iEndFreeSlot - e.iStartFreeSlot;                      // :1343-1349

    // CalcHashTableStats, src/load_balancer.fpp:646-805
    double TotParts = 0, norm2 = 0, norm_ss2 = 0, inst_hf = 0, highest = 0;
    for (int64_t i = 0; i < e.TotWalkers; ++i) {
        double s = e.sign(i);
        const bool tDet = c.t_semi_stochastic && e.test_flag(i, NECI_FLAG_DETERMINISTIC);
        if (unocc(s) && !tDet) { /* AnnihilatedDet */ }
        else {
            if (!tDet && std::fabs(s) > 1.e-12 && std::fabs(s) < c.occupied_thresh) {
                const double pRemove = (c.occupied_thresh - std::fabs(s)) / c.occupied_thresh;
                Stream rng(c.seed, iter, det_hash64(e.orb(i), nw), 0, RNG_PRUNE);
                if (pRemove > rng.draw()) {
                    e.stats[NECI_ST_NOREMOVED] += std::fabs(s);
                    s = 0.0; e.set_sign(i, 0.0);
                    RemoveHashDet(e, i);
                    e.HolesInList += 1;
                } else {
                    e.stats[NECI_ST_NOBORN] += c.occupied_thresh - std::fabs(s);
                    s = dsign(c.occupied_thresh, s); e.set_sign(i, s);
                }
            }
            TotParts += std::fabs(s);
            norm2 += s * s;
            if (tDet) norm_ss2 += s * s;
            if (std::fabs(s) > highest) highest = (double)(int64_t)std::fabs(s);
        }
        if (excit_level_hphf(e.S, e.ilut_ref.data(), e.orb(i)) == 0) inst_hf = s;
    }
    e.stats[NECI_ST_TOTPARTS] = TotParts;
    e.stats[NECI_ST_NORM_PSI_SQ] = norm2;
    e.stats[NECI_ST_NORM_SEMISTOCH_SQ] = norm_ss2;
    e.stats[NECI_ST_INSTNOATHF] = inst_hf;
    e.stats[NECI_ST_HIGHEST_POP] = highest;
    e.stats[NECI_ST_TOTWALKERS] = (double)e.TotWalkers;
    e.stats[NECI_ST_HOLESINLIST] = (double)e.HolesInList;
}

}  // namespace

// ============================================================================
// C entry points (same shapes as include/neci_gpu.h)
// ============================================================================
extern "C" {

int orc_init(const neci_gpu_config *cfg, orc_engine **out) {
    orc_engine *e = new orc_engine();
    e->cfg = *cfg;
    e->nwords = cfg->nifd + 1; e->W = cfg->niftot + 1;
    e->random_orb_index.assign(cfg->random_orb_index, cfg->random_orb_index + cfg->nbasis);
    e->random_hash2.assign(cfg->random_hash2, cfg->random_hash2 + cfg->nbasis);
    e->lb_mapping.assign(cfg->load_balance_mapping, cfg->load_balance_mapping + cfg->balance_blocks);
    e->ilut_ref.assign((const uint64_t *)cfg->ilut_ref, (const uint64_t *)cfg->ilut_ref + e->nwords);
    e->cfg.random_orb_index = nullptr; e->cfg.random_hash2 = nullptr; e->cfg.load_balance_mapping = nullptr; e->cfg.ilut_ref = nullptr;
    e->S.type = cfg->system_type; e->S.nel = cfg->nel; e->S.nbasis = cfg->nbasis; e->S.nwords = e->nwords;
    e->S.nocc_alpha = cfg->nocc_alpha; e->S.nocc_beta = cfg->nocc_beta;
    e->S.t_exch = cfg->t_exch != 0; e->S.t_no_brillouin = cfg->t_no_brillouin != 0; e->S.ecore = cfg->ecore;
    e->S.t_hphf = cfg->t_hphf != 0;
    e->dets.assign((size_t)e->W * cfg->max_walkers, 0);
    e->diagH.assign(cfg->max_walkers, 0.0); e->offdiagH.assign(cfg->max_walkers, 0.0);
    e->FreeSlot.assign(cfg->max_walkers + 1, 0);
    e->spawned.resize(cfg->nranks);
    zero_stats(*e);
    *out = e;
    return 0;
}
int orc_finalize(orc_engine *e) { delete e; return 0; }

int orc_set_system_fcidump(orc_engine *e, const double *umat, int64_t n_umat, const double *tmat2d) {
    e->S.umat.assign(umat, umat + n_umat);
    e->S.tmat.assign(tmat2d, tmat2d + (size_t)e->cfg.nbasis * e->cfg.nbasis);
    return 0;
}
int orc_set_pchb(orc_engine *e, int32_t n_spat, int32_t ij_max, int32_t ab_max, const double *probs,
                 const double *bias, const int32_t *alias, const double *p_exch, const int32_t *tgt_orbs,
                 double p_singles, double p_doubles, double p_parallel, int32_t n_classes,
                 const int32_t *class_of_spinorb) {
    System &S = e->S;
    S.n_spat = n_spat; S.ij_max = ij_max; S.ab_max = ab_max;
    const size_t n = (size_t)ij_max * 3 * ab_max;
    S.probs.assign(probs, probs + n); S.bias.assign(bias, bias + n); S.alias.assign(alias, alias + n);
    S.p_exch.assign(p_exch, p_exch + ij_max); S.tgt_orbs.assign(tgt_orbs, tgt_orbs + 2 * (size_t)ab_max);
    S.p_singles = p_singles; S.p_doubles = p_doubles; S.p_parallel = p_parallel;
    S.n_classes = n_classes;
    S.class_of_spinorb.assign(class_of_spinorb, class_of_spinorb + e->cfg.nbasis);
    S.class_orbs.assign(n_classes, {});
    for (int o = 1; o <= e->cfg.nbasis; ++o) S.class_orbs[class_of_spinorb[o - 1]].push_back(o);
    return 0;
}
int orc_set_pchb_particles(orc_engine *e, int32_t mode, const double *p_first, const double *p_second) {
    System &S = e->S;
    if (mode < 0 || mode > 2) return 1;
    if (mode != 0 && e->cfg.nbasis > 128) return 1;
    S.pchb_particles = mode;
    if (mode != 0) {
        S.p_first.assign(p_first, p_first + e->cfg.nbasis);
        S.p_second.assign(p_second, p_second + (size_t)e->cfg.nbasis * e->cfg.nbasis);
    }
    return 0;
}
int orc_set_system_hubbard_rs(orc_engine *e, int32_t max_neigh, const int32_t *neighbours,
                              const double *tmat2d, double uhub) {
    e->S.max_neigh = max_neigh;
    e->S.neighbours.assign(neighbours, neighbours + (size_t)max_neigh * e->cfg.nbasis);
    e->S.tmat.assign(tmat2d, tmat2d + (size_t)e->cfg.nbasis * e->cfg.nbasis);
    e->S.uhub = uhub;
    return 0;
}
int orc_set_excit_probs(orc_engine *e, double p_singles, double p_doubles, double p_parallel) {
    e->S.p_singles = p_singles; e->S.p_doubles = p_doubles; e->S.p_parallel = p_parallel;
    return 0;
}
int orc_set_system_hubbard_k(orc_engine *e, int32_t n_k, const int32_t *ksum, const int32_t *kdiff,
                             const double *eps_k, double u_over_n) {
    e->S.n_k = n_k;
    e->S.ksum.assign(ksum, ksum + (size_t)n_k * n_k); e->S.kdiff.assign(kdiff, kdiff + (size_t)n_k * n_k);
    e->S.eps_k.assign(eps_k, eps_k + n_k); e->S.u_over_n = u_over_n;
    return 0;
}
int orc_set_core_space(orc_engine *e, int64_t n_local, const int64_t *row_ptr, const int32_t *col,
                       const double *val, const int32_t *sizes, const int32_t *displs,
                       const int64_t *core_iluts) {
    e->n_core_local = n_local;
    e->row_ptr.assign(row_ptr, row_ptr + n_local + 1);
    e->col.assign(col, col + row_ptr[n_local]); e->val.assign(val, val + row_ptr[n_local]);
    e->core_sizes.assign(sizes, sizes + e->cfg.nranks); e->core_displs.assign(displs, displs + e->cfg.nranks);
    e->n_core_total = 0; for (int r = 0; r < e->cfg.nranks; ++r) e->n_core_total += sizes[r];
    e->core_set.clear();
    for (int64_t i = 0; i < e->n_core_total; ++i)
        e->core_set.insert({(uint64_t)core_iluts[i * e->nwords], (e->nwords > 1) ? (uint64_t)core_iluts[i * e->nwords + 1] : 0ull});
    e->indices_of_determ_states.assign(n_local, 0);
    e->core_ham_diag.assign(n_local, 0.0);
    for (int64_t i = 0; i < n_local; ++i)
        for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
            if (col[k] == i + displs[e->cfg.rank]) e->core_ham_diag[i] = val[k];
    e->partial_determ_vecs.assign(n_local, 0.0); e->full_determ_vecs.assign(e->n_core_total, 0.0);
    return 0;
}

int orc_upload_walkers(orc_engine *e, const int64_t *current_dets, int64_t n, const double *gd, const double *go) {
    if (n > e->cfg.max_walkers) return 1;
    std::memcpy(e->dets.data(), current_dets, (size_t)n * e->W * 8);
    e->TotWalkers = n;
    e->hash.clear();
    for (int64_t j = 0; j < n; ++j) {
        const bool core = e->test_flag(j, NECI_FLAG_DETERMINISTIC);
        if (!unocc(e->sign(j)) || core) e->hash[e->key(e->orb(j))] = j;
        if (gd) e->diagH[j] = gd[j]; else e->diagH[j] = get_diagonal_matel(e->S, e->orb(j)) - e->cfg.hii;
        if (go) e->offdiagH[j] = go[j]; else e->offdiagH[j] = get_off_diagonal_matel(e->S, e->orb(j), e->ilut_ref.data());
        if (e->t_trial) trial_search_and_flag(*e, j);
    }
    return 0;
}
// init_trial_wf (src/trial_wf_gen.F90): install the two hash tables and flag the resident list
int orc_set_trial_space(orc_engine *e, int64_t n_trial, const int64_t *trial_iluts, const double *trial_amps,
                        int64_t n_con, const int64_t *con_iluts, const double *con_amps) {
    const int nw = e->nwords;
    e->trial_ht.clear(); e->con_ht.clear();
    for (int64_t i = 0; i < n_trial; ++i)
        e->trial_ht.insert({{(uint64_t)trial_iluts[i * nw], nw > 1 ? (uint64_t)trial_iluts[i * nw + 1] : 0ull}, trial_amps[i]});
    for (int64_t i = 0; i < n_con; ++i)
        e->con_ht.insert({{(uint64_t)con_iluts[i * nw], nw > 1 ? (uint64_t)con_iluts[i * nw + 1] : 0ull}, con_amps[i]});
    e->t_trial = true;
    e->current_trial_amps.assign((size_t)e->cfg.max_walkers, 0.0);
    for (int64_t j = 0; j < e->TotWalkers; ++j) trial_search_and_flag(*e, j);
    return 0;
}
int orc_download_walkers(orc_engine *e, int64_t *current_dets, int64_t *n, double *gd, double *go) {
    if (n) *n = e->TotWalkers;
    if (current_dets) std::memcpy(current_dets, e->dets.data(), (size_t)e->TotWalkers * e->W * 8);
    if (gd) std::memcpy(gd, e->diagH.data(), (size_t)e->TotWalkers * 8);
    if (go) std::memcpy(go, e->offdiagH.data(), (size_t)e->TotWalkers * 8);
    return 0;
}

// write_pops_det, src/Popsfile.F90:2054-2107: occupied determinants above binarypops_min_weight, in list order
int orc_download_occupied(orc_engine *e, double min_weight, int64_t *dets_out, int64_t *n, double *gd, double *go) {
    int64_t k = 0;
    for (int64_t j = 0; j < e->TotWalkers; ++j) {
        if (!(std::fabs(e->sign(j)) > min_weight)) continue;
        if (dets_out) {
            std::memcpy(dets_out + (size_t)k * e->W, &e->dets[(size_t)j * e->W], (size_t)e->W * 8);
            dets_out[(size_t)k * e->W + e->nwords + 1] &= ~(int64_t)1;
            if (gd) gd[k] = e->diagH[j];
            if (go) go[k] = e->offdiagH[j];
        }
        ++k;
    }
    if (n) *n = k;
    return 0;
}

// ---- phases, exposed so that a test harness can play the role of the exchange
int orc_spawn_phase(orc_engine *e, double tau, double diag_sft, int64_t iter) {
    spawn_phase(*e, tau, diag_sft, iter);
    return 0;
}
int64_t orc_spawned_count(orc_engine *e, int32_t dest) { return (int64_t)(e->spawned[dest].size() / e->W); }
int orc_spawned_get(orc_engine *e, int32_t dest, int64_t *out) {
    std::memcpy(out, e->spawned[dest].data(), e->spawned[dest].size() * 8);
    return 0;
}
int64_t orc_core_local(orc_engine *e) { return e->n_core_local; }
int orc_partial_vec_get(orc_engine *e, double *out) { std::memcpy(out, e->partial_determ_vecs.data(), e->n_core_local * 8); return 0; }
int orc_determ_projection(orc_engine *e, const double *full_vec, double tau, double diag_sft) {
    std::memcpy(e->full_determ_vecs.data(), full_vec, e->n_core_total * 8);
    determ_projection(*e, tau, diag_sft);
    return 0;
}
// receives the already exchanged spawns of this rank; adds to the stats of the spawn phase
int orc_annihilate_phase(orc_engine *e, const int64_t *spawned_parts, int64_t n_spawned, int64_t iter, double *stats_out) {
    std::vector<int64_t> sp(spawned_parts, spawned_parts + (size_t)n_spawned * e->W);
    annihilate_phase(*e, sp, iter);
    if (stats_out) std::memcpy(stats_out, e->stats, sizeof(e->stats));
    return 0;
}
// standalone annihilation of a fixed list (stats zeroed first; FreeSlot rebuilt from the list)
int orc_annihilate(orc_engine *e, const int64_t *spawned_parts, int64_t n_spawned, int64_t iter, double *stats_out) {
    zero_stats(*e);
    e->iStartFreeSlot = e->iEndFreeSlot = 0;
    for (int64_t j = 0; j < e->TotWalkers; ++j)
        if (unocc(e->sign(j)) && !e->test_flag(j, NECI_FLAG_DETERMINISTIC)) e->FreeSlot[e->iEndFreeSlot++] = j;
    return orc_annihilate_phase(e, spawned_parts, n_spawned, iter, stats_out);
}

// whole iteration on a "world" of engines (ranks 0..n-1 of one job), threads play MPI ranks
int orc_world_iterate(orc_engine **es, int32_t n, double tau, double diag_sft, int64_t iter, double *stats_out, int32_t nthreads) {
    auto par = [&](auto fn) {
        if (nthreads <= 1 || n == 1) { for (int r = 0; r < n; ++r) fn(r); return; }
        std::vector<std::thread> th;
        for (int r = 0; r < n; ++r) th.emplace_back(fn, r);
        for (auto &t : th) t.join();
    };
    par([&](int r) { spawn_phase(*es[r], tau, diag_sft, iter); });
    if (es[0]->cfg.t_semi_stochastic && es[0]->n_core_total > 0) {     // no core space handed over yet: nothing to project
        std::vector<double> full(es[0]->n_core_total);
        for (int r = 0; r < n; ++r)
            std::memcpy(&full[es[r]->core_displs[r]], es[r]->partial_determ_vecs.data(), es[r]->n_core_local * 8);
        par([&](int r) { es[r]->full_determ_vecs = full; determ_projection(*es[r], tau, diag_sft); });
    }
    std::vector<std::vector<int64_t>> recv(n);
    for (int r = 0; r < n; ++r)       // SendProcNewParts: contiguous by source rank
        for (int s = 0; s < n; ++s) recv[r].insert(recv[r].end(), es[s]->spawned[r].begin(), es[s]->spawned[r].end());
    par([&](int r) { annihilate_phase(*es[r], recv[r], iter); });
    if (stats_out)
        for (int r = 0; r < n; ++r) std::memcpy(stats_out + (size_t)r * NECI_ST_COUNT, es[r]->stats, sizeof(es[r]->stats));
    return 0;
}
int orc_iterate(orc_engine *e, double tau, double diag_sft, int64_t iter, double *stats_out) {
    orc_engine *es[1] = {e};
    return orc_world_iterate(es, 1, tau, diag_sft, iter, stats_out, 1);
}

// ---- dynamic load balancing -------------------------------------------------
// block populations, src/load_balancer.fpp:216-235: sum(ceiling(abs(sgn))) per block
int orc_block_populations(orc_engine *e, double *block_parts) {
    for (int b = 0; b < e->cfg.balance_blocks; ++b) block_parts[b] = 0.0;
    int nI[128];
    for (int64_t j = 0; j < e->TotWalkers; ++j) {
        const double s = e->sign(j);
        if (unocc(s)) continue;
        decode(e->orb(j), e->cfg.nbasis, nI);
        block_parts[det_block(*e, nI) - 1] += std::ceil(std::fabs(s));
    }
    return 0;
}
// move_block, src/load_balancer.fpp:353-512, for a whole new mapping at once: determinants whose block now
// belongs to another rank are removed here (RemoveHashDet + null sign) and added there (AddNewHashDet with
// recomputed H_ii / H_0i).
int orc_world_rebalance(orc_engine **es, int32_t n, const int32_t *new_mapping) {
    std::vector<std::vector<int64_t>> moved(n);
    int nI[128];
    for (int r = 0; r < n; ++r) {
        orc_engine &e = *es[r];
        e.lb_mapping.assign(new_mapping, new_mapping + e.cfg.balance_blocks);
        e.iStartFreeSlot = e.iEndFreeSlot = 0;
        for (int64_t j = 0; j < e.TotWalkers; ++j) {
            const double s = e.sign(j);
            if (unocc(s)) { if (!e.test_flag(j, NECI_FLAG_DETERMINISTIC)) e.FreeSlot[e.iEndFreeSlot++] = j; continue; }
            decode(e.orb(j), e.cfg.nbasis, nI);
            const int owner = det_node(e, nI);
            if (owner != r) {
                const int64_t *rec = &e.dets[(size_t)j * e.W];
                std::vector<int64_t> tmp(rec, rec + e.W);
                tmp[e.nwords + 1] &= ~1ll;
                moved[owner].insert(moved[owner].end(), tmp.begin(), tmp.end());
                RemoveHashDet(e, j);
                e.set_sign(j, 0.0);
            }
        }
    }
    for (int r = 0; r < n; ++r) {
        orc_engine &e = *es[r];
        const int64_t m = (int64_t)(moved[r].size() / e.W);
        for (int64_t i = 0; i < m; ++i) {
            const int64_t *rec = &moved[r][(size_t)i * e.W];
            const double diagH = get_diagonal_matel(e.S, (const uint64_t *)rec);
            const double offdiagH = get_off_diagonal_matel(e.S, (const uint64_t *)rec, e.ilut_ref.data());
            if (!AddNewHashDet(e, rec, sign_to_double(rec[e.nwords]), diagH, offdiagH)) return 1;
        }
    }
    return 0;
}

// ---- probes ---------------------------------------------------------------
int orc_probe_det_node(orc_engine *e, int64_t n, const int64_t *iluts, int32_t *block_out, int32_t *node_out) {
    int nI[128];
    for (int64_t i = 0; i < n; ++i) {
        decode((const uint64_t *)&iluts[(size_t)i * e->nwords], e->cfg.nbasis, nI);
        const int b = det_block(*e, nI);
        block_out[i] = b; node_out[i] = e->lb_mapping[b - 1];
    }
    return 0;
}
int orc_probe_walker_hash(orc_engine *e, int64_t n, const int64_t *iluts, int32_t table_len, int32_t *out) {
    int nI[128];
    for (int64_t i = 0; i < n; ++i) {
        decode((const uint64_t *)&iluts[(size_t)i * e->nwords], e->cfg.nbasis, nI);
        out[i] = find_walker_hash(*e, nI, table_len);
    }
    return 0;
}
int orc_probe_helement(orc_engine *e, int64_t n, const int64_t *ii, const int64_t *ij, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t *a = (const uint64_t *)&ii[(size_t)i * e->nwords], *b = (const uint64_t *)&ij[(size_t)i * e->nwords];
        if (!e->S.t_hphf) out[i] = get_helement(e->S, a, b);
        else out[i] = (DetBitLT(a, b, e->nwords) == 0) ? hphf_diag_helement(e->S, a) : hphf_off_diag_helement(e->S, a, b);
    }
    return 0;
}
int orc_probe_gen_excit(orc_engine *e, int64_t n, const int64_t *iluts, const int32_t *attempt, int64_t iter,
                        int64_t *ilut_j_out, int32_t *ic_out, int32_t *ex_out, int32_t *parity_out,
                        double *pgen_out, double *hel_out) {
    int nI[128];
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t *il = (const uint64_t *)&iluts[(size_t)i * e->nwords];
        decode(il, e->cfg.nbasis, nI);
        Stream rng(e->cfg.seed, iter, det_hash64(il, e->nwords), (uint32_t)attempt[i], RNG_ATTEMPT);
        Excitation E;
        double HElGen = 0.0;
        if (e->S.t_hphf) gen_hphf_excit(e->S, nI, il, rng, E, HElGen);
        else generate_excitation(e->S, nI, il, rng, E);
        for (int w = 0; w < e->nwords; ++w) ilut_j_out[(size_t)i * e->nwords + w] = E.valid ? (int64_t)E.ilutJ[w] : 0;
        ic_out[i] = E.ic;
        for (int k = 0; k < 4; ++k) ex_out[4 * i + k] = E.valid ? E.ex[k] : 0;
        parity_out[i] = E.valid ? (E.parity ? 1 : 0) : 0;
        pgen_out[i] = E.valid ? E.pgen : 0.0;
        hel_out[i] = E.valid ? (e->S.t_hphf ? HElGen : get_spawn_helement(e->S, nI, E)) : 0.0;
    }
    return 0;
}
// get_pgen recomputation for PCHB doubles (reference test: get_pgen == returned pgen)
int orc_probe_pchb_pgen(orc_engine *e, int64_t n, const int32_t *ex, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = e->S.p_doubles * pchb_double_get_pgen(e->S, &ex[4 * i]);
    return 0;
}
// the same for a selector whose probability depends on the determinant (FULL-FULL particle selection)
int orc_probe_pchb_pgen_det(orc_engine *e, int64_t n, const int64_t *iluts, const int32_t *ex, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t il[2] = {(uint64_t)iluts[(size_t)i * e->nwords], e->nwords > 1 ? (uint64_t)iluts[(size_t)i * e->nwords + 1] : 0ull};
        int nI[128]; decode(il, e->cfg.nbasis, nI);
        out[i] = e->S.p_doubles * pchb_double_get_pgen(e->S, &ex[4 * i], nI);
    }
    return 0;
}
// direct known-answer probes of the lattice elements with an explicit excitation matrix
// (the shapes the reference's unit tests call: get_offdiag_helement_k_sp_hub(nI, ex, tpar),
//  get_offdiag_helement_rs_hub(nI, ex, tpar), get_diag_helement_k_sp_hub(nI))
int orc_probe_offdiag_k(orc_engine *e, const int32_t *ex4, int32_t tpar, double *out) {
    *out = e->S.offdiag_k_hub(ex4, tpar != 0); return 0;
}
int orc_probe_offdiag_rs(orc_engine *e, int32_t src, int32_t tgt, int32_t tpar, double *out) {
    *out = e->S.offdiag_rs_hub(src, tgt, tpar != 0); return 0;
}
int orc_probe_diag(orc_engine *e, const int64_t *ilut, double *out) {
    *out = get_diagonal_matel(e->S, (const uint64_t *)ilut); return 0;
}
// UMatInd (src/UMatCache.F90:257-296) and the alias sampler in isolation
int64_t orc_probe_umat_ind(int32_t i, int32_t j, int32_t k, int32_t l) { return System::UMatInd(i, j, k, l); }
// draws n samples from one alias table (probs/bias/alias of length len, 1-based alias) with the
// stream (seed, iter=0, h=stream_id, attempt = draw index): histogram of the 1-based results.
int orc_probe_alias_hist(const double *bias, const int32_t *alias, int32_t len, uint64_t seed, int64_t ndraw, int64_t *hist) {
    for (int i = 0; i < len; ++i) hist[i] = 0;
    for (int64_t d = 0; d < ndraw; ++d) {
        Stream rng(seed, d >> 32, 0x1234567ull, (uint32_t)d, RNG_ATTEMPT);
        const double r = rng.draw();
        const int pos = (int)(len * r) + 1;
        const double b = std::max(len * r + 1 - pos, 0.0);
        const int ind = (b < bias[pos - 1]) ? pos : alias[pos - 1];
        hist[ind - 1] += 1;
    }
    return 0;
}
// raw Philox block (known-answer test) and the double stream
int orc_probe_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { Philox::gen(ctr, key, out); return 0; }
int orc_probe_stream(uint64_t seed, int64_t iter, const int64_t *ilut, int32_t nwords, int32_t attempt, int32_t purpose, int32_t n, double *out) {
    Stream s(seed, iter, det_hash64((const uint64_t *)ilut, nwords), (uint32_t)attempt, (uint32_t)purpose);
    for (int i = 0; i < n; ++i) out[i] = s.draw();
    return 0;
}
int orc_probe_make_double(int32_t nel, const int32_t *nI, int32_t e1, int32_t e2, int32_t t1, int32_t t2, int32_t *nJ, int32_t *ex, int32_t *par) {
    bool p; make_double(nI, nel, e1, e2, t1, t2, nJ, ex, p); *par = p; return 0;
}
int orc_probe_make_single(int32_t nel, const int32_t *nI, int32_t e1, int32_t t1, int32_t *nJ, int32_t *ex, int32_t *par) {
    bool p; make_single(nI, nel, e1, t1, nJ, ex, p); *par = p; return 0;
}

}  // extern "C"
