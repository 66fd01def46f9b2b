// K1: the loop over determinants of PerformFCIMCycPar (src/FciMCPar.F90:1294-1758), fused:
//   CalcParentFlag (fcimc_helper.F90:1036-1243), SumEContrib (:518-802), decide_num_to_spawn (:2160-2174),
//   generate_excitation + attempt_create_normal (fcimc_pointed_fns.F90:178-491), create_particle
//   (fcimc_helper.F90:152-308), walker_death / attempt_die_normal (:2279-2407, fcimc_pointed_fns.F90:573-705).
//
// B200 design (round 2: every warp is its own pipeline).  Walkers per determinant vary from 1 to 1e5+, a third of
// the PCHB draws are null excitations and a tenth are singles with an O(nel) matrix element, so attempts -- not
// determinants -- are the unit of parallelism and work moves between *stages* through queues so that every stage
// runs with full warps.  In round 1 the queues were shared by a 256-thread CTA: block barriers between the stages
// (14 % of the stall samples), atomics on the queue counters, a block-wide prefix sum per tile.  Now a warp owns
// everything it needs -- its chunk of the list, its attempt map, its two queues, all in its private slice of shared
// memory -- so the stages are separated by __syncwarp only, queue counters are warp-uniform registers (a push is a
// ballot and a popcount), and no warp ever waits for another:
//
//   stage A  the warp loads a chunk of 32 x K1_SPT consecutive slots (all SoA streams of all its slots requested up
//            front), one lane per slot: flags, energy sums, death, attempt count; parents (det, stream id, info)
//            and the warp-wide prefix sum of the attempt counts go to the warp's shared memory
//   stage B1 one lane per attempt (parents expand their index into an attempt -> parent map): draw the excitation from
//            ONE Philox block.  valid doubles / lattice excitations -> queue QE {parent det, stream id, attempt,
//            orbitals, pgen};  PCHB singles -> queue QS {parent det, stream id, attempt};  null draws stop here
//   stage B2 whenever QE holds >= 32 entries: parity (popc), matrix element (2 UMAT loads), spawn weight, the
//            rounding draw (its own Philox block, computed here with every lane busy), stochastic rounding,
//            warp-aggregated append to the destination rank's segment
//   stage B3 whenever QS holds >= 32 entries: uniform single + sltcnd_1 (loads batched 4 at a time), then as B2
//
// Statistics of stage A are reduced across the warp only when some lane has something to add (death, energy and
// sign-flip events are rare per chunk), counts go through the integer reduction instruction.  Determinants above
// NG_HEAVY attempts go to k_spawn_heavy, which spreads their attempts over all warps of the grid.  On several ranks
// the appends of B2/B3 go to a staging list and k_partition_push routes them afterwards (kernels.cuh).  HPHF runs are
// their own compile-time variant (NG_SYS_PCHB_HPHF).
//
// Random numbers are counter-based (device_common.cuh: Stream), so the result does not depend on the order in
// which the queues are served, nor on which warp handles which chunk.
#pragma once
#include "device_system.cuh"

namespace ng {

#define NG_BLOCK 256         /* block size of the streaming kernels */
#ifndef K1_BLOCK
#define K1_BLOCK 256         /* threads per CTA of the spawning kernel (8 independent warps) */
#endif
#define K1_WARPS (K1_BLOCK / 32)
#ifndef K1_CTAS_PER_SM       /* launch bound: resident CTAs per SM the register allocation must allow */
#define K1_CTAS_PER_SM (1024 / K1_BLOCK)
#endif
#define NG_HEAVY 1024        /* attempts per determinant handled inside a chunk */
#ifndef K1_SPT1
#define K1_SPT1 4            /* slots per lane and chunk, one-word determinants */
#endif
#ifndef K1_SPT2
#define K1_SPT2 2            /* slots per lane and chunk, two-word determinants */
#endif
template <int NW> __host__ __device__ constexpr int k1_spt() { return NW == 1 ? K1_SPT1 : K1_SPT2; }
#define K1_QCAP 64           /* queue capacity per warp: < 32 entries before a push of at most 32 */
#define K1_MAPW 128          /* attempts per window of the attempt -> parent map */

struct SpawnBuf {
    long long *buf;          // SpawnedParts: nranks segments of seg_cap records (W words each)
    long long *recv;         // received records (contiguous)
    unsigned long long *cnt; // ValidSpawnedList - InitialSpawnedSlots, per destination rank
    long long seg_cap;
    int W;
    // spawn-merge hash table, entries [stamp:16][tag:16][index:32]
    u64 *sht; u64 sht_cap;
    long long *acc_hi, *acc_lo;       // real coefficients: order-independent fixed-point sums of the records merged into
                                      // a representative (k_compress adds, k_annihilate reads and re-zeroes)
    int *ins_idx;            // records that become new determinants
    long long *heavy;        // (slot, nspawn) pairs
    long long heavy_cap;
    unsigned long long *n_recv_dev;   // received-record count left on the device by the peer-memory exchange
    long long *stage;                 // nranks > 1: spawns of the spawning kernel before they are routed (k_partition)
    unsigned long long *stage_cnt;
    long long stage_cap;
};

struct IterArgs {
    double tau, diag_sft;
    long long iter;
    long long n_recv;        // -1: read SB.cnt[0] on the device (single rank); -2: read *SB.n_recv_dev (peer-memory exchange)
    u32 stamp;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// statistics accumulated by K1, in the order of their rows in K1Shared::wacc
enum { W_NODIED = 0, W_NOBORN_D, W_ABORT, W_HF, W_DOUBS, W_ENUM, W_ENUMABS, W_INITSENUM, W_INITD, W_NINITD, W_INITW,
       W_NINITW, W_ADDED, W_CHILD, W_CHILD_SING, W_VALID, W_INVALID, W_MAXSP, W_COUNT };

// a warp's private slice of shared memory
template <int NW> struct K1Warp {
    static constexpr int CHUNK = 32 * k1_spt<NW>();
    // parents of the current chunk
    u64 p_d0[CHUNK];
    u64 p_d1[(NW > 1) ? CHUNK : 1];
    u64 p_h[CHUNK];
    int p_off[CHUNK + 1];             // exclusive prefix sum of the attempt counts; [CHUNK] = total
    unsigned char p_info[CHUNK];
    unsigned char map[K1_MAPW];      // attempt (within the current window) -> parent index in the chunk
    // QE: generated excitations waiting for their matrix element
    u64 q_d0[K1_QCAP];
    u64 q_d1[(NW > 1) ? K1_QCAP : 1];
    u64 q_h[K1_QCAP];
    double q_pgen[K1_QCAP];
    u32 q_att[K1_QCAP];
    u32 q_orbs[K1_QCAP];     // src1 | src2 << 8 | tgt1 << 16 | tgt2 << 24
    u32 q_misc[K1_QCAP];     // info | ic << 8
    // QS: PCHB single excitations still to be generated
    u64 s_d0[K1_QCAP];
    u64 s_d1[(NW > 1) ? K1_QCAP : 1];
    u64 s_h[K1_QCAP];
    u32 s_att[K1_QCAP];
    u32 s_misc[K1_QCAP];
};
template <int NW> struct K1Shared {
    K1Warp<NW> w[K1_WARPS];
    double wacc[K1_WARPS][W_COUNT];
    int roi[NG_MAX_BASIS];           // RandomOrbIndex (append_spawn's DetermineDetNode; unused on one rank)
    int bloom_cnt[2];
    unsigned long long bloom_max[2];
    // tau search (log_spawn_magnitude): classes 0 singles, 1 doubles, 2 parallel doubles, 3 opposite-spin doubles
    int tau_cnt[4];
    unsigned long long tau_gamma[4];
};
// a warp's queue fill levels: identical in all lanes, kept in registers
struct K1Queues { int qe, qs; };

// stochastic_round (src/lib/util_mod.fpp:182-204) with the random number drawn by the caller
__device__ __forceinline__ double stochastic_round_r(double r, double u) {
    int i = (int)r;
    const double res = r - (double)i;
    if (fabs(res) >= 1.0e-12) {
        if (fabs(res) > u) i += (r < 0.0 || (r == 0.0 && signbit(r))) ? -1 : 1;
    }
    return (double)i;
}

// create_particle (src/fcimc_helper.F90:152-308): warp-aggregated append of (ilutJ, child, flags) to the
// destination rank's segment of SpawnedParts.  Must be called by all 32 lanes.
template <int NW>
__device__ __forceinline__ void append_spawn(const Params &P, const SpawnBuf &SB, const WalkerList &L, const int *roi,
                                             bool has, const Det<NW> &detJ, double child, long long flags) {
    const u32 lane = threadIdx.x & 31;
    int proc = 0;
    if (has && P.nranks > 1) proc = __ldg(&P.lb_mapping[det_block<NW>(P, roi, detJ) - 1]);
    const u32 active = __ballot_sync(0xffffffffu, has);
    if (!has) return;
    u32 peers = active;
    if (P.nranks > 1) peers = __match_any_sync(active, proc);
    const int leader = __ffs(peers) - 1;
    const int rank_in = __popc(peers & ((1u << lane) - 1u));
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(&SB.cnt[proc], (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const long long pos = (long long)base + rank_in;
    if (pos >= SB.seg_cap) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 1ull); return; }
    long long *rec = SB.buf + ((size_t)proc * SB.seg_cap + pos) * SB.W;
    rec[0] = (long long)detJ.w[0];
    if (NW > 1) rec[NW - 1] = (long long)detJ.w[NW - 1];
    rec[NW] = __double_as_longlong(child);
    rec[NW + 1] = flags;
}
// The spawning kernel's own append.  On one rank this is create_particle itself.  On several ranks the spawn goes
// to a staging list first and k_partition_push routes it afterwards: DetermineDetNode costs ~230 instructions, and
// the partition kernel hashes with every lane busy and the record already on its way over NVLink.
template <int NW>
__device__ __forceinline__ void append_spawn_k1(const Params &P, const SpawnBuf &SB, const WalkerList &L, const int *roi,
                                                bool has, const Det<NW> &detJ, double child, long long flags) {
    if (P.nranks == 1) { append_spawn<NW>(P, SB, L, roi, has, detJ, child, flags); return; }
    const u32 lane = threadIdx.x & 31;
    const u32 active = __ballot_sync(0xffffffffu, has);
    if (!has) return;
    const int leader = __ffs(active) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(SB.stage_cnt, (unsigned long long)__popc(active));
    base = __shfl_sync(active, base, leader);
    const long long pos = (long long)base + __popc(active & ((1u << lane) - 1u));
    if (pos >= SB.stage_cap) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 1ull); return; }
    long long *rec = SB.stage + (size_t)pos * SB.W;
    rec[0] = (long long)detJ.w[0];
    if (NW > 1) rec[NW - 1] = (long long)detJ.w[NW - 1];
    rec[NW] = __double_as_longlong(child);
    rec[NW + 1] = flags;
}

// per-thread accumulators of the attempt stages
struct AttAcc {
    double child, child_sing, maxsp;
    int valid, invalid;
};

// attempt_create_normal (fcimc_pointed_fns.F90:178-491) for an excitation whose orbitals and pgen are known, followed
// by create_particle.  `active` lanes hold a real entry; all lanes must call.  (h, att) name the attempt's stream:
// the rounding number is the first of its RNG_ATT_ROUND stream.
template <int NW, int SYS>
__device__ __forceinline__ void evaluate_and_append(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                                    K1Shared<NW> &S, bool active, const Det<NW> &d, Excit<NW> &E, int info,
                                                    u64 h, u32 att, AttAcc &acc) {
    bool has = false;
    double child = 0.0;
    long long cflags = 0;
    int tau_cls = -1;
    if (active) {
        finalize_excit(d, E);
        bool cancelled = false;
        double rh_hphf = 0.0;
        if (sys_hphf(SYS)) {
            // gen_hphf_excit wraps the generator (fcimc_initialisation.fpp:2162-2165): representative, pgen, element
            if (!hphf_fixup<NW, SYS>(P, d, E, rh_hphf)) { cancelled = true; acc.valid -= 1; acc.invalid += 1; }
        }
        if (!cancelled && P.t_semi_stochastic && (info & 4)) {
            // core -> core spawning is done by determ_projection (FciMCPar.F90:1651-1670)
            if (is_core_state<NW>(P, E.detJ)) cancelled = true;
            cflags = F_DPARENT;
        }
        if (!cancelled) {
            const double prob = E.pgen * P.av_mc_excits;
            const double rh = sys_hphf(SYS) ? rh_hphf : spawn_helement<NW, SYS>(P, d, E);
            const double ww = (info & 1) ? -1.0 : 1.0;
            if (P.t_tau_search) {
                // log_spawn_magnitude (tau/tau_search_conventional.F90:138-260): gamma = |H_ij| / (prob / p_class)
                double tp;
                if (E.ic == 1) { tp = prob / P.p_singles; tau_cls = 0; }
                else {
                    tp = prob / P.p_doubles; tau_cls = 1;
                    if (P.t_consider_par_bias) {
                        if (((E.src1 ^ E.src2) & 1) == 0) { tp = tp / P.p_parallel; tau_cls = 2; }
                        else { tp = tp / (1.0 - P.p_parallel); tau_cls = 3; }
                    }
                }
                const double g = fabs(rh) / tp;
                if (tau_cls < 2 && !(g > 0.0)) tau_cls = -1;       // singles / plain doubles are counted when gamma > 0
                else {
                    const unsigned long long gb = (unsigned long long)__double_as_longlong(g);
                    if (gb > S.tau_gamma[tau_cls]) atomicMax(&S.tau_gamma[tau_cls], gb);
                }
            }
            double nSpawn = -A.tau * rh * ww / prob;
            acc.maxsp = fmax(acc.maxsp, fabs(nSpawn));
            if (P.t_all_real_coeff) {
                if (P.t_real_spawn_cutoff && fabs(nSpawn) < P.real_spawn_cutoff) {
                    Stream rr(P.seed, A.iter, h, att, RNG_ATT_ROUND);
                    nSpawn = P.real_spawn_cutoff * stochastic_round_r(nSpawn / P.real_spawn_cutoff, rr.draw53());
                }
            } else {
                Stream rr(P.seed, A.iter, h, att, RNG_ATT_ROUND);
                nSpawn = stochastic_round_r(nSpawn, rr.draw53());
            }
            if (fabs(nSpawn) > NG_EPS) {
                const double ac = fabs(nSpawn);
                acc.child += ac;                               // NoBorn and acceptances
                if (E.ic == 1) acc.child_sing += ac;           // SpawnFromSing
                if (ac > P.initiator_walk_no) {                // bloom statistics (rare)
                    const int b = (E.ic == 1) ? 0 : 1;
                    atomicAdd(&S.bloom_cnt[b], 1);
                    atomicMax(&S.bloom_max[b], (unsigned long long)__double_as_longlong(ac));
                }
                has = true; child = nSpawn;
                if (P.t_trunc_initiator && (info & 2)) cflags |= F_INIT;
            }
        }
    }
    if (P.t_tau_search) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const u32 m = __ballot_sync(0xffffffffu, tau_cls == c);
            if (m && (threadIdx.x & 31) == 0) atomicAdd(&S.tau_cnt[c], __popc(m));
        }
    }
    append_spawn_k1<NW>(P, SB, L, S.roi, has, E.detJ, child, cflags);
}

// stage B1: one spawning attempt of parent (d, h, info), attempt index p.  All lanes must call.
template <int NW, int SYS>
__device__ __forceinline__ void stage_generate(const Params &P, const WalkerList &L, const IterArgs &A, K1Warp<NW> &W,
                                               K1Queues &Q, bool active, const Det<NW> &d, u64 h, int info, u32 p, AttAcc &acc) {
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    bool push_e = false, push_s = false;
    Excit<NW> E;
    E.ic = 2; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = false; E.err = 0; E.parity = false;
    if (active) {
        Stream rng(P.seed, A.iter, h, p, RNG_ATTEMPT);
        if (sys_pchb(SYS)) {
            const double u = rng.draw53();                               // gen_exc_sd
            if (u < P.p_singles) push_s = true;                          // single: generated in stage B3
            else { gen_pchb_double(P, d, (u - P.p_singles) * P.inv_1m_ps, rng, E); E.pgen = E.pgen * P.p_doubles; }
        } else generate_excitation_core<NW, SYS>(P, d, rng, E);
        if (!push_s) {
            if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
            if (E.valid) { acc.valid += 1; push_e = true; }
            else acc.invalid += 1;
        }
    }
    const u32 me = __ballot_sync(0xffffffffu, push_e);
    if (push_e) {
        const int q = Q.qe + __popc(me & lt);
        W.q_d0[q] = d.w[0]; if (NW > 1) W.q_d1[q] = d.w[NW - 1];
        W.q_h[q] = h; W.q_att[q] = p; W.q_pgen[q] = E.pgen;
        W.q_orbs[q] = (u32)E.src1 | ((u32)E.src2 << 8) | ((u32)E.tgt1 << 16) | ((u32)E.tgt2 << 24);
        W.q_misc[q] = (u32)info | ((u32)E.ic << 8);
    }
    Q.qe += __popc(me);
    if (sys_pchb(SYS)) {
        const u32 ms = __ballot_sync(0xffffffffu, push_s);
        if (push_s) {
            const int q = Q.qs + __popc(ms & lt);
            W.s_d0[q] = d.w[0]; if (NW > 1) W.s_d1[q] = d.w[NW - 1];
            W.s_h[q] = h; W.s_att[q] = p; W.s_misc[q] = (u32)info;
        }
        Q.qs += __popc(ms);
    }
    __syncwarp();
}

// stage B2: serve n <= 32 entries from the top of QE
template <int NW, int SYS>
__device__ __forceinline__ void stage_evaluate(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                               K1Shared<NW> &S, K1Warp<NW> &W, int first, int n, AttAcc &acc) {
    const int lane = threadIdx.x & 31;
    const int i = first + lane;
    const bool active = lane < n;
    Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
    Excit<NW> E; E.ic = 2; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = active; E.err = 0; E.detJ = d; E.parity = false;
    int info = 0; u64 h = 0; u32 att = 0;
    if (active) {
        d.w[0] = W.q_d0[i]; if (NW > 1) d.w[NW - 1] = W.q_d1[i];
        const u32 o = W.q_orbs[i], m = W.q_misc[i];
        E.src1 = o & 0xff; E.src2 = (o >> 8) & 0xff; E.tgt1 = (o >> 16) & 0xff; E.tgt2 = o >> 24;
        E.ic = (m >> 8) & 0xff; info = m & 0xff;
        E.pgen = W.q_pgen[i]; h = W.q_h[i]; att = W.q_att[i];
    }
    __syncwarp();                                           // the entries are in registers: the queue may be refilled
    evaluate_and_append<NW, SYS>(P, L, SB, A, S, active, d, E, info, h, att, acc);
}

// stage B3 (PCHB only): serve n <= 32 deferred singles from the top of QS
template <int NW, int SYS>
__device__ __forceinline__ void stage_singles(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                              K1Shared<NW> &S, K1Warp<NW> &W, int first, int n, AttAcc &acc) {
    const int lane = threadIdx.x & 31;
    const int i = first + lane;
    bool active = lane < n;
    Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
    Excit<NW> E; E.ic = 1; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = false; E.err = 0; E.detJ = d; E.parity = false;
    int info = 0; u64 h = 0; u32 att = 0;
    if (active) {
        d.w[0] = W.s_d0[i]; if (NW > 1) d.w[NW - 1] = W.s_d1[i];
        info = (int)W.s_misc[i]; h = W.s_h[i]; att = W.s_att[i];
    }
    __syncwarp();
    if (active) {
        Stream rng(P.seed, A.iter, h, att, RNG_ATTEMPT, 4);     // the first block chose "single"; singles draw from word 4
        gen_uniform_single(P, d, rng, E);
        E.pgen = E.pgen * P.p_singles;
        if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
        if (E.valid) acc.valid += 1;
        else { acc.invalid += 1; active = false; }
    }
    evaluate_and_append<NW, SYS>(P, L, SB, A, S, active, d, E, info, h, att, acc);
}

// serve the queues while they hold at least `level` entries (level = 32 keeps every stage at full width,
// level = 1 drains).  Must be called by the whole warp.
template <int NW, int SYS>
__device__ __forceinline__ void serve_queues(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                             K1Shared<NW> &S, K1Warp<NW> &W, K1Queues &Q, int level, AttAcc &acc) {
    while (Q.qe >= level) {
        const int n = min(Q.qe, 32);
        Q.qe -= n;
        stage_evaluate<NW, SYS>(P, L, SB, A, S, W, Q.qe, n, acc);
    }
    if (sys_pchb(SYS)) {
        while (Q.qs >= level) {
            const int n = min(Q.qs, 32);
            Q.qs -= n;
            stage_singles<NW, SYS>(P, L, SB, A, S, W, Q.qs, n, acc);
        }
    }
}

template <int NW>
__device__ __forceinline__ void k1_init_shared(const Params &P, K1Shared<NW> &S) {
#pragma unroll 1
    for (int i = threadIdx.x; i < P.nbasis; i += K1_BLOCK) S.roi[i] = P.random_orb_index[i];
    for (int i = threadIdx.x; i < K1_WARPS * W_COUNT; i += K1_BLOCK) (&S.wacc[0][0])[i] = 0.0;
    if (threadIdx.x == 0) { S.bloom_cnt[0] = S.bloom_cnt[1] = 0; S.bloom_max[0] = S.bloom_max[1] = 0ull; }
    if (threadIdx.x < 4) { S.tau_cnt[threadIdx.x] = 0; S.tau_gamma[threadIdx.x] = 0ull; }
}

// end of kernel: per-thread attempt accumulators -> per-warp rows -> one partial row per CTA
template <int NW>
__device__ __forceinline__ void k1_flush(const WalkerList &L, K1Shared<NW> &S, const AttAcc &acc, double *partials, bool with_stage_a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double c = warp_sum(acc.child), cs = warp_sum(acc.child_sing), mx = warp_max(acc.maxsp);
    const int v = __reduce_add_sync(0xffffffffu, acc.valid), iv = __reduce_add_sync(0xffffffffu, acc.invalid);
    if (lane == 0) {
        S.wacc[warp][W_CHILD] = c; S.wacc[warp][W_CHILD_SING] = cs; S.wacc[warp][W_VALID] = (double)v; S.wacc[warp][W_INVALID] = (double)iv;
        S.wacc[warp][W_MAXSP] = mx;
    }
    double *row = partials + (size_t)blockIdx.x * NECI_ST_COUNT;
    for (int k = threadIdx.x; k < NECI_ST_COUNT; k += K1_BLOCK) row[k] = 0.0;
    __syncthreads();
    if (threadIdx.x < W_COUNT) {
        const int k = threadIdx.x;
        double t = S.wacc[0][k];
        for (int w = 1; w < K1_WARPS; ++w) t = (k == W_MAXSP) ? fmax(t, S.wacc[w][k]) : t + S.wacc[w][k];
        S.wacc[0][k] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *t = S.wacc[0];
        row[NECI_ST_NOBORN] = t[W_CHILD] + (with_stage_a ? t[W_NOBORN_D] : 0.0);
        row[NECI_ST_ACCEPTANCES] = t[W_CHILD];
        row[NECI_ST_SPAWNFROMSING] = t[W_CHILD_SING];
        row[NECI_ST_NVALIDEXCITS] = t[W_VALID];
        row[NECI_ST_NINVALIDEXCITS] = t[W_INVALID];
        row[NECI_ST_MAX_CYC_SPAWN] = t[W_MAXSP];
        row[NECI_ST_BLOOM_COUNT_1] = (double)S.bloom_cnt[0];
        row[NECI_ST_BLOOM_COUNT_2] = (double)S.bloom_cnt[1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            row[NECI_ST_TAU_GAMMA_SING + c] = __longlong_as_double((long long)S.tau_gamma[c]);
            row[NECI_ST_TAU_CNT_SING + c] = (double)S.tau_cnt[c];
        }
        if (with_stage_a) {
            row[NECI_ST_NODIED] = t[W_NODIED]; row[NECI_ST_NOABORTED] = t[W_ABORT]; row[NECI_ST_HFCYC] = t[W_HF];
            row[NECI_ST_NOATDOUBS] = t[W_DOUBS]; row[NECI_ST_ENUMCYC] = t[W_ENUM]; row[NECI_ST_ENUMCYCABS] = t[W_ENUMABS];
            row[NECI_ST_INITSENUMCYC] = t[W_INITSENUM]; row[NECI_ST_NOINITDETS] = t[W_INITD];
            row[NECI_ST_NONONINITDETS] = t[W_NINITD]; row[NECI_ST_NOINITWALK] = t[W_INITW];
            row[NECI_ST_NONONINITWALK] = t[W_NINITW]; row[NECI_ST_NOADDEDINITIATORS] = t[W_ADDED];
        }
        if (S.bloom_max[0]) atomicMax((unsigned long long *)&L.ctr[C_COUNT - 2], S.bloom_max[0]);
        if (S.bloom_max[1]) atomicMax((unsigned long long *)&L.ctr[C_COUNT - 1], S.bloom_max[1]);
    }
}

// a statistic of stage A summed over the warp into the warp's row -- only when some lane has something to add
// (the sums are taken in chunk order by one lane, so they are reproducible run to run)
__device__ __forceinline__ void k1_add_stat(double *row, int k, double v) {
    if (__any_sync(0xffffffffu, v != 0.0)) {
        const double t = warp_sum(v);
        if ((threadIdx.x & 31) == 0) row[k] += t;
    }
}

// stage A of one chunk: flags, energy sums, death and attempt counts of 32 x SPT slots; leaves the parents and the
// exclusive prefix sum of their attempt counts in the warp's shared memory.  nsp_k / off_k: this lane's slots.
template <int NW, int SYS>
__device__ __forceinline__ void k1_stage_a(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A, K1Shared<NW> &S,
                                           K1Warp<NW> &W, const Det<NW> &ref, long long chunk, long long n_list, int &T) {
    constexpr int SPT = k1_spt<NW>();
    constexpr int CHUNK = 32 * SPT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double c_died = 0.0, c_bornd = 0.0, c_abort = 0.0, c_hf = 0.0, c_doubs = 0.0, c_enum = 0.0, c_enumabs = 0.0,
           c_initsenum = 0.0, c_initw = 0.0, c_ninitw = 0.0;
    int c_initd = 0, c_ninitd = 0, c_added = 0;
    // the slots of a lane are taken two at a time: all streams of both are requested before anything is consumed (one
    // HBM round trip per pair; four slots at once cost 36 live registers and spilled)
    constexpr int G = (SPT >= 2) ? 2 : 1;
#pragma unroll 1
    for (int k0 = 0; k0 < SPT; k0 += G) {
    double ld_s[G], ld_K[G], ld_O[G]; int ld_f[G]; Det<NW> ld_d[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const long long slot = chunk * CHUNK + (k0 + g) * 32 + lane;
        ld_s[g] = 0.0; ld_K[g] = 0.0; ld_O[g] = 0.0; ld_f[g] = 0; ld_d[g].w[0] = 0; if (NW > 1) ld_d[g].w[NW - 1] = 0;
        if (slot < n_list) {
            ld_s[g] = __ldcs(&L.sgn[slot]); ld_d[g].w[0] = __ldcs(&L.det0[slot]);
            if (NW > 1) ld_d[g].w[NW - 1] = __ldcs(&L.det1[slot]);
            ld_f[g] = __ldcs(&L.flg[slot]); ld_K[g] = __ldcs(&L.diagH[slot]); ld_O[g] = __ldcs(&L.offH[slot]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int kk = k0 + g;
        const int idx = kk * 32 + lane;
        const long long slot = chunk * CHUNK + idx;
        int nsp = 0;
        unsigned char info = 0;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        u64 h = 0;
        const double s = ld_s[g];
        if (slot < n_list && fabs(s) >= 1.0e-12) {
            d = ld_d[g];
            int f = ld_f[g];
            const int f0 = f;
            const double K = ld_K[g], O = ld_O[g];
            const bool core = (f & F_DETERM) != 0;
            const int exl = excit_level_ref<NW, sys_hphf(SYS)>(ref, d);        // FindBitExcitLevel(..., t_hphf_ic = .true.)
            const double as = fabs(s);
            // CalcParentFlag / TestInitiator_explicit (fcimc_helper.F90:1036-1243)
            if (P.t_trunc_initiator) {
                const bool was = (f & F_INIT) != 0;
                const bool initiator = parent_is_initiator(P, was, as, exl, core);
                if (initiator != was) c_added += initiator ? 1 : -1;
                if (initiator) { c_initd += 1; c_initw += as; f |= F_INIT; }
                else { c_ninitd += 1; c_ninitw += as; f &= ~F_INIT; }
            }
            // SumEContrib (fcimc_helper.F90:518-802)
            if (exl == 0) c_hf += s;
            if (exl == 2) c_doubs += as;
            const double dE = O * s;
            c_enum += dE; c_enumabs += fabs(dE);
            if (f & F_INIT) c_initsenum += dE;
            h = det_hash64(d);
            // decide_num_to_spawn (fcimc_helper.F90:2160-2174)
            {
                const double x = s * P.av_mc_excits;
                nsp = abs((int)x);
                if (fabs(fabs(x) - (double)nsp) > 1.e-12) {
                    Stream rng(P.seed, A.iter, h, 0, RNG_NSPAWN);
                    if ((fabs(x) - (double)nsp) > rng.draw53()) ++nsp;
                }
            }
            info = (unsigned char)((s < 0.0 ? 1 : 0) | ((f & F_INIT) ? 2 : 0) | (core ? 4 : 0));
            // walker_death / attempt_die_normal (fcimc_helper.F90:2279-2407, fcimc_pointed_fns.F90:573-705)
            // tDeathBeforeComms: here with t_core_die_ = .false. (FciMCPar.F90:1752-1756); otherwise
            // perform_death_all_walkers (fcimc_helper.F90:2253-2277) would run it after the loop for every
            // determinant, core ones included -- death of slot j touches only slot j, so it is fused here too
            double news = s;
            if (!core || !P.t_death_before_comms) {
                const double fac = A.tau * (K - A.diag_sft);
                if (fac > 2.0) atomicOr((unsigned long long *)&L.ctr[C_ERR], 4ull);
                double iDie;
                if (P.t_all_real_coeff) iDie = fac * as;
                else {
                    double rat = fac * as;
                    iDie = (double)(long long)rat;
                    rat = rat - iDie;
                    Stream rng(P.seed, A.iter, h, 0, RNG_DEATH);
                    if (fabs(rat) > rng.draw53()) iDie += (rat < 0.0 || (rat == 0.0 && signbit(rat))) ? -1.0 : 1.0;
                }
                c_died += fmin(iDie, as);
                c_bornd += fmax(iDie - as, 0.0);
                news = s - (iDie * dsign(1.0, s));
                if (P.t_trunc_initiator && fabs(news) > 1.0e-12 && ((news > 0.0) != (s > 0.0))) {
                    c_abort += fabs(news);
                    if (f & F_INIT) c_added -= 1;
                    news = 0.0;
                }
                if (!(fabs(news) > 1.0e-12) && !core) {
                    if (P.t_trunc_initiator && (f & F_INIT)) c_added -= 1;
                    ht_remove<NW>(L, d, h, slot);
                    f |= F_REMOVED;
                    news = 0.0;
                }
            }
            if (news != s) L.sgn[slot] = news;
            if (f != f0) L.flg[slot] = f;
            if (nsp > NG_HEAVY) {
                const long long k = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NHEAVY], 1ull);
                if (k < SB.heavy_cap) { SB.heavy[2 * k] = slot; SB.heavy[2 * k + 1] = ((long long)nsp << 8) | info; }
                else atomicOr((unsigned long long *)&L.ctr[C_ERR], 32ull);
                nsp = 0;
            }
        }
        W.p_d0[idx] = d.w[0]; if (NW > 1) W.p_d1[idx] = d.w[NW - 1];
        W.p_h[idx] = h; W.p_info[idx] = info;
        W.p_off[idx] = nsp;                              // attempt count; turned into the prefix sum below
    }
    }
    // statistics of this chunk into the warp's row
    {
        double *row = S.wacc[warp];
        k1_add_stat(row, W_NODIED, c_died); k1_add_stat(row, W_NOBORN_D, c_bornd); k1_add_stat(row, W_ABORT, c_abort);
        k1_add_stat(row, W_HF, c_hf); k1_add_stat(row, W_DOUBS, c_doubs); k1_add_stat(row, W_ENUM, c_enum);
        k1_add_stat(row, W_ENUMABS, c_enumabs); k1_add_stat(row, W_INITSENUM, c_initsenum);
        if (P.t_trunc_initiator) {
            k1_add_stat(row, W_INITW, c_initw); k1_add_stat(row, W_NINITW, c_ninitw);
            const int a = __reduce_add_sync(0xffffffffu, c_initd), b = __reduce_add_sync(0xffffffffu, c_ninitd),
                      c = __reduce_add_sync(0xffffffffu, c_added);
            if (lane == 0) { row[W_INITD] += (double)a; row[W_NINITD] += (double)b; row[W_ADDED] += (double)c; }
        }
    }
    // exclusive prefix sum of the attempt counts over the chunk (index order kk * 32 + lane)
    int run = 0;
#pragma unroll
    for (int kk = 0; kk < SPT; ++kk) {
        const int mine = W.p_off[kk * 32 + lane];        // written by this lane above
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        W.p_off[kk * 32 + lane] = run + incl - mine;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    T = run;
    if (lane == 0) W.p_off[CHUNK] = run;
    __syncwarp();
}

template <int NW, int SYS>
__global__ void __launch_bounds__(K1_BLOCK, K1_CTAS_PER_SM) k_spawn(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Shared<NW> &S = *reinterpret_cast<K1Shared<NW> *>(k1_smem);
    constexpr int SPT = k1_spt<NW>();
    constexpr int CHUNK = 32 * SPT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K1Warp<NW> &W = S.w[warp];
    k1_init_shared<NW>(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    K1Queues Q; Q.qe = 0; Q.qs = 0;
    const Det<NW> ref = ref_det<NW>(P);
    const long long n_list = L.ctr[C_NLIST];
    __syncthreads();

    // Chunks are dealt to the warps of the grid round-robin (static, so every sum is taken in the same order every run).
    // One loop of rounds with a single call site for the attempt stages (they are the bulk of the kernel's code, and
    // every additional inlined copy costs instruction-cache misses): a round generates up to 32 attempts of the
    // current window of the current chunk and serves the queues; when the window is exhausted the next window's map
    // is filled, when the chunk is exhausted the next chunk is loaded (stage A); after the last chunk one draining
    // round ends the loop.
    const long long n_chunks = (n_list + CHUNK - 1) / CHUNK;
    const long long gwarp = (long long)blockIdx.x * K1_WARPS + warp, nwarps = (long long)gridDim.x * K1_WARPS;
    long long chunk = gwarp;
    int T = 0, wb = 0, we = 0, base = 0;
    for (;;) {
        bool drain = false;
        if (base >= we) {
            if (we >= T) {                                         // chunk exhausted
                if (chunk >= n_chunks) drain = true;
                else {
                    k1_stage_a<NW, SYS>(P, L, SB, A, S, W, ref, chunk, n_list, T);
                    chunk += nwarps;
                    wb = 0; we = 0; base = 0;
                    if (T == 0) continue;
                }
            } else { __syncwarp(); wb = we; }                      // next window: the map is overwritten
            if (!drain) {
                // Attempts are numbered 0..T-1 over the chunk; each parent writes its chunk index into the map entries
                // of its own attempts (window by window): an attempt finds its parent with one load.
                we = min(T, wb + K1_MAPW); base = wb;
#pragma unroll
                for (int kk = 0; kk < SPT; ++kk) {
                    const int idx = kk * 32 + lane;
                    const int lo = max(W.p_off[idx], wb), hi = min(W.p_off[idx + 1], we);
                    const bool big = hi - lo > 4;
                    if (!big) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) if (lo + u < hi) W.map[lo + u - wb] = (unsigned char)idx;
                    }
                    u32 m = __ballot_sync(0xffffffffu, big);       // long ranges are filled by the whole warp
                    while (m) {
                        const int src = __ffs(m) - 1; m &= m - 1u;
                        const int l2 = __shfl_sync(0xffffffffu, lo, src), h2 = __shfl_sync(0xffffffffu, hi, src);
                        const int i2 = __shfl_sync(0xffffffffu, idx, src);
#pragma unroll 1
                        for (int a = l2 + lane; a < h2; a += 32) W.map[a - wb] = (unsigned char)i2;
                    }
                }
                __syncwarp();
            }
        }
        {
            const int a = base + lane;
            const bool active = !drain && a < we;
            Det<NW> dp; dp.w[0] = 0; if (NW > 1) dp.w[NW - 1] = 0;
            u64 h = 0; int info = 0; u32 p = 0;
            if (active) {
                const int lo = W.map[a - wb];
                dp.w[0] = W.p_d0[lo]; if (NW > 1) dp.w[NW - 1] = W.p_d1[lo];
                h = W.p_h[lo]; info = W.p_info[lo]; p = (u32)(a - W.p_off[lo]);
            }
            if (!drain) { stage_generate<NW, SYS>(P, L, A, W, Q, active, dp, h, info, p, acc); base += 32; }
        }
        serve_queues<NW, SYS>(P, L, SB, A, S, W, Q, drain ? 1 : 32, acc);
        if (drain) break;
    }
    k1_flush<NW>(L, S, acc, partials, true);
}

// Attempts of the deferred heavy determinants (> NG_HEAVY walkers): rounds of 32 attempts dealt to all warps of the grid.
template <int NW, int SYS>
__global__ void __launch_bounds__(K1_BLOCK, K1_CTAS_PER_SM) k_spawn_heavy(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Shared<NW> &S = *reinterpret_cast<K1Shared<NW> *>(k1_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K1Warp<NW> &W = S.w[warp];
    k1_init_shared<NW>(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    K1Queues Q; Q.qe = 0; Q.qs = 0;
    __syncthreads();
    long long nh = L.ctr[C_NHEAVY];
    if (nh > SB.heavy_cap) nh = SB.heavy_cap;
    const long long gwarp = (long long)blockIdx.x * K1_WARPS + warp, nwarps = (long long)gridDim.x * K1_WARPS;
    // same shape as k_spawn's loop: one call site for the attempt stages, a last draining round
    long long e = 0, rd = gwarp, rounds = 0;
    Det<NW> dp; dp.w[0] = 0; if (NW > 1) dp.w[NW - 1] = 0;
    u64 h = 0; int nsp = 0, info = 0;
    bool loaded = false;
    for (;;) {
        bool drain = false;
        while (!loaded || rd >= rounds) {                          // next heavy determinant with a round for this warp
            if (loaded) { ++e; rd = gwarp; loaded = false; }
            if (e >= nh) { drain = true; break; }
            const long long slot = SB.heavy[2 * e];
            const long long packed = SB.heavy[2 * e + 1];
            nsp = (int)(packed >> 8); info = (int)(packed & 0xff);
            dp = load_det<NW>(L, slot);
            h = det_hash64(dp);
            rounds = (nsp + 31) / 32;
            loaded = true;
        }
        if (!drain) {
            const int a = (int)(rd * 32) + lane;
            stage_generate<NW, SYS>(P, L, A, W, Q, a < nsp, dp, h, info, (u32)a, acc);
            rd += nwarps;
        }
        serve_queues<NW, SYS>(P, L, SB, A, S, W, Q, drain ? 1 : 32, acc);
        if (drain) break;
    }
    k1_flush<NW>(L, S, acc, partials, false);
}

}  // namespace ng
