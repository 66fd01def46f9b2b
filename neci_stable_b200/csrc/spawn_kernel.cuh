// K1: the loop over determinants of PerformFCIMCycPar (src/FciMCPar.F90:1294-1758), fused:
//   CalcParentFlag (fcimc_helper.F90:1036-1243), SumEContrib (:518-802), decide_num_to_spawn (:2160-2174),
//   generate_excitation + attempt_create_normal (fcimc_pointed_fns.F90:178-491), create_particle
//   (fcimc_helper.F90:152-308), walker_death / attempt_die_normal (:2279-2407, fcimc_pointed_fns.F90:573-705).
//
// B200 design.  Walkers per determinant vary from 1 to 1e5+, a third of the PCHB draws are null
// excitations and a tenth are singles with an O(nel) matrix element, so "one thread walks one
// determinant" leaves two thirds of every warp idle (measured: 10.9 of 32 lanes active, profiles/).
// The kernel is therefore a persistent CTA that moves work between *stages* through shared-memory
// queues, so that every stage runs with full warps:
//
//   stage A  one thread per slot of a 512-slot tile: flags, energy sums, death, attempt count;
//            parents (det, stream id, info) and the prefix sum of the attempt counts go to shared memory
//   stage B1 one thread per attempt (parents expand their index into an attempt -> parent map): draw the excitation.
//            valid doubles / lattice excitations -> queue QE {parent det, orbitals, pgen, rounding draw}
//            PCHB singles -> queue QS {parent det, stream id, attempt}        null draws stop here
//   stage B2 whenever QE holds >= 256 entries: parity (popc), matrix element (2 UMAT loads), spawn
//            weight, stochastic rounding, warp-aggregated append to the destination rank's segment
//   stage B3 whenever QS holds >= 256 entries: uniform single + sltcnd_1 (loads batched 4 at a time), then as B2
//
// The kernel body is one loop of rounds (serve the queues, then generate up to 256 attempts) with a single
// queue-serving call site.  On several ranks the appends of B2/B3 go to a staging list and k_partition_push routes
// them afterwards (kernels.cuh).  HPHF runs are their own compile-time variant (NG_SYS_PCHB_HPHF).
//
// Random numbers are counter-based (device_common.cuh: Stream), so the result does not depend on the
// order in which the queues are served.
#pragma once
#include "device_system.cuh"

namespace ng {

#define NG_BLOCK 256         /* block size of the streaming kernels */
#ifndef K1_BLOCK
#define K1_BLOCK 256         /* block size of the spawning kernel (measured: 128 -> 0.87 ms, 256 -> 0.78 ms, 512 -> 0.80 ms per launch) */
#endif
#ifndef K1_CTAS_PER_SM       /* launch bound: resident CTAs per SM the register allocation must allow (variant builds: _build.build_gpu_variant) */
#define K1_CTAS_PER_SM (1024 / K1_BLOCK)
#endif
#define NG_HEAVY 4096        /* attempts per determinant handled inside a tile */
#ifndef K1_SPT
#define K1_SPT 2             /* slots per thread and tile */
#endif
#define K1_TILE (K1_BLOCK * K1_SPT)
#define K1_QCAP (2 * K1_BLOCK)
#define K1_MAPW 1024         /* attempts per window of the attempt -> parent map */

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
enum { W_STAGE_A = W_CHILD };   // [0, W_STAGE_A) are flushed once per tile, the rest once per kernel

template <int NW> struct K1Shared {
    // parents of the current tile
    u64 p_d0[K1_TILE];
    u64 p_d1[(NW > 1) ? K1_TILE : 1];
    u64 p_h[K1_TILE];
    int p_off[K1_TILE + 1];
    unsigned short p_map[K1_MAPW];   // attempt (within the current window) -> parent index in the tile
    unsigned char p_info[K1_TILE];
    // QE: generated excitations waiting for their matrix element
    u64 q_d0[K1_QCAP];
    u64 q_d1[(NW > 1) ? K1_QCAP : 1];
    double q_pgen[K1_QCAP];
    double q_r[K1_QCAP];
    u32 q_orbs[K1_QCAP];     // src1 | src2 << 8 | tgt1 << 16 | tgt2 << 24
    u32 q_misc[K1_QCAP];     // info | ic << 8
    // QS: PCHB single excitations still to be generated
    u64 s_d0[K1_QCAP];
    u64 s_d1[(NW > 1) ? K1_QCAP : 1];
    u64 s_h[K1_QCAP];
    u32 s_att[K1_QCAP];
    u32 s_misc[K1_QCAP];
    int roi[NG_MAX_BASIS];
    double wacc[K1_BLOCK / 32][W_COUNT];
    int wsum[K1_BLOCK / 32];
    int q_count, s_count;
    int bloom_cnt[2];
    unsigned long long bloom_max[2];
    // tau search (log_spawn_magnitude): classes 0 singles, 1 doubles, 2 parallel doubles, 3 opposite-spin doubles
    int tau_cnt[4];
    unsigned long long tau_gamma[4];
};

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
// to a staging list first and k_partition routes it afterwards: DetermineDetNode costs ~230 instructions, and inside
// the spawning kernel only ~5 of 32 lanes hold a successful spawn, so hashing there wasted 85 % of the issue slots
// it used (11 % of the kernel); the partition kernel hashes with every lane busy.
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

// attempt_create_normal (fcimc_pointed_fns.F90:178-491) for an excitation whose orbitals, pgen and rounding
// draw are known, followed by create_particle.  `active` lanes hold a real entry; all lanes must call.
template <int NW, int SYS>
__device__ __forceinline__ void evaluate_and_append(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                                    K1Shared<NW> &S, bool active, const Det<NW> &d, Excit<NW> &E, int info,
                                                    double r_round, AttAcc &acc) {
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
                if (P.t_real_spawn_cutoff && fabs(nSpawn) < P.real_spawn_cutoff)
                    nSpawn = P.real_spawn_cutoff * stochastic_round_r(nSpawn / P.real_spawn_cutoff, r_round);
            } else nSpawn = stochastic_round_r(nSpawn, r_round);
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

// push helpers: warp-aggregated reservation in a shared-memory stack
__device__ __forceinline__ int queue_reserve(int *count, bool push) {
    const u32 lane = threadIdx.x & 31;
    const u32 m = __ballot_sync(0xffffffffu, push);
    if (m == 0) return -1;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return push ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// stage B1: one spawning attempt of parent (d, h, info), attempt index p
template <int NW, int SYS>
__device__ __forceinline__ void stage_generate(const Params &P, const WalkerList &L, const IterArgs &A, K1Shared<NW> &S,
                                               bool active, const Det<NW> &d, u64 h, int info, u32 p, AttAcc &acc) {
    bool push_e = false, push_s = false;
    Excit<NW> E;
    double r_round = 0.0;
    if (active) {
        Stream rng(P.seed, A.iter, h, p, RNG_ATTEMPT);
        if (sys_pchb(SYS)) {
            if (rng.draw() < P.p_singles) push_s = true;                 // gen_exc_sd: single, generated in stage B3
            else { gen_pchb_double(P, d, rng, E); E.pgen = E.pgen * P.p_doubles; }
        } else generate_excitation_core<NW, SYS>(P, d, rng, E);
        if (!push_s) {
            if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
            if (E.valid) { acc.valid += 1; r_round = rng.draw(); push_e = true; }
            else acc.invalid += 1;
        }
    }
    const int qe = queue_reserve(&S.q_count, push_e);
    if (push_e) {
        S.q_d0[qe] = d.w[0]; if (NW > 1) S.q_d1[qe] = d.w[NW - 1];
        S.q_pgen[qe] = E.pgen; S.q_r[qe] = r_round;
        S.q_orbs[qe] = (u32)E.src1 | ((u32)E.src2 << 8) | ((u32)E.tgt1 << 16) | ((u32)E.tgt2 << 24);
        S.q_misc[qe] = (u32)info | ((u32)E.ic << 8);
    }
    if (sys_pchb(SYS)) {
        const int qs = queue_reserve(&S.s_count, push_s);
        if (push_s) {
            S.s_d0[qs] = d.w[0]; if (NW > 1) S.s_d1[qs] = d.w[NW - 1];
            S.s_h[qs] = h; S.s_att[qs] = p; S.s_misc[qs] = (u32)info;
        }
    }
}

// stage B2: serve up to 256 entries from the top of QE
template <int NW, int SYS>
__device__ __forceinline__ void stage_evaluate(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                               K1Shared<NW> &S, int first, int n, AttAcc &acc) {
    const int i = first + threadIdx.x;
    const bool active = (int)threadIdx.x < n;
    Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
    Excit<NW> E; E.ic = 2; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = active; E.err = 0; E.detJ = d; E.parity = false;
    int info = 0; double r = 0.0;
    if (active) {
        d.w[0] = S.q_d0[i]; if (NW > 1) d.w[NW - 1] = S.q_d1[i];
        const u32 o = S.q_orbs[i], m = S.q_misc[i];
        E.src1 = o & 0xff; E.src2 = (o >> 8) & 0xff; E.tgt1 = (o >> 16) & 0xff; E.tgt2 = o >> 24;
        E.ic = (m >> 8) & 0xff; info = m & 0xff;
        E.pgen = S.q_pgen[i]; r = S.q_r[i];
    }
    evaluate_and_append<NW, SYS>(P, L, SB, A, S, active, d, E, info, r, acc);
}

// stage B3 (PCHB only): serve up to 256 deferred singles from the top of QS
template <int NW, int SYS>
__device__ __forceinline__ void stage_singles(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                              K1Shared<NW> &S, int first, int n, AttAcc &acc) {
    const int i = first + threadIdx.x;
    bool active = (int)threadIdx.x < n;
    Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
    Excit<NW> E; E.ic = 1; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = false; E.err = 0; E.detJ = d; E.parity = false;
    int info = 0; double r = 0.0;
    if (active) {
        d.w[0] = S.s_d0[i]; if (NW > 1) d.w[NW - 1] = S.s_d1[i];
        info = (int)S.s_misc[i];
        Stream rng(P.seed, A.iter, S.s_h[i], S.s_att[i], RNG_ATTEMPT, 1);      // draw 0 chose "single"
        gen_uniform_single(P, d, rng, E);
        E.pgen = E.pgen * P.p_singles;
        if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
        if (E.valid) { acc.valid += 1; r = rng.draw(); }
        else { acc.invalid += 1; active = false; }
    }
    evaluate_and_append<NW, SYS>(P, L, SB, A, S, active, d, E, info, r, acc);
}

// serve the queues while they hold at least `level` entries (level = 256 keeps every stage at full width,
// level = 1 drains).  Must be called by the whole CTA; ends with the counters published.
template <int NW, int SYS>
__device__ __forceinline__ void serve_queues(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                             K1Shared<NW> &S, int level, AttAcc &acc) {
    __syncthreads();
    for (;;) {
        const int qc = S.q_count;
        if (qc < level) break;
        const int n = min(qc, K1_BLOCK);
        __syncthreads();
        if (threadIdx.x == 0) S.q_count = qc - n;
        stage_evaluate<NW, SYS>(P, L, SB, A, S, qc - n, n, acc);
        __syncthreads();
    }
    if (sys_pchb(SYS)) {
        for (;;) {
            const int sc = S.s_count;
            if (sc < level) break;
            const int n = min(sc, K1_BLOCK);
            __syncthreads();
            if (threadIdx.x == 0) S.s_count = sc - n;
            stage_singles<NW, SYS>(P, L, SB, A, S, sc - n, n, acc);
            __syncthreads();
        }
    }
}

template <int NW>
__device__ __forceinline__ void k1_init_shared(const Params &P, K1Shared<NW> &S) {
#pragma unroll 1
    for (int i = threadIdx.x; i < P.nbasis; i += K1_BLOCK) S.roi[i] = P.random_orb_index[i];
    for (int i = threadIdx.x; i < (K1_BLOCK / 32) * W_COUNT; i += K1_BLOCK) (&S.wacc[0][0])[i] = 0.0;
    if (threadIdx.x == 0) { S.q_count = 0; S.s_count = 0; S.bloom_cnt[0] = S.bloom_cnt[1] = 0; S.bloom_max[0] = S.bloom_max[1] = 0ull; }
    if (threadIdx.x < 4) { S.tau_cnt[threadIdx.x] = 0; S.tau_gamma[threadIdx.x] = 0ull; }
}

// end of kernel: per-thread attempt accumulators -> per-warp rows -> one partial row per CTA
template <int NW>
__device__ __forceinline__ void k1_flush(const WalkerList &L, K1Shared<NW> &S, const AttAcc &acc, double *partials, bool with_stage_a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double c = warp_sum(acc.child), cs = warp_sum(acc.child_sing), v = warp_sum((double)acc.valid),
                 iv = warp_sum((double)acc.invalid), mx = warp_max(acc.maxsp);
    if (lane == 0) {
        S.wacc[warp][W_CHILD] = c; S.wacc[warp][W_CHILD_SING] = cs; S.wacc[warp][W_VALID] = v; S.wacc[warp][W_INVALID] = iv;
        S.wacc[warp][W_MAXSP] = mx;
    }
    double *row = partials + (size_t)blockIdx.x * NECI_ST_COUNT;
    for (int k = threadIdx.x; k < NECI_ST_COUNT; k += K1_BLOCK) row[k] = 0.0;
    __syncthreads();
    if (threadIdx.x < W_COUNT) {
        const int k = threadIdx.x;
        double t = S.wacc[0][k];
        for (int w = 1; w < K1_BLOCK / 32; ++w) t = (k == W_MAXSP) ? fmax(t, S.wacc[w][k]) : t + S.wacc[w][k];
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

// stage A of one tile: flags, energy sums, death and attempt counts of 512 slots; leaves the parents and the
// exclusive prefix sum of their attempt counts in shared memory.  nsp_k / off_k: this thread's two slots.
template <int NW, int SYS>
__device__ __forceinline__ void k1_stage_a(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A, K1Shared<NW> &S,
                                           const Det<NW> &ref, long long tile, long long n_list, int (&nsp_k)[K1_SPT],
                                           int (&off_k)[K1_SPT], int &T) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        // ---------------- stage A: one thread per slot --------------------------------------------
        double sa[W_STAGE_A];
#pragma unroll
        for (int k = 0; k < W_STAGE_A; ++k) sa[k] = 0.0;
        // all five streams of both slots are requested before anything is consumed: one HBM round trip per
        // tile (empty slots cost bandwidth, which this kernel has to spare, not latency)
        double ld_s[K1_SPT], ld_K[K1_SPT], ld_O[K1_SPT]; int ld_f[K1_SPT]; Det<NW> ld_d[K1_SPT];
#pragma unroll
        for (int kk = 0; kk < K1_SPT; ++kk) {
            const long long slot = tile * K1_TILE + kk * K1_BLOCK + tid;
            ld_s[kk] = 0.0; ld_K[kk] = 0.0; ld_O[kk] = 0.0; ld_f[kk] = 0; ld_d[kk].w[0] = 0; if (NW > 1) ld_d[kk].w[NW - 1] = 0;
            if (slot < n_list) {
                ld_s[kk] = __ldcs(&L.sgn[slot]); ld_d[kk].w[0] = __ldcs(&L.det0[slot]);
                if (NW > 1) ld_d[kk].w[NW - 1] = __ldcs(&L.det1[slot]);
                ld_f[kk] = __ldcs(&L.flg[slot]); ld_K[kk] = __ldcs(&L.diagH[slot]); ld_O[kk] = __ldcs(&L.offH[slot]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < K1_SPT; ++kk) {
            const int idx = kk * K1_BLOCK + tid;
            const long long slot = tile * K1_TILE + idx;
            int nsp = 0;
            unsigned char info = 0;
            Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
            u64 h = 0;
            if (slot < n_list) {
                const double s = ld_s[kk];
                if (fabs(s) >= 1.0e-12) {
                    d = ld_d[kk];
                    int f = ld_f[kk];
                    const int f0 = f;
                    const double K = ld_K[kk], O = ld_O[kk];
                    const bool core = (f & F_DETERM) != 0;
                    const int exl = excit_level_ref<NW, sys_hphf(SYS)>(ref, d);        // FindBitExcitLevel(..., t_hphf_ic = .true.)
                    const double as = fabs(s);
                    // CalcParentFlag / TestInitiator_explicit (fcimc_helper.F90:1036-1243)
                    if (P.t_trunc_initiator) {
                        const bool was = (f & F_INIT) != 0;
                        const bool initiator = parent_is_initiator(P, was, as, exl, core);
                        if (initiator != was) sa[W_ADDED] += initiator ? 1.0 : -1.0;
                        if (initiator) { sa[W_INITD] += 1.0; sa[W_INITW] += as; f |= F_INIT; }
                        else { sa[W_NINITD] += 1.0; sa[W_NINITW] += as; f &= ~F_INIT; }
                    }
                    // SumEContrib (fcimc_helper.F90:518-802)
                    if (exl == 0) sa[W_HF] += s;
                    if (exl == 2) sa[W_DOUBS] += as;
                    const double dE = O * s;
                    sa[W_ENUM] += dE; sa[W_ENUMABS] += fabs(dE);
                    if (f & F_INIT) sa[W_INITSENUM] += dE;
                    h = det_hash64(d);
                    // decide_num_to_spawn (fcimc_helper.F90:2160-2174)
                    {
                        const double x = s * P.av_mc_excits;
                        nsp = abs((int)x);
                        if (fabs(fabs(x) - (double)nsp) > 1.e-12) {
                            Stream rng(P.seed, A.iter, h, 0, RNG_NSPAWN);
                            if ((fabs(x) - (double)nsp) > rng.draw()) ++nsp;
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
                            if (fabs(rat) > rng.draw()) iDie += (rat < 0.0 || (rat == 0.0 && signbit(rat))) ? -1.0 : 1.0;
                        }
                        sa[W_NODIED] += fmin(iDie, as);
                        sa[W_NOBORN_D] += fmax(iDie - as, 0.0);
                        news = s - (iDie * dsign(1.0, s));
                        if (P.t_trunc_initiator && fabs(news) > 1.0e-12 && ((news > 0.0) != (s > 0.0))) {
                            sa[W_ABORT] += fabs(news);
                            if (f & F_INIT) sa[W_ADDED] -= 1.0;
                            news = 0.0;
                        }
                        if (!(fabs(news) > 1.0e-12) && !core) {
                            if (P.t_trunc_initiator && (f & F_INIT)) sa[W_ADDED] -= 1.0;
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
            }
            S.p_d0[idx] = d.w[0]; if (NW > 1) S.p_d1[idx] = d.w[NW - 1];
            S.p_h[idx] = h; S.p_info[idx] = info;
            nsp_k[kk] = nsp;
        }
        // per-tile flush of the stage-A sums into this warp's row (fixed order => reproducible sums)
#pragma unroll
        for (int k = 0; k < W_STAGE_A; ++k) {
            const double t = warp_sum(sa[k]);
            if (lane == 0) S.wacc[warp][k] += t;
        }
        // exclusive prefix sum of the attempt counts over the tile (index order kk * 256 + tid)
        int run = 0;
#pragma unroll
        for (int kk = 0; kk < K1_SPT; ++kk) {
            int incl = nsp_k[kk];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) S.wsum[warp] = incl;
            __syncthreads();
            int wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < K1_BLOCK / 32; ++w) { const int v = S.wsum[w]; if (w < warp) wbase += v; total += v; }
            off_k[kk] = run + wbase + incl - nsp_k[kk];
            S.p_off[kk * K1_BLOCK + tid] = off_k[kk];
            run += total;
            __syncthreads();
        }
        T = run;
}

template <int NW, int SYS>
__global__ void __launch_bounds__(K1_BLOCK, K1_CTAS_PER_SM) k_spawn(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Shared<NW> &S = *reinterpret_cast<K1Shared<NW> *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    k1_init_shared<NW>(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    const Det<NW> ref = ref_det<NW>(P);
    const long long n_list = L.ctr[C_NLIST];
    __syncthreads();

    // One loop of rounds: a round serves the queues (single call site: the evaluate / singles stages are the bulk
    // of the kernel's code, and instruction-cache misses showed up in the profile when they were inlined twice) and
    // then generates up to 256 attempts of the current window of the current tile.  When the tile's attempts are
    // exhausted the next tile is loaded (stage A); after the last tile one draining round ends the kernel.
    long long tile = blockIdx.x;
    int T = 0, wb = 0, we = 0, base = 0;
    int nsp_k[K1_SPT], off_k[K1_SPT];
#pragma unroll
    for (int kk = 0; kk < K1_SPT; ++kk) { nsp_k[kk] = 0; off_k[kk] = 0; }
    for (;;) {
        bool drain = false;
        if (base >= we) {
            bool new_window = false;
            if (we >= T) {
                if (tile * K1_TILE >= n_list) drain = true;
                else {
                    __syncthreads();        // parents and the map are overwritten
                    k1_stage_a<NW, SYS>(P, L, SB, A, S, ref, tile, n_list, nsp_k, off_k, T);
                    tile += gridDim.x;
                    wb = 0; we = 0; base = 0;
                    if (T == 0) continue;
                    new_window = true;
                }
            } else { __syncthreads(); wb = we; new_window = true; }
            if (new_window) {
                // Attempts are numbered 0..T-1 over the tile; each parent writes its tile index into the map
                // entries of its own attempts (window by window): an attempt finds its parent with one load.
                we = min(T, wb + K1_MAPW); base = wb;
#pragma unroll
                for (int kk = 0; kk < K1_SPT; ++kk) {
                    const int idx = kk * K1_BLOCK + tid;
                    const int lo = max(off_k[kk], wb), hi = min(off_k[kk] + nsp_k[kk], we);
                    const bool big = hi - lo > 4;
                    if (!big) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) if (lo + u < hi) S.p_map[lo + u - wb] = (unsigned short)idx;
                    }
                    u32 m = __ballot_sync(0xffffffffu, big);           // long ranges are filled by the whole warp
                    while (m) {
                        const int src = __ffs(m) - 1; m &= m - 1u;
                        const int l2 = __shfl_sync(0xffffffffu, lo, src), h2 = __shfl_sync(0xffffffffu, hi, src);
                        const int i2 = __shfl_sync(0xffffffffu, idx, src);
#pragma unroll 1
                        for (int a = l2 + lane; a < h2; a += 32) S.p_map[a - wb] = (unsigned short)i2;
                    }
                }
            }
        }
        serve_queues<NW, SYS>(P, L, SB, A, S, drain ? 1 : K1_BLOCK, acc);   // starts with a barrier (publishes p_* and the map)
        if (drain) break;
        {
            const int a = base + tid;
            const bool active = a < we;
            Det<NW> dp; dp.w[0] = 0; if (NW > 1) dp.w[NW - 1] = 0;
            u64 h = 0; int info = 0; u32 p = 0;
            if (active) {
                const int lo = S.p_map[a - wb];
                dp.w[0] = S.p_d0[lo]; if (NW > 1) dp.w[NW - 1] = S.p_d1[lo];
                h = S.p_h[lo]; info = S.p_info[lo]; p = (u32)(a - S.p_off[lo]);
            }
            stage_generate<NW, SYS>(P, L, A, S, active, dp, h, info, p, acc);
            base += K1_BLOCK;
        }
    }
    serve_queues<NW, SYS>(P, L, SB, A, S, 1, acc);
    k1_flush<NW>(L, S, acc, partials, true);
}

// Attempts of the deferred heavy determinants (> NG_HEAVY walkers), spread over the whole grid.
template <int NW, int SYS>
__global__ void __launch_bounds__(K1_BLOCK, K1_CTAS_PER_SM) k_spawn_heavy(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    K1Shared<NW> &S = *reinterpret_cast<K1Shared<NW> *>(k1_smem);
    const int tid = threadIdx.x;
    k1_init_shared<NW>(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    __syncthreads();
    long long nh = L.ctr[C_NHEAVY];
    if (nh > SB.heavy_cap) nh = SB.heavy_cap;
    for (long long e = 0; e < nh; ++e) {
        const long long slot = SB.heavy[2 * e];
        const long long packed = SB.heavy[2 * e + 1];
        const int nsp = (int)(packed >> 8);
        const int info = (int)(packed & 0xff);
        const Det<NW> dp = load_det<NW>(L, slot);
        const u64 h = det_hash64(dp);
        const int rounds = (nsp + K1_BLOCK - 1) / K1_BLOCK;
        for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
            serve_queues<NW, SYS>(P, L, SB, A, S, K1_BLOCK, acc);
            const int a = rd * K1_BLOCK + tid;
            stage_generate<NW, SYS>(P, L, A, S, a < nsp, dp, h, info, (u32)a, acc);
        }
    }
    serve_queues<NW, SYS>(P, L, SB, A, S, 1, acc);
    k1_flush<NW>(L, S, acc, partials, false);
}

}  // namespace ng
