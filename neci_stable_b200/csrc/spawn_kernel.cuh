// K1: the loop over determinants of PerformFCIMCycPar (src/FciMCPar.F90:1294-1758):
//   CalcParentFlag (fcimc_helper.F90:1036-1243), SumEContrib (:518-802), decide_num_to_spawn (:2160-2174),
//   generate_excitation + attempt_create_normal (fcimc_pointed_fns.F90:178-491), create_particle
//   (fcimc_helper.F90:152-308), walker_death / attempt_die_normal (:2279-2407, fcimc_pointed_fns.F90:573-705).
//
// B200 design (round 2).  Walkers per determinant vary from 1 to 1e5+, a third of the PCHB draws are null
// excitations and a tenth are singles with an O(nel) matrix element, so attempts -- not determinants -- are the unit
// of parallelism and work moves between *stages* through queues so that every stage runs with full warps.  Round 1
// ran all stages inside one persistent kernel (shared-memory queues, block barriers); its profile and that of a
// warp-autonomous rewrite (profiles/r02_k1_*) show what bounds such a kernel on this chip: 110 KB of code against a
// 32 KB instruction cache (L1.5) -- with the warps of an SM spread over the stages, "no instruction" became the first
// stall reason (46 % of the samples in the warp-autonomous version).  So the stages are now four small kernels, each
// with a code footprint that fits the instruction cache, connected by queues in global memory.  The queues are
// written and read once, coalesced, and what of them does not stay in the 126 MB L2 goes over an HBM link this path
// uses to a few per cent:
//
//   k_walk      one thread per slot of the list (all SoA streams of its slots requested up front): flags, energy
//               sums, death, attempt count; occupied determinants -> parent list {det, attempts, info}
//   k_generate  a CTA takes 512 parents, expands them into attempts (attempt -> parent map in shared memory), one
//               thread per attempt draws the excitation from ONE Philox block.  valid doubles / lattice excitations ->
//               queue QE {parent det, orbitals, pgen, attempt, info}; PCHB singles -> queue QS {parent det, attempt,
//               info}; null draws stop here.  k_generate_heavy does the same for determinants above NG_HEAVY attempts.
//   k_evaluate  one thread per QE entry: parity (popc), matrix element (2 UMAT loads), spawn weight, the rounding draw
//               (its own Philox block), stochastic rounding, warp-aggregated append to the destination's segment
//   k_singles   one thread per QS entry: uniform single + sltcnd_1 (loads batched 4 at a time), then as k_evaluate
//
// Every stage runs with all lanes busy (the queues are compacted), no stage waits for another inside a kernel, and
// each kernel gets the register allocation and occupancy that suit it.  On several ranks every flush of a warp's
// spawn stage routes its records (DetermineDetNode) and stores them into the owners' inboxes over NVLink
// (spawn_stage_push); with the NCCL exchange they go to a staging list that k_partition routes (kernels.cuh).  HPHF
// runs are their own compile-time variant (NG_SYS_PCHB_HPHF).
//
// Random numbers are counter-based (device_common.cuh: Stream), so the result does not depend on the order in which
// queue entries are written or served.
#pragma once
#include "device_system.cuh"

namespace ng {

#define NG_BLOCK 256         /* block size of the streaming kernels */
#ifndef K1_WALK_CTAS
#define K1_WALK_CTAS 4       /* resident CTAs per SM the register allocation of k_walk aims at */
#endif
#ifndef K1_SING_CTAS
#define K1_SING_CTAS 4       /* the same for k_singles */
#endif
#define NG_HEAVY 4096        /* attempts per determinant expanded inside a tile of k_generate */
#define K1_GEN_BLOCK 256     /* threads per CTA of k_generate */
#define K1_GEN_TILE 512      /* parents per tile */
#define K1_MAPW 1024         /* attempts per window of the attempt -> parent map */
#define K1_GEN_TSEG 256      /* tiles per CTA whose segment is looked up ahead of the tile loop */

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
    long long *stage;                 // nranks > 1, NCCL exchange: spawns of the spawning kernels before they are routed (k_partition)
    unsigned long long *stage_cnt;
    long long stage_cap;
    // nranks > 1, peer-memory exchange: the spawning kernels route and push their spawns themselves (spawn_stage_push).
    // push_seg[r] = rank r's inbox mapped here, push_off = first record of this rank's segment there for the
    // current exchange, push_cnt[16 * r] = records reserved for rank r so far (one counter per 128-byte line: the L2
    // atomic unit serialises atomics on one line).  Null otherwise.
    long long *const *push_seg;
    unsigned long long *push_cnt;
    long long push_off;
};
#define NG_PUSH_CNT_STRIDE 16

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

// the queues between the K1 kernels, in global memory
struct K1Queues {
    // parent list: occupied determinants with 1 <= attempts <= NG_HEAVY.  Every CTA of k_walk owns the segment
    // [cta * par_seg_cap, (cta + 1) * par_seg_cap) and leaves its count in par_cnt[cta]: no global atomics at all.
    u64 *par_d0, *par_d1; u32 *par_meta;            // meta = attempts << 8 | info
    u32 *par_cnt; int par_nseg; long long par_seg_cap;
    // QE: generated excitations waiting for their matrix element, records of qe_rec<NW>() words:
    //   det word(s), pgen, orbitals | attempt << 32 | info << 61 (orbitals = src1 | src2 << 8 | tgt1 << 16 | tgt2 << 24;
    //   k_walk refuses more than 2^29 - 1 attempts per determinant)
    u64 *qe;
    // QS: PCHB single excitations still to be generated, records of qs_rec<NW>() words: det word(s), attempt | info << 32
    u64 *qs;
    long long qe_cap, qs_cap;
    unsigned long long *cnt;                        // [Q_NQE], [Q_NQS]
};
enum { Q_NQE = 0, Q_NQS };
template <int NW> __host__ __device__ constexpr int qe_rec() { return NW + 2; }
template <int NW> __host__ __device__ constexpr int qs_rec() { return NW + 1; }
// info bits of a parent: 1 negative sign, 2 initiator, 4 core determinant

// shared-memory scratch of the attempt kernels: bloom and tau-search statistics (rare atomics)
struct AttShared {
    int bloom_cnt[2];
    unsigned long long bloom_max[2];
    // tau search (log_spawn_magnitude): classes 0 singles, 1 doubles, 2 parallel doubles, 3 opposite-spin doubles
    int tau_cnt[4];
    unsigned long long tau_gamma[4];
};

// ---- block-level reduction of per-thread statistics into per-block partials ---
// acc[k] for the statistics listed in idx[k]; writes out[blockIdx.x * NECI_ST_COUNT + idx[k]]
// (all other entries of the block row are zeroed here).  Fixed tree: reproducible for a fixed launch shape.
template <int N>
__device__ __forceinline__ void block_flush_stats(const double (&acc)[N], const int (&idx)[N], double *out, double *s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < NECI_ST_COUNT; k += blockDim.x) out[(size_t)blockIdx.x * NECI_ST_COUNT + k] = 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const bool is_max = (idx[k] >= NECI_ST_FIRST_MAX && idx[k] <= NECI_ST_LAST_MAX) || idx[k] == NECI_ST_HIGHEST_POP;
        const double v = is_max ? warp_max(acc[k]) : warp_sum(acc[k]);
        if (lane == 0) s_red[k * 32 + warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        const int k = threadIdx.x;
        const bool is_max = (idx[k] >= NECI_ST_FIRST_MAX && idx[k] <= NECI_ST_LAST_MAX) || idx[k] == NECI_ST_HIGHEST_POP;
        double v = s_red[k * 32];
        for (int w = 1; w < nw; ++w) v = is_max ? fmax(v, s_red[k * 32 + w]) : v + s_red[k * 32 + w];
        out[(size_t)blockIdx.x * NECI_ST_COUNT + idx[k]] = v;
    }
}

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

// CTA-wide reservation of queue entries.  The L2 atomic unit serialises atomics on one address (~1 ns each): a
// reservation per warp and round made k_generate a 0.5 ms queue for its two counters.  Here the warps of a CTA post
// their counts in shared memory and ONE thread reserves for all of them; up to NQ queues are served by the same two
// barriers.  All threads of the CTA must call, the same number of times (the scratch is double-buffered by call
// parity, so a fast warp entering the next call cannot overwrite what a slow one still reads).
template <int NQ> struct CtaReserveScratch {
    int cnt[2][NQ][32];
    unsigned long long off[2][NQ][32];       // global position of the first entry of every warp
};
template <int NQ>
__device__ __forceinline__ void cta_reserve(CtaReserveScratch<NQ> &R, int &parity, unsigned long long *const (&counter)[NQ],
                                            const bool (&push)[NQ], long long (&index)[NQ]) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    u32 m[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        m[q] = __ballot_sync(0xffffffffu, push[q]);
        if (lane == 0) R.cnt[parity][q][warp] = __popc(m[q]);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        const int q = threadIdx.x;
        int tot = 0;
        for (u32 w = 0; w < nwarp; ++w) tot += R.cnt[parity][q][w];
        unsigned long long run = tot ? atomicAdd(counter[q], (unsigned long long)tot) : 0ull;
        for (u32 w = 0; w < nwarp; ++w) { R.off[parity][q][w] = run; run += (unsigned long long)R.cnt[parity][q][w]; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q)
        index[q] = push[q] ? (long long)R.off[parity][q][warp] + __popc(m[q] & ((1u << lane) - 1u)) : -1;
    parity ^= 1;
}

// Warp-private staging of output records in shared memory.  A warp collects what its lanes produce and writes it out
// in bulk: one global atomic per FLUSH records instead of one per trip, no block barrier, and the records of a
// flush are contiguous (coalesced 8-byte stores).  The fill level is identical in all lanes (a register).
//   REC: 64-bit words per record, CAP: records the buffer holds (>= FLUSH + 31).
template <int REC, int CAP> struct WarpStage { unsigned long long w[CAP][REC]; };
template <int REC, int CAP>
__device__ __forceinline__ void warp_stage_flush(WarpStage<REC, CAP> &B, int &fill, unsigned long long *counter, long long cap,
                                                 unsigned long long *dst, const WalkerList &L, unsigned long long err_bit) {
    const u32 lane = threadIdx.x & 31;
    if (fill == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)fill);
    base = __shfl_sync(0xffffffffu, base, 0);
    int n = fill;
    if ((long long)base + n > cap) { if (lane == 0) atomicOr((unsigned long long *)&L.ctr[C_ERR], err_bit); n = (int)max(0ll, cap - (long long)base); }
    const unsigned long long *src = &B.w[0][0];
    unsigned long long *out = dst + (size_t)base * REC;
    for (int j = lane; j < n * REC; j += 32) out[j] = src[j];
    fill = 0;
    __syncwarp();
}

// The attempt kernels' own append (all lanes of the warp must call).  Records are staged per warp in shared memory and
// leave ~128 at a time: on one rank to SpawnedParts (create_particle itself, one global atomic per flush), on several
// ranks through spawn_stage_push below (peer-memory exchange) or to the staging list of the NCCL exchange.
#define NG_SPAWN_STAGE_CAP 160     /* one-word determinants: flushed at >= 128; two words: 128 records, flushed at >= 96 (48 KB of static shared memory) */
template <int NW> __host__ __device__ constexpr int spawn_stage_cap() { return NW == 1 ? NG_SPAWN_STAGE_CAP : 128; }
#define NG_MAX_PUSH_RANKS 64
template <int NW> using SpawnStage = WarpStage<NW + 2, spawn_stage_cap<NW>()>;
// what a warp needs besides its stage to push spawns to their owners
struct PushScratch {
    int hist[NG_MAX_PUSH_RANKS];                       // records of this flush per destination
    u32 base[NG_MAX_PUSH_RANKS];                       // first position reserved in the destination's segment
    unsigned short start[NG_MAX_PUSH_RANKS];           // first index of the destination's run in `perm`
    unsigned short fill[NG_MAX_PUSH_RANKS];
    unsigned char proc[NG_SPAWN_STAGE_CAP];            // owner of staged record j
    unsigned char perm[NG_SPAWN_STAGE_CAP];            // staged records grouped by destination
};
struct PushCtx { PushScratch *R; const int *roi; };
// create_particle's routing (DetermineDetNode, src/fcimc_helper.F90:152-308) and SendProcNewParts (src/Annihilation.F90:150-247)
// fused into the flush of a warp's spawn stage: every lane hashes one staged record (all lanes busy, unlike hashing at
// the point of the spawn where ~5 of 32 lanes hold one), the warp reserves positions in the owners' inbox segments
// with ONE global atomic per destination and flush (~128 records), and the records go straight over NVLink, in runs
// of ~128 / nranks consecutive records per destination.  The transfer thus overlaps the spawning kernels; what is
// left of the exchange afterwards is the mailbox hand-shake.
// (the arguments are scalars and pointers: a reference to the kernel's Params would make the compiler copy the
// whole parameter block to the stack for this out-of-line function)
struct PushArgs {
    int nranks, balance_blocks, W; u64 bb_magic; const int *lb_mapping;
    long long *const *push_seg; unsigned long long *push_cnt; long long push_off, seg_cap; long long *err;
};
template <int NW>
__device__ __noinline__ void spawn_stage_push(const PushArgs A, SpawnStage<NW> *Bp, int fill, PushScratch *Rp, const int *roi) {
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    SpawnStage<NW> &B = *Bp;
    PushScratch &R = *Rp;
    for (int d = lane; d < A.nranks; d += 32) R.hist[d] = 0;
    __syncwarp();
#pragma unroll 1
    for (int j0 = 0; j0 < fill; j0 += 32) {                    // owners and their record counts
        const int j = j0 + lane;
        int proc = -1;
        if (j < fill) {
            Det<NW> d; d.w[0] = B.w[j][0]; if (NW > 1) d.w[NW - 1] = B.w[j][NW - 1];
            proc = __ldg(&A.lb_mapping[det_block<NW>(A.balance_blocks, A.bb_magic, roi, d) - 1]);
            R.proc[j] = (unsigned char)proc;
        }
        const u32 peers = __match_any_sync(0xffffffffu, proc);
        if (proc >= 0 && lane == (u32)(__ffs(peers) - 1)) R.hist[proc] += __popc(peers);      // one lane per destination
        __syncwarp();
    }
    for (int d = lane; d < A.nranks; d += 32) {
        const int c = R.hist[d];
        R.base[d] = c ? (u32)atomicAdd(&A.push_cnt[NG_PUSH_CNT_STRIDE * d], (unsigned long long)c) : 0u;
    }
    __syncwarp();
    if (lane == 0) {                                            // runs of the destinations in the grouped order
        int run = 0;
        for (int d = 0; d < A.nranks; ++d) { R.start[d] = (unsigned short)run; R.fill[d] = 0; run += R.hist[d]; }
    }
    __syncwarp();
#pragma unroll 1
    for (int j0 = 0; j0 < fill; j0 += 32) {                    // counting sort of the record indices by destination
        const int j = j0 + lane;
        const int proc = (j < fill) ? (int)R.proc[j] : -1;
        const u32 peers = __match_any_sync(0xffffffffu, proc);
        if (proc >= 0) R.perm[R.start[proc] + R.fill[proc] + __popc(peers & lt)] = (unsigned char)j;
        __syncwarp();
        if (proc >= 0 && lane == (u32)(__ffs(peers) - 1)) R.fill[proc] += (unsigned short)__popc(peers);
        __syncwarp();
    }
    // Remote stores, word by word in the grouped order: consecutive lanes write consecutive 8-byte words of a
    // destination's run, so a store instruction covers ten records of (mostly) one destination in one or two contiguous
    // pieces.  A lane storing its own record word by word sent 8-byte fragments 24 bytes apart: three NVLink packets
    // per record, and at N = 8 the spawning kernels were bound by the packet rate (+0.19 ms).
    constexpr int RW = NW + 2;
    const int nwords = fill * RW;
#pragma unroll 1
    for (int i = lane; i < nwords; i += 32) {
        const int k = i / RW, w = i - k * RW;
        const int j = R.perm[k], d = R.proc[j];
        const long long pos = (long long)R.base[d] + (k - (int)R.start[d]);
        if (pos >= A.seg_cap) { if (w == 0) atomicOr((unsigned long long *)A.err, 1ull); continue; }
        A.push_seg[d][(size_t)(A.push_off + pos) * RW + w] = (long long)B.w[j][w];
    }
    __syncwarp();
}
template <int NW>
__device__ __forceinline__ void spawn_stage_flush(const Params &P, const SpawnBuf &SB, const WalkerList &L, SpawnStage<NW> &B, int &fill,
                                                  const PushCtx &X) {
    if (SB.push_seg) {
        if (fill) {
            PushArgs A;
            A.nranks = P.nranks; A.balance_blocks = P.balance_blocks; A.W = SB.W; A.bb_magic = P.bb_magic; A.lb_mapping = P.lb_mapping;
            A.push_seg = SB.push_seg; A.push_cnt = SB.push_cnt; A.push_off = SB.push_off; A.seg_cap = SB.seg_cap; A.err = &L.ctr[C_ERR];
            spawn_stage_push<NW>(A, &B, fill, X.R, X.roi);
        }
        fill = 0;
        return;
    }
    const bool staged = P.nranks > 1;
    warp_stage_flush<NW + 2, spawn_stage_cap<NW>()>(B, fill, staged ? SB.stage_cnt : &SB.cnt[0], staged ? SB.stage_cap : SB.seg_cap,
                                                 (unsigned long long *)(staged ? SB.stage : SB.buf), L, 1ull);
}
template <int NW>
__device__ __forceinline__ void append_spawn_k1(const Params &P, const SpawnBuf &SB, const WalkerList &L, SpawnStage<NW> &B, int &fill,
                                                const PushCtx &X, bool has, const Det<NW> &detJ, double child, long long flags) {
    const u32 lane = threadIdx.x & 31;
    const u32 m = __ballot_sync(0xffffffffu, has);
    if (m == 0) return;
    if (has) {
        unsigned long long *rec = B.w[fill + __popc(m & ((1u << lane) - 1u))];
        rec[0] = detJ.w[0];
        if (NW > 1) rec[NW - 1] = detJ.w[NW - 1];
        rec[NW] = (unsigned long long)__double_as_longlong(child);
        rec[NW + 1] = (unsigned long long)flags;
    }
    fill += __popc(m);
    __syncwarp();
    if (fill >= spawn_stage_cap<NW>() - 32) spawn_stage_flush<NW>(P, SB, L, B, fill, X);
}
// the CTA's routing scratch: RandomOrbIndex in shared memory and one PushScratch per warp (peer-memory exchange only)
struct PushShared { int roi[NG_MAX_BASIS]; PushScratch R[NG_BLOCK / 32]; };
__device__ __forceinline__ PushCtx push_ctx_init(const Params &P, const SpawnBuf &SB, PushShared &S) {
    if (SB.push_seg) {
        for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) S.roi[i] = P.random_orb_index[i];
        __syncthreads();
    }
    PushCtx X; X.R = &S.R[threadIdx.x >> 5]; X.roi = S.roi;
    return X;
}

// per-thread accumulators of the attempt stages
struct AttAcc {
    double child, child_sing, maxsp;
    int valid, invalid;
};

__device__ __forceinline__ void att_shared_init(const Params &P, AttShared &S) {
    if (threadIdx.x == 0) { S.bloom_cnt[0] = S.bloom_cnt[1] = 0; S.bloom_max[0] = S.bloom_max[1] = 0ull; }
    if (threadIdx.x < 4) { S.tau_cnt[threadIdx.x] = 0; S.tau_gamma[threadIdx.x] = 0ull; }
    __syncthreads();
}
// end of an attempt kernel: per-thread accumulators and the shared bloom / tau statistics -> this CTA's partial row
__device__ __forceinline__ void att_flush(const WalkerList &L, AttShared &S, const AttAcc &a, double *partials, double *s_red) {
    const double acc[6] = {a.child, a.child, a.child_sing, (double)a.valid, (double)a.invalid, a.maxsp};
    const int idx[6] = {NECI_ST_NOBORN, NECI_ST_ACCEPTANCES, NECI_ST_SPAWNFROMSING, NECI_ST_NVALIDEXCITS, NECI_ST_NINVALIDEXCITS,
                        NECI_ST_MAX_CYC_SPAWN};
    block_flush_stats<6>(acc, idx, partials, s_red);
    __syncthreads();
    if (threadIdx.x == 0) {
        double *row = partials + (size_t)blockIdx.x * NECI_ST_COUNT;
        row[NECI_ST_BLOOM_COUNT_1] = (double)S.bloom_cnt[0];
        row[NECI_ST_BLOOM_COUNT_2] = (double)S.bloom_cnt[1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            row[NECI_ST_TAU_GAMMA_SING + c] = __longlong_as_double((long long)S.tau_gamma[c]);
            row[NECI_ST_TAU_CNT_SING + c] = (double)S.tau_cnt[c];
        }
        if (S.bloom_max[0]) atomicMax((unsigned long long *)&L.ctr[C_COUNT - 2], S.bloom_max[0]);
        if (S.bloom_max[1]) atomicMax((unsigned long long *)&L.ctr[C_COUNT - 1], S.bloom_max[1]);
    }
}

// attempt_create_normal (fcimc_pointed_fns.F90:178-491) for an excitation of level IC whose orbitals and pgen are
// known, followed by create_particle.  `active` lanes hold a real entry; all lanes must call.  (h, att) name the
// attempt's stream: the rounding number is the first of its RNG_ATT_ROUND stream.  B / fill: the warp's spawn stage.
template <int NW, int SYS, int IC>
__device__ __forceinline__ void evaluate_and_append(const Params &P, const WalkerList &L, const SpawnBuf &SB, const IterArgs &A,
                                                    AttShared &S, SpawnStage<NW> &B, int &fill, const PushCtx &X, bool active, const Det<NW> &d,
                                                    Excit<NW> &E, int info, u64 h, u32 att, AttAcc &acc) {
    bool has = false;
    double child = 0.0;
    long long cflags = 0;
    int tau_cls = -1;
    E.ic = IC;
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
                if (IC == 1) { tp = prob / P.p_singles; tau_cls = 0; }
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
                if (IC == 1) acc.child_sing += ac;             // SpawnFromSing
                if (ac > P.initiator_walk_no) {                // bloom statistics (rare)
                    const int b = (IC == 1) ? 0 : 1;
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
    append_spawn_k1<NW>(P, SB, L, B, fill, X, has, E.detJ, child, cflags);
}

// ---------------------------------------------------------------------------------------------------------------
// k_walk: CalcParentFlag, SumEContrib, decide_num_to_spawn and walker_death of every slot; the parents of this
// iteration's attempts go to the parent list.  One thread per slot, two slots per trip (ten loads in flight).
// ---------------------------------------------------------------------------------------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK, K1_WALK_CTAS) k_walk(Params P, WalkerList L, SpawnBuf SB, K1Queues K, IterArgs A, double *partials) {
    __shared__ double s_red[13 * 32];
    __shared__ int s_npar;                                  // parents written by this CTA so far
    __shared__ WarpStage<1, 128> s_free[NG_BLOCK / 32];     // slots emptied by this warp, flushed to the FreeSlot stack in bulk
    WarpStage<1, 128> &FB = s_free[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    int n_free = 0, n_tomb = 0;
    if (threadIdx.x == 0) s_npar = 0;
    __syncthreads();
    const long long seg0 = (long long)blockIdx.x * K.par_seg_cap;
    const Det<NW> ref = ref_det<NW>(P);
    const long long n_list = L.ctr[C_NLIST];
    double c_died = 0.0, c_bornd = 0.0, c_abort = 0.0, c_hf = 0.0, c_doubs = 0.0, c_enum = 0.0, c_enumabs = 0.0,
           c_initsenum = 0.0, c_initw = 0.0, c_ninitw = 0.0, c_initd = 0.0, c_ninitd = 0.0, c_added = 0.0;
    constexpr int G = 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n_list + G * stride - 1) / (G * stride)) * (G * stride);
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nloop; i0 += G * stride) {
        double ld_s[G], ld_K[G], ld_O[G]; int ld_f[G]; Det<NW> ld_d[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const long long slot = i0 + g * stride;
            ld_s[g] = 0.0; ld_K[g] = 0.0; ld_O[g] = 0.0; ld_f[g] = 0; ld_d[g].w[0] = 0; if (NW > 1) ld_d[g].w[NW - 1] = 0;
            if (slot < n_list) {
                ld_s[g] = __ldcs(&L.sgn[slot]); ld_d[g].w[0] = __ldcs(&L.det0[slot]);
                if (NW > 1) ld_d[g].w[NW - 1] = __ldcs(&L.det1[slot]);
                ld_f[g] = __ldcs(&L.flg[slot]); ld_K[g] = __ldcs(&L.diagH[slot]); ld_O[g] = __ldcs(&L.offH[slot]);
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const long long slot = i0 + g * stride;
            int nsp = 0, info = 0;
            bool removed = false;
            Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
            const double s = ld_s[g];
            if (slot < n_list && fabs(s) >= 1.0e-12) {
                d = ld_d[g];
                int f = ld_f[g];
                const int f0 = f;
                const double Kd = ld_K[g], O = ld_O[g];
                const bool core = (f & F_DETERM) != 0;
                const int exl = excit_level_ref<NW, sys_hphf(SYS)>(ref, d);        // FindBitExcitLevel(..., t_hphf_ic = .true.)
                const double as = fabs(s);
                // CalcParentFlag / TestInitiator_explicit (fcimc_helper.F90:1036-1243)
                if (P.t_trunc_initiator) {
                    const bool was = (f & F_INIT) != 0;
                    const bool initiator = parent_is_initiator(P, was, as, exl, core);
                    if (initiator != was) c_added += initiator ? 1.0 : -1.0;
                    if (initiator) { c_initd += 1.0; c_initw += as; f |= F_INIT; }
                    else { c_ninitd += 1.0; c_ninitw += as; f &= ~F_INIT; }
                }
                // SumEContrib (fcimc_helper.F90:518-802)
                if (exl == 0) c_hf += s;
                if (exl == 2) c_doubs += as;
                const double dE = O * s;
                c_enum += dE; c_enumabs += fabs(dE);
                if (f & F_INIT) c_initsenum += dE;
                const u64 h = det_hash64(d);
                // decide_num_to_spawn (fcimc_helper.F90:2160-2174)
                {
                    const double x = s * P.av_mc_excits;
                    nsp = abs((int)x);
                    if (fabs(fabs(x) - (double)nsp) > 1.e-12) {
                        Stream rng(P.seed, A.iter, h, 0, RNG_NSPAWN);
                        if ((fabs(x) - (double)nsp) > rng.draw53()) ++nsp;
                    }
                }
                info = (s < 0.0 ? 1 : 0) | ((f & F_INIT) ? 2 : 0) | (core ? 4 : 0);
                // walker_death / attempt_die_normal (fcimc_helper.F90:2279-2407, fcimc_pointed_fns.F90:573-705)
                // tDeathBeforeComms: here with t_core_die_ = .false. (FciMCPar.F90:1752-1756); otherwise
                // perform_death_all_walkers (fcimc_helper.F90:2253-2277) would run it after the loop for every
                // determinant, core ones included -- death of slot j touches only slot j, so it is fused here too
                double news = s;
                if (!core || !P.t_death_before_comms) {
                    const double fac = A.tau * (Kd - A.diag_sft);
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
                        if (f & F_INIT) c_added -= 1.0;
                        news = 0.0;
                    }
                    if (!(fabs(news) > 1.0e-12) && !core) {
                        if (P.t_trunc_initiator && (f & F_INIT)) c_added -= 1.0;
                        if (ht_tombstone<NW>(L, d, h, slot)) ++n_tomb;     // RemoveHashDet; the free-slot push follows below
                        removed = true;
                        f |= F_REMOVED;
                        news = 0.0;
                    }
                }
                if (news != s) { L.sgn[slot] = news; mirror_sign<NW>(L, slot, news); }
                if (f != f0) { L.flg[slot] = f; mirror_flags<NW>(L, slot, f); }
                if (nsp >= (1 << 29)) atomicOr((unsigned long long *)&L.ctr[C_ERR], 512ull);     // attempt index field of the queues
                if (nsp > NG_HEAVY) {
                    const long long k = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NHEAVY], 1ull);
                    if (k < SB.heavy_cap) { SB.heavy[2 * k] = slot; SB.heavy[2 * k + 1] = ((long long)nsp << 8) | info; }
                    else atomicOr((unsigned long long *)&L.ctr[C_ERR], 32ull);
                    nsp = 0;
                }
            }
            // parent list: the CTA's own segment, one shared-memory atomic per warp
            {
                const u32 m = __ballot_sync(0xffffffffu, nsp > 0);
                if (m) {
                    int base = 0;
                    if (lane == (u32)(__ffs(m) - 1)) base = atomicAdd(&s_npar, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                    if (nsp > 0) {
                        const long long ql = base + __popc(m & lt);
                        if (ql >= K.par_seg_cap) atomicOr((unsigned long long *)&L.ctr[C_ERR], 128ull);
                        else {
                            const long long q = seg0 + ql;
                            K.par_d0[q] = d.w[0]; if (NW > 1) K.par_d1[q] = d.w[NW - 1];
                            K.par_meta[q] = ((u32)nsp << 8) | (u32)info;
                        }
                    }
                }
            }
            // FreeSlot stack: staged per warp
            {
                const u32 m = __ballot_sync(0xffffffffu, removed);
                if (m) {
                    if (removed) FB.w[n_free + __popc(m & lt)][0] = (unsigned long long)slot;
                    n_free += __popc(m);
                    __syncwarp();
                    if (n_free >= 96) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd((unsigned long long *)&L.ctr[C_NFREEB], (unsigned long long)n_free);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        for (int j = lane; j < n_free; j += 32) L.freeB[base + j] = (int)FB.w[j][0];
                        n_free = 0;
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (n_free) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd((unsigned long long *)&L.ctr[C_NFREEB], (unsigned long long)n_free);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int j = lane; j < n_free; j += 32) L.freeB[base + j] = (int)FB.w[j][0];
    }
    ht_settle_tombs(L, n_tomb);
    __syncthreads();
    if (threadIdx.x == 0) K.par_cnt[blockIdx.x] = (u32)min((long long)s_npar, K.par_seg_cap);
    const double acc[13] = {c_died, c_bornd, c_abort, c_hf, c_doubs, c_enum, c_enumabs, c_initsenum, c_initd, c_ninitd, c_initw,
                            c_ninitw, c_added};
    const int idx[13] = {NECI_ST_NODIED, NECI_ST_NOBORN, NECI_ST_NOABORTED, NECI_ST_HFCYC, NECI_ST_NOATDOUBS, NECI_ST_ENUMCYC,
                         NECI_ST_ENUMCYCABS, NECI_ST_INITSENUMCYC, NECI_ST_NOINITDETS, NECI_ST_NONONINITDETS, NECI_ST_NOINITWALK,
                         NECI_ST_NONONINITWALK, NECI_ST_NOADDEDINITIATORS};
    block_flush_stats<13>(acc, idx, partials, s_red);
}

// ---------------------------------------------------------------------------------------------------------------
// k_generate: one spawning attempt per thread
// ---------------------------------------------------------------------------------------------------------------
// stage B1: attempt index p of parent (d, h, info): draw the excitation, stage it for QE / QS.  All lanes must call.
#define NG_QE_STAGE_CAP 112      /* flushed at >= 80 */
#define NG_QS_STAGE_CAP 64       /* flushed at >= 32 */
template <int NW> struct GenStage { WarpStage<qe_rec<NW>(), NG_QE_STAGE_CAP> e; WarpStage<qs_rec<NW>(), NG_QS_STAGE_CAP> s; };
template <int NW, int SYS>
__device__ __forceinline__ void generate_and_push(const Params &P, const WalkerList &L, const K1Queues &K, const IterArgs &A,
                                                  GenStage<NW> &G, int &fill_e, int &fill_s, bool active,
                                                  const Det<NW> &d, u64 h, int info, u32 p, AttAcc &acc) {
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    bool push_e = false, push_s = false;
    Excit<NW> E;
    E.ic = 2; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = false; E.err = 0; E.parity = false;
    if (active) {
        Stream rng(P.seed, A.iter, h, p, RNG_ATTEMPT);
        if (sys_pchb(SYS)) {
            const double u = rng.draw53();                               // gen_exc_sd
            if (u < P.p_singles) push_s = true;                          // single: generated by k_singles
            else {
                if (SYS == NG_SYS_PCHB_FULL) gen_pchb_double_full(P, d, (u - P.p_singles) * P.inv_1m_ps, rng, E);
                else gen_pchb_double(P, d, (u - P.p_singles) * P.inv_1m_ps, rng, E);
                E.pgen = E.pgen * P.p_doubles;
            }
        } else generate_excitation_core<NW, SYS>(P, d, rng, E);
        if (!push_s) {
            if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
            if (E.valid) { acc.valid += 1; push_e = true; }
            else acc.invalid += 1;
        }
    }
    {
        const u32 m = __ballot_sync(0xffffffffu, push_e);
        if (push_e) {
            unsigned long long *rec = G.e.w[fill_e + __popc(m & lt)];
            rec[0] = d.w[0]; if (NW > 1) rec[NW - 1] = d.w[NW - 1];
            rec[NW] = (unsigned long long)__double_as_longlong(E.pgen);
            rec[NW + 1] = (unsigned long long)((u32)E.src1 | ((u32)E.src2 << 8) | ((u32)E.tgt1 << 16) | ((u32)E.tgt2 << 24)) |
                          ((unsigned long long)p << 32) | ((unsigned long long)info << 61);
        }
        fill_e += __popc(m);
        __syncwarp();
        if (fill_e >= NG_QE_STAGE_CAP - 32)
            warp_stage_flush<qe_rec<NW>(), NG_QE_STAGE_CAP>(G.e, fill_e, &K.cnt[Q_NQE], K.qe_cap, K.qe, L, 128ull);
    }
    if (sys_pchb(SYS)) {
        const u32 m = __ballot_sync(0xffffffffu, push_s);
        if (push_s) {
            unsigned long long *rec = G.s.w[fill_s + __popc(m & lt)];
            rec[0] = d.w[0]; if (NW > 1) rec[NW - 1] = d.w[NW - 1];
            rec[NW] = (unsigned long long)p | ((unsigned long long)info << 32);
        }
        fill_s += __popc(m);
        __syncwarp();
        if (fill_s >= 32)
            warp_stage_flush<qs_rec<NW>(), NG_QS_STAGE_CAP>(G.s, fill_s, &K.cnt[Q_NQS], K.qs_cap, K.qs, L, 128ull);
    }
}

#define NG_MAX_PAR_SEG 1024     /* CTAs of k_walk (segments of the parent list) */
template <int NW> struct GenShared {
    u64 p_d0[K1_GEN_TILE];
    u64 p_d1[(NW > 1) ? K1_GEN_TILE : 1];
    int p_off[K1_GEN_TILE + 1];          // exclusive prefix sum of the attempt counts; [TILE] = total
    unsigned short p_map[K1_MAPW];       // attempt (within the current window) -> parent index in the tile
    unsigned char p_info[K1_GEN_TILE];
    int wsum[K1_GEN_BLOCK / 32];
    int seg_tile0[NG_MAX_PAR_SEG + 1];   // first tile of every segment of the parent list
    unsigned short tile_seg[K1_GEN_TSEG]; // segment of this CTA's t-th tile, found by K1_GEN_TSEG threads at once
    int cur_seg;
};

template <int NW, int SYS>
__global__ void __launch_bounds__(K1_GEN_BLOCK) k_generate(Params P, WalkerList L, K1Queues K, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    GenShared<NW> &S = *reinterpret_cast<GenShared<NW> *>(gen_smem);
    GenStage<NW> *stages = reinterpret_cast<GenStage<NW> *>(gen_smem + ((sizeof(GenShared<NW>) + 15) & ~(size_t)15));
    __shared__ double s_red[2 * 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GenStage<NW> &G = stages[warp];
    int fill_e = 0, fill_s = 0;
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    // tiles of K1_GEN_TILE parents inside the segments k_walk's CTAs have written
    const int nseg = K.par_nseg;
    for (int sg = tid; sg < nseg; sg += K1_GEN_BLOCK) S.seg_tile0[sg + 1] = ((int)K.par_cnt[sg] + K1_GEN_TILE - 1) / K1_GEN_TILE;
    __syncthreads();
    if (warp == 0) {                                             // inclusive prefix sum over the segments: a contiguous share per lane
        const int per = (nseg + 31) / 32, s0 = min(nseg, lane * per), s1 = min(nseg, s0 + per);
        int tot = 0;
        for (int sg = s0; sg < s1; ++sg) tot += S.seg_tile0[sg + 1];
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int run = incl - tot;
        for (int sg = s0; sg < s1; ++sg) { run += S.seg_tile0[sg + 1]; S.seg_tile0[sg + 1] = run; }
        if (lane == 0) S.seg_tile0[0] = 0;
    }
    __syncthreads();
    const int n_tiles = S.seg_tile0[nseg];
    constexpr int SPT = K1_GEN_TILE / K1_GEN_BLOCK;
    // segment of a tile: last sg with seg_tile0[sg] <= tile.  One thread per tile of this CTA searches ahead of the loop
    // (a search by one thread per tile, the rest waiting at the barrier, was 11 % of the kernel's stall samples)
    auto seg_of = [&](int tile) {
        int lo_s = 0, hi_s = nseg;
        while (hi_s - lo_s > 1) { const int mid = (lo_s + hi_s) >> 1; if (S.seg_tile0[mid] <= tile) lo_s = mid; else hi_s = mid; }
        return lo_s;
    };
    if (tid < K1_GEN_TSEG && (long long)blockIdx.x + (long long)tid * gridDim.x < n_tiles) S.tile_seg[tid] = (unsigned short)seg_of(blockIdx.x + tid * gridDim.x);
    int ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        if (ti >= K1_GEN_TSEG && tid == 0) S.cur_seg = seg_of(tile);
        __syncthreads();                                         // also: the previous tile's parents and map are overwritten
        const int lo_s = (ti < K1_GEN_TSEG) ? (int)S.tile_seg[ti] : S.cur_seg;
        const long long q0 = (long long)lo_s * K.par_seg_cap + (long long)(tile - S.seg_tile0[lo_s]) * K1_GEN_TILE;
        const long long q_end = (long long)lo_s * K.par_seg_cap + (long long)K.par_cnt[lo_s];
        int nsp_k[SPT], off_k[SPT];
#pragma unroll
        for (int kk = 0; kk < SPT; ++kk) {
            const int idx = kk * K1_GEN_BLOCK + tid;
            const long long q = q0 + idx;
            Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
            u32 meta = 0;
            if (q < q_end) { d.w[0] = __ldcs(&K.par_d0[q]); if (NW > 1) d.w[NW - 1] = __ldcs(&K.par_d1[q]); meta = __ldcs(&K.par_meta[q]); }
            S.p_d0[idx] = d.w[0]; if (NW > 1) S.p_d1[idx] = d.w[NW - 1];
            S.p_info[idx] = (unsigned char)(meta & 0xffu);
            nsp_k[kk] = (int)(meta >> 8);
        }
        // exclusive prefix sum of the attempt counts over the tile (index order kk * BLOCK + tid)
        int run = 0;
#pragma unroll
        for (int kk = 0; kk < SPT; ++kk) {
            int incl = nsp_k[kk];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) S.wsum[warp] = incl;
            __syncthreads();
            int wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < K1_GEN_BLOCK / 32; ++w) { const int v = S.wsum[w]; if (w < warp) wbase += v; total += v; }
            off_k[kk] = run + wbase + incl - nsp_k[kk];
            S.p_off[kk * K1_GEN_BLOCK + tid] = off_k[kk];
            run += total;
            __syncthreads();
        }
        const int T = run;
        for (int wb = 0; wb < T; wb += K1_MAPW) {
            // Attempts are numbered 0..T-1 over the tile; each parent writes its tile index into the map entries of
            // its own attempts (window by window): an attempt finds its parent with one load.
            const int we = min(T, wb + K1_MAPW);
            if (wb) __syncthreads();                             // the previous window's map is overwritten
#pragma unroll
            for (int kk = 0; kk < SPT; ++kk) {
                const int idx = kk * K1_GEN_BLOCK + tid;
                const int lo = max(off_k[kk], wb), hi = min(off_k[kk] + nsp_k[kk], we);
                const bool big = hi - lo > 4;
                if (!big) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (lo + u < hi) S.p_map[lo + u - wb] = (unsigned short)idx;
                }
                u32 m = __ballot_sync(0xffffffffu, big);         // long ranges are filled by the whole warp
                while (m) {
                    const int src = __ffs(m) - 1; m &= m - 1u;
                    const int l2 = __shfl_sync(0xffffffffu, lo, src), h2 = __shfl_sync(0xffffffffu, hi, src);
                    const int i2 = __shfl_sync(0xffffffffu, idx, src);
#pragma unroll 1
                    for (int a = l2 + lane; a < h2; a += 32) S.p_map[a - wb] = (unsigned short)i2;
                }
            }
            __syncthreads();
            // one attempt per thread and trip; the warps run through the window independently (no barrier inside)
#pragma unroll 1
            for (int base = wb + warp * 32; base < we; base += K1_GEN_BLOCK) {
                const int a = base + lane;
                const bool active = a < we;
                Det<NW> dp; dp.w[0] = 0; if (NW > 1) dp.w[NW - 1] = 0;
                u64 h = 0; int info = 0; u32 p = 0;
                if (active) {
                    const int lo = S.p_map[a - wb];
                    dp.w[0] = S.p_d0[lo]; if (NW > 1) dp.w[NW - 1] = S.p_d1[lo];
                    h = det_hash64(dp); info = S.p_info[lo]; p = (u32)(a - S.p_off[lo]);    // the hash again per attempt: 4 KB of shared memory buy a fifth CTA per SM
                }
                generate_and_push<NW, SYS>(P, L, K, A, G, fill_e, fill_s, active, dp, h, info, p, acc);
            }
        }
    }
    warp_stage_flush<qe_rec<NW>(), NG_QE_STAGE_CAP>(G.e, fill_e, &K.cnt[Q_NQE], K.qe_cap, K.qe, L, 128ull);
    if (sys_pchb(SYS)) warp_stage_flush<qs_rec<NW>(), NG_QS_STAGE_CAP>(G.s, fill_s, &K.cnt[Q_NQS], K.qs_cap, K.qs, L, 128ull);
    const double a2[2] = {(double)acc.valid, (double)acc.invalid};
    const int idx[2] = {NECI_ST_NVALIDEXCITS, NECI_ST_NINVALIDEXCITS};
    block_flush_stats<2>(a2, idx, partials, s_red);
}
template <int NW> __host__ __device__ constexpr size_t gen_smem_bytes() {
    return ((sizeof(GenShared<NW>) + 15) & ~(size_t)15) + (K1_GEN_BLOCK / 32) * sizeof(GenStage<NW>);
}

// Attempts of the deferred heavy determinants (> NG_HEAVY walkers), spread over the whole grid.
template <int NW, int SYS>
__global__ void __launch_bounds__(K1_GEN_BLOCK) k_generate_heavy(Params P, WalkerList L, SpawnBuf SB, K1Queues K, IterArgs A, double *partials) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    GenStage<NW> *stages = reinterpret_cast<GenStage<NW> *>(gen_smem);
    __shared__ double s_red[2 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GenStage<NW> &G = stages[warp];
    int fill_e = 0, fill_s = 0;
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    long long nh = L.ctr[C_NHEAVY];
    if (nh > SB.heavy_cap) nh = SB.heavy_cap;
    const long long gwarp = (long long)blockIdx.x * (K1_GEN_BLOCK / 32) + warp, nwarps = (long long)gridDim.x * (K1_GEN_BLOCK / 32);
    for (long long e = 0; e < nh; ++e) {
        const long long slot = SB.heavy[2 * e];
        const long long packed = SB.heavy[2 * e + 1];
        const long long nsp = packed >> 8;
        const int info = (int)(packed & 0xff);
        const Det<NW> dp = load_det<NW>(L, slot);
        const u64 h = det_hash64(dp);
        const long long rounds = (nsp + 31) / 32;
#pragma unroll 1
        for (long long rd = gwarp; rd < rounds; rd += nwarps) {
            const long long a = rd * 32 + lane;
            generate_and_push<NW, SYS>(P, L, K, A, G, fill_e, fill_s, a < nsp, dp, h, info, (u32)a, acc);
        }
    }
    warp_stage_flush<qe_rec<NW>(), NG_QE_STAGE_CAP>(G.e, fill_e, &K.cnt[Q_NQE], K.qe_cap, K.qe, L, 128ull);
    if (sys_pchb(SYS)) warp_stage_flush<qs_rec<NW>(), NG_QS_STAGE_CAP>(G.s, fill_s, &K.cnt[Q_NQS], K.qs_cap, K.qs, L, 128ull);
    const double a2[2] = {(double)acc.valid, (double)acc.invalid};
    const int idx[2] = {NECI_ST_NVALIDEXCITS, NECI_ST_NINVALIDEXCITS};
    block_flush_stats<2>(a2, idx, partials, s_red);
}

// ---------------------------------------------------------------------------------------------------------------
// k_evaluate: one thread per QE entry (stage B2)
// ---------------------------------------------------------------------------------------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK, sys_hphf(SYS) ? 4 : 5) k_evaluate(Params P, WalkerList L, SpawnBuf SB, K1Queues K, IterArgs A, double *partials) {
    __shared__ AttShared S;
    __shared__ SpawnStage<NW> s_stage[NG_BLOCK / 32];
    __shared__ double s_red[6 * 32];
    constexpr int IC = (SYS == NECI_SYS_HUBBARD_RS) ? 1 : 2;
    __shared__ PushShared s_push;
    SpawnStage<NW> &B = s_stage[threadIdx.x >> 5];
    int fill = 0;
    const PushCtx X = push_ctx_init(P, SB, s_push);
    att_shared_init(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    long long n = (long long)K.cnt[Q_NQE]; if (n > K.qe_cap) n = K.qe_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        const bool active = i < n;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        Excit<NW> E; E.ic = IC; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = active; E.err = 0; E.detJ = d; E.parity = false;
        int info = 0; u32 att = 0;
        if (active) {
            const u64 *rec = K.qe + (size_t)i * qe_rec<NW>();
            d.w[0] = __ldcs(&rec[0]); if (NW > 1) d.w[NW - 1] = __ldcs(&rec[NW - 1]);
            E.pgen = __longlong_as_double((long long)__ldcs(&rec[NW]));
            const u64 oa = __ldcs(&rec[NW + 1]);
            const u32 o = (u32)oa;
            E.src1 = o & 0xff; E.src2 = (o >> 8) & 0xff; E.tgt1 = (o >> 16) & 0xff; E.tgt2 = o >> 24;
            att = (u32)(oa >> 32) & 0x1FFFFFFFu; info = (int)(oa >> 61);
        }
        const u64 h = det_hash64(d);
        evaluate_and_append<NW, SYS, IC>(P, L, SB, A, S, B, fill, X, active, d, E, info, h, att, acc);
    }
    spawn_stage_flush<NW>(P, SB, L, B, fill, X);
    if (SB.push_seg) __threadfence_system();                  // the pushed records are visible at their owners before the mailboxes are posted
    att_flush(L, S, acc, partials, s_red);
}

// ---------------------------------------------------------------------------------------------------------------
// k_singles: one thread per QS entry (stage B3, PCHB only)
// ---------------------------------------------------------------------------------------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK, K1_SING_CTAS) k_singles(Params P, WalkerList L, SpawnBuf SB, K1Queues K, IterArgs A, double *partials) {
    __shared__ AttShared S;
    __shared__ SpawnStage<NW> s_stage[NG_BLOCK / 32];
    __shared__ double s_red[6 * 32];
    __shared__ PushShared s_push;
    SpawnStage<NW> &B = s_stage[threadIdx.x >> 5];
    int fill = 0;
    const PushCtx X = push_ctx_init(P, SB, s_push);
    att_shared_init(P, S);
    AttAcc acc; acc.child = acc.child_sing = acc.maxsp = 0.0; acc.valid = acc.invalid = 0;
    long long n = (long long)K.cnt[Q_NQS]; if (n > K.qs_cap) n = K.qs_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        bool active = i < n;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        Excit<NW> E; E.ic = 1; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 1; E.pgen = 1.0; E.valid = false; E.err = 0; E.detJ = d; E.parity = false;
        int info = 0; u32 att = 0; u64 h = 0;
        if (active) {
            const u64 *rec = K.qs + (size_t)i * qs_rec<NW>();
            d.w[0] = __ldcs(&rec[0]); if (NW > 1) d.w[NW - 1] = __ldcs(&rec[NW - 1]);
            const u64 pm = __ldcs(&rec[NW]);
            att = (u32)pm; info = (int)(pm >> 32);
            h = det_hash64(d);
            Stream rng(P.seed, A.iter, h, att, RNG_ATTEMPT, 4);     // the first block chose "single"; singles draw from word 4
            gen_uniform_single(P, d, rng, E);
            E.pgen = E.pgen * P.p_singles;
            if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
            if (E.valid) acc.valid += 1;
            else { acc.invalid += 1; active = false; }
        }
        evaluate_and_append<NW, SYS, 1>(P, L, SB, A, S, B, fill, X, active, d, E, info, h, att, acc);
    }
    spawn_stage_flush<NW>(P, SB, L, B, fill, X);
    if (SB.push_seg) __threadfence_system();                  // the pushed records are visible at their owners before the mailboxes are posted
    att_flush(L, S, acc, partials, s_red);
}

}  // namespace ng
