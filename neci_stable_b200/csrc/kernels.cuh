// The hot-path kernels of one FCIQMC iteration (PerformFCIMCycPar,
// src/FciMCPar.F90:1177-1920), B200-first:
//
//   k_spawn          loop over determinants (:1294-1758): initiator flags
//                    (CalcParentFlag), energy accumulators (SumEContrib),
//                    spawning (generate_excitation + attempt_create +
//                    create_particle) and death (walker_death), fused in one
//                    pass over the SoA walker list.  Attempts are distributed
//                    over the threads of a CTA per tile (prefix sum + search),
//                    very heavy determinants are deferred to k_spawn_heavy.
//   k_compress       CompressSpawnedList (Annihilation.F90:249-515) as an
//                    in-place hash merge of the received spawn records.
//   k_annihilate     AnnihilateSpawnedParts (:965-1352): probe the main hash
//                    table, merge signs, abort / round, queue new determinants.
//   k_insert         AddNewHashDet (load_balancer.fpp:514-629) incl.
//                    get_diagonal_matel / get_off_diagonal_matel.
//   k_list_stats     CalcHashTableStats (load_balancer.fpp:646-805).
//   k_determ_spmv    determ_projection (semi_stoch_procs.F90:105-241).
#pragma once
#include "device_system.cuh"

namespace ng {

#define NG_BLOCK 256
#define NG_HEAVY 4096        /* attempts per determinant handled inside a tile */

struct SpawnBuf {
    long long *buf;          // SpawnedParts: nranks segments of seg_cap records (W words each)
    long long *recv;         // received records (contiguous)
    unsigned long long *cnt; // ValidSpawnedList - InitialSpawnedSlots, per destination rank
    long long seg_cap;
    int W;
    // spawn-merge hash table, entries [stamp:16][tag:16][index:32]
    u64 *sht; u64 sht_cap;
    int *ins_idx;            // records that become new determinants
    long long *heavy;        // (slot, nspawn) pairs
    long long heavy_cap;
};

struct IterArgs {
    double tau, diag_sft;
    long long iter;
    long long n_recv;        // < 0: read SB.cnt[0] on the device (single rank)
    u32 stamp;
};

// ---- block-level reduction of per-thread statistics into per-block partials ---
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
// acc[k] for the statistics listed in idx[k]; writes out[blockIdx.x * NECI_ST_COUNT + idx[k]]
// (all other entries of the block row are zeroed by the caller's prologue).
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

// final reduction over the partial rows of all kernels: one CTA per statistic, fixed
// (launch-independent) summation tree, so results are reproducible run to run
__global__ void __launch_bounds__(256) k_reduce_stats(const double *partials, int nrows, double *stats) {
    __shared__ double s_v[256];
    const int k = blockIdx.x;
    const bool is_max = (k >= NECI_ST_FIRST_MAX && k <= NECI_ST_LAST_MAX) || k == NECI_ST_HIGHEST_POP;
    double v = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 256) {
        const double x = partials[(size_t)r * NECI_ST_COUNT + k];
        v = is_max ? fmax(v, x) : v + x;
    }
    s_v[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_v[threadIdx.x] = is_max ? fmax(s_v[threadIdx.x], s_v[threadIdx.x + o]) : s_v[threadIdx.x] + s_v[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) stats[k] = s_v[0];
}

// ---- stochastic_round (src/lib/util_mod.fpp:182-204) ---------------------------
__device__ __forceinline__ double stochastic_round(double r, Stream &rng) {
    int i = (int)r;
    const double res = r - (double)i;
    if (fabs(res) >= 1.0e-12) {
        if (fabs(res) > rng.draw()) i += (r < 0.0 || (r == 0.0 && signbit(r))) ? -1 : 1;
    }
    return (double)i;
}

// One spawning attempt: generate_excitation + attempt_create_normal
// (src/fcimc_pointed_fns.F90:178-491).  Returns the child weight (0 = none).
enum { A_NOBORN = 0, A_SING, A_ACC, A_VALID, A_INVALID, A_BC1, A_BC2, A_MAXSP, A_BS1, A_BS2, A_COUNT };

template <int NW, int SYS, int NA>
__device__ __forceinline__ double do_attempt(const Params &P, const WalkerList &L, const IterArgs &A,
                                             const Det<NW> &d, u64 h, u32 p, bool neg, bool core,
                                             Det<NW> &detJ, double (&acc)[NA], int &child_extra_flags) {
    Stream rng(P.seed, A.iter, h, p, RNG_ATTEMPT);
    Excit<NW> E;
    generate_excitation<NW, SYS>(P, d, rng, E);
    if (E.err) atomicOr((unsigned long long *)&L.ctr[C_ERR], 16ull);
    if (!E.valid) { acc[A_INVALID] += 1.0; return 0.0; }
    acc[A_VALID] += 1.0;
    child_extra_flags = 0;
    if (P.t_semi_stochastic && core) {
        // core -> core spawning is done by determ_projection (FciMCPar.F90:1651-1670)
        const long long s = ht_lookup<NW>(L, E.detJ, det_hash64(E.detJ));
        if (s >= 0 && (L.flg[s] & F_DETERM)) return 0.0;
        child_extra_flags = F_DPARENT;
    }
    const double prob = E.pgen * P.av_mc_excits;
    const double rh = spawn_helement<NW, SYS>(P, d, E);
    const double ww = neg ? -1.0 : 1.0;
    double nSpawn = -A.tau * rh * ww / prob;
    acc[A_MAXSP] = fmax(acc[A_MAXSP], fabs(nSpawn));
    if (P.t_all_real_coeff) {
        if (P.t_real_spawn_cutoff && fabs(nSpawn) < P.real_spawn_cutoff)
            nSpawn = P.real_spawn_cutoff * stochastic_round(nSpawn / P.real_spawn_cutoff, rng);
    } else nSpawn = stochastic_round(nSpawn, rng);
    if (fabs(nSpawn) <= NG_EPS) return 0.0;
    const double ac = fabs(nSpawn);
    acc[A_NOBORN] += ac;
    if (E.ic == 1) acc[A_SING] += ac;
    if (ac > P.initiator_walk_no) {
        if (E.ic == 1) { acc[A_BC1] += 1.0; acc[A_BS1] = fmax(acc[A_BS1], ac); }
        else { acc[A_BC2] += 1.0; acc[A_BS2] = fmax(acc[A_BS2], ac); }
    }
    acc[A_ACC] += ac;
    detJ = E.detJ;
    return nSpawn;
}

// create_particle (src/fcimc_helper.F90:152-308): warp-aggregated append of
// (ilutJ, child, flags) to the destination rank's segment of SpawnedParts.
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

enum { S_NODIED = A_COUNT, S_NOBORN_D, S_ABORT, S_HF, S_DOUBS, S_ENUM, S_ENUMABS, S_INITSENUM,
       S_INITD, S_NINITD, S_INITW, S_NINITW, S_ADDED, S_COUNT };

template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK) k_spawn(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    __shared__ int s_roi[NG_MAX_BASIS];
    __shared__ u64 s_d0[NG_BLOCK];
    __shared__ u64 s_d1[(NW > 1) ? NG_BLOCK : 1];
    __shared__ u64 s_h[NG_BLOCK];
    __shared__ int s_off[NG_BLOCK + 1];
    __shared__ unsigned char s_info[NG_BLOCK];
    __shared__ int s_wsum[NG_BLOCK / 32];
    __shared__ double s_red[S_COUNT * 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < P.nbasis; i += NG_BLOCK) s_roi[i] = P.random_orb_index[i];
    double acc[S_COUNT];
#pragma unroll
    for (int k = 0; k < S_COUNT; ++k) acc[k] = 0.0;
    const Det<NW> ref = ref_det<NW>(P);
    const long long n_list = L.ctr[C_NLIST];
    __syncthreads();

    for (long long tile = blockIdx.x; tile * NG_BLOCK < n_list; tile += gridDim.x) {
        const long long slot = tile * NG_BLOCK + tid;
        int nsp = 0;
        unsigned char info = 0;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        u64 h = 0;
        if (slot < n_list) {
            const double s = L.sgn[slot];
            if (fabs(s) >= 1.0e-12) {
                d = load_det<NW>(L, slot);
                int f = L.flg[slot];
                const int f0 = f;
                const double K = L.diagH[slot], O = L.offH[slot];
                const bool core = (f & F_DETERM) != 0;
                const int exl = excit_level(ref, d);
                const double as = fabs(s);
                // CalcParentFlag / TestInitiator_explicit (fcimc_helper.F90:1036-1243)
                if (P.t_trunc_initiator) {
                    bool initiator = (f & F_INIT) != 0;
                    const bool popInit = as > P.initiator_walk_no;
                    if (!initiator) { if (popInit) { initiator = true; acc[S_ADDED] += 1.0; } }
                    else if (exl != 0 && !(core && P.t_core_inits) && !popInit) { initiator = false; acc[S_ADDED] -= 1.0; }
                    if (initiator) { acc[S_INITD] += 1.0; acc[S_INITW] += as; f |= F_INIT; }
                    else { acc[S_NINITD] += 1.0; acc[S_NINITW] += as; f &= ~F_INIT; }
                }
                // SumEContrib (fcimc_helper.F90:518-802)
                if (exl == 0) acc[S_HF] += s;
                if (exl == 2) acc[S_DOUBS] += as;
                const double dE = O * s;
                acc[S_ENUM] += dE; acc[S_ENUMABS] += fabs(dE);
                if (f & F_INIT) acc[S_INITSENUM] += dE;
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
                double news = s;
                if (!core) {
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
                    acc[S_NODIED] += fmin(iDie, as);
                    acc[S_NOBORN_D] += fmax(iDie - as, 0.0);
                    news = s - (iDie * dsign(1.0, s));
                    if (P.t_trunc_initiator && fabs(news) > 1.0e-12 && ((news > 0.0) != (s > 0.0))) {
                        acc[S_ABORT] += fabs(news);
                        if (f & F_INIT) acc[S_ADDED] -= 1.0;
                        news = 0.0;
                    }
                    if (!(fabs(news) > 1.0e-12)) {
                        if (P.t_trunc_initiator && (f & F_INIT)) acc[S_ADDED] -= 1.0;
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
        // ---- distribute the tile's attempts over the CTA -------------------------
        s_d0[tid] = d.w[0]; if (NW > 1) s_d1[tid] = d.w[NW - 1];
        s_h[tid] = h; s_info[tid] = info;
        int incl = nsp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_wsum[w];
        s_off[tid] = wbase + incl - nsp;
        if (tid == NG_BLOCK - 1) s_off[NG_BLOCK] = wbase + incl;
        __syncthreads();
        const int T = s_off[NG_BLOCK];
        for (int base = 0; base < T; base += NG_BLOCK) {
            const int a = base + tid;
            bool has = false; Det<NW> detJ; double child = 0.0; long long cflags = 0;
            detJ.w[0] = 0; if (NW > 1) detJ.w[NW - 1] = 0;
            if (a < T) {
                int lo = 0, hi = NG_BLOCK - 1;            // last index with s_off[idx] <= a
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= a) lo = mid; else hi = mid - 1; }
                Det<NW> dp; dp.w[0] = s_d0[lo]; if (NW > 1) dp.w[NW - 1] = s_d1[lo];
                const unsigned char inf = s_info[lo];
                int extra = 0;
                child = do_attempt<NW, SYS>(P, L, A, dp, s_h[lo], (u32)(a - s_off[lo]), inf & 1, (inf & 4) != 0,
                                            detJ, acc, extra);
                if (child != 0.0) {
                    has = true;
                    cflags = (long long)extra;
                    if (P.t_trunc_initiator && (inf & 2)) cflags |= F_INIT;
                }
            }
            append_spawn<NW>(P, SB, L, s_roi, has, detJ, child, cflags);
        }
        __syncthreads();
    }
    // ---- flush statistics -------------------------------------------------------
    double out[20];
    out[0] = acc[A_NOBORN] + acc[S_NOBORN_D]; out[1] = acc[S_NODIED]; out[2] = acc[S_ABORT]; out[3] = acc[A_SING];
    out[4] = acc[A_ACC]; out[5] = acc[S_HF]; out[6] = acc[S_DOUBS]; out[7] = acc[S_ENUM]; out[8] = acc[S_ENUMABS];
    out[9] = acc[S_INITSENUM]; out[10] = acc[S_INITD]; out[11] = acc[S_NINITD]; out[12] = acc[S_INITW];
    out[13] = acc[S_NINITW]; out[14] = acc[S_ADDED]; out[15] = acc[A_VALID]; out[16] = acc[A_INVALID];
    out[17] = acc[A_BC1]; out[18] = acc[A_BC2]; out[19] = acc[A_MAXSP];
    const int idx[20] = {NECI_ST_NOBORN, NECI_ST_NODIED, NECI_ST_NOABORTED, NECI_ST_SPAWNFROMSING, NECI_ST_ACCEPTANCES,
                         NECI_ST_HFCYC, NECI_ST_NOATDOUBS, NECI_ST_ENUMCYC, NECI_ST_ENUMCYCABS, NECI_ST_INITSENUMCYC,
                         NECI_ST_NOINITDETS, NECI_ST_NONONINITDETS, NECI_ST_NOINITWALK, NECI_ST_NONONINITWALK,
                         NECI_ST_NOADDEDINITIATORS, NECI_ST_NVALIDEXCITS, NECI_ST_NINVALIDEXCITS,
                         NECI_ST_BLOOM_COUNT_1, NECI_ST_BLOOM_COUNT_2, NECI_ST_MAX_CYC_SPAWN};
    block_flush_stats<20>(out, idx, partials, s_red);
    // bloom sizes and error bits: rarely non-zero, merged with atomics
    if (acc[A_BS1] > 0.0 || acc[A_BS2] > 0.0) {
        atomicMax((unsigned long long *)&L.ctr[C_COUNT - 2], (unsigned long long)__double_as_longlong(acc[A_BS1]));
        atomicMax((unsigned long long *)&L.ctr[C_COUNT - 1], (unsigned long long)__double_as_longlong(acc[A_BS2]));
    }
}

// Attempts of the deferred heavy determinants, spread over the whole grid.
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK) k_spawn_heavy(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    __shared__ int s_roi[NG_MAX_BASIS];
    __shared__ double s_red[A_COUNT * 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < P.nbasis; i += NG_BLOCK) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    double acc[A_COUNT];
#pragma unroll
    for (int k = 0; k < A_COUNT; ++k) acc[k] = 0.0;
    long long nh = L.ctr[C_NHEAVY];
    if (nh > SB.heavy_cap) nh = SB.heavy_cap;
    for (long long e = 0; e < nh; ++e) {
        const long long slot = SB.heavy[2 * e];
        const long long packed = SB.heavy[2 * e + 1];
        const int nsp = (int)(packed >> 8);
        const unsigned char inf = (unsigned char)(packed & 0xff);
        const Det<NW> dp = load_det<NW>(L, slot);
        const u64 h = det_hash64(dp);
        const int rounds = (nsp + NG_BLOCK - 1) / NG_BLOCK;
        for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
            const int a = rd * NG_BLOCK + tid;
            bool has = false; Det<NW> detJ; double child = 0.0; long long cflags = 0;
            detJ.w[0] = 0; if (NW > 1) detJ.w[NW - 1] = 0;
            if (a < nsp) {
                int extra = 0;
                child = do_attempt<NW, SYS>(P, L, A, dp, h, (u32)a, inf & 1, (inf & 4) != 0, detJ, acc, extra);
                if (child != 0.0) { has = true; cflags = (long long)extra; if (P.t_trunc_initiator && (inf & 2)) cflags |= F_INIT; }
            }
            append_spawn<NW>(P, SB, L, s_roi, has, detJ, child, cflags);
        }
    }
    double out[8] = {acc[A_NOBORN], acc[A_SING], acc[A_ACC], acc[A_VALID], acc[A_INVALID], acc[A_BC1], acc[A_BC2], acc[A_MAXSP]};
    const int idx[8] = {NECI_ST_NOBORN, NECI_ST_SPAWNFROMSING, NECI_ST_ACCEPTANCES, NECI_ST_NVALIDEXCITS,
                        NECI_ST_NINVALIDEXCITS, NECI_ST_BLOOM_COUNT_1, NECI_ST_BLOOM_COUNT_2, NECI_ST_MAX_CYC_SPAWN};
    block_flush_stats<8>(out, idx, partials, s_red);
    if (acc[A_BS1] > 0.0 || acc[A_BS2] > 0.0) {
        atomicMax((unsigned long long *)&L.ctr[C_COUNT - 2], (unsigned long long)__double_as_longlong(acc[A_BS1]));
        atomicMax((unsigned long long *)&L.ctr[C_COUNT - 1], (unsigned long long)__double_as_longlong(acc[A_BS2]));
    }
}

// freeB -> freeA, clamp counters (runs with one block per 256 entries + 1)
__global__ void k_merge_free(WalkerList L) {
    __shared__ long long s_a, s_b;
    if (threadIdx.x == 0) { long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0; s_a = a; s_b = L.ctr[C_NFREEB]; }
    __syncthreads();
    const long long a = s_a, b = s_b;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (long long)gridDim.x * blockDim.x)
        L.freeA[a + i] = L.freeB[i];
}
__global__ void k_merge_free_finish(WalkerList L) {
    long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0;
    L.ctr[C_NFREEA] = a + L.ctr[C_NFREEB];
    L.ctr[C_NFREEB] = 0;
}

// ---- CompressSpawnedList as an in-place hash merge ------------------------------
__device__ __forceinline__ long long recv_count(const SpawnBuf &SB, const IterArgs &A) {
    return (A.n_recv >= 0) ? A.n_recv : (long long)SB.cnt[0];
}
__device__ __forceinline__ u64 sht_mask_for(long long n, u64 cap) {
    u64 m = 1024;
    while (m < 2ull * (u64)n && m < cap) m <<= 1;
    return m - 1;
}

template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_compress(Params P, SpawnBuf SB, IterArgs A, double *partials) {
    __shared__ double s_red[32];
    const long long n = recv_count(SB, A);
    const u64 mask = sht_mask_for(n, SB.sht_cap);
    const u64 stamp = (u64)(A.stamp & 0xFFFFu) << 48;
    double acc[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long *rec = SB.recv + (size_t)i * SB.W;
        Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
        const double s = __longlong_as_double(rec[NW]);
        const long long f = rec[NW + 1];
        acc[0] += fabs(s);
        const u64 h = det_hash64(d);
        const u64 mine = stamp | ((h >> 48) << 32) | (u64)(u32)i;
        u64 pos = h & mask;
        u64 e = __ldcg(&SB.sht[pos]);                     // L2 reads: slots are claimed concurrently
        for (;;) {
            if ((e >> 48) != (stamp >> 48)) {
                const u64 old = atomicCAS((unsigned long long *)&SB.sht[pos], e, mine);
                if (old == e) break;                      // this record represents its determinant
                e = old;                                  // lost the race: judge the winner's entry
                continue;
            }
            if (((e >> 32) & 0xFFFFu) == (h >> 48)) {
                const long long j = (long long)(e & 0xFFFFFFFFull);
                const long long *rj = SB.recv + (size_t)j * SB.W;
                bool same = ((u64)rj[0] == d.w[0]);
                if (NW > 1) same = same && ((u64)rj[NW - 1] == d.w[NW - 1]);
                if (same) {
                    // FindResidualParticle (Annihilation.F90:551-634): sign sum, flag union
                    atomicAdd((double *)&rj[NW], s);
                    atomicOr((unsigned long long *)&rj[NW + 1], (unsigned long long)((f & F_INIT) | SF_MULTI));
                    rec[NW + 1] = SF_DEAD;
                    break;
                }
            }
            pos = (pos + 1) & mask;
            e = __ldcg(&SB.sht[pos]);
        }
    }
    const int idx[1] = {NECI_ST_ANNIHILATED};
    block_flush_stats<1>(acc, idx, partials, s_red);
}

// ---- AnnihilateSpawnedParts ------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_annihilate(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    __shared__ double s_red[6 * 32];
    const long long n = recv_count(SB, A);
    double acc[6] = {0, 0, 0, 0, 0, 0};     // annihilated, aborted, removed, born(round), merged, recv
    const u32 lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        bool ins = false;
        if (i < n) {
            long long *rec = SB.recv + (size_t)i * SB.W;
            long long f = rec[NW + 1];
            acc[5] += 1.0;
            if (!(f & SF_DEAD)) {
                Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
                double s = __longlong_as_double(rec[NW]);
                const bool multi = (f & SF_MULTI) != 0;
                acc[0] -= fabs(s);                       // Annihilated(compress) = sum|s_i| - |sum s_i|
                if (fabs(s) > 1.e-12 || (!multi && fabs(s) >= 1.e-12)) {
                    acc[4] += 1.0;
                    bool spawn_init = (f & F_INIT) != 0;
                    if (multi) {
                        if (P.t_trunc_initiator && P.t_init_coherent_rule) spawn_init = true;
                        f = spawn_init ? (long long)F_INIT : 0ll;      // cum_det carries initiator flags only
                    } else f &= (long long)(F_INIT | F_DPARENT);
                    const u64 h = det_hash64(d);
                    u64 pos;
                    const long long slot = ht_lookup<NW>(L, d, h, &pos);
                    if (slot >= 0) {
                        const double cur = L.sgn[slot];
                        const int fl = L.flg[slot];
                        const bool tDet = (fl & F_DETERM) != 0;
                        if (fabs(cur) >= 1.e-12 || tDet) {
                            if (cur * s < 0.0) acc[0] += 2.0 * fmin(fabs(cur), fabs(s));
                            const double ns = s + cur;
                            L.sgn[slot] = ns;
                            if (!tDet && fabs(ns) < 1.0e-12) {
                                L.ht[pos] = HT_TOMB;
                                atomicAdd((unsigned long long *)&L.ctr[C_NTOMB], 1ull);
                                const long long k = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NFREEB], 1ull);
                                L.freeB[k] = (int)slot;
                                L.flg[slot] = fl | F_REMOVED;
                            }
                        }
                    } else {
                        if (P.t_trunc_initiator && !spawn_init) { acc[1] += fabs(s); s = 0.0; }   // test_abort_spawn
                        if (fabs(s) >= 1.0e-12) {
                            const double thr = P.occupied_thresh;        // stochRoundSpawn
                            if (fabs(s) > 1.e-12 && fabs(s) < thr) {
                                const double pRemove = 1.0 - fabs(s) / thr;
                                Stream rng(P.seed, A.iter, h, 0, RNG_ROUND_SPAWN);
                                if (pRemove > rng.draw()) { acc[2] += fabs(s); s = 0.0; }
                                else { acc[3] += thr - fabs(s); s = dsign(thr, s); }
                            }
                            if (fabs(s) >= 1.0e-12) {
                                rec[NW] = __double_as_longlong(s);
                                rec[NW + 1] = f;
                                ins = true;
                            }
                        }
                    }
                }
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, ins);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd((unsigned long long *)&L.ctr[C_NINSERT], (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ins) SB.ins_idx[base + __popc(m & ((1u << lane) - 1u))] = (int)i;
        }
    }
    const int idx[6] = {NECI_ST_ANNIHILATED, NECI_ST_NOABORTED, NECI_ST_NOREMOVED, NECI_ST_NOBORN,
                        NECI_ST_NSPAWNED_MERGED, NECI_ST_NSPAWNED_RECV};
    block_flush_stats<6>(acc, idx, partials, s_red);
}

// ---- AddNewHashDet ----------------------------------------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK) k_insert(Params P, WalkerList L, SpawnBuf SB, double *partials) {
    __shared__ double s_red[32];
    const long long n = L.ctr[C_NINSERT];
    double acc[1] = {0.0};
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const long long i = SB.ins_idx[k];
        const long long *rec = SB.recv + (size_t)i * SB.W;
        Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
        const double s = __longlong_as_double(rec[NW]);
        const int f = (int)(rec[NW + 1] & 0x7fffffffll) & ~F_REMOVED;
        const double hd = diagonal_matel<NW, SYS>(P, d) - P.hii;
        const double ho = off_diagonal_matel<NW, SYS>(P, d);
        long long slot;
        const long long fi = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NFREEA], (unsigned long long)-1ll) - 1;
        if (fi >= 0) slot = L.freeA[fi];
        else {
            slot = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NLIST], 1ull);
            if (slot + 1 >= L.cap) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 2ull); continue; }
        }
        store_det<NW>(L, slot, d);
        L.sgn[slot] = s; L.flg[slot] = f; L.diagH[slot] = hd; L.offH[slot] = ho;
        const u64 h = det_hash64(d);
        ht_insert(L, h, slot, h & L.ht_mask);
        acc[0] += 1.0;
    }
    const int idx[1] = {NECI_ST_NINSERTED};
    block_flush_stats<1>(acc, idx, partials, s_red);
}
__global__ void k_fix_counters(WalkerList L) {
    if (L.ctr[C_NFREEA] < 0) L.ctr[C_NFREEA] = 0;
    if (L.ctr[C_NLIST] > L.cap - 1) L.ctr[C_NLIST] = L.cap - 1;
}

// ---- CalcHashTableStats ---------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_list_stats(Params P, WalkerList L, IterArgs A, double *partials) {
    __shared__ double s_red[6 * 32];
    const long long n = L.ctr[C_NLIST];
    double acc[6] = {0, 0, 0, 0, 0, 0};     // totparts, norm2, norm_ss2, removed, born, highest
    const bool need_flags = P.t_semi_stochastic != 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = L.sgn[i];
        const bool tDet = need_flags && (L.flg[i] & F_DETERM);
        if (fabs(s) < 1.0e-12 && !tDet) continue;
        if (!tDet && fabs(s) > 1.e-12 && fabs(s) < P.occupied_thresh) {
            const Det<NW> d = load_det<NW>(L, i);
            const u64 h = det_hash64(d);
            const double pRemove = (P.occupied_thresh - fabs(s)) / P.occupied_thresh;
            Stream rng(P.seed, A.iter, h, 0, RNG_PRUNE);
            if (pRemove > rng.draw()) {
                acc[3] += fabs(s);
                s = 0.0; L.sgn[i] = 0.0;
                ht_remove<NW>(L, d, h, i);
                L.flg[i] |= F_REMOVED;
            } else {
                acc[4] += P.occupied_thresh - fabs(s);
                s = dsign(P.occupied_thresh, s); L.sgn[i] = s;
            }
        }
        acc[0] += fabs(s); acc[1] += s * s;
        if (tDet) acc[2] += s * s;
        acc[5] = fmax(acc[5], (double)(long long)fabs(s));
    }
    const int idx[6] = {NECI_ST_TOTPARTS, NECI_ST_NORM_PSI_SQ, NECI_ST_NORM_SEMISTOCH_SQ, NECI_ST_NOREMOVED,
                        NECI_ST_NOBORN, NECI_ST_HIGHEST_POP};
    block_flush_stats<6>(acc, idx, partials, s_red);
}

// final bookkeeping: InstNoatHF, counters -> stats (single thread)
template <int NW>
__global__ void k_finish_stats(Params P, WalkerList L, SpawnBuf SB, double *stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Det<NW> ref = ref_det<NW>(P);
    const long long slot = ht_lookup<NW>(L, ref, det_hash64(ref));
    stats[NECI_ST_INSTNOATHF] = (slot >= 0) ? L.sgn[slot] : 0.0;
    stats[NECI_ST_TOTWALKERS] = (double)L.ctr[C_NLIST];
    long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0;
    stats[NECI_ST_HOLESINLIST] = (double)(a + L.ctr[C_NFREEB]);
    unsigned long long sent = 0;
    for (int r = 0; r < P.nranks; ++r) sent += SB.cnt[r];
    stats[NECI_ST_NSPAWNED_SENT] = (double)sent;
    stats[NECI_ST_ERR_FLAGS] = (double)L.ctr[C_ERR];
    stats[NECI_ST_BLOOM_SIZE_1] = __longlong_as_double(L.ctr[C_COUNT - 2]);
    stats[NECI_ST_BLOOM_SIZE_2] = __longlong_as_double(L.ctr[C_COUNT - 1]);
}

// ---- hash-table maintenance -------------------------------------------------------
__global__ void k_fill_u64(u64 *p, u64 v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
template <int NW>
__global__ void k_ht_rebuild(Params P, WalkerList L) {
    const long long n = L.ctr[C_NLIST];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double s = L.sgn[i];
        const int f = L.flg[i];
        if (fabs(s) >= 1.0e-12 || (f & F_DETERM)) {
            const u64 h = det_hash64(load_det<NW>(L, i));
            ht_insert(L, h, i, h & L.ht_mask);
        }
    }
}

// ---- upload / download (AoS ilut(0:NIfTot) <-> SoA) ------------------------------
template <int NW, int SYS>
__global__ void k_upload(Params P, WalkerList L, const long long *aos, long long n, const double *gd, const double *go, int W) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long *rec = aos + (size_t)i * W;
        Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
        const double s = __longlong_as_double(rec[NW]);
        const int f = (int)(rec[NW + 1] & 0x7fffffffll);
        store_det<NW>(L, i, d);
        L.sgn[i] = s; L.flg[i] = f;
        const bool live = fabs(s) >= 1.0e-12 || (f & F_DETERM);
        L.diagH[i] = gd ? gd[i] : (live ? diagonal_matel<NW, SYS>(P, d) - P.hii : 0.0);
        L.offH[i] = go ? go[i] : (live ? off_diagonal_matel<NW, SYS>(P, d) : 0.0);
        if (live) { const u64 h = det_hash64(d); ht_insert(L, h, i, h & L.ht_mask); }
        else { const long long k = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NFREEA], 1ull); L.freeA[k] = (int)i; }
    }
}
template <int NW>
__global__ void k_download(WalkerList L, long long *aos, long long n, int W) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long *rec = aos + (size_t)i * W;
        rec[0] = (long long)L.det0[i]; if (NW > 1) rec[NW - 1] = (long long)L.det1[i];
        rec[NW] = __double_as_longlong(L.sgn[i]);
        rec[NW + 1] = (long long)L.flg[i];
    }
}

// ---- semi-stochastic ---------------------------------------------------------------
// gather of partial_determ_vecs (FciMCPar.F90:1387-1411)
__global__ void k_core_gather(WalkerList L, const int *core_slots, long long n, double *v_part) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v_part[i] = L.sgn[core_slots[i]];
}
// determ_projection: one warp per row; out_i = tau * (-sum_j H_ij v_j + S * v_{i+displ}),
// fused with deterministic_annihilation (Annihilation.F90:930-963): sign += out_i.
__global__ void __launch_bounds__(NG_BLOCK) k_determ_spmv(WalkerList L, const long long *row_ptr, const int *col, const double *val,
                                                          const double *v_full, long long n_local, long long displ,
                                                          double tau, double diag_sft, const int *core_slots, double *out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n_local; i += nwarps) {
        const long long b = row_ptr[i], e = row_ptr[i + 1];
        double acc = 0.0;
        for (long long k = b + lane; k < e; k += 32) acc -= __ldg(&val[k]) * __ldg(&v_full[__ldg(&col[k])]);
        acc = warp_sum(acc);
        if (lane == 0) {
            acc = (acc + diag_sft * v_full[i + displ]) * tau;
            out[i] = acc;
        }
    }
}
__global__ void k_determ_apply(WalkerList L, const int *core_slots, const double *out, long long n_local) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x)
        L.sgn[core_slots[i]] += out[i];
}

// locate the core determinants in the list and flag them (check_determ_flag)
template <int NW>
__global__ void k_core_locate(Params P, WalkerList L, const long long *iluts, long long n, int *slots) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> d; d.w[0] = (u64)iluts[i * NW]; if (NW > 1) d.w[NW - 1] = (u64)iluts[i * NW + NW - 1];
        const long long s = ht_lookup<NW>(L, d, det_hash64(d));
        if (s < 0) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 64ull); slots[i] = 0; }
        else { slots[i] = (int)s; L.flg[s] |= F_DETERM; }
    }
}

// ---- probes ---------------------------------------------------------------------------
template <int NW>
__global__ void k_probe_det_node(Params P, const long long *iluts, long long n, int *block_out, int *node_out) {
    __shared__ int s_roi[NG_MAX_BASIS];
    for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> d; d.w[0] = (u64)iluts[i * NW]; if (NW > 1) d.w[NW - 1] = (u64)iluts[i * NW + NW - 1];
        const int b = det_block<NW>(P, s_roi, d);
        block_out[i] = b; node_out[i] = P.lb_mapping[b - 1];
    }
}
template <int NW, int SYS>
__global__ void k_probe_helement(Params P, const long long *ii, const long long *ij, long long n, double *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> a, b;
        a.w[0] = (u64)ii[i * NW]; b.w[0] = (u64)ij[i * NW];
        if (NW > 1) { a.w[NW - 1] = (u64)ii[i * NW + NW - 1]; b.w[NW - 1] = (u64)ij[i * NW + NW - 1]; }
        out[i] = helement<NW, SYS>(P, a, b);
    }
}
template <int NW, int SYS>
__global__ void k_probe_gen_excit(Params P, const long long *iluts, const int *attempt, long long iter, long long n,
                                  long long *ilut_j, int *ic, int *ex, int *par, double *pgen, double *hel) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> d; d.w[0] = (u64)iluts[i * NW]; if (NW > 1) d.w[NW - 1] = (u64)iluts[i * NW + NW - 1];
        Stream rng(P.seed, iter, det_hash64(d), (u32)attempt[i], RNG_ATTEMPT);
        Excit<NW> E;
        generate_excitation<NW, SYS>(P, d, rng, E);
        ic[i] = E.ic;
        if (E.valid) {
            ilut_j[i * NW] = (long long)E.detJ.w[0]; if (NW > 1) ilut_j[i * NW + NW - 1] = (long long)E.detJ.w[NW - 1];
            ex[4 * i] = E.src1; ex[4 * i + 1] = E.src2; ex[4 * i + 2] = E.tgt1; ex[4 * i + 3] = E.tgt2;
            par[i] = E.parity ? 1 : 0; pgen[i] = E.pgen; hel[i] = spawn_helement<NW, SYS>(P, d, E);
        } else {
            for (int w = 0; w < NW; ++w) ilut_j[i * NW + w] = 0;
            ex[4 * i] = ex[4 * i + 1] = ex[4 * i + 2] = ex[4 * i + 3] = 0;
            par[i] = 0; pgen[i] = 0.0; hel[i] = 0.0;
        }
    }
}

// per-block walker populations for adjust_load_balance (load_balancer.fpp:216-235)
template <int NW>
__global__ void k_block_pops(Params P, WalkerList L, double *block_parts) {
    __shared__ int s_roi[NG_MAX_BASIS];
    for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    const long long n = L.ctr[C_NLIST];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double s = L.sgn[i];
        if (fabs(s) < 1.0e-12) continue;
        const int b = det_block<NW>(P, s_roi, load_det<NW>(L, i));
        atomicAdd(&block_parts[b - 1], fabs(s));
    }
}

}  // namespace ng
