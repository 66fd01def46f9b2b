// The hot-path kernels of one FCIQMC iteration (PerformFCIMCycPar,
// src/FciMCPar.F90:1177-1920), B200-first:
//
//   k_walk, k_generate, k_evaluate, k_singles
//                      (spawn_kernel.cuh) loop over determinants (:1294-1758): initiator flags, energy
//                      accumulators, death; spawning attempts staged through global-memory queues;
//                      k_generate_heavy takes determinants with > 4096 attempts.
//   k_trial_energy     trial part of SumEContrib (fcimc_helper.F90:586-648).
//   k_compress         CompressSpawnedList (Annihilation.F90:249-515) as an in-place hash merge of the
//                      received spawn records; also merges the FreeSlot lists.
//   k_annihilate       AnnihilateSpawnedParts (:965-1352): probe the main hash table, merge signs,
//                      abort / round, queue new determinants.
//   k_insert           AddNewHashDet (load_balancer.fpp:514-629) incl. get_diagonal_matel /
//                      get_off_diagonal_matel and hash_search_trial.
//   k_list_stats       CalcHashTableStats (load_balancer.fpp:646-805).
//   k_reduce_stats     the iteration's statistics vector (communicate_estimates' per-rank inputs).
//   k_determ_spmv_blocked (+ k_determ_finish)  determ_projection / determ_projection_no_death (semi_stoch_procs.F90:105-374).
//   k_partition, k_push, k_wait, k_gather
//                      SendProcNewParts over NVLink peer memory (the spawning kernels route and push their spawns
//                      themselves: spawn_stage_push in spawn_kernel.cuh) / routing for the NCCL exchange.
//   k_rebalance_pack   move_block, sender side (load_balancer.fpp:353-512).
//   k_pops_*           POPSFILE gather (Popsfile.F90:2054-2107).
//   k_upload / k_download / k_probe_*   list transfer and the parity probes.
#pragma once
#include "spawn_kernel.cuh"

namespace ng {

// final reduction over the partial rows of all kernels: one CTA per statistic, fixed
// (launch-independent) summation tree, so results are reproducible run to run
template <int NW>
__global__ void __launch_bounds__(256) k_reduce_stats(Params P, WalkerList L, SpawnBuf SB, const double *partials, int nrows, double *stats) {
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
    // the device counters ride along behind the statistics (one copy to the host instead of two)
    if (k == 0 && threadIdx.x < C_COUNT) reinterpret_cast<long long *>(stats)[NECI_ST_COUNT + threadIdx.x] = L.ctr[threadIdx.x];
    if (threadIdx.x != 0) return;
    double out = s_v[0];
    // the statistics that are counters or look-ups rather than sums over the list (InstNoatHF, list length, holes, ...)
    if (k == NECI_ST_INSTNOATHF) {
        const Det<NW> ref = ref_det<NW>(P);
        const long long slot = ht_lookup<NW>(L, ref, det_hash64(ref));
        out = (slot >= 0) ? L.sgn[slot] : 0.0;
    } else if (k == NECI_ST_TOTWALKERS) out = (double)min(L.ctr[C_NLIST], L.cap - 1);
    else if (k == NECI_ST_HOLESINLIST) { long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0; out = (double)(a + L.ctr[C_NFREEB]); }
    else if (k == NECI_ST_NSPAWNED_SENT) { unsigned long long sent = 0; for (int r = 0; r < P.nranks; ++r) sent += SB.cnt[r]; out = (double)sent; }
    else if (k == NECI_ST_ERR_FLAGS) out = (double)L.ctr[C_ERR];
    else if (k == NECI_ST_BLOOM_SIZE_1) out = __longlong_as_double(L.ctr[C_COUNT - 2]);
    else if (k == NECI_ST_BLOOM_SIZE_2) out = __longlong_as_double(L.ctr[C_COUNT - 1]);
    stats[k] = out;
}

// ValidSpawnedList = InitialSpawnedSlots etc. (FciMCPar.F90:1237-1248): every per-iteration counter in one launch
__global__ void k_begin_iteration(WalkerList L, SpawnBuf SB, K1Queues K, int nranks) {
    const int t = threadIdx.x;
    if (t < nranks) { SB.cnt[t] = 0ull; if (SB.push_cnt) SB.push_cnt[NG_PUSH_CNT_STRIDE * t] = 0ull; }
    if (t >= C_NHEAVY && t < C_COUNT) L.ctr[t] = 0;
    if (t < 4) K.cnt[t] = 0ull;
    if (t == 0 && SB.stage_cnt) *SB.stage_cnt = 0ull;
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
    if (A.n_recv >= 0) return A.n_recv;
    // append_spawn keeps counting past seg_cap after an overflow (the error bit reports it): never walk past the buffer
    if (A.n_recv == -1) return (long long)min(SB.cnt[0], (unsigned long long)SB.seg_cap);
    return (long long)*SB.n_recv_dev;
}
__device__ __forceinline__ u64 sht_mask_for(long long n, u64 cap) {
    u64 m = 1024;
    while (m < 2ull * (u64)n && m < cap) m <<= 1;
    return m - 1;
}

// Order-independent sum of real-coefficient spawns onto one determinant.  atomicAdd(double) would make the last
// bits of a merged amplitude depend on the order in which the contributors arrive, i.e. on the schedule.  Instead
// every contribution is split exactly into a multiple of 2^-24 and a remainder resolved to 2^-64, the two parts are
// accumulated with 64-bit integer atomics (associative), and the sums are joined once: the same bits whatever the
// order, within one ulp of the sequential sum (exact split for |s| >= 2^-11; |sum| < 2^39).  Integer amplitudes
// add exactly in any order and keep the plain fp64 atomic.
__device__ __forceinline__ void fixed_split(double s, long long &hi, long long &lo) {
    const double x = s * 16777216.0;
    hi = (long long)x;
    lo = (long long)((x - (double)hi) * 1099511627776.0);
}
__device__ __forceinline__ double fixed_join(long long hi, long long lo) {
    return ((double)hi + (double)lo * (1.0 / 1099511627776.0)) * (1.0 / 16777216.0);
}

template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_compress(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials, unsigned int *ticket) {
    __shared__ double s_red[32];
    __shared__ bool s_last;
    // FreeSlot bookkeeping that used to be two launches: the slots freed by the spawning pass (freeB) are appended
    // to the pop side (freeA) by all CTAs; the CTA that finishes last publishes the new counters.  Nothing in this
    // kernel touches the free lists otherwise, and the next kernel starts after this one has ended.
    {
        long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0;
        const long long b = L.ctr[C_NFREEB];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (long long)gridDim.x * blockDim.x)
            L.freeA[a + i] = L.freeB[i];
    }
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
                    if (P.t_all_real_coeff) {
                        long long hi, lo; fixed_split(s, hi, lo);
                        atomicAdd((unsigned long long *)&SB.acc_hi[j], (unsigned long long)hi);
                        atomicAdd((unsigned long long *)&SB.acc_lo[j], (unsigned long long)lo);
                    } else atomicAdd((double *)&rj[NW], s);
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
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        long long a = L.ctr[C_NFREEA]; if (a < 0) a = 0;
        L.ctr[C_NFREEA] = a + L.ctr[C_NFREEB];
        L.ctr[C_NFREEB] = 0;
        *ticket = 0u;
    }
}

// ---- AnnihilateSpawnedParts ------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_annihilate(Params P, WalkerList L, SpawnBuf SB, IterArgs A, double *partials) {
    __shared__ double s_red[6 * 32];
    __shared__ CtaReserveScratch<2> R;
    int parity = 0, n_tomb = 0;
    const long long n = recv_count(SB, A);
    double acc[6] = {0, 0, 0, 0, 0, 0};     // annihilated, aborted, removed, born(round), merged, recv
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        bool ins = false;
        long long freed = -1;                // slot emptied by this record (RemoveHashDet)
        if (i < n) {
            long long *rec = SB.recv + (size_t)i * SB.W;
            long long f = rec[NW + 1];
            acc[5] += 1.0;
            if (!(f & SF_DEAD)) {
                Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
                double s = __longlong_as_double(rec[NW]);
                const bool multi = (f & SF_MULTI) != 0;
                if (multi && P.t_all_real_coeff) {
                    // the representative's own amplitude joins the fixed-point sums of the others; the accumulators
                    // go back to zero for the next iteration
                    long long hi, lo; fixed_split(s, hi, lo);
                    hi += SB.acc_hi[i]; lo += SB.acc_lo[i];
                    SB.acc_hi[i] = 0; SB.acc_lo[i] = 0;
                    s = fixed_join(hi, lo);
                }
                acc[0] -= fabs(s);                       // Annihilated(compress) = sum|s_i| - |sum s_i|
                if (fabs(s) > 1.e-12 || (!multi && fabs(s) >= 1.e-12)) {
                    acc[4] += 1.0;
                    bool spawn_init = (f & F_INIT) != 0;
                    if (multi) {
                        // FindResidualParticle touches the flags of a merged block only under tTruncInitiator
                        // (Annihilation.F90:576-590); without it cum_det's flag word stays zero
                        if (!P.t_trunc_initiator) spawn_init = false;
                        else if (P.t_init_coherent_rule) spawn_init = true;
                        f = spawn_init ? (long long)F_INIT : 0ll;      // cum_det carries initiator flags only
                    } else f &= (long long)(F_INIT | F_DPARENT);
                    const u64 h = det_hash64(d);
                    u64 pos;
                    const long long slot = ht_lookup<NW>(L, d, h, &pos);
                    if (slot >= 0) {
                        const double cur = L.sgn[slot];
                        // the flag word is a separate random sector: only semi-stochastic runs need it here (core
                        // determinants are never removed); a removal below reads it when it happens
                        int fl = P.t_semi_stochastic ? L.flg[slot] : 0;
                        const bool tDet = (fl & F_DETERM) != 0;
                        if (fabs(cur) >= 1.e-12 || tDet) {
                            if (cur * s < 0.0) acc[0] += 2.0 * fmin(fabs(cur), fabs(s));
                            const double ns = s + cur;
                            L.sgn[slot] = ns; mirror_sign<NW>(L, slot, ns);
                            if (!tDet && fabs(ns) < 1.0e-12) {
                                L.ht[pos] = HT_TOMB;
                                ++n_tomb;
                                freed = slot;
                                if (!P.t_semi_stochastic) fl = L.flg[slot];
                                L.flg[slot] = fl | F_REMOVED; mirror_flags<NW>(L, slot, fl | F_REMOVED);
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
        // queue of new determinants and FreeSlot stack: one reservation per CTA and trip for both
        unsigned long long *const counter[2] = {(unsigned long long *)&L.ctr[C_NINSERT], (unsigned long long *)&L.ctr[C_NFREEB]};
        const bool push[2] = {ins, freed >= 0};
        long long q[2];
        cta_reserve<2>(R, parity, counter, push, q);
        if (ins) SB.ins_idx[q[0]] = (int)i;
        if (freed >= 0) L.freeB[q[1]] = (int)freed;
    }
    ht_settle_tombs(L, n_tomb);
    const int idx[6] = {NECI_ST_ANNIHILATED, NECI_ST_NOABORTED, NECI_ST_NOREMOVED, NECI_ST_NOBORN,
                        NECI_ST_NSPAWNED_MERGED, NECI_ST_NSPAWNED_RECV};
    block_flush_stats<6>(acc, idx, partials, s_red);
}

// ---- AddNewHashDet ----------------------------------------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(NG_BLOCK) k_insert(Params P, WalkerList L, SpawnBuf SB, double *partials) {
    __shared__ double s_red[32];
    __shared__ int s_cnt[2][32];
    __shared__ long long s_old[2], s_new[2];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long n = L.ctr[C_NINSERT];
    double acc[1] = {0.0};
    int n_reused = 0, parity = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nloop; k += stride) {
        const bool active = k < n;
        // slots for the determinants of this trip: from the top of the FreeSlot stack while it lasts, then from the
        // end of the list -- one pop and one extension per CTA instead of one atomic per determinant
        const u32 m = __ballot_sync(0xffffffffu, active);
        if (lane == 0) s_cnt[parity][warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = 0;
            for (u32 w = 0; w < nwarp; ++w) tot += s_cnt[parity][w];
            long long old = 0, base_new = 0;
            if (tot) {
                old = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NFREEA], (unsigned long long)(-tot));
                const long long avail = min(max(old, 0ll), tot);
                if (tot > avail) base_new = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NLIST], (unsigned long long)(tot - avail));
            }
            s_old[parity] = old; s_new[parity] = base_new;
        }
        __syncthreads();
        const int par = parity; parity ^= 1;
        if (!active) continue;
        long long rank = __popc(m & ((1u << lane) - 1u));
        for (u32 w = 0; w < warp; ++w) rank += s_cnt[par][w];
        const long long old = s_old[par], avail = max(old, 0ll);
        long long slot;
        if (rank < avail) slot = L.freeA[old - 1 - rank];
        else {
            slot = s_new[par] + (rank - avail);
            if (slot + 1 >= L.cap) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 2ull); continue; }
        }
        const long long i = SB.ins_idx[k];
        const long long *rec = SB.recv + (size_t)i * SB.W;
        Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
        const double s = __longlong_as_double(rec[NW]);
        const int f = (int)(rec[NW + 1] & 0x7fffffffll) & ~F_REMOVED;
        const double hd = diagonal_matel<NW, SYS>(P, d) - P.hii;
        const double ho = off_diagonal_matel<NW, SYS>(P, d);
        int ft = f;
        if (P.trial_ht) {                                   // hash_search_trial, load_balancer.fpp:586-611
            double amp;
            ft = (f & ~(F_TRIAL | F_CONNECTED)) | trial_lookup<NW>(P, d, &amp);
            L.trial_amp[slot] = amp;
        }
        store_det<NW>(L, slot, d);
        L.sgn[slot] = s; L.flg[slot] = ft; L.diagH[slot] = hd; L.offH[slot] = ho;
        mirror_record<NW>(L, slot, d, s, ft, hd, ho);
        const u64 h = det_hash64(d);
        if (ht_insert(L, h, slot, h & L.ht_mask)) --n_reused;
        acc[0] += 1.0;
    }
    ht_settle_tombs(L, n_reused);
    const int idx[1] = {NECI_ST_NINSERTED};
    block_flush_stats<1>(acc, idx, partials, s_red);
}
__global__ void k_fix_counters(WalkerList L) {
    if (L.ctr[C_NFREEA] < 0) L.ctr[C_NFREEA] = 0;
    if (L.ctr[C_NLIST] > L.cap - 1) L.ctr[C_NLIST] = L.cap - 1;
}

// ---- CalcHashTableStats ---------------------------------------------------------
// the stochastic pruning of one under-threshold determinant: kept out of line so that the streaming loop of
// k_list_stats stays at a register count that allows full occupancy.  Returns the new sign; *removed is set when the
// determinant left the list (its hash-table entry is a tombstone then; the caller returns the slot to the FreeSlot
// stack in bulk and counts the tombstone).
template <int NW>
__device__ __noinline__ double prune_slot(const Params &P, const WalkerList &L, long long iter, long long i, double s, bool *removed,
                                          int *n_tomb) {
    const Det<NW> d = load_det<NW>(L, i);
    const u64 h = det_hash64(d);
    const double pRemove = (P.occupied_thresh - fabs(s)) / P.occupied_thresh;
    Stream rng(P.seed, iter, h, 0, RNG_PRUNE);
    if (pRemove > rng.draw()) {
        L.sgn[i] = 0.0; mirror_sign<NW>(L, i, 0.0);
        if (ht_tombstone<NW>(L, d, h, i)) *n_tomb += 1;
        const int f = L.flg[i] | F_REMOVED;
        L.flg[i] = f; mirror_flags<NW>(L, i, f);
        *removed = true;
        return 0.0;
    }
    s = dsign(P.occupied_thresh, s); L.sgn[i] = s; mirror_sign<NW>(L, i, s);
    return s;
}
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK, 4) k_list_stats(const __grid_constant__ Params P, const __grid_constant__ WalkerList L, IterArgs A,
                                                            double *partials) {
    __shared__ double s_red[6 * 32];
    __shared__ WarpStage<1, 128> s_free[NG_BLOCK / 32];     // slots emptied by pruning (real coefficients: many), pushed in bulk
    WarpStage<1, 128> &FB = s_free[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    int n_free = 0, n_tomb = 0;
    // k_insert may have overshot the counters when the list overflowed (error already flagged): clamp, as every CTA
    // does for itself; the stored values are repaired by one thread
    long long n = L.ctr[C_NLIST]; if (n > L.cap - 1) n = L.cap - 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (L.ctr[C_NFREEA] < 0) L.ctr[C_NFREEA] = 0;
        if (L.ctr[C_NLIST] > L.cap - 1) L.ctr[C_NLIST] = L.cap - 1;
    }
    double acc[6] = {0, 0, 0, 0, 0, 0};     // totparts, norm2, norm_ss2, removed, born, highest
    const bool need_flags = P.t_semi_stochastic != 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + 4 * stride - 1) / (4 * stride)) * (4 * stride);
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nloop; i0 += 4 * stride) {
        double sv[4]; int fv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                // four independent loads in flight per thread
            const long long i = i0 + u * stride;
            sv[u] = (i < n) ? L.sgn[i] : 0.0;
            fv[u] = (need_flags && i < n) ? L.flg[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            bool removed = false;
            double s = sv[u];
            const bool tDet = need_flags && (fv[u] & F_DETERM);
            if (i < n && !(fabs(s) < 1.0e-12 && !tDet)) {
                if (!tDet && fabs(s) > 1.e-12 && fabs(s) < P.occupied_thresh) {
                    const double old = fabs(s);
                    s = prune_slot<NW>(P, L, A.iter, i, s, &removed, &n_tomb);
                    if (s == 0.0) acc[3] += old; else acc[4] += P.occupied_thresh - old;       // NoRemoved / NoBorn
                }
                acc[0] += fabs(s); acc[1] += s * s;
                if (tDet) acc[2] += s * s;
                acc[5] = fmax(acc[5], (double)(long long)fabs(s));
            }
            const u32 m = __ballot_sync(0xffffffffu, removed);
            if (m) {
                if (removed) FB.w[n_free + __popc(m & lt)][0] = (unsigned long long)i;
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
    if (n_free) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd((unsigned long long *)&L.ctr[C_NFREEB], (unsigned long long)n_free);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int j = lane; j < n_free; j += 32) L.freeB[base + j] = (int)FB.w[j][0];
    }
    ht_settle_tombs(L, n_tomb);
    const int idx[6] = {NECI_ST_TOTPARTS, NECI_ST_NORM_PSI_SQ, NECI_ST_NORM_SEMISTOCH_SQ, NECI_ST_NOREMOVED,
                        NECI_ST_NOBORN, NECI_ST_HIGHEST_POP};
    block_flush_stats<6>(acc, idx, partials, s_red);
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
            ht_insert(L, h, i, h & L.ht_mask);              // the table was cleared: no tombstones to recycle
        }
    }
}

// ---- frozen synthetic walker list, generated on the device (benchmark set-up) ------------------
// Candidate c of the GLOBAL list is a function of (seed, c) alone: n_alpha of the n_spat spatial orbitals for the alpha
// electrons and n_beta for the beta electrons, uniformly (sequential selection sampling), sign +-round(1 + Exp(1))
// drawn from a stream keyed by the determinant.
// Every rank walks all candidates and keeps the determinants it owns (DetermineDetNode), so the global list does not
// depend on the number of ranks.  Records go to the AoS staging buffer; duplicates are nulled by k_synth_dedupe, and
// the list is then taken in by k_upload like an uploaded CurrentDets.
template <int NW>
__global__ void __launch_bounds__(256) k_synth_records(Params P, u64 seed, long long n_cand, int n_spat, long long *aos, int W, long long cap,
                                                       unsigned long long *count) {
    __shared__ int s_roi[NG_MAX_BASIS];
    for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n_cand + stride - 1) / stride) * stride;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nloop; c += stride) {
        bool keep = false;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        double sgn = 0.0;
        if (c < n_cand) {
            Stream rng(seed, 0x5EED, mix64((u64)c + 0x9E3779B97F4A7C15ull), 0, RNG_ATTEMPT);
#pragma unroll 1
            for (int spin = 0; spin < 2; ++spin) {              // 0: alpha (even orbitals), 1: beta (odd orbitals)
                int need = spin ? P.nocc_beta : P.nocc_alpha;
#pragma unroll 1
                for (int o = 0; o < n_spat && need > 0; ++o) {
                    const u32 u = rng.next_u32();
                    if ((u64)u * (u64)(n_spat - o) < ((u64)need << 32)) { set_orb(d, 2 * (o + 1) - spin); --need; }
                }
            }
            // the sign is a function of the determinant, not of the candidate: whichever record of a determinant drawn
            // twice survives k_synth_dedupe, the list is the same
            Stream rs(seed, 0x5EED, det_hash64(d), 1, RNG_ATTEMPT);
            const double mag = rint(1.0 - log(1.0 - rs.draw53()));
            sgn = (rs.next_u32() & 1u) ? -mag : mag;
            keep = (P.nranks == 1) || (__ldg(&P.lb_mapping[det_block<NW>(P, s_roi, d) - 1]) == P.rank);
        }
        const u32 m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            unsigned long long base = 0;
            if (lane == (u32)(__ffs(m) - 1)) base = atomicAdd(count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            const long long pos = (long long)base + __popc(m & lt);
            if (keep && pos < cap) {
                long long *rec = aos + (size_t)pos * W;
                rec[0] = (long long)d.w[0]; if (NW > 1) rec[NW - 1] = (long long)d.w[NW - 1];
                rec[NW] = __double_as_longlong(sgn); rec[NW + 1] = 0;
            }
        }
    }
}
// a determinant drawn twice keeps its first-inserted record; the others get a zero sign (holes at upload).  `ht` is the
// (cleared) main hash table used as a set over record indices.
template <int NW>
__global__ void __launch_bounds__(256) k_synth_dedupe(WalkerList L, long long *aos, int W, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long *rec = aos + (size_t)i * W;
        Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
        const u64 h = det_hash64(d);
        const u64 entry = ((u64)(u32)(h >> 32) << 32) | (u64)(u32)i;
        u64 pos = h & L.ht_mask;
        u64 e = __ldcg(&L.ht[pos]);
        for (;;) {
            if (e == HT_EMPTY) {
                const u64 old = atomicCAS((unsigned long long *)&L.ht[pos], e, entry);
                if (old == e) break;
                e = old;
                continue;
            }
            if ((u32)(e >> 32) == (u32)(h >> 32)) {
                const long long *rj = aos + (size_t)(e & 0xFFFFFFFFull) * W;
                bool same = (u64)rj[0] == d.w[0];
                if (NW > 1) same = same && (u64)rj[NW - 1] == d.w[NW - 1];
                if (same) { rec[NW] = 0; break; }
            }
            pos = (pos + 1) & L.ht_mask;
            e = __ldcg(&L.ht[pos]);
        }
    }
}

// ---- upload / download (AoS ilut(0:NIfTot) <-> SoA) ------------------------------
template <int NW, int SYS>
__global__ void __launch_bounds__(256) k_upload(Params P, WalkerList L, const long long *aos, long long i_begin, long long n, const double *gd,
                                                const double *go, int W, long long n_prev) {      // slots [i_begin, n)
    // empty slots go to the FreeSlot stack through a per-warp stage: one global atomic per ~100 holes, not one per hole
    __shared__ WarpStage<1, 128> s_free[8];
    WarpStage<1, 128> &FB = s_free[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    int n_free = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = i_begin + ((n - i_begin + stride - 1) / stride) * stride;
    for (long long i = i_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        bool hole = false;
        if (i < n) {
            const long long *rec = aos + (size_t)i * W;
            Det<NW> d; d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
            const double s = __longlong_as_double(rec[NW]);
            int f = (int)(rec[NW + 1] & 0x7fffffffll);
            if (P.trial_ht) { double amp; f = (f & ~(F_TRIAL | F_CONNECTED)) | trial_lookup<NW>(P, d, &amp); L.trial_amp[i] = amp; }
            const bool live = fabs(s) >= 1.0e-12 || (f & F_DETERM);
            // global_determinant_data: taken from the host when it passes it; otherwise kept when this slot already
            // holds the same determinant from an earlier call (H_ii and H_0i are functions of the determinant), else
            // recomputed (get_diagonal_matel / get_off_diagonal_matel)
            double hd = 0.0, ho = 0.0;
            if (gd && go) { hd = gd[i]; ho = go[i]; }
            else if (live) {
                const bool same = L.det0[i] == d.w[0] && (NW == 1 || L.det1[i] == d.w[NW - 1]) && (L.flg[i] & F_REMOVED) == 0 &&
                                  i < n_prev;
                hd = gd ? gd[i] : (same ? L.diagH[i] : diagonal_matel<NW, SYS>(P, d) - P.hii);
                ho = go ? go[i] : (same ? L.offH[i] : off_diagonal_matel<NW, SYS>(P, d));
            }
            store_det<NW>(L, i, d);
            L.sgn[i] = s; L.flg[i] = f; L.diagH[i] = hd; L.offH[i] = ho;
            if (live) { const u64 h = det_hash64(d); ht_insert(L, h, i, h & L.ht_mask); }      // fresh table: no tombstones
            else hole = true;
        }
        const u32 m = __ballot_sync(0xffffffffu, hole);
        if (m) {
            if (hole) FB.w[n_free + __popc(m & lt)][0] = (unsigned long long)i;
            n_free += __popc(m);
            __syncwarp();
            if (n_free >= 96) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd((unsigned long long *)&L.ctr[C_NFREEA], (unsigned long long)n_free);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int j = lane; j < n_free; j += 32) L.freeA[base + j] = (int)FB.w[j][0];
                n_free = 0;
                __syncwarp();
            }
        }
    }
    if (n_free) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd((unsigned long long *)&L.ctr[C_NFREEA], (unsigned long long)n_free);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int j = lane; j < n_free; j += 32) L.freeA[base + j] = (int)FB.w[j][0];
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

// ---- POPSFILE gather (write_pops_det, src/Popsfile.F90:2054-2107) ----------------------
// The binary POPSFILE record of a determinant is det(0:NIfD), sign, flags [, gdata]: the AoS ilut itself.  Occupied
// determinants above binarypops_min_weight are compacted in slot order: per-CTA counts over fixed chunks, a scan
// of the counts by one CTA, then the ordered write.
#define NG_POPS_CHUNK 4096
__global__ void __launch_bounds__(256) k_pops_count(WalkerList L, double min_weight, int *chunk_cnt) {
    __shared__ int s_c;
    const long long n = L.ctr[C_NLIST];
    const long long nchunks = (n + NG_POPS_CHUNK - 1) / NG_POPS_CHUNK;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x == 0) s_c = 0;
        __syncthreads();
        int mine = 0;
        for (int k = threadIdx.x; k < NG_POPS_CHUNK; k += 256) {
            const long long i = c * NG_POPS_CHUNK + k;
            if (i < n && fabs(L.sgn[i]) > min_weight) ++mine;
        }
        atomicAdd(&s_c, mine);
        __syncthreads();
        if (threadIdx.x == 0) chunk_cnt[c] = s_c;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_pops_scan(WalkerList L, int *chunk_cnt, long long *total) {
    __shared__ long long s_part[1024];
    const long long n = L.ctr[C_NLIST];
    const long long nchunks = (n + NG_POPS_CHUNK - 1) / NG_POPS_CHUNK;
    const long long per = (nchunks + 1023) / 1024, lo = threadIdx.x * per, hi = min(nchunks, lo + per);
    long long sum = 0;
    for (long long c = lo; c < hi; ++c) sum += chunk_cnt[c];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { long long run = 0; for (int t = 0; t < 1024; ++t) { const long long v = s_part[t]; s_part[t] = run; run += v; } *total = run; }
    __syncthreads();
    long long run = s_part[threadIdx.x];
    // counts -> exclusive offsets in place (the list holds < 2^31 slots, neci_gpu_init checks max_walkers)
    for (long long c = lo; c < hi; ++c) { const int v = chunk_cnt[c]; chunk_cnt[c] = (int)run; run += v; }
}
template <int NW>
__global__ void __launch_bounds__(256) k_pops_write(WalkerList L, double min_weight, const int *chunk_off, long long *aos, int W,
                                                    double *gd, double *go) {
    __shared__ int s_w[8];
    const long long n = L.ctr[C_NLIST];
    const long long nchunks = (n + NG_POPS_CHUNK - 1) / NG_POPS_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        long long base = chunk_off[c];
        for (int k0 = 0; k0 < NG_POPS_CHUNK; k0 += 256) {
            const long long i = c * NG_POPS_CHUNK + k0 + threadIdx.x;
            const bool keep = i < n && fabs(L.sgn[i]) > min_weight;
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_w[warp] = __popc(m);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int v = s_w[w]; if (w < warp) woff += v; tot += v; }
            if (keep) {
                const long long o = base + woff + __popc(m & ((1u << lane) - 1u));
                long long *rec = aos + (size_t)o * W;
                rec[0] = (long long)L.det0[i]; if (NW > 1) rec[NW - 1] = (long long)L.det1[i];
                rec[NW] = __double_as_longlong(L.sgn[i]);
                rec[NW + 1] = (long long)(L.flg[i] & ~F_REMOVED);
                if (gd) gd[o] = L.diagH[i];
                if (go) go[o] = L.offH[i];
            }
            base += tot;
            __syncthreads();
        }
    }
}

// ---- semi-stochastic ---------------------------------------------------------------
// gather of partial_determ_vecs (FciMCPar.F90:1387-1411)
__global__ void k_core_gather(WalkerList L, const int *core_slots, long long n, double *v_part) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v_part[i] = L.sgn[core_slots[i]];
}
// determ_projection: out_i = tau * (-sum_j H_ij v_j + S * v_{i+displ}), fused with the shift / diagonal term of
// determ_projection_no_death.
//
// What bounds a CSR SpMV here is not HBM but the gather of v_full: rows hold ~2000 elements spread over 1e5 columns, so
// the 32 lanes of a gather touch 32 different cache lines and the L1 tag stage retires about one of them per cycle
// and SM.  2.2e8 elements / 148 SMs / 1.97 GHz = 0.75 ms -- exactly where both earlier kernels (register-staged loads,
// and a bulk-copy ring in shared memory) stopped, with `L1/TEX throughput 89 %` in ncu and HBM at 45-54 %
// (profiles/r02b_k3_ncu_summary.txt).  Shared memory serves a random 8-byte gather five to six times faster, but the
// vector (0.8 MB at 1e5 core determinants) does not fit.  So the matrix is cut into COLUMN BLOCKS of at most
// NG_SPMV_CB_MAX columns whose slice of v_full (<= 216 KB) does fit:
//   * set-up (k_spmv_block_count / k_spmv_block_fill): every row is partitioned stably by column block and the
//     matrix is re-laid block-major -- for block c the chunks of rows 0..n_local-1 one after the other -- as
//     {fp64 value, 16-bit column inside the block}: 10 bytes per element instead of CSR's 12;
//   * k_determ_spmv_blocked: one persistent CTA of 1024 threads per SM owns a contiguous 1/G-th of the elements
//     (so of the bytes), loads the v slice of its block into shared memory (200 KB from L2, 1 % of the CTA's traffic)
//     and its warps walk the (block, row) chunks: coalesced streaming loads of value and column, gather from shared
//     memory, one partial sum per chunk;
//   * k_determ_finish adds the partial sums of a row in block order (fixed order: bit-reproducible), the shift or
//     diagonal term, and multiplies by tau.
#ifndef NG_SPMV_CB_MAX
#define NG_SPMV_CB_MAX 27648          /* columns per block: 216 KB of fp64 in shared memory */
#endif
#ifndef NG_SPMV_CTAS
#define NG_SPMV_CTAS 1                /* resident CTAs per SM (2 needs NG_SPMV_CB_MAX <= 13824) */
#endif
#define NG_SPMV_THREADS 1024
#define NG_SPMV_NB_MAX 256            /* column blocks (set-up histogram); 7e6 core determinants */
// set-up, pass 1: elements of every (block, row) chunk.  One warp per row; cnt is block-major [c * n_local + i].
__global__ void __launch_bounds__(256) k_spmv_block_count(const long long *__restrict__ row_ptr, const int *__restrict__ col, long long n_local,
                                                          int cb, int nb, long long *__restrict__ cnt) {
    __shared__ int s_cnt[8][NG_SPMV_NB_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n_local; i += nwarps) {
        for (int c = lane; c < nb; c += 32) s_cnt[wib][c] = 0;
        __syncwarp();
        for (long long k = row_ptr[i] + lane; k < row_ptr[i + 1]; k += 32) atomicAdd(&s_cnt[wib][col[k] / cb], 1);
        __syncwarp();
        for (int c = lane; c < nb; c += 32) cnt[(size_t)c * n_local + i] = s_cnt[wib][c];
        __syncwarp();
    }
}
// set-up, pass 2: stable scatter of every row into its chunks (bptr = exclusive prefix sum of cnt in block-major order)
__global__ void __launch_bounds__(256) k_spmv_block_fill(const long long *__restrict__ row_ptr, const int *__restrict__ col,
                                                         const double *__restrict__ val, long long n_local, int cb, int nb,
                                                         const long long *__restrict__ bptr, double *__restrict__ bval,
                                                         unsigned short *__restrict__ bcol) {
    __shared__ long long s_off[8][NG_SPMV_NB_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n_local; i += nwarps) {
        for (int c = lane; c < nb; c += 32) s_off[wib][c] = bptr[(size_t)c * n_local + i];
        __syncwarp();
        const long long b = row_ptr[i], e = row_ptr[i + 1];
        for (long long k0 = b; k0 < e; k0 += 32) {
            const long long k = k0 + lane;
            const bool has = k < e;
            const int cj = has ? col[k] : 0;
            const int c = has ? cj / cb : -1;
            const u32 peers = __match_any_sync(0xffffffffu, c);
            if (has) {
                const long long pos = s_off[wib][c] + __popc(peers & lt);
                bval[pos] = val[k]; bcol[pos] = (unsigned short)(cj - c * cb);
            }
            __syncwarp();
            if (has && lane == __ffs(peers) - 1) s_off[wib][c] += __popc(peers);
            __syncwarp();
        }
        // chunks are padded to whole quadruples with zero elements (bptr holds the padded positions)
        for (int c = lane; c < nb; c += 32)
            for (long long pos = s_off[wib][c]; pos < bptr[(size_t)c * n_local + i + 1]; ++pos) { bval[pos] = 0.0; bcol[pos] = 0; }
        __syncwarp();
    }
}
// set-up, pass 3: bank-aware order inside every chunk.  The kernel's gathers read 8 bytes per lane from the vector slice
// in shared memory; the 16 lanes of a half-warp conflict when their columns agree modulo 16 (same pair of banks), and
// random columns do so 4-way on average.  Within a chunk the order of the elements is free (only the order of the
// additions changes, deterministically), so they are dealt out round-robin over the 16 residues: the j-th element of
// every residue class forms one group of 16 consecutive elements -- conflict-free as long as all classes still have
// elements -- and a half-warp of a trip reads exactly one such group (trips start at multiples of 16 from the chunk's
// start).  One warp per chunk; sizes of the residue classes first, then positions.
__global__ void __launch_bounds__(256) k_spmv_bank_order(const long long *__restrict__ bptr, long long nchunk, const double *__restrict__ bval,
                                                         const unsigned short *__restrict__ bcol, double *__restrict__ oval,
                                                         unsigned short *__restrict__ ocol) {
    __shared__ int s_cnt[8][16], s_run[8][16];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < nchunk; c += nwarps) {
        const long long s = bptr[c], e = bptr[c + 1];
        if (lane < 16) { s_cnt[wib][lane] = 0; s_run[wib][lane] = 0; }
        __syncwarp();
        for (long long k = s + lane; k < e; k += 32) atomicAdd(&s_cnt[wib][bcol[k] & 15], 1);
        __syncwarp();
        int cnt[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) cnt[r] = s_cnt[wib][r];
        for (long long k0 = s; k0 < e; k0 += 32) {
            const long long k = k0 + lane;
            const bool has = k < e;
            const unsigned short cj = has ? bcol[k] : 0;
            const int r = has ? (int)(cj & 15) : -1;
            const u32 peers = __match_any_sync(0xffffffffu, r);
            if (has) {
                const int j = s_run[wib][r] + __popc(peers & lt);     // rank inside the residue class, in chunk order
                int pos = 0;
#pragma unroll
                for (int q = 0; q < 16; ++q) pos += min(cnt[q], j) + ((q < r && cnt[q] > j) ? 1 : 0);
                oval[s + pos] = bval[k]; ocol[s + pos] = cj;
            }
            __syncwarp();
            if (has && lane == __ffs(peers) - 1) s_run[wib][r] += __popc(peers);
            __syncwarp();
        }
    }
}
// One trip of a warp = 256 consecutive elements, 8 per lane (lane, lane + 32, ...): sixteen coalesced streaming loads.
// They are volatile asm so that ptxas keeps them together in program order; trips go in pairs, the loads of the next
// one issued before the current one is consumed.  (Tried and measured slower on the same matrix, 0.50 ms for this
// kernel against: 256-bit value loads with packed column pairs, 4 to 8 quadruples per lane, 0.53-0.63 ms; a
// trip stream pipelined across chunk boundaries, 0.57 ms -- profiles/r02k_*, r02m_*.  What is left is latency: 32 warps
// per SM is all that 200 KB of shared memory per CTA allow, and the gathers see 4-way bank conflicts on average.)
struct SpmvTrip { double a[8]; unsigned short c[8]; };
__device__ __forceinline__ void spmv_trip_load(SpmvTrip &T, const double *bval, const unsigned short *bcol, long long k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) asm volatile("ld.global.cs.u16 %0, [%1];" : "=h"(T.c[u]) : "l"(bcol + k + 32 * u));
#pragma unroll
    for (int u = 0; u < 8; ++u) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(T.a[u]) : "l"(bval + k + 32 * u));
}
__device__ __forceinline__ double spmv_trip_dot(const SpmvTrip &T, u32 vs_addr) {
    double g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(g[u]) : "r"(vs_addr + 8u * (u32)T.c[u]));
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int u = 0; u < 8; u += 2) { t0 += T.a[u] * g[u]; t1 += T.a[u + 1] * g[u + 1]; }
    return t0 + t1;
}
// the tail of a chunk (< 256 elements), predicated
__device__ __forceinline__ double spmv_tail(const double *__restrict__ bval, const unsigned short *__restrict__ bcol, const double *vs,
                                            long long k, long long e) {
    double a[8]; int c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const long long kk = k + 32 * u;
        const bool p = kk < e;
        a[u] = p ? __ldcs(&bval[kk]) : 0.0;
        c[u] = p ? (int)__ldcs(&bcol[kk]) : 0;
    }
    double t = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) t += a[u] * vs[c[u]];
    return t;
}
__global__ void __launch_bounds__(NG_SPMV_THREADS, NG_SPMV_CTAS) k_determ_spmv_blocked(const long long *__restrict__ bptr, const unsigned short *__restrict__ bcol,
                                                                         const double *__restrict__ bval, const double *__restrict__ v_full,
                                                                         const long long *__restrict__ work, long long n_local,
                                                                         long long n_core, int cb, double *__restrict__ partial) {
    extern __shared__ __align__(16) double spmv_vs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = NG_SPMV_THREADS / 32;
    const u32 vs_addr = (u32)__cvta_generic_to_shared(spmv_vs);
    // chunks (flat block-major index c * n_local + i) whose first element lies in this CTA's share of the elements
    const long long idx_lo = work[blockIdx.x], idx_hi = work[blockIdx.x + 1];
    if (idx_lo >= idx_hi) return;
    for (long long c = idx_lo / n_local; c * n_local < idx_hi; ++c) {
        const long long a = max(idx_lo, c * n_local), b = min(idx_hi, (c + 1) * n_local);
        const long long col0 = c * cb;
        const int ncol = (int)min((long long)cb, n_core - col0);
        __syncthreads();                                       // the previous block's slice is still being read
        for (int j = threadIdx.x; j < ncol; j += NG_SPMV_THREADS) spmv_vs[j] = __ldg(&v_full[col0 + j]);
        __syncthreads();
        for (long long idx = a + warp; idx < b; idx += nwarp) {
            const long long s = __ldg(&bptr[idx]), e = __ldg(&bptr[idx + 1]);
            const long long nfull = (e - s) >> 8;
            double t = 0.0;
            long long k = s + lane;
            if (nfull > 0) {                                   // trips in pairs: the next one is requested before this one is consumed
                SpmvTrip A, B;
                spmv_trip_load(A, bval, bcol, k);
                long long n = 1;
                for (; n + 1 < nfull; n += 2) {
                    spmv_trip_load(B, bval, bcol, k + 256 * n); t += spmv_trip_dot(A, vs_addr);
                    spmv_trip_load(A, bval, bcol, k + 256 * (n + 1)); t += spmv_trip_dot(B, vs_addr);
                }
                if (n < nfull) { spmv_trip_load(B, bval, bcol, k + 256 * n); t += spmv_trip_dot(A, vs_addr); t += spmv_trip_dot(B, vs_addr); }
                else t += spmv_trip_dot(A, vs_addr);
                k += 256 * nfull;
            }
            if (k - lane < e) t += spmv_tail(bval, bcol, spmv_vs, k, e);
            t = warp_sum(t);
            if (lane == 0) partial[idx] = t;
        }
    }
}
// sum of a row's partial sums in block order + the shift term (determ_projection) or the diagonal element
// (determ_projection_no_death, semi_stoch_procs.F90:285-374: death then acts on the core determinants as well)
__global__ void k_determ_finish(const double *__restrict__ partial, const double *__restrict__ v_full, long long n_local, int nb,
                                long long displ, double tau, double diag_sft, const double *__restrict__ core_ham_diag,
                                double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x) {
        double t = 0.0;
        for (int c = 0; c < nb; ++c) t += partial[(size_t)c * n_local + i];
        const double d = core_ham_diag ? core_ham_diag[i] : diag_sft;
        out[i] = (-t + d * v_full[i + displ]) * tau;
    }
}
__global__ void k_determ_apply(WalkerList L, const int *core_slots, const double *out, long long n_local) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x)
        L.sgn[core_slots[i]] += out[i];
}

// hash table over the replicated core space (entry = index + 1)
template <int NW>
__global__ void k_core_ht_build(const long long *iluts, long long n, int *ht, u64 mask) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> d; d.w[0] = (u64)iluts[i * NW]; if (NW > 1) d.w[NW - 1] = (u64)iluts[i * NW + NW - 1];
        u64 pos = det_hash64(d) & mask;
        while (atomicCAS(&ht[pos], 0, (int)i + 1) != 0) pos = (pos + 1) & mask;
    }
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

// Sparse core Hamiltonian on the device (calc_determ_hamil_sparse, src/sparse_arrays.F90:426-572; row contents as
// calc_determ_hamil_opt leaves them, src/fast_determ_hamil.F90:1421-1507): one warp per local row, the lanes sweep
// the replicated core space 32 determinants at a time (coalesced), a popcount test drops everything more than a
// double excitation away, the survivors get their matrix element, and a ballot keeps the row in ascending column
// order.  FILL = false counts the non-zero elements of every row (+1 for the diagonal), FILL = true writes them at
// row_ptr[i] and closes the row with H_ii - Hii (also stored in core_ham_diag).
template <int NW, int SYS, bool FILL>
__global__ void __launch_bounds__(NG_BLOCK) k_core_ham(Params P, const long long *iluts, long long n_core, long long displ,
                                                       long long n_local, double hii, long long *row_ptr, int *col,
                                                       double *val, double *core_ham_diag) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n_local; i += nwarps) {
        const long long gi = displ + i;
        Det<NW> I; I.w[0] = (u64)iluts[gi * NW]; if (NW > 1) I.w[NW - 1] = (u64)iluts[gi * NW + NW - 1];
        long long pos = FILL ? row_ptr[i] : 0;
        for (long long j0 = 0; j0 < n_core; j0 += 32) {
            const long long j = j0 + lane;
            double h = 0.0;
            if (j < n_core && j != gi) {
                Det<NW> J; J.w[0] = (u64)iluts[j * NW]; if (NW > 1) J.w[NW - 1] = (u64)iluts[j * NW + NW - 1];
                if (sys_hphf(SYS)) h = hphf_off_diag<NW, SYS>(P, I, J);     // partner determinants: no cheap prefilter
                else {
                    int x = __popcll(I.w[0] ^ J.w[0]); if (NW > 1) x += __popcll(I.w[NW - 1] ^ J.w[NW - 1]);
                    if (x <= 4) h = helement<NW, SYS>(P, I, J);
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, h != 0.0);
            if (FILL && h != 0.0) {
                const long long k = pos + __popc(m & ((1u << lane) - 1u));
                col[k] = (int)j; val[k] = h;
            }
            pos += __popc(m);
        }
        if (lane == 0) {
            if (FILL) {
                const double d = diagonal_matel<NW, SYS>(P, I) - hii;
                col[pos] = (int)gi; val[pos] = d; core_ham_diag[i] = d;
            } else row_ptr[i] = pos + 1;
        }
    }
}

// ---- trial wavefunction -----------------------------------------------------------------
template <int NW>
__global__ void k_trial_ht_build(const long long *iluts, long long n, int *ht, u64 mask) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Det<NW> d; d.w[0] = (u64)iluts[i * NW]; if (NW > 1) d.w[NW - 1] = (u64)iluts[i * NW + NW - 1];
        u64 pos = det_hash64(d) & mask;
        while (atomicCAS(&ht[pos], 0, (int)i + 1) != 0) pos = (pos + 1) & mask;
    }
}
// flags + current_trial_amps for the resident list (what init_trial_wf does for CurrentDets)
template <int NW>
__global__ void k_trial_locate(Params P, WalkerList L) {
    const long long n = L.ctr[C_NLIST];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double amp;
        const int k = trial_lookup<NW>(P, load_det<NW>(L, i), &amp);
        L.flg[i] = (L.flg[i] & ~(F_TRIAL | F_CONNECTED)) | k;
        L.trial_amp[i] = amp;
    }
}
// trial part of SumEContrib (src/fcimc_helper.F90:586-648, ntrial_excits = 1, no qmc_trial_wf): runs before the
// spawning pass, i.e. on the signs the walker loop sees
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_trial_energy(Params P, WalkerList L, double *partials) {
    __shared__ double s_red[4 * 32];
    const Det<NW> ref = ref_det<NW>(P);
    const long long n = L.ctr[C_NLIST];
    double acc[4] = {0, 0, 0, 0};           // numerator, denominator, initiator numerator, initiator denominator
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int f = L.flg[i];
        if (!(f & (F_TRIAL | F_CONNECTED))) continue;
        const double s = L.sgn[i];
        if (fabs(s) < 1.0e-12) continue;
        // CalcParentFlag runs before SumEContrib, so the initiator flag is this iteration's
        bool init = (f & F_INIT) != 0;
        if (P.t_trunc_initiator) {
            const int exl = P.t_hphf ? excit_level_ref<NW, true>(ref, load_det<NW>(L, i)) : excit_level(ref, load_det<NW>(L, i));
            init = parent_is_initiator(P, init, fabs(s), exl, (f & F_DETERM) != 0);
        }
        const double c = L.trial_amp[i] * s;
        if (f & F_TRIAL) { acc[1] += c; if (init) acc[3] += c; }
        else { acc[0] += c; if (init) acc[2] += c; }
    }
    const int idx[4] = {NECI_ST_TRIAL_NUMERATOR, NECI_ST_TRIAL_DENOM, NECI_ST_INIT_TRIAL_NUMERATOR, NECI_ST_INIT_TRIAL_DENOM};
    block_flush_stats<4>(acc, idx, partials, s_red);
}

// log_death_magnitude (src/tau/tau_main.F90:198-207, called from attempt_die for every determinant, core ones
// included): max (K_ii - S) over the list the walker loop is about to see.  Its own streaming pass, launched only
// when the tau search is on -- inside the spawning kernel it cost 3 % of every run.
__global__ void __launch_bounds__(NG_BLOCK) k_death_magnitude(WalkerList L, double diag_sft, double *partials) {
    __shared__ double s_red[32];
    const long long n = L.ctr[C_NLIST];
    double acc[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (fabs(L.sgn[i]) >= 1.0e-12) acc[0] = fmax(acc[0], L.diagH[i] - diag_sft);
    const int idx[1] = {NECI_ST_TAU_MAX_DEATH_CPT};
    block_flush_stats<1>(acc, idx, partials, s_red);
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
        if (sys_hphf(SYS) && !det_eq(a, b)) out[i] = hphf_off_diag<NW, SYS>(P, a, b);
        else out[i] = helement<NW, SYS>(P, a, b);
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
        double rh_hphf = 0.0;
        if (sys_hphf(SYS) && E.valid) E.valid = hphf_fixup<NW, SYS>(P, d, E, rh_hphf);
        if (E.valid) {
            ilut_j[i * NW] = (long long)E.detJ.w[0]; if (NW > 1) ilut_j[i * NW + NW - 1] = (long long)E.detJ.w[NW - 1];
            ex[4 * i] = E.src1; ex[4 * i + 1] = E.src2; ex[4 * i + 2] = E.tgt1; ex[4 * i + 3] = E.tgt2;
            par[i] = E.parity ? 1 : 0; pgen[i] = E.pgen;
            hel[i] = sys_hphf(SYS) ? rh_hphf : spawn_helement<NW, SYS>(P, d, E);
        } else {
            for (int w = 0; w < NW; ++w) ilut_j[i * NW + w] = 0;
            ex[4 * i] = ex[4 * i + 1] = ex[4 * i + 2] = ex[4 * i + 3] = 0;
            par[i] = 0; pgen[i] = 0.0; hel[i] = 0.0;
        }
    }
}

// per-block walker populations for adjust_load_balance (load_balancer.fpp:216-235): sum(ceiling(abs(sgn)))
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
        atomicAdd(&block_parts[b - 1], ceil(fabs(s)));
    }
}

// Reservation of positions in per-destination segments for a CTA's records: the counts per destination are taken in
// shared memory (one shared atomic per warp and destination) and ONE thread per destination reserves in the global
// counter.  A global atomic per warp and destination -- eight of them on one cache line per warp at N = 8 -- is what
// made the round-1 exchange grow with the number of ranks (0.08 / 0.14 / 0.21 ms at 2 / 4 / 8 GPUs): the L2 atomic unit
// serialises them.  Records of one CTA and destination also become one contiguous run (~32 records at N = 8).
// All threads of the CTA must call.  Returns the position, or -1 without a record.
struct DestReserveScratch { int hist[NG_MAX_PUSH_RANKS]; unsigned long long base[NG_MAX_PUSH_RANKS]; };
__device__ __forceinline__ long long dest_reserve(DestReserveScratch &R, unsigned long long *cnt, int nranks, bool has, int proc) {
    const u32 lane = threadIdx.x & 31;
    if ((int)threadIdx.x < nranks) R.hist[threadIdx.x] = 0;
    __syncthreads();
    int rank_in = 0;
    {
        const u32 active = __ballot_sync(0xffffffffu, has);
        if (has) {
            const u32 peers = __match_any_sync(active, proc);
            const int leader = __ffs(peers) - 1;
            int wbase = 0;
            if ((int)lane == leader) wbase = atomicAdd(&R.hist[proc], __popc(peers));
            wbase = __shfl_sync(peers, wbase, leader);
            rank_in = wbase + __popc(peers & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nranks) {
        const int c = R.hist[threadIdx.x];
        R.base[threadIdx.x] = c ? atomicAdd(&cnt[threadIdx.x], (unsigned long long)c) : 0ull;
    }
    __syncthreads();
    return has ? (long long)R.base[proc] + rank_in : -1;
}

// move_block (load_balancer.fpp:353-512), sender side, for all moved blocks at once: every occupied
// determinant whose owner under the NEW mapping is another rank is appended to that rank's segment of the
// spawn buffer (the wire format of the reference's MPISend: ilut incl. sign and flags) and removed here.
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_rebalance_pack(Params P, WalkerList L, SpawnBuf SB) {
    __shared__ int s_roi[NG_MAX_BASIS];
    __shared__ DestReserveScratch R;
    for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    const long long n = L.ctr[C_NLIST];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        bool move = false;
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        double s = 0.0; int f = 0;
        if (i < n) {
            s = L.sgn[i];
            if (fabs(s) >= 1.0e-12) {
                d = load_det<NW>(L, i);
                f = L.flg[i];
                const int owner = __ldg(&P.lb_mapping[det_block<NW>(P, s_roi, d) - 1]);
                if (owner != P.rank) {
                    move = true;
                    ht_remove<NW>(L, d, det_hash64(d), i);
                    L.sgn[i] = 0.0; L.flg[i] = f | F_REMOVED;
                }
            }
        }
        int proc = 0;
        if (move) proc = __ldg(&P.lb_mapping[det_block<NW>(P, s_roi, d) - 1]);
        const long long pos = dest_reserve(R, SB.cnt, P.nranks, move, proc);
        if (move) {
            if (pos >= SB.seg_cap) atomicOr((unsigned long long *)&L.ctr[C_ERR], 1ull);
            else {
                long long *rec = SB.buf + ((size_t)proc * SB.seg_cap + pos) * SB.W;
                rec[0] = (long long)d.w[0]; if (NW > 1) rec[NW - 1] = (long long)d.w[NW - 1];
                rec[NW] = __double_as_longlong(s); rec[NW + 1] = (long long)(f & ~F_REMOVED);
            }
        }
    }
}
// DetermineDetNode for the staged spawns of one iteration (nranks > 1): every lane hashes one spawn and appends it
// to its destination's segment of SpawnedParts (create_particle's routing, src/fcimc_helper.F90:152-308).
template <int NW>
__global__ void __launch_bounds__(NG_BLOCK) k_partition(Params P, WalkerList L, SpawnBuf SB) {
    __shared__ int s_roi[NG_MAX_BASIS];
    __shared__ DestReserveScratch R;
    for (int i = threadIdx.x; i < P.nbasis; i += blockDim.x) s_roi[i] = P.random_orb_index[i];
    __syncthreads();
    long long n = (long long)*SB.stage_cnt; if (n > SB.stage_cap) n = SB.stage_cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nloop = ((n + stride - 1) / stride) * stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nloop; i += stride) {
        Det<NW> d; d.w[0] = 0; if (NW > 1) d.w[NW - 1] = 0;
        double s = 0.0; long long f = 0;
        const bool has = i < n;
        if (has) {
            const long long *rec = SB.stage + (size_t)i * SB.W;
            d.w[0] = (u64)rec[0]; if (NW > 1) d.w[NW - 1] = (u64)rec[NW - 1];
            s = __longlong_as_double(rec[NW]); f = rec[NW + 1];
        }
        int proc = 0;
        if (has) proc = __ldg(&P.lb_mapping[det_block<NW>(P, s_roi, d) - 1]);
        const long long pos = dest_reserve(R, SB.cnt, P.nranks, has, proc);
        if (has) {
            if (pos >= SB.seg_cap) atomicOr((unsigned long long *)&L.ctr[C_ERR], 1ull);
            else {
                long long *rec = SB.buf + ((size_t)proc * SB.seg_cap + pos) * SB.W;
                rec[0] = (long long)d.w[0]; if (NW > 1) rec[NW - 1] = (long long)d.w[NW - 1];
                rec[NW] = __double_as_longlong(s); rec[NW + 1] = f;
            }
        }
    }
}
// receiver side: every received record is a new determinant here
__global__ void k_iota_insert(WalkerList L, SpawnBuf SB, long long n) {
    if (n < 0) n = (long long)*SB.n_recv_dev;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) SB.ins_idx[i] = (int)i;
    if (blockIdx.x == 0 && threadIdx.x == 0) L.ctr[C_NINSERT] = n;
}

// ---- spawn exchange over NVLink peer memory (SendProcNewParts, src/Annihilation.F90:150-247) -----------------
// Every rank owns an inbox that its peers can write: per parity (two exchanges can be in flight) and per source rank
// one segment of seg_cap records plus one 64-bit mailbox word (sequence number << 32 | record count).
//   k_push    copies this rank's per-destination segments of SpawnedParts straight into the destinations' inbox
//             segments (coalesced remote stores), then -- once every CTA has fenced its stores -- posts the mailboxes
//   k_wait    spins until all sources have posted this exchange's sequence number; leaves counts / offsets on the device
//   k_gather  compacts the inbox segments into the contiguous receive list (source-rank order, like MPI_Alltoallv)
// No host synchronisation and no count round trip: the kernels that follow read the count from device memory.
struct PeerBox {
    long long **peer_seg;            // [nranks] base of rank r's inbox segments  (device array of peer pointers)
    unsigned long long **peer_mail;  // [nranks] base of rank r's mailboxes
    long long *my_seg;               // this rank's inbox segments [2][nranks][seg_cap][W]
    unsigned long long *my_mail;     // this rank's mailboxes [2][nranks]
    unsigned long long *cnt_in;      // [nranks] counts of the current exchange, [nranks .. 2 nranks) offsets
    unsigned int *ticket;
};
__global__ void __launch_bounds__(256) k_push(SpawnBuf SB, PeerBox X, int nranks, int rank, unsigned int seq) {
    __shared__ bool s_last;
    const int par = (int)(seq & 1u);
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
    for (int dst = 0; dst < nranks; ++dst) {
        long long n = (long long)SB.cnt[dst]; if (n > SB.seg_cap) n = SB.seg_cap;      // overflow is reported by K1
        const long long words = n * SB.W;
        const long long *src = SB.buf + (size_t)dst * SB.seg_cap * SB.W;
        long long *out = X.peer_seg[dst] + ((size_t)par * nranks + rank) * SB.seg_cap * SB.W;
        for (long long i = gtid; i < words; i += gstride) out[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(X.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if ((int)threadIdx.x < nranks) {
            const int dst = threadIdx.x;
            unsigned long long n = SB.cnt[dst]; if (n > (unsigned long long)SB.seg_cap) n = (unsigned long long)SB.seg_cap;
            volatile unsigned long long *mail = X.peer_mail[dst] + (size_t)par * nranks + rank;
            *mail = ((unsigned long long)seq << 32) | n;
        }
        if (threadIdx.x == 0) *X.ticket = 0u;
        __threadfence_system();
    }
}
// Mailbox hand-shake of an exchange.  post = true (spawning pass): the spawning kernels have already routed and pushed
// their spawns into the owners' inboxes (spawn_stage_push, spawn_kernel.cuh), so this rank first posts its counts --
// the stores of the finished kernels were fenced system-wide by their threads -- and then waits like k_push's partner.
__global__ void k_wait(WalkerList L, SpawnBuf SB, PeerBox X, int nranks, int rank, unsigned int seq, bool post, long long timeout_cycles) {
    __shared__ unsigned long long s_cnt[64];
    const int par = (int)(seq & 1u);
    if (post && (int)threadIdx.x < nranks) {
        const int dst = threadIdx.x;
        const unsigned long long c = SB.push_cnt[NG_PUSH_CNT_STRIDE * dst];
        SB.cnt[dst] = c;                                    // ValidSpawnedList - InitialSpawnedSlots, for the statistics
        if (c > (unsigned long long)SB.seg_cap) atomicOr((unsigned long long *)&L.ctr[C_ERR], 1ull);
        volatile unsigned long long *mail = X.peer_mail[dst] + (size_t)par * nranks + rank;
        *mail = ((unsigned long long)seq << 32) | min(c, (unsigned long long)SB.seg_cap);
        __threadfence_system();
    }
    if ((int)threadIdx.x < nranks) {
        volatile unsigned long long *mail = X.my_mail + (size_t)par * nranks + threadIdx.x;
        const long long t0 = clock64();
        unsigned long long v = *mail;
        while ((unsigned int)(v >> 32) != seq) {
            if (clock64() - t0 > timeout_cycles) { atomicOr((unsigned long long *)&L.ctr[C_ERR], 256ull); v = (unsigned long long)seq << 32; break; }
            __nanosleep(200);
            v = *mail;
        }
        s_cnt[threadIdx.x] = v & 0xFFFFFFFFull;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int s = 0; s < nranks; ++s) { X.cnt_in[s] = s_cnt[s]; X.cnt_in[nranks + s] = run; run += s_cnt[s]; }
        *SB.n_recv_dev = run;
    }
}
__global__ void __launch_bounds__(256) k_gather(SpawnBuf SB, PeerBox X, int nranks, unsigned int seq) {
    const int par = (int)(seq & 1u);
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < nranks; ++s) {
        const long long words = (long long)X.cnt_in[s] * SB.W;
        const long long *src = X.my_seg + ((size_t)par * nranks + s) * SB.seg_cap * SB.W;
        long long *out = SB.recv + (size_t)X.cnt_in[nranks + s] * SB.W;
        for (long long i = gtid; i < words; i += gstride) out[i] = src[i];
    }
}

}  // namespace ng
