// Device-side Hamiltonian matrix elements and excitation generators, written
// directly on the occupation bit-strings (no nI lists, no per-thread arrays):
//   Slater-Condon rules          src/sltcnd.fpp:585-708
//   UMAT / TMAT access           src/UMatCache.F90:257-296, src/OneEInts.F90:188-226
//   PCHB doubles + uniform singles
//        src/gasci_pchb_doubles_spatorb_fastweighted.fpp:155-277,
//        src/excit_gens_int_weighted.F90:722-840, src/aliasSampling.F90:310-332,
//        src/GenRandSymExcitNUMod.F90:1118-1286, src/excitation_generators.F90:112-138
//   real-space Hubbard           src/real_space_hubbard.F90:1938-2142,2303-2321,2425-2457
//   k-space Hubbard              src/k_space_hubbard.F90:356-374,535-625,1712-1817,2620-2671
//   DetermineDetNode             src/load_balance_calcnodes.F90:25-117
#pragma once
#include "device_common.cuh"

namespace ng {

#define NG_EPS 1e-13   /* src/lib/constants.F90:28 */

template <int NW> struct Excit {
    bool valid;
    int ic;
    int src1, src2, tgt1, tgt2;    // sorted as make_single / make_double return them
    bool parity;
    double pgen;
    Det<NW> detJ;
    int err;
};

// ---- integrals ---------------------------------------------------------------
// 1-based triangular index in unsigned 32-bit arithmetic (spatial orbitals <= 64 => pair index <= 2080,
// UMatInd <= 2 164 240: neci_gpu_set_system_fcidump checks n_umat < 2^31)
__device__ __forceinline__ u32 tri(u32 a, u32 b) {
    const u32 hi = max(a, b), lo = min(a, b);
    return ((hi * (hi - 1u)) >> 1) + lo;
}
// <ij|kl> over spatial orbitals (1-based): UMAT(UMatInd(i,j,k,l))
__device__ __forceinline__ double umat_el(const Params &P, int i, int j, int k, int l) {
    const u32 ind = tri(tri((u32)i, (u32)k), tri((u32)j, (u32)l));
    return __ldg(&P.umat[ind - 1u]);
}
__device__ __forceinline__ double tmat_el(const Params &P, int i, int j) {
    return __ldg(&P.tmat[(size_t)(i - 1) + (size_t)P.nbasis * (j - 1)]);
}
__device__ __forceinline__ double umat_k(const Params &P, int i, int j, int k, int l) {   // get_umat_kspace
    const int a = __ldg(&P.ksum[(i - 1) * P.n_k + (j - 1)]), b = __ldg(&P.ksum[(k - 1) * P.n_k + (l - 1)]);
    return (a == b) ? P.u_over_n : 0.0;
}

// parity of the double excitation {s1<s2} -> {t1<t2} (make_double, src/excit_parity.F90:78-170):
// orbitals jumped by s1->t1 in D, then by s2->t2 in D - s1 + t1.
template <int NW>
__device__ __forceinline__ bool parity_double(const Det<NW> &d, int s1, int s2, int t1, int t2) {
    int c = count_between(d, s1, t1);
    Det<NW> d2 = d; clr_orb(d2, s1); set_orb(d2, t1);
    c += count_between(d2, s2, t2);
    return c & 1;
}
template <int NW>
__device__ __forceinline__ bool parity_single(const Det<NW> &d, int s, int t) { return count_between(d, s, t) & 1; }

// ---- Slater-Condon -----------------------------------------------------------
// sltcnd_0 (src/sltcnd.fpp:585-622) in the reference's summation order.  The Coulomb <ij|ij> and exchange <ij|ji>
// integrals come from two n_spat x n_spat tables gathered from UMAT at upload (same values, so the sums are
// bit-identical to indexing UMAT through UMatInd) -- they stay in L1 and need one multiply-add of index math.
template <int NW>
__device__ double sltcnd_0(const Params &P, const Det<NW> &d) {
    double hel_sing = 0.0, hel_doub = 0.0, hel_tmp = 0.0;
    const int ns = P.n_spat_sys;
    Det<NW> a = d;
    while (det_any(a)) {
        const int oi = pop_lowest(a);
        hel_sing += tmat_el(P, oi, oi);
        const int idi = gtid(oi);
        const double *jrow = P.jmat + (size_t)(idi - 1) * ns - 1;
        const double *krow = P.kmat + (size_t)(idi - 1) * ns - 1;
        Det<NW> b = a;
        double s = 0.0;
        while (det_any(b)) {
            const int oj = pop_lowest(b);
            const int idj = gtid(oj);
            s += __ldg(&jrow[idj]);
            if (P.t_exch && ((oi ^ oj) & 1) == 0) hel_tmp -= __ldg(&krow[idj]);
        }
        hel_doub += s;
    }
    return hel_doub + hel_tmp + hel_sing;
}
// sltcnd_1_kernel.  The reference walks the occupied orbitals one by one; here the integral loads of four
// orbitals are issued together (predicated) so that the L2 round trips overlap, Coulomb and exchange terms
// are accumulated separately (agrees with the sequential order to ~1 ulp).
template <int NW>
__device__ double sltcnd_1(const Params &P, const Det<NW> &d, int src, int tgt) {
    const int id1 = gtid(src), id2 = gtid(tgt);
    double hc = 0.0, hx = 0.0;
    if (((src ^ tgt) & 1) == 0) {
        Det<NW> a = d; clr_orb(a, src);
        const bool exch = P.t_exch != 0;
        while (det_any(a)) {
            double c[4], x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = det_any(a);
                int o = src;
                if (v) o = pop_lowest(a);
                const int id = gtid(o);
                c[u] = v ? umat_el(P, id1, id, id2, id) : 0.0;
                x[u] = (v && exch && ((o ^ src) & 1) == 0) ? umat_el(P, id1, id, id, id2) : 0.0;
            }
            hc += (c[0] + c[1]) + (c[2] + c[3]);
            hx += (x[0] + x[1]) + (x[2] + x[3]);
        }
    }
    return (hc - hx) + tmat_el(P, src, tgt);
}
__device__ __forceinline__ double sltcnd_2(const Params &P, int s1, int s2, int t1, int t2) {
    double hel = 0.0;
    if ((((s1 ^ t1) | (s2 ^ t2)) & 1) == 0) hel = umat_el(P, gtid(s1), gtid(s2), gtid(t1), gtid(t2));
    if ((((s1 ^ t2) | (s2 ^ t1)) & 1) == 0) hel -= umat_el(P, gtid(s1), gtid(s2), gtid(t2), gtid(t1));
    return hel;
}
// k-space Hubbard double, get_offdiag_helement_k_sp_hub
__device__ __forceinline__ double offdiag_k(const Params &P, int s1, int s2, int t1, int t2, bool par) {
    if (((s1 ^ s2) & 1) == 0 || ((t1 ^ t2) & 1) == 0) return 0.0;
    double hel = umat_k(P, gtid(s1), gtid(s2), gtid(t1), gtid(t2));
    if (((s1 ^ t1) & 1) != 0) hel = -hel;
    if (fabs(hel) < NG_EPS) return hel;
    return par ? -hel : hel;
}
template <int NW>
__device__ double diag_k(const Params &P, const Det<NW> &d) {      // sltcnd_0 with get_umat_kspace, tExch
    double hs = 0.0;
    Det<NW> a = d;
    while (det_any(a)) hs += __ldg(&P.eps_k[gtid(pop_lowest(a)) - 1]);
    // every pair contributes U/N (Coulomb), same-spin pairs cancel by exchange:
    // summed explicitly in the reference's order to stay within 1e-12.
    const int n = popc(d);
    int nb = __popcll(d.w[0] & NG_BETA_MASK); if (NW > 1) nb += __popcll(d.w[NW - 1] & NG_BETA_MASK);
    const int na = n - nb;
    const double coul = (double)(n * (n - 1) / 2) * P.u_over_n;
    const double exch = (double)(na * (na - 1) / 2 + nb * (nb - 1) / 2) * P.u_over_n;
    return (coul - exch) + hs;
}
template <int NW>
__device__ __forceinline__ double diag_rs(const Params &P, const Det<NW> &d) {   // U * double occupancies
    int nd = __popcll(d.w[0] & (d.w[0] >> 1) & NG_BETA_MASK);
    if (NW > 1) nd += __popcll(d.w[NW - 1] & (d.w[NW - 1] >> 1) & NG_BETA_MASK);
    return P.uhub * nd;
}

// (defined with the other HPHF functions below; FCIDUMP systems only)
template <int NW, bool HPHF> __device__ __forceinline__ int excit_level_ref(const Det<NW> &ref, const Det<NW> &d);
template <int NW> __device__ double hphf_diag_dispatch(const Params &P, const Det<NW> &d);
template <int NW> __device__ double hphf_off_diag_dispatch(const Params &P, const Det<NW> &I, const Det<NW> &J);

// get_diagonal_matel (src/matel_getter.F90:30-58): full H_ii (ECore included)
template <int NW, int SYS>
__device__ __forceinline__ double diagonal_matel(const Params &P, const Det<NW> &d) {
    if (sys_hphf(SYS)) return hphf_diag_dispatch<NW>(P, d);
    if (SYS == NECI_SYS_HUBBARD_RS) return diag_rs(P, d);
    if (SYS == NECI_SYS_HUBBARD_K) return diag_k(P, d) + P.ecore;
    return sltcnd_0(P, d) + P.ecore;
}

// get_helement between two arbitrary determinants (ic <= 2), with the
// excitation I -> J and its parity derived from the bit-strings.
template <int NW, int SYS>
__device__ double helement(const Params &P, const Det<NW> &I, const Det<NW> &J) {
    Det<NW> S, T;
    S.w[0] = I.w[0] & ~J.w[0]; T.w[0] = J.w[0] & ~I.w[0];
    if (NW > 1) { S.w[NW - 1] = I.w[NW - 1] & ~J.w[NW - 1]; T.w[NW - 1] = J.w[NW - 1] & ~I.w[NW - 1]; }
    const int ic = popc(S);
    if (ic != popc(T) || ic > 2) return 0.0;
    if (ic == 0) return diagonal_matel<NW, SYS>(P, I);
    if (ic == 1) {
        const int s = pop_lowest(S), t = pop_lowest(T);
        const bool par = parity_single(I, s, t);
        double h;
        if (SYS == NECI_SYS_HUBBARD_RS) h = tmat_el(P, s, t);
        else if (SYS == NECI_SYS_HUBBARD_K) return 0.0;
        else h = sltcnd_1(P, I, s, t);
        return par ? -h : h;
    }
    const int s1 = pop_lowest(S), s2 = pop_lowest(S), t1 = pop_lowest(T), t2 = pop_lowest(T);
    const bool par = parity_double(I, s1, s2, t1, t2);
    if (SYS == NECI_SYS_HUBBARD_RS) return 0.0;
    if (SYS == NECI_SYS_HUBBARD_K) return offdiag_k(P, s1, s2, t1, t2, par);
    const double h = sltcnd_2(P, s1, s2, t1, t2);
    return par ? -h : h;
}
// get_off_diagonal_matel (src/matel_getter.F90:61-105)
template <int NW, int SYS>
__device__ __forceinline__ double off_diagonal_matel(const Params &P, const Det<NW> &d) {
    const Det<NW> ref = ref_det<NW>(P);
    const int ex = excit_level_ref<NW, sys_hphf(SYS)>(ref, d);
    if (ex == 2 || (ex == 1 && P.t_no_brillouin))
        return sys_hphf(SYS) ? hphf_off_diag_dispatch<NW>(P, ref, d) : helement<NW, SYS>(P, d, ref);
    return 0.0;
}

// get_det_block / DetermineDetNode with the RandomOrbIndex table staged in
// shared memory (roi).  Two's-complement wrap, Fortran mod and abs reproduced.
template <int NW>
__device__ __forceinline__ int det_block(int balance_blocks, u64 bb_magic, const int *roi, Det<NW> d) {
    u64 acc = 0; int i = 1;
    while (det_any(d)) {
        const int o = pop_lowest(d);
        acc = 1099511628211ull * acc + (u64)(long long)(roi[o - 1] * i);
        ++i;
    }
    // abs(mod(acc, balance_blocks)) with Fortran's truncating mod == |acc| mod balance_blocks; the 64-bit remainder
    // is taken with the host-computed reciprocal floor((2^64 - 1) / balance_blocks): q is the quotient or one less
    const u64 a = ((long long)acc < 0) ? (0ull - acc) : acc;       // |INT64_MIN| = 2^63 fits
    const u64 B = (u64)balance_blocks;
    const u64 q = __umul64hi(a, bb_magic);
    u64 r = a - q * B;
    if (r >= B) r -= B;
    return (int)r + 1;
}
template <int NW>
__device__ __forceinline__ int det_block(const Params &P, const int *roi, const Det<NW> &d) { return det_block<NW>(P.balance_blocks, P.bb_magic, roi, d); }

// TestInitiator_explicit (src/fcimc_helper.F90:1142-1243): the initiator flag of a parent for this iteration
__device__ __forceinline__ bool parent_is_initiator(const Params &P, bool initiator, double as, int exl, bool core) {
    const bool popInit = as > P.initiator_walk_no;
    if (!initiator) return popInit;
    if (exl != 0 && !(core && P.t_core_inits) && !popInit) return false;
    return true;
}

// ---- generators ----------------------------------------------------------------
// pick_from_cum_list over an on-the-fly cumulative list of `n` equal or unequal
// weights is specialised per generator below.

// gen_excit_rs_hubbard
template <int NW>
__device__ void gen_rs_hubbard(const Params &P, const Det<NW> &d, Stream &rng, Excit<NW> &E) {
    E.ic = 1; E.valid = false; E.err = 0; E.pgen = 0.0;
    const int elec = 1 + (int)(rng.draw32() * P.nel);
    const double p_elec = 1.0 / (double)P.nel;
    const int src = select_orb(d, ~0ull, elec);
    const int *ng = P.neighbours + (size_t)(src - 1) * P.max_neigh;
    double cum[8]; int nb[8];
    double cum_sum = 0.0; int nn = 0;
    for (int i = 0; i < P.max_neigh && i < 8; ++i) {
        const int o = __ldg(&ng[i]);
        if (o == 0) break;
        double elem = 0.0;
        if (!occ(d, o)) elem = fabs(tmat_el(P, src, o));
        cum_sum += elem; cum[i] = cum_sum; nb[i] = o; nn = i + 1;
    }
    if (cum_sum < NG_EPS) return;
    const double r = rng.draw53() * cum_sum;
    if (cum[nn - 1] < r) return;
    // binary_search_first_ge over <= 8 entries == first index with cum >= r
    int ind = 0;
    while (cum[ind] < r) ++ind;
    const double p_orb = (ind == 0) ? cum[0] / cum_sum : (cum[ind] - cum[ind - 1]) / cum_sum;
    const int orb = nb[ind];
    E.pgen = p_elec * p_orb;
    E.src1 = src; E.src2 = 0; E.tgt1 = orb; E.tgt2 = 0;
    E.valid = true;
}

// bits at even positions of x gathered into the low 32 bits / the inverse
__device__ __forceinline__ u64 compress_even(u64 x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return x;
}
__device__ __forceinline__ u64 spread_even(u64 x) {
    x &= 0x00000000FFFFFFFFull;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// create_ab_list_hubbard + pick_from_cum_list exactly as the reference walks them (two sweeps over all orbitals), for
// the one case the closed form below does not cover: a random number of exactly zero.
template <int NW>
__device__ __noinline__ void gen_k_hubbard_sweep(const Params &P, const Det<NW> &d, int s1, int s2, double r_draw, Excit<NW> &E) {
    E.ic = 2; E.valid = false; E.err = 0; E.pgen = 0.0;
    const double p_elec = 1.0 / (double)(P.nocc_beta * P.nocc_alpha);
    const int kij = __ldg(&P.ksum[(gtid(s1) - 1) * P.n_k + (gtid(s2) - 1)]);
    // create_ab_list_hubbard: cumulative list over a = 1..nbasis; elem = excit_cache(i,j,a)
    // = U/N when a, b are empty (b = k_i + k_j - k_a with the spin opposite to a)
    const double w = fabs(P.u_over_n);
    double cum_sum = 0.0;
    for (int a = 1; a <= P.nbasis; ++a) {
        double elem = 0.0;
        if (!occ(d, a)) {
            const int kb = __ldg(&P.kdiff[kij * P.n_k + (gtid(a) - 1)]);
            const int b = 2 * (kb + 1) - ((a & 1) ? 0 : 1);
            if (b != a && !occ(d, b)) elem = w;
        }
        cum_sum += elem;
    }
    if (cum_sum < NG_EPS) return;
    const double r = r_draw * cum_sum;
    // second sweep: first a with cum(a) >= r  (binary_search_first_ge on the same sums)
    double c = 0.0, prev = 0.0; int ind = 0, bsel = 0;
    for (int a = 1; a <= P.nbasis; ++a) {
        double elem = 0.0; int b = -1;
        if (!occ(d, a)) {
            const int kb = __ldg(&P.kdiff[kij * P.n_k + (gtid(a) - 1)]);
            b = 2 * (kb + 1) - ((a & 1) ? 0 : 1);
            if (b != a && !occ(d, b)) elem = w;
        }
        prev = c; c += elem;
        if (!(c < r)) { ind = a; bsel = b; break; }
    }
    if (ind == 0) return;
    double p_orb = (ind == 1) ? c / cum_sum : (c - prev) / cum_sum;
    p_orb = 2.0 * p_orb;
    if (bsel <= 0) { E.pgen = 0.0; return; }       // r == 0 corner: zero-weight first entry
    const int t1 = min(ind, bsel), t2 = max(ind, bsel);
    E.src1 = s1; E.src2 = s2; E.tgt1 = t1; E.tgt2 = t2;
    E.pgen = p_elec * p_orb;
    E.valid = true;
}

// gen_excit_k_space_hub (src/k_space_hubbard.F90:535-625).  The reference builds a cumulative list over all nBasis
// orbitals a (weight |U/N| when a and its momentum partner b = k_i + k_j - k_a of the other spin are both empty) and
// picks from it.  All weights are equal, so the k-th partial sum depends on k alone: the host tabulates the running
// sums C[k] = C[k-1] + |U/N| in the reference's own order of additions, the allowed orbitals are a bit mask
// (empty, and not the partner of an occupied orbital: one table look-up per ELECTRON instead of two per ORBITAL), and
// the pick is "the k-th set bit with C[k] the first partial sum >= r" -- the same orbital and the same probability,
// bit for bit, in ~400 instead of ~2200 instructions.
template <int NW>
__device__ void gen_k_hubbard(const Params &P, const Det<NW> &d, Stream &rng, Excit<NW> &E) {
    E.ic = 2; E.valid = false; E.err = 0; E.pgen = 0.0;
    // pick_spin_opp_elecs (src/lattice_models_utils.F90:123-148) without its rejection loop: the pair index from one
    // number, alpha electron = index mod nOccAlpha, beta electron = index / nOccAlpha, counted in orbital order --
    // the distribution (every opposite-spin pair with 1 / (nOccAlpha nOccBeta)) and p_elec of the reference's loop,
    // which on a warp ran until the slowest lane had its pair (a third of the lanes active, 60 % of this generator's
    // instructions in profiles/r02m_hubk_k1_*).  The CPU checker draws the same way.
    int s1, s2;
    {
        const int nA = P.nocc_alpha, npair = nA * P.nocc_beta;
        const int idx = min((int)(rng.draw32() * (double)npair), npair - 1);
        const int ib = idx / nA, ia = idx - ib * nA;
        const int oa = select_orb(d, NG_ALPHA_MASK, ia + 1), ob = select_orb(d, NG_BETA_MASK, ib + 1);
        s1 = min(oa, ob); s2 = max(oa, ob);
    }
    const bool failed = false;
    const double p_elec = 1.0 / (double)(P.nocc_beta * P.nocc_alpha);
    const int kij = __ldg(&P.ksum[(gtid(s1) - 1) * P.n_k + (gtid(s2) - 1)]);
    const int *kd = P.kdiff + (size_t)kij * P.n_k;
    // Orbitals whose partner is occupied.  In spatial-orbital masks (bit k = k-point k) the beta orbitals blocked are
    // the image of the occupied ALPHA k-points under k -> k_i + k_j - k, and vice versa; the host tabulates that
    // permutation per pair momentum and per byte of the mask, so the image is ceil(n_k / 8) table look-ups instead
    // of one look-up per electron.
    u64 occB = compress_even(d.w[0]), occA = compress_even(d.w[0] >> 1);
    if (NW > 1) { occB |= compress_even(d.w[NW - 1]) << 32; occA |= compress_even(d.w[NW - 1] >> 1) << 32; }
    const u64 *perm = P.kperm + (size_t)kij * P.kperm_bytes * 256;
    u64 blockedB = 0, blockedA = 0;
    for (int j = 0; j < P.kperm_bytes; ++j) {
        blockedB |= __ldg(&perm[j * 256 + (int)((occA >> (8 * j)) & 255ull)]);
        blockedA |= __ldg(&perm[j * 256 + (int)((occB >> (8 * j)) & 255ull)]);
    }
    const u64 all_k = (P.n_k < 64) ? ((1ull << P.n_k) - 1ull) : ~0ull;
    const u64 allowB = ~occB & ~blockedB & all_k, allowA = ~occA & ~blockedA & all_k;
    Det<NW> allowed;
    allowed.w[0] = spread_even(allowB & 0xFFFFFFFFull) | (spread_even(allowA & 0xFFFFFFFFull) << 1);
    if (NW > 1) allowed.w[NW - 1] = spread_even(allowB >> 32) | (spread_even(allowA >> 32) << 1);
    if (failed) { E.err = 1; return; }
    const int n = popc(allowed);
    if (n == 0) return;
    const double cum_sum = __ldg(&P.kcum[n]);
    if (cum_sum < NG_EPS) return;
    const double u = rng.draw53();
    const double r = u * cum_sum;
    if (!(r > 0.0)) { gen_k_hubbard_sweep<NW>(P, d, s1, s2, u, E); return; }
    // smallest k >= 1 with C[k] >= r (binary_search_first_ge on the reference's list)
    int k = (int)(r / fabs(P.u_over_n));
    k = max(1, min(k, n));
    while (k < n && __ldg(&P.kcum[k]) < r) ++k;
    while (k > 1 && !(__ldg(&P.kcum[k - 1]) < r)) --k;
    const double c = __ldg(&P.kcum[k]), prev = __ldg(&P.kcum[k - 1]);
    const int ta = select_orb(allowed, ~0ull, k);
    const int kb = __ldg(&kd[gtid(ta) - 1]);
    const int tb = 2 * (kb + 1) - ((ta & 1) ? 0 : 1);
    double p_orb = (ta == 1) ? c / cum_sum : (c - prev) / cum_sum;
    p_orb = 2.0 * p_orb;
    E.src1 = s1; E.src2 = s2; E.tgt1 = min(ta, tb); E.tgt2 = max(ta, tb);
    E.pgen = p_elec * p_orb;
    E.valid = true;
}

// CreateSingleExcit (uniform singles)
template <int NW>
__device__ void gen_uniform_single(const Params &P, const Det<NW> &d, Stream &rng, Excit<NW> &E) {
    E.ic = 1; E.valid = false; E.err = 0; E.pgen = 0.0;
    // construct_class_counts + CheckIfSingleExcits via class masks (no per-thread arrays: the count of
    // empty orbitals of a class is recomputed from the masks in the constant bank when needed)
    int ElecsWNoExcits = 0;
#pragma unroll 1
    for (int c = 0; c < P.n_classes; ++c) {
        int o = __popcll(d.w[0] & P.class_mask[c][0]);
        int t = __popcll(P.class_mask[c][0]);
        if (NW > 1) { o += __popcll(d.w[NW - 1] & P.class_mask[c][1]); t += __popcll(P.class_mask[c][1]); }
        if (t - o == 0) ElecsWNoExcits += o;
    }
    if (ElecsWNoExcits == P.nel) return;
    // rejection loops without a `return` inside and with explicit re-convergence (see gen_k_hubbard)
    const unsigned conv = __activemask();
    bool failed = false;
    int src = 0, cls = 0, NExcit = 0, attempts = 0;
    for (;;) {
        const int Eleci = (int)(P.nel * rng.draw32()) + 1;
        src = select_orb(d, ~0ull, Eleci);
        cls = __ldg(&P.class_of_spinorb[src - 1]);
        NExcit = __popcll(P.class_mask[cls][0]) - __popcll(d.w[0] & P.class_mask[cls][0]);
        if (NW > 1) NExcit += __popcll(P.class_mask[cls][1]) - __popcll(d.w[NW - 1] & P.class_mask[cls][1]);
        if (NExcit != 0) break;
        if (attempts > 250) { failed = true; break; }
        ++attempts;
    }
    __syncwarp(conv);
    const int cs = __ldg(&P.class_start[cls]), nOrbs = __ldg(&P.class_start[cls + 1]) - cs;
    int Orb = 0; attempts = 0;
    for (;;) {
        if (failed) break;
        const int ChosenUnocc = (int)(nOrbs * rng.draw32());
        Orb = __ldg(&P.class_orbs[cs + ChosenUnocc]);
        if (!occ(d, Orb)) break;
        if (attempts > 250) { failed = true; break; }
        ++attempts;
    }
    __syncwarp(conv);
    if (failed) { E.err = 1; return; }
    E.src1 = src; E.src2 = 0; E.tgt1 = Orb; E.tgt2 = 0;
    const double pDoubNew = 1.0 - P.p_singles;
    double pgen = (1 - pDoubNew) / ((double)(NExcit * (P.nel - ElecsWNoExcits)));
    pgen = pgen / P.p_singles;
    E.pgen = pgen;
    E.valid = true;
}

// pick_biased_elecs + GAS_doubles_PCHB_gen_exc.
// Random numbers.  The reference draws four numbers per double excitation (single or double, the electron pair,
// exchange or not, the alias sample).  Here one attempt takes ONE Philox block: the first 53-bit number decides
// single / double and, rescaled to [0,1) by the caller (r = (u - pSingles) / (1 - pSingles)), picks the pair --
// the rescaling pick_biased_elecs itself applies to its own number, (r / pParallel) * nPairs -- and the fraction
// left over after the pair index has been taken off decides exchange; the second 53-bit number is the alias sample
// (position and bias from one number, as AliasSampler_t does it).  Every choice keeps the reference's probabilities
// (resolution 2^-53 / 2^-45 / 2^-53), pgen is unchanged.  The CPU checker of the test suite draws the same way.
template <int NW>
__device__ __forceinline__ void pchb_pick_holes(const Params &P, const Det<NW> &d, int s1, int s2, double pGen, double u2, Stream &rng, Excit<NW> &E);
template <int NW>
__device__ void gen_pchb_double(const Params &P, const Det<NW> &d, double r, Stream &rng, Excit<NW> &E) {
    E.ic = 2; E.valid = false; E.err = 0;
    const int nA = P.nocc_alpha, nB = P.nocc_beta;
    const int AA = nA * (nA - 1) / 2, BB = nB * (nB - 1) / 2, par = AA + BB, AB = nA * nB;
    u64 m1, m2; int k1, k2;                        // the pair = k1-th orbital of mask m1 and k2-th of mask m2
    const bool is_par = r < P.p_parallel;
    const double x = is_par ? r * P.c_par : (r - P.p_parallel) * P.c_opp;
    int idx = min((int)x, (is_par ? par : AB) - 1);
    const double u2 = x - (double)idx;             // uniform on [0,1), independent of idx
    double pGen = is_par ? P.pgen_pair_par : P.pgen_pair_opp;      // p_parallel / par, (1 - p_parallel) / AB: host quotients
    if (is_par) {
        u64 mask = NG_ALPHA_MASK;
        if (idx >= AA) { idx -= AA; mask = NG_BETA_MASK; }
        // n1 = ceil((1 + sqrt(9 + 8 idx)) / 2), n2 = idx + 1 - (n1 - 1)(n1 - 2) / 2, tabulated by the host
        const u32 t = __ldg(&P.tri_tab[idx]);
        m1 = mask; k1 = (int)(t >> 8); m2 = mask; k2 = (int)(t & 0xffu);
    } else {
        const int q = (nA == 1) ? idx : (int)__umulhi((u32)idx, P.magic_nalpha);   // idx / nA
        m1 = NG_ALPHA_MASK; k1 = 1 + idx - q * nA;
        m2 = NG_BETA_MASK;  k2 = 1 + q;             // == 1 + floor(idx / real(nA))
    }
    // the two orbital selections are common to both branches (kept out of the divergent part)
    const int oa = select_orb(d, m1, k1), ob = select_orb(d, m2, k2);
    pchb_pick_holes(P, d, min(oa, ob), max(oa, ob), pGen, u2, rng, E);
}
// the part of GAS_doubles_PCHB_gen_exc after the particles are chosen (:204-262): sampler by spin and exchange, alias
// sample of the hole pair, validity.  u2: the uniform number of the exchange decision.
template <int NW>
__device__ __forceinline__ void pchb_pick_holes(const Params &P, const Det<NW> &d, int s1, int s2, double pGen, double u2, Stream &rng, Excit<NW> &E) {
    const int ij = (int)tri((u32)gtid(s1), (u32)gtid(s2));      // fuse_index
    int spin1 = s1 & 1, spin2 = s2 & 1;           // getSpinIndex: 0 alpha, 1 beta
    const int4 pi = __ldg(reinterpret_cast<const int4 *>(P.pchb_pair + (ij - 1)));    // {p_exch, nonempty, pad}
    int sampler;
    if (spin1 == spin2) sampler = 0;
    else {
        const double pe = __hiloint2double(pi.y, pi.x);
        if (u2 < pe) { sampler = 2; pGen *= pe; const int t = spin1; spin1 = spin2; spin2 = t; }
        else { sampler = 1; pGen *= (1.0 - pe); }
    }
    E.src1 = s1; E.src2 = s2; E.tgt1 = 0; E.tgt2 = 0; E.pgen = pGen;
    // AliasSampler_t::sample
    if (((pi.z >> sampler) & 1) == 0) return;                        // empty sampler: ab = 0
    const PchbEntry *tab = P.pchb + ((size_t)(ij - 1) * 3 + sampler) * P.ab_max;
    const double rr = rng.draw53();
    const int pos = (int)(P.ab_max * rr) + 1;
    const double bias = fmax(P.ab_max * rr + 1 - pos, 0.0);
    // the whole entry with one 256-bit load: {bias, prob, prob_alias, tgt | tgt_alias << 32}
    double en_bias, en_prob, en_palias; u64 en_tgt;
    asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=d"(en_bias), "=d"(en_prob), "=d"(en_palias), "=l"(en_tgt) : "l"(tab + (pos - 1)));
    const bool own = bias < en_bias;
    const double pGenHoles = own ? en_prob : en_palias;
    const u32 tg = own ? (u32)en_tgt : (u32)(en_tgt >> 32);
    const int o1 = 2 * (int)(tg & 0xffffu) - spin1, o2 = 2 * (int)(tg >> 16) - spin2;
    E.tgt1 = o1; E.tgt2 = o2;
    bool invalid = (o1 == 0 || o2 == 0) || occ(d, o1) || occ(d, o2);
    if (!invalid && fabs(pGenHoles) <= NG_EPS) invalid = true;
    if (invalid) return;
    const int t1 = min(o1, o2), t2 = max(o1, o2);
    E.tgt1 = t1; E.tgt2 = t2;
    E.pgen = pGen * pGenHoles;
    E.valid = true;
}

// CDF_Sampler_t over the occupied orbitals with the weights row[orbital - 1] (src/CDF_sampling.fpp:57-118, as
// constrained_sample uses it, src/aliasSampling.F90:519-525): the first occupied orbital, in ascending order, whose
// running sum of weights reaches r * total; the last one with a non-zero weight if rounding leaves the sum below.
template <int NW>
__device__ __forceinline__ int cdf_pick_occupied(const Det<NW> &d, const double *row, double total, double r, double &renorm_out) {
    const double thr = r * total;
    double cum = 0.0;
    int chosen = 0, last = 0;
    Det<NW> a = d;
    while (det_any(a)) {
        const int o = pop_lowest(a);
        const double w = __ldg(&row[o - 1]);
        cum += w;
        if (w > 0.0) { last = o; if (chosen == 0 && cum >= thr) chosen = o; }
    }
    renorm_out = cum;
    return chosen ? chosen : last;
}
template <int NW>
__device__ __forceinline__ double sum_occupied(const Det<NW> &d, const double *row) {
    double s = 0.0;
    Det<NW> a = d;
    while (det_any(a)) s += __ldg(&row[pop_lowest(a) - 1]);
    return s;
}
// GAS_doubles_PCHB_gen_exc with PC_FullyWeightedParticles_t (src/gasci_pchb_doubles_select_particles.fpp:330-384):
// first particle with p_first restricted to the occupied orbitals, second with p(J | I) likewise, p({I, J}) summed over
// both orders.  Random numbers: r (from the attempt's first number) picks the first particle, the next 53-bit number
// the second; the exchange decision takes a 32-bit number and the alias sample a 53-bit number of the second block.
template <int NW>
__device__ void gen_pchb_double_full(const Params &P, const Det<NW> &d, double r, Stream &rng, Excit<NW> &E) {
    E.ic = 2; E.valid = false; E.err = 0; E.src1 = E.src2 = E.tgt1 = E.tgt2 = 0; E.pgen = 1.0;
    const int nb = P.nbasis;
    const bool unif_first = P.pchb_particles == 2;      // UNIF-FULL (draw_PC_WeightedParticles_t, :440-478): the first particle uniformly
    const double renorm_first = sum_occupied(d, P.pchb_pfirst);
    const double r2 = rng.draw53();
    if (!unif_first && fabs(renorm_first) <= NG_EPS) return;
    double dummy;
    const int s1 = unif_first ? select_orb(d, ~0ull, min((int)(r * P.nel), P.nel - 1) + 1)
                              : cdf_pick_occupied(d, P.pchb_pfirst, renorm_first, r, dummy);
    const double *row1 = P.pchb_psecond + (size_t)(s1 - 1) * nb;
    const double renorm_second1 = sum_occupied(d, row1);
    if (fabs(renorm_second1) <= NG_EPS) return;
    const int s2 = cdf_pick_occupied(d, row1, renorm_second1, r2, dummy);
    const double p_second1 = __ldg(&row1[s2 - 1]) / renorm_second1;
    const double *row2 = P.pchb_psecond + (size_t)(s2 - 1) * nb;
    const double renorm_second2 = sum_occupied(d, row2);
    const double p_second2 = (fabs(renorm_second2) <= NG_EPS) ? 0.0 : __ldg(&row2[s1 - 1]) / renorm_second2;
    double pGen;
    if (unif_first) pGen = (p_second1 + p_second2) / (double)P.nel;
    else {
        const double p_first1 = __ldg(&P.pchb_pfirst[s1 - 1]) / renorm_first, p_first2 = __ldg(&P.pchb_pfirst[s2 - 1]) / renorm_first;
        pGen = p_first1 * p_second1 + p_first2 * p_second2;
    }
    const double u2 = rng.draw32();
    pchb_pick_holes(P, d, min(s1, s2), max(s1, s2), pGen, u2, rng, E);
}

// make_single / make_double results derived from the bit-strings: parity and the excited determinant
template <int NW>
__device__ __forceinline__ void finalize_excit(const Det<NW> &d, Excit<NW> &E) {
    E.detJ = d;
    if (E.ic == 1) {
        E.parity = parity_single(d, E.src1, E.tgt1);
        clr_orb(E.detJ, E.src1); set_orb(E.detJ, E.tgt1);
    } else {
        E.parity = parity_double(d, E.src1, E.src2, E.tgt1, E.tgt2);
        clr_orb(E.detJ, E.src1); clr_orb(E.detJ, E.src2); set_orb(E.detJ, E.tgt1); set_orb(E.detJ, E.tgt2);
    }
}

// generation proper (orbitals + pgen); parity and detJ are added by finalize_excit
template <int NW, int SYS>
__device__ __forceinline__ void generate_excitation_core(const Params &P, const Det<NW> &d, Stream &rng, Excit<NW> &E) {
    if (SYS == NECI_SYS_HUBBARD_RS) gen_rs_hubbard(P, d, rng, E);
    else if (SYS == NECI_SYS_HUBBARD_K) gen_k_hubbard(P, d, rng, E);
    else {
        // gen_exc_sd: the first number of the attempt's block; singles continue in the second block (word 4)
        const double u = rng.draw53();
        if (u < P.p_singles) { rng.pos = 4; gen_uniform_single(P, d, rng, E); E.pgen = E.pgen * P.p_singles; }
        else {
            if (SYS == NG_SYS_PCHB_FULL) gen_pchb_double_full(P, d, (u - P.p_singles) * P.inv_1m_ps, rng, E);
            else gen_pchb_double(P, d, (u - P.p_singles) * P.inv_1m_ps, rng, E);
            E.pgen = E.pgen * P.p_doubles;
        }
    }
}
template <int NW, int SYS>
__device__ __forceinline__ void generate_excitation(const Params &P, const Det<NW> &d, Stream &rng, Excit<NW> &E) {
    generate_excitation_core<NW, SYS>(P, d, rng, E);
    if (E.valid) finalize_excit(d, E);
}

// ---- HPHF functions (src/HPHFIntegrals.fpp, src/HPHFRandExcit.F90, src/DetBitOps.F90:648-740,819-848), even S ------
template <int NW> __device__ __forceinline__ Det<NW> spin_sym(const Det<NW> &a) {           // spin_sym_ilut
    Det<NW> b;
    b.w[0] = ((a.w[0] & NG_ALPHA_MASK) >> 1) | ((a.w[0] & NG_BETA_MASK) << 1);
    if (NW > 1) b.w[NW - 1] = ((a.w[NW - 1] & NG_ALPHA_MASK) >> 1) | ((a.w[NW - 1] & NG_BETA_MASK) << 1);
    return b;
}
template <int NW> __device__ __forceinline__ bool closed_shell(const Det<NW> &a) {           // TestClosedShellDet
    u64 x = ((a.w[0] & NG_ALPHA_MASK) >> 1) ^ (a.w[0] & NG_BETA_MASK);
    if (NW > 1) x |= ((a.w[NW - 1] & NG_ALPHA_MASK) >> 1) ^ (a.w[NW - 1] & NG_BETA_MASK);
    return x == 0ull;
}
template <int NW> __device__ __forceinline__ int open_orbs(const Det<NW> &a) {                // CalcOpenOrbs
    int n = __popcll(~((a.w[0] & NG_ALPHA_MASK) >> 1) & (a.w[0] & NG_BETA_MASK));
    if (NW > 1) n += __popcll(~((a.w[NW - 1] & NG_ALPHA_MASK) >> 1) & (a.w[NW - 1] & NG_BETA_MASK));
    return n;
}
// DetBitLT(a, b) == 1: signed comparison, word 0 first
template <int NW> __device__ __forceinline__ bool det_less(const Det<NW> &a, const Det<NW> &b) {
    if ((long long)a.w[0] != (long long)b.w[0] || NW == 1) return (long long)a.w[0] < (long long)b.w[0];
    return (long long)a.w[NW - 1] < (long long)b.w[NW - 1];
}
// FindBitExcitLevel(ref, det, t_hphf_ic = .true.): the smallest level over the spin-flipped partners
template <int NW, bool HPHF> __device__ __forceinline__ int excit_level_ref(const Det<NW> &ref, const Det<NW> &d) {
    int ic = excit_level(ref, d);
    if (HPHF && !(closed_shell(ref) && closed_shell(d))) {
        const Det<NW> r2 = spin_sym(ref), d2 = spin_sym(d);
        ic = min(min(ic, excit_level(ref, d2)), min(excit_level(r2, d), excit_level(r2, d2)));
    }
    return ic;
}
// hphf_off_diag_helement_norm (src/HPHFIntegrals.fpp:62-150)
template <int NW, int SYS>
__device__ double hphf_off_diag(const Params &P, const Det<NW> &I, const Det<NW> &J) {
    if (det_eq(I, J)) return 0.0;
    double hel = helement<NW, SYS>(P, I, J);
    if (closed_shell(I)) { if (!closed_shell(J)) hel = hel * sqrt(2.0); }
    else if (closed_shell(J)) hel = hel * sqrt(2.0);
    else {
        const Det<NW> I2 = spin_sym(I);
        if (excit_level(I2, J) <= 2) {
            const double m2 = helement<NW, SYS>(P, I2, J);
            hel = (open_orbs(I) % 2 == 0) ? hel + m2 : hel - m2;
        }
    }
    return hel;
}
// hphf_diag_helement (src/HPHFIntegrals.fpp:348-411); ECore included
template <int NW, int SYS>
__device__ double hphf_diag(const Params &P, const Det<NW> &I) {
    double hel = sltcnd_0(P, I) + P.ecore;        // the determinant's own diagonal element (not diagonal_matel: that dispatches here)
    if (!closed_shell(I)) {
        const Det<NW> I2 = spin_sym(I);
        if (excit_level(I, I2) <= 2) {
            const double m2 = helement<NW, SYS>(P, I, I2);
            hel = (open_orbs(I) % 2 == 1) ? hel - m2 : hel + m2;
        }
    }
    return hel;
}
// CalcNonUniPGen for the PCHB class generator (src/HPHFRandExcit.F90:686-821 -> get_pgen_sd,
// src/excitation_generators.F90:141-158): probability with which I -> K is drawn, K given by its excitation
template <int NW>
__device__ double calc_pgen_pchb(const Params &P, const Det<NW> &d, int ic, int s1, int s2, int t1, int t2) {
    if (ic == 1) {
        int ElecsWNoExcits = 0, NExcitA = 0;
        const int cls = __ldg(&P.class_of_spinorb[s1 - 1]);
#pragma unroll 1
        for (int c = 0; c < P.n_classes; ++c) {
            int o = __popcll(d.w[0] & P.class_mask[c][0]), t = __popcll(P.class_mask[c][0]);
            if (NW > 1) { o += __popcll(d.w[NW - 1] & P.class_mask[c][1]); t += __popcll(P.class_mask[c][1]); }
            if (t - o == 0) ElecsWNoExcits += o;
            if (c == cls) NExcitA = t - o;
        }
        double pgen = (1 - P.p_doubles) / ((double)(NExcitA * (P.nel - ElecsWNoExcits)));
        pgen = pgen / P.p_singles;
        return P.p_singles * pgen;
    }
    if (ic != 2) return 0.0;
    // GAS_doubles_PCHB_get_pgen (src/gasci_pchb_doubles_spatorb_fastweighted.fpp:286-326)
    const int ij = (int)tri((u32)gtid(s1), (u32)gtid(s2)), ab = (int)tri((u32)gtid(t1), (u32)gtid(t2));
    const bool same = ((s1 ^ s2) & 1) == 0;
    double pgen = same ? P.pgen_pair_par : P.pgen_pair_opp;
    const int4 pi = __ldg(reinterpret_cast<const int4 *>(P.pchb_pair + (ij - 1)));
    const double pe = __hiloint2double(pi.y, pi.x);
    int sampler = 0;
    if (!same) {
        if (((s1 ^ t1) & 1) == 0 || gtid(t1) == gtid(t2)) { sampler = 1; pgen *= (1.0 - pe); }
        else { sampler = 2; pgen *= pe; }
    }
    if (((pi.z >> sampler) & 1) == 0) return 0.0;
    const PchbEntry *tab = P.pchb + ((size_t)(ij - 1) * 3 + sampler) * P.ab_max;
    return (1.0 - P.p_singles) * (pgen * __ldg(&tab[ab - 1].prob));
}
// gen_hphf_excit (src/HPHFRandExcit.F90:175-476) applied to an excitation the PCHB generator has produced:
// replaces detJ by the allowed representative of its HPHF function, adds the probability of having drawn the
// partner determinant, and returns the HPHF matrix element.  false: the excitation stays inside I's own function.
template <int NW, int SYS>
__device__ bool hphf_fixup(const Params &P, const Det<NW> &d, Excit<NW> &E, double &hel) {
    if (!closed_shell(E.detJ)) {
        const Det<NW> J2 = spin_sym(E.detJ);                      // ReturnAlphaOpenDet
        const int exl = excit_level(d, J2);                       // to the determinant that was NOT generated
        if (exl == 0) return false;
        if (exl <= 2) {
            Det<NW> S, T;
            S.w[0] = d.w[0] & ~J2.w[0]; T.w[0] = J2.w[0] & ~d.w[0];
            if (NW > 1) { S.w[NW - 1] = d.w[NW - 1] & ~J2.w[NW - 1]; T.w[NW - 1] = J2.w[NW - 1] & ~d.w[NW - 1]; }
            const int s1 = pop_lowest(S), t1 = pop_lowest(T);
            const int s2 = (exl == 2) ? pop_lowest(S) : 0, t2 = (exl == 2) ? pop_lowest(T) : 0;
            E.pgen = E.pgen + calc_pgen_pchb<NW>(P, d, exl, s1, s2, t1, t2);
        }
        if (det_less(E.detJ, J2)) E.detJ = J2;
    }
    hel = hphf_off_diag<NW, SYS>(P, d, E.detJ);
    return true;
}

template <int NW> __device__ double hphf_diag_dispatch(const Params &P, const Det<NW> &d) { return hphf_diag<NW, NG_SYS_PCHB_HPHF>(P, d); }
template <int NW> __device__ double hphf_off_diag_dispatch(const Params &P, const Det<NW> &I, const Det<NW> &J) {
    return hphf_off_diag<NW, NG_SYS_PCHB_HPHF>(P, I, J);
}

// get_spawn_helement = get_helement_det_only (src/Determinants.F90:508-554)
template <int NW, int SYS>
__device__ __forceinline__ double spawn_helement(const Params &P, const Det<NW> &d, const Excit<NW> &E) {
    if (SYS == NECI_SYS_HUBBARD_RS) { const double h = tmat_el(P, E.src1, E.tgt1); return E.parity ? -h : h; }
    if (SYS == NECI_SYS_HUBBARD_K) return offdiag_k(P, E.src1, E.src2, E.tgt1, E.tgt2, E.parity);
    double h;
    if (E.ic == 1) h = sltcnd_1(P, d, E.src1, E.tgt1);
    else h = sltcnd_2(P, E.src1, E.src2, E.tgt1, E.tgt2);
    return E.parity ? -h : h;
}

}  // namespace ng
