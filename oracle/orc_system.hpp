// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see orc_common.hpp).
// Hamiltonian matrix elements and excitation generators, restated from the
// reference in its own list-based (nI) style.
#pragma once
#include "orc_common.hpp"

namespace orc {

struct System {
    int type = 0;
    int nel = 0, nbasis = 0, nwords = 1;
    int nocc_alpha = 0, nocc_beta = 0;
    bool t_exch = true, t_no_brillouin = false, t_hphf = false;
    double ecore = 0.0;
    // FCIDUMP
    std::vector<double> umat, tmat;
    // PCHB
    int n_spat = 0, ij_max = 0, ab_max = 0;
    std::vector<double> probs, bias, p_exch;
    std::vector<int32_t> alias, tgt_orbs;
    double p_singles = 1.0, p_doubles = 1.0, p_parallel = 1.0;   // lattice models: a single excitation class
    // PCHB particle selection: 0 UNIF-UNIF (pick_biased_elecs), 1 FULL-FULL (PC_FullyWeightedParticles_t) with the
    // normalised probabilities of its I_sampler (p_first[I]) and J_sampler (p_second[I][J] = p(J | I))
    int pchb_particles = 0;
    std::vector<double> p_first, p_second;
    int n_classes = 0;
    std::vector<int32_t> class_of_spinorb;          // 0-based class id per spin orbital (index orb-1)
    std::vector<std::vector<int32_t>> class_orbs;   // members of each class, ascending
    // r-space Hubbard
    int max_neigh = 0;
    std::vector<int32_t> neighbours;
    double uhub = 0;
    // k-space Hubbard
    int n_k = 0;
    std::vector<int32_t> ksum, kdiff;
    std::vector<double> eps_k;
    double u_over_n = 0;

    // ---- integrals ---------------------------------------------------------
    // UMatInd_base, src/UMatCache.F90:257-296 (real, hermitian 2-body)
    static int64_t UMatInd(int i, int j, int k, int l) {
        int64_t A = (i > k) ? (int64_t)i * (i - 1) / 2 + k : (int64_t)k * (k - 1) / 2 + i;
        int64_t B = (j > l) ? (int64_t)j * (j - 1) / 2 + l : (int64_t)l * (l - 1) / 2 + j;
        return (A > B) ? A * (A - 1) / 2 + B : B * (B - 1) / 2 + A;
    }
    // get_umat_el: <ij|kl> over spatial orbitals.
    //   FCIDUMP: get_umat_el_normal, src/Integrals_neci.F90:1875-1898
    //   k-space Hubbard: get_umat_kspace, src/k_space_hubbard.F90:356-374
    double get_umat_el(int i, int j, int k, int l) const {
        if (type == NECI_SYS_HUBBARD_K) {
            // momentum conservation k_i + k_j == k_k + k_l
            const int a = ksum[(i - 1) * n_k + (j - 1)], b = ksum[(k - 1) * n_k + (l - 1)];
            return (a == b) ? u_over_n : 0.0;
        }
        return umat[UMatInd(i, j, k, l) - 1];
    }
    // GetTMatEl, src/OneEInts.F90:188-226 (TMat2D(i,j))
    double GetTMatEl(int i, int j) const {
        if (type == NECI_SYS_HUBBARD_K) return (i == j) ? eps_k[gtID(i) - 1] : 0.0;
        return tmat[(size_t)(i - 1) + (size_t)nbasis * (j - 1)];
    }

    // ---- Slater-Condon rules, src/sltcnd.fpp:585-708 -----------------------
    double sltcnd_0(const int *nI) const {
        double hel_sing = 0.0;
        for (int i = 0; i < nel; ++i) hel_sing += GetTMatEl(nI[i], nI[i]);
        double hel_doub = 0.0;
        for (int i = 0; i < nel - 1; ++i) {
            double s = 0.0;  // sum(get_2el(id(i), id(i+1:), id(i), id(i+1:)))
            for (int j = i + 1; j < nel; ++j) s += get_umat_el(gtID(nI[i]), gtID(nI[j]), gtID(nI[i]), gtID(nI[j]));
            hel_doub += s;
        }
        double hel_tmp = 0.0;
        if (t_exch) {
            for (int i = 0; i < nel - 1; ++i)
                for (int j = i + 1; j < nel; ++j)
                    if (G1_Ms(nI[i]) == G1_Ms(nI[j]))
                        hel_tmp -= get_umat_el(gtID(nI[i]), gtID(nI[j]), gtID(nI[j]), gtID(nI[i]));
        }
        return hel_doub + hel_tmp + hel_sing;
    }
    double sltcnd_1_kernel(const int *nI, int src, int tgt) const {
        const int id1 = gtID(src), id2 = gtID(tgt);
        double hel = 0.0;
        if (G1_Ms(src) == G1_Ms(tgt)) {
            for (int i = 0; i < nel; ++i)
                if (src != nI[i]) { const int id = gtID(nI[i]); hel += get_umat_el(id1, id, id2, id); }
        }
        if (t_exch && G1_Ms(src) == G1_Ms(tgt)) {
            for (int i = 0; i < nel; ++i)
                if (src != nI[i] && G1_Ms(src) == G1_Ms(nI[i])) {
                    const int id = gtID(nI[i]);
                    hel -= get_umat_el(id1, id, id, id2);
                }
        }
        return hel + GetTMatEl(src, tgt);
    }
    // ex = {src1, src2, tgt1, tgt2}
    double sltcnd_2_kernel(const int *ex) const {
        const int s1 = ex[0], s2 = ex[1], t1 = ex[2], t2 = ex[3];
        double hel = 0.0;
        if (G1_Ms(s1) == G1_Ms(t1) && G1_Ms(s2) == G1_Ms(t2))
            hel = get_umat_el(gtID(s1), gtID(s2), gtID(t1), gtID(t2));
        if (G1_Ms(s1) == G1_Ms(t2) && G1_Ms(s2) == G1_Ms(t1))
            hel -= get_umat_el(gtID(s1), gtID(s2), gtID(t2), gtID(t1));
        return hel;
    }

    // ---- lattice-model elements --------------------------------------------
    // get_diag_helemen_rs_hub: U * (# doubly occupied sites), src/real_space_hubbard.F90:2303-2321
    double diag_rs_hub(const uint64_t *ilut) const {
        int nd = 0;
        for (int s = 1; s <= nbasis / 2; ++s) nd += (is_occ(ilut, 2 * s - 1) && is_occ(ilut, 2 * s));
        return uhub * nd;
    }
    // get_offdiag_helement_rs_hub, src/real_space_hubbard.F90:2425-2457
    double offdiag_rs_hub(int src, int tgt, bool tpar) const {
        double hel = tmat[(size_t)(src - 1) + (size_t)nbasis * (tgt - 1)];
        return tpar ? -hel : hel;
    }
    // get_offdiag_helement_k_sp_hub, src/k_space_hubbard.F90:2620-2671,2790
    double offdiag_k_hub(const int *ex, bool tpar) const {
        const int s1 = ex[0], s2 = ex[1], t1 = ex[2], t2 = ex[3];
        if (is_beta(s1) == is_beta(s2) || is_beta(t1) == is_beta(t2)) return 0.0;
        double hel = 0.0;
        if (is_beta(s1) == is_beta(t1) && is_beta(s2) == is_beta(t2))
            hel = get_umat_el(gtID(s1), gtID(s2), gtID(t1), gtID(t2));
        else if (is_beta(s1) == is_beta(t2) && is_beta(s2) == is_beta(t1))
            hel = -get_umat_el(gtID(s1), gtID(s2), gtID(t1), gtID(t2));
        if (std::fabs(hel) < EPS) return hel;
        return tpar ? -hel : hel;
    }
    // get_orb_from_kpoints, src/lattice_models_utils.F90:1881-1925 (opposite-spin pair)
    int get_orb_from_kpoints(int orbi, int orbj, int orba) const {
        const int ki = gtID(orbi) - 1, kj = gtID(orbj) - 1, ka = gtID(orba) - 1;
        const int kb = kdiff[ksum[ki * n_k + kj] * n_k + ka];
        // spin of b: opposite of a for an alpha/beta pair
        const bool same = (is_beta(orbi) == is_beta(orbj));
        bool b_beta;
        if (same) b_beta = is_beta(orbi); else b_beta = !is_beta(orba);
        return 2 * (kb + 1) - (b_beta ? 1 : 0);
    }
};

// ---- determinant helpers -----------------------------------------------------
// decode_bit_det, src/BitReps.F90:931-957 (result: ascending orbital list)
inline int decode(const uint64_t *ilut, int nbasis, int *nI) {
    int n = 0;
    for (int o = 1; o <= nbasis; ++o) if (is_occ(ilut, o)) nI[n++] = o;
    return n;
}
// FindBitExcitLevel, src/DetBitOps.F90:140-180: popcount(I & (I xor J))
inline int excit_level(const uint64_t *a, const uint64_t *b, int nwords) {
    int n = 0;
    for (int w = 0; w < nwords; ++w) n += __builtin_popcountll(a[w] & (a[w] ^ b[w]));
    return n;
}

// make_single, src/excit_parity.F90:15-76 (nI sorted ascending, elec 1-based)
inline void make_single(const int *nI, int nel, int elec, int tgt, int *nJ, int *ex, bool &tParity) {
    for (int i = 0; i < nel; ++i) nJ[i] = nI[i];
    const int src = nI[elec - 1];
    ex[0] = src; ex[1] = tgt;
    int i;
    if (src < tgt) {
        for (i = elec + 1; i <= nel; ++i) {
            if (tgt < nJ[i - 1]) { nJ[i - 2] = tgt; break; }
            nJ[i - 2] = nJ[i - 1];
        }
        if (i == nel + 1) nJ[nel - 1] = tgt;
    } else {
        for (i = elec - 1; i >= 1; --i) {
            if (tgt > nJ[i - 1]) { nJ[i] = tgt; break; }
            nJ[i] = nJ[i - 1];
        }
        if (i == 0) nJ[0] = tgt;
    }
    tParity = ((elec - i) % 2 == 0);
}

// make_double, src/excit_parity.F90:78-170.  ex = {src1,src2,tgt1,tgt2}
inline void make_double(const int *nI, int nel, int elec1, int elec2, int tgt1, int tgt2,
                        int *nJ, int *ex, bool &tParity) {
    int elecs[2] = {std::min(elec1, elec2), std::max(elec1, elec2)};
    int tgts[2] = {std::min(tgt1, tgt2), std::max(tgt1, tgt2)};
    int srcs[2] = {nI[elecs[0] - 1], nI[elecs[1] - 1]};
    ex[0] = srcs[0]; ex[1] = srcs[1]; ex[2] = tgts[0]; ex[3] = tgts[1];
    for (int i = 0; i < nel; ++i) nJ[i] = nI[i];
    if (srcs[0] < tgts[0] && srcs[1] < tgts[0]) elecs[1] -= 1;
    int pos_moved = 0;
    for (int k = 0; k < 2; ++k) {
        int i;
        if (srcs[k] < tgts[k]) {
            if (elecs[k] == nel) { i = nel + 1; nJ[nel - 1] = tgts[k]; }
            else {
                for (i = elecs[k] + 1; i <= nel; ++i) {
                    if (tgts[k] < nJ[i - 1]) { nJ[i - 2] = tgts[k]; break; }
                    nJ[i - 2] = nJ[i - 1];
                }
                if (i == nel + 1) nJ[nel - 1] = tgts[k];
            }
        } else {
            if (elecs[k] == 1) { i = 0; nJ[0] = tgts[k]; }
            else {
                for (i = elecs[k] - 1; i >= 1; --i) {
                    if (tgts[k] > nJ[i - 1]) { nJ[i] = tgts[k]; break; }
                    nJ[i] = nJ[i - 1];
                }
                if (i == 0) nJ[0] = tgts[k];
            }
        }
        pos_moved += elecs[k] - i + 1;
    }
    tParity = (pos_moved & 1) != 0;
}

// GetBitExcitation-style excitation matrix + parity between two determinants
// differing by ic <= 2 (used by get_helement, src/Determinants.F90:340-420):
// ex = {src..., tgt...}, parity from moving src->tgt through nI (make_single/double).
inline int excitation_between(const uint64_t *iI, const uint64_t *iJ, int nbasis, int nel,
                              int *ex, bool &tParity) {
    int nI[128], nJ[128], tmp[128];
    decode(iI, nbasis, nI);
    int tgts[4], ns = 0, nt = 0, elec[4];
    for (int i = 0; i < nel; ++i) if (!is_occ(iJ, nI[i])) { if (ns < 4) elec[ns] = i + 1; ++ns; }
    for (int o = 1; o <= nbasis; ++o) if (is_occ(iJ, o) && !is_occ(iI, o)) { if (nt < 4) tgts[nt] = o; ++nt; }
    if (ns != nt || ns > 2) return (ns == nt) ? ns : -1;
    (void)nJ;
    if (ns == 1) { int e[2]; make_single(nI, nel, elec[0], tgts[0], tmp, e, tParity); ex[0] = e[0]; ex[1] = 0; ex[2] = e[1]; ex[3] = 0; }
    else if (ns == 2) make_double(nI, nel, elec[0], elec[1], tgts[0], tgts[1], tmp, ex, tParity);
    else tParity = false;
    return ns;
}

// get_helement(nI,nJ): src/Determinants.F90:340-506 -> sltcnd / lattice routines
inline double get_helement(const System &S, const uint64_t *iI, const uint64_t *iJ) {
    int ex[4]; bool par = false;
    const int ic = excitation_between(iI, iJ, S.nbasis, S.nel, ex, par);
    int nI[128];
    decode(iI, S.nbasis, nI);
    if (ic == 0) {
        if (S.type == NECI_SYS_HUBBARD_RS) return S.diag_rs_hub(iI);
        return S.sltcnd_0(nI) + S.ecore;
    }
    if (ic == 1) {
        if (S.type == NECI_SYS_HUBBARD_RS) return S.offdiag_rs_hub(ex[0], ex[2], par);
        if (S.type == NECI_SYS_HUBBARD_K) return 0.0;   // momentum forbids singles
        const double h = S.sltcnd_1_kernel(nI, ex[0], ex[2]);
        return par ? -h : h;
    }
    if (ic == 2) {
        if (S.type == NECI_SYS_HUBBARD_RS) return 0.0;
        if (S.type == NECI_SYS_HUBBARD_K) return S.offdiag_k_hub(ex, par);
        const double h = S.sltcnd_2_kernel(ex);
        return par ? -h : h;
    }
    return 0.0;
}

// ---- excitation generators ---------------------------------------------------
struct Excitation {
    bool valid = false;
    int ic = 0;
    int ex[4] = {0, 0, 0, 0};     // src1,src2,tgt1,tgt2
    bool parity = false;
    double pgen = 0.0;
    int nJ[128];
    uint64_t ilutJ[2] = {0, 0};
    int err = 0;
};

// pick_from_cum_list, src/lattice_models_utils.F90:150-172
// + binary_search_first_ge, src/lib/util_mod_numerical.F90.template:35-85
inline int pick_from_cum_list(const double *cum_arr, int n, double cum_sum, Stream &rng, double &pgen) {
    if (cum_sum < EPS) { pgen = 0.0; return -1; }
    const double r = rng.draw53() * cum_sum;
    int lo = 1, hi = n, pos = -1;
    if (cum_arr[hi - 1] < r) { pgen = 0.0; return -1; }
    while (hi != lo) {
        pos = (int)((float)(hi + lo) / 2.0f);
        if (!(cum_arr[pos - 1] < r)) hi = pos; else lo = pos + 1;
    }
    const int ind = hi;
    pgen = (ind == 1) ? cum_arr[0] / cum_sum : (cum_arr[ind - 1] - cum_arr[ind - 2]) / cum_sum;
    return ind;
}

// gen_excit_rs_hubbard, src/real_space_hubbard.F90:1938-2033
inline void gen_excit_rs_hubbard(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E) {
    E.ic = 1;
    const int elec = 1 + (int)(rng.draw32() * S.nel);
    const double p_elec = 1.0 / (double)S.nel;
    const int src = nI[elec - 1];
    const int32_t *neigh = &S.neighbours[(size_t)(src - 1) * S.max_neigh];
    int nn = 0; while (nn < S.max_neigh && neigh[nn] != 0) ++nn;
    double cum_arr[16], cum_sum = 0.0;
    // create_cum_list_rs_hubbard, :2080-2142 : weight = |t| for empty neighbours
    for (int i = 0; i < nn; ++i) {
        double elem = 0.0;
        if (!is_occ(ilutI, neigh[i])) elem = std::fabs(S.offdiag_rs_hub(src, neigh[i], false));
        cum_sum += elem; cum_arr[i] = cum_sum;
    }
    if (cum_sum < EPS) { E.valid = false; E.pgen = 0.0; return; }
    double p_orb;
    const int ind = pick_from_cum_list(cum_arr, nn, cum_sum, rng, p_orb);
    if (ind < 0) { E.valid = false; E.pgen = 0.0; return; }
    const int orb = neigh[ind - 1];
    E.pgen = p_elec * p_orb;
    int ex2[2];
    make_single(nI, S.nel, elec, orb, E.nJ, ex2, E.parity);
    E.ex[0] = ex2[0]; E.ex[1] = 0; E.ex[2] = ex2[1]; E.ex[3] = 0;
    E.ilutJ[0] = ilutI[0]; E.ilutJ[1] = (S.nwords > 1) ? ilutI[1] : 0;
    clr_orb(E.ilutJ, src); set_orb(E.ilutJ, orb);
    E.valid = true;
}

// gen_excit_k_space_hub, src/k_space_hubbard.F90:535-625
inline void gen_excit_k_space_hub(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E) {
    E.ic = 2;
    // pick_spin_opp_elecs, src/lattice_models_utils.F90:123-148.  The reference draws ordered electron pairs until the
    // two spins differ ("i think i could do that way more efficiently, but do it in the simple loop way for now"):
    // every one of the nOccAlpha * nOccBeta opposite-spin pairs comes out with probability p_elec = 1 / (nOccAlpha *
    // nOccBeta).  The engine and this checker take the pair from ONE number instead -- index = int(r * nOccAlpha *
    // nOccBeta), alpha electron index mod nOccAlpha, beta electron index / nOccAlpha, both counted in orbital
    // order -- the same distribution and the same p_elec without a data-dependent loop (on the GPU the loop ran until
    // the slowest of 32 lanes had its pair: a third of the lanes active, 60 % of the generator's instructions).
    int elecs[2];
    {
        const int npair = S.nocc_alpha * S.nocc_beta;
        int idx = (int)(rng.draw32() * (double)npair);
        if (idx > npair - 1) idx = npair - 1;
        const int ia = idx % S.nocc_alpha, ib = idx / S.nocc_alpha;       // 0-based among the alpha / beta electrons
        int ca = 0, cb = 0, ea = 0, eb = 0;
        for (int e = 0; e < S.nel; ++e) {
            if (is_beta(nI[e])) { if (cb == ib) eb = e + 1; ++cb; }
            else { if (ca == ia) ea = e + 1; ++ca; }
        }
        elecs[0] = std::min(ea, eb); elecs[1] = std::max(ea, eb);
    }
    const double p_elec = 1.0 / (double)(S.nocc_beta * S.nocc_alpha);
    const int src[2] = {nI[elecs[0] - 1], nI[elecs[1] - 1]};
    // create_ab_list_hubbard, :1795-1817 ; excit_cache(i,j,a) = |<ij|H|ab>|, :485-520
    double cum_arr[128], cum_sum = 0.0;
    int orb_b[128];
    for (int a = 1; a <= S.nbasis; ++a) {
        double elem = 0.0; int b = -1;
        if (!is_occ(ilutI, a)) {
            b = S.get_orb_from_kpoints(src[0], src[1], a);
            if (b != a && !is_occ(ilutI, b)) {
                // excit_cache(src1,src2,a): zero when a or b coincide with i/j
                if (a != src[0] && a != src[1] && b != src[0] && b != src[1]) {
                    int ex[4] = {src[0], src[1], a, b};
                    elem = std::fabs(S.offdiag_k_hub(ex, false));
                }
            }
        }
        cum_sum += elem; cum_arr[a - 1] = cum_sum; orb_b[a - 1] = b;
    }
    if (cum_sum < EPS) { E.valid = false; E.pgen = 0.0; return; }
    double p_orb;
    const int ind = pick_from_cum_list(cum_arr, S.nbasis, cum_sum, rng, p_orb);
    if (ind < 0) { E.valid = false; E.pgen = 0.0; return; }
    p_orb = 2.0 * p_orb;                       // pick_ab_orbitals_hubbard :1612
    const int orbs[2] = {ind, orb_b[ind - 1]};
    make_double(nI, S.nel, elecs[0], elecs[1], orbs[0], orbs[1], E.nJ, E.ex, E.parity);
    E.ilutJ[0] = ilutI[0]; E.ilutJ[1] = (S.nwords > 1) ? ilutI[1] : 0;
    clr_orb(E.ilutJ, src[0]); clr_orb(E.ilutJ, src[1]); set_orb(E.ilutJ, orbs[0]); set_orb(E.ilutJ, orbs[1]);
    E.pgen = p_elec * p_orb;
    E.valid = true;
}

// pick_biased_elecs, src/excit_gens_int_weighted.F90:722-840 (no pAA bias).
// Random numbers: the reference draws its own number here and rescales it to a pair index,
// (r / pParallel) * nPairs or ((r - pParallel) / (1 - pParallel)) * nPairs.  The engine hands in the number that
// decided "double" in gen_exc_sd, rescaled to [0,1) by the caller in the same way, and multiplies by the
// host-computed constants nPairs / pParallel and nPairs / (1 - pParallel); the fraction left after the pair index
// has been taken off is returned as `rest`, a uniform number independent of the index, which decides exchange in
// GAS_doubles_PCHB_gen_exc.  Same probabilities, one Philox block per attempt instead of two.
inline void pick_biased_elecs(const System &S, const int *nI, double r, int *elecs, int *src, double &pgen, double &rest) {
    const int nA = S.nocc_alpha, nB = S.nocc_beta;
    const int AA = nA * (nA - 1) / 2, BB = nB * (nB - 1) / 2, par = AA + BB, AB = nA * nB;
    const double c_par = (S.p_parallel > 0.0) ? (double)par / S.p_parallel : 0.0;
    const double c_opp = (S.p_parallel < 1.0) ? (double)AB / (1.0 - S.p_parallel) : 0.0;
    int al_req, be_req, al_num[2] = {0, 0}, be_num[2] = {0, 0};
    if (r < S.p_parallel) {
        r = r * c_par;
        int idx = std::min((int)r, par - 1);
        rest = r - (double)idx;
        if (idx < AA) {
            al_req = 2; be_req = 0;
            al_num[0] = (int)std::ceil((1 + std::sqrt(9 + 8 * (double)idx)) / 2);
            al_num[1] = idx + 1 - ((al_num[0] - 1) * (al_num[0] - 2)) / 2;
        } else {
            al_req = 0; be_req = 2;
            idx -= AA;
            be_num[0] = (int)std::ceil((1 + std::sqrt(9 + 8 * (double)idx)) / 2);
            be_num[1] = idx + 1 - ((be_num[0] - 1) * (be_num[0] - 2)) / 2;
        }
        pgen = S.p_parallel / (double)par;
    } else {
        al_req = 1; be_req = 1;
        pgen = (1.0 - S.p_parallel) / (double)AB;
        r = (r - S.p_parallel) * c_opp;
        const int idx = std::min((int)r, AB - 1);
        rest = r - (double)idx;
        al_num[0] = 1 + idx % nA;
        be_num[0] = 1 + (int)std::floor(idx / (double)nA);
    }
    int al_count = 0, be_count = 0, found = 0;
    elecs[0] = elecs[1] = 0;
    for (int i = 1; i <= S.nel; ++i) {
        if (is_alpha(nI[i - 1])) {
            ++al_count;
            if (al_req > 0 && al_count == al_num[al_req - 1]) { elecs[found++] = i; --al_req; }
        } else {
            ++be_count;
            if (be_req > 0 && be_count == be_num[be_req - 1]) { elecs[found++] = i; --be_req; }
        }
        if (al_req == 0 && be_req == 0) break;
    }
    src[0] = nI[elecs[0] - 1]; src[1] = nI[elecs[1] - 1];
}

// sample_AliasTable_t + AliasSampler_t::sample, src/aliasSampling.F90:310-332,433-448
// tables flattened as in include/neci_gpu.h; returns 1-based ab or 0 (empty sampler)
inline int alias_sample(const System &S, int ij, int sampler, Stream &rng, double &prob) {
    const size_t base = ((size_t)(ij - 1) * 3 + sampler) * S.ab_max;
    if (S.alias[base] == 0) { prob = 1.0; return 0; }
    const int sizeArr = S.ab_max;
    const double r = rng.draw53();
    const int pos = (int)(sizeArr * r) + 1;
    const double b = std::max(sizeArr * r + 1 - pos, 0.0);
    const int ind = (b < S.bias[base + pos - 1]) ? pos : S.alias[base + pos - 1];
    prob = S.probs[base + ind - 1];
    return ind;
}

// CreateSingleExcit (uniform singles), src/GenRandSymExcitNUMod.F90:1118-1286
inline void gen_uniform_single(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E) {
    E.ic = 1;
    // construct_class_counts + CheckIfSingleExcits
    std::vector<int> unocc(S.n_classes);
    for (int c = 0; c < S.n_classes; ++c) unocc[c] = (int)S.class_orbs[c].size();
    for (int i = 0; i < S.nel; ++i) unocc[S.class_of_spinorb[nI[i] - 1]] -= 1;
    int ElecsWNoExcits = 0;
    for (int i = 0; i < S.nel; ++i) if (unocc[S.class_of_spinorb[nI[i] - 1]] == 0) ++ElecsWNoExcits;
    if (ElecsWNoExcits == S.nel) { E.valid = false; E.pgen = 0.0; return; }
    int Eleci = 0, cls = 0, NExcit = 0, attempts = 0;
    for (;;) {
        const double r = rng.draw32();
        Eleci = (int)(S.nel * r) + 1;
        cls = S.class_of_spinorb[nI[Eleci - 1] - 1];
        NExcit = unocc[cls];
        if (NExcit != 0) break;
        if (attempts > 250) { E.valid = false; E.err = 1; return; }
        ++attempts;
    }
    const int nOrbs = (int)S.class_orbs[cls].size();
    int Orb = 0; attempts = 0;
    for (;;) {
        const double r = rng.draw32();
        const int ChosenUnocc = (int)(nOrbs * r);
        Orb = S.class_orbs[cls][ChosenUnocc];
        if (!is_occ(ilutI, Orb)) break;
        if (attempts > 250) { E.valid = false; E.err = 1; return; }
        ++attempts;
    }
    int ex2[2];
    make_single(nI, S.nel, Eleci, Orb, E.nJ, ex2, E.parity);
    E.ex[0] = ex2[0]; E.ex[1] = 0; E.ex[2] = ex2[1]; E.ex[3] = 0;
    const double pDoubNew = 1.0 - S.p_singles;                       // :1103
    double pgen = (1 - pDoubNew) / ((double)(NExcit * (S.nel - ElecsWNoExcits)));
    pgen = pgen / S.p_singles;                                        // exc_gen_class_wrappers.F90:53
    E.pgen = pgen;
    E.ilutJ[0] = ilutI[0]; E.ilutJ[1] = (S.nwords > 1) ? ilutI[1] : 0;
    clr_orb(E.ilutJ, ex2[0]); set_orb(E.ilutJ, ex2[1]);
    E.valid = true;
}

// CDF_Sampler_t over the probabilities w[k] of the occupied orbitals (src/CDF_sampling.fpp:57-118): p = w / total,
// cum_p = cumsum(p), the draw is the first position with cum_p >= r (binary_search_first_ge) -- taken here as the
// first position whose running sum of w reaches r * total (the same condition without a division per element).
// Returns the 0-based position (the last one with a non-zero weight if rounding leaves the sum below the threshold).
inline int cdf_pick(const double *w, int n, double total, double r) {
    const double thr = r * total;
    double cum = 0.0; int last = 0;
    for (int k = 0; k < n; ++k) {
        cum += w[k];
        if (w[k] > 0.0) { last = k; if (cum >= thr) return k; }
    }
    return last;
}
// draw_PC_FullyWeightedParticles_t, src/gasci_pchb_doubles_select_particles.fpp:330-384.  The reference's
// constrained_sample (src/aliasSampling.F90:500-535) draws from the alias table until the result is occupied, or, when
// the occupied orbitals hold less than redrawing_cutoff = 0.1 of the weight, from a CDF sampler over them; both give
// p(x) = probs(x) / renormalization.  Engine and checker always take the CDF branch (no data-dependent loop).
// r_first: the number that picks the first particle; the second particle takes the next 53-bit number of the stream.
// Returns false when no second particle can be drawn (srcs = 0 in the reference: a null excitation).
inline bool pick_weighted_elecs(const System &S, const int *nI, double r_first, Stream &rng, int *elecs, int *src, double &pgen) {
    const int nb = (int)S.p_first.size();
    double w[128];
    double renorm_first = 0.0;
    for (int k = 0; k < S.nel; ++k) { w[k] = S.p_first[nI[k] - 1]; renorm_first += w[k]; }
    elecs[0] = elecs[1] = 0; src[0] = src[1] = 0; pgen = 1.0;
    const bool unif_first = S.pchb_particles == 2;      // UNIF-FULL (draw_PC_WeightedParticles_t, :440-478): elecs(1) = int(r * nEl) + 1
    if (!unif_first && near_zero(renorm_first)) return false;
    const int e1 = unif_first ? std::min((int)(r_first * S.nel), S.nel - 1) : cdf_pick(w, S.nel, renorm_first, r_first);
    const int s1 = nI[e1];
    const double p_first1 = unif_first ? 1.0 / (double)S.nel : S.p_first[s1 - 1] / renorm_first;
    const double *row1 = &S.p_second[(size_t)(s1 - 1) * nb];
    double renorm_second1 = 0.0;
    for (int k = 0; k < S.nel; ++k) { w[k] = row1[nI[k] - 1]; renorm_second1 += w[k]; }
    const double r2 = rng.draw53();
    if (near_zero(renorm_second1)) return false;
    const int e2 = cdf_pick(w, S.nel, renorm_second1, r2);
    const int s2 = nI[e2];
    const double p_second1 = row1[s2 - 1] / renorm_second1;
    // the other order (constrained_getProb: 0 when the renormalisation vanishes)
    const double p_first2 = unif_first ? p_first1 : S.p_first[s2 - 1] / renorm_first;
    const double *row2 = &S.p_second[(size_t)(s2 - 1) * nb];
    double renorm_second2 = 0.0;
    for (int k = 0; k < S.nel; ++k) renorm_second2 += row2[nI[k] - 1];
    const double p_second2 = near_zero(renorm_second2) ? 0.0 : row2[s1 - 1] / renorm_second2;
    pgen = unif_first ? (p_second1 + p_second2) / (double)S.nel          // sum(p_second) / nEl
                      : p_first1 * p_second1 + p_first2 * p_second2;
    if (s1 < s2) { elecs[0] = e1 + 1; elecs[1] = e2 + 1; src[0] = s1; src[1] = s2; }
    else { elecs[0] = e2 + 1; elecs[1] = e1 + 1; src[0] = s2; src[1] = s1; }
    return true;
}
// get_pgen_PC_FullyWeightedParticles_t, :386-438
inline double weighted_elecs_pgen(const System &S, const int *nI, int I, int J) {
    const int nb = (int)S.p_first.size();
    double renorm_first = 0.0, rs1 = 0.0, rs2 = 0.0;
    const double *row1 = &S.p_second[(size_t)(I - 1) * nb], *row2 = &S.p_second[(size_t)(J - 1) * nb];
    for (int k = 0; k < S.nel; ++k) { renorm_first += S.p_first[nI[k] - 1]; rs1 += row1[nI[k] - 1]; rs2 += row2[nI[k] - 1]; }
    const double ps1 = near_zero(rs1) ? 0.0 : row1[J - 1] / rs1, ps2 = near_zero(rs2) ? 0.0 : row2[I - 1] / rs2;
    if (S.pchb_particles == 2) return (ps1 + ps2) / (double)S.nel;           // get_pgen_PC_WeightedParticles_t, :476-506
    if (near_zero(renorm_first)) return 0.0;
    const double pf1 = S.p_first[I - 1] / renorm_first, pf2 = S.p_first[J - 1] / renorm_first;
    return pf1 * ps1 + pf2 * ps2;
}

// GAS_doubles_PCHB_gen_exc, src/gasci_pchb_doubles_spatorb_fastweighted.fpp:155-277
inline void gen_pchb_double(const System &S, const int *nI, const uint64_t *ilutI, double r_pair, Stream &rng, Excitation &E) {
    E.ic = 2;
    int elecs[2], src[2];
    double pGen, rest;
    if (S.pchb_particles != 0) {
        // FULL-FULL / UNIF-FULL: particles from the selector's tables; the exchange decision takes a 32-bit number of the
        // attempt's second block, the alias sample the 53-bit number after it (block 0: single / double + first
        // particle, second particle)
        if (!pick_weighted_elecs(S, nI, r_pair, rng, elecs, src, pGen)) {
            E.ex[0] = E.ex[1] = E.ex[2] = E.ex[3] = 0; E.valid = false; E.pgen = 1.0; return;
        }
        rest = rng.draw32();
    } else
    pick_biased_elecs(S, nI, r_pair, elecs, src, pGen, rest);
    const int ij = (int)fuseIndex(gtID(src[0]), gtID(src[1]));
    int spin[2] = {is_beta(src[0]) ? 1 : 0, is_beta(src[1]) ? 1 : 0};   // getSpinIndex: 0 alpha, 1 beta
    int sampler;
    if (spin[0] == spin[1]) sampler = 0;                                  // SAME_SPIN
    else {
        const double pe = S.p_exch[ij - 1];
        if (rest < pe) { sampler = 2; pGen *= pe; std::swap(spin[0], spin[1]); }          // OPP_SPIN_EXCH
        else { sampler = 1; pGen *= (1.0 - pe); }                                          // OPP_SPIN_NO_EXCH
    }
    double pGenHoles;
    const int ab = alias_sample(S, ij, sampler, rng, pGenHoles);
    bool invalid;
    int orbs[2] = {0, 0};
    if (ab == 0) invalid = true;
    else {
        orbs[0] = 2 * S.tgt_orbs[2 * (ab - 1)] - spin[0];
        orbs[1] = 2 * S.tgt_orbs[2 * (ab - 1) + 1] - spin[1];
        invalid = (orbs[0] == 0 || orbs[1] == 0);
        for (int i = 0; i < S.nel; ++i) if (nI[i] == orbs[0] || nI[i] == orbs[1]) invalid = true;
    }
    if (!invalid && near_zero(pGenHoles)) invalid = true;
    E.ex[0] = src[0]; E.ex[1] = src[1]; E.ex[2] = orbs[0]; E.ex[3] = orbs[1];
    if (invalid) { E.valid = false; E.pgen = pGen; return; }
    make_double(nI, S.nel, elecs[0], elecs[1], orbs[0], orbs[1], E.nJ, E.ex, E.parity);
    E.ilutJ[0] = ilutI[0]; E.ilutJ[1] = (S.nwords > 1) ? ilutI[1] : 0;
    clr_orb(E.ilutJ, src[0]); clr_orb(E.ilutJ, src[1]); set_orb(E.ilutJ, orbs[0]); set_orb(E.ilutJ, orbs[1]);
    E.pgen = pGen * pGenHoles;
    E.valid = true;
}

// GAS_doubles_PCHB_get_pgen, :286-326 (+ get_pgen_pick_biased_elecs, excit_gens_int_weighted.F90:849)
inline double pchb_double_get_pgen(const System &S, const int *ex, const int *nI = nullptr) {
    const int nex[4] = {gtID(ex[0]), gtID(ex[1]), gtID(ex[2]), gtID(ex[3])};
    const int ij = (int)fuseIndex(nex[0], nex[1]), ab = (int)fuseIndex(nex[2], nex[3]);
    const int nA = S.nocc_alpha, nB = S.nocc_beta;
    const int par = nA * (nA - 1) / 2 + nB * (nB - 1) / 2, AB = nA * nB;
    const bool same = (is_beta(ex[0]) == is_beta(ex[1]));
    double pgen = same ? S.p_parallel / (double)par : (1.0 - S.p_parallel) / (double)AB;
    if (S.pchb_particles != 0) pgen = nI ? weighted_elecs_pgen(S, nI, ex[0], ex[1]) : 0.0;   // depends on the determinant
    int sampler;
    if (same) sampler = 0;
    else if ((is_beta(ex[0]) == is_beta(ex[2])) || nex[2] == nex[3]) { sampler = 1; pgen *= (1.0 - S.p_exch[ij - 1]); }
    else { sampler = 2; pgen *= S.p_exch[ij - 1]; }
    const size_t base = ((size_t)(ij - 1) * 3 + sampler) * S.ab_max;
    if (S.alias[base] == 0) return 0.0;
    return pgen * S.probs[base + ab - 1];
}

// gen_exc_sd, src/excitation_generators.F90:112-138
// One number decides single / double; a double re-uses it, rescaled to [0,1), for the electron pair (see
// pick_biased_elecs); a single continues with the second block of the attempt's stream (word 4).
inline void gen_excit_pchb(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E) {
    const double u = rng.draw53();
    if (u < S.p_singles) {
        rng.pos = 4;
        gen_uniform_single(S, nI, ilutI, rng, E);
        E.pgen = E.pgen * S.p_singles;
    } else {
        const double inv_1m_ps = 1.0 / (1.0 - S.p_singles);
        gen_pchb_double(S, nI, ilutI, (u - S.p_singles) * inv_1m_ps, rng, E);
        E.pgen = E.pgen * S.p_doubles;
    }
}

inline void generate_excitation(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E) {
    E = Excitation();
    switch (S.type) {
        case NECI_SYS_FCIDUMP_PCHB: gen_excit_pchb(S, nI, ilutI, rng, E); break;
        case NECI_SYS_HUBBARD_RS:   gen_excit_rs_hubbard(S, nI, ilutI, rng, E); break;
        case NECI_SYS_HUBBARD_K:    gen_excit_k_space_hub(S, nI, ilutI, rng, E); break;
        default: E.valid = false;
    }
}

// get_spawn_helement = get_helement_det_only, src/Determinants.F90:508-554
inline double get_spawn_helement(const System &S, const int *nI, const Excitation &E) {
    if (S.type == NECI_SYS_HUBBARD_RS) return S.offdiag_rs_hub(E.ex[0], E.ex[2], E.parity);
    if (S.type == NECI_SYS_HUBBARD_K) return S.offdiag_k_hub(E.ex, E.parity);
    double h;
    if (E.ic == 1) h = S.sltcnd_1_kernel(nI, E.ex[0], E.ex[2]);
    else h = S.sltcnd_2_kernel(E.ex);
    return E.parity ? -h : h;
}

// ---- HPHF functions (src/HPHFIntegrals.fpp, src/HPHFRandExcit.F90, src/DetBitOps.F90:648-740,819-848) ----------
// spin_sym_ilut: swap the alpha and beta occupation of every spatial orbital
inline void spin_sym_ilut(const uint64_t *a, uint64_t *b, int nw) {
    for (int w = 0; w < nw; ++w) b[w] = ((a[w] & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((a[w] & 0x5555555555555555ull) << 1);
}
inline bool TestClosedShellDet(const uint64_t *a, int nw) {
    for (int w = 0; w < nw; ++w) if (((a[w] & 0xAAAAAAAAAAAAAAAAull) >> 1) ^ (a[w] & 0x5555555555555555ull)) return false;
    return true;
}
// CalcOpenOrbs: beta electrons whose alpha partner is empty (= half the singly occupied orbitals for Ms = 0)
inline int CalcOpenOrbs(const uint64_t *a, int nw) {
    int n = 0;
    for (int w = 0; w < nw; ++w) n += __builtin_popcountll(~((a[w] & 0xAAAAAAAAAAAAAAAAull) >> 1) & (a[w] & 0x5555555555555555ull));
    return n;
}
// DetBitLT (src/DetBitOps.F90:502-534): signed word comparison, word 0 first
inline int DetBitLT(const uint64_t *a, const uint64_t *b, int nw) {
    for (int w = 0; w < nw; ++w) {
        if ((int64_t)a[w] < (int64_t)b[w]) return 1;
        if ((int64_t)a[w] > (int64_t)b[w]) return -1;
    }
    return 0;
}
// FindBitExcitLevel(..., t_hphf_ic = .true.) (src/DetBitOps.F90:140-170, 819-848)
inline int excit_level_hphf(const System &S, const uint64_t *a, const uint64_t *b) {
    const int nw = S.nwords;
    if (!S.t_hphf || (TestClosedShellDet(a, nw) && TestClosedShellDet(b, nw))) return excit_level(a, b, nw);
    uint64_t a2[2], b2[2];
    spin_sym_ilut(a, a2, nw); spin_sym_ilut(b, b2, nw);
    return std::min(std::min(excit_level(a, b, nw), excit_level(a, b2, nw)), std::min(excit_level(a2, b, nw), excit_level(a2, b2, nw)));
}
// hphf_off_diag_helement_norm (src/HPHFIntegrals.fpp:62-150), even S
inline double hphf_off_diag_helement(const System &S, const uint64_t *iI, const uint64_t *iJ) {
    const int nw = S.nwords;
    if (DetBitLT(iI, iJ, nw) == 0) return 0.0;
    double hel = get_helement(S, iI, iJ);
    if (TestClosedShellDet(iI, nw)) {
        if (!TestClosedShellDet(iJ, nw)) hel = hel * std::sqrt(2.0);
    } else if (TestClosedShellDet(iJ, nw)) {
        hel = hel * std::sqrt(2.0);
    } else {
        uint64_t iI2[2] = {0, 0};
        spin_sym_ilut(iI, iI2, nw);                                  // FindExcitBitDetSym
        if (excit_level(iI2, iJ, nw) <= 2) {
            const int OpenOrbsI = CalcOpenOrbs(iI, nw);
            const double MatEl2 = get_helement(S, iI2, iJ);
            if (OpenOrbsI % 2 == 0) hel = hel + MatEl2; else hel = hel - MatEl2;
        }
    }
    return hel;
}
// hphf_diag_helement (src/HPHFIntegrals.fpp:348-411), even S; ECore included by get_helement
inline double hphf_diag_helement(const System &S, const uint64_t *iI) {
    const int nw = S.nwords;
    double hel = get_helement(S, iI, iI);
    if (!TestClosedShellDet(iI, nw)) {
        uint64_t iI2[2] = {0, 0};
        spin_sym_ilut(iI, iI2, nw);
        if (excit_level(iI, iI2, nw) <= 2) {
            const double MatEl2 = get_helement(S, iI, iI2);
            if (CalcOpenOrbs(iI, nw) % 2 == 1) hel = hel - MatEl2; else hel = hel + MatEl2;
        }
    }
    return hel;
}
// CalcNonUniPGen (src/HPHFRandExcit.F90:686-821) for the PCHB class generator: get_pgen_sd
// (src/excitation_generators.F90:141-158) with UniformSingles_get_pgen / calc_pgen_symrandexcit2 and
// GAS_doubles_PCHB_get_pgen
inline double calc_pgen_pchb(const System &S, const int *nI, const uint64_t *ilutI, const int *ex, int ic) {
    if (ic == 1) {
        int ElecsWNoExcits = 0;
        std::vector<int> occ(S.n_classes, 0), unocc(S.n_classes, 0);
        for (int c = 0; c < S.n_classes; ++c) {
            for (int o : S.class_orbs[c]) { if (is_occ(ilutI, o)) ++occ[c]; else ++unocc[c]; }
            if (unocc[c] == 0) ElecsWNoExcits += occ[c];
        }
        const int NExcitA = unocc[S.class_of_spinorb[ex[0] - 1]];
        double pgen = (1 - S.p_doubles) / ((double)(NExcitA * (S.nel - ElecsWNoExcits)));
        pgen = pgen / S.p_singles;
        return S.p_singles * pgen;
    }
    if (ic == 2) return (1.0 - S.p_singles) * pchb_double_get_pgen(S, ex);
    return 0.0;
}
// gen_hphf_excit (src/HPHFRandExcit.F90:175-476) around the PCHB generator.  The matrix element between the two
// HPHF functions is hphf_off_diag_helement_norm (what get_spawn_helement returns with tGenMatHEl = .false.,
// src/fcimc_initialisation.fpp:2198-2203; the in-generator evaluation is the same number).
inline void gen_hphf_excit(const System &S, const int *nI, const uint64_t *ilutI, Stream &rng, Excitation &E, double &HEl) {
    const int nw = S.nwords;
    HEl = 0.0;
    gen_excit_pchb(S, nI, ilutI, rng, E);
    if (!E.valid) return;
    if (!TestClosedShellDet(E.ilutJ, nw)) {
        uint64_t iJ2[2] = {0, 0};
        spin_sym_ilut(E.ilutJ, iJ2, nw);                             // ReturnAlphaOpenDet
        const bool tSwapped = DetBitLT(E.ilutJ, iJ2, nw) == 1;
        const int ExcitLevel = excit_level(ilutI, iJ2, nw);          // to the determinant that was NOT generated
        if (ExcitLevel == 0) { E.valid = false; return; }            // excitation inside one HPHF function: null
        if (ExcitLevel <= 2) {
            int ex2[4]; bool tSign;
            excitation_between(ilutI, iJ2, S.nbasis, S.nel, ex2, tSign);
            E.pgen = E.pgen + calc_pgen_pchb(S, nI, ilutI, ex2, ExcitLevel);
        }
        if (tSwapped) { E.ilutJ[0] = iJ2[0]; E.ilutJ[1] = iJ2[1]; decode(E.ilutJ, S.nbasis, E.nJ); }
    }
    HEl = hphf_off_diag_helement(S, ilutI, E.ilutJ);
}

// get_diagonal_matel (src/matel_getter.F90:30-58) minus nothing; caller subtracts Hii
inline double get_diagonal_matel(const System &S, const uint64_t *ilut) {
    if (S.t_hphf) return hphf_diag_helement(S, ilut);
    if (S.type == NECI_SYS_HUBBARD_RS) return S.diag_rs_hub(ilut);
    int nI[128]; decode(ilut, S.nbasis, nI);
    return S.sltcnd_0(nI) + S.ecore;
}
// get_off_diagonal_matel, src/matel_getter.F90:61-105
inline double get_off_diagonal_matel(const System &S, const uint64_t *ilut, const uint64_t *ilut_ref) {
    const int exlevel = excit_level_hphf(S, ilut_ref, ilut);
    if (exlevel == 2 || (exlevel == 1 && S.t_no_brillouin))
        return S.t_hphf ? hphf_off_diag_helement(S, ilut_ref, ilut) : get_helement(S, ilut, ilut_ref);
    return 0.0;
}

}  // namespace orc
