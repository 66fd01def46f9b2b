// Host-side mirror of the parts of NECI's Fortran host that feed the engine:
// system setup (synthetic FCIDUMP integrals, Hubbard lattices), the PCHB table
// initialisation, the hashing tables and the shift update.  In a real
// deployment these stay in the Fortran host (SURVEY.md §2b "stays Fortran");
// this library exists so that the engine can be driven stand-alone from C++ or
// Python (tests, bench) with inputs of exactly the shapes the Fortran host
// would hand over.  CPU-only, no CUDA.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>

namespace {

// small counter-based generator for setup-time random numbers (the reference
// uses dSFMT here; the sequence differs, the recipes do not)
struct SetupRng {
    uint64_t s;
    explicit SetupRng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double real2() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
};

inline int64_t tri(int64_t a, int64_t b) { return (a > b) ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; }
// UMatInd over spatial orbitals, src/UMatCache.F90:257-296
inline int64_t umat_ind(int i, int j, int k, int l) { return tri(tri(i, k), tri(j, l)); }
inline int64_t fuse(int64_t x, int64_t y) { return (x < y) ? x + y * (y - 1) / 2 : y + x * (x - 1) / 2; }
inline bool is_beta(int o) { return o & 1; }
inline int gtid(int o) { return (o - 1) / 2 + 1; }

}  // namespace

extern "C" {

// number of UMAT entries for n_spat spatial orbitals
int64_t neci_host_umat_size(int32_t n_spat) {
    const int64_t np = (int64_t)n_spat * (n_spat + 1) / 2;
    return np * (np + 1) / 2;
}

// generate_random_integrals, src/unit_test_helper_excitgen.F90:371-485 (RHF,
// hermitian): umatRand(i,j) = r^2 if r < sparse; (ij|kl) = sqrt(umatRand(i,j) umatRand(k,l));
// h_ij = r if r < sparseT.  diag_shift adds diag_shift*i to h_ii (SURVEY §8d:
// gives an aufbau reference).  umat: packed UMAT (0-based storage of the
// 1-based UMatInd); tmat: nbasis x nbasis spin-orbital TMAT2D, column-major.
int neci_host_random_fcidump(int32_t n_spat, double sparse, double sparseT, uint64_t seed,
                             double diag_shift, double *umat, double *tmat) {
    SetupRng rng(seed);
    const int n = n_spat;
    std::vector<double> ur((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double r = rng.real2();
            if (r < sparse) { ur[(size_t)i * n + j] = r * r; ur[(size_t)j * n + i] = r * r; }
        }
    const int64_t nu = neci_host_umat_size(n_spat);
    std::fill(umat, umat + nu, 0.0);
    // write_4index: i, j<=i, k>=i, l<=k ; FCIDUMP line (ij|kl) -> UMAT(UMatInd(i,k,j,l))
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= i; ++j)
            for (int k = i; k <= n; ++k)
                for (int l = 1; l <= k; ++l) {
                    const double m = std::sqrt(ur[(size_t)(i - 1) * n + (j - 1)] * ur[(size_t)(k - 1) * n + (l - 1)]);
                    if (m > 1e-13) umat[umat_ind(i, k, j, l) - 1] = m;
                }
    const int nb = 2 * n;
    std::fill(tmat, tmat + (size_t)nb * nb, 0.0);
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= i; ++j) {
            const double r = rng.real2();
            double v = (r < sparseT) ? r : 0.0;
            if (i == j) v += diag_shift * i;
            for (int s = 0; s < 2; ++s) {        // same-spin blocks only
                const int a = 2 * i - s, b = 2 * j - s;
                tmat[(size_t)(a - 1) + (size_t)nb * (b - 1)] = v;
                tmat[(size_t)(b - 1) + (size_t)nb * (a - 1)] = v;
            }
        }
    return 0;
}

// RandomOrbIndex / RandomHash2, src/fcimc_initialisation.fpp:862-942:
// distinct values INT(nBasis*r*1000)+1.
int neci_host_random_hash_tables(int32_t nbasis, uint64_t seed, int32_t *random_orb_index, int32_t *random_hash2) {
    SetupRng rng(seed ^ 0xA5A5A5A5ull);
    for (int pass = 0; pass < 2; ++pass) {
        int32_t *t = pass ? random_hash2 : random_orb_index;
        std::fill(t, t + nbasis, 0);
        for (int i = 0; i < nbasis; ++i) {
            for (;;) {
                const int chosen = (int)(nbasis * rng.real2() * 1000) + 1;
                bool used = false;
                for (int j = 0; j < nbasis; ++j) if (t[j] == chosen) { used = true; break; }
                if (!used) { t[i] = chosen; break; }
            }
        }
    }
    return 0;
}

// ----------------------------------------------------------------------------
// PCHB tables, GAS_doubles_PCHB_compute_samplers
// (src/gasci_pchb_doubles_spatorb_fastweighted.fpp:329-445) for one supergroup,
// alias tables per init_AliasTable_t (src/aliasSampling.F90:209-297).
// Layout documented in include/neci_gpu.h (neci_gpu_set_pchb).
// ----------------------------------------------------------------------------
static void build_alias(const std::vector<double> &w, double *probs, double *bias, int32_t *alias) {
    const int n = (int)w.size();
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += w[i];
    if (std::fabs(sum) <= 1e-13) {           // near_zero(sum(arr)): sampler left unassociated
        for (int i = 0; i < n; ++i) { probs[i] = 0.0; bias[i] = 0.0; alias[i] = 0; }
        return;
    }
    for (int i = 0; i < n; ++i) { bias[i] = w[i] / sum * n; probs[i] = w[i] / sum; alias[i] = i + 1; }
    std::vector<int> overfull(n), underfull(n);
    int cV = 0, cU = 0;
    auto assign = [&](int i) { if (bias[i - 1] > 1) overfull[cV++] = i; else underfull[cU++] = i; };
    for (int i = 1; i <= n; ++i) assign(i);
    std::reverse(overfull.begin(), overfull.begin() + cV);
    while (cV != 0 && cU != 0) {
        const int i = overfull[cV - 1], j = underfull[cU - 1];
        alias[j - 1] = i;
        bias[i - 1] = bias[i - 1] + bias[j - 1] - 1.0;
        --cU; --cV;
        assign(i);
    }
    for (int k = 0; k < cV; ++k) { bias[overfull[k] - 1] = 1.0; alias[overfull[k] - 1] = overfull[k]; }
    for (int k = 0; k < cU; ++k) { bias[underfull[k] - 1] = 1.0; alias[underfull[k] - 1] = underfull[k]; }
}

// one alias table (init_AliasTable_t, src/aliasSampling.F90:209-297), exposed for the sampler tests
int neci_host_alias_build(int32_t n, const double *w, double *probs, double *bias, int32_t *alias) {
    build_alias(std::vector<double>(w, w + n), probs, bias, alias);
    return 0;
}

int neci_host_pchb_dims(int32_t n_spat, int32_t *ij_max, int32_t *ab_max) {
    *ij_max = (int32_t)fuse(n_spat, n_spat); *ab_max = *ij_max; return 0;
}

int neci_host_pchb_build(int32_t n_spat, const double *umat, double *probs, double *bias,
                         int32_t *alias, double *p_exch, int32_t *tgt_orbs) {
    const int nBI = n_spat;
    const int ijMax = (int)fuse(nBI, nBI), abMax = ijMax;
    for (int a = 1; a <= nBI; ++a)
        for (int b = 1; b <= a; ++b) {
            const int ab = (int)fuse(a, b);
            tgt_orbs[2 * (ab - 1)] = b; tgt_orbs[2 * (ab - 1) + 1] = a;
        }
    std::fill(probs, probs + (size_t)ijMax * 3 * abMax, 0.0);
    std::fill(bias, bias + (size_t)ijMax * 3 * abMax, 0.0);
    std::fill(alias, alias + (size_t)ijMax * 3 * abMax, 0);
    std::vector<double> pExch(ijMax, 0.0), pNoExch(ijMax, 1.0), w(abMax);
    auto Ms = [](int o) { return is_beta(o) ? -1 : 1; };
    auto umat_el = [&](int i, int j, int k, int l) { return umat[umat_ind(i, j, k, l) - 1]; };
    // nI_invariant_sltcnd_excit -> sltcnd_2_kernel, src/sltcnd.fpp:690-708
    auto weight = [&](const int ex[4]) {
        double hel = 0.0;
        if (Ms(ex[0]) == Ms(ex[2]) && Ms(ex[1]) == Ms(ex[3])) hel = umat_el(gtid(ex[0]), gtid(ex[1]), gtid(ex[2]), gtid(ex[3]));
        if (Ms(ex[0]) == Ms(ex[3]) && Ms(ex[1]) == Ms(ex[2])) hel -= umat_el(gtid(ex[0]), gtid(ex[1]), gtid(ex[3]), gtid(ex[2]));
        return std::fabs(hel);
    };
    auto to_spin_orb = [](int orb, bool alpha) { return alpha ? 2 * orb : 2 * orb - 1; };
    enum { SAME_SPIN = 1, OPP_SPIN_NO_EXCH = 2, OPP_SPIN_EXCH = 3 };
    for (int i_exch = 1; i_exch <= 3; ++i_exch)
        for (int i = 1; i <= nBI; ++i) {
            int ex[4];
            ex[0] = to_spin_orb(i, true);
            for (int j = i; j <= nBI; ++j) {
                if (i_exch == SAME_SPIN && i == j) continue;
                const int ij = (int)fuse(i, j);
                std::fill(w.begin(), w.end(), 0.0);
                ex[1] = to_spin_orb(j, i_exch == SAME_SPIN);
                for (int a = 1; a <= nBI; ++a) {
                    ex[2] = to_spin_orb(a, i_exch == SAME_SPIN || i_exch == OPP_SPIN_NO_EXCH);
                    if (ex[2] == ex[0] || ex[2] == ex[1]) continue;
                    for (int b = a; b <= nBI; ++b) {
                        if (i_exch == OPP_SPIN_EXCH && a == b) continue;
                        const int ab = (int)fuse(a, b);
                        ex[3] = to_spin_orb(b, i_exch == SAME_SPIN || i_exch == OPP_SPIN_EXCH);
                        if (ex[3] == ex[0] || ex[3] == ex[1] || ex[3] == ex[2]) continue;
                        // canonicalize: sort sources and targets (sign irrelevant under abs)
                        int c[4] = {std::min(ex[0], ex[1]), std::max(ex[0], ex[1]), std::min(ex[2], ex[3]), std::max(ex[2], ex[3])};
                        w[ab - 1] = weight(c);
                    }
                }
                const size_t base = ((size_t)(ij - 1) * 3 + (i_exch - 1)) * abMax;
                build_alias(w, probs + base, bias + base, alias + base);
                double s = 0.0; for (double x : w) s += x;
                if (i_exch == OPP_SPIN_EXCH) pExch[ij - 1] = s;
                if (i_exch == OPP_SPIN_NO_EXCH) pNoExch[ij - 1] = s;
            }
        }
    for (int ij = 0; ij < ijMax; ++ij) {
        const double d = pExch[ij] + pNoExch[ij];
        p_exch[ij] = (std::fabs(d) <= 1e-13) ? 0.0 : pExch[ij] / d;
    }
    return 0;
}

// Particle-selection tables of PCHB_ParticleSelection FULL-FULL (PC_FullyWeightedParticles_t,
// src/gasci_pchb_doubles_select_particles.fpp:268-328): IJ_weights(I, J) = sum over the hole pairs of the weights
// of the samplers of the spin-orbital pair (I, J), accumulated as GAS_doubles_PCHB_compute_samplers does it
// (src/gasci_pchb_doubles_spatorb_fastweighted.fpp:374-420), then normalised the way AliasSampler_t::setup_entry
// normalises (probs = w / sum(w)): p_first[I] from sum_J IJ_weights(J, I), p_second[I][J] = p(J | I) from column I.
// Only the probabilities are needed: the engine draws from them restricted to the occupied orbitals (constrained
// sampling).  Spin orbitals 1-based as in NECI (2 * spatial = alpha, 2 * spatial - 1 = beta), arrays 0-based.
int neci_host_pchb_particle_probs(int32_t n_spat, const double *umat, double *p_first, double *p_second) {
    const int nBI = n_spat, nb = 2 * n_spat;
    const int abMax = (int)fuse(nBI, nBI);
    std::vector<double> IJ((size_t)nb * nb, 0.0), w(abMax);
    auto at = [&](int I, int J) -> double & { return IJ[(size_t)(I - 1) * nb + (J - 1)]; };
    auto Ms = [](int o) { return is_beta(o) ? -1 : 1; };
    auto umat_el = [&](int i, int j, int k, int l) { return umat[umat_ind(i, j, k, l) - 1]; };
    auto weight = [&](const int ex[4]) {
        double hel = 0.0;
        if (Ms(ex[0]) == Ms(ex[2]) && Ms(ex[1]) == Ms(ex[3])) hel = umat_el(gtid(ex[0]), gtid(ex[1]), gtid(ex[2]), gtid(ex[3]));
        if (Ms(ex[0]) == Ms(ex[3]) && Ms(ex[1]) == Ms(ex[2])) hel -= umat_el(gtid(ex[0]), gtid(ex[1]), gtid(ex[3]), gtid(ex[2]));
        return std::fabs(hel);
    };
    auto to_spin_orb = [](int orb, bool alpha) { return alpha ? 2 * orb : 2 * orb - 1; };
    enum { SAME_SPIN = 1, OPP_SPIN_NO_EXCH = 2, OPP_SPIN_EXCH = 3 };
    for (int i_exch = 1; i_exch <= 3; ++i_exch)
        for (int i = 1; i <= nBI; ++i) {
            int ex[4];
            ex[0] = to_spin_orb(i, true);
            for (int j = i; j <= nBI; ++j) {
                if (i_exch == SAME_SPIN && i == j) continue;
                std::fill(w.begin(), w.end(), 0.0);
                ex[1] = to_spin_orb(j, i_exch == SAME_SPIN);
                for (int a = 1; a <= nBI; ++a) {
                    ex[2] = to_spin_orb(a, i_exch == SAME_SPIN || i_exch == OPP_SPIN_NO_EXCH);
                    if (ex[2] == ex[0] || ex[2] == ex[1]) continue;
                    for (int b = a; b <= nBI; ++b) {
                        if (i_exch == OPP_SPIN_EXCH && a == b) continue;
                        ex[3] = to_spin_orb(b, i_exch == SAME_SPIN || i_exch == OPP_SPIN_EXCH);
                        if (ex[3] == ex[0] || ex[3] == ex[1] || ex[3] == ex[2]) continue;
                        int c[4] = {std::min(ex[0], ex[1]), std::max(ex[0], ex[1]), std::min(ex[2], ex[3]), std::max(ex[2], ex[3])};
                        w[fuse(a, b) - 1] = weight(c);
                    }
                }
                double sw = 0.0; for (double x : w) sw += x;
                {
                    const int I = ex[0], J = ex[1];
                    at(I, J) = at(I, J) + sw; at(J, I) = at(I, J);
                }
                if (i != j) {                                   // the same pair of spatial orbitals with the spins flipped
                    const int I = ex[0] - 1, J = (i_exch == SAME_SPIN) ? ex[1] - 1 : ex[1] + 1;
                    at(I, J) = at(I, J) + sw; at(J, I) = at(I, J);
                }
            }
        }
    // I_sampler: weights sum(IJ_weights(:, :), dim = 1), i.e. the column sums; J_sampler entry I: IJ_weights(:, I)
    std::vector<double> col(nb, 0.0);
    double tot = 0.0;
    for (int I = 1; I <= nb; ++I) {
        double c = 0.0;
        for (int J = 1; J <= nb; ++J) c += at(J, I);
        col[I - 1] = c; tot += c;
    }
    for (int I = 1; I <= nb; ++I) {
        p_first[I - 1] = (std::fabs(tot) <= 1e-13) ? 0.0 : col[I - 1] / tot;
        for (int J = 1; J <= nb; ++J)
            p_second[(size_t)(I - 1) * nb + (J - 1)] = (std::fabs(col[I - 1]) <= 1e-13) ? 0.0 : at(J, I) / col[I - 1];
    }
    return 0;
}

// ----------------------------------------------------------------------------
// Lattices.  Site s = x + lx*y (0-based) is spatial orbital s+1; spin orbitals
// 2(s+1)-1 (beta) and 2(s+1) (alpha).
// ----------------------------------------------------------------------------
// Real-space Hubbard on an lx x ly square lattice (periodic if pbc): neighbour
// lists per spin orbital (padded with 0, width max_neigh = 4) and TMAT2D with
// bhub = -t between neighbours of equal spin (src/real_space_hubbard.F90:151-160).
int neci_host_hubbard_rs_setup(int32_t lx, int32_t ly, int32_t pbc, double t, int32_t *neighbours, double *tmat) {
    const int ns = lx * ly, nb = 2 * ns, mx = 4;
    std::fill(neighbours, neighbours + (size_t)nb * mx, 0);
    std::fill(tmat, tmat + (size_t)nb * nb, 0.0);
    for (int y = 0; y < ly; ++y)
        for (int x = 0; x < lx; ++x) {
            const int s = x + lx * y;
            int cand[4], nc = 0;
            const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
            for (int d = 0; d < 4; ++d) {
                int xx = x + dx[d], yy = y + dy[d];
                if (pbc) { xx = (xx + lx) % lx; yy = (yy + ly) % ly; }
                else if (xx < 0 || xx >= lx || yy < 0 || yy >= ly) continue;
                const int s2 = xx + lx * yy;
                if (s2 == s) continue;
                bool dup = false;
                for (int q = 0; q < nc; ++q) if (cand[q] == s2) dup = true;
                if (!dup) cand[nc++] = s2;
            }
            std::sort(cand, cand + nc);
            for (int spin = 0; spin < 2; ++spin) {
                const int o = 2 * (s + 1) - spin;      // spin=1 beta (odd), 0 alpha (even)
                for (int q = 0; q < nc; ++q) {
                    const int o2 = 2 * (cand[q] + 1) - spin;
                    neighbours[(size_t)(o - 1) * mx + q] = o2;
                    tmat[(size_t)(o - 1) + (size_t)nb * (o2 - 1)] = -t;
                }
            }
        }
    return 0;
}

// k-space Hubbard on an lx x ly periodic mesh.  k-points are ordered by
// kinetic energy (ties: by mesh index) so that the lowest orbitals form the
// reference.  eps_k = -2t (cos kx + cos ky); ksum/kdiff: index tables of
// k1+k2 / k1-k2 modulo the mesh.
int neci_host_hubbard_k_setup(int32_t lx, int32_t ly, double t, int32_t *ksum, int32_t *kdiff, double *eps_k) {
    const int nk = lx * ly;
    std::vector<int> order(nk), rank(nk);
    std::vector<double> e(nk);
    const double pi = 3.14159265358979323846;
    for (int ky = 0; ky < ly; ++ky)
        for (int kx = 0; kx < lx; ++kx) {
            double v = -2.0 * t * std::cos(2 * pi * kx / lx);
            if (ly > 1) v += -2.0 * t * std::cos(2 * pi * ky / ly);
            e[kx + lx * ky] = v;
        }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return e[a] < e[b] - 1e-12; });
    for (int i = 0; i < nk; ++i) { rank[order[i]] = i; eps_k[i] = e[order[i]]; }
    for (int a = 0; a < nk; ++a)
        for (int b = 0; b < nk; ++b) {
            const int ma = order[a], mb = order[b];
            const int ax = ma % lx, ay = ma / lx, bx = mb % lx, by = mb / lx;
            ksum[a * nk + b] = rank[(ax + bx) % lx + lx * ((ay + by) % ly)];
            kdiff[a * nk + b] = rank[(ax - bx + lx) % lx + lx * ((ay - by + ly) % ly)];
        }
    return 0;
}

// ----------------------------------------------------------------------------
// update_shift, src/fcimc_iter_utilities.F90:1063-1072,1156-1158 (single run,
// fixed tau, no target-growth refinements):
//   S <- S - SftDamp * ln(AllGrowRate) / (tau * StepsSft)
// ----------------------------------------------------------------------------
double neci_host_update_shift(double diag_sft, double sft_damp, double tau, int32_t steps_sft,
                              double all_tot_parts, double all_tot_parts_old) {
    if (all_tot_parts_old <= 0.0 || all_tot_parts <= 0.0) return diag_sft;
    const double grow = all_tot_parts / all_tot_parts_old;
    return diag_sft - (std::log(grow) * sft_damp) / (tau * steps_sft);
}

}  // extern "C"
