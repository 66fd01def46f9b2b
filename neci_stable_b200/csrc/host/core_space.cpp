// Host-side mirror of the semi-stochastic set-up that the Fortran host runs before the iteration loop
// (SURVEY.md §8 row a17 consumes its output through neci_gpu_set_core_space):
//
//   neci_host_get_helement      get_helement for determinant pairs of an FCIDUMP system
//                               (src/Determinants.F90:508-554 -> src/sltcnd.fpp:585-708)
//   neci_host_sd_space          generate_sing_doub_determinants (src/semi_stoch_gen.F90:537-604), `doubles-core`
//   neci_host_det_node          DetermineDetNode / get_det_block (src/load_balance_calcnodes.F90:25-117)
//   neci_host_ham_apply         sum_j H_ij v_j over two determinant lists (con_space_vecs of init_trial_wf,
//                               src/trial_wf_gen.F90)
//   neci_host_core_ham_*        the sparse core Hamiltonian of one rank (calc_determ_hamil_sparse and
//                               calc_determ_hamil_sparse_hphf, src/sparse_arrays.F90:426-690; row contents as calc_determ_hamil_opt,
//                               src/fast_determ_hamil.F90:1421-1507: non-zero off-diagonal elements, then the
//                               diagonal H_ii - Hii as the last entry of the row)
//
// Everything works on occupation words (two 64-bit words cover nBasis <= 128), never on orbital lists: holes and
// particles of a pair are the set bits of I & (I ^ J) and J & (I ^ J), parities are popcounts of masked words.
// The Hamiltonian build is a brute-force sweep of all pairs with a popcount prefilter, rows dealt out to host
// threads in chunks; it is set-up code, not part of the iteration.  CPU only, no CUDA.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include <algorithm>

namespace {

inline int64_t tri(int64_t a, int64_t b) { return (a > b) ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; }

struct Det2 { uint64_t w[2]; };

inline int popc(uint64_t x) { return __builtin_popcountll(x); }
inline bool has(const Det2 &d, int b) { return (d.w[b >> 6] >> (b & 63)) & 1ull; }
inline void flip(Det2 &d, int b) { d.w[b >> 6] ^= 1ull << (b & 63); }
// lowest set bit position (0-based over both words); d must be non-empty
inline int low_bit(const Det2 &d) { return d.w[0] ? __builtin_ctzll(d.w[0]) : 64 + __builtin_ctzll(d.w[1]); }
inline int pop_low(Det2 &d) {
    const int b = low_bit(d);
    d.w[b >> 6] &= d.w[b >> 6] - 1ull;
    return b;
}
// number of occupied orbitals strictly between bit positions x and y
inline int between(const Det2 &d, int x, int y) {
    if (x > y) std::swap(x, y);
    int n = 0;
    for (int w = 0; w < 2; ++w) {
        const int lo = x + 1 - 64 * w, hi = y - 64 * w;          // bits [lo, hi) of this word
        if (hi <= 0 || lo >= 64) continue;
        const uint64_t below_hi = (hi >= 64) ? ~0ull : ((1ull << hi) - 1ull);
        const uint64_t below_lo = (lo <= 0) ? 0ull : ((1ull << lo) - 1ull);
        n += popc(d.w[w] & below_hi & ~below_lo);
    }
    return n;
}

struct Ham {
    int nel, nb, nw;
    const double *umat, *tmat;
    double ecore;
    // k-space Hubbard instead of tabulated integrals (ksum != NULL): <ij|kl> = U/N if k_i + k_j = k_k + k_l
    // (get_umat_kspace, UMAT(1) = UHUB/OMEGA), h_pp = eps(k_p); spatial orbital s has k index s - 1
    const int32_t *ksum = nullptr; int nk = 0; double u_over_n = 0.0; const double *eps_k = nullptr;
    // <ij|kl> over 1-based spatial orbitals, UMatInd (src/UMatCache.F90:257-296)
    inline double um(int i, int j, int k, int l) const {
        if (ksum) return (ksum[(i - 1) * nk + (j - 1)] == ksum[(k - 1) * nk + (l - 1)]) ? u_over_n : 0.0;
        return umat[tri(tri(i, k), tri(j, l)) - 1];
    }
    inline double tm(int a, int b) const {                                                // 0-based spin orbitals
        if (ksum) return (a == b) ? eps_k[a >> 1] : 0.0;
        return tmat[(size_t)a + (size_t)nb * b];
    }
    static inline int sp(int b) { return (b >> 1) + 1; }                                 // spatial index of bit b
    static inline bool same_spin(int a, int b) { return ((a ^ b) & 1) == 0; }

    // optional dense tables for sltcnd_1: cj[(i, a, j)] = <ij|aj>, ck[(i, a, j)] = <ij|ja> over spatial orbitals
    std::vector<double> cj, ck;
    void build_single_tables() {
        const int ns = nb / 2;
        cj.assign((size_t)ns * ns * ns, 0.0); ck.assign((size_t)ns * ns * ns, 0.0);
        for (int i = 1; i <= ns; ++i)
            for (int a = 1; a <= ns; ++a)
                for (int j = 1; j <= ns; ++j) {
                    const size_t k = ((size_t)(i - 1) * ns + (a - 1)) * ns + (j - 1);
                    cj[k] = um(i, j, a, j); ck[k] = um(i, j, j, a);
                }
    }

    inline Det2 load(const int64_t *il) const {
        Det2 d; d.w[0] = (uint64_t)il[0]; d.w[1] = (nw > 1) ? (uint64_t)il[1] : 0ull; return d;
    }

    // sltcnd_0, src/sltcnd.fpp:585-622
    double diag(const Det2 &I) const {
        int occ[128], n = 0;
        Det2 t = I;
        while (t.w[0] | t.w[1]) occ[n++] = pop_low(t);
        double h1 = 0.0, coul = 0.0, exch = 0.0;
        for (int a = 0; a < n; ++a) h1 += tm(occ[a], occ[a]);
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b) {
                const int i = sp(occ[a]), j = sp(occ[b]);
                coul += um(i, j, i, j);
                if (same_spin(occ[a], occ[b])) exch += um(i, j, j, i);
            }
        return h1 + coul - exch + ecore;
    }
    // sltcnd_1, src/sltcnd.fpp:643-678: I -> I - i + a
    double single(const Det2 &I, int i, int a) const {
        if (!same_spin(i, a)) return 0.0;
        const int si = sp(i), sa = sp(a);
        double h = 0.0;
        Det2 t = I;
        flip(t, i);
        if (!cj.empty()) {                       // same terms in the same order, looked up instead of indexed
            const int ns = nb / 2;
            const double *pj = &cj[((size_t)(si - 1) * ns + (sa - 1)) * ns], *pk = &ck[((size_t)(si - 1) * ns + (sa - 1)) * ns];
            while (t.w[0] | t.w[1]) {
                const int j = pop_low(t), sj = j >> 1;
                h += pj[sj];
                if (same_spin(i, j)) h -= pk[sj];
            }
        } else {
            while (t.w[0] | t.w[1]) {
                const int j = pop_low(t), sj = sp(j);
                h += um(si, sj, sa, sj);
                if (same_spin(i, j)) h -= um(si, sj, sj, sa);
            }
        }
        h += tm(i, a);
        return (between(I, i, a) & 1) ? -h : h;
    }
    // sltcnd_2, src/sltcnd.fpp:690-708: I -> I - i - j + a + b, paired (i -> a), (j -> b)
    double dbl(const Det2 &I, int i, int j, int a, int b) const {
        double h = 0.0;
        const int si = sp(i), sj = sp(j), sa = sp(a), sb = sp(b);
        if (same_spin(i, a) && same_spin(j, b)) h += um(si, sj, sa, sb);
        if (same_spin(i, b) && same_spin(j, a)) h -= um(si, sj, sb, sa);
        if (h == 0.0) return 0.0;
        Det2 t = I;
        int p = between(t, i, a);
        flip(t, i); flip(t, a);
        p += between(t, j, b);
        return (p & 1) ? -h : h;
    }
    // ---- HPHF functions, even S (src/HPHFIntegrals.fpp:62-150, 348-411) ----
    static inline Det2 spin_flip(const Det2 &d) {
        Det2 f;
        for (int w = 0; w < 2; ++w) f.w[w] = ((d.w[w] & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((d.w[w] & 0x5555555555555555ull) << 1);
        return f;
    }
    static inline bool closed_shell(const Det2 &d) { const Det2 f = spin_flip(d); return f.w[0] == d.w[0] && f.w[1] == d.w[1]; }
    // CalcOpenOrbs: beta electrons without their alpha partner
    static inline int open_orbs(const Det2 &d) {
        int n = 0;
        for (int w = 0; w < 2; ++w) n += popc(~((d.w[w] & 0xAAAAAAAAAAAAAAAAull) >> 1) & (d.w[w] & 0x5555555555555555ull));
        return n;
    }
    static inline int level(const Det2 &a, const Det2 &b) { return (popc(a.w[0] ^ b.w[0]) + popc(a.w[1] ^ b.w[1])) / 2; }
    // hphf_off_diag_helement_norm
    double hphf_off_diag(const Det2 &I, const Det2 &J) const {
        if (I.w[0] == J.w[0] && I.w[1] == J.w[1]) return 0.0;
        double hel = element(I, J);
        const bool ci = closed_shell(I), cj = closed_shell(J);
        if (ci != cj) return hel * std::sqrt(2.0);
        if (ci) return hel;
        const Det2 I2 = spin_flip(I);
        if (level(I2, J) <= 2) {
            const double m2 = element(I2, J);
            hel = (open_orbs(I) % 2 == 0) ? hel + m2 : hel - m2;
        }
        return hel;
    }
    // hphf_diag_helement
    double hphf_diag(const Det2 &I) const {
        double hel = diag(I);
        if (!closed_shell(I)) {
            const Det2 I2 = spin_flip(I);
            if (level(I, I2) <= 2) {
                const double m2 = element(I, I2);
                hel = (open_orbs(I) % 2 == 1) ? hel - m2 : hel + m2;
            }
        }
        return hel;
    }

    // get_helement(nI, nJ, iLutI, iLutJ)
    double element(const Det2 &I, const Det2 &J) const {
        Det2 hole, part;
        for (int w = 0; w < 2; ++w) { const uint64_t x = I.w[w] ^ J.w[w]; hole.w[w] = I.w[w] & x; part.w[w] = J.w[w] & x; }
        const int ic = popc(hole.w[0]) + popc(hole.w[1]);
        if (ic == 0) return diag(I);
        if (ic == 1) return single(I, low_bit(hole), low_bit(part));
        if (ic == 2) {
            const int i = pop_low(hole), j = low_bit(hole), a = pop_low(part), b = low_bit(part);
            return dbl(I, i, j, a, b);
        }
        return 0.0;
    }
};

struct CoreHamJob {
    int64_t n_local = 0, nnz = 0;
    // per row chunk: the non-zero entries of its rows, rows back to back
    static constexpr int64_t CHUNK = 32;
    std::vector<std::vector<int32_t>> col;
    std::vector<std::vector<double>> val;
    std::vector<int64_t> row_len;
};

}  // namespace

extern "C" {

// get_helement for n pairs of an FCIDUMP system (umat: packed UMAT, tmat: nbasis x nbasis TMAT2D column-major).
// iluts_*: n x nw occupation words.
int neci_host_get_helement(int32_t nel, int32_t nbasis, const double *umat, const double *tmat, double ecore,
                           const int64_t *iluts_i, const int64_t *iluts_j, int64_t n, double *out) {
    if (nbasis > 128) return 1;
    const Ham H{nel, nbasis, nbasis / 64 + 1, umat, tmat, ecore};
    for (int64_t k = 0; k < n; ++k) out[k] = H.element(H.load(iluts_i + k * H.nw), H.load(iluts_j + k * H.nw));
    return 0;
}
// the same between HPHF functions given by their allowed representatives (hphf_diag_helement when both are equal,
// hphf_off_diag_helement otherwise)
int neci_host_get_helement_hphf(int32_t nel, int32_t nbasis, const double *umat, const double *tmat, double ecore,
                                const int64_t *iluts_i, const int64_t *iluts_j, int64_t n, double *out) {
    if (nbasis > 128) return 1;
    const Ham H{nel, nbasis, nbasis / 64 + 1, umat, tmat, ecore};
    for (int64_t k = 0; k < n; ++k) {
        const Det2 I = H.load(iluts_i + k * H.nw), J = H.load(iluts_j + k * H.nw);
        out[k] = (I.w[0] == J.w[0] && I.w[1] == J.w[1]) ? H.hphf_diag(I) : H.hphf_off_diag(I, J);
    }
    return 0;
}

// generate_sing_doub_determinants: the reference determinant, then its spin- and symmetry-conserving single and
// double excitations, which is what GenExcitations3 enumerates.  orbsym: the FCIDUMP's ORBSYM label of every spatial
// orbital (abelian point groups: labels 1..8, product = ((a-1) xor (b-1)) + 1), or NULL when all irreps are equal as
// in the synthetic FCIDUMPs.  only_keep_conn additionally drops determinants without a matrix element to the
// reference (:585-594).  Returns the number of determinants written, or
// -(needed) if `capacity` is too small.  Order: ascending hole (pair), ascending particle (pair); the caller
// sorts per rank as init_semi_stochastic does (:227).
int64_t neci_host_sd_space(int32_t nel, int32_t nbasis, const double *umat, const double *tmat,
                           const int64_t *ilut_ref, int32_t only_keep_conn, const int32_t *orbsym,
                           int64_t capacity, int64_t *out) {
    if (nbasis > 128) return 0;
    const Ham H{nel, nbasis, nbasis / 64 + 1, umat, tmat, 0.0};
    const Det2 R = H.load(ilut_ref);
    std::vector<int> occ, vir;
    for (int b = 0; b < nbasis; ++b) (has(R, b) ? occ : vir).push_back(b);
    int64_t n = 0;
    auto put = [&](const Det2 &d) {
        if (n < capacity) { out[n * H.nw] = (int64_t)d.w[0]; if (H.nw > 1) out[n * H.nw + 1] = (int64_t)d.w[1]; }
        ++n;
    };
    put(R);
    auto irr = [&](int b) { return orbsym ? (orbsym[b >> 1] - 1) : 0; };
    for (int i : occ)
        for (int a : vir) {
            if (!Ham::same_spin(i, a) || irr(i) != irr(a)) continue;
            if (only_keep_conn && std::fabs(H.single(R, i, a)) < 1e-12) continue;
            Det2 d = R; flip(d, i); flip(d, a); put(d);
        }
    for (size_t x = 0; x < occ.size(); ++x)
        for (size_t y = x + 1; y < occ.size(); ++y)
            for (size_t p = 0; p < vir.size(); ++p)
                for (size_t q = p + 1; q < vir.size(); ++q) {
                    const int i = occ[x], j = occ[y], a = vir[p], b = vir[q];
                    if (((i & 1) + (j & 1)) != ((a & 1) + (b & 1))) continue;          // Ms conserved
                    if ((irr(i) ^ irr(j)) != (irr(a) ^ irr(b))) continue;              // total irrep conserved
                    if (only_keep_conn && std::fabs(H.dbl(R, i, j, a, b)) < 1e-12) continue;
                    Det2 d = R; flip(d, i); flip(d, j); flip(d, a); flip(d, b); put(d);
                }
    return (n <= capacity) ? n : -n;
}

// enumerate_sing_doub_kpnt (src/semi_stoch_gen.F90:1429-1500) for the k-space Hubbard model: the reference determinant
// and every double excitation conserving spin and total momentum (same-spin pairs included: they are symmetry-allowed
// although their matrix element to the reference vanishes); momentum conservation leaves no single excitations.
int64_t neci_host_sd_space_hubbard_k(int32_t nbasis, int32_t n_k, const int32_t *ksum, const int64_t *ilut_ref,
                                     int64_t capacity, int64_t *out) {
    if (nbasis > 128 || nbasis != 2 * n_k) return 0;
    const int nw = nbasis / 64 + 1;
    Det2 R; R.w[0] = (uint64_t)ilut_ref[0]; R.w[1] = (nw > 1) ? (uint64_t)ilut_ref[1] : 0ull;
    std::vector<int> occ, vir;
    for (int b = 0; b < nbasis; ++b) (has(R, b) ? occ : vir).push_back(b);
    int64_t n = 0;
    auto put = [&](const Det2 &d) {
        if (n < capacity) { out[n * nw] = (int64_t)d.w[0]; if (nw > 1) out[n * nw + 1] = (int64_t)d.w[1]; }
        ++n;
    };
    put(R);
    for (size_t x = 0; x < occ.size(); ++x)
        for (size_t y = x + 1; y < occ.size(); ++y)
            for (size_t p = 0; p < vir.size(); ++p)
                for (size_t q = p + 1; q < vir.size(); ++q) {
                    const int i = occ[x], j = occ[y], a = vir[p], b = vir[q];
                    if (((i & 1) + (j & 1)) != ((a & 1) + (b & 1))) continue;
                    if (ksum[(i >> 1) * n_k + (j >> 1)] != ksum[(a >> 1) * n_k + (b >> 1)]) continue;
                    Det2 d = R; flip(d, i); flip(d, j); flip(d, a); flip(d, b); put(d);
                }
    return (n <= capacity) ? n : -n;
}

// DetermineDetNode for n determinants (hash_iter = 0, no unique HF node: the defaults).
// blocks[k] = get_det_block (1-based, as neci_gpu_probe_det_node reports it), nodes[k] = LoadBalanceMapping(block).
int neci_host_det_node(int32_t nbasis, const int32_t *random_orb_index, int32_t balance_blocks,
                       const int32_t *load_balance_mapping, const int64_t *iluts, int64_t n,
                       int32_t *blocks, int32_t *nodes) {
    if (nbasis > 128 || balance_blocks <= 0) return 1;
    const int nw = nbasis / 64 + 1;
    for (int64_t k = 0; k < n; ++k) {
        Det2 d; d.w[0] = (uint64_t)iluts[k * nw]; d.w[1] = (nw > 1) ? (uint64_t)iluts[k * nw + 1] : 0ull;
        uint64_t acc = 0;                       // int64 arithmetic of the reference wraps the same way
        uint64_t i = 0;
        while (d.w[0] | d.w[1]) {
            const int b = pop_low(d);
            acc = 1099511628211ull * acc + (uint64_t)(int64_t)random_orb_index[b] * (++i);
        }
        const int64_t m = (int64_t)acc % (int64_t)balance_blocks;          // Fortran mod: sign of the dividend
        const int32_t blk = (int32_t)(m < 0 ? -m : m);
        if (blocks) blocks[k] = blk + 1;
        if (nodes) nodes[k] = load_balance_mapping[blk];
    }
    return 0;
}

// Sparse core Hamiltonian rows [displ, displ + n_local) over the whole core space `iluts` (n_core x nw, the
// rank-major order of store_whole_core_space).  Returns an opaque job (NULL on error) and its nnz; the rows are
// copied out and the job freed by neci_host_core_ham_fetch.  n_threads <= 0: all hardware threads.
static void *core_ham_job(Ham &H, double hii, const int64_t *iluts, int64_t n_core, int64_t displ, int64_t n_local,
                          int32_t n_threads, int32_t hphf, int64_t *nnz_out);

void *neci_host_core_ham_build(int32_t nel, int32_t nbasis, const double *umat, const double *tmat, double ecore,
                               double hii, const int64_t *iluts, int64_t n_core, int64_t displ, int64_t n_local,
                               int32_t n_threads, int32_t hphf, int64_t *nnz_out) {
    if (nbasis > 128 || displ < 0 || n_local < 0 || displ + n_local > n_core || n_core > 0x7fffffffll) return nullptr;
    Ham H{nel, nbasis, nbasis / 64 + 1, umat, tmat, ecore};
    return core_ham_job(H, hii, iluts, n_core, displ, n_local, n_threads, hphf, nnz_out);
}

// the same for the k-space Hubbard Hamiltonian (tables as neci_gpu_set_system_hubbard_k takes them)
void *neci_host_core_ham_build_hubbard_k(int32_t nel, int32_t nbasis, int32_t n_k, const int32_t *ksum, const double *eps_k,
                                         double u_over_n, double hii, const int64_t *iluts, int64_t n_core, int64_t displ,
                                         int64_t n_local, int32_t n_threads, int64_t *nnz_out) {
    if (nbasis > 128 || nbasis != 2 * n_k || displ < 0 || n_local < 0 || displ + n_local > n_core || n_core > 0x7fffffffll) return nullptr;
    Ham H{nel, nbasis, nbasis / 64 + 1, nullptr, nullptr, 0.0};
    H.ksum = ksum; H.nk = n_k; H.u_over_n = u_over_n; H.eps_k = eps_k;
    return core_ham_job(H, hii, iluts, n_core, displ, n_local, n_threads, 0, nnz_out);
}
int neci_host_get_helement_hubbard_k(int32_t nel, int32_t nbasis, int32_t n_k, const int32_t *ksum, const double *eps_k,
                                     double u_over_n, const int64_t *iluts_i, const int64_t *iluts_j, int64_t n, double *out) {
    if (nbasis > 128 || nbasis != 2 * n_k) return 1;
    Ham H{nel, nbasis, nbasis / 64 + 1, nullptr, nullptr, 0.0};
    H.ksum = ksum; H.nk = n_k; H.u_over_n = u_over_n; H.eps_k = eps_k;
    for (int64_t k = 0; k < n; ++k) out[k] = H.element(H.load(iluts_i + k * H.nw), H.load(iluts_j + k * H.nw));
    return 0;
}

static void *core_ham_job(Ham &H, double hii, const int64_t *iluts, int64_t n_core, int64_t displ, int64_t n_local,
                          int32_t n_threads, int32_t hphf, int64_t *nnz_out) {
    H.build_single_tables();
    std::vector<Det2> D((size_t)n_core);
    for (int64_t k = 0; k < n_core; ++k) D[k] = H.load(iluts + k * H.nw);
    auto *job = new CoreHamJob;
    job->n_local = n_local;
    const int64_t n_chunks = (n_local + CoreHamJob::CHUNK - 1) / CoreHamJob::CHUNK;
    job->col.resize(n_chunks); job->val.resize(n_chunks); job->row_len.assign(n_local, 0);
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        std::vector<int32_t> cc;               // scratch of this thread, reused from chunk to chunk: growing the
        std::vector<double> vv;                // chunk's own vectors page-faults under the process-wide mmap lock
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            cc.clear(); vv.clear();
            const int64_t r0 = c * CoreHamJob::CHUNK, r1 = std::min(n_local, r0 + CoreHamJob::CHUNK);
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t gi = displ + r;
                const Det2 I = D[gi];
                const Det2 I2 = hphf ? Ham::spin_flip(I) : I;
                const size_t start = cc.size();
                for (int64_t j = 0; j < n_core; ++j) {
                    if (j == gi) continue;
                    // more than a double excitation apart (for HPHF functions: from the determinant and from its partner)
                    if (Ham::level(I, D[j]) > 2 && (!hphf || Ham::level(I2, D[j]) > 2)) continue;
                    const double h = hphf ? H.hphf_off_diag(I, D[j]) : H.element(I, D[j]);
                    if (std::fabs(h) > 0.0) { cc.push_back((int32_t)j); vv.push_back(h); }
                }
                cc.push_back((int32_t)gi); vv.push_back((hphf ? H.hphf_diag(I) : H.diag(I)) - hii);  // the diagonal closes the row
                job->row_len[r] = (int64_t)(cc.size() - start);
            }
            job->col[c].assign(cc.begin(), cc.end());
            job->val[c].assign(vv.begin(), vv.end());
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n_chunks));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    job->nnz = 0;
    for (int64_t r = 0; r < n_local; ++r) job->nnz += job->row_len[r];
    if (nnz_out) *nnz_out = job->nnz;
    return job;
}

// row_ptr[n_local + 1] (0-based offsets), col[nnz] (0-based core-space index), val[nnz]; frees the job.
int neci_host_core_ham_fetch(void *handle, int64_t *row_ptr, int32_t *col, double *val) {
    auto *job = static_cast<CoreHamJob *>(handle);
    if (!job) return 1;
    if (row_ptr) {
        row_ptr[0] = 0;
        for (int64_t r = 0; r < job->n_local; ++r) row_ptr[r + 1] = row_ptr[r] + job->row_len[r];
        int64_t pos = 0;
        for (size_t c = 0; c < job->col.size(); ++c) {
            const size_t m = job->col[c].size();
            if (m) {
                std::memcpy(col + pos, job->col[c].data(), m * sizeof(int32_t));
                std::memcpy(val + pos, job->val[c].data(), m * sizeof(double));
            }
            pos += (int64_t)m;
            std::vector<int32_t>().swap(job->col[c]);
            std::vector<double>().swap(job->val[c]);
        }
    }
    delete job;
    return 0;
}

// out_i = sum_j <row_i|H|col_j> vec_j: the connected-space vector of a trial wavefunction, con_space_vecs
// (generate_connected_space_vector, src/trial_wf_gen.F90) with rows = connected space, cols = trial space.
int neci_host_ham_apply(int32_t nel, int32_t nbasis, const double *umat, const double *tmat, double ecore,
                        const int64_t *rows, int64_t n_rows, const int64_t *cols, int64_t n_cols,
                        const double *vec, int32_t n_threads, double *out) {
    if (nbasis > 128) return 1;
    Ham H{nel, nbasis, nbasis / 64 + 1, umat, tmat, ecore};
    H.build_single_tables();
    std::vector<Det2> Cc((size_t)n_cols);
    for (int64_t k = 0; k < n_cols; ++k) Cc[k] = H.load(cols + k * H.nw);
    std::atomic<int64_t> next{0};
    const int64_t CH = 256;
    auto work = [&]() {
        for (;;) {
            const int64_t r0 = next.fetch_add(CH);
            if (r0 >= n_rows) break;
            for (int64_t r = r0; r < std::min(n_rows, r0 + CH); ++r) {
                const Det2 I = H.load(rows + r * H.nw);
                double acc = 0.0;
                for (int64_t j = 0; j < n_cols; ++j) {
                    if (popc(I.w[0] ^ Cc[j].w[0]) + popc(I.w[1] ^ Cc[j].w[1]) > 4) continue;
                    acc += H.element(I, Cc[j]) * vec[j];
                }
                out[r] = acc;
            }
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n_rows + CH - 1) / CH));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return 0;
}

}  // extern "C"
