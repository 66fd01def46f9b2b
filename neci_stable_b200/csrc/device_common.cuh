// Device-side building blocks of the FCIQMC engine: determinant bit-string
// arithmetic (popc / ffs based, no orbital lists), the counter-based Philox
// streams, and the parameter block every kernel receives by value.
//
// Orbital numbering follows NECI: 1-based spin orbitals, odd = beta,
// even = alpha (src/macros.h:16,21); orbital o is bit (o-1)%64 of word (o-1)/64
// (src/BitReps.F90:164-306).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/neci_gpu.h"

namespace ng {

typedef unsigned long long u64;
typedef unsigned int u32;

#define NG_MAX_BASIS 128
// Kernel variant selector: the three system types of the ABI plus the HPHF flavour of the FCIDUMP/PCHB system.  HPHF
// is a compile-time variant because its code inside the spawning kernel costs the determinant runs 20 % when it is
// merely present behind a run-time flag (registers and instruction cache).
#define NG_SYS_PCHB_HPHF 4
// the FCIDUMP/PCHB system with PCHB_ParticleSelection FULL-FULL (neci_gpu_set_pchb_particles): its own variant for the
// same reason
#define NG_SYS_PCHB_FULL 5
__host__ __device__ constexpr bool sys_pchb(int s) { return s == NECI_SYS_FCIDUMP_PCHB || s == NG_SYS_PCHB_HPHF || s == NG_SYS_PCHB_FULL; }
__host__ __device__ constexpr bool sys_hphf(int s) { return s == NG_SYS_PCHB_HPHF; }
#define NG_MAX_CLASSES 16

// ---------------------------------------------------------------------------
// Determinant occupation bits in registers.  NW = nIfD + 1 = 1 or 2 words.
// ---------------------------------------------------------------------------
template <int NW> struct Det { u64 w[NW]; };

#define NG_BETA_MASK  0x5555555555555555ull   /* odd orbitals  -> even bit index */
#define NG_ALPHA_MASK 0xAAAAAAAAAAAAAAAAull   /* even orbitals -> odd bit index  */

template <int NW> __device__ __forceinline__ bool det_eq(const Det<NW> &a, const Det<NW> &b) {
    bool e = a.w[0] == b.w[0];
    if (NW > 1) e = e && (a.w[1] == b.w[1]);
    return e;
}
// Two-word determinants are addressed with selects, never with a run-time word index: d.w[b >> 6] makes the
// compiler either spill the determinant to local memory or branch, and lanes of one warp routinely sit in different words.
template <int NW> __device__ __forceinline__ bool occ(const Det<NW> &d, int orb) {
    const int b = orb - 1;
    if (NW == 1) return (d.w[0] >> b) & 1ull;
    const u64 w = (b < 64) ? d.w[0] : d.w[NW - 1];
    return (w >> (b & 63)) & 1ull;
}
template <int NW> __device__ __forceinline__ void set_orb(Det<NW> &d, int orb) {
    const int b = orb - 1;
    if (NW == 1) { d.w[0] |= 1ull << b; return; }
    const u64 bit = 1ull << (b & 63);
    d.w[0] |= (b < 64) ? bit : 0ull;
    d.w[NW - 1] |= (b < 64) ? 0ull : bit;
}
template <int NW> __device__ __forceinline__ void clr_orb(Det<NW> &d, int orb) {
    const int b = orb - 1;
    if (NW == 1) { d.w[0] &= ~(1ull << b); return; }
    const u64 bit = 1ull << (b & 63);
    d.w[0] &= ~((b < 64) ? bit : 0ull);
    d.w[NW - 1] &= ~((b < 64) ? 0ull : bit);
}
template <int NW> __device__ __forceinline__ int popc(const Det<NW> &d) {
    int n = __popcll(d.w[0]);
    if (NW > 1) n += __popcll(d.w[1]);
    return n;
}
// FindBitExcitLevel (src/DetBitOps.F90:140-180): popcount(ref & (ref ^ det))
template <int NW> __device__ __forceinline__ int excit_level(const Det<NW> &ref, const Det<NW> &d) {
    int n = __popcll(ref.w[0] & (ref.w[0] ^ d.w[0]));
    if (NW > 1) n += __popcll(ref.w[1] & (ref.w[1] ^ d.w[1]));
    return n;
}
// bit index (0-based) of the n-th (1-based) set bit of a 64-bit word; n <= popc(x).
// Branch-free popcount bisection (the __fns intrinsic is a divergent software loop).
__device__ __forceinline__ int select64(u64 x, int n) {
    u32 w = (u32)x;
    int pos = 0;
    int c = __popc(w);
    if (n > c) { n -= c; w = (u32)(x >> 32); pos = 32; }
    c = __popc(w & 0xFFFFu); if (n > c) { n -= c; w >>= 16; pos += 16; }
    c = __popc(w & 0xFFu);   if (n > c) { n -= c; w >>= 8;  pos += 8; }
    c = __popc(w & 0xFu);    if (n > c) { n -= c; w >>= 4;  pos += 4; }
    c = __popc(w & 0x3u);    if (n > c) { n -= c; w >>= 2;  pos += 2; }
    if (n > (int)(w & 1u)) pos += 1;
    return pos;
}
// orbital (1-based) holding the n-th set bit of det & mask
template <int NW> __device__ __forceinline__ int select_orb(const Det<NW> &d, u64 mask, int n) {
    const u64 a = d.w[0] & mask;
    if (NW == 1) return select64(a, n) + 1;
    const int c = __popcll(a);
    const bool first = n <= c;                             // one select64 for both words
    return (first ? 0 : 64) + select64(first ? a : (d.w[NW - 1] & mask), first ? n : n - c) + 1;
}
// number of occupied orbitals with index < orb  (== position in nI, 0-based)
template <int NW> __device__ __forceinline__ int count_below(const Det<NW> &d, int orb) {
    const int b = orb - 1;
    if (NW == 1) return __popcll(d.w[0] & ((1ull << b) - 1ull));
    const u64 low = (1ull << (b & 63)) - 1ull;             // bits below b within its word
    const u64 m0 = (b < 64) ? low : ~0ull;
    const u64 m1 = (b < 64) ? 0ull : low;
    return __popcll(d.w[0] & m0) + __popcll(d.w[NW - 1] & m1);
}
// occupied orbitals strictly between orbitals a and b (a != b)
template <int NW> __device__ __forceinline__ int count_between(const Det<NW> &d, int a, int b) {
    const int lo = min(a, b), hi = max(a, b);
    // below(hi) counts orbitals < hi, below(lo+1) counts orbitals <= lo
    return count_below(d, hi) - count_below(d, lo + 1);
}
// iterate set bits in ascending orbital order: returns next orbital and clears it
template <int NW> __device__ __forceinline__ int pop_lowest(Det<NW> &d) {
    if (NW == 1) { const int b = __ffsll((long long)d.w[0]) - 1; d.w[0] &= d.w[0] - 1ull; return b + 1; }
    const bool first = d.w[0] != 0ull;                     // branch-free: lanes of a warp are in different words
    u64 x = first ? d.w[0] : d.w[NW - 1];
    const int b = __ffsll((long long)x) - 1;
    x &= x - 1ull;
    d.w[0] = first ? x : 0ull;
    d.w[NW - 1] = first ? d.w[NW - 1] : x;
    return (first ? 0 : 64) + b + 1;
}
template <int NW> __device__ __forceinline__ bool det_any(const Det<NW> &d) {
    u64 x = d.w[0]; if (NW > 1) x |= d.w[1]; return x != 0ull;
}

__device__ __forceinline__ bool is_beta(int orb) { return orb & 1; }
__device__ __forceinline__ int gtid(int orb) { return (orb + 1) >> 1; }     // spatial index, 1-based
__device__ __forceinline__ int fuse_index(int x, int y) {                  // src/lib/util_mod.fpp:429-441
    return (x < y) ? x + y * (y - 1) / 2 : y + x * (x - 1) / 2;
}
__device__ __forceinline__ double dsign(double a, double b) { return copysign(a, b); }

// ---------------------------------------------------------------------------
// Philox4x32, counter based; stream keyed by (seed, iteration, determinant hash, attempt, purpose) -- DESIGN.md
// §RNG.  A stream is a sequence of 32-bit words, four per Philox block; draw53() takes two consecutive words and
// maps them to [0,1) with 53 bits, draw32() takes one word (coarse choices: an electron, an orbital of a class).
// Seven rounds: the smallest count for which Philox4x32 passes BigCrush (Salmon et al., SC'11, table 2; ten is that
// paper's default with a safety margin) -- the engine draws ~4e10 blocks per second, so the rounds are paid for.
// The ten-round function is kept for the Random123 known-answer test.
// ---------------------------------------------------------------------------
enum : u32 { RNG_NSPAWN = 0, RNG_ATTEMPT = 1, RNG_DEATH = 2, RNG_ROUND_SPAWN = 3, RNG_PRUNE = 4, RNG_ATT_ROUND = 5 };
#define NG_PHILOX_ROUNDS 7

__host__ __device__ __forceinline__ u64 mix64(u64 z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
template <int NW> __device__ __forceinline__ u64 det_hash64(const Det<NW> &d) {
    u64 h = mix64(d.w[0] + 0x9E3779B97F4A7C15ull);
    if (NW > 1) h = mix64((h ^ (d.w[NW - 1] * 0xC2B2AE3D27D4EB4Full)) + 0x165667B19E3779F9ull);
    return h;
}

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_rounds(u32 v0, u32 v1, u32 v2, u32 v3, u32 q0, u32 q1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (r) { q0 += 0x9E3779B9u; q1 += 0xBB67AE85u; }
        const u32 hi0 = __umulhi(0xD2511F53u, v0), lo0 = 0xD2511F53u * v0;
        const u32 hi1 = __umulhi(0xCD9E8D57u, v2), lo1 = 0xCD9E8D57u * v2;
        const u32 n0 = hi1 ^ v1 ^ q0, n2 = hi0 ^ v3 ^ q1;
        v0 = n0; v1 = lo1; v2 = n2; v3 = lo0;
    }
    return make_uint4(v0, v1, v2, v3);
}
// One block of the engine's generator.  Deliberately NOT inlined: a stream is drawn from at a dozen places of the
// spawning kernel and a dozen inlined copies pushed its code past the instruction cache.
__device__ __noinline__ uint4 philox_block(u32 v0, u32 v1, u32 v2, u32 v3, u32 q0, u32 q1) {
    return philox4x32_rounds<NG_PHILOX_ROUNDS>(v0, v1, v2, v3, q0, q1);
}

struct Stream {
    u32 c0, c1, c2, c3, k0, k1;
    uint4 blk;         // block `cur` when cur >= 0
    int cur, pos;      // pos: index of the next word
    __device__ __forceinline__ Stream(u64 seed, long long iter, u64 h, u32 attempt, u32 purpose, int start_word = 0) {
        c0 = (u32)h; c1 = (u32)(h >> 32); c2 = attempt; c3 = purpose << 24;
        k0 = (u32)seed ^ (u32)(seed >> 32); k1 = (u32)iter; pos = start_word; cur = -1; blk = make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ u32 next_u32() {
        const int b = pos >> 2, w = pos & 3;
        if (b != cur) { blk = philox_block(c0, c1, c2, c3 | (u32)b, k0, k1); cur = b; }
        ++pos;
        return (w == 0) ? blk.x : (w == 1) ? blk.y : (w == 2) ? blk.z : blk.w;
    }
    __device__ __forceinline__ double draw32() { return (double)next_u32() * (1.0 / 4294967296.0); }
    __device__ __forceinline__ double draw53() {
        const u32 a = next_u32(), b = next_u32();
        const u64 u = (u64)a | ((u64)b << 32);
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double draw() { return draw53(); }
};

// ---------------------------------------------------------------------------
// Parameter block (passed by value; lives in the kernel's constant bank).
// ---------------------------------------------------------------------------
// one entry of an alias sampler, 32 bytes = one L2 sector: bias threshold and probability of pair index ab with its
// two spatial target orbitals (tgtOrbs(:, ab)) packed as lo | hi << 16, and -- instead of the alias INDEX the
// reference's tables hold -- the alias's own probability and target orbitals, so that taking the alias costs no
// second (dependent) table load
struct __align__(32) PchbEntry { double bias, prob, prob_alias; u32 tgt, tgt_alias; };
// per electron-pair index ij: exchange probability, and bit s set when sampler s of this pair is non-empty
struct __align__(16) PchbPair { double p_exch; int nonempty; int pad; };

struct Params {
    // configuration scalars
    int nel, nbasis, nocc_alpha, nocc_beta;
    int nranks, rank, balance_blocks;
    u64 bb_magic;                     // floor((2^64 - 1) / balance_blocks), see det_block
    int system_type;
    int t_trunc_initiator, t_all_real_coeff, t_real_spawn_cutoff, t_death_before_comms;
    int t_init_coherent_rule, t_no_brillouin, t_exch, t_semi_stochastic, t_core_inits;
    int t_tau_search, t_consider_par_bias, t_hphf;
    double initiator_walk_no, real_spawn_cutoff, occupied_thresh, av_mc_excits;
    double hii, ecore;
    u64 seed;
    u64 ref[2];
    // hashing tables (global memory copies)
    const int *random_orb_index;      // [nbasis]
    const int *lb_mapping;            // [balance_blocks]
    // FCIDUMP
    const double *umat, *tmat;
    const double *jmat, *kmat;        // <ij|ij>, <ij|ji> over spatial orbitals, [n_spat_sys][n_spat_sys] (sltcnd_0)
    int n_spat_sys;
    // PCHB: the host's probs / bias / alias / tgtOrbs arrays interleaved into one 32-byte entry per
    // (ij, sampler, ab) so that an alias draw costs one L2 sector (two when the alias is taken)
    int n_spat, ij_max, ab_max;
    const struct PchbEntry *pchb;     // [ij_max * 3 * ab_max]
    const struct PchbPair *pchb_pair; // [ij_max]
    const double *pchb_pfirst, *pchb_psecond;       // FULL-FULL / UNIF-FULL particle selection: p_first[nbasis], p_second[nbasis][nbasis]
    int pchb_particles;                             // 0 UNIF-UNIF, 1 FULL-FULL, 2 UNIF-FULL
    double p_singles, p_doubles, p_parallel;
    double pgen_pair_par, pgen_pair_opp;            // p_parallel / #parallel pairs, (1 - p_parallel) / #alpha-beta pairs
    // host-computed rescaling constants of the first draw of an attempt (see gen_pchb_double): 1 / (1 - p_singles),
    // #parallel pairs / p_parallel, #alpha-beta pairs / (1 - p_parallel)
    double inv_1m_ps, c_par, c_opp;
    const unsigned short *tri_tab;                  // pair index -> n1 | n2 << 8 of a same-spin electron pair
    u32 magic_nalpha;                               // floor(2^32 / nocc_alpha) + 1 (exact quotients for idx < 2^32 / nocc_alpha)
    int n_classes;
    const unsigned char *class_of_spinorb;          // [nbasis]
    const int *class_start, *class_orbs;            // CSR of class members
    u64 class_mask[NG_MAX_CLASSES][2];
    // r-space Hubbard
    int max_neigh;
    const int *neighbours;
    double uhub;
    // k-space Hubbard
    int n_k;
    const int *ksum, *kdiff;
    const double *eps_k;
    const u64 *kperm;                 // [n_k][kperm_bytes][256]: image of a byte of a k-point mask under k -> kij - k
    int kperm_bytes;
    const double *kcum;               // [nbasis + 1] running sums k * |U/N| accumulated as the reference does (gen_k_hubbard)
    double u_over_n;
    // semi-stochastic: the whole core space, replicated (is_core_state, src/semi_stoch_procs.F90:547)
    const long long *core_iluts;      // [n_core_total][nw]
    const int *core_ht;               // open addressing, entry = index + 1, 0 = empty
    u64 core_ht_mask;
    // trial wavefunction: trial space followed by connected space (trial_ht / con_ht of src/searching.F90:182-223)
    const long long *trial_iluts;     // [n_trial + n_con][nw]
    const double *trial_amps;         // trial_wfs / con_space_vecs
    const int *trial_ht;              // entry = index + 1, 0 = empty
    u64 trial_ht_mask;
    long long n_trial;
};

// main walker list, structure of arrays in HBM
struct WalkerList {
    u64 *det0, *det1;        // occupation words (det1 unused for NW = 1)
    double *sgn;
    int *flg;
    double *diagH, *offH;
    double *trial_amp;       // current_trial_amps(1, :), allocated by neci_gpu_set_trial_space
    long long cap;
    // open-addressing hash table: entry = (tag32 << 32) | slot ; EMPTY / TOMB sentinels
    u64 *ht; u64 ht_mask;
    // free-slot stacks: A is popped, B is pushed (merged between kernels)
    int *freeA, *freeB;
    // device counters: [0] n_list, [1] n_freeA, [2] n_freeB, [3] n_tomb, [4] n_heavy,
    //                  [5] n_insert, [6] err flags, [7] n_merged
    long long *ctr;
    // Host mirror (neci_gpu_iterate_host with page-locked host arrays): device pointers of the HOST's CurrentDets and
    // global_determinant_data rows.  While set, every kernel that changes a slot also writes the change through to
    // the host arrays over PCIe, so the host list is current when the iteration ends without a bulk download of
    // the ~85 % of the slots an iteration does not touch.  Null otherwise.
    long long *h_rec; double *h_gd, *h_go; int h_W;
};
// write-through of a slot's sign / flag words (and, for a new determinant, of the whole record) to the host mirror
template <int NW> __device__ __forceinline__ void mirror_sign(const WalkerList &L, long long slot, double s) {
    if (L.h_rec) L.h_rec[(size_t)slot * L.h_W + NW] = __double_as_longlong(s);
}
template <int NW> __device__ __forceinline__ void mirror_flags(const WalkerList &L, long long slot, int f) {
    if (L.h_rec) L.h_rec[(size_t)slot * L.h_W + NW + 1] = (long long)f;
}
template <int NW> __device__ __forceinline__ void mirror_record(const WalkerList &L, long long slot, const Det<NW> &d, double s, int f,
                                                               double hd, double ho) {
    if (!L.h_rec) return;
    long long *rec = L.h_rec + (size_t)slot * L.h_W;
    rec[0] = (long long)d.w[0]; if (NW > 1) rec[NW - 1] = (long long)d.w[NW - 1];
    rec[NW] = __double_as_longlong(s); rec[NW + 1] = (long long)f;
    if (L.h_gd) L.h_gd[slot] = hd;
    if (L.h_go) L.h_go[slot] = ho;
}
enum { C_NLIST = 0, C_NFREEA, C_NFREEB, C_NTOMB, C_NHEAVY, C_NINSERT, C_ERR, C_NMERGED, C_COUNT = 16 };

#define HT_EMPTY 0xFFFFFFFFFFFFFFFFull
#define HT_TOMB  0xFFFFFFFFFFFFFFFEull

#define F_INIT   (1 << NECI_FLAG_INITIATOR)
#define F_DETERM (1 << NECI_FLAG_DETERMINISTIC)
#define F_REMOVED (1 << NECI_FLAG_REMOVED)
#define F_DPARENT (1 << NECI_FLAG_DETERM_PARENT)
#define F_TRIAL (1 << NECI_FLAG_TRIAL)
#define F_CONNECTED (1 << NECI_FLAG_CONNECTED)
// engine-internal marker bits inside spawn-record flag words (never leave the device)
#define SF_MULTI (1ll << 40)   /* record merged from >= 2 spawns */
#define SF_DEAD  (1ll << 41)   /* record folded into its representative */

template <int NW> __device__ __forceinline__ Det<NW> load_det(const WalkerList &L, long long i) {
    Det<NW> d; d.w[0] = L.det0[i]; if (NW > 1) d.w[NW - 1] = L.det1[i]; return d;
}
template <int NW> __device__ __forceinline__ void store_det(const WalkerList &L, long long i, const Det<NW> &d) {
    L.det0[i] = d.w[0]; if (NW > 1) L.det1[i] = d.w[NW - 1];
}
template <int NW> __device__ __forceinline__ Det<NW> ref_det(const Params &P) {
    Det<NW> d; d.w[0] = P.ref[0]; if (NW > 1) d.w[NW - 1] = P.ref[1]; return d;
}

// hash-table probe: returns slot or -1.  If ht_pos is given it receives the
// table position of the hit (for tombstoning) or of the terminating EMPTY.
template <int NW>
__device__ __forceinline__ long long ht_lookup(const WalkerList &L, const Det<NW> &d, u64 h, u64 *ht_pos = nullptr,
                                               long long *first_tomb = nullptr) {
    const u32 tag = (u32)(h >> 32);
    u64 pos = h & L.ht_mask;
    if (first_tomb) *first_tomb = -1;
    for (;;) {
        const u64 e = L.ht[pos];
        if (e == HT_EMPTY) { if (ht_pos) *ht_pos = pos; return -1; }
        if (e == HT_TOMB) { if (first_tomb && *first_tomb < 0) *first_tomb = (long long)pos; }
        else if ((u32)(e >> 32) == tag) {
            const long long slot = (long long)(e & 0xFFFFFFFFull);
            if (det_eq(load_det<NW>(L, slot), d)) { if (ht_pos) *ht_pos = pos; return slot; }
        }
        pos = (pos + 1) & L.ht_mask;
    }
}
// is_core_state: membership in the replicated core space
template <int NW>
__device__ __forceinline__ bool is_core_state(const Params &P, const Det<NW> &d) {
    if (!P.core_ht) return false;
    u64 pos = det_hash64(d) & P.core_ht_mask;
    for (;;) {
        const int e = __ldg(&P.core_ht[pos]);
        if (e == 0) return false;
        const long long *il = P.core_iluts + (size_t)(e - 1) * NW;
        bool same = ((u64)__ldg(&il[0]) == d.w[0]);
        if (NW > 1) same = same && ((u64)__ldg(&il[NW - 1]) == d.w[NW - 1]);
        if (same) return true;
        pos = (pos + 1) & P.core_ht_mask;
    }
}
// hash_search_trial (src/searching.F90:182-223): flag bit (trial / connected / none) and amplitude of a determinant
template <int NW>
__device__ __forceinline__ int trial_lookup(const Params &P, const Det<NW> &d, double *amp) {
    *amp = 0.0;
    if (!P.trial_ht) return 0;
    u64 pos = det_hash64(d) & P.trial_ht_mask;
    for (;;) {
        const int e = __ldg(&P.trial_ht[pos]);
        if (e == 0) return 0;
        const long long *il = P.trial_iluts + (size_t)(e - 1) * NW;
        bool same = ((u64)__ldg(&il[0]) == d.w[0]);
        if (NW > 1) same = same && ((u64)__ldg(&il[NW - 1]) == d.w[NW - 1]);
        if (same) { *amp = __ldg(&P.trial_amps[e - 1]); return (e - 1 < P.n_trial) ? F_TRIAL : F_CONNECTED; }
        pos = (pos + 1) & P.trial_ht_mask;
    }
}
// insert a key known to be absent (unique among concurrent inserters).  Returns true when a tombstone was recycled:
// the caller counts those and lowers L.ctr[C_NTOMB] once per warp (a per-insert atomic on one address serialises).
__device__ __forceinline__ bool ht_insert(const WalkerList &L, u64 h, long long slot, u64 start_pos) {
    const u64 entry = ((u64)(u32)(h >> 32) << 32) | (u64)(u32)slot;
    u64 pos = start_pos;
    u64 e = __ldcg(&L.ht[pos]);                 // L2 reads: other CTAs insert concurrently
    for (;;) {
        if (e == HT_EMPTY || e == HT_TOMB) {
            const u64 old = atomicCAS((unsigned long long *)&L.ht[pos], e, entry);
            if (old == e) return e == HT_TOMB;
            e = old;                            // lost the race: judge the winner's value, no re-read
            continue;
        }
        pos = (pos + 1) & L.ht_mask;
        e = __ldcg(&L.ht[pos]);
    }
}
// RemoveHashDet, table part only: tombstones the entry; returns true when an entry was tombstoned.  The caller
// pushes the slot on the free stack and counts the tombstone (k_walk aggregates both per CTA).
template <int NW>
__device__ __forceinline__ bool ht_tombstone(const WalkerList &L, const Det<NW> &d, u64 h, long long slot) {
    u64 pos;
    const long long s = ht_lookup<NW>(L, d, h, &pos);
    if (s == slot) { L.ht[pos] = HT_TOMB; return true; }
    return false;
}
// tombstones recycled by the inserts of this thread -> L.ctr[C_NTOMB], one atomic per warp (call with all lanes)
__device__ __forceinline__ void ht_settle_tombs(const WalkerList &L, int delta) {
    const int t = __reduce_add_sync(0xffffffffu, delta);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd((unsigned long long *)&L.ctr[C_NTOMB], (unsigned long long)(long long)t);
}
// RemoveHashDet (src/load_balancer.fpp:631-644): tombstone + push the slot on the free stack
template <int NW>
__device__ __forceinline__ void ht_remove(const WalkerList &L, const Det<NW> &d, u64 h, long long slot) {
    u64 pos;
    const long long s = ht_lookup<NW>(L, d, h, &pos);
    if (s == slot) {
        L.ht[pos] = HT_TOMB;
        atomicAdd((unsigned long long *)&L.ctr[C_NTOMB], 1ull);
    }
    const long long k = (long long)atomicAdd((unsigned long long *)&L.ctr[C_NFREEB], 1ull);
    L.freeB[k] = (int)slot;
}

}  // namespace ng
