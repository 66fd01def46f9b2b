// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement ("oracle") of the FCIQMC walker-propagation hot path of
// ghb24/NECI_STABLE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (neci_stable_b200/) never does.
//
// Parity pinning: the reference (Fortran + MPI) cannot be built in this image
// (no gfortran/MPI), so the oracle is pinned against the reference's own
// known-answer tests, fixtures and printed run outputs instead (DESIGN.md section 2 has
// the whole table; tests/test_oracle_golden.py, test_core_space_cpu.py, test_hphf_cpu.py, ...):
//   - Hubbard matrix-element and generator known answers
//     (unit_tests/real_space_hubbard/test_real_space_hubbard.F90, unit_tests/k_space_hubbard/test_k_space_hubbard.F90)
//   - sltcnd property test  (unit_tests/sltcnd/test_sltcnd.F90:25-67)
//   - alias-table L1 test   (unit_tests/sampler/test_aliasTables.F90:45-110)
//   - PCHB sum(1/pgen) test (unit_tests/excitgen/pchb_excitgen_test_helper.F90:40-118), for the UNIF-UNIF,
//     FULL-FULL and UNIF-FULL particle selections
//   - DetermineDetNode: the `Reference processor` 58 reference runs printed, with the reference's own dSFMT
//     (oracle/_ref, compiled in place from src/lib/dSFMT.cpp)
//   - reference energies, deterministic-space sizes and core correlation energies printed by the reference's
//     regression runs; determ_projection and SumEContrib by the first lines of the determ_doubles iteration table
//   - exact diagonalisation energies of small lattices.
// Parity UNPINNED by any reference vector (restatement reviewed against the cited lines, and checked against a
// second restatement in the reference's sort-and-merge form, tests/literal_annihilation.py): FindWalkerHash values,
// CompressSpawnedList / AnnihilateSpawnedParts outputs on a fixed list, CalcHashTableStats.  The reference holds no
// test for those.
//
// The random stream is NOT the reference's dSFMT: both the oracle and the CUDA
// engine draw from the same counter-based Philox4x32 streams (seven rounds) keyed by
// (seed, iteration, determinant, attempt, purpose) -- see DESIGN.md section 3 -- so
// a whole iteration of the engine can be compared with the oracle bit for bit.  Where the engine takes a choice
// from fewer random numbers than the reference (PCHB: single/double, electron pair and exchange from one number;
// k-space Hubbard: the opposite-spin electron pair from one number instead of a rejection loop; weighted particle
// selection: the CDF branch of constrained_sample instead of redrawing), the oracle draws the same way; each such
// place says so and keeps the reference's probabilities, which the acceptance tests above check.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <array>
#include <string>
#include <unordered_map>
#include <algorithm>
#include "../include/neci_gpu.h"

namespace orc {

// ----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: no state.
// ----------------------------------------------------------------------------
struct Philox {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k[0];
        const uint32_t n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    // rounds = 10 is the Random123 default (known-answer test); the streams of the engine and of this oracle use
    // ROUNDS = 7, the smallest count that passes BigCrush (Salmon et al., SC'11)
    static constexpr int ROUNDS = 7;
    static inline void gen(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds = 10) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int r = 0; r < rounds; ++r) {
            if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
            round(c, k);
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// purposes of a random stream
enum : uint32_t { RNG_NSPAWN = 0, RNG_ATTEMPT = 1, RNG_DEATH = 2, RNG_ROUND_SPAWN = 3, RNG_PRUNE = 4, RNG_ATT_ROUND = 5 };

inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
// 64-bit identity of a determinant used to key its random streams.
inline uint64_t det_hash64(const uint64_t *w, int nwords) {
    uint64_t h = mix64(w[0] + 0x9E3779B97F4A7C15ull);
    if (nwords > 1) h = mix64((h ^ (w[1] * 0xC2B2AE3D27D4EB4Full)) + 0x165667B19E3779F9ull);
    return h;
}

// One (determinant, attempt, purpose, iteration) stream: a sequence of 32-bit words, four per Philox block.
// draw53() takes two consecutive words and maps them to [0,1) with 53 bits; draw32() takes one word (coarse
// choices: an electron, an orbital of a class).  draw() == draw53().
struct Stream {
    uint32_t ctr[4], key[2];
    uint32_t cache[4];
    int cur = -1, pos = 0;
    Stream(uint64_t seed, int64_t iter, uint64_t h, uint32_t attempt, uint32_t purpose, int start_word = 0) {
        ctr[0] = (uint32_t)h; ctr[1] = (uint32_t)(h >> 32); ctr[2] = attempt; ctr[3] = purpose << 24;
        key[0] = (uint32_t)seed ^ (uint32_t)(seed >> 32); key[1] = (uint32_t)iter;
        pos = start_word;
    }
    uint32_t next_u32() {
        const int b = pos >> 2;
        if (b != cur) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3] | (uint32_t)b};
            Philox::gen(c, key, cache, Philox::ROUNDS);
            cur = b;
        }
        return cache[(pos++) & 3];
    }
    double draw32() { return (double)next_u32() * (1.0 / 4294967296.0); }
    double draw53() {
        const uint32_t a = next_u32(), b = next_u32();
        const uint64_t u = (uint64_t)a | ((uint64_t)b << 32);
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    double draw() { return draw53(); }
};

// ----------------------------------------------------------------------------
// basic orbital helpers (src/macros.h:16-30)
// ----------------------------------------------------------------------------
inline bool is_beta(int orb) { return (orb & 1) == 1; }   // odd = beta
inline bool is_alpha(int orb) { return (orb & 1) == 0; }
inline int  gtID(int orb) { return (orb - 1) / 2 + 1; }    // spatial index, 1-based
inline int  G1_Ms(int orb) { return is_alpha(orb) ? 1 : -1; }
inline bool is_occ(const uint64_t *ilut, int orb) { return (ilut[(orb - 1) / 64] >> ((orb - 1) % 64)) & 1ull; }
inline void set_orb(uint64_t *ilut, int orb) { ilut[(orb - 1) / 64] |= (1ull << ((orb - 1) % 64)); }
inline void clr_orb(uint64_t *ilut, int orb) { ilut[(orb - 1) / 64] &= ~(1ull << ((orb - 1) % 64)); }
inline int64_t fuseIndex(int64_t x, int64_t y) {          // src/lib/util_mod.fpp:429-441
    return (x < y) ? x + y * (y - 1) / 2 : y + x * (x - 1) / 2;
}
inline double dsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
inline double sign_to_double(int64_t w) { double d; std::memcpy(&d, &w, 8); return d; }
inline int64_t double_to_sign(double d) { int64_t w; std::memcpy(&w, &d, 8); return w; }

constexpr double EPS = 1e-13;                              // src/lib/constants.F90:28
inline bool near_zero(double x) { return std::fabs(x) <= EPS; }
inline bool unocc(double s) { return std::fabs(s) < 1.0e-12; }   // IsUnoccDet, src/macros.h:18

}  // namespace orc
