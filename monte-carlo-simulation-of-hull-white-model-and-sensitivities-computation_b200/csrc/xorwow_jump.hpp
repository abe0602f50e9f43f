// xorwow_jump.hpp -- host-side GF(2) jump algebra for the stateless XORWOW streams.
//
// cuRAND seeds path p with curand_init(seed, p, 0): v = J^p v0(seed), J = T^(2^67), T the
// one-draw xorshift map on the 160-bit vector v (curand_kernel.h:720-736, :863-874).  cuRAND
// ships J^(4^k) as 32 precomputed 160x160 bit matrices; this engine instead rebuilds the
// binary powers T^(2^k) and J^(2^k) from the recurrence itself at start-up (about 115
// squarings, a few milliseconds) and uploads them once per device.
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace hw1f {

constexpr int kXwWords = 5;
constexpr int kXwBits = 160;
constexpr int kXwRowWords = kXwBits * kXwWords;  // 800 uint32 per matrix, row-major rows = images of e_b
constexpr int kNumPow = 48;                       // powers 2^0 .. 2^47 of both T and J

using BitVec = std::array<uint32_t, kXwWords>;

struct BitMatrix {
    // col[b] = M e_b (the image of basis vector b); bit b = 32*word + bit
    std::array<BitVec, kXwBits> col;
    BitVec apply(const BitVec& x) const;
    BitMatrix after(const BitMatrix& first) const;  // this o first
    static BitMatrix identity();
    static BitMatrix xorwow_step();                  // T
};

struct JumpTables {
    std::vector<BitMatrix> step_pow2;  // T^(2^k)
    std::vector<BitMatrix> seq_pow2;   // J^(2^k)
    JumpTables();
    // flat uint32 images for upload: [kNumPow][160][5]
    std::vector<uint32_t> flat_step() const;
    std::vector<uint32_t> flat_seq() const;
    // J^(n * 16^j) for the four low nibbles j of a subsequence index, n = 0..15 (n = 0: identity): flat
    // [4][16][160][5].  prep_lo_kernel applies at most four of them in a row instead of one J^(2^k) per set bit.
    std::vector<uint32_t> flat_seq_nibbles() const;
};

const JumpTables& jump_tables();  // process-wide singleton

// curand_init's seed scramble (curand_kernel.h:805-819): returns d0, fills v0
uint32_t seed_scramble(uint64_t seed, BitVec& v0);

}  // namespace hw1f
