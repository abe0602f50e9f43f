#include "xorwow_jump.hpp"

#include <mutex>

namespace hw1f {

BitVec BitMatrix::apply(const BitVec& x) const
{
    BitVec acc{0, 0, 0, 0, 0};
    for (int w = 0; w < kXwWords; ++w) {
        uint32_t bits = x[w];
        while (bits) {
            const int j = __builtin_ctz(bits);
            bits &= bits - 1;
            const BitVec& c = col[32 * w + j];
            for (int k = 0; k < kXwWords; ++k) acc[k] ^= c[k];
        }
    }
    return acc;
}

BitMatrix BitMatrix::after(const BitMatrix& first) const
{
    BitMatrix out;
    for (int b = 0; b < kXwBits; ++b) out.col[b] = apply(first.col[b]);
    return out;
}

BitMatrix BitMatrix::identity()
{
    BitMatrix m;
    for (int b = 0; b < kXwBits; ++b) {
        m.col[b] = BitVec{0, 0, 0, 0, 0};
        m.col[b][b >> 5] = 1u << (b & 31);
    }
    return m;
}

BitMatrix BitMatrix::xorwow_step()
{
    // one application of the xorshift recurrence to every basis vector
    BitMatrix m;
    for (int b = 0; b < kXwBits; ++b) {
        BitVec v{0, 0, 0, 0, 0};
        v[b >> 5] = 1u << (b & 31);
        const uint32_t t = v[0] ^ (v[0] >> 2);
        const uint32_t n4 = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
        m.col[b] = BitVec{v[1], v[2], v[3], v[4], n4};
    }
    return m;
}

JumpTables::JumpTables()
{
    step_pow2.reserve(kNumPow);
    seq_pow2.reserve(kNumPow);
    BitMatrix m = BitMatrix::xorwow_step();
    for (int k = 0; k < 67 + kNumPow; ++k) {
        if (k < kNumPow) step_pow2.push_back(m);
        if (k >= 67) seq_pow2.push_back(m);
        m = m.after(m);
    }
}

static std::vector<uint32_t> flatten(const std::vector<BitMatrix>& ms)
{
    std::vector<uint32_t> out;
    out.reserve(ms.size() * kXwRowWords);
    for (const auto& m : ms)
        for (int b = 0; b < kXwBits; ++b)
            for (int k = 0; k < kXwWords; ++k) out.push_back(m.col[b][k]);
    return out;
}

std::vector<uint32_t> JumpTables::flat_step() const { return flatten(step_pow2); }
std::vector<uint32_t> JumpTables::flat_seq() const { return flatten(seq_pow2); }

std::vector<uint32_t> JumpTables::flat_seq_nibbles() const
{
    std::vector<BitMatrix> ms;
    ms.reserve(64);
    for (int j = 0; j < 4; ++j)
        for (int n = 0; n < 16; ++n) {
            BitMatrix m = BitMatrix::identity();
            for (int b = 0; b < 4; ++b)
                if ((n >> b) & 1) m = seq_pow2[4 * j + b].after(m);   // powers of one matrix commute
            ms.push_back(m);
        }
    return flatten(ms);
}

const JumpTables& jump_tables()
{
    static const JumpTables t;
    return t;
}

uint32_t seed_scramble(uint64_t seed, BitVec& v0)
{
    const uint32_t s0 = static_cast<uint32_t>(seed) ^ 0xaad26b49u;
    const uint32_t s1 = static_cast<uint32_t>(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    v0 = BitVec{123456789u + t0, 362436069u ^ t0, 521288629u + t1, 88675123u ^ t1, 5783321u + t0};
    return 6615241u + t1 + t0;
}

}  // namespace hw1f
