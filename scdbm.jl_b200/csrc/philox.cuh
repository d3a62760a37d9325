// Philox4x32-10 and the element->counter addressing shared with
// oracle/philox.py (Random123 KAT-checked there and in tests/).
#pragma once
#include "common.cuh"

namespace scdbm {

enum Purpose : uint32_t { P_GIBBS = 0, P_INIT = 1, P_INIT2 = 2, P_GAMMA = 3, P_POIS = 4, P_DROPOUT = 5, P_PERM = 6, P_DATA = 7, P_SIZE = 8 };

__host__ __device__ inline uint32_t stream_id(uint32_t layer, uint32_t purpose = P_GIBBS, uint32_t attempt = 0) {
    return (layer & 0xFFu) | ((purpose & 0xFFu) << 8) | ((attempt & 0xFFFFu) << 16);
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// 23 random bits + 1/2: exact in fp32, 2^-24 <= u <= 1-2^-24
__device__ __forceinline__ float u23(uint32_t w) {
    return fmaf(__uint2float_rn(w >> 9), 1.1920928955078125e-07f, 5.9604644775390625e-08f);
}

// grouped addressing: words for columns [4g, 4g+3] of `row`
__device__ __forceinline__ uint4 philox_group(const EpiParams& e, uint32_t row, uint32_t colgroup) {
    return philox4x32_10(make_uint4(colgroup, row + e.row_offset, e.tick, e.stream), e.k0, e.k1);
}
// the same words with the round keys precomputed on the host (EpiParams::rk lives in the kernel's constant bank, so each
// key is a direct operand of the round's LOP3 instead of two uniform-datapath additions per round)
__device__ __forceinline__ uint4 philox_group_rk(const EpiParams& e, uint32_t row, uint32_t colgroup) {
    uint4 c = make_uint4(colgroup, row + e.row_offset, e.tick, e.stream);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ e.rk[2 * i], lo1, hi0 ^ c.w ^ e.rk[2 * i + 1], lo0);
    }
    return c;
}
__host__ inline void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk)[20]) {
    for (int i = 0; i < 10; ++i) {
        rk[2 * i] = k0;
        rk[2 * i + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// element addressing (rejection samplers)
__device__ __forceinline__ uint4 philox_element(uint32_t k0, uint32_t k1, uint32_t row, uint32_t col, uint32_t tick,
                                                uint32_t stream) {
    return philox4x32_10(make_uint4(col, row, tick, stream), k0, k1);
}

}  // namespace scdbm
