// tcgen05 / TMA / TMEM sweep GEMM for sm_100a with the fused epilogue.
//
//   out[M x N] = epi( sum_src sum_(pa,pb) A_src,pa[M x K] * B_src,pb[N x K]^T )
//
// A: state planes (bf16, K-major, row = chain), B: weight planes (bf16, K-major:
// W[V x ldH] for the down-sweep, Wt[H x ldV] for the up-sweep).  The hi/lo
// operand planes are extra MMA passes into the same fp32 TMEM accumulator.
//
// Persistent CTAs (one per SM), warp-specialised:
//   warp 0      TMA producer      (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      MMA issuer        (tcgen05.mma kind::f16, M=128, N=BN, K=16 per instruction)
//   warp 2      TMEM allocator
//   warps 4-19  epilogue          (tcgen05.ld 32x32b -> bias/activation/Philox draw -> bf16 planes)
// Two TMEM accumulator stages (2 x BN columns) let the MMAs of tile j+1 run
// under the (sampler-bound) epilogue of tile j.
#pragma once
#include <cuda.h>

#include "epilogue.cuh"

namespace scdbm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
    PFN_tmapEncodeTiled encode = nullptr;
    bool attr_set[6] = {false, false, false, false, false, false};
};

struct TcOperandPlanes {
    MatView a;                       // state (1 or 2 planes)
    const __nv_bfloat16 *b0, *b1;    // weight planes [N x ldb]
    int64_t ldb, N, K;
};

struct TcGradArgs {
    MatView v, h, vm, hm;
    float pos_scale, neg_scale;
    float* out;
    int64_t V, H;
};

constexpr int TC_BM = 128, TC_BK = 64, TC_EPW = 16, TC_THREADS = 32 * (4 + TC_EPW);
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_BUDGET = 200 * 1024;

struct alignas(64) TcMaps {
    CUtensorMap a[2][2];
    CUtensorMap b[2][2];
};

struct TcParams {
    int32_t nsrc;
    int32_t kblocks[2];
    int32_t nA[2];
    uint32_t combos[2];   // bit (pa*2+pb)
    int64_t M, N;
    int32_t m_tiles, n_tiles;
    int32_t stages, stage_bytes;
    EpiParams epi;
};

inline void tc_context_init(TcContext& t) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        t.encode = (PFN_tmapEncodeTiled)fn;
}

// ------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();  // never hang the GPU: a protocol bug becomes a launch error
    } while (!done);
}
// service warps (TMA producer, MMA issuer) share an SM sub-partition with epilogue warps: back off
// between polls so that their spin does not take issue slots from the sampler
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(200);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128B-swizzled operand tile: 8-row groups of 1024 B (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                    // version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ kernel
// EK: epilogue class -- 0 generic (all modes, accurate math, explicit uniforms), 1 Philox-fed NB draw
// (fast math), 2 Philox-fed Bernoulli draw (fast math).  Separate instantiations keep each kernel's
// code small enough for the instruction cache.
template <int BN, int EK>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    constexpr int NS = TC_EPW / 4;       // column slices per tile
    constexpr int CW = BN / NS;          // columns per epilogue warp per tile (16 or 32)
    constexpr int B_TILE = BN * TC_BK * 2;
    constexpr uint32_t TMEM_COLS = 2 * BN;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    static_assert(CW == 16 || CW == 32, "epilogue slice must be 16 or 32 columns");
    (void)tmem_ld16;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_holder;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&bar_full[i]), 1);
            mbar_init(smem_u32(&bar_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_tfull[i]), 1);
            mbar_init(smem_u32(&bar_tempty[i]), TC_EPW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * TC_BM, n0 = (tile % p.n_tiles) * BN;   // n fastest: CTAs sharing an A tile run together (L2)
                for (int s = 0; s < p.nsrc; ++s) {
                    const uint32_t tx = (uint32_t)(p.nA[s] * TC_A_TILE + 2 * B_TILE);
                    for (int kb = 0; kb < p.kblocks[s]; ++kb) {
                        mbar_wait_backoff(smem_u32(&bar_empty[stage]), phase ^ 1u);
                        const uint32_t full = smem_u32(&bar_full[stage]);
                        const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                        mbar_expect_tx(full, tx);
                        for (int pa = 0; pa < p.nA[s]; ++pa) tma_load_2d(sa + pa * TC_A_TILE, &maps.a[s][pa], full, kb * TC_BK, m0);
                        for (int pb = 0; pb < 2; ++pb)
                            tma_load_2d(sa + 2 * TC_A_TILE + pb * B_TILE, &maps.b[s][pb], full, kb * TC_BK, n0);
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait_backoff(smem_u32(&bar_tempty[astage]), aphase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)astage * BN;
                uint32_t accumulate = 0;
                for (int s = 0; s < p.nsrc; ++s) {
                    for (int kb = 0; kb < p.kblocks[s]; ++kb) {
                        mbar_wait_backoff(smem_u32(&bar_full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                        const uint32_t sb = sa + 2 * TC_A_TILE;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (!((p.combos[s] >> c) & 1u)) continue;
                            const uint32_t a0 = sa + (c >> 1) * TC_A_TILE, b0 = sb + (c & 1) * B_TILE;
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k) {
                                tc_mma_f16(tmem_d, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), IDESC, accumulate);
                                accumulate = 1;
                            }
                        }
                        tc_commit(smem_u32(&bar_empty[stage]));   // smem slot is free once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
                tc_commit(smem_u32(&bar_tfull[astage]));          // accumulator ready for the epilogue
                if (++astage == 2) { astage = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps
        const int q = warp & 3, slice = (warp - 4) >> 2;
        int astage = 0;
        uint32_t aphase = 0;
        float delta_local = 0.0f;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t m0 = (int64_t)(tile / p.n_tiles) * TC_BM, n0 = (int64_t)(tile % p.n_tiles) * BN;
            mbar_wait(smem_u32(&bar_tfull[astage]), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * BN + slice * CW);
            const int64_t row = m0 + q * 32 + lane;
            float rowsum = 0.0f;
            // one 4-column group at a time straight out of TMEM: a single copy of the epilogue body
#pragma unroll 1
            for (int g = 0; g < CW / 4; ++g) {
                uint32_t r[4];
                tmem_ld4(taddr + g * 4, r);          // warp-collective: outside any row/column predicate
                tmem_wait_ld();
                const int64_t col0 = n0 + slice * CW + g * 4;
                const int nvalid = (int)max((int64_t)0, min((int64_t)4, p.N - col0));
                if (row < p.M && nvalid > 0) {
                    const float acc[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
                    if constexpr (EK == 1) epi_nb_draw4(p.epi, row, col0, acc, nvalid);
                    else if constexpr (EK == 2) epi_bern4(p.epi, row, col0, acc, nvalid);
                    else rowsum += epi_apply4(p.epi, row, col0, acc, nvalid, delta_local);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[astage]));   // TMEM stage can be overwritten
            if (++astage == 2) { astage = 0; aphase ^= 1u; }
            if constexpr (EK == 0) {
                if (p.epi.mode == EPI_AIS_RATIO && row < p.M) atomicAdd(&p.epi.logw[row], (double)rowsum);
            }
        }
        if (p.epi.prev.p0 && p.epi.delta_bits) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) delta_local = fmaxf(delta_local, __shfl_xor_sync(0xffffffffu, delta_local, off));
            if (lane == 0 && delta_local > 0.0f) atomicMax(p.epi.delta_bits, __float_as_uint(delta_local));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------- host
inline bool tc_make_map(TcContext& t, CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return t.encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the tcgen05 kernel tiles any sweep whose operands exist as bf16 planes
inline bool tc_supported(const GemmArgs& g, int nsrc) {
    if (nsrc < 1 || nsrc > 2 || g.M < 1 || g.N < 1) return false;
    for (int s = 0; s < nsrc; ++s)
        if (g.src[s].K < 1) return false;
    return true;
}
inline bool tc_grad_supported(const TcGradArgs&) { return false; }
inline cudaError_t launch_grad_tc(TcContext&, const TcGradArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }

template <int BN, int EK>
inline cudaError_t tc_launch_bn(TcContext& t, const TcMaps& maps, TcParams& p, int num_sms, cudaStream_t stream) {
    const int slot = (BN == 64 ? 0 : 3) + EK;
    constexpr int B_TILE = BN * TC_BK * 2;
    p.stage_bytes = 2 * TC_A_TILE + 2 * B_TILE;
    p.stages = std::min(TC_MAX_STAGES, TC_SMEM_BUDGET / p.stage_bytes);
    const int smem = p.stages * p.stage_bytes + 1024;
    if (!t.attr_set[slot]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET + 1024);
        if (e != cudaSuccess) return e;
        t.attr_set[slot] = true;
    }
    const int grid = std::min(p.m_tiles * p.n_tiles, num_sms);
    k_gemm_tc<BN, EK><<<grid, TC_THREADS, smem, stream>>>(maps, p);
    return cudaGetLastError();
}

inline cudaError_t launch_gemm_tc(TcContext& t, const TcOperandPlanes* ops, int nsrc, const GemmArgs& g, int num_sms,
                                  cudaStream_t stream) {
    if (!t.encode) return cudaErrorNotSupported;
    const int BN = g.N <= 64 ? 64 : 128;
    TcMaps maps;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.nsrc = nsrc;
    p.M = g.M;
    p.N = g.N;
    p.m_tiles = (int)((g.M + TC_BM - 1) / TC_BM);
    p.n_tiles = (int)((g.N + BN - 1) / BN);
    p.epi = g.epi;
    for (int s = 0; s < nsrc; ++s) {
        const MatView& a = ops[s].a;
        p.kblocks[s] = (int)((ops[s].K + TC_BK - 1) / TC_BK);
        p.nA[s] = a.p1 ? 2 : 1;
        // plane products: binary x (hi,lo); count (lo,hi) x (hi,lo); real (hi,lo) x (hi,lo) minus lo*lo
        p.combos[s] = a.p1 ? (a.kind == MAT_COUNT ? 0xFu : 0x7u) : 0x3u;
        bool ok = tc_make_map(t, &maps.a[s][0], a.p0, g.M, ops[s].K, a.ld, TC_BM);
        ok = ok && tc_make_map(t, &maps.a[s][1], a.p1 ? a.p1 : a.p0, g.M, ops[s].K, a.ld, TC_BM);
        ok = ok && tc_make_map(t, &maps.b[s][0], ops[s].b0, ops[s].N, ops[s].K, ops[s].ldb, BN);
        ok = ok && tc_make_map(t, &maps.b[s][1], ops[s].b1, ops[s].N, ops[s].K, ops[s].ldb, BN);
        if (!ok) return cudaErrorInvalidValue;
    }
    if (nsrc == 1) { maps.a[1][0] = maps.a[0][0]; maps.a[1][1] = maps.a[0][1]; maps.b[1][0] = maps.b[0][0]; maps.b[1][1] = maps.b[0][1]; }
    // epilogue class: the fast sampling kernels need Philox-fed draws into the canonical plane layout
    const EpiParams& e = g.epi;
    int ek = 0;
    if (!e.uniforms && !e.addend && !e.prev.p0) {
        if (e.mode == EPI_NB_DRAW && e.factor == 1.0f && e.nbc && !e.bias1 && e.out.p1 && !e.out.f && e.out.kind == MAT_COUNT) ek = 1;
        if (e.mode == EPI_SIGM_BERN && e.bias0 && !e.out.p1 && !e.out.f) ek = 2;
    }
    if (BN == 64) {
        if (ek == 1) return tc_launch_bn<64, 1>(t, maps, p, num_sms, stream);
        if (ek == 2) return tc_launch_bn<64, 2>(t, maps, p, num_sms, stream);
        return tc_launch_bn<64, 0>(t, maps, p, num_sms, stream);
    }
    if (ek == 1) return tc_launch_bn<128, 1>(t, maps, p, num_sms, stream);
    if (ek == 2) return tc_launch_bn<128, 2>(t, maps, p, num_sms, stream);
    return tc_launch_bn<128, 0>(t, maps, p, num_sms, stream);
}

}  // namespace scdbm
