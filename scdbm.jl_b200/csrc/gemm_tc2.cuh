// Two-SM (cta_group::2) variant of the fused sweep GEMM for the L2-bandwidth-bound deterministic / Bernoulli sweeps.
//
// The single-SM kernel streams 64 B/clk/SM of operands from L2 (256 x 128 tile, two weight planes) and the chip's L2
// delivers ~43 B/clk/SM: the tensor pipe idles half of the time.  Here a cluster of two CTAs (an SM pair) computes a
// 256-row x 256-column tile with ONE tcgen05.mma.cta_group::2 per k-step: each CTA stages its own 128 rows of A and
// only ITS HALF (128 of 256 output columns) of each weight-plane tile, the tensor cores of both SMs read both halves,
// and each CTA's TMEM receives its 128 rows x 256 columns.  Operand traffic: 16 KB (A) + 2 x 16 KB (B halves) per
// k-block and SM for 1024 MMA cycles = 48 B/clk/SM, with double-buffered TMEM (2 x 256 columns).
//
// Protocol (rank 0 of the cluster = leader):
//   full[s]   (leader's)   TMA loads of BOTH CTAs complete_tx on it (cta_group::2 loads, peer bit of the barrier address
//                          cleared); the leader's producer arms it with the byte count of both CTAs
//   empty[s]  (each CTA)   the leader's tcgen05.commit, multicast to both CTAs: the stage's MMAs have retired
//   tfull[a]  (each CTA)   commit multicast: accumulator stage a is complete -> the epilogue warps of both CTAs
//   tempty[a] (leader's)   every epilogue warp of both CTAs arrives (remote arrive through mapa for rank 1)
// Epilogues: the Bernoulli draw (EK = 2) and the deterministic class (EK = 3) of gemm_tc.cuh, same per-element code.
#pragma once
#include "gemm_tc.cuh"

namespace scdbm {

constexpr int C2_BN = 256, C2_EPW = 16, C2_CW = 64;
constexpr int C2_B_TILE = 128 * TC_BK * 2;     // this CTA's half of one weight-plane tile: 128 output columns x 64 K
template <int EK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (4 + C2_EPW), 1)
    k_gemm_cg2(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    static_assert(EK == 2 || EK == 3, "the two-SM kernel serves the Bernoulli and the deterministic epilogue class");
    constexpr int ASTAGES = 2;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(C2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // c = F32, a = b = F16, K-major, M = 256
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_holder;
    if constexpr (EK == 3) {
        if (p.epi.skip && *p.epi.skip) return;   // both CTAs of a cluster read the same (stable) flag
    }
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int svc = warp - C2_EPW;
    const uint32_t rank = cluster_ctarank();
    const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);
    const int nwork = p.m_tiles * p.n_tiles;
    const bool tposed = EK == 3 && p.tposed != 0;

    if (svc == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&bar_full[i]), 1);
            mbar_init(smem_u32(&bar_empty[i]), 1);
        }
        for (int i = 0; i < ASTAGES; ++i) {
            mbar_init(smem_u32(&bar_tfull[i]), 1);
            mbar_init(smem_u32(&bar_tempty[i]), 2 * C2_EPW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (svc == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();            // barriers of BOTH CTAs are initialised before any remote signal can arrive
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (svc == 0) {
        if (lane == 0) {
            // ===================== TMA producer (both CTAs: own A rows, own half of the B columns)
            int stage = 0;
            uint32_t phase = 0;
            for (int w = cl; w < nwork; w += ncl) {
                const int mt = w / p.n_tiles, nt = w % p.n_tiles;
                const int m0 = mt * 256 + (int)rank * TC_BM, n0 = nt * C2_BN + (int)rank * 128;
                for (int s = 0; s < p.nsrc; ++s) {
                    const bool hi = tc_use_hi(p, s, mt);
                    const int npass = p.split_hi[s] ? (hi ? 2 : 1) : 1;
                    const int nA = p.split_hi[s] ? 1 : (hi ? p.nA[s] : 1);
                    const uint32_t tx = 2u * ((uint32_t)nA * TC_A_TILE + 2u * C2_B_TILE);     // both CTAs
                    for (int pass = 0; pass < npass; ++pass)
                        for (int kb = 0; kb < p.kblocks[s]; ++kb) {
                            mbar_wait_parked(smem_u32(&bar_empty[stage]), phase ^ 1u);
                            const uint32_t full_leader = smem_u32(&bar_full[stage]) & C2_PEER_MASK;
                            const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                            if (rank == 0) mbar_expect_tx(smem_u32(&bar_full[stage]), tx);
                            for (int pa = 0; pa < nA; ++pa) tma_load_2d_cg2(sa + pa * TC_A_TILE, &maps.a[s][pa + pass], full_leader, kb * TC_BK, m0);
                            for (int pb = 0; pb < 2; ++pb)
                                tma_load_2d_cg2(sa + (uint32_t)p.a_tiles * TC_A_TILE + pb * C2_B_TILE, &maps.b[s][pb], full_leader, kb * TC_BK, n0);
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                }
            }
        }
    } else if (svc == 1) {
        if (lane == 0 && rank == 0) {
            // ===================== MMA issuer (leader CTA only)
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int w = cl; w < nwork; w += ncl) {
                const int mt = w / p.n_tiles;
                const int tkb = tc_tile_kblocks(p, mt);
                const int nch = tc_chunks(tkb), per = (tkb + nch - 1) / nch;
                int done = 0, in_chunk = 0;
                uint32_t tmem_d = 0, accumulate = 0;
                for (int s = 0; s < p.nsrc; ++s) {
                    const bool hi = tc_use_hi(p, s, mt);
                    const uint32_t combos = (p.split_hi[s] || !hi) ? (p.combos[s] & 0x3u) : p.combos[s];
                    const int nkb = p.kblocks[s] * ((p.split_hi[s] && hi) ? 2 : 1);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (in_chunk == 0) {
                            mbar_wait_parked(smem_u32(&bar_tempty[astage]), aphase ^ 1u);
                            tc_fence_after();
                            tmem_d = tmem_base + (uint32_t)astage * C2_BN;
                            accumulate = 0;
                        }
                        mbar_wait_parked(smem_u32(&bar_full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                        const uint32_t sb = sa + (uint32_t)p.a_tiles * TC_A_TILE;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (!((combos >> c) & 1u)) continue;
                            const uint32_t a0 = sa + (c >> 1) * TC_A_TILE, b0 = sb + (c & 1) * C2_B_TILE;
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k)     // tposed: the weight tile is the MMA's "A" (rows of D = output units)
                                tc_mma_f16_cg2(tmem_d, umma_desc_k_sw128((tposed ? b0 : a0) + k * 32), umma_desc_k_sw128((tposed ? a0 : b0) + k * 32), IDESC,
                                               (accumulate | (uint32_t)(c != __ffs(combos) - 1) | (uint32_t)k) ? 1u : 0u);
                        }
                        accumulate = 1;
                        tc_commit_cg2(smem_u32(&bar_empty[stage]), 3);        // the smem slots of both CTAs are free once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        ++done;
                        if (++in_chunk == per || done == tkb) {
                            tc_commit_cg2(smem_u32(&bar_tfull[astage]), 3);   // accumulator (chunk) ready in both CTAs
                            if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
                            in_chunk = 0;
                        }
                    }
                }
            }
        }
    } else if (svc < 0) {
        // ===================== epilogue warps (both CTAs): quadrant = warp % 4 (32 rows), slice = warp / 4 (64 columns)
        const int q = warp & 3, slice = warp >> 2;
        int astage = 0;
        uint32_t aphase = 0;
        float delta_local = 0.0f;
        const float acc_scale = p.epi.acc_scale ? __ldg(p.epi.acc_scale) : 1.0f;
        for (int w = cl; w < nwork; w += ncl) {
            const int mt = w / p.n_tiles, nt = w % p.n_tiles;
            const int64_t m0 = (int64_t)mt * 256 + (int64_t)rank * TC_BM, n0 = (int64_t)nt * C2_BN;
            // this lane's two columns of the warp's slice: summed biases, inverse dispersion
            float bias_l[2] = {0.0f, 0.0f}, r_l[2] = {1.0f, 1.0f};
            // transposed accumulator: this lane's output unit, its summed biases and inverse dispersion
            const int64_t unit = n0 + (int64_t)rank * 128 + q * 32 + lane;
            if (tposed) {
                if (unit < p.N) {
                    if (p.epi.bias0) bias_l[0] = __ldg(p.epi.bias0 + unit);
                    if (p.epi.bias1) bias_l[0] += __ldg(p.epi.bias1 + unit);
                    if (p.epi.r && p.epi.mode == EPI_NBMEAN) r_l[0] = __ldg(p.epi.r + unit);
                }
            } else
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int64_t c = n0 + slice * C2_CW + t * 32 + lane;
                if (c < p.N) {
                    if (p.epi.bias0) bias_l[t] = __ldg(p.epi.bias0 + c);
                    if (p.epi.bias1) bias_l[t] += __ldg(p.epi.bias1 + c);
                    if (p.epi.r && p.epi.mode == EPI_NBMEAN) r_l[t] = __ldg(p.epi.r + c);
                }
            }
            if constexpr (EK == 3) {
                if (!tposed) epi_prefetch_extras(p.epi, m0 + q * 32 + lane, p.M, n0 + slice * C2_CW, p.N, C2_CW);
            }
            const int nch = tc_chunks(tc_tile_kblocks(p, mt));
            if (nch > 1) {   // chunked accumulation with fp32 promotion (see tc_chunks)
                float sum[C2_CW];
#pragma unroll
                for (int i = 0; i < C2_CW; ++i) sum[i] = 0.0f;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * C2_BN + slice * C2_CW);
#pragma unroll
                    for (int g = 0; g < C2_CW / 16; ++g) {
                        uint32_t r[16];
                        tmem_ld16(ta + g * 16, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) sum[g * 16 + i] += __uint_as_float(r[i]);
                    }
                    if (c + 1 < nch) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[astage]), 0);
                        if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
                    } else {
#pragma unroll
                        for (int g = 0; g < C2_CW / 16; ++g) {
                            uint32_t r[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(sum[g * 16 + i]);
                            tmem_st16(ta + g * 16, r);
                        }
                        tmem_wait_st();
                    }
                }
            } else {
                mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                tc_fence_after();
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * C2_BN + slice * C2_CW);
            const int64_t row = m0 + q * 32 + lane;
            if (tposed) {
                if constexpr (EK == 3) {
                    // TMEM column = row of the 256-row cluster tile (both CTAs' rows), lane = output unit
                    const int64_t row0 = (int64_t)mt * 256 + slice * C2_CW;
#pragma unroll 1
                    for (int g = 0; g < C2_CW / 4; ++g) {
                        uint32_t r[4];
                        tmem_ld4(taddr + g * 4, r);
                        tmem_wait_ld();
                        const int nrows = (int)max((int64_t)0, min((int64_t)4, p.M - (row0 + g * 4)));
                        if (unit < p.N && nrows > 0) {
                            const float acc[4] = {__uint_as_float(r[0]) * acc_scale, __uint_as_float(r[1]) * acc_scale, __uint_as_float(r[2]) * acc_scale,
                                                  __uint_as_float(r[3]) * acc_scale};
                            epi_det4t(p.epi, row0 + g * 4, nrows, unit, acc, delta_local, bias_l[0], r_l[0]);
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int g = 0; g < C2_CW / 4; ++g) {
                uint32_t r[4];
                tmem_ld4(taddr + g * 4, r);
                float bias4[4], r4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = g * 4 + j;
                    const float b0 = __shfl_sync(0xffffffffu, bias_l[0], c & 31), b1 = __shfl_sync(0xffffffffu, bias_l[1], c & 31);
                    const float q0 = __shfl_sync(0xffffffffu, r_l[0], c & 31), q1 = __shfl_sync(0xffffffffu, r_l[1], c & 31);
                    bias4[j] = c < 32 ? b0 : b1;
                    r4[j] = c < 32 ? q0 : q1;
                }
                tmem_wait_ld();
                const int64_t col0 = n0 + slice * C2_CW + g * 4;
                const int nvalid = (int)max((int64_t)0, min((int64_t)4, p.N - col0));
                if (row < p.M && nvalid > 0) {
                    const float acc[4] = {__uint_as_float(r[0]) * acc_scale, __uint_as_float(r[1]) * acc_scale, __uint_as_float(r[2]) * acc_scale,
                                          __uint_as_float(r[3]) * acc_scale};
                    if constexpr (EK == 2) epi_bern4(p.epi, row, col0, acc, nvalid, bias4);
                    else epi_det4(p.epi, row, col0, acc, nvalid, delta_local, bias4, r4);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[astage]), 0);   // the leader's MMA issuer may overwrite this TMEM stage
            if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
        }
        if (p.epi.prev.p0 && p.epi.delta_bits) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) delta_local = fmaxf(delta_local, __shfl_xor_sync(0xffffffffu, delta_local, off));
            if (lane == 0 && delta_local > 0.0f) atomicMax(p.epi.delta_bits, __float_as_uint(delta_local));
        }
    }
    tc_fence_before();
    cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still signal it or read its shared memory
    if (svc == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------- host
// The two-SM kernel pays off where the single-SM one is L2-operand-bound and there are enough 256 x 256 cluster tiles to
// fill the GPU: Bernoulli / deterministic sweeps over many rows with a wide output.
inline bool tc_cg2_eligible(const GemmArgs& g, const TcOperandPlanes* ops, int nsrc, int num_sms) {
    const int ek = tc_epilogue_class(g, ops, nsrc);
    if (ek != 2 && ek != 3) return false;
    if (g.N < 192) return false;
    const int64_t tiles = ((g.M + 255) / 256) * ((g.N + C2_BN - 1) / C2_BN);
    return tiles >= num_sms / 2;
}

template <int EK>
inline cudaError_t cg2_launch(TcContext& t, const TcMaps& maps, TcParams& p, int num_sms, cudaStream_t stream) {
    p.a_tiles = std::max(p.split_hi[0] ? 1 : p.nA[0], p.nsrc > 1 ? (p.split_hi[1] ? 1 : p.nA[1]) : 1);
    p.stage_bytes = p.a_tiles * TC_A_TILE + 2 * C2_B_TILE;
    p.stages = std::max(1, std::min(TC_MAX_STAGES, (TC_SMEM_BUDGET + 20 * 1024) / p.stage_bytes));
    const int smem = p.stages * p.stage_bytes + 1024;
    const int slot = 12 + (EK == 2 ? 0 : 1);
    if (!t.attr_set[slot]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_cg2<EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        t.attr_set[slot] = true;
    }
    const int nwork = p.m_tiles * p.n_tiles;
    const int clusters = std::min(nwork, num_sms / 2);
    k_gemm_cg2<EK><<<2 * clusters, 32 * (4 + C2_EPW), smem, stream>>>(maps, p);
    return cudaGetLastError();
}

inline cudaError_t launch_gemm_cg2(TcContext& t, const TcOperandPlanes* ops, int nsrc, const GemmArgs& g, int num_sms, cudaStream_t stream) {
    if (!t.encode) return cudaErrorNotSupported;
    const int ek = tc_epilogue_class(g, ops, nsrc);
    TcMaps maps;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.nsrc = nsrc;
    p.M = g.M;
    p.N = g.N;
    p.m_tiles = (int)((g.M + 255) / 256);
    p.n_tiles = (int)((g.N + C2_BN - 1) / C2_BN);
    p.pair = 1;            // hi-plane flags: a 256-row cluster tile covers two flagged 128-row tiles (tc_use_hi)
    p.epi = g.epi;
    for (int s = 0; s < nsrc; ++s) {
        const MatView& a = ops[s].a;
        p.kblocks[s] = (int)((ops[s].K + TC_BK - 1) / TC_BK);
        p.a_hiflag[s] = (a.kind == MAT_COUNT) ? a.hiflag : nullptr;
        p.nA[s] = a.p1 ? 2 : 1;
        p.combos[s] = a.p1 ? (a.kind == MAT_COUNT ? 0xFu : 0x7u) : 0x3u;
        p.split_hi[s] = (a.p1 && a.kind == MAT_COUNT) ? 1 : 0;
        bool ok = tc_make_map(t, &maps.a[s][0], a.p0, g.M, ops[s].K, a.ld, TC_BM);
        ok = ok && tc_make_map(t, &maps.a[s][1], a.p1 ? a.p1 : a.p0, g.M, ops[s].K, a.ld, TC_BM);
        ok = ok && tc_make_map(t, &maps.b[s][0], ops[s].b0, ops[s].N, ops[s].K, ops[s].ldb, 128);
        ok = ok && tc_make_map(t, &maps.b[s][1], ops[s].b1, ops[s].N, ops[s].K, ops[s].ldb, 128);
        if (!ok) return cudaErrorInvalidValue;
    }
    if (nsrc == 1) { maps.a[1][0] = maps.a[0][0]; maps.a[1][1] = maps.a[0][1]; maps.b[1][0] = maps.b[0][0]; maps.b[1][1] = maps.b[0][1]; }
    if (ek == 2) return cg2_launch<2>(t, maps, p, num_sms, stream);
    static const bool tposed_on = [] { const char* v = getenv("SCDBM_CG2T"); return !(v && v[0] == '0'); }();
    p.tposed = tposed_on ? 1 : 0;
    return cg2_launch<3>(t, maps, p, num_sms, stream);
}

// =====================================================================================
// Two-SM gradient GEMM (K5): dW[V x H] = s0 v'h + s1 vm'hm for the large shapes (C4: 20 000 x 1 024 over 108 192 samples).
// A cluster computes a 256 x 256 tile of dW; each CTA stages its own 128 visible units (one plane per pass) and HALF of the
// hidden units' planes (128 of 256): 16 KB + 2 x 16 KB per 1 024 MMA cycles = 48 B/clk/SM of operand traffic against 64 for
// the single-SM kernel with 256-row tiles.  Pass structure, chunked accumulation and scaling as in k_grad_tc; barrier
// protocol as in k_gemm_cg2.  Each epilogue warp owns 32 visible units x 64 hidden units of its CTA's 128 x 256 block.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TG_THREADS, 1)
    k_grad_cg2(const __grid_constant__ TcGradMaps maps, const __grid_constant__ TcGradParams p) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // F32 acc, F16 MN-major operands, M = N = 256
    constexpr uint32_t TMEM_COLS = 512;        // 2 accumulator stages x 256 columns
    constexpr int STAGE = tg_stage_bytes(1);   // 16 KB A plane + 2 x 16 KB B half-planes
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_holder;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int svc = warp - 16;
    const uint32_t rank = cluster_ctarank();
    const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);
    const int nwork = p.m_tiles * p.n_tiles * p.splits;

    if (svc == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(smem_u32(&bar_full[i]), 1); mbar_init(smem_u32(&bar_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (svc == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    auto kb_range = [&](int s, int sp, int& lo, int& hi) {
        lo = (int)((int64_t)p.kblocks[s] * sp / p.splits);
        hi = (int)((int64_t)p.kblocks[s] * (sp + 1) / p.splits);
    };
    auto pass_masks = [&](int s, int kb, uint32_t& m0, uint32_t& m1) {
        m0 = p.combos[s] & 3u;
        m1 = (p.combos[s] >> 2) & 3u;
        if (p.a_hiflag[s] && (p.a_hiflag[s][kb / 2] & 1u) == 0u) m1 = 0u;      // COUNT operand: hi plane unused in this 128-sample block
    };

    if (svc == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = cl; w < nwork; w += ncl) {
                const int tile = w / p.splits, sp = w % p.splits;
                const int m0 = (tile / p.n_tiles) * 256 + (int)rank * 128, n0 = (tile % p.n_tiles) * 256 + (int)rank * 128;
                for (int s = 0; s < 2; ++s) {
                    int lo, hi;
                    kb_range(s, sp, lo, hi);
                    for (int kb = lo; kb < hi; ++kb) {
                        uint32_t pm[2];
                        pass_masks(s, kb, pm[0], pm[1]);
                        for (int pa = 0; pa < 2; ++pa) {
                            if (!pm[pa]) continue;
                            const int nb = (pm[pa] & 2u) ? 2 : 1;
                            mbar_wait_parked(smem_u32(&bar_empty[stage]), phase ^ 1u);
                            const uint32_t full_leader = smem_u32(&bar_full[stage]) & C2_PEER_MASK;
                            const uint32_t sa = smem_base + (uint32_t)stage * STAGE, sb = sa + TG_PLANE;
                            if (rank == 0) mbar_expect_tx(smem_u32(&bar_full[stage]), 2u * (uint32_t)((1 + nb) * TG_PLANE));
                            for (int blk = 0; blk < 2; ++blk) tma_load_2d_cg2(sa + blk * 8192, &maps.a[s][pa], full_leader, m0 + blk * 64, kb * TC_BK);
                            for (int pb = 0; pb < nb; ++pb)
                                for (int blk = 0; blk < 2; ++blk)
                                    tma_load_2d_cg2(sb + pb * TG_PLANE + blk * 8192, &maps.b[s][pb], full_leader, n0 + blk * 64, kb * TC_BK);
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (svc == 1) {
        if (lane == 0 && rank == 0) {
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int w = cl; w < nwork; w += ncl) {
                const int sp = w % p.splits;
                for (int s = 0; s < 2; ++s) {
                    int lo, hi;
                    kb_range(s, sp, lo, hi);
                    const int n = hi - lo;
                    if (n <= 0) continue;
                    const int nch = (n + TC_CHUNK_KB - 1) / TC_CHUNK_KB, per = (n + nch - 1) / nch;
                    int in_chunk = 0;
                    uint32_t accumulate = 0;
                    for (int kb = lo; kb < hi; ++kb) {
                        if (in_chunk == 0) {
                            mbar_wait_parked(smem_u32(&bar_tempty[astage]), aphase ^ 1u);
                            tc_fence_after();
                            accumulate = 0;
                        }
                        const uint32_t tmem_d = tmem_base + (uint32_t)astage * 256u;
                        uint32_t pm[2];
                        pass_masks(s, kb, pm[0], pm[1]);
                        for (int pa = 0; pa < 2; ++pa) {
                            if (!pm[pa]) continue;
                            mbar_wait_parked(smem_u32(&bar_full[stage]), phase);
                            tc_fence_after();
                            const uint32_t sa = smem_base + (uint32_t)stage * STAGE, sb = sa + TG_PLANE;
#pragma unroll
                            for (int pb = 0; pb < 2; ++pb) {
                                if (!((pm[pa] >> pb) & 1u)) continue;
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; ++k) {
                                    tc_mma_f16_cg2(tmem_d, umma_desc_mn_sw128(sa + k * 2048), umma_desc_mn_sw128(sb + pb * TG_PLANE + k * 2048), IDESC, accumulate);
                                    accumulate = 1;
                                }
                            }
                            tc_commit_cg2(smem_u32(&bar_empty[stage]), 3);
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                        if (++in_chunk == per || kb + 1 == hi) {
                            tc_commit_cg2(smem_u32(&bar_tfull[astage]), 3);
                            if (++astage == 2) { astage = 0; aphase ^= 1u; }
                            in_chunk = 0;
                        }
                    }
                }
            }
        }
    } else if (svc < 0) {
        const int q = warp & 3, slice = warp >> 2;
        int astage = 0;
        uint32_t aphase = 0;
        for (int w = cl; w < nwork; w += ncl) {
            const int tile = w / p.splits, sp = w % p.splits;
            const int64_t m0 = (int64_t)(tile / p.n_tiles) * 256 + (int64_t)rank * 128, nn0 = (int64_t)(tile % p.n_tiles) * 256;
            float sum[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) sum[i] = 0.0f;
            for (int s = 0; s < 2; ++s) {
                int lo, hi;
                kb_range(s, sp, lo, hi);
                const int n = hi - lo;
                if (n <= 0) continue;
                const int nch = (n + TC_CHUNK_KB - 1) / TC_CHUNK_KB;
                const float sc = p.scale[s];
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)astage * 256u + (uint32_t)(slice * 64);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t r[16];
                        tmem_ld16(ta + g * 16, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) sum[g * 16 + i] = fmaf(sc, __uint_as_float(r[i]), sum[g * 16 + i]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[astage]), 0);
                    if (++astage == 2) { astage = 0; aphase ^= 1u; }
                }
            }
            float* dst = (p.splits > 1 ? p.partial + (int64_t)sp * p.V * p.H : p.out);
            const int64_t c0 = nn0 + slice * 64;
            const bool vec = c0 + 64 <= p.H && (p.H & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0u;
            const int64_t row = m0 + q * 32 + lane;
            if (row < p.V) {
                float* d = dst + row * p.H + c0;
                if (vec) {
#pragma unroll
                    for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (c0 + i < p.H) d[i] = sum[i];
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (svc == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// the two-SM gradient kernel serves shapes with at least one wave of 256 x 256 cluster tiles and many samples
inline bool tc_grad_cg2_eligible(const TcGradArgs& a, int num_sms) {
    if (a.da || a.V < 256 || a.H < 256 || a.v.rows + a.vm.rows < 32 * TC_BK) return false;
    return ((a.V + 255) / 256) * ((a.H + 255) / 256) >= num_sms / 2;     // one wave of clusters; K splits even out the rest
}
// K splits that even out the last wave of cluster tiles (partial sums are reduced in fixed order by k_grad_reduce)
inline int tc_grad_cg2_splits(const TcGradArgs& a, int num_sms) {
    const int64_t tiles = ((a.V + 255) / 256) * ((a.H + 255) / 256);
    const int ncl = num_sms / 2;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 4; ++s) {
        const double waves = (double)(tiles * s) / ncl;
        const double eff = waves / std::ceil(waves) - 0.01 * (s - 1);     // each extra split costs a slab of traffic
        if (eff > best_eff) { best_eff = eff; best = s; }
    }
    return best;
}

inline cudaError_t launch_grad_cg2(TcContext& t, const TcGradMaps& maps, TcGradParams& p, int num_sms, cudaStream_t stream) {
    p.m_tiles = (int)((p.V + 255) / 256);
    p.n_tiles = (int)((p.H + 255) / 256);
    p.stages = 4;
    p.nbias = 0;
    const int smem = p.stages * tg_stage_bytes(1) + 1024;
    if (!t.attr_set[16]) {
        cudaError_t e = cudaFuncSetAttribute(k_grad_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        t.attr_set[16] = true;
    }
    const int nwork = p.m_tiles * p.n_tiles * p.splits;
    const int clusters = std::min(nwork, num_sms / 2);
    k_grad_cg2<<<2 * clusters, TG_THREADS, smem, stream>>>(maps, p);
    return cudaGetLastError();
}

}  // namespace scdbm
