// CUDA-core (FFMA) sweep GEMM with the fused epilogue.  Used for shapes the
// tcgen05 engine does not tile (tiny / unaligned RBMs, the latency-bound C1
// config) and as the fp32 on-device cross-check of the tensor-core engine.
//   out[M x N] = epi( sum_src scale_src * A_src[M x K] * B_src[K x N] )
#pragma once
#include "epilogue.cuh"

namespace scdbm {

constexpr int ST_M = 64, ST_N = 64, ST_K = 16, ST_THREADS = 256;

__device__ __forceinline__ float operand_load(const Operand& o, int64_t mn, int64_t k) {
    const int64_t i = mn * o.s_mn + k * o.s_k;
    if (o.f) return o.f[i];
    float v = plane_decode(o.p0[i]);
    if (o.p1) v += plane_decode(o.p1[i]);
    return v;
}

__global__ void __launch_bounds__(ST_THREADS) k_gemm_simt(const GemmArgs g) {
    pdl_trigger();
    pdl_wait();
    __shared__ float As[ST_K][ST_M + 4];
    __shared__ float Bs[ST_K][ST_N + 4];
    if (g.epi.skip && *g.epi.skip) return;   // speculatively launched mean-field iteration after convergence
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;  // thread computes rows ty*4.., cols tx*4..
    const int64_t m0 = (int64_t)blockIdx.y * ST_M, n0 = (int64_t)blockIdx.x * ST_N;
    float acc[4][4] = {};
    for (int s = 0; s < g.nsrc; ++s) {
        const GemmSource& src = g.src[s];
        const bool a_kc = src.a.s_k == 1, b_kc = src.b.s_k == 1;
        for (int64_t k0 = 0; k0 < src.K; k0 += ST_K) {
#pragma unroll
            for (int i = 0; i < (ST_M * ST_K) / ST_THREADS; ++i) {
                const int idx = tid + i * ST_THREADS;
                const int kk = a_kc ? idx % ST_K : idx / ST_M;
                const int mm = a_kc ? idx / ST_K : idx % ST_M;
                float v = 0.0f;
                if (m0 + mm < g.M && k0 + kk < src.K) v = src.scale * operand_load(src.a, m0 + mm, k0 + kk);
                As[kk][mm] = v;
            }
#pragma unroll
            for (int i = 0; i < (ST_N * ST_K) / ST_THREADS; ++i) {
                const int idx = tid + i * ST_THREADS;
                const int kk = b_kc ? idx % ST_K : idx / ST_N;
                const int nn = b_kc ? idx / ST_K : idx % ST_N;
                float v = 0.0f;
                if (n0 + nn < g.N && k0 + kk < src.K) v = operand_load(src.b, n0 + nn, k0 + kk);
                Bs[kk][nn] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < ST_K; ++kk) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    float delta_local = 0.0f;
    const int64_t col0 = n0 + tx * 4;
    const int nvalid = (int)max((int64_t)0, min((int64_t)4, g.N - col0));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + ty * 4 + i;
        float s = 0.0f;
        if (row < g.M && nvalid > 0) s = epi_apply4(g.epi, row, col0, acc[i], nvalid, delta_local);
        if (g.epi.mode == EPI_AIS_RATIO) {
            // the 16 threads sharing `ty` cover the 64 columns of this row
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (tx == 0 && row < g.M) atomicAdd(&g.epi.logw[row], (double)s);
        }
    }
    if (g.epi.prev.p0 && g.epi.delta_bits) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) delta_local = fmaxf(delta_local, __shfl_xor_sync(0xffffffffu, delta_local, off));
        if ((tid & 31) == 0 && delta_local > 0.0f) atomicMax(g.epi.delta_bits, __float_as_uint(delta_local));
    }
}

inline cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
    dim3 grid((unsigned)((g.N + ST_N - 1) / ST_N), (unsigned)((g.M + ST_M - 1) / ST_M));
    { cudaError_t le = launch_pdl(k_gemm_simt, dim3(grid), dim3(ST_THREADS), 0, stream, g); if (le != cudaSuccess) return le; }
    return cudaGetLastError();
}

}  // namespace scdbm
