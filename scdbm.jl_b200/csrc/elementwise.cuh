// Elementwise / reduction kernels around the sweep GEMMs: layout+dtype
// conversion at the boundary, fused parameter update (+clamps, +operand
// refresh), column sums, particle init, row gather, column statistics.
// All are HBM-bound: coalesced, vectorised where alignment allows.
#pragma once
#include "epilogue.cuh"

namespace scdbm {

// ---- host (column-major fp64, leading dimension ld) <-> device planes
// 32x32 smem transpose so both sides are coalesced.
__global__ void k_upload_f64(const double* __restrict__ src, int64_t lds, MatView dst) {
    __shared__ float t[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + threadIdx.x, c = c0 + j;
        t[j][threadIdx.x] = (r < dst.rows && c < dst.cols) ? (float)src[c * lds + r] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < dst.rows && c < dst.cols) {
            plane_t a, b;
            split_value(dst.kind, t[threadIdx.x][j], a, b);
            dst.p0[r * dst.ld + c] = a;
            if (dst.p1) dst.p1[r * dst.ld + c] = b;
            if (dst.f) dst.f[r * dst.ld + c] = t[threadIdx.x][j];
            if (dst.hiflag && (__half_as_ushort(b) & 0x7FFFu) != 0) dst.hiflag[r / 128] = 3u;
        }
    }
}

__global__ void k_download_f64(MatView src, double* __restrict__ dst, int64_t ldd) {
    __shared__ float t[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        t[j][threadIdx.x] = (r < src.rows && c < src.cols) ? mat_load(src, r, c) : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + threadIdx.x, c = c0 + j;
        if (r < src.rows && c < src.cols) dst[c * ldd + r] = (double)t[threadIdx.x][j];
    }
}

// row-major compact download of counts / binary states: block (32, 8) = 8 rows, a lane converts 8 columns
// (one 16-byte load per plane, one 16-byte store); the hi count plane is read only where its row tile is flagged
__global__ void k_download_u16(MatView src, uint16_t* __restrict__ dst) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= src.rows) return;
    const bool use_hi = src.p1 && (!src.hiflag || (src.hiflag[r / 128] & 1u));
    const bool vec_out = (src.cols & 7) == 0;
    for (int64_t c0 = (int64_t)threadIdx.x * 8; c0 < src.cols; c0 += 32 * 8) {
        float v[8];
        if (src.f) {
            for (int j = 0; j < 8; ++j) v[j] = (c0 + j < src.cols) ? src.f[r * src.ld + c0 + j] : 0.0f;
        } else {
            const uint4 a = *reinterpret_cast<const uint4*>(src.p0 + r * src.ld + c0);      // ld % 64 == 0: in bounds, aligned
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
                v[2 * j] = f.x; v[2 * j + 1] = f.y;
            }
            if (use_hi) {
                const uint4 b = *reinterpret_cast<const uint4*>(src.p1 + r * src.ld + c0);
                const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&bw[j]));
                    v[2 * j] += f.x; v[2 * j + 1] += f.y;
                }
            }
        }
        uint16_t o[8];
        for (int j = 0; j < 8; ++j) o[j] = (uint16_t)fminf(fmaxf(v[j], 0.0f), 65535.0f);
        uint16_t* d = dst + r * src.cols + c0;
        if (vec_out) {
            uint4 w;
            w.x = o[0] | ((uint32_t)o[1] << 16); w.y = o[2] | ((uint32_t)o[3] << 16);
            w.z = o[4] | ((uint32_t)o[5] << 16); w.w = o[6] | ((uint32_t)o[7] << 16);
            *reinterpret_cast<uint4*>(d) = w;
        } else {
            for (int j = 0; j < 8 && c0 + j < src.cols; ++j) d[j] = o[j];
        }
    }
}

__global__ void k_upload_uniforms(const double* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t rows,
                                  int64_t cols, int64_t ldd) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % rows, c = i / rows;
        dst[r * ldd + c] = (float)src[c * lds + r];
    }
}

// ---- weights: fp32 master W[V x H] row-major -> fp16 hi/lo operand planes of s*W in both orientations
// (W[V x ldH] for the down-sweep, Wt[H x ldV] for the up-sweep).  s = scale[0] is the model's power-of-two
// weight scale (device scalar, see k_weight_scale): it keeps |s W| <= 1024 and small weights inside fp16's
// normal range, so hi + lo carries ~22 bits; the tensor-core epilogue multiplies the accumulator by scale[1] = 1/s.
__global__ void k_absmax(const float* __restrict__ W, int64_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(W[i]));
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out_bits, __float_as_uint(m));
}
// exponent e of the power-of-two scale s = 2^e for a model whose largest |W| is m:  |s W| <= 2^10
__device__ __forceinline__ int weight_scale_exp(float m) {
    int e = 0;
    if (m > 0.0f && isfinite(m)) e = 10 - (int)ceilf(log2f(m));
    return max(-20, min(20, e));
}
__global__ void k_weight_scale(const unsigned int* __restrict__ max_bits, float* __restrict__ scale) {
    const int e = weight_scale_exp(__uint_as_float(*max_bits));
    scale[0] = exp2f((float)e);
    scale[1] = exp2f((float)-e);
}
__global__ void k_refresh_operands(const float* __restrict__ W, int V, int H, const float* __restrict__ scale, plane_t* w0,
                                   plane_t* w1, int64_t ldh, plane_t* t0, plane_t* t1, int64_t ldv) {
    __shared__ float t[32][33];
    const float s = scale[0];
    const int v0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int v = v0 + j, h = h0 + threadIdx.x;
        float x = 0.0f;
        if (v < V && h < H) {
            x = s * W[(int64_t)v * H + h];
            const plane_t a = plane_encode(x);
            w0[(int64_t)v * ldh + h] = a;
            w1[(int64_t)v * ldh + h] = plane_encode(x - plane_decode(a));
        }
        t[j][threadIdx.x] = x;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int h = h0 + j, v = v0 + threadIdx.x;
        if (v < V && h < H) {
            const float x = t[threadIdx.x][j];
            const plane_t a = plane_encode(x);
            t0[(int64_t)h * ldv + v] = a;
            t1[(int64_t)h * ldv + v] = plane_encode(x - plane_decode(a));
        }
    }
}

// ---- fused update: theta += lr * grad, NB clamps (src/optimizers.jl:478-485,513-520)
__global__ void k_update(float* __restrict__ theta, const float* __restrict__ grad, int64_t n, float lr, float hi_clamp,
                         int clamp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float t = fmaf(lr, grad[i], theta[i]);
        if (clamp) t = fminf(t, hi_clamp);
        theta[i] = t;
    }
}

// ---- second half of a split-K sweep (latency-bound shapes: one or a few output tiles, long contraction): sums the S
// slabs of raw fp32 partial sums in fixed order, adds the biases, applies the deterministic activation and writes the
// state planes.  One thread per 4 consecutive columns of a row.
__global__ void k_det_finish(const float* __restrict__ ws, int S, int64_t stride, int64_t ldw, EpiParams e, int64_t M, int64_t N) {
    pdl_trigger();
    pdl_wait();
    const int64_t groups = (N + 3) / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * groups) return;
    const int64_t row = i / groups, col0 = (i % groups) * 4;
    const int nvalid = (int)min((int64_t)4, N - col0);
    float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (nvalid == 4) {
        float4 t = *reinterpret_cast<const float4*>(ws + row * ldw + col0);
        for (int sp = 1; sp < S; ++sp) {
            const float4 u = *reinterpret_cast<const float4*>(ws + (int64_t)sp * stride + row * ldw + col0);
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
        for (int j = 0; j < nvalid; ++j)
            for (int sp = 0; sp < S; ++sp) o[j] += ws[(int64_t)sp * stride + row * ldw + col0 + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nvalid) { o[j] = 0.0f; continue; }
        float b = 0.0f;
        if (e.bias0) b += e.bias0[col0 + j];
        if (e.bias1) b += e.bias1[col0 + j];
        const float x = e.factor * fmaf(e.wscale, o[j], b);
        o[j] = e.mode == EPI_SIGM ? sigmf(x) : e.mode == EPI_NBMEAN ? nb_mean(x, e.r[col0 + j]) : x;
    }
    if (e.mode == EPI_SIGM_BERN) {      // Philox-fed Bernoulli draw, the counters of epi_bern4 (o[] holds the inputs)
        const uint4 w = philox_group_rk(e, (uint32_t)row, (uint32_t)(col0 >> 2));
        const float u[4] = {u23(w.x), u23(w.y), u23(w.z), u23(w.w)};
        for (int j = 0; j < nvalid; ++j) e.out.p0[row * e.out.ld + col0 + j] = plane_encode(bern_draw_fast(o[j], u[j]));
        return;
    }
    mat_store4(e.out, row, col0, o, nvalid);
}

// ---- K6: the whole parameter update of one RBM layer in one pass (src/optimizers.jl:478-485,513-520):
//   W += lr dW (NB: W = min(W, 0)),  a += lr da (NB: a = min(a, -1e-10)),  b += lr db,
// and, from the updated values still in registers, every derived operand the sweeps read: the fp16 hi/lo planes of s W in
// both orientations (32 x 32 shared-memory transpose tile) and, for an NB layer, the per-gene sampling constants and
// pass-A thresholds (W <= 0 after the clamp, so the zero-probability bound depends on a and r only).  The block also
// folds max |W| into `wmax` for the next round's scale.  HBM traffic: 8 B read + 12 B written per weight.
// grad: the layer's block {dW [V x H] row-major, da [V], db [H]}.
__global__ void k_update_layer(float* __restrict__ W, float* __restrict__ a, float* __restrict__ b, const float* __restrict__ r,
                               const float* __restrict__ gW, const float* __restrict__ ga, const float* __restrict__ gb, int V, int H,
                               float lr, int nb, float* __restrict__ scale, plane_t* w0, plane_t* w1, int64_t ldh,
                               plane_t* t0, plane_t* t1, int64_t ldv, unsigned int* __restrict__ wmax3, int round, int first, float mu_inv_max,
                               float4* __restrict__ nbc, uint32_t* __restrict__ nbT) {
    pdl_trigger();
    pdl_wait();
    __shared__ float t[32][33];
    __shared__ float red[8];
    // The round's weight scale follows from the max |W| the previous round left in wmax3[round % 3] (weight_scale_exp;
    // every thread derives it: no extra launch, no block-level broadcast).  This round's maximum goes to the next slot;
    // the slot after that is cleared for the round after (three slots: no block reads a slot another block writes).
    const int e2 = weight_scale_exp(__uint_as_float(wmax3[round % 3]));
    const float s = exp2f((float)e2);
    unsigned int* wmax = wmax3 + (round + 1) % 3;
    if (first && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
        scale[0] = s;
        scale[1] = exp2f((float)-e2);
        wmax3[(round + 2) % 3] = 0u;
    }
    const int v0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    float m = 0.0f;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int v = v0 + j, h = h0 + threadIdx.x;
        float x = 0.0f;
        if (v < V && h < H) {
            const int64_t i = (int64_t)v * H + h;
            float w = fmaf(lr, gW[i], W[i]);
            if (nb) w = fminf(w, 0.0f);
            W[i] = w;
            m = fmaxf(m, fabsf(w));
            x = s * w;
            const plane_t hi = plane_encode(x);
            w0[(int64_t)v * ldh + h] = hi;
            w1[(int64_t)v * ldh + h] = plane_encode(x - plane_decode(hi));
        }
        t[j][threadIdx.x] = x;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int h = h0 + j, v = v0 + threadIdx.x;
        if (v < V && h < H) {
            const float x = t[threadIdx.x][j];
            const plane_t hi = plane_encode(x);
            t0[(int64_t)h * ldv + v] = hi;
            t1[(int64_t)h * ldv + v] = plane_encode(x - plane_decode(hi));
        }
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (threadIdx.x == 0) red[threadIdx.y] = m;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int j = 1; j < (int)blockDim.y; ++j) m = fmaxf(m, red[j]);
        if (m > 0.0f) atomicMax(wmax, __float_as_uint(m));
    }
    // biases: one column / row of blocks each
    if (blockIdx.y == 0 && threadIdx.y == 0) {
        const int v = v0 + threadIdx.x;
        if (v < V) {
            float av = fmaf(lr, ga[v], a[v]);
            if (nb) av = fminf(av, -1e-10f);
            a[v] = av;
            if (nb && nbc) {      // as k_refresh_nbc with sum_h max(W, 0) = 0
                const float rr = r[v], c = 1e-6f / rr;
                nbc[v] = make_float4(av, rr, c, rr - 1.0f);
                const float om = -expm1f(av);
                const float p = om / (1.0f + c * om) - 1e-6f;
                const float mu_max = rr * expf(av) / om + 1e-6f;
                float L = 0.0f;
                if (p > 0.0f && mu_max < 0.99f * mu_inv_max) L = expf(rr * logf(p)) * (1.0f - 2e-5f);
                const double mt = ceil(((double)L * 16777216.0 - 1.0) * 0.5);
                nbT[v] = mt >= 8388608.0 ? 0xFFFFFFFFu : (mt <= 0.0 ? 0u : ((uint32_t)mt << 9));
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.y == 1) {
        const int h = h0 + threadIdx.x;
        if (h < H) b[h] = fmaf(lr, gb[h], b[h]);
    }
}
// ---- bias gradients of one RBM in one launch (small batches; the two-pass column sums below serve large ones):
//   da[c] = ps sum_rows v - ns sum_rows vm   (c < V),   db[c - V] = ps sum_rows h - ns sum_rows hm.   Deterministic.
__global__ void k_bias_grad(MatView v, MatView vm, MatView h, MatView hm, float ps, float ns, float* __restrict__ da, float* __restrict__ db) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32][33];
    const int64_t V = v.cols, H = h.cols;
    const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const bool vis = c < V;
    const MatView& P = vis ? v : h;
    const MatView& N = vis ? vm : hm;
    const int64_t cc = vis ? c : c - V;
    float sp = 0.0f, sn = 0.0f;
    if (c < V + H) {      // 32 row lanes: few dependent loads per thread (the launch is latency-bound)
        for (int64_t r = threadIdx.y; r < P.rows; r += 32) sp += mat_load(P, r, cc);
        for (int64_t r = threadIdx.y; r < N.rows; r += 32) sn += mat_load(N, r, cc);
    }
    red[threadIdx.y][threadIdx.x] = ps * sp - ns * sn;
    __syncthreads();
    if (threadIdx.y == 0 && c < V + H) {
        float tot = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) tot += red[j][threadIdx.x];
        (vis ? da : db)[cc] = tot;
    }
}

// ---- column sums of a state matrix, deterministic (no float atomics):
// pass 1: partial[slab][c] = sum over the slab's rows; pass 2: out[c] (+)= scale * sum_slabs (fixed order)
__global__ void k_colsum_partial(MatView m, float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    // block = 32 chunks of 8 columns x 8 row lanes (16-byte plane loads: rows are padded to multiples of 64 columns);
    // grid.x over 256-column blocks, grid.y over row slabs
    __shared__ float red[8][32][9];
    const int64_t c0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 8;
    float s[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    auto add8 = [&](const plane_t* p) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(h[j]);
            s[2 * j] += t.x;
            s[2 * j + 1] += t.y;
        }
    };
    if (c0 < m.ld) {
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < m.rows; r += (int64_t)gridDim.y * 8) {
            const int64_t i = r * m.ld + c0;
            if (m.f) {
                const float4 x = *reinterpret_cast<const float4*>(m.f + i), y = *reinterpret_cast<const float4*>(m.f + i + 4);
                s[0] += x.x; s[1] += x.y; s[2] += x.z; s[3] += x.w; s[4] += y.x; s[5] += y.y; s[6] += y.z; s[7] += y.w;
            } else {
                add8(m.p0 + i);
                // hi count plane only where the row tile uses it (half the HBM reads)
                if (m.p1 && !(m.kind == MAT_COUNT && m.hiflag && !(m.hiflag[r >> 7] & 1u))) add8(m.p1 + i);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = s[j];
    __syncthreads();
    // 256 columns of the block, one per thread: fixed summation order over the 8 row lanes
    const int t = threadIdx.y * 32 + threadIdx.x;
    const int64_t c = (int64_t)blockIdx.x * 256 + t;
    if (c < m.cols) {
        float tot = 0.0f;
        for (int j = 0; j < 8; ++j) tot += red[j][t >> 3][t & 7];
        partial[(int64_t)blockIdx.y * m.cols + c] = tot;
    }
}
__global__ void k_colsum_final(const float* __restrict__ partial, int nslabs, int64_t cols, float scale,
                               float* __restrict__ out, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float tot = 0.0f;
    for (int s = 0; s < nslabs; ++s) tot += partial[(int64_t)s * cols + c];
    out[c] = (accumulate ? out[c] : 0.0f) + scale * tot;
}

// per-gene constants of the fast NB epilogue: {a, r, 1e-6/r, r-1}, and a lower bound L on the zero
// probability P0 = p^r valid for every hidden state h in [0,1]^H and every temperature in [0,1]:
// x = a + t * sum_h W h <= a + sum_h max(W, 0) =: xmax, and P0 is decreasing in x.  (The reference
// clamps W <= 0 for NB layers, src/optimizers.jl:482, so xmax = a there.)  u < L  =>  the draw is 0.
// The shortcut only holds where the draw is the single-uniform inversion: a gene whose mean can exceed
// mu_inv_max (gamma-Poisson branch, which does not consume u) gets L = 0.
__global__ void k_refresh_nbc(const float* __restrict__ a, const float* __restrict__ r, const float* __restrict__ W, int V,
                              int H, float mu_inv_max, float4* __restrict__ nbc, uint32_t* __restrict__ nbT,
                              unsigned int* __restrict__ invalid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float rr = r[i], c = 1e-6f / rr;
    nbc[i] = make_float4(a[i], rr, c, rr - 1.0f);
    float pos = 0.0f;
    for (int h = 0; h < H; ++h) pos += fmaxf(W[(int64_t)i * H + h], 0.0f);
    const float xmax = a[i] + pos;
    if (!(xmax < 0.0f) && invalid) *invalid = 1u;               // may be mapped host memory: plain store   // NB validity: x = a + W h must stay negative
    float L = 0.0f;
    if (xmax < 0.0f) {
        const float om = -expm1f(xmax);
        const float p = om / (1.0f + c * om) - 1e-6f;          // largest pshift in use
        const float mu_max = rr * expf(xmax) / om + 1e-6f;      // the mean is increasing in x
        if (p > 0.0f && mu_max < 0.99f * mu_inv_max) L = expf(rr * logf(p)) * (1.0f - 2e-5f);   // margin: fast-math P0 of pass B
    }
    // u = (m + 1/2) 2^-23 with m = word >> 9 is exact in fp32, so  u < L  <=>  2m + 1 < L 2^24  <=>  m < ceil((L 2^24 - 1) / 2)
    const double mt = ceil(((double)L * 16777216.0 - 1.0) * 0.5);
    nbT[i] = mt >= 8388608.0 ? 0xFFFFFFFFu : (mt <= 0.0 ? 0u : ((uint32_t)mt << 9));
}

__global__ void k_fill_nbc_pad(float4* nbc, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nbc[i] = make_float4(-1.0f, 1.0f, 1e-6f, 0.0f);
}

// before an NB sampling launch into `flags`' matrix: tiles whose hi plane was in use (or marked) must be zero-filled
// by this launch (-> 2); tiles that were zero-filled by the previous launch and got no new entry are clean (-> 0)
__global__ void k_flags_advance(uint32_t* __restrict__ flags, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (flags[i] & 1u) ? 2u : 0u;
}

__global__ void k_fill_f32(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- row gather: dst[i, :] = src[idx[i], :]   (minibatch, src/rbmtraining.jl:544)
// does any row tile of a COUNT matrix use its hi plane?  (one block; out = 3 or 0, the value k_gather_rows gives its tiles)
__global__ void k_any_flag(const uint32_t* __restrict__ flags, int64_t n, uint32_t* __restrict__ out) {
    uint32_t any = 0u;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) any |= flags[i];
    any = __syncthreads_or((int)(any & 1u));
    if (threadIdx.x == 0) *out = any ? 3u : 0u;
}
// any_hi: device word written by k_any_flag for `src` (nullptr: assume counts >= 2048 may occur)
// One launch gathers a run of minibatches: dst row (j * Bpad + i) = src row idx[j * B + i]; Bpad (a multiple of 128)
// keeps every minibatch on its own hi-plane flag tiles.  Rows are copied 16 bytes at a time (ld is a multiple of 64).
__global__ void k_gather_rows(MatView src, const int32_t* __restrict__ idx, MatView dst, const uint32_t* __restrict__ any_hi, int B, int Bpad) {
    pdl_trigger();
    pdl_wait();
    const int64_t j = blockIdx.x / B, i = blockIdx.x % B;
    const int64_t r = j * Bpad + i;
    const int64_t s = idx[blockIdx.x];
    if (dst.hiflag && threadIdx.x == 0 && (i & 127) == 0) dst.hiflag[r >> 7] = any_hi ? *any_hi : 3u;
    if (src.ld == dst.ld && (!dst.f || src.f)) {
        const int64_t n16 = dst.ld / 8;       // 16-byte chunks of one plane row
        const uint4* s0 = reinterpret_cast<const uint4*>(src.p0 + s * src.ld);
        uint4* d0 = reinterpret_cast<uint4*>(dst.p0 + r * dst.ld);
        for (int64_t c = threadIdx.x; c < n16; c += blockDim.x) d0[c] = s0[c];
        if (dst.p1) {
            uint4* d1 = reinterpret_cast<uint4*>(dst.p1 + r * dst.ld);
            if (src.p1) {
                const uint4* s1 = reinterpret_cast<const uint4*>(src.p1 + s * src.ld);
                for (int64_t c = threadIdx.x; c < n16; c += blockDim.x) d1[c] = s1[c];
            } else {
                for (int64_t c = threadIdx.x; c < n16; c += blockDim.x) d1[c] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (dst.f) {
            const float4* sf = reinterpret_cast<const float4*>(src.f + s * src.ld);
            float4* df = reinterpret_cast<float4*>(dst.f + r * dst.ld);
            for (int64_t c = threadIdx.x; c < dst.ld / 4; c += blockDim.x) df[c] = sf[c];
        }
        return;
    }
    for (int64_t c = threadIdx.x; c < dst.ld; c += blockDim.x) {
        dst.p0[r * dst.ld + c] = src.p0[s * src.ld + c];
        if (dst.p1) dst.p1[r * dst.ld + c] = src.p1 ? src.p1[s * src.ld + c] : __ushort_as_half(0);
        if (dst.f) dst.f[r * dst.ld + c] = src.f ? src.f[s * src.ld + c] : (c < src.cols ? mat_load(src, s, c) : 0.0f);
    }
}

__global__ void k_copy_rows(MatView src, MatView dst, int64_t rows) {
    pdl_trigger();
    pdl_wait();
    const int64_t r = blockIdx.x;
    if (r >= rows) return;
    for (int64_t c = threadIdx.x; c < dst.ld; c += blockDim.x) {
        dst.p0[r * dst.ld + c] = src.p0[r * src.ld + c];
        if (dst.p1) dst.p1[r * dst.ld + c] = src.p1 ? src.p1[r * src.ld + c] : __ushort_as_half(0);
        if (dst.f) dst.f[r * dst.ld + c] = src.f ? src.f[r * src.ld + c] : (c < src.cols ? mat_load(src, r, c) : 0.0f);
    }
}

// ---- in-place Bernoulli sampling of a matrix of probabilities (bernoulli!,
// src/gibbssampling.jl:40-45) -> dst (BINARY or REAL kind)
struct DrawAddr {
    uint32_t k0, k1, tick, stream, row_offset;
    const float* uniforms;
    int64_t ldu;
};
__global__ void k_bernoulli(MatView prob, MatView dst, DrawAddr d) {
    pdl_trigger();
    pdl_wait();
    const int64_t groups = (prob.cols + 3) / 4;
    const int64_t n = prob.rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups, gidx = i % groups, c0 = gidx * 4;
        const int nvalid = (int)min((int64_t)4, prob.cols - c0);
        float u[4];
        if (d.uniforms) {
            for (int j = 0; j < 4; ++j) u[j] = j < nvalid ? d.uniforms[r * d.ldu + c0 + j] : 1.0f;
        } else {
            const uint4 w = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)r + d.row_offset, d.tick, d.stream), d.k0, d.k1);
            u[0] = u23(w.x); u[1] = u23(w.y); u[2] = u23(w.z); u[3] = u23(w.w);
        }
        float o[4] = {0, 0, 0, 0};
        for (int j = 0; j < nvalid; ++j) o[j] = (u[j] < mat_load(prob, r, c0 + j)) ? 1.0f : 0.0f;
        mat_store4(dst, r, c0, o, nvalid);
    }
}

// ---- particle / chain-state initialisation
// mode 0: fair coin (u < 0.5); 1: Bernoulli(sigm(bias[c])); 2: raw uniform (PCD chain state,
// src/rbmtraining.jl:141); 3: NB visible init (src/gibbssampling.jl:512-531)
__global__ void k_init(MatView dst, int mode, const float* __restrict__ bias, const float* __restrict__ r, uint32_t k0,
                       uint32_t k1, uint32_t layer, uint32_t row_offset) {
    const int64_t groups = (dst.cols + 3) / 4;
    const int64_t n = dst.rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / groups, gidx = i % groups, c0 = gidx * 4;
        const int nvalid = (int)min((int64_t)4, dst.cols - c0);
        const uint4 w = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)row + row_offset, 0u, stream_id(layer, P_INIT)), k0, k1);
        const float u[4] = {u23(w.x), u23(w.y), u23(w.z), u23(w.w)};
        float o[4] = {0, 0, 0, 0};
        if (mode == 3) {
            const uint4 w2 = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)row + row_offset, 0u, stream_id(layer, P_INIT2)), k0, k1);
            const float u2[4] = {u23(w2.x), u23(w2.y), u23(w2.z), u23(w2.w)};
            const uint32_t wj[4] = {w.x, w.y, w.z, w.w};
            for (int j = 0; j < nvalid; ++j) {
                // stage 1, NB(1, 1/2) by inversion: P(k) = 2^-(k+1), every chop-down subtraction is exact on the
                // 24-bit uniform (m + 1/2) 2^-23, so k = number of leading one bits of m (24 if all 23 are set)
                const uint32_t t = ~(wj[j] >> 9) & 0x7FFFFFu;
                const float v0 = t ? (float)(__clz((int)t) - 9) : 24.0f;
                // stage 2, NB(r, p = r/(v0 + r) - 1e-6) by inversion (same recurrence as nb_inversion, MUFU math)
                const float rr = r[c0 + j];
                const float p = rr * fast_rcp(v0 + rr) - 1e-6f, q = 1.0f - p, qr1 = q * (rr - 1.0f);
                float P = fast_ex2(rr * fast_lg2(p)), uu = u2[j], k = 0.0f, rk = 1.0f;
                while (uu >= P && P > PMIN_F) {
                    uu -= P;
                    k += 1.0f;
                    P *= fmaf(qr1, rk, q);
                    rk = fast_rcp(k + 1.0f);
                }
                o[j] = k;
                if (k >= COUNT_BASE && dst.hiflag) dst.hiflag[row / 128] = 3u;
            }
        } else {
            for (int j = 0; j < nvalid; ++j) {
                if (mode == 0) o[j] = u[j] < 0.5f ? 1.0f : 0.0f;
                else if (mode == 1) o[j] = u[j] < sigmf(bias[c0 + j]) ? 1.0f : 0.0f;
                else o[j] = u[j];
            }
        }
        mat_store4(dst, row, c0, o, nvalid);
    }
}

// ---- per-column mean / variance / zero fraction (the paper's sample-quality
// statistics, src/NegativeBinomialRBMPlots.jl:148-182,428-462)
__global__ void k_colstats(MatView m, double* __restrict__ sum, double* __restrict__ sumsq, double* __restrict__ zeros) {
    __shared__ double rs[8][33], rq[8][33], rz[8][33];
    const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    double s = 0, q = 0, z = 0;
    if (c < m.cols)
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < m.rows; r += (int64_t)gridDim.y * 8) {
            const double v = mat_load(m, r, c);
            s += v; q += v * v; z += (v == 0.0);
        }
    rs[threadIdx.y][threadIdx.x] = s; rq[threadIdx.y][threadIdx.x] = q; rz[threadIdx.y][threadIdx.x] = z;
    __syncthreads();
    if (threadIdx.y == 0 && c < m.cols) {
        for (int j = 1; j < 8; ++j) { s += rs[j][threadIdx.x]; q += rq[j][threadIdx.x]; z += rz[j][threadIdx.x]; }
        atomicAdd(&sum[c], s); atomicAdd(&sumsq[c], q); atomicAdd(&zeros[c], z);
    }
}

// ---- conditional Gibbs: reset clamped visible entries to their initial values
__global__ void k_restore_cols(MatView dst, MatView orig, const int32_t* __restrict__ cols, int ncols) {
    const int64_t n = dst.rows * (int64_t)ncols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncols, c = cols[i % ncols], j = r * dst.ld + c;
        dst.p0[j] = orig.p0[j];
        if (dst.p1) dst.p1[j] = orig.p1[j];
        if (dst.hiflag && dst.p1 && (__half_as_ushort(orig.p1[j]) & 0x7FFFu) != 0) dst.hiflag[r / 128] = 3u;
    }
}
__global__ void k_restore_mask(MatView dst, MatView orig, const uint8_t* __restrict__ mask) {   // mask row-major [rows x cols]
    const int64_t n = dst.rows * dst.cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!mask[i]) continue;
        const int64_t r = i / dst.cols, j = r * dst.ld + i % dst.cols;
        dst.p0[j] = orig.p0[j];
        if (dst.p1) dst.p1[j] = orig.p1[j];
        if (dst.hiflag && dst.p1 && (__half_as_ushort(orig.p1[j]) & 0x7FFFu) != 0) dst.hiflag[r / 128] = 3u;
    }
}

// ---- zero-inflation mask (K12): v *= 1 - bernoulli(logistic(shape * (x' - mid)))
__global__ void k_dropout(MatView v, MatView x, const float* __restrict__ mid, const float* __restrict__ shape, int nparams,
                          int log1p_form, DrawAddr d) {
    const int64_t groups = (v.cols + 3) / 4;
    const int64_t n = v.rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups, gidx = i % groups, c0 = gidx * 4;
        const int nvalid = (int)min((int64_t)4, v.cols - c0);
        const uint4 w = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)r + d.row_offset, d.tick, d.stream), d.k0, d.k1);
        const float u[4] = {u23(w.x), u23(w.y), u23(w.z), u23(w.w)};
        float o[4] = {0, 0, 0, 0};
        for (int j = 0; j < nvalid; ++j) {
            const int pi = nparams == 1 ? 0 : (int)(c0 + j);
            float xt = mat_load(x, r, c0 + j);
            if (log1p_form) xt = log1pf(xt);
            const float p = 1.0f / (1.0f + expf(-shape[pi] * (xt - mid[pi])));
            o[j] = (u[j] < p) ? 0.0f : mat_load(v, r, c0 + j);
        }
        mat_store4(v, r, c0, o, nvalid);
    }
}

// ---- NB inverse dispersion by penalised Fisher scoring (mle_for_theta), one block per gene, fp64
__device__ inline double digamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}
__device__ inline double trigamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    const double f = 1.0 / (x * x), ix = 1.0 / x;
    return r + ix + 0.5 * f + ix * f * (1.0 / 6.0 - f * (1.0 / 30.0 - f * (1.0 / 42.0 - f * (1.0 / 30.0))));
}
__device__ inline double block_sum(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}
__global__ void __launch_bounds__(256) k_mle_theta(MatView x, MatView mu, double lam, int maxiter, double tol, double* __restrict__ theta) {
    __shared__ double red[8];
    const int64_t g = blockIdx.x, n = x.rows;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
        const double y = mat_load(x, r, g), m = mat_load(mu, r, g);
        const double t = y / m - 1.0;
        s += t * t;
    }
    double th = (double)n / block_sum(s, red);
    for (int it = 0; it < maxiter; ++it) {
        th = fmax(th, 0.01);
        double score = 0.0, fisher = 0.0;
        const double dg = digamma_d(th), tg = trigamma_d(th), lth = log(th);
        for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
            const double y = mat_load(x, r, g), m = mat_load(mu, r, g);
            score += -((y + th) / (m + th) + log(m + th) - 1.0 - lth - digamma_d(th + y) + dg);
            fisher += -(y + th) / ((m + th) * (m + th)) + 2.0 / (m + th) - 1.0 / th - trigamma_d(th + y) + tg;
        }
        score = block_sum(score, red);
        fisher = block_sum(fisher, red);
        const double delta = (score + lam * 2.0 / (th * th * th)) / (fisher + lam * 6.0 / (th * th * th * th));
        if (fabs(delta) <= tol) break;     // block-uniform: every thread holds the same sums
        th += delta;
    }
    if (threadIdx.x == 0) theta[g] = th;
}

}  // namespace scdbm
