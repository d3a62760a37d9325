// Sparse data path: scRNA-seq count matrices are 80-92 % zeros, so the host <-> device
// boundary can carry compressed rows/columns instead of dense Float64 (16 GB for 10^6 x 2000 cells), and
// synthetic "10x-like" data can be generated where it is consumed.  The device state stays the dense fp16
// plane layout the tcgen05 sweeps contract over; these kernels convert at the boundary.  All HBM-bound.
#pragma once
#include "elementwise.cuh"

namespace scdbm {

// ---- compressed -> planes.  `major` = 0: CSR (ptr over rows, idx = column); 1: CSC (ptr over columns,
// idx = row; Julia's SparseMatrixCSC).  One warp per compressed row/column; the planes were zeroed before.
// err: bit 0 = index out of range, bit 1 = value not an integer in [0, 65535] (COUNT) / not 0 or 1 (BINARY).
__global__ void k_scatter_compressed(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                     const float* __restrict__ val, int64_t nmajor, int major, MatView dst,
                                     unsigned int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); s < nmajor; s += warps) {
        const int64_t b = ptr[s], e = ptr[s + 1];
        for (int64_t i = b + lane; i < e; i += 32) {
            const int64_t o = idx[i];
            const int64_t r = major ? o : s, c = major ? s : o;
            const float v = val[i];
            if (o < 0 || r >= dst.rows || c >= dst.cols) { atomicOr(err, 1u); continue; }
            const bool ok = dst.kind == MAT_COUNT ? (v >= 0.0f && v <= 65535.0f && v == floorf(v))
                          : dst.kind == MAT_BINARY ? (v == 0.0f || v == 1.0f) : isfinite(v);
            if (!ok) { atomicOr(err, 2u); continue; }
            plane_t a, hb;
            split_value(dst.kind, v, a, hb);
            const int64_t j = r * dst.ld + c;
            dst.p0[j] = a;
            if (dst.p1) dst.p1[j] = hb;
            if (dst.f) dst.f[j] = v;
            if (dst.hiflag && (__half_as_ushort(hb) & 0x7FFFu) != 0) dst.hiflag[r / 128] = 3u;
        }
    }
}

// ---- planes -> CSR.  Pass 1: non-zeros per row (one warp per row).
__global__ void k_row_nnz(MatView src, int64_t* __restrict__ nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < src.rows; r += warps) {
        int n = 0;
        for (int64_t c = lane; c < src.cols; c += 32) n += mat_load(src, r, c) != 0.0f;
        for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
        if (lane == 0) nnz[r] = n;
    }
}

// exclusive scan of int64 counts, three small kernels (1024 elements per block)
__global__ void k_scan_block(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ block_total) {
    __shared__ int64_t ws[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t v = i < n ? in[i] : 0;
    int64_t x = v;
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
        int64_t t = ws[lane];
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t += y;
        }
        ws[lane] = t;
    }
    __syncthreads();
    const int64_t incl = x + (w ? ws[w - 1] : 0);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == 1023) block_total[blockIdx.x] = incl;
}
__global__ void k_scan_totals(int64_t* __restrict__ block_total, int64_t nblocks, int64_t* __restrict__ grand_total) {
    // single thread block, sequential over chunks of 1024 block totals (10^9 rows = 10^6 totals = 1000 chunks)
    __shared__ int64_t ws[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int64_t v = i < nblocks ? block_total[i] : 0;
        int64_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            int64_t t = ws[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, t, off);
                if (lane >= off) t += y;
            }
            ws[lane] = t;
        }
        __syncthreads();
        const int64_t incl = x + (w ? ws[w - 1] : 0) + carry;
        if (i < nblocks) block_total[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void k_scan_add(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ block_total, const int64_t* __restrict__ grand_total) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_total[blockIdx.x];
    if (i == 0) out[n] = *grand_total;
}

// Pass 2: one warp per row appends (column, value) in column order (ballot compaction).
template <class Index>
__global__ void k_csr_fill(MatView src, const int64_t* __restrict__ indptr, Index* __restrict__ indices, uint16_t* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < src.rows; r += warps) {
        int64_t pos = indptr[r];
        for (int64_t c0 = 0; c0 < src.cols; c0 += 32) {
            const int64_t c = c0 + lane;
            const float v = c < src.cols ? mat_load(src, r, c) : 0.0f;
            const uint32_t mask = __ballot_sync(0xffffffffu, v != 0.0f);
            if (v != 0.0f) {
                const int64_t j = pos + __popc(mask & lt);
                indices[j] = (Index)c;
                values[j] = (uint16_t)fminf(fmaxf(v, 0.0f), 65535.0f);
            }
            pos += __popc(mask);
        }
    }
}

// ---- synthetic "10x-like" counts generated on the device:
//   size factor s_c = exp(sigma z_c), z_c = sqrt(-2 ln u1) cos(2 pi u2)   (Philox element (col 0, row c), stream (0, P_SIZE), tick 0)
//   x_cg ~ NB(r_g, p = r_g / (s_c m_g + r_g))                              (uniform of stream (0, P_DATA), tick 0; nb_draw contract)
__global__ void k_generate_counts(MatView dst, const float* __restrict__ gene_mean, const float* __restrict__ r, float sigma,
                                  float mu_inv_max, uint32_t k0, uint32_t k1, uint32_t row_offset) {
    const int64_t groups = (dst.cols + 3) / 4;
    const int64_t n = dst.rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / groups, gidx = i % groups, c0 = gidx * 4;
        const int nvalid = (int)min((int64_t)4, dst.cols - c0);
        const uint32_t grow = (uint32_t)row + row_offset;
        const uint4 ws = philox_element(k0, k1, grow, 0u, 0u, stream_id(0, P_SIZE));
        const float z = sqrtf(-2.0f * logf(u23(ws.x))) * cospif(2.0f * u23(ws.y));
        const float s = expf(sigma * z);
        const uint4 w = philox4x32_10(make_uint4((uint32_t)gidx, grow, 0u, stream_id(0, P_DATA)), k0, k1);
        const float u[4] = {u23(w.x), u23(w.y), u23(w.z), u23(w.w)};
        float o[4] = {0, 0, 0, 0};
        bool big = false;
        for (int j = 0; j < nvalid; ++j) {
            const RngAddr g{k0, k1, grow, (uint32_t)(c0 + j), 0u, 0u};
            o[j] = nb_draw(s * gene_mean[c0 + j], r[c0 + j], u[j], 0.0f, mu_inv_max, g);
            big = big || o[j] >= COUNT_BASE;
        }
        mat_store4(dst, row, c0, o, nvalid);
        if (big && dst.hiflag) dst.hiflag[row / 128] = 3u;
    }
}

}  // namespace scdbm
