// The fused epilogue shared by the SIMT and tcgen05 sweep GEMMs.
// Processes 4 consecutive columns (col0 % 4 == 0) of one row: adds biases,
// applies the factor, activation and (optionally) the Philox draw, and writes
// the result as bf16 operand planes for the next contraction.
#pragma once
#include "samplers.cuh"

namespace scdbm {

// 16-bit plane word <-> float (IEEE fp16, saturating so that no plane ever holds Inf)
__device__ __forceinline__ float plane_decode(plane_t w) { return __half2float(w); }
__device__ __forceinline__ plane_t plane_encode(float v) {
    return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
}

__device__ __forceinline__ float mat_load(const MatView& m, int64_t row, int64_t col) {
    const int64_t i = row * m.ld + col;
    if (m.f) return m.f[i];
    float v = plane_decode(m.p0[i]);
    if (m.p1) v += plane_decode(m.p1[i]);
    return v;
}

__device__ __forceinline__ void split_value(int kind, float v, plane_t& a, plane_t& b) {
    if (kind == MAT_COUNT) {
        const float hi = floorf(v * (1.0f / COUNT_BASE)) * COUNT_BASE;
        a = plane_encode(v - hi);
        b = plane_encode(hi);
    } else {
        a = plane_encode(v);
        b = plane_encode(v - plane_decode(a));
    }
}

// write up to 4 consecutive values; 8-byte vector store when all 4 are in range
__device__ __forceinline__ void mat_store4(const MatView& m, int64_t row, int64_t col0, const float (&v)[4], int nvalid) {
    plane_t a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_value(m.kind, v[j], a[j], b[j]);
    const int64_t i = row * m.ld + col0;
    if (m.f) {
        if (nvalid == 4) *reinterpret_cast<float4*>(m.f + i) = make_float4(v[0], v[1], v[2], v[3]);
        else for (int j = 0; j < nvalid; ++j) m.f[i + j] = v[j];
    }
    if (nvalid == 4) {
        uint2 pa, pb;
        pa.x = (uint32_t)__half_as_ushort(a[0]) | ((uint32_t)__half_as_ushort(a[1]) << 16);
        pa.y = (uint32_t)__half_as_ushort(a[2]) | ((uint32_t)__half_as_ushort(a[3]) << 16);
        *reinterpret_cast<uint2*>(m.p0 + i) = pa;
        if (m.p1) {
            pb.x = (uint32_t)__half_as_ushort(b[0]) | ((uint32_t)__half_as_ushort(b[1]) << 16);
            pb.y = (uint32_t)__half_as_ushort(b[2]) | ((uint32_t)__half_as_ushort(b[3]) << 16);
            *reinterpret_cast<uint2*>(m.p1 + i) = pb;
        }
    } else {
        for (int j = 0; j < nvalid; ++j) {
            m.p0[i + j] = a[j];
            if (m.p1) m.p1[i + j] = b[j];
        }
    }
}

// log((1+e^{y1})/(1+e^{y2})) without cancellation: log1p(sigm(y2) * expm1(y1-y2))
// The transcendental-heavy bodies of the generic epilogue are out of line: inlined four times per column group they
// made its loop body 52 KB of SASS -- larger than the 32 KB instruction cache level, i.e. an instruction-fetch stall on
// every group (1.7 us per 4-column group in a latency-bound launch).
__device__ __noinline__ float softplus_diff(float y2, float dy) { return log1pf(sigmf(y2) * expm1f(dy)); }
__device__ __noinline__ float nb_mean_ool(float x, float r) { return nb_mean(x, r); }
__device__ __noinline__ float nb_draw_ool(float x, float r, float u, float pshift, float mu_inv_max, RngAddr g) {
    return nb_draw(nb_mean(x, r), r, u, pshift, mu_inv_max, g);
}

// Returns the AIS partial row sum (0 for other modes).  `row` is local to the
// matrix; nvalid = number of in-range columns among col0..col0+3.
// pre_bias / pre_r: the four columns' summed biases and inverse dispersions when the caller has them in registers
// already (tensor-core engine: loaded once per tile and broadcast by shuffle); nullptr: loaded here.
__device__ __forceinline__ float epi_apply4(const EpiParams& e, int64_t row, int64_t col0, const float (&acc)[4],
                                            int nvalid, float& delta_local, const float* pre_bias = nullptr, const float* pre_r = nullptr) {
    float x[4], out[4], bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float b = 0.0f;
        if (pre_bias) b = pre_bias[j];
        else if (j < nvalid) {
            if (e.bias0) b += e.bias0[col0 + j];
            if (e.bias1) b += e.bias1[col0 + j];
        }
        bias[j] = b;
        if (e.addend && j < nvalid) {
            float t = e.addend[row * e.ld_add + col0 + j];
#pragma unroll 1
            for (int sp = 1; sp < e.addend_splits; ++sp) t += e.addend[(int64_t)sp * e.addend_stride + row * e.ld_add + col0 + j];   // split-K slabs, fixed order
            b += e.addend_scaled ? t * e.wscale : t;
        }
        x[j] = e.factor * (e.wscale * acc[j] + b);
    }
    if (e.mode == EPI_AIS_RATIO) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nvalid) s += softplus_diff(e.t2 * acc[j] + bias[j], (e.t1 - e.t2) * acc[j]);
        }
        return s;
    }
    if (e.mode == EPI_RAW_F32) {
        for (int j = 0; j < nvalid; ++j) e.out_f32[row * e.ldo + col0 + j] = acc[j];
        return 0.0f;
    }
    uint4 w = make_uint4(0, 0, 0, 0);
    float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (e.mode == EPI_SIGM_BERN || e.mode == EPI_NB_DRAW) {
        if (e.uniforms) {
            for (int j = 0; j < nvalid; ++j) u[j] = e.uniforms[row * e.ldu + col0 + j];
        } else {
            w = philox_group(e, (uint32_t)row, (uint32_t)(col0 >> 2));
            u[0] = u23(w.x); u[1] = u23(w.y); u[2] = u23(w.z); u[3] = u23(w.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float o = 0.0f;
        if (j < nvalid) {
            switch (e.mode) {
                case EPI_INPUT: o = x[j]; break;
                case EPI_SIGM: o = sigmf(x[j]); break;
                case EPI_SIGM_BERN: o = (u[j] < sigmf(x[j])) ? 1.0f : 0.0f; break;
                case EPI_NBMEAN: o = nb_mean_ool(x[j], pre_r ? pre_r[j] : e.r[col0 + j]); break;
                case EPI_NB_DRAW: {
                    const float r = pre_r ? pre_r[j] : e.r[col0 + j];
                    const RngAddr g{e.k0, e.k1, (uint32_t)row + e.row_offset, (uint32_t)(col0 + j), e.tick,
                                    e.stream & 0xFFu};
                    o = nb_draw_ool(x[j], r, u[j], e.pshift, e.mu_inv_max, g);
                } break;
                default: break;
            }
            if (e.prev.p0) delta_local = fmaxf(delta_local, fabsf(mat_load(e.prev, row, col0 + j) - o));
        }
        out[j] = o;
    }
    mat_store4(e.out, row, col0, out, nvalid);
    return 0.0f;
}

// ---- the deterministic epilogue class of the tensor-core engine (EK = 3): inputs, sigmoid potentials, NB means and raw
// fp32 accumulators, with the optional mean-field extras (hoisted addend, previous value -> max |change|).  No draws,
// no AIS: its loop body is a few hundred instructions, so latency-bound launches (training minibatches) and the
// mean-field / positive-phase sweeps do not stall on instruction fetch the way the all-modes epilogue does.
__device__ __forceinline__ void epi_det4(const EpiParams& e, int64_t row, int64_t col0, const float (&acc)[4], int nvalid,
                                         float& delta_local, const float (&bias)[4], const float (&r4)[4]) {
    float o[4], add[4] = {0.0f, 0.0f, 0.0f, 0.0f}, old[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    // mean-field extras first, as two 16-byte loads in flight together (rows are 16-byte aligned: ld % 4 == 0)
    const bool vec = nvalid == 4;
    if (e.addend) {
        const float* a = e.addend + row * e.ld_add + col0;
        if (vec && (e.ld_add & 3) == 0 && (reinterpret_cast<uintptr_t>(e.addend) & 15u) == 0u) {
            const float4 t = *reinterpret_cast<const float4*>(a);
            add[0] = t.x; add[1] = t.y; add[2] = t.z; add[3] = t.w;
        } else {
            for (int j = 0; j < nvalid; ++j) add[j] = a[j];
        }
#pragma unroll 1
        for (int sp = 1; sp < e.addend_splits; ++sp)
            for (int j = 0; j < nvalid; ++j) add[j] += a[(int64_t)sp * e.addend_stride + j];
        if (e.addend_scaled)
            for (int j = 0; j < 4; ++j) add[j] *= e.wscale;
    }
    if (e.prev.p0) {
        if (e.prev.f && vec) {
            const float4 t = *reinterpret_cast<const float4*>(e.prev.f + row * e.prev.ld + col0);
            old[0] = t.x; old[1] = t.y; old[2] = t.z; old[3] = t.w;
        } else {
            for (int j = 0; j < nvalid; ++j) old[j] = mat_load(e.prev, row, col0 + j);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x = e.factor * fmaf(e.wscale, acc[j], bias[j] + add[j]);
        o[j] = e.mode == EPI_SIGM ? sigmf(x) : e.mode == EPI_NBMEAN ? nb_mean(x, r4[j]) : x;
        if (j >= nvalid) o[j] = 0.0f;
    }
    if (e.mode == EPI_RAW_F32) {
        float* d = e.out_f32 + row * e.ldo + col0;
        if (nvalid == 4 && (reinterpret_cast<uintptr_t>(d) & 15u) == 0u) *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else for (int j = 0; j < nvalid; ++j) d[j] = acc[j];
        return;
    }
    if (e.prev.p0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nvalid) delta_local = fmaxf(delta_local, fabsf(old[j] - o[j]));
    }
    mat_store4(e.out, row, col0, o, nvalid);
}

// The same class for a TRANSPOSED accumulator (two-SM kernel, lane = output unit, the 4 values are 4 consecutive rows of
// one column): every global access of a warp is then 32 consecutive units of one row -- coalesced 128-byte segments --
// where the row-per-lane form touches 32 different lines per instruction (the mean-field sweeps at 100 000 cells were
// bound by that: 129 M sector transactions for 66 M sectors of data, ncu capture prof_r02_c4mf).
__device__ __forceinline__ void epi_det4t(const EpiParams& e, int64_t row0, int nrows, int64_t col, const float (&acc)[4],
                                          float& delta_local, float bias, float r) {
    float add[4] = {0.0f, 0.0f, 0.0f, 0.0f}, old[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (e.mode == EPI_RAW_F32) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nrows) e.out_f32[(row0 + j) * e.ldo + col] = acc[j];
        return;
    }
    if (e.addend) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nrows) add[j] = e.addend[(row0 + j) * e.ld_add + col];
#pragma unroll 1
        for (int sp = 1; sp < e.addend_splits; ++sp)
            for (int j = 0; j < nrows; ++j) add[j] += e.addend[(int64_t)sp * e.addend_stride + (row0 + j) * e.ld_add + col];
        if (e.addend_scaled)
            for (int j = 0; j < 4; ++j) add[j] *= e.wscale;
    }
    if (e.prev.p0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nrows) old[j] = e.prev.f ? e.prev.f[(row0 + j) * e.prev.ld + col] : mat_load(e.prev, row0 + j, col);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nrows) continue;
        const float x = e.factor * fmaf(e.wscale, acc[j], bias + add[j]);
        const float o = e.mode == EPI_SIGM ? sigmf(x) : e.mode == EPI_NBMEAN ? nb_mean(x, r) : x;
        if (e.prev.p0) delta_local = fmaxf(delta_local, fabsf(old[j] - o));
        plane_t a, b;
        split_value(e.out.kind, o, a, b);
        const int64_t i = (row0 + j) * e.out.ld + col;
        if (e.out.f) e.out.f[i] = o;
        e.out.p0[i] = a;
        if (e.out.p1) e.out.p1[i] = b;
    }
}

// ---- specialised sampling epilogues (tensor-core engine, Philox-fed): 4 consecutive columns
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// ---- two-pass NB epilogue (tensor-core engine).  Pass B: one queued element per lane.
__device__ __forceinline__ float nb_entry_draw(const EpiParams& e, float acc, float u, uint32_t grow, uint32_t col) {
    const float4 c = __ldg(e.nbc + col);                             // {a, r, 1e-6/r, r-1}
    const float x = fmaf(e.wscale, acc, c.x);
    const float E = fast_ex2(x * 1.4426950408889634f);
    float om = 1.0f - E;
    // x -> 0-: 1 - e^x cancels; -expm1(x) = -x (1 + x/2 + x^2/6 + x^3/24 + x^4/120 + x^5/720), |err| < 2e-9 on (-1/8, 0)
    if (x > -0.125f)
        om = -x * fmaf(x, fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 720.0f, 1.0f / 120.0f), 1.0f / 24.0f), 1.0f / 6.0f), 0.5f), 1.0f);
    if (c.y * E > (e.mu_inv_max - 1e-6f) * om) {                     // mean above mu_inv_max: gamma-Poisson, out of line
        const RngAddr g{e.k0, e.k1, grow, col, e.tick, e.stream & 0xFFu};
        return nb_draw_slow(x, c.y, u, e.pshift, e.mu_inv_max, g);
    }
    const float p = om * fast_rcp(fmaf(c.z, om, 1.0f)) - e.pshift;
    const float q = 1.0f - p, qr1 = q * c.w;
    float P = fast_ex2(c.y * fast_lg2(p));
    float k = 0.0f, rk = 1.0f;
    while (u >= P && P > PMIN_F) {
        u -= P;
        k += 1.0f;
        const float ratio = fmaf(qr1, rk, q);
        rk = fast_rcp(k + 1.0f);
        P *= ratio;
    }
    return k;
}

__device__ __forceinline__ void epi_bern4(const EpiParams& e, int64_t row, int64_t col0, const float (&acc)[4], int nvalid, const float (&bias)[4]) {
    const uint4 w = philox_group_rk(e, (uint32_t)row, (uint32_t)(col0 >> 2));
    const float u[4] = {u23(w.x), u23(w.y), u23(w.z), u23(w.w)};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (j < nvalid) ? bern_draw_fast(e.factor * fmaf(e.wscale, acc[j], bias[j]), u[j]) : 0.0f;
    const int64_t i = row * e.out.ld + col0;
    if (nvalid == 4) {
        *reinterpret_cast<uint2*>(e.out.p0 + i) = make_uint2(pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]));
    } else {
        for (int j = 0; j < nvalid; ++j) e.out.p0[i + j] = plane_encode(o[j]);
    }
}

}  // namespace scdbm
