// Elementwise draws: Bernoulli and Negative-Binomial (fp32 mirror of
// oracle/sampling.py; reference: src/gibbssampling.jl:40-45, 796-820, 852-857).
#pragma once
#include "philox.cuh"

namespace scdbm {

constexpr float KMAX_F = 65535.0f;
constexpr float PMIN_F = 1.4901161193847656e-08f;  // 2^-26
constexpr int MAX_ATTEMPTS = 64;

__device__ __forceinline__ float sigmf(float x) { return 1.0f / (1.0f + expf(-x)); }

// src/gibbssampling.jl:1018 : r e^x/(1-e^x) + 1e-6, with 1-e^x = -expm1(x)
__device__ __forceinline__ float nb_mean(float x, float r) {
    // numerator e^x directly (expm1(x)+1 cancels for very negative x); denominator via expm1 only
    // where 1-e^x itself would cancel (x -> 0-)
    const float ex = expf(x);
    const float den = (x > -0.6931472f) ? -expm1f(x) : 1.0f - ex;
    return r * ex / den + 1e-6f;
}

// chop-down inversion of a pmf with P_{k+1} = P_k * ratio(k)
template <class Ratio>
__device__ __forceinline__ float invert_pmf(float u, float P, Ratio ratio) {
    float k = 0.0f;
    while (u >= P && P > PMIN_F && k < KMAX_F) {
        u -= P;
        P *= ratio(k);
        k += 1.0f;
    }
    return k;
}

__device__ __forceinline__ float nb_inversion(float mu, float r, float u, float pshift) {
    const float p = r / (mu + r) - pshift;
    const float q = 1.0f - p;
    const float P0 = expf(r * logf(p));
    return invert_pmf(u, P0, [=](float k) { return q * (k + r) / (k + 1.0f); });
}

struct RngAddr {
    uint32_t k0, k1, row, col, tick, layer;
};

__device__ __noinline__ float gamma_mt(float a, const RngAddr& g) {
    const bool small = a < 1.0f;
    const float d = (small ? a + 1.0f : a) - (1.0f / 3.0f);
    const float c = 1.0f / sqrtf(9.0f * d);
    float boost = 1.0f;
    for (int attempt = 0; attempt < MAX_ATTEMPTS; ++attempt) {
        const uint4 w = philox_element(g.k0, g.k1, g.row, g.col, g.tick, stream_id(g.layer, P_GAMMA, attempt));
        if (attempt == 0 && small) boost = powf(u23(w.w), 1.0f / a);
        const float z = sqrtf(-2.0f * logf(u23(w.x))) * cospif(2.0f * u23(w.y));
        const float t = 1.0f + c * z;
        const float v = t * t * t;
        if (v > 0.0f && logf(u23(w.z)) < 0.5f * z * z + d - d * v + d * logf(v)) return d * v * boost;
    }
    return a;
}

__device__ __noinline__ float poisson_draw(float lam, const RngAddr& g) {
    if (lam < 10.0f) {
        const uint4 w = philox_element(g.k0, g.k1, g.row, g.col, g.tick, stream_id(g.layer, P_POIS, 0));
        return invert_pmf(u23(w.x), expf(-lam), [=](float k) { return lam / (k + 1.0f); });
    }
    const float slam = sqrtf(lam), loglam = logf(lam);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
    const float vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (int attempt = 0; attempt < MAX_ATTEMPTS; ++attempt) {
        const uint4 w = philox_element(g.k0, g.k1, g.row, g.col, g.tick, stream_id(g.layer, P_POIS, attempt));
        const float U = u23(w.x) - 0.5f, V = u23(w.y);
        const float us = 0.5f - fabsf(U);
        const float k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return fminf(k, KMAX_F);
        if (k < 0.0f || (us < 0.013f && V > us)) continue;
        if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -lam + k * loglam - lgammaf(k + 1.0f))
            return fminf(k, KMAX_F);
    }
    return floorf(lam);
}

// NB(r, p = r/(mu+r) - pshift) draw.  u: uniform for the inversion branch.
__device__ __forceinline__ float nb_draw(float mu, float r, float u, float pshift, float mu_inv_max, const RngAddr& g) {
    if (mu <= mu_inv_max) return nb_inversion(mu, r, u, pshift);
    const float p = r / (mu + r) - pshift;
    const float lam = gamma_mt(r, g) * ((1.0f - p) / p);
    return fminf(poisson_draw(lam, g), KMAX_F);
}

// ---- fast-math variants used by the tensor-core engine's sampling epilogues (Philox-fed draws).
// Same algorithm and the same uniforms as above; MUFU ex2/lg2/rcp approximations instead of the
// IEEE library calls.  A draw can differ from the accurate path only where u sits within ~1e-6
// (relative) of a CDF step.
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// rare path of the fast epilogue: large means (gamma-Poisson) or x -> 0- (1 - e^x cancels); accurate math
__device__ __noinline__ float nb_draw_slow(float x, float r, float u, float pshift, float mu_inv_max, RngAddr g) {
    return nb_draw(nb_mean(x, r), r, u, pshift, mu_inv_max, g);
}

// u < sigm(x)  <=>  u (1 + e^-x) < 1
__device__ __forceinline__ float bern_draw_fast(float x, float u) {
    return (u * (1.0f + fast_ex2(-x * 1.4426950408889634f)) < 1.0f) ? 1.0f : 0.0f;
}

}  // namespace scdbm
