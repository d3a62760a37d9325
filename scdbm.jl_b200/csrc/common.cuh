// Shared definitions for libscdbm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <string>

namespace scdbm {

// ---- device state matrix: [rows x cols], row = chain/cell, unit-contiguous.
// Stored as one or two FP16 "planes" whose SUM is the value, so that every state is directly a tcgen05
// kind::f16 operand without conversion (A and B operands must share one 16-bit format; FP16 holds
// integers to 2048 exactly, which keeps the second count plane empty for counts < 2048).
//   BINARY : p0 in {0,1}                                   (exact)
//   COUNT  : p0 = v mod 2048, p1 = v - p0 (multiple of 2048) -- exact to 65535, p1 != 0 only for counts >= 2048
//   REAL   : p0 = fp16(v),   p1 = fp16(v - p0)             (22 significant bits in fp16's normal range)
//            plus an fp32 master copy `f` (what downloads, mean-field deltas and the
//            CUDA-core engine read; the planes are the tensor-core operands)
using plane_t = __half;   // one 16-bit word of a state / weight plane: IEEE fp16

enum MatKind : int32_t { MAT_BINARY = 0, MAT_COUNT = 1, MAT_REAL = 2 };

struct MatView {
    plane_t* p0;
    plane_t* p1;     // nullptr for BINARY
    float* f;              // fp32 master copy, REAL only (else nullptr)
    uint32_t* hiflag;      // COUNT only, per 128-row tile: bit 0 = the p1 (hi) plane may hold a non-zero entry (consumers read it),
                           // bit 1 = it may hold stale non-zeros (the NB sampling kernel zero-fills it before patching); 0 = all zero
    int64_t rows, cols, ld; // ld (elements) is a multiple of 64
    int32_t kind;
};

// ---- programmatic dependent launch.  The latency-bound training loops (BASELINE configs C1 / C2) are chains of ~10 small
// dependent kernels per minibatch; a plain stream launch leaves the GPU idle for the launch latency between each pair.
// Kernels launched through launch_pdl may become resident while their predecessor still runs: they set up (barriers,
// TMEM), call pdl_wait() -- which returns once the predecessor grid has completed and its writes are visible -- and only then
// touch global memory.  pdl_trigger() lets the successor start its own set-up; it is issued only after this CTA holds every
// resource it needs (TMEM), so a successor can never starve a predecessor.  Both are no-ops in a plainly launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool& pdl_enabled() {
    static bool on = [] { const char* v = getenv("SCDBM_PDL"); return !(v && v[0] == '0'); }();
    return on;
}
// launch `kernel` with programmatic stream serialization; the kernel MUST call pdl_wait() before its first global access
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr float COUNT_BASE = 2048.0f;   // COUNT planes: p0 = v mod 2048, p1 = v - p0

__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// epilogue modes of the fused sweep GEMM
enum EpiMode : int32_t {
    EPI_INPUT = 0,      // x                         (hiddeninput!/visibleinput!)
    EPI_SIGM = 1,       // sigm(x)                   (hiddenpotential!, Bernoulli visiblepotential!, mean-field)
    EPI_SIGM_BERN = 2,  // u < sigm(x)               (samplehidden!, sigm_bernoulli!)
    EPI_NBMEAN = 3,     // r e^x/(1-e^x) + 1e-6      (NB visiblepotential!)
    EPI_NB_DRAW = 4,    // NB(r, r/(mu+r)) draw      (NB samplevisiblepotential!)
    EPI_AIS_RATIO = 5,  // row-sum of log((1+e^{t1 I+a})/(1+e^{t2 I+a}))
    EPI_RAW_F32 = 6     // acc -> fp32 matrix (gradient GEMM)
};

struct EpiParams {
    int32_t mode;
    const float* bias0;     // per-column bias (nullable)
    const float* bias1;     // second per-column bias (DBM intermediate layers; nullable)
    const float* addend;    // optional fp32 [rows x ld_add] added to the pre-activation
    int64_t ld_add;         //   (mean-field: the hoisted x*W1)
    int32_t addend_splits;  // > 1: the addend is the sum of this many slabs, addend_stride elements apart (split-K partial sums)
    int32_t addend_scaled;  // 1: the addend stands for the contraction itself (split-K): wscale multiplies it, as it would the accumulator
    int64_t addend_stride;
    float factor;           // multiplies (wscale*acc + biases): up/down factor
    float wscale;           // multiplies acc only: AIS temperature (weights-only annealing)
    const float* acc_scale; // tensor-core engine: device scalar 1/s undoing the power-of-two scale s of the fp16 weight planes
    const float* r;         // NB inverse dispersion per column
    const float4* nbc;      // NB per-gene constants {a, r, 1e-6/r, r-1} (fast sampling epilogue)
    const uint32_t* nbT;    // NB per-gene threshold: Philox word < T  <=>  u < lower bound of P(v = 0 | h) over all h (two-pass epilogue)
    float pshift;           // 1e-6 for the gamma-Poisson form of the reference, else 0
    float mu_inv_max;       // inversion / gamma-Poisson switch
    uint32_t k0, k1;        // Philox key (seed)
    uint32_t rk[20];        // Philox round keys {k0 + i W0, k1 + i W1}: constant-bank operands of the sampling epilogues
    uint32_t tick, stream;  // Philox counter words 2,3
    uint32_t row_offset;    // global index of row 0 (chain sharding)
    const float* uniforms;  // optional explicit uniforms [rows x ldu], row-major fp32
    int64_t ldu;
    MatView out;            // output state (not used by EPI_AIS_RATIO)
    MatView prev;           // mean-field: previous value of the same layer (p0 == nullptr: off)
    unsigned int* delta_bits; // mean-field: max |prev-new| as float bits (atomicMax)
    const unsigned int* skip; // mean-field: device flag "the fixed point has converged" -- a set flag turns the launch into a no-op
    float t1, t2;           // AIS temperatures
    double* logw;           // AIS: per-row accumulator
    float* out_f32;         // EPI_RAW_F32 destination [M x ldo]
    int64_t ldo;
};

// One GEMM operand, element (mn, k) at base[mn * s_mn + k * s_k]; value is
// p0 + p1 (bf16 planes) or f (fp32).  `scale` multiplies the A operand of a
// source (gradient: +1/n_pos, -1/n_neg).
struct Operand {
    const plane_t* p0;
    const plane_t* p1;
    const float* f;
    int64_t s_mn, s_k;
    int32_t kind;          // MatKind of the planes
};
struct GemmSource {
    Operand a, b;
    int64_t K;
    float scale;
};
struct GemmArgs {
    GemmSource src[2];
    int32_t nsrc;
    int64_t M, N;
    EpiParams epi;
};

}  // namespace scdbm
