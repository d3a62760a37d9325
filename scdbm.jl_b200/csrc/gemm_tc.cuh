// tcgen05 / TMA / TMEM sweep GEMM (sm_100a) -- interface.  (stub: replaced below)
#pragma once
#include "epilogue.cuh"
namespace scdbm {
struct TcContext { void* encode = nullptr; };
struct TcOperandPlanes { MatView a; const __nv_bfloat16 *b0, *b1; int64_t ldb, N, K; };
struct TcGradArgs { MatView v, h, vm, hm; float pos_scale, neg_scale; float* out; int64_t V, H; };
inline void tc_context_init(TcContext&) {}
inline bool tc_supported(const GemmArgs&, int) { return false; }
inline bool tc_grad_supported(const TcGradArgs&) { return false; }
inline cudaError_t launch_gemm_tc(TcContext&, const TcOperandPlanes*, int, const GemmArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t launch_grad_tc(TcContext&, const TcGradArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
}
