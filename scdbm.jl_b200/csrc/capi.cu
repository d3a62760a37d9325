// libscdbm_b200: C ABI + host orchestration (handles, chain drivers, mean-field,
// CD/PCD epoch, DBM PCD step, AIS).  Kernels live in the included headers.
#include "../../include/scdbm_b200.h"
#include "elementwise.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "sparse.cuh"
#include "comm.cuh"

#include <algorithm>
#include <functional>
#include <atomic>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

using namespace scdbm;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
struct Err : std::runtime_error {
    int code;
    Err(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            throw Err(e__ == cudaErrorMemoryAllocation ? SCDBM_ENOMEM : SCDBM_ECUDA,                     \
                      std::string(#call) + ": " + cudaGetErrorString(e__));                             \
    } while (0)
#define REQUIRE(cond, msg)                          \
    do {                                            \
        if (!(cond)) throw Err(SCDBM_EINVAL, msg);  \
    } while (0)
#define REQUIRE_OR_RETURN(cond)                     \
    do {                                            \
        if (!(cond)) {                              \
            g_err = "NULL argument";                \
            return SCDBM_EINVAL;                    \
        }                                           \
    } while (0)
#define API_BEGIN try {
#define API_END                          \
    return SCDBM_OK;                     \
    }                                    \
    catch (const Err& e) {               \
        g_err = e.what();                \
        return e.code;                   \
    }                                    \
    catch (const std::exception& e) {    \
        g_err = e.what();                \
        return SCDBM_EINVAL;             \
    }

// ------------------------------------------------------------------ handles
struct scdbm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t seed = 0;
    float mu_inv_max = 8.0f;
    float nb_pshift = 0.0f;
    int engine = SCDBM_ENGINE_AUTO;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned int* d_delta = nullptr;
    unsigned int *h_flag = nullptr, *d_flag = nullptr;   // zero-copy status word (mapped pinned host memory): no copy-engine read-back
    int64_t launches = 0;
    int64_t tc_launches = 0;
    int num_sms = 148;
    bool use_cg2 = true;    // two-SM (cta_group::2) sweep kernel for the L2-operand-bound Bernoulli / deterministic sweeps
    TcContext tc;
    float* scratch = nullptr;   // reduction workspace (column-sum partials)
    int64_t scratch_cap = 0;
    // asynchronous downloads: dense uint16 snapshot in HBM, drained to the host by the copy engine on its own stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_conv = nullptr, ev_copy = nullptr;
    uint16_t* stage16 = nullptr;
    size_t stage16_cap = 0;
    bool copy_pending = false;
    // mean-field convergence state {delta bits, converged, iterations, last delta bits} on the device, mirrored into mapped
    // pinned host memory by k_mf_check so that the host reads it without a copy
    unsigned int* d_mf = nullptr;
    unsigned int *h_mf = nullptr, *d_hmf = nullptr;   // mapped mirror of the state: two 64-bit words, see k_mf_check
    unsigned int mf_gen = 0;                          // generation tag of the current mean-field call
    // data-parallel communicator (scdbm_comm_init): NCCL over NVLink, collectives ordered on the context's streams
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t comm_stream = nullptr;     // gradient allreduce overlapping the remaining gradient GEMMs
    cudaEvent_t ev_grad = nullptr, ev_comm = nullptr;
    double* d_red = nullptr;                // small reduction buffer (scdbm_comm_allreduce_f64)
    int64_t live = 0;      // child handles still alive
    bool closed = false;   // scdbm_ctx_destroy was called; freed when `live` drops to 0
};
static void ctx_free(scdbm_ctx* c) {
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        cudaEventDestroy(c->ev_conv);
        cudaEventDestroy(c->ev_copy);
    }
    if (c->comm) nccl_api().CommDestroy(c->comm);
    if (c->comm_stream) { cudaStreamDestroy(c->comm_stream); cudaEventDestroy(c->ev_grad); cudaEventDestroy(c->ev_comm); }
    cudaFree(c->d_red);
    cudaFree(c->d_mf);
    if (c->h_mf) cudaFreeHost(c->h_mf);
    cudaFree(c->stage16);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaFree(c->d_delta);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    cudaFree(c->scratch);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}
static void ctx_retain(scdbm_ctx* c) { c->live++; }
static void ctx_release(scdbm_ctx* c) {
    if (--c->live == 0 && c->closed) ctx_free(c);
}

struct Layer {
    int V = 0, H = 0, kind = SCDBM_VIS_BERNOULLI;
    float *W = nullptr, *a = nullptr, *b = nullptr, *r = nullptr;
    plane_t *w0 = nullptr, *w1 = nullptr, *t0 = nullptr, *t1 = nullptr;  // W[V x ldh], Wt[H x ldv] hi/lo
    float4* nbc = nullptr;  // NB layer: per-gene {a, r, 1e-6/r, r-1}
    uint32_t* nbT = nullptr;   // NB layer: per-gene Philox-word threshold of the zero shortcut, padded with 0xFFFFFFFF
    float nbL_mu = -1.0f;   // mu_inv_max the bound was computed for
    const float* scale = nullptr;   // model's device scalars {s, 1/s}: power-of-two scale of the fp16 weight planes
    int64_t ldh = 0, ldv = 0;
};

struct scdbm_model {
    scdbm_ctx* ctx;
    std::vector<Layer> L;
    int64_t live = 0;
    bool closed = false;
    float* d_scale = nullptr;          // {s, 1/s}
    unsigned int* d_wmax = nullptr;    // 3 slots of max |W| over all layers (float bits); slot wmax_round holds the current maximum,
    int wmax_round = 0;                //   the fused update kernel (k_update_layer) rotates them
    int units(int layer) const { return layer == 0 ? L[0].V : L[layer - 1].H; }
    int nlayers() const { return (int)L.size() + 1; }
};

static void model_free(scdbm_model* m) {
    // stream-ordered frees: destroying a model never waits for the copy stream (an asynchronous download in flight)
    cudaStream_t st = m->ctx->stream;
    for (auto& L : m->L) {
        float* f[] = {L.W, L.a, L.b, L.r};
        for (float* p : f) if (p) cudaFreeAsync(p, st);
        plane_t* h[] = {L.w0, L.w1, L.t0, L.t1};
        for (plane_t* p : h) if (p) cudaFreeAsync(p, st);
        if (L.nbc) cudaFreeAsync(L.nbc, st);
        if (L.nbT) cudaFreeAsync(L.nbT, st);
    }
    if (m->d_scale) cudaFreeAsync(m->d_scale, st);
    if (m->d_wmax) cudaFreeAsync(m->d_wmax, st);
    scdbm_ctx* c = m->ctx;
    delete m;
    ctx_release(c);
}
static void model_retain(scdbm_model* m) { m->live++; }
static void model_release(scdbm_model* m) {
    if (--m->live == 0 && m->closed) model_free(m);
}

// Reference-counted: the creator holds one reference; every handle handed out by scdbm_particles_layer /
// scdbm_trainer_chainstate / scdbm_trainer_estimated_mean adds one, so a handle can never dangle when its
// particles / trainer object is destroyed first.  scdbm_mat_destroy drops one reference.
struct scdbm_mat {
    scdbm_ctx* ctx;
    MatView v;
    bool owns;
    int64_t* csr_indptr = nullptr;   // device row pointers of the pending CSR download (scdbm_mat_csr_prepare)
    int64_t csr_nnz = -1;
    int64_t refs = 1;
    scdbm_mat* parent = nullptr;     // view (owns == false): the matrix whose planes `v` aliases, retained
};

struct scdbm_particles {
    scdbm_model* model;
    int64_t n;
    uint64_t row_offset;
    int64_t clock;
    bool real_valued;
    std::vector<scdbm_mat*> cur, nxt;
    scdbm_mat* alias0 = nullptr;  // mean-field: layer 0 aliases the data matrix
    float* hoist = nullptr;       // mean-field: x * W1 (fp32 [n x H1])
};

struct scdbm_opt {
    scdbm_model* model;
    float* buf = nullptr;
    int64_t n = 0;
    std::vector<int64_t> offW, offA, offB;
};

struct scdbm_trainer {
    scdbm_model* model;
    int l, B, pcd;
    int64_t clock;
    scdbm_mat *chain = nullptr, *v = nullptr, *h = nullptr, *hm = nullptr, *hs = nullptr, *vm = nullptr, *vs = nullptr;
    scdbm_opt* opt = nullptr;
    int32_t* d_perm = nullptr;
    int64_t perm_cap = 0;
    int64_t v_groups = 0;          // minibatches gathered per launch into `v` (rows: v_groups x round_up(B, 128))
    uint32_t* d_anyhi = nullptr;   // does the data matrix of this epoch hold counts >= 2048 anywhere?
    bool track_mean = false;      // record `estimatedmean` (model means in minibatch order) each epoch
    scdbm_mat* est = nullptr;
};

// ------------------------------------------------------------------ helpers
static inline int grid1d(int64_t n, int block = 256) { return (int)std::min<int64_t>((n + block - 1) / block, 148 * 16); }

// out[c] (+)= scale * sum_r m(r, c), deterministic
static void colsum(scdbm_ctx* c, const MatView& m, float scale, float* out, bool accumulate) {
    if (m.cols == 0) return;
    const int nslabs = (int)std::max<int64_t>(1, std::min<int64_t>((m.rows + 255) / 256, 256));
    const int64_t need = (int64_t)nslabs * m.cols;
    if (c->scratch_cap < need) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->scratch);
        c->scratch = nullptr;
        if (cudaMalloc(&c->scratch, need * sizeof(float)) != cudaSuccess) throw Err(SCDBM_ENOMEM, "column-sum workspace");
        c->scratch_cap = need;
    }
    dim3 grid((unsigned)((m.cols + 255) / 256), (unsigned)nslabs), block(32, 8);
    CK(launch_pdl(k_colsum_partial, dim3(grid), dim3(block), 0, c->stream, m, c->scratch));
    CK(launch_pdl(k_colsum_final, dim3((unsigned)((m.cols + 127) / 128)), dim3(128), 0, c->stream, c->scratch, nslabs, m.cols, scale, out, accumulate ? 1 : 0));
    c->launches += 2;
}

static scdbm_mat* mat_new(scdbm_ctx* ctx, int64_t rows, int64_t cols, int kind) {
    REQUIRE(rows >= 0 && cols >= 0, "matrix dimensions must be non-negative");
    REQUIRE(kind >= 0 && kind <= 2, "unknown matrix kind");
    auto* m = new scdbm_mat{ctx, {}, true};
    ctx_retain(ctx);
    m->v.rows = rows;
    m->v.cols = cols;
    m->v.ld = round_up(std::max<int64_t>(cols, 1), 64);
    m->v.kind = kind;
    const size_t bytes = (size_t)std::max<int64_t>(rows, 1) * m->v.ld * sizeof(plane_t);
    // stream-ordered pool allocations: repeated samples()/particles calls re-use the same memory
    CK(cudaMallocAsync(&m->v.p0, bytes, ctx->stream));
    CK(cudaMemsetAsync(m->v.p0, 0, bytes, ctx->stream));
    if (kind != MAT_BINARY) {
        CK(cudaMallocAsync(&m->v.p1, bytes, ctx->stream));
        CK(cudaMemsetAsync(m->v.p1, 0, bytes, ctx->stream));
    }
    if (kind == MAT_COUNT) {
        const size_t fb = (size_t)((std::max<int64_t>(rows, 1) + 127) / 128) * sizeof(uint32_t);
        CK(cudaMallocAsync(&m->v.hiflag, fb, ctx->stream));
        CK(cudaMemsetAsync(m->v.hiflag, 0, fb, ctx->stream));   // planes start zeroed
    }
    if (kind == MAT_REAL) {
        CK(cudaMallocAsync(&m->v.f, bytes * 2, ctx->stream));
        CK(cudaMemsetAsync(m->v.f, 0, bytes * 2, ctx->stream));
    }
    return m;
}
static scdbm_mat* mat_retain(scdbm_mat* m) {
    if (m) m->refs++;
    return m;
}
// drops one reference; the last one frees (stream-ordered for the device memory)
static void mat_free(scdbm_mat* m) {
    if (!m || --m->refs > 0) return;
    if (m->csr_indptr) cudaFreeAsync(m->csr_indptr, m->ctx->stream);
    if (m->owns) {
        cudaStream_t st = m->ctx->stream;
        if (m->v.p0) cudaFreeAsync(m->v.p0, st);
        if (m->v.p1) cudaFreeAsync(m->v.p1, st);
        if (m->v.f) cudaFreeAsync(m->v.f, st);
        if (m->v.hiflag) cudaFreeAsync(m->v.hiflag, st);
        scdbm_ctx* c = m->ctx;
        delete m;
        ctx_release(c);
        return;
    }
    scdbm_mat* parent = m->parent;
    delete m;
    mat_free(parent);
}
// non-owning view of another matrix' planes (mean-field: layer 0 of mu is the data matrix); keeps `of` alive
static scdbm_mat* mat_view_of(scdbm_mat* of) {
    auto* m = new scdbm_mat{of->ctx, of->v, false};
    m->parent = mat_retain(of);
    return m;
}
static MatView rows_view(const MatView& v, int64_t rows) {
    MatView r = v;
    r.rows = std::min(rows, v.rows);
    return r;
}

// rows [r0, r0 + rows) of a matrix; r0 is a multiple of 128 (the hi-plane flags are per 128-row tile)
static MatView rows_at(const MatView& v, int64_t r0, int64_t rows) {
    MatView r = v;
    r.p0 += r0 * v.ld;
    if (r.p1) r.p1 += r0 * v.ld;
    if (r.f) r.f += r0 * v.ld;
    if (r.hiflag) r.hiflag += r0 / 128;
    r.rows = std::min(rows, v.rows - r0);
    return r;
}

// hi-plane flags of a COUNT matrix (see MatView): `all` = true after any writer that does not maintain them
// (gather, copies, generic epilogues); false before a writer that rewrites both planes and raises the flags itself
static void set_hiflags(scdbm_ctx* ctx, const MatView& v, bool all) {
    if (!v.hiflag) return;
    const size_t fb = (size_t)((std::max<int64_t>(v.rows, 1) + 127) / 128) * sizeof(uint32_t);
    cudaMemsetAsync(v.hiflag, all ? 0x03 : 0x00, fb, ctx->stream);   // 3 = in use + may hold stale entries; 0 = all zero
}

static EpiParams epi_base(scdbm_ctx* ctx) {
    EpiParams e;
    memset(&e, 0, sizeof(e));
    e.factor = 1.0f;
    e.wscale = 1.0f;
    e.pshift = ctx->nb_pshift;
    e.mu_inv_max = ctx->mu_inv_max;
    e.k0 = (uint32_t)(ctx->seed & 0xFFFFFFFFu);
    e.k1 = (uint32_t)(ctx->seed >> 32);
    philox_round_keys(e.k0, e.k1, e.rk);
    return e;
}

// One contraction source of a sweep: state matrix x one RBM's weights.
struct SweepSrc {
    const MatView* a;
    const Layer* layer;
    bool up;  // up: out units = hidden of `layer` (K = V); down: out units = visible (K = H)
};

static void run_gemm(scdbm_ctx* ctx, GemmArgs& g, const SweepSrc* srcs, int nsrc) {
    bool use_tc = false;
    if (ctx->engine != SCDBM_ENGINE_SIMT && srcs) {
        use_tc = tc_supported(g, nsrc);
        // AUTO: only toy shapes go to the FFMA kernel (no descriptor set-up, fp32 operands).  A one-tile tcgen05 launch
        // costs ~11 us; the FFMA kernel needs 75 us for 100 x 256 -> 64 (C2's top layer, measured)
        double work = 0.0;
        for (int s = 0; s < nsrc; ++s) work += (double)g.M * (double)g.N * (double)g.src[s].K;
        if (ctx->engine == SCDBM_ENGINE_AUTO && work < (double)(1 << 17)) use_tc = false;
    }
    if (ctx->engine == SCDBM_ENGINE_TC && !use_tc && srcs)
        throw Err(SCDBM_EUNSUP, "engine=tc requested but the shape is not tiled by the tcgen05 kernel");
    if (use_tc) {
        TcOperandPlanes ops[2];
        for (int s = 0; s < nsrc; ++s) {
            const Layer& L = *srcs[s].layer;
            ops[s].a = *srcs[s].a;
            ops[s].b0 = srcs[s].up ? L.t0 : L.w0;
            ops[s].b1 = srcs[s].up ? L.t1 : L.w1;
            ops[s].ldb = srcs[s].up ? L.ldv : L.ldh;
            ops[s].N = srcs[s].up ? L.H : L.V;
            ops[s].K = srcs[s].up ? L.V : L.H;
        }
        // the Philox-fed tensor-core NB epilogue maintains the hi-plane flags itself: row tiles whose hi plane was in
        // use are marked for zero-filling, all others are known to be all zero and are not written at all.  Every
        // other writer of a COUNT matrix marks all row tiles as "hi plane possibly in use"
        if (tc_epilogue_class(g, ops, nsrc) == 1 && g.epi.out.hiflag) {
            const int64_t nt = (std::max<int64_t>(g.epi.out.rows, 1) + 127) / 128;
            CK(launch_pdl(k_flags_advance, dim3((unsigned)((nt + 255) / 256)), dim3(256), 0, ctx->stream, g.epi.out.hiflag, nt));
            ctx->launches++;
        } else {
            set_hiflags(ctx, g.epi.out, true);
        }
        const int S = tc_plan_split(g, ops, nsrc, ctx->num_sms);
        if (S > 1) {
            // split-K: S partial sums of the contraction as raw fp32 slabs [S][M][ldw], then one finishing launch
            const int64_t ldw = round_up(g.N, 4), stride = g.M * ldw, need = (int64_t)S * stride;
            if (ctx->scratch_cap < need) {
                CK(cudaStreamSynchronize(ctx->stream));
                cudaFree(ctx->scratch);
                ctx->scratch = nullptr;
                ctx->scratch_cap = 0;
                CK(cudaMalloc(&ctx->scratch, need * sizeof(float)));
                ctx->scratch_cap = need;
            }
            GemmArgs gp = g;
            gp.epi.mode = EPI_RAW_F32;
            gp.epi.out_f32 = ctx->scratch;
            gp.epi.ldo = ldw;
            gp.epi.bias0 = gp.epi.bias1 = nullptr;
            CK(launch_gemm_tc(ctx->tc, ops, nsrc, gp, ctx->num_sms, ctx->stream, S));
            const int64_t nthreads = g.M * ((g.N + 3) / 4);
            CK(launch_pdl(k_det_finish, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, ctx->stream, ctx->scratch, S, stride, ldw, g.epi, g.M, g.N));
            CK(cudaGetLastError());
            ctx->launches++;
        } else if (ctx->use_cg2 && tc_cg2_eligible(g, ops, nsrc, ctx->num_sms)) {
            CK(launch_gemm_cg2(ctx->tc, ops, nsrc, g, ctx->num_sms, ctx->stream));
        } else {
            CK(launch_gemm_tc(ctx->tc, ops, nsrc, g, ctx->num_sms, ctx->stream));
        }
        ctx->tc_launches++;
    } else if (!srcs && nsrc == 0 && g.epi.addend && (g.epi.ld_add & 3) == 0 && (reinterpret_cast<uintptr_t>(g.epi.addend) & 15u) == 0u &&
               !g.epi.prev.p0 && !g.epi.uniforms && !g.epi.skip &&
               (g.epi.mode == EPI_INPUT || g.epi.mode == EPI_SIGM || g.epi.mode == EPI_NBMEAN)) {
        // no contraction at all (mean-field initialisation from the hoisted x W1): an elementwise pass, coalesced
        set_hiflags(ctx, g.epi.out, true);
        EpiParams e = g.epi;
        e.wscale = e.addend_scaled ? e.wscale : 1.0f;
        e.addend = nullptr;
        const int64_t nthreads = g.M * ((g.N + 3) / 4);
        CK(launch_pdl(k_det_finish, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), 0, ctx->stream, g.epi.addend, std::max(g.epi.addend_splits, 1),
                      g.epi.addend_stride, g.epi.ld_add, e, g.M, g.N));
    } else {
        set_hiflags(ctx, g.epi.out, true);
        CK(launch_gemm_simt(g, ctx->stream));
    }
    ctx->launches++;
}

static void run_sweep(scdbm_ctx* ctx, const SweepSrc* srcs, int nsrc, int64_t M, int64_t N, EpiParams& epi) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.nsrc = nsrc;
    g.M = M;
    g.N = N;
    for (int s = 0; s < nsrc; ++s) {
        const MatView& a = *srcs[s].a;
        const Layer& L = *srcs[s].layer;
        const int64_t K = srcs[s].up ? L.V : L.H;
        REQUIRE(a.cols == K, "state width does not match the weight matrix");
        REQUIRE(a.rows >= M, "state has fewer rows than requested");
        REQUIRE((srcs[s].up ? L.H : L.V) == N, "output width does not match the weight matrix");
        g.src[s].a = a.f ? Operand{nullptr, nullptr, a.f, a.ld, 1, a.kind} : Operand{a.p0, a.p1, nullptr, a.ld, 1, a.kind};
        g.src[s].b = srcs[s].up ? Operand{nullptr, nullptr, L.W, 1, (int64_t)L.H, MAT_REAL} : Operand{nullptr, nullptr, L.W, (int64_t)L.H, 1, MAT_REAL};
        g.src[s].K = K;
        g.src[s].scale = 1.0f;
    }
    g.epi = epi;
    if (nsrc > 0) g.epi.acc_scale = srcs[0].layer->scale + 1;   // used by the tensor-core engine only
    if (M == 0 || N == 0) return;
    run_gemm(ctx, g, srcs, nsrc);
}

// uniforms override: host column-major fp64 -> device row-major fp32
struct DevUniforms {
    float* d = nullptr;
    int64_t ld = 0;
    cudaStream_t st = nullptr;
    ~DevUniforms() { if (d) cudaFreeAsync(d, st); }
};
static void upload_uniforms(scdbm_ctx* ctx, const double* u, int64_t ldu, int64_t rows, int64_t cols, DevUniforms& out) {
    if (!u) return;
    REQUIRE(ldu >= rows, "uniforms: leading dimension smaller than the row count");
    double* stage = nullptr;
    const size_t nb = (size_t)ldu * cols * sizeof(double);
    out.st = ctx->stream;
    CK(cudaMallocAsync(&stage, std::max<size_t>(nb, 8), ctx->stream));
    CK(cudaMemcpyAsync(stage, u, nb, cudaMemcpyHostToDevice, ctx->stream));
    out.ld = cols;
    CK(cudaMallocAsync(&out.d, std::max<size_t>((size_t)rows * cols * sizeof(float), 8), ctx->stream));
    k_upload_uniforms<<<grid1d(rows * cols), 256, 0, ctx->stream>>>(stage, ldu, out.d, rows, cols, out.ld);
    ctx->launches++;
    CK(cudaFreeAsync(stage, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // the caller's host buffer must not be retained
}

static void refresh_nbc(scdbm_ctx* ctx, Layer& L, unsigned int* invalid = nullptr) {
    if (!L.nbc) return;
    k_refresh_nbc<<<(L.V + 127) / 128, 128, 0, ctx->stream>>>(L.a, L.r, L.W, L.V, L.H, ctx->mu_inv_max, L.nbc, L.nbT, invalid);
    L.nbL_mu = ctx->mu_inv_max;
    ctx->launches++;
}

// All weight planes of a model share one power-of-two scale (dual-source sweeps accumulate two layers' products
// in one TMEM accumulator), kept on the device so that training never syncs for it: max |W| over all layers ->
// scale -> fp16 hi/lo planes of every layer.  Called whenever any fp32 master weight changed.
static void refresh_operands(scdbm_model* m) {
    scdbm_ctx* ctx = m->ctx;
    CK(cudaMemsetAsync(m->d_wmax, 0, 3 * sizeof(unsigned int), ctx->stream));
    m->wmax_round = 0;
    for (auto& L : m->L) {
        const int64_t n = (int64_t)L.V * L.H;
        k_absmax<<<grid1d(n), 256, 0, ctx->stream>>>(L.W, n, m->d_wmax);
    }
    k_weight_scale<<<1, 1, 0, ctx->stream>>>(m->d_wmax, m->d_scale);
    for (auto& L : m->L) {
        dim3 grid((L.V + 31) / 32, (L.H + 31) / 32), block(32, 8);
        k_refresh_operands<<<grid, block, 0, ctx->stream>>>(L.W, L.V, L.H, m->d_scale, L.w0, L.w1, L.ldh, L.t0, L.t1, L.ldv);
    }
    CK(cudaGetLastError());
    ctx->launches += 2 * (int64_t)m->L.size() + 1;
}

// ---- the three conditional operators, with explicit Philox stream/tick
static int epi_mode_for(int what, bool nb_visible) {
    if (what == SCDBM_OUT_INPUT) return EPI_INPUT;
    if (what == SCDBM_OUT_POTENTIAL) return nb_visible ? EPI_NBMEAN : EPI_SIGM;
    if (what == SCDBM_OUT_SAMPLE) return nb_visible ? EPI_NB_DRAW : EPI_SIGM_BERN;
    throw Err(SCDBM_EINVAL, "`what` must be SCDBM_OUT_INPUT/POTENTIAL/SAMPLE");
}

struct DrawCfg {
    uint32_t stream = 0, tick = 0, row_offset = 0;
    const float* uniforms = nullptr;
    int64_t ldu = 0;
    float wscale = 1.0f;
    float pshift = -1.0f;  // <0: context default
};

static void op_hidden(scdbm_model* m, int l, const MatView& v, const MatView& out, float factor, int what, const DrawCfg& d,
                      int64_t rows = -1) {
    Layer& L = m->L[l];
    EpiParams e = epi_base(m->ctx);
    e.mode = epi_mode_for(what, false);
    e.bias0 = L.b;
    e.factor = factor;
    e.wscale = d.wscale;
    e.stream = d.stream; e.tick = d.tick; e.row_offset = d.row_offset;
    e.uniforms = d.uniforms; e.ldu = d.ldu;
    e.out = out;
    SweepSrc s{&v, &L, true};
    const int64_t M = rows < 0 ? v.rows : rows;
    REQUIRE(out.rows >= M && out.cols == L.H, "hidden output has the wrong shape");
    run_sweep(m->ctx, &s, 1, M, L.H, e);
}

static void op_visible(scdbm_model* m, int l, const MatView& h, const MatView& out, float factor, int what, const DrawCfg& d,
                       int64_t rows = -1) {
    Layer& L = m->L[l];
    const bool nb = L.kind == SCDBM_VIS_NEGBIN;
    EpiParams e = epi_base(m->ctx);
    e.mode = epi_mode_for(what, nb);
    e.bias0 = L.a;
    e.factor = (nb && what != SCDBM_OUT_INPUT) ? 1.0f : factor;  // NB visiblepotential! ignores `factor` (quirk B1)
    if (what == SCDBM_OUT_INPUT) e.factor = 1.0f;
    e.wscale = d.wscale;
    e.r = L.r;
    if (nb && L.nbL_mu != m->ctx->mu_inv_max) refresh_nbc(m->ctx, L);   // option changed since the last refresh
    e.nbc = L.nbc;
    e.nbT = L.nbT;
    if (d.pshift >= 0.0f) e.pshift = d.pshift;
    e.stream = d.stream; e.tick = d.tick; e.row_offset = d.row_offset;
    e.uniforms = d.uniforms; e.ldu = d.ldu;
    e.out = out;
    SweepSrc s{&h, &L, false};
    const int64_t M = rows < 0 ? h.rows : rows;
    REQUIRE(out.rows >= M && out.cols == L.V, "visible output has the wrong shape");
    run_sweep(m->ctx, &s, 1, M, L.V, e);
}

// intermediate DBM layer `layer` (1 <= layer <= nrbms-1) from both neighbours
static void op_dbm_layer(scdbm_model* m, int layer, const MatView* below, const float* below_hoist, int64_t ld_hoist,
                         const MatView& above, const MatView& out, int what, const DrawCfg& d, const MatView* prev,
                         unsigned int* delta_bits) {
    Layer& Lb = m->L[layer - 1];
    Layer& La = m->L[layer];
    EpiParams e = epi_base(m->ctx);
    e.mode = epi_mode_for(what, false);
    e.bias0 = Lb.b;
    e.bias1 = La.a;
    e.wscale = d.wscale;
    e.stream = d.stream; e.tick = d.tick; e.row_offset = d.row_offset;
    e.uniforms = d.uniforms; e.ldu = d.ldu;
    e.out = out;
    if (prev) { e.prev = *prev; e.delta_bits = delta_bits; }
    SweepSrc s[2];
    int ns = 0;
    if (below_hoist) { e.addend = below_hoist; e.ld_add = ld_hoist; }
    else s[ns++] = SweepSrc{below, &Lb, true};
    s[ns++] = SweepSrc{&above, &La, false};
    REQUIRE(out.cols == Lb.H && La.V == Lb.H, "DBM layer widths are inconsistent");
    run_sweep(m->ctx, s, ns, above.rows, Lb.H, e);
}

// ================================================================== C ABI
extern "C" {

const char* scdbm_last_error(void) { return g_err.c_str(); }
int32_t scdbm_version(void) { return 100; }

int32_t scdbm_ctx_create(int32_t device, uint64_t seed, void* stream, scdbm_ctx** out) {
    API_BEGIN
    REQUIRE(out, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Err(SCDBM_ECUDA, std::string("no CUDA device available (libscdbm_b200 has no CPU fallback): ") +
                                   cudaGetErrorString(e));
    REQUIRE(device >= 0 && device < count, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        throw Err(SCDBM_EUNSUP, std::string("libscdbm_b200 is built for sm_100a only; device is ") + prop.name);
    auto* c = new scdbm_ctx();
    c->device = device;
    c->seed = seed;
    c->num_sms = prop.multiProcessorCount;
    if (const char* v = getenv("SCDBM_CG2")) c->use_cg2 = v[0] == '1';
    c->tc.nb_cg2 = c->use_cg2;
    if (stream) c->stream = (cudaStream_t)stream;
    else { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaMalloc(&c->d_delta, sizeof(unsigned int)));
    CK(cudaHostAlloc(&c->h_flag, sizeof(unsigned int), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&c->d_flag, c->h_flag, 0));
    *c->h_flag = 0u;
    CK(cudaMalloc(&c->d_mf, 8 * sizeof(unsigned int)));
    CK(cudaHostAlloc(&c->h_mf, 4 * sizeof(unsigned int), cudaHostAllocMapped));
    memset(c->h_mf, 0, 4 * sizeof(unsigned int));
    CK(cudaHostGetDevicePointer(&c->d_hmf, c->h_mf, 0));
    {   // keep freed state matrices in the device's default memory pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    tc_context_init(c->tc);
    *out = c;
    API_END
}

int32_t scdbm_ctx_destroy(scdbm_ctx* c) {
    API_BEGIN
    if (!c) return SCDBM_OK;
    c->closed = true;   // children (models, matrices, ...) may outlive the Python/Julia wrapper of the context
    if (c->live == 0) ctx_free(c);
    API_END
}

int32_t scdbm_sync(scdbm_ctx* c) {
    API_BEGIN
    REQUIRE(c, "ctx is NULL");
    CK(cudaStreamSynchronize(c->stream));
    API_END
}

int32_t scdbm_ctx_set_option(scdbm_ctx* c, const char* name, double value) {
    API_BEGIN
    REQUIRE(c && name, "NULL argument");
    const std::string n(name);
    if (n == "engine") { REQUIRE(value >= 0 && value <= 2, "engine must be 0, 1 or 2"); c->engine = (int)value; }
    else if (n == "mu_inv_max") c->mu_inv_max = (float)value;
    else if (n == "nb_pshift") c->nb_pshift = (float)value;
    else if (n == "seed") c->seed = (uint64_t)value;
    else if (n == "two_sm") { c->use_cg2 = value != 0.0; c->tc.nb_cg2 = c->use_cg2; }
    else if (n == "pdl") pdl_enabled() = value != 0.0;          // process-wide: programmatic dependent launch on / off
    else throw Err(SCDBM_EINVAL, "unknown option: " + n);
    API_END
}

int32_t scdbm_ctx_get_option(scdbm_ctx* c, const char* name, double* value) {
    API_BEGIN
    REQUIRE(c && name && value, "NULL argument");
    const std::string n(name);
    if (n == "engine") *value = c->engine;
    else if (n == "mu_inv_max") *value = c->mu_inv_max;
    else if (n == "nb_pshift") *value = c->nb_pshift;
    else if (n == "seed") *value = (double)c->seed;
    else if (n == "two_sm") *value = c->use_cg2 ? 1.0 : 0.0;
    else if (n == "pdl") *value = pdl_enabled() ? 1.0 : 0.0;
    else if (n == "num_sms") *value = c->num_sms;
    else if (n == "tc_launches") *value = (double)c->tc_launches;
    else if (n == "comm_rank") *value = c->rank;
    else if (n == "comm_size") *value = c->nranks;
    else throw Err(SCDBM_EINVAL, "unknown option: " + n);
    API_END
}

int32_t scdbm_timer_start(scdbm_ctx* c) {
    API_BEGIN
    REQUIRE(c, "ctx is NULL");
    CK(cudaEventRecord(c->ev0, c->stream));
    API_END
}
int32_t scdbm_timer_stop(scdbm_ctx* c, float* ms) {
    API_BEGIN
    REQUIRE(c && ms, "NULL argument");
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    API_END
}
int32_t scdbm_launch_count(scdbm_ctx* c, int64_t* count, int32_t reset) {
    API_BEGIN
    REQUIRE(c, "ctx is NULL");
    if (count) *count = c->launches;
    if (reset) { c->launches = 0; c->tc_launches = 0; }
    API_END
}

// ------------------------------------------------------------------- model
int32_t scdbm_model_create(scdbm_ctx* c, int32_t nrbms, const int32_t* units, int32_t visible_kind, scdbm_model** out) {
    API_BEGIN
    REQUIRE(c && units && out, "NULL argument");
    REQUIRE(nrbms >= 1 && nrbms <= 16, "nrbms must be in 1..16");
    REQUIRE(visible_kind == SCDBM_VIS_BERNOULLI || visible_kind == SCDBM_VIS_NEGBIN, "unknown visible kind");
    for (int i = 0; i <= nrbms; ++i) REQUIRE(units[i] > 0, "layer sizes must be positive");
    auto* m = new scdbm_model{c, {}};
    ctx_retain(c);
    CK(cudaMallocAsync(&m->d_scale, 2 * sizeof(float), c->stream));
    CK(cudaMallocAsync(&m->d_wmax, 3 * sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(m->d_wmax, 0, 3 * sizeof(unsigned int), c->stream));
    {
        const float one[2] = {1.0f, 1.0f};
        CK(cudaMemcpyAsync(m->d_scale, one, sizeof(one), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    m->L.resize(nrbms);
    for (int l = 0; l < nrbms; ++l) {
        Layer& L = m->L[l];
        L.V = units[l];
        L.H = units[l + 1];
        L.kind = (l == 0) ? visible_kind : SCDBM_VIS_BERNOULLI;
        L.ldh = round_up(L.H, 64);
        L.ldv = round_up(L.V, 64);
        L.scale = m->d_scale;
        CK(cudaMallocAsync(&L.W, (size_t)L.V * L.H * sizeof(float), c->stream));
        CK(cudaMallocAsync(&L.a, (size_t)L.V * sizeof(float), c->stream));
        CK(cudaMallocAsync(&L.b, (size_t)L.H * sizeof(float), c->stream));
        CK(cudaMallocAsync(&L.r, (size_t)L.V * sizeof(float), c->stream));
        if (L.kind == SCDBM_VIS_NEGBIN) {
            // both per-gene tables are padded past the last column tile of any tile width (the padding never queues a draw;
            // a 2^-32 event -- Philox word == 0xFFFFFFFF -- reads benign constants there)
            CK(cudaMallocAsync(&L.nbc, (size_t)(round_up(L.V, 256) + 256) * sizeof(float4), c->stream));
            k_fill_nbc_pad<<<grid1d(round_up(L.V, 256) + 256), 256, 0, c->stream>>>(L.nbc, round_up(L.V, 256) + 256);
            CK(cudaMallocAsync(&L.nbT, (size_t)(round_up(L.V, 256) + 256) * sizeof(uint32_t), c->stream));
            CK(cudaMemsetAsync(L.nbT, 0xFF, (size_t)(round_up(L.V, 256) + 256) * sizeof(uint32_t), c->stream));
        }
        CK(cudaMemsetAsync(L.W, 0, (size_t)L.V * L.H * sizeof(float), c->stream));
        CK(cudaMemsetAsync(L.a, 0, (size_t)L.V * sizeof(float), c->stream));
        CK(cudaMemsetAsync(L.b, 0, (size_t)L.H * sizeof(float), c->stream));
        k_fill_f32<<<grid1d(L.V), 256, 0, c->stream>>>(L.r, L.V, 1.0f);
        const size_t wb = (size_t)L.V * L.ldh * sizeof(plane_t), tb = (size_t)L.H * L.ldv * sizeof(plane_t);
        CK(cudaMallocAsync(&L.w0, wb, c->stream)); CK(cudaMallocAsync(&L.w1, wb, c->stream));
        CK(cudaMallocAsync(&L.t0, tb, c->stream)); CK(cudaMallocAsync(&L.t1, tb, c->stream));
        CK(cudaMemsetAsync(L.w0, 0, wb, c->stream)); CK(cudaMemsetAsync(L.w1, 0, wb, c->stream));
        CK(cudaMemsetAsync(L.t0, 0, tb, c->stream)); CK(cudaMemsetAsync(L.t1, 0, tb, c->stream));
        refresh_nbc(c, L);
    }
    *out = m;
    API_END
}

int32_t scdbm_model_destroy(scdbm_model* m) {
    API_BEGIN
    if (!m) return SCDBM_OK;
    m->closed = true;   // particles / optimizers / trainers keep the parameters alive
    if (m->live == 0) model_free(m);
    API_END
}

int32_t scdbm_model_set_layer(scdbm_model* m, int32_t l, const double* W, int64_t ldw, const double* a, const double* b,
                              const double* r) {
    API_BEGIN
    REQUIRE(m, "model is NULL");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    Layer& L = m->L[l];
    cudaStream_t st = m->ctx->stream;
    if (W) {
        REQUIRE(ldw >= L.V, "ldw smaller than the number of visible units");
        std::vector<float> w((size_t)L.V * L.H);
        for (int h = 0; h < L.H; ++h)
            for (int v = 0; v < L.V; ++v) {
                const double x = W[(size_t)h * ldw + v];
                REQUIRE(std::isfinite(x), "weights contain NaN/Inf");
                w[(size_t)v * L.H + h] = (float)x;
            }
        CK(cudaMemcpyAsync(L.W, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        refresh_operands(m);
    }
    auto up = [&](float* dst, const double* src, int n, const char* what, bool need_pos, bool need_neg) {
        if (!src) return;
        std::vector<float> t(n);
        for (int i = 0; i < n; ++i) {
            REQUIRE(std::isfinite(src[i]), std::string(what) + " contains NaN/Inf");
            if (need_pos) REQUIRE(src[i] > 0.0, std::string(what) + " must be positive");
            if (need_neg) REQUIRE(src[i] < 0.0, std::string(what) + " must be negative for an NB visible layer");
            t[i] = (float)src[i];
        }
        CK(cudaMemcpyAsync(dst, t.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    };
    up(L.a, a, L.V, "visbias", false, L.kind == SCDBM_VIS_NEGBIN);
    up(L.b, b, L.H, "hidbias", false, false);
    if (r) {
        REQUIRE(L.kind == SCDBM_VIS_NEGBIN, "inversedispersion given for a non-NB layer");
        up(L.r, r, L.V, "inversedispersion", true, false);
    }
    if (L.kind == SCDBM_VIS_NEGBIN) {
        // validity of the NB conditional (Appendix A of SURVEY.md): x_i = a_i + sum_j W_ij h_j < 0 for every binary h,
        // i.e. a_i + sum_j max(W_ij, 0) < 0 (the reference keeps W <= 0, a <= -1e-10: src/optimizers.jl:482-483)
        // the status word lives in mapped host memory: a device->host copy would queue behind an asynchronous
        // download in flight on the copy engine
        CK(cudaStreamSynchronize(st));
        *m->ctx->h_flag = 0u;
        refresh_nbc(m->ctx, L, m->ctx->d_flag);
        CK(cudaStreamSynchronize(st));
        const unsigned int bad = *m->ctx->h_flag;
        if (bad && W && a) throw Err(SCDBM_EINVAL, "invalid NB parameters: visbias[i] + sum_j max(weights[i,j], 0) must be negative for every gene");
    }
    API_END
}

int32_t scdbm_model_get_layer(scdbm_model* m, int32_t l, double* W, int64_t ldw, double* a, double* b, double* r) {
    API_BEGIN
    REQUIRE(m, "model is NULL");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    Layer& L = m->L[l];
    cudaStream_t st = m->ctx->stream;
    CK(cudaStreamSynchronize(st));
    if (W) {
        REQUIRE(ldw >= L.V, "ldw smaller than the number of visible units");
        std::vector<float> w((size_t)L.V * L.H);
        CK(cudaMemcpy(w.data(), L.W, w.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int h = 0; h < L.H; ++h)
            for (int v = 0; v < L.V; ++v) W[(size_t)h * ldw + v] = (double)w[(size_t)v * L.H + h];
    }
    auto down = [&](double* dst, const float* src, int n) {
        if (!dst) return;
        std::vector<float> t(n);
        CK(cudaMemcpy(t.data(), src, n * sizeof(float), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) dst[i] = (double)t[i];
    };
    down(a, L.a, L.V);
    down(b, L.b, L.H);
    down(r, L.r, L.V);
    API_END
}

int32_t scdbm_model_copy_layer(scdbm_model* dst, int32_t dl, scdbm_model* src, int32_t sl) {
    API_BEGIN
    REQUIRE(dst && src, "NULL argument");
    REQUIRE(dl >= 0 && dl < (int)dst->L.size() && sl >= 0 && sl < (int)src->L.size(), "layer index out of range");
    Layer& D = dst->L[dl];
    Layer& S = src->L[sl];
    REQUIRE(D.V == S.V && D.H == S.H, "layer shapes differ");
    cudaStream_t st = dst->ctx->stream;
    CK(cudaStreamSynchronize(src->ctx->stream));
    CK(cudaMemcpyAsync(D.W, S.W, (size_t)D.V * D.H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(D.a, S.a, (size_t)D.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(D.b, S.b, (size_t)D.H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(D.r, S.r, (size_t)D.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    refresh_operands(dst);
    refresh_nbc(dst->ctx, D);
    API_END
}

__global__ void k_init_weights(float* W, int V, int H, int nb, uint32_t k0, uint32_t k1) {
    const int64_t groups = (H + 3) / 4;
    const int64_t n = (int64_t)V * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / groups, g = i % groups;
        const uint4 w = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)v, 0xFFFF0000u, stream_id(0, P_INIT)), k0, k1);
        const uint4 w2 = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)v, 0xFFFF0000u, stream_id(0, P_INIT2)), k0, k1);
        const uint32_t a[4] = {w.x, w.y, w.z, w.w}, b[4] = {w2.x, w2.y, w2.z, w2.w};
        for (int j = 0; j < 4 && g * 4 + j < H; ++j) {
            float x;
            if (nb) x = -(0.0005f + (0.002f - 0.0005f) * u23(a[j]));
            else x = sqrtf(-2.0f * logf(u23(a[j]))) * cospif(2.0f * u23(b[j])) * rsqrtf((float)V);
            W[v * H + g * 4 + j] = x;
        }
    }
}

int32_t scdbm_model_init_layer(scdbm_model* m, int32_t l, const scdbm_mat* x, const double* r, uint64_t seed) {
    API_BEGIN
    REQUIRE(m && x, "NULL argument");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    Layer& L = m->L[l];
    REQUIRE(x->v.cols == L.V, "data width does not match the visible layer");
    REQUIRE(x->v.rows > 0, "data has no rows");
    scdbm_ctx* c = m->ctx;
    const bool nb = L.kind == SCDBM_VIS_NEGBIN;
    k_init_weights<<<grid1d((int64_t)L.V * ((L.H + 3) / 4)), 256, 0, c->stream>>>(L.W, L.V, L.H, nb ? 1 : 0,
                                                                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    CK(cudaGetLastError());
    c->launches++;
    refresh_operands(m);
    // column means -> visible bias on the host (V numbers)
    float* d_sum = nullptr;
    CK(cudaMallocAsync(&d_sum, L.V * sizeof(float), c->stream));
    CK(cudaMemsetAsync(d_sum, 0, L.V * sizeof(float), c->stream));
    colsum(c, x->v, 1.0f, d_sum, false);
    std::vector<float> s(L.V);
    CK(cudaMemcpyAsync(s.data(), d_sum, L.V * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d_sum, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double> a(L.V), rr(L.V, 0.5), b(L.H, 0.0);
    if (r) for (int i = 0; i < L.V; ++i) rr[i] = r[i];
    for (int i = 0; i < L.V; ++i) {
        const double mean = (double)s[i] / (double)x->v.rows;
        if (nb) {
            // log(mean/(mean+r)); a zero-mean gene gives -Inf in the reference, kept finite here
            a[i] = std::log(std::max(mean, 1e-30) / (std::max(mean, 1e-30) + rr[i]));
            a[i] = std::min(a[i], -1e-10);
        } else a[i] = mean > 0.0 ? std::log(mean / (1.0 - mean)) : 0.0;
    }
    return scdbm_model_set_layer(m, l, nullptr, 0, a.data(), b.data(), nb ? rr.data() : nullptr);
    API_END
}

// --------------------------------------------------------------------- mat
int32_t scdbm_mat_create(scdbm_ctx* c, int64_t rows, int64_t cols, int32_t kind, scdbm_mat** out) {
    API_BEGIN
    REQUIRE(c && out, "NULL argument");
    *out = mat_new(c, rows, cols, kind);
    API_END
}
int32_t scdbm_mat_destroy(scdbm_mat* m) {
    API_BEGIN
    if (m) mat_free(m);   // stream-ordered free: safe behind any work already queued on the context's stream
    API_END
}
int32_t scdbm_mat_shape(const scdbm_mat* m, int64_t* rows, int64_t* cols, int32_t* kind) {
    API_BEGIN
    REQUIRE(m, "mat is NULL");
    if (rows) *rows = m->v.rows;
    if (cols) *cols = m->v.cols;
    if (kind) *kind = m->v.kind;
    API_END
}

int32_t scdbm_mat_upload(scdbm_mat* m, const double* x, int64_t ld) {
    API_BEGIN
    REQUIRE(m && x, "NULL argument");
    REQUIRE(ld >= m->v.rows, "leading dimension smaller than the row count");
    if (m->v.rows == 0 || m->v.cols == 0) return SCDBM_OK;
    scdbm_ctx* c = m->ctx;
    // validate on the host: NaN never enters; counts / binary states must be valid
    for (int64_t j = 0; j < m->v.cols; ++j)
        for (int64_t i = 0; i < m->v.rows; ++i) {
            const double v = x[j * ld + i];
            if (!std::isfinite(v)) throw Err(SCDBM_EINVAL, "matrix contains NaN/Inf");
            if (m->v.kind == MAT_COUNT && (v < 0.0 || v > 65535.0 || v != std::floor(v)))
                throw Err(SCDBM_EINVAL, "count matrix entries must be integers in [0, 65535]");
            if (m->v.kind == MAT_BINARY && v != 0.0 && v != 1.0)
                throw Err(SCDBM_EINVAL, "binary matrix entries must be 0 or 1");
        }
    // column-block staging keeps the fp64 staging buffer small
    const int64_t colblk = std::max<int64_t>(32, std::min<int64_t>(m->v.cols, (int64_t)(256ll << 20) / (8 * std::max<int64_t>(ld, 1)) / 32 * 32));
    double* stage = nullptr;
    CK(cudaMallocAsync(&stage, (size_t)ld * colblk * sizeof(double), c->stream));
    for (int64_t c0 = 0; c0 < m->v.cols; c0 += colblk) {
        const int64_t nc = std::min(colblk, m->v.cols - c0);
        CK(cudaMemcpyAsync(stage, x + c0 * ld, (size_t)ld * nc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        MatView sub = m->v;
        sub.p0 += c0;
        if (sub.p1) sub.p1 += c0;
        if (sub.f) sub.f += c0;
        sub.cols = nc;
        dim3 grid((unsigned)((m->v.rows + 31) / 32), (unsigned)((nc + 31) / 32)), block(32, 8);
        if (c0 == 0) set_hiflags(c, m->v, false);   // the upload kernel raises the flag of a row tile that gets a count >= 256
        k_upload_f64<<<grid, block, 0, c->stream>>>(stage, ld, sub);
        CK(cudaGetLastError());
        c->launches++;
        CK(cudaStreamSynchronize(c->stream));
    }
    cudaFreeAsync(stage, c->stream);
    API_END
}

int32_t scdbm_mat_download(const scdbm_mat* m, double* x, int64_t ld) {
    API_BEGIN
    REQUIRE(m && x, "NULL argument");
    REQUIRE(ld >= m->v.rows, "leading dimension smaller than the row count");
    if (m->v.rows == 0 || m->v.cols == 0) return SCDBM_OK;
    scdbm_ctx* c = m->ctx;
    const int64_t colblk = std::max<int64_t>(32, std::min<int64_t>(m->v.cols, (int64_t)(256ll << 20) / (8 * std::max<int64_t>(ld, 1)) / 32 * 32));
    double* stage = nullptr;
    CK(cudaMallocAsync(&stage, (size_t)ld * colblk * sizeof(double), c->stream));
    CK(cudaMemsetAsync(stage, 0, (size_t)ld * colblk * sizeof(double), c->stream));
    for (int64_t c0 = 0; c0 < m->v.cols; c0 += colblk) {
        const int64_t nc = std::min(colblk, m->v.cols - c0);
        MatView sub = m->v;
        sub.p0 += c0;
        if (sub.p1) sub.p1 += c0;
        if (sub.f) sub.f += c0;
        sub.cols = nc;
        dim3 grid((unsigned)((m->v.rows + 31) / 32), (unsigned)((nc + 31) / 32)), block(32, 8);
        k_download_f64<<<grid, block, 0, c->stream>>>(sub, stage, ld);
        CK(cudaGetLastError());
        c->launches++;
        if (ld == m->v.rows) {
            CK(cudaMemcpyAsync(x + c0 * ld, stage, (size_t)ld * nc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        } else {
            CK(cudaMemcpy2DAsync(x + c0 * ld, ld * sizeof(double), stage, ld * sizeof(double), m->v.rows * sizeof(double), nc,
                                 cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
    }
    cudaFreeAsync(stage, c->stream);
    API_END
}

int32_t scdbm_mat_download_u16_async(const scdbm_mat* m, uint16_t* x);
int32_t scdbm_wait_downloads(scdbm_ctx* ctx);

// synchronous form: the asynchronous snapshot + copy-engine drain below (context-owned stream, events and snapshot
// buffer: nothing is created or allocated per call once the buffer has grown), then wait
int32_t scdbm_mat_download_u16(const scdbm_mat* m, uint16_t* x) {
    REQUIRE_OR_RETURN(m && x);
    const int32_t st = scdbm_mat_download_u16_async(m, x);
    if (st != SCDBM_OK) return st;
    return scdbm_wait_downloads(m->ctx);
}

int32_t scdbm_mat_download_u16_async(const scdbm_mat* m, uint16_t* x) {
    API_BEGIN
    REQUIRE(m && x, "NULL argument");
    const int64_t n = m->v.rows * m->v.cols;
    if (n == 0) return SCDBM_OK;
    scdbm_ctx* c = m->ctx;
    if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_conv, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    }
    if (c->copy_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));   // the snapshot buffer is still being drained
    const size_t bytes = (size_t)n * sizeof(uint16_t);
    if (c->stage16_cap < bytes) {
        CK(cudaStreamSynchronize(c->copy_stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->stage16);
        c->stage16 = nullptr;
        c->stage16_cap = 0;
        if (cudaMalloc(&c->stage16, bytes) != cudaSuccess) throw Err(SCDBM_ENOMEM, "download snapshot buffer");
        c->stage16_cap = bytes;
    }
    // the snapshot is taken in stream order, so the caller may overwrite the matrix (next batch of chains) right away
    k_download_u16<<<(unsigned)((m->v.rows + 7) / 8), dim3(32, 8), 0, c->stream>>>(m->v, c->stage16);
    CK(cudaGetLastError());
    c->launches++;
    CK(cudaEventRecord(c->ev_conv, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_conv, 0));
    // 32 MB pieces: small device->host read-backs of later calls (status flags, scalars) share the copy engine
    // and would otherwise queue behind the whole transfer
    for (size_t off = 0; off < bytes; off += (size_t)32 << 20)
        CK(cudaMemcpyAsync(reinterpret_cast<char*>(x) + off, reinterpret_cast<const char*>(c->stage16) + off,
                           std::min<size_t>((size_t)32 << 20, bytes - off), cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaEventRecord(c->ev_copy, c->copy_stream));
    c->copy_pending = true;
    API_END
}

int32_t scdbm_wait_downloads(scdbm_ctx* ctx) {
    API_BEGIN
    REQUIRE(ctx, "ctx is NULL");
    if (ctx->copy_pending) {
        CK(cudaStreamSynchronize(ctx->copy_stream));
        ctx->copy_pending = false;
    }
    API_END
}

// ------------------------------------------------------------------ sparse data path
int32_t scdbm_mat_upload_compressed(scdbm_mat* m, int32_t major, const int64_t* ptr, const int32_t* idx, const float* val) {
    API_BEGIN
    REQUIRE(m && ptr, "NULL argument");
    REQUIRE(major == 0 || major == 1, "major must be 0 (CSR) or 1 (CSC)");
    scdbm_ctx* c = m->ctx;
    const int64_t nmajor = major ? m->v.cols : m->v.rows;
    const int64_t nnz = ptr[nmajor];
    REQUIRE(ptr[0] == 0 && nnz >= 0, "pointer array must start at 0 and end at nnz");
    for (int64_t i = 0; i < nmajor; ++i) REQUIRE(ptr[i] <= ptr[i + 1], "pointer array must be non-decreasing");
    REQUIRE(nnz == 0 || (idx && val), "NULL index/value array");
    if (m->v.rows == 0 || m->v.cols == 0) return SCDBM_OK;
    const size_t plane = (size_t)m->v.rows * m->v.ld * sizeof(plane_t);
    CK(cudaMemsetAsync(m->v.p0, 0, plane, c->stream));
    if (m->v.p1) CK(cudaMemsetAsync(m->v.p1, 0, plane, c->stream));
    if (m->v.f) CK(cudaMemsetAsync(m->v.f, 0, plane * 2, c->stream));
    set_hiflags(c, m->v, false);
    if (nnz == 0) return SCDBM_OK;
    int64_t* d_ptr = nullptr;
    int32_t* d_idx = nullptr;
    float* d_val = nullptr;
    unsigned int* d_err = nullptr;
    cudaStream_t st = c->stream;
    auto cleanup = [&]() { for (void* q : {(void*)d_ptr, (void*)d_idx, (void*)d_val, (void*)d_err}) if (q) cudaFreeAsync(q, st); };
    try {
        CK(cudaMallocAsync(&d_ptr, (size_t)(nmajor + 1) * sizeof(int64_t), st));
        CK(cudaMallocAsync(&d_idx, (size_t)nnz * sizeof(int32_t), st));
        CK(cudaMallocAsync(&d_val, (size_t)nnz * sizeof(float), st));
        CK(cudaMallocAsync(&d_err, sizeof(unsigned int), st));
        CK(cudaMemsetAsync(d_err, 0, sizeof(unsigned int), c->stream));
        CK(cudaMemcpyAsync(d_ptr, ptr, (size_t)(nmajor + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_idx, idx, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_val, val, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        k_scatter_compressed<<<grid1d(nmajor * 32), 256, 0, c->stream>>>(d_ptr, d_idx, d_val, nmajor, major, m->v, d_err);
        CK(cudaGetLastError());
        c->launches++;
        unsigned int err = 0;
        CK(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (err & 1u) throw Err(SCDBM_EINVAL, "compressed matrix holds an index outside the matrix");
        if (err & 2u) throw Err(SCDBM_EINVAL, m->v.kind == MAT_COUNT ? "count matrix entries must be integers in [0, 65535]"
                                              : m->v.kind == MAT_BINARY ? "binary matrix entries must be 0 or 1" : "matrix contains NaN/Inf");
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
    API_END
}

int32_t scdbm_mat_csr_prepare(scdbm_mat* m, int64_t* nnz) {
    API_BEGIN
    REQUIRE(m && nnz, "NULL argument");
    REQUIRE(m->v.kind != MAT_REAL, "CSR download is for count / binary matrices");
    scdbm_ctx* c = m->ctx;
    const int64_t R = m->v.rows;
    if (m->csr_indptr) { CK(cudaFreeAsync(m->csr_indptr, c->stream)); m->csr_indptr = nullptr; }
    m->csr_nnz = -1;
    CK(cudaMallocAsync(&m->csr_indptr, (size_t)(R + 1) * sizeof(int64_t), c->stream));
    const int64_t nblocks = std::max<int64_t>(1, (R + 1023) / 1024);
    int64_t *d_cnt = nullptr, *d_tot = nullptr;
    CK(cudaMallocAsync(&d_cnt, (size_t)std::max<int64_t>(R, 1) * sizeof(int64_t), c->stream));
    CK(cudaMallocAsync(&d_tot, (size_t)(nblocks + 1) * sizeof(int64_t), c->stream));
    if (R > 0 && m->v.cols > 0) k_row_nnz<<<grid1d(R * 32), 256, 0, c->stream>>>(m->v, d_cnt);
    else if (R > 0) CK(cudaMemsetAsync(d_cnt, 0, (size_t)R * sizeof(int64_t), c->stream));
    k_scan_block<<<(unsigned)nblocks, 1024, 0, c->stream>>>(d_cnt, R, m->csr_indptr, d_tot);
    k_scan_totals<<<1, 1024, 0, c->stream>>>(d_tot, nblocks, d_tot + nblocks);
    k_scan_add<<<(unsigned)nblocks, 1024, 0, c->stream>>>(m->csr_indptr, R, d_tot, d_tot + nblocks);
    CK(cudaGetLastError());
    c->launches += 4;
    int64_t total = 0;
    CK(cudaMemcpyAsync(&total, d_tot + nblocks, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d_cnt, c->stream));
    CK(cudaFreeAsync(d_tot, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    m->csr_nnz = total;
    *nnz = total;
    API_END
}

int32_t scdbm_mat_download_csr(scdbm_mat* m, int64_t* indptr, void* indices, int32_t index_bytes, uint16_t* values) {
    API_BEGIN
    REQUIRE(m && indptr, "NULL argument");
    REQUIRE(m->csr_indptr && m->csr_nnz >= 0, "call scdbm_mat_csr_prepare first");
    REQUIRE(index_bytes == 2 || index_bytes == 4, "index_bytes must be 2 or 4");
    REQUIRE(index_bytes == 4 || m->v.cols <= 65536, "16-bit column indices need at most 65536 columns");
    const int64_t nnz = m->csr_nnz, R = m->v.rows;
    REQUIRE(nnz == 0 || (indices && values), "NULL index/value array");
    scdbm_ctx* c = m->ctx;
    void* d_idx = nullptr;
    uint16_t* d_val = nullptr;
    if (nnz > 0) {
        CK(cudaMallocAsync(&d_idx, (size_t)nnz * index_bytes, c->stream));
        CK(cudaMallocAsync(&d_val, (size_t)nnz * sizeof(uint16_t), c->stream));
        if (index_bytes == 2) k_csr_fill<uint16_t><<<grid1d(R * 32), 256, 0, c->stream>>>(m->v, m->csr_indptr, (uint16_t*)d_idx, d_val);
        else k_csr_fill<int32_t><<<grid1d(R * 32), 256, 0, c->stream>>>(m->v, m->csr_indptr, (int32_t*)d_idx, d_val);
        CK(cudaGetLastError());
        c->launches++;
        CK(cudaMemcpyAsync(indices, d_idx, (size_t)nnz * index_bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(values, d_val, (size_t)nnz * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaMemcpyAsync(indptr, m->csr_indptr, (size_t)(R + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    if (d_idx) CK(cudaFreeAsync(d_idx, c->stream));
    if (d_val) CK(cudaFreeAsync(d_val, c->stream));
    CK(cudaFreeAsync(m->csr_indptr, c->stream));
    m->csr_indptr = nullptr;
    m->csr_nnz = -1;
    CK(cudaStreamSynchronize(c->stream));
    API_END
}

int32_t scdbm_mat_generate_counts(scdbm_mat* m, const double* gene_mean, const double* r, double sizefactor_sigma, uint64_t seed,
                                  uint64_t row_offset) {
    API_BEGIN
    REQUIRE(m && gene_mean && r, "NULL argument");
    REQUIRE(m->v.kind == MAT_COUNT, "synthetic counts need a COUNT matrix");
    REQUIRE(sizefactor_sigma >= 0.0, "size-factor sigma must be non-negative");
    const int64_t V = m->v.cols;
    if (m->v.rows == 0 || V == 0) return SCDBM_OK;
    scdbm_ctx* c = m->ctx;
    std::vector<float> h(2 * (size_t)V);
    for (int64_t j = 0; j < V; ++j) {
        REQUIRE(std::isfinite(gene_mean[j]) && gene_mean[j] >= 0.0, "gene means must be finite and non-negative");
        REQUIRE(std::isfinite(r[j]) && r[j] > 0.0, "inverse dispersions must be positive");
        h[j] = (float)gene_mean[j];
        h[V + j] = (float)r[j];
    }
    float* d = nullptr;
    CK(cudaMallocAsync(&d, h.size() * sizeof(float), c->stream));
    CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    set_hiflags(c, m->v, false);
    k_generate_counts<<<grid1d(m->v.rows * ((V + 3) / 4)), 256, 0, c->stream>>>(m->v, d, d + V, (float)sizefactor_sigma, c->mu_inv_max,
                                                                                 (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)row_offset);
    CK(cudaGetLastError());
    c->launches++;
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // h goes out of scope
    API_END
}

int32_t scdbm_mat_copy(scdbm_mat* dst, const scdbm_mat* src) {
    API_BEGIN
    REQUIRE(dst && src, "NULL argument");
    REQUIRE(dst->v.rows == src->v.rows && dst->v.cols == src->v.cols, "shapes differ");
    REQUIRE(dst->v.kind == src->v.kind, "kinds differ");
    if (dst->v.rows == 0) return SCDBM_OK;
    set_hiflags(dst->ctx, dst->v, true);
    CK(launch_pdl(k_copy_rows, dim3((unsigned)dst->v.rows), dim3(128), 0, dst->ctx->stream, src->v, dst->v, dst->v.rows));
    CK(cudaGetLastError());
    dst->ctx->launches++;
    API_END
}

int32_t scdbm_mat_column_stats(const scdbm_mat* m, double* mean, double* var, double* zerofrac) {
    API_BEGIN
    REQUIRE(m, "mat is NULL");
    const int64_t C = m->v.cols, R = m->v.rows;
    REQUIRE(R > 0 && C > 0, "matrix is empty");
    scdbm_ctx* c = m->ctx;
    double* d = nullptr;
    CK(cudaMallocAsync(&d, 3 * C * sizeof(double), c->stream));
    CK(cudaMemsetAsync(d, 0, 3 * C * sizeof(double), c->stream));
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)std::min<int64_t>((R + 7) / 8, 256)), block(32, 8);
    k_colstats<<<grid, block, 0, c->stream>>>(m->v, d, d + C, d + 2 * C);
    c->launches++;
    std::vector<double> h(3 * C);
    CK(cudaMemcpyAsync(h.data(), d, 3 * C * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int64_t j = 0; j < C; ++j) {
        const double mu = h[j] / R;
        if (mean) mean[j] = mu;
        if (var) var[j] = R > 1 ? (h[C + j] - R * mu * mu) / (R - 1) : 0.0;
        if (zerofrac) zerofrac[j] = h[2 * C + j] / R;
    }
    API_END
}

// ------------------------------------------------- single conditional ops
int32_t scdbm_hidden(scdbm_model* m, int32_t l, const scdbm_mat* v, scdbm_mat* out, double factor, int32_t what,
                     uint32_t tick, const double* uniforms, int64_t ldu) {
    API_BEGIN
    REQUIRE(m && v && out, "NULL argument");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    REQUIRE(out->v.rows == v->v.rows, "row counts differ");
    DevUniforms du;
    upload_uniforms(m->ctx, uniforms, ldu, out->v.rows, out->v.cols, du);
    DrawCfg d;
    d.stream = stream_id(l + 1); d.tick = tick; d.uniforms = du.d; d.ldu = du.ld;
    op_hidden(m, l, v->v, out->v, (float)factor, what, d);
    if (du.d) CK(cudaStreamSynchronize(m->ctx->stream));
    API_END
}

int32_t scdbm_visible(scdbm_model* m, int32_t l, const scdbm_mat* h, scdbm_mat* out, double factor, int32_t what,
                      uint32_t tick, const double* uniforms, int64_t ldu) {
    API_BEGIN
    REQUIRE(m && h && out, "NULL argument");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    REQUIRE(out->v.rows == h->v.rows, "row counts differ");
    DevUniforms du;
    upload_uniforms(m->ctx, uniforms, ldu, out->v.rows, out->v.cols, du);
    DrawCfg d;
    d.stream = stream_id(l); d.tick = tick; d.uniforms = du.d; d.ldu = du.ld;
    op_visible(m, l, h->v, out->v, (float)factor, what, d);
    if (du.d) CK(cudaStreamSynchronize(m->ctx->stream));
    API_END
}

int32_t scdbm_dbm_layer(scdbm_model* m, int32_t layer, const scdbm_mat* below, const scdbm_mat* above, scdbm_mat* out,
                        int32_t what, uint32_t tick, const double* uniforms, int64_t ldu) {
    API_BEGIN
    REQUIRE(m && below && above && out, "NULL argument");
    REQUIRE(layer >= 1 && layer <= (int)m->L.size() - 1, "not an intermediate layer");
    REQUIRE(below->v.rows == above->v.rows && out->v.rows == above->v.rows, "row counts differ");
    DevUniforms du;
    upload_uniforms(m->ctx, uniforms, ldu, out->v.rows, out->v.cols, du);
    DrawCfg d;
    d.stream = stream_id(layer); d.tick = tick; d.uniforms = du.d; d.ldu = du.ld;
    op_dbm_layer(m, layer, &below->v, nullptr, 0, above->v, out->v, what, d, nullptr, nullptr);
    if (du.d) CK(cudaStreamSynchronize(m->ctx->stream));
    API_END
}

// --------------------------------------------------------------- particles
int32_t scdbm_particles_create(scdbm_model* m, int64_t n, uint64_t row_offset, int32_t real_valued, scdbm_particles** out) {
    API_BEGIN
    REQUIRE(m && out, "NULL argument");
    REQUIRE(n >= 0, "nparticles must be non-negative");
    REQUIRE(row_offset + (uint64_t)n <= 0xFFFFFFFFull, "global particle index exceeds 2^32");
    auto* p = new scdbm_particles{m, n, row_offset, 1, real_valued != 0, {}, {}};
    model_retain(m);
    const int nl = m->nlayers();
    for (int i = 0; i < nl; ++i) {
        int kind = i == 0 ? (m->L[0].kind == SCDBM_VIS_NEGBIN ? MAT_COUNT : MAT_BINARY) : MAT_BINARY;
        if (real_valued && i > 0) kind = MAT_REAL;
        if (real_valued && i == 0) { p->cur.push_back(nullptr); p->nxt.push_back(nullptr); continue; }
        p->cur.push_back(mat_new(m->ctx, n, m->units(i), kind));
        p->nxt.push_back(real_valued ? nullptr : mat_new(m->ctx, n, m->units(i), kind));
    }
    *out = p;
    API_END
}

int32_t scdbm_particles_destroy(scdbm_particles* p) {
    API_BEGIN
    if (!p) return SCDBM_OK;
    for (auto* x : p->cur) mat_free(x);      // stream-ordered frees: work queued on the context's stream still sees the buffers
    for (auto* x : p->nxt) mat_free(x);
    mat_free(p->alias0);
    if (p->hoist) cudaFreeAsync(p->hoist, p->model->ctx->stream);
    scdbm_model* pm = p->model;
    delete p;
    model_release(pm);
    API_END
}

int32_t scdbm_particles_init(scdbm_particles* p, int32_t biased) {
    API_BEGIN
    REQUIRE(p, "particles is NULL");
    REQUIRE(!p->real_valued, "mean-field particles are initialised by scdbm_meanfield");
    scdbm_model* m = p->model;
    scdbm_ctx* c = m->ctx;
    const uint32_t k0 = (uint32_t)c->seed, k1 = (uint32_t)(c->seed >> 32);
    for (int i = 0; i < m->nlayers(); ++i) {
        MatView& v = p->cur[i]->v;
        if (v.rows == 0) continue;
        int mode;
        const float* bias = nullptr;
        const float* r = nullptr;
        if (i == 0) {
            if (m->L[0].kind == SCDBM_VIS_NEGBIN) { mode = 3; r = m->L[0].r; }
            else { mode = biased ? 1 : 0; bias = m->L[0].a; }
        } else { mode = biased ? 1 : 0; bias = m->L[i - 1].b; }
        set_hiflags(c, v, false);   // k_init rewrites both planes and raises the flag of a row tile that gets a count >= 2048
        k_init<<<grid1d(v.rows * ((v.cols + 3) / 4)), 256, 0, c->stream>>>(v, mode, bias, r, k0, k1, (uint32_t)i,
                                                                            (uint32_t)p->row_offset);
        CK(cudaGetLastError());
        c->launches++;
    }
    p->clock = 1;
    API_END
}

int32_t scdbm_particles_layer(scdbm_particles* p, int32_t layer, scdbm_mat** borrowed) {
    API_BEGIN
    REQUIRE(p && borrowed, "NULL argument");
    REQUIRE(layer >= 0 && layer < (int)p->cur.size(), "layer index out of range");
    scdbm_mat* x = (layer == 0 && p->real_valued) ? p->alias0 : p->cur[layer];
    REQUIRE(x, "layer 0 of mean-field particles is set by scdbm_meanfield");
    *borrowed = mat_retain(x);
    API_END
}

int32_t scdbm_particles_clock(scdbm_particles* p, int64_t* clock, int64_t set_to) {
    API_BEGIN
    REQUIRE(p, "particles is NULL");
    if (set_to >= 0) p->clock = set_to;
    if (clock) *clock = p->clock;
    API_END
}

static void gibbs_dbm_steps(scdbm_model* m, scdbm_particles* p, int nsteps, float temperature) {
    const int nl = m->nlayers();
    for (int step = 0; step < nsteps; ++step) {
        DrawCfg d;
        d.tick = (uint32_t)p->clock;
        d.row_offset = (uint32_t)p->row_offset;
        d.wscale = temperature;
        d.stream = stream_id(0);
        op_visible(m, 0, p->cur[1]->v, p->nxt[0]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
        for (int i = 1; i < nl - 1; ++i) {
            d.stream = stream_id(i);
            op_dbm_layer(m, i, &p->cur[i - 1]->v, nullptr, 0, p->cur[i + 1]->v, p->nxt[i]->v, SCDBM_OUT_SAMPLE, d, nullptr, nullptr);
        }
        d.stream = stream_id(nl - 1);
        op_hidden(m, nl - 2, p->cur[nl - 2]->v, p->nxt[nl - 1]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
        std::swap(p->cur, p->nxt);
        p->clock++;
    }
}

int32_t scdbm_gibbs(scdbm_model* m, scdbm_particles* p, int32_t nsteps, double temperature, double upfactor,
                    double downfactor) {
    API_BEGIN
    REQUIRE(m && p, "NULL argument");
    REQUIRE(p->model == m, "particles belong to a different model");
    REQUIRE(!p->real_valued, "cannot Gibbs-sample mean-field particles");
    REQUIRE(nsteps >= 0, "nsteps must be non-negative");
    if (p->n == 0) return SCDBM_OK;
    if (m->L.size() == 1) {
        // gibbssample!(particles, rbm, nsteps, up, down): v | h, then h | new v
        for (int s = 0; s < nsteps; ++s) {
            DrawCfg d;
            d.row_offset = (uint32_t)p->row_offset;
            d.wscale = (float)temperature;
            d.tick = (uint32_t)p->clock; d.stream = stream_id(0);
            op_visible(m, 0, p->cur[1]->v, p->cur[0]->v, (float)downfactor, SCDBM_OUT_SAMPLE, d);
            p->clock++;
            d.tick = (uint32_t)p->clock; d.stream = stream_id(1);
            op_hidden(m, 0, p->cur[0]->v, p->cur[1]->v, (float)upfactor, SCDBM_OUT_SAMPLE, d);
            p->clock++;
        }
    } else {
        gibbs_dbm_steps(m, p, nsteps, (float)temperature);
    }
    API_END
}

int32_t scdbm_gibbs_negbin_rbm(scdbm_model* m, scdbm_particles* p, int32_t nsteps, double upfactor, double downfactor) {
    API_BEGIN
    REQUIRE(m && p, "NULL argument");
    REQUIRE(p->model == m && m->L.size() == 1, "needs particles of a single-RBM model");
    REQUIRE(m->L[0].kind == SCDBM_VIS_NEGBIN, "model is not an NB-Bernoulli RBM");
    if (p->n == 0) return SCDBM_OK;
    for (int s = 0; s < nsteps; ++s) {
        DrawCfg d;
        d.row_offset = (uint32_t)p->row_offset;
        d.tick = (uint32_t)p->clock;
        d.stream = stream_id(0);
        d.pshift = 1e-6f;
        op_visible(m, 0, p->cur[1]->v, p->nxt[0]->v, (float)downfactor, SCDBM_OUT_SAMPLE, d);
        d.stream = stream_id(1);
        d.pshift = -1.0f;
        op_hidden(m, 0, p->cur[0]->v, p->nxt[1]->v, (float)upfactor, SCDBM_OUT_SAMPLE, d);
        std::swap(p->cur, p->nxt);
        p->clock++;
    }
    API_END
}

int32_t scdbm_particles_visible_potential(scdbm_model* m, scdbm_particles* p, scdbm_mat* out) {
    API_BEGIN
    REQUIRE(m && p && out, "NULL argument");
    DrawCfg d;
    op_visible(m, 0, p->cur[1]->v, out->v, 1.0f, SCDBM_OUT_POTENTIAL, d);
    API_END
}

// -------------------------------------------------------------- mean-field
// after one mean-field iteration: fold this iteration's max |change| into the convergence state.  One thread.
//   st[0] delta bits of the iteration (reset), st[1] converged flag, st[2] iterations run, st[3] delta bits of the last one,
//   st[4] number of checks executed in this mean-field call (skipped iterations included)
// The state is mirrored into mapped host memory as two 64-bit words: {iterations, delta bits} first, then -- after a
// system-wide fence -- {generation of the call, checks executed, converged}.  The host polls the second word instead of
// synchronising the stream: it can move on the moment convergence is visible while the speculatively launched
// iterations behind it drain as no-ops; the generation tag keeps a late check of the previous call from being
// mistaken for this one's.
__global__ void k_mf_check(unsigned int* st, float eps, unsigned long long* host_mirror, unsigned int gen) {
    pdl_trigger();
    pdl_wait();
    if (st[1] == 0u) {
        const float delta = __uint_as_float(st[0]);
        st[2] += 1u;
        st[3] = st[0];
        if (!(delta > eps)) st[1] = 1u;
    }
    st[0] = 0u;
    st[4] += 1u;
    volatile unsigned long long* hm = host_mirror;
    hm[1] = ((unsigned long long)st[2] << 32) | (unsigned long long)st[3];
    __threadfence_system();
    hm[0] = ((unsigned long long)(gen & 0xFFFFu) << 48) | ((unsigned long long)st[4] << 8) | (unsigned long long)(st[1] & 1u);
}

static void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw Err(SCDBM_ECUDA, std::string(what) + ": " + nccl_api().GetErrorString(r));
}

// after_first_group: work that does not depend on the mean-field result, queued behind the first group of iterations
// and before the host looks at the stopping rule -- while the host then decides and queues what follows, the GPU
// still has that work to do (traindbm!: the Gibbs sweeps over the persistent particles)
static void meanfield_impl(scdbm_model* m, const scdbm_mat* x, double eps, int maxiter, scdbm_particles* mu, int* niter,
                           double* delta_out, const std::function<void()>& after_first_group = nullptr) {
    scdbm_ctx* c = m->ctx;
    const int nl = m->nlayers();
    const int64_t n = x->v.rows;
    REQUIRE(mu->real_valued, "mu must be created with real_valued = 1");
    REQUIRE(mu->n == n, "mu must have as many particles as x has rows");
    REQUIRE(x->v.cols == m->L[0].V, "data width does not match the visible layer");
    if (!mu->alias0 || mu->alias0->parent != x) {
        mat_free(mu->alias0);   // handles given out for the previous data matrix stay valid (they hold references)
        mu->alias0 = mat_view_of(const_cast<scdbm_mat*>(x));
    }
    mu->alias0->v = x->v;
    // data-parallel: every rank iterates (an empty shard too) -- the stopping rule is the GLOBAL max over all rows
    // (src/dbmtraining.jl:194-225), one scalar allreduce(max) per iteration
    const bool global_delta = c->comm != nullptr;
    if (n == 0 && !global_delta) { if (niter) *niter = 0; if (delta_out) *delta_out = 0; return; }
    // hoist x*W1 (loop invariant; the reference recomputes it, quirk B10)
    Layer& L0 = m->L[0];
    if (!mu->hoist && n > 0) CK(cudaMallocAsync(&mu->hoist, (size_t)n * L0.H * sizeof(float), c->stream));
    if (n > 0) {
        EpiParams e = epi_base(c);
        e.mode = EPI_RAW_F32;
        e.out_f32 = mu->hoist;
        e.ldo = L0.H;
        SweepSrc s{&x->v, &L0, true};
        run_sweep(c, &s, 1, n, L0.H, e);
    }
    const unsigned int* skip = nullptr;
    auto from_hoist = [&](const MatView& out, float factor, const MatView* prev) {
        // sigm(factor * (x W1 + b1)) with no contraction
        EpiParams e = epi_base(c);
        e.mode = EPI_SIGM;
        e.bias0 = L0.b;
        e.addend = mu->hoist; e.ld_add = L0.H;
        e.factor = factor;
        e.out = out;
        e.skip = skip;
        if (prev) { e.prev = *prev; e.delta_bits = c->d_mf; }
        run_sweep(c, nullptr, 0, n, L0.H, e);
    };
    // bottom-up initialisation with doubled weights except for the top layer (dbmtraining.jl:182-186)
    DrawCfg d;
    for (int i = 1; i < nl && n > 0; ++i) {
        const float factor = (i < nl - 1) ? 2.0f : 1.0f;
        if (i == 1) from_hoist(mu->cur[1]->v, factor, nullptr);
        else op_hidden(m, i - 1, mu->cur[i - 1]->v, mu->cur[i]->v, factor, SCDBM_OUT_POTENTIAL, d);
    }
    // Fixed-point iterations.  The stopping rule lives on the device: k_mf_check raises a flag once the max |change| of an
    // iteration is <= eps, and the sweeps of iterations launched after that are no-ops.  Iterations are launched in groups
    // of MF_GROUP without any host round trip; the host looks at the (mapped) state once per group.
    // The reference loops without a cap; fp32 round-off can keep delta above a tiny eps forever, so stop after MF_ITER_CAP
    // sweeps and report the delta reached (the caller sees delta > eps).
    const int MF_ITER_CAP = 10000, MF_GROUP = 4;
    const int cap = maxiter > 0 ? maxiter : MF_ITER_CAP;
    CK(cudaMemsetAsync(c->d_mf, 0, 8 * sizeof(unsigned int), c->stream));
    const unsigned int gen = (c->mf_gen = (c->mf_gen % 0xFFFFu) + 1u);     // 1 .. 65535
    volatile unsigned long long* hm = reinterpret_cast<volatile unsigned long long*>(c->h_mf);
    unsigned long long state = 0ull;
    skip = c->d_mf + 1;
    int launched = 0;
    bool queued_extra = false;
    bool done = !(1.0 > eps);   // `delta = 1.0; while delta > eps` (dbmtraining.jl:194)
    float last_delta = done ? 1.0f : 0.0f;
    while (!done && launched < cap) {
        const int group = std::min(MF_GROUP, cap - launched);
        for (int gi = 0; gi < group; ++gi) {
            if (n > 0) {
                for (int i = 1; i < nl - 1; ++i) {
                    const MatView& cur = mu->cur[i]->v;
                    // intermediate layers from both neighbours (Gauss-Seidel over layers, dbmtraining.jl:197-207)
                    Layer& Lb = m->L[i - 1];
                    Layer& La = m->L[i];
                    EpiParams e = epi_base(c);
                    e.mode = EPI_SIGM;
                    e.bias0 = Lb.b;
                    e.bias1 = La.a;
                    e.out = cur;
                    e.prev = cur;
                    e.delta_bits = c->d_mf;
                    e.skip = skip;
                    SweepSrc srcs[2];
                    int ns = 0;
                    if (i == 1) { e.addend = mu->hoist; e.ld_add = L0.H; }
                    else srcs[ns++] = SweepSrc{&mu->cur[i - 1]->v, &Lb, true};
                    srcs[ns++] = SweepSrc{&mu->cur[i + 1]->v, &La, false};
                    run_sweep(c, srcs, ns, n, Lb.H, e);
                }
                // last layer from the NEW layer below (Gauss-Seidel, dbmtraining.jl:218-224)
                const MatView& cur = mu->cur[nl - 1]->v;
                if (nl == 2) from_hoist(cur, 1.0f, &cur);
                else {
                    Layer& L = m->L[nl - 2];
                    EpiParams e = epi_base(c);
                    e.mode = EPI_SIGM;
                    e.bias0 = L.b;
                    e.out = cur;
                    e.prev = cur;
                    e.delta_bits = c->d_mf;
                    e.skip = skip;
                    SweepSrc s{&mu->cur[nl - 2]->v, &L, true};
                    run_sweep(c, &s, 1, n, L.H, e);
                }
            }
            // non-negative floats order like their bit patterns: max over ranks of the uint32 word
            if (global_delta) nccl_check(nccl_api().AllReduce(c->d_mf, c->d_mf, 1, ncclUint32, ncclMax, c->comm, c->stream), "ncclAllReduce(max delta)");
            CK(launch_pdl(k_mf_check, dim3(1), dim3(1), 0, c->stream, c->d_mf, (float)eps, reinterpret_cast<unsigned long long*>(c->d_hmf), gen));
            c->launches++;
        }
        launched += group;
        if (launched == group && after_first_group) { after_first_group(); queued_extra = true; }
        // poll the mapped state word: converged, or every check of this group executed
        for (unsigned int spins = 1;; ++spins) {
            const unsigned long long w = hm[0];
            if ((w >> 48) == gen && ((w & 1ull) || ((w >> 8) & 0xFFFFFFFFull) >= (unsigned long long)launched)) break;
            if ((spins & 0x3FFu) == 0u) {      // a failed launch must not hang the host
                const cudaError_t q = cudaStreamQuery(c->stream);
                if (q == cudaSuccess) break;   // drained: the word below is final
                if (q != cudaErrorNotReady) CK(q);
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        done = (hm[0] & 1ull) != 0ull;
        state = hm[1];
    }
    if (!queued_extra && after_first_group) after_first_group();     // no iteration was launched at all
    const unsigned int delta_bits = (unsigned int)(state & 0xFFFFFFFFull);
    if (launched > 0) memcpy(&last_delta, &delta_bits, sizeof(float));
    if (niter) *niter = (int)(state >> 32);
    if (delta_out) *delta_out = last_delta;
}

int32_t scdbm_meanfield(scdbm_model* m, const scdbm_mat* x, double eps, int32_t maxiter, scdbm_particles* mu,
                        int32_t* niter, double* delta) {
    API_BEGIN
    REQUIRE(m && x && mu, "NULL argument");
    REQUIRE(mu->model == m, "mu belongs to a different model");
    int it = 0;
    meanfield_impl(m, x, eps, maxiter, mu, &it, delta);
    if (niter) *niter = it;
    API_END
}

// ---------------------------------------------------------------- learning
int32_t scdbm_opt_create(scdbm_model* m, scdbm_opt** out) {
    API_BEGIN
    REQUIRE(m && out, "NULL argument");
    auto* o = new scdbm_opt{m};
    model_retain(m);
    int64_t off = 0;
    for (auto& L : m->L) {
        o->offW.push_back(off); off += (int64_t)L.V * L.H;
        o->offA.push_back(off); off += L.V;
        o->offB.push_back(off); off += L.H;
    }
    o->n = off;
    CK(cudaMallocAsync(&o->buf, off * sizeof(float), m->ctx->stream));       // stream-ordered: no device-wide stall per traindbm! call
    CK(cudaMemsetAsync(o->buf, 0, off * sizeof(float), m->ctx->stream));
    *out = o;
    API_END
}
int32_t scdbm_opt_destroy(scdbm_opt* o) {
    API_BEGIN
    if (o) {
        // (every exchange on the communication stream has been joined into the context's stream by the step that queued it)
        cudaFreeAsync(o->buf, o->model->ctx->stream);
        scdbm_model* om_ = o->model;
        delete o;
        model_release(om_);
    }
    API_END
}

// dW rows [v0, v0 + vn) of layer l:  pos_scale * v'h - neg_scale * vm'hm  (vn < 0: all rows)
// fuse_bias: also produce da, db in the same launch when the batch is small (returns whether it did)
static bool gradient_weights(scdbm_opt* o, int l, const MatView& v_, const MatView& vm_, const MatView& h, const MatView& hm,
                             float pos_scale, float neg_scale, int64_t v0 = 0, int64_t vn = -1, bool fuse_bias = false) {
    scdbm_model* m = o->model;
    scdbm_ctx* c = m->ctx;
    Layer& L = m->L[l];
    REQUIRE(v_.cols == L.V && vm_.cols == L.V && h.cols == L.H && hm.cols == L.H, "gradient: widths do not match the RBM");
    REQUIRE(v_.rows == h.rows && vm_.rows == hm.rows, "gradient: row counts of v/h (or vmodel/hmodel) differ");
    if (vn < 0) vn = L.V;
    auto cols_view = [&](const MatView& x) {
        MatView r = x;
        r.p0 += v0;
        if (r.p1) r.p1 += v0;
        if (r.f) r.f += v0;
        r.cols = vn;
        return r;
    };
    const MatView v = cols_view(v_), vm = cols_view(vm_);
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.nsrc = 2;
    g.M = vn;
    g.N = L.H;
    auto opT = [](const MatView& x) { return x.f ? Operand{nullptr, nullptr, x.f, 1, x.ld, x.kind} : Operand{x.p0, x.p1, nullptr, 1, x.ld, x.kind}; };
    g.src[0].a = opT(v);
    g.src[0].b = opT(h);
    g.src[0].K = v.rows;
    g.src[0].scale = pos_scale;
    g.src[1].a = opT(vm);
    g.src[1].b = opT(hm);
    g.src[1].K = vm.rows;
    g.src[1].scale = -neg_scale;
    g.epi = epi_base(c);
    g.epi.mode = EPI_RAW_F32;
    g.epi.out_f32 = o->buf + o->offW[l] + v0 * L.H;
    g.epi.ldo = L.H;
    bool done = false, fused = false;
    const double work = (double)vn * (double)L.H * (double)(v.rows + vm.rows);
    if (c->engine == SCDBM_ENGINE_TC || (c->engine == SCDBM_ENGINE_AUTO && work >= (double)(1 << 20))) {
        TcGradArgs ga{v, h, vm, hm, pos_scale, -neg_scale, g.epi.out_f32, vn, L.H, nullptr, 1};
        if (tc_grad_supported(ga)) {
            // bias sums inside the gradient launch for minibatches (<= 1 024 rows: ~50 loads per thread); larger batches go
            // through the two-pass column sums (5 000 rows inside the launch: 157 us instead of 95; 13 500 x 21 000: 2.4 ms)
            if (fuse_bias && v0 == 0 && vn == L.V && std::max(v.rows, vm.rows) <= 1024) {
                ga.da = o->buf + o->offA[l];
                ga.db = o->buf + o->offB[l];
                fused = true;
            }
            ga.splits = tc_grad_splits(ga, c->num_sms);
            if (ga.splits > 1) {
                const int64_t need = (int64_t)ga.splits * vn * L.H;
                if (c->scratch_cap < need) {
                    CK(cudaStreamSynchronize(c->stream));
                    cudaFree(c->scratch);
                    c->scratch = nullptr;
                    c->scratch_cap = 0;
                    CK(cudaMalloc(&c->scratch, need * sizeof(float)));
                    c->scratch_cap = need;
                }
                ga.partial = c->scratch;
            }
            CK(launch_grad_tc(c->tc, ga, c->num_sms, c->stream));
            c->tc_launches++;
            done = true;
        }
    }
    if (!done) CK(launch_gemm_simt(g, c->stream));
    c->launches++;
    return done && fused;
}

// da, db of layer l from column sums (deterministic two-pass reduction)
static void gradient_biases(scdbm_opt* o, int l, const MatView& v, const MatView& vm, const MatView& h, const MatView& hm,
                            float pos_scale, float neg_scale) {
    scdbm_model* m = o->model;
    scdbm_ctx* c = m->ctx;
    Layer& L = m->L[l];
    float* ga = o->buf + o->offA[l];
    float* gb = o->buf + o->offB[l];
    if (std::max(v.rows, vm.rows) <= 2048) {
        CK(launch_pdl(k_bias_grad, dim3((unsigned)((L.V + L.H + 31) / 32)), dim3(32, 32), 0, c->stream, v, vm, h, hm, pos_scale, neg_scale, ga, gb));
        CK(cudaGetLastError());
        c->launches++;
        return;
    }
    CK(cudaMemsetAsync(ga, 0, (size_t)(L.V + L.H) * sizeof(float), c->stream));  // offA and offB are adjacent
    colsum(c, v, pos_scale, ga, true);
    colsum(c, vm, -neg_scale, ga, true);
    colsum(c, h, pos_scale, gb, true);
    colsum(c, hm, -neg_scale, gb, true);
    CK(cudaGetLastError());
}

static void gradient_layer(scdbm_opt* o, int l, const MatView& v, const MatView& vm, const MatView& h, const MatView& hm,
                           float pos_scale, float neg_scale) {
    if (!gradient_weights(o, l, v, vm, h, hm, pos_scale, neg_scale, 0, -1, true)) gradient_biases(o, l, v, vm, h, hm, pos_scale, neg_scale);
}

// ---- data-parallel exchange: allreduce(sum) of gradient ranges on the communication stream, ordered behind the work
// queued so far on the context's stream; comm_join() makes the context's stream wait for all of them
static void comm_allreduce_range(scdbm_ctx* c, float* ptr, int64_t n) {
    if (!c->comm || n <= 0) return;
    CK(cudaEventRecord(c->ev_grad, c->stream));
    CK(cudaStreamWaitEvent(c->comm_stream, c->ev_grad, 0));
    nccl_check(nccl_api().AllReduce(ptr, ptr, (size_t)n, ncclFloat32, ncclSum, c->comm, c->comm_stream), "ncclAllReduce(gradient)");
}
static void comm_join(scdbm_ctx* c) {
    if (!c->comm) return;
    CK(cudaEventRecord(c->ev_comm, c->comm_stream));
    CK(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
}

// StackedOptimizer gradient of all layers; with a communicator the block is summed over ranks while it is being
// produced: the weight gradient of a large layer is computed in row blocks and the allreduce of block i runs on the
// communication stream (NVLink) under the GEMM of block i + 1
static void gradient_dbm(scdbm_opt* o, scdbm_model* m, scdbm_particles* mu, scdbm_particles* p, float ps, float ns) {
    scdbm_ctx* c = m->ctx;
    for (int l = 0; l < (int)m->L.size(); ++l) {
        Layer& L = m->L[l];
        const MatView& v = l == 0 ? mu->alias0->v : mu->cur[l]->v;
        const MatView &vm = p->cur[l]->v, &h = mu->cur[l + 1]->v, &hm = p->cur[l + 1]->v;
        const int64_t nW = (int64_t)L.V * L.H;
        // two row blocks: the allreduce of the first runs under the GEMM of the second; more blocks leave too few cluster
        // tiles per launch (four blocks of the 20 000 x 1 024 layer: 80 tiles on 74 clusters)
        static const int dp_blocks = [] { const char* v = getenv("SCDBM_DP_BLOCKS"); return v ? std::max(1, atoi(v)) : 2; }();
        const int nblk = (c->comm && nW >= (int64_t)(4 << 20)) ? dp_blocks : 1;
        const int64_t step = round_up((L.V + nblk - 1) / nblk, 128);
        bool bias_done = false;
        for (int64_t v0 = 0; v0 < L.V; v0 += step) {
            const int64_t vn = std::min<int64_t>(step, L.V - v0);
            bias_done = gradient_weights(o, l, v, vm, h, hm, ps, ns, v0, vn, nblk == 1);
            comm_allreduce_range(c, o->buf + o->offW[l] + v0 * L.H, vn * L.H);
        }
        if (!bias_done) gradient_biases(o, l, v, vm, h, hm, ps, ns);
        comm_allreduce_range(c, o->buf + o->offA[l], L.V + L.H);
    }
    comm_join(c);
}

int32_t scdbm_compute_gradient(scdbm_opt* o, scdbm_model* m, int32_t l, const scdbm_mat* v, const scdbm_mat* vmodel,
                               const scdbm_mat* h, const scdbm_mat* hmodel, int64_t nrows_used) {
    API_BEGIN
    REQUIRE(o && m && v && vmodel && h && hmodel, "NULL argument");
    REQUIRE(o->model == m, "optimizer belongs to a different model");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    const int64_t npos = v->v.rows;
    const int64_t nneg = nrows_used > 0 ? std::min<int64_t>(nrows_used, vmodel->v.rows) : vmodel->v.rows;
    REQUIRE(npos > 0 && nneg > 0, "gradient needs at least one positive and one negative sample");
    // the reference skips the division when the count is 1 (src/optimizers.jl:272-279)
    gradient_layer(o, l, v->v, rows_view(vmodel->v, nneg), h->v, rows_view(hmodel->v, nneg), 1.0f / (float)npos, 1.0f / (float)nneg);
    API_END
}

int32_t scdbm_compute_gradient_dbm(scdbm_opt* o, scdbm_model* m, scdbm_particles* mu, scdbm_particles* p, double pos_scale,
                                   double neg_scale) {
    API_BEGIN
    REQUIRE(o && m && mu && p, "NULL argument");
    REQUIRE(o->model == m && mu->model == m && p->model == m, "objects belong to different models");
    REQUIRE(mu->alias0, "mu has not been filled by scdbm_meanfield");
    const float ps = pos_scale > 0 ? (float)pos_scale : 1.0f / (float)std::max<int64_t>(mu->n, 1);
    const float ns = neg_scale > 0 ? (float)neg_scale : 1.0f / (float)std::max<int64_t>(p->n, 1);
    for (int l = 0; l < (int)m->L.size(); ++l) {
        const MatView& v = l == 0 ? mu->alias0->v : mu->cur[l]->v;
        gradient_layer(o, l, v, p->cur[l]->v, mu->cur[l + 1]->v, p->cur[l + 1]->v, ps, ns);
    }
    API_END
}

static void update_layer(scdbm_model* m, scdbm_opt* o, int l, float lr) {
    scdbm_ctx* c = m->ctx;
    Layer& L = m->L[l];
    const bool nb = L.kind == SCDBM_VIS_NEGBIN;
    const int64_t nW = (int64_t)L.V * L.H;
    k_update<<<grid1d(nW), 256, 0, c->stream>>>(L.W, o->buf + o->offW[l], nW, lr, 0.0f, nb ? 1 : 0);
    k_update<<<grid1d(L.V), 256, 0, c->stream>>>(L.a, o->buf + o->offA[l], L.V, lr, -1e-10f, nb ? 1 : 0);
    k_update<<<grid1d(L.H), 256, 0, c->stream>>>(L.b, o->buf + o->offB[l], L.H, lr, 0.0f, 0);
    CK(cudaGetLastError());
    c->launches += 3;
    refresh_nbc(c, L);   // (weight planes: refresh_operands(m) once after all layers are updated)
}

// K6: update of ALL layers of the model, one fused kernel per layer (parameters, clamps, fp16 operand planes in both
// orientations, NB sampling constants, the round's weight scale and the next round's max |W|)
static void update_all_layers(scdbm_model* m, scdbm_opt* o, float lr) {
    scdbm_ctx* c = m->ctx;
    for (int l = 0; l < (int)m->L.size(); ++l) {
        Layer& L = m->L[l];
        const bool nb = L.kind == SCDBM_VIS_NEGBIN;
        dim3 grid((L.V + 31) / 32, (L.H + 31) / 32), block(32, 8);
        CK(launch_pdl(k_update_layer, dim3(grid), dim3(block), 0, c->stream, L.W, L.a, L.b, L.r, o->buf + o->offW[l], o->buf + o->offA[l], o->buf + o->offB[l], L.V, L.H,
                                                     lr, nb ? 1 : 0, m->d_scale, L.w0, L.w1, L.ldh, L.t0, L.t1, L.ldv, m->d_wmax, m->wmax_round, l == 0 ? 1 : 0, c->mu_inv_max,
                                                     L.nbc, L.nbT));
        if (nb) L.nbL_mu = c->mu_inv_max;
    }
    m->wmax_round = (m->wmax_round + 1) % 3;
    CK(cudaGetLastError());
    c->launches += (int64_t)m->L.size();
}

int32_t scdbm_update_parameters(scdbm_model* m, scdbm_opt* o, int32_t l, double lr) {
    API_BEGIN
    REQUIRE(m && o, "NULL argument");
    REQUIRE(o->model == m, "optimizer belongs to a different model");
    REQUIRE(l < (int)m->L.size(), "layer index out of range");
    if (l < 0 || m->L.size() == 1) update_all_layers(m, o, (float)lr);
    else { update_layer(m, o, l, (float)lr); refresh_operands(m); }    // one layer of a stack: the shared scale needs the full refresh
    API_END
}

int32_t scdbm_opt_get_gradient(scdbm_opt* o, int32_t l, double* W, int64_t ldw, double* a, double* b) {
    API_BEGIN
    REQUIRE(o, "optimizer is NULL");
    scdbm_model* m = o->model;
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    Layer& L = m->L[l];
    CK(cudaStreamSynchronize(m->ctx->stream));
    std::vector<float> t((size_t)L.V * L.H + L.V + L.H);
    CK(cudaMemcpy(t.data(), o->buf + o->offW[l], t.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (W) {
        REQUIRE(ldw >= L.V, "ldw smaller than the number of visible units");
        for (int h = 0; h < L.H; ++h)
            for (int v = 0; v < L.V; ++v) W[(size_t)h * ldw + v] = t[(size_t)v * L.H + h];
    }
    const size_t oa = (size_t)L.V * L.H, ob = oa + L.V;
    if (a) for (int i = 0; i < L.V; ++i) a[i] = t[oa + i];
    if (b) for (int i = 0; i < L.H; ++i) b[i] = t[ob + i];
    API_END
}

int32_t scdbm_opt_buffer(scdbm_opt* o, void** ptr, int64_t* nfloats) {
    API_BEGIN
    REQUIRE(o, "optimizer is NULL");
    if (ptr) *ptr = o->buf;
    if (nfloats) *nfloats = o->n;
    API_END
}

// ------------------------------------------------------------- RBM trainer
int32_t scdbm_trainer_create(scdbm_model* m, int32_t l, int32_t B, int32_t pcd, scdbm_trainer** out) {
    API_BEGIN
    REQUIRE(m && out, "NULL argument");
    REQUIRE(l >= 0 && l < (int)m->L.size(), "layer index out of range");
    REQUIRE(B >= 1, "batchsize must be positive");
    scdbm_ctx* c = m->ctx;
    Layer& L = m->L[l];
    auto* t = new scdbm_trainer{m, l, B, pcd, 1};
    model_retain(m);
    const int vkind = L.kind == SCDBM_VIS_NEGBIN ? MAT_COUNT : MAT_REAL;  // upper layers train on potentials
    t->v = mat_new(c, round_up(B, 128), L.V, vkind);
    t->v_groups = 1;
    t->h = mat_new(c, B, L.H, MAT_REAL);
    t->hm = mat_new(c, B, L.H, MAT_REAL);
    t->hs = mat_new(c, B, L.H, MAT_BINARY);
    t->vm = mat_new(c, B, L.V, MAT_REAL);
    t->vs = mat_new(c, B, L.V, L.kind == SCDBM_VIS_NEGBIN ? MAT_COUNT : MAT_BINARY);
    CK(scdbm_opt_create(m, &t->opt) == SCDBM_OK ? cudaSuccess : cudaErrorUnknown);
    if (pcd) {
        // chainstate = rand(batchsize, nhidden): uniforms used as probabilities (quirk B9)
        t->chain = mat_new(c, B, L.H, MAT_REAL);
        k_init<<<grid1d((int64_t)B * ((L.H + 3) / 4)), 256, 0, c->stream>>>(t->chain->v, 2, nullptr, nullptr, (uint32_t)c->seed,
                                                                             (uint32_t)(c->seed >> 32), 1u, 0u);
        CK(cudaGetLastError());
        c->launches++;
    }
    *out = t;
    API_END
}

int32_t scdbm_trainer_destroy(scdbm_trainer* t) {
    API_BEGIN
    if (!t) return SCDBM_OK;
    cudaStreamSynchronize(t->model->ctx->stream);
    mat_free(t->chain); mat_free(t->v); mat_free(t->h); mat_free(t->hm); mat_free(t->hs); mat_free(t->vm); mat_free(t->vs);
    mat_free(t->est);
    scdbm_opt_destroy(t->opt);
    cudaFree(t->d_perm);
    cudaFree(t->d_anyhi);
    scdbm_model* tm = t->model;
    delete t;
    model_release(tm);
    API_END
}

int32_t scdbm_trainer_clock(scdbm_trainer* t, int64_t* clock, int64_t set_to) {
    API_BEGIN
    REQUIRE(t, "trainer is NULL");
    if (set_to >= 0) t->clock = set_to;
    if (clock) *clock = t->clock;
    API_END
}

int32_t scdbm_trainer_chainstate(scdbm_trainer* t, scdbm_mat** borrowed) {
    API_BEGIN
    REQUIRE(t && borrowed, "NULL argument");
    REQUIRE(t->chain, "trainer was created without PCD");
    *borrowed = mat_retain(t->chain);
    API_END
}

int32_t scdbm_rbm_train_epoch(scdbm_trainer* t, const scdbm_mat* x, const int32_t* perm, double lr, int32_t cdsteps,
                              double upfactor, double downfactor) {
    API_BEGIN
    REQUIRE(t && x && perm, "NULL argument");
    REQUIRE(cdsteps >= 1, "cdsteps must be >= 1");
    scdbm_model* m = t->model;
    scdbm_ctx* c = m->ctx;
    Layer& L = m->L[t->l];
    const int64_t n = x->v.rows;
    const int B = t->B;
    REQUIRE(x->v.cols == L.V, "data width does not match the visible layer");
    REQUIRE(x->v.kind == t->v->v.kind, "data kind does not match the trainer (COUNT for NB, REAL otherwise)");
    if (B > n) throw Err(SCDBM_EINVAL, "Batchsize (" + std::to_string(B) + ") exceeds number of samples (" + std::to_string(n) + ").");
    for (int64_t i = 0; i < n; ++i) REQUIRE(perm[i] >= 0 && perm[i] < n, "perm entry out of range");
    if (t->perm_cap < n) {
        cudaFree(t->d_perm);
        CK(cudaMalloc(&t->d_perm, n * sizeof(int32_t)));
        t->perm_cap = n;
    }
    CK(cudaMemcpyAsync(t->d_perm, perm, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (x->v.hiflag) {   // once per epoch: minibatches gathered from data without large counts skip the hi count plane
        if (!t->d_anyhi) CK(cudaMalloc(&t->d_anyhi, sizeof(uint32_t)));
        k_any_flag<<<1, 256, 0, c->stream>>>(x->v.hiflag, (n + 127) / 128, t->d_anyhi);
        c->launches++;
    }
    if (t->track_mean && (!t->est || t->est->v.rows != n)) {
        mat_free(t->est);
        t->est = nullptr;
        t->est = mat_new(c, n, L.V, MAT_REAL);
    }
    const float up = (float)upfactor, down = (float)downfactor;
    // minibatches are gathered a run at a time (one launch per <= 64 MB of rows), each on its own 128-row tiles
    const int64_t Bpad = round_up(B, 128);
    const int64_t G = std::max<int64_t>(1, std::min<int64_t>((n + B - 1) / B, ((int64_t)64 << 20) / (Bpad * t->v->v.ld * 8)));
    if (t->v_groups != G) {
        const int vkind = t->v->v.kind;
        mat_free(t->v);
        t->v = nullptr;
        t->v = mat_new(c, G * Bpad, L.V, vkind);
        t->v_groups = G;
    }
    for (int64_t s = 0, mb = 0; s < n; s += B, ++mb) {
        const int64_t thisb = std::min<int64_t>(B, n - s);
        const int64_t mrows = t->pcd ? B : thisb;  // PCD keeps the full chain (rbmtraining.jl:548-553)
        const int64_t grp = mb % G;
        if (grp == 0) {
            const int64_t cnt = std::min<int64_t>(n - s, G * B);
            CK(launch_pdl(k_gather_rows, dim3((unsigned)cnt), dim3(128), 0, c->stream, x->v, t->d_perm + s, t->v->v,
                          x->v.hiflag ? t->d_anyhi : nullptr, B, (int)Bpad));   // sets the hi-plane flags of the gathered tiles
            c->launches++;
        }
        MatView v = rows_at(t->v->v, grp * Bpad, thisb), h = rows_view(t->h->v, thisb);
        DrawCfg d;
        op_hidden(m, t->l, v, h, up, SCDBM_OUT_POTENTIAL, d);                         // :556
        MatView prob = t->pcd ? t->chain->v : h;                                       // :561-565
        MatView hs = rows_view(t->hs->v, mrows);
        DrawAddr da{(uint32_t)c->seed, (uint32_t)(c->seed >> 32), (uint32_t)t->clock, stream_id(1), 0u, nullptr, 0};
        CK(launch_pdl(k_bernoulli, dim3(grid1d(mrows * ((L.H + 3) / 4))), dim3(256), 0, c->stream, rows_view(prob, mrows), hs, da));  // :567
        c->launches++;
        t->clock++;
        MatView vs = rows_view(t->vs->v, mrows);
        for (int k = 0; k < cdsteps - 1; ++k) {                                        // :570, samplers.jl:47-56
            d.tick = (uint32_t)t->clock; d.stream = stream_id(0);
            op_visible(m, t->l, hs, vs, down, SCDBM_OUT_SAMPLE, d);
            t->clock++;
            d.tick = (uint32_t)t->clock; d.stream = stream_id(1);
            op_hidden(m, t->l, vs, hs, up, SCDBM_OUT_SAMPLE, d);
            t->clock++;
        }
        MatView vm = rows_view(t->vm->v, mrows);
        op_visible(m, t->l, hs, vm, down, SCDBM_OUT_POTENTIAL, d);                     // :573
        if (t->track_mean) {                                                           // :575-576 (first n rows are used)
            const int64_t cnt = std::min<int64_t>(mrows, n - s);
            MatView dstv = t->est->v;
            dstv.p0 += s * dstv.ld; dstv.p1 += s * dstv.ld; dstv.f += s * dstv.ld;
            dstv.rows = cnt;
            CK(launch_pdl(k_copy_rows, dim3((unsigned)cnt), dim3(128), 0, c->stream, vm, dstv, cnt));
            c->launches++;
        }
        MatView hm = t->pcd ? t->chain->v : rows_view(t->hm->v, mrows);
        op_hidden(m, t->l, vm, hm, up, SCDBM_OUT_POTENTIAL, d);                        // :578
        gradient_layer(t->opt, t->l, v, rows_view(vm, thisb), h, rows_view(hm, thisb), 1.0f / (float)thisb, 1.0f / (float)thisb);
        if (m->L.size() == 1) update_all_layers(m, t->opt, (float)lr);                 // :586-587
        else { update_layer(m, t->opt, t->l, (float)lr); refresh_operands(m); }
    }
    CK(cudaGetLastError());
    API_END
}

static void dbm_pcd_step_impl(scdbm_model* m, const scdbm_mat* x, scdbm_particles* p, scdbm_particles* mu, scdbm_opt* o, double lr,
                              int32_t nsteps, double eps, int64_t global_npos, int64_t global_nneg, int32_t* mf_iters) {
    REQUIRE(m && x && p && mu && o, "NULL argument");
    REQUIRE(p->model == m && mu->model == m && o->model == m, "objects belong to different models");
    REQUIRE(m->L.size() >= 2, "traindbm! needs at least two RBM layers");
    // dbmtraining.jl:290-291: gibbssample!(particles), then meanfield(x).  The two are independent (particles vs mu), so
    // the sweeps are queued behind the first group of mean-field iterations: same results, and the host's look at the
    // stopping rule no longer leaves the GPU idle
    int it = 0;
    meanfield_impl(m, x, eps, 0, mu, &it, nullptr, [&] { gibbs_dbm_steps(m, p, nsteps, 1.0f); });
    if (mf_iters) *mf_iters = it;
    const int64_t npos = global_npos > 0 ? global_npos : mu->n, nneg = global_nneg > 0 ? global_nneg : p->n;
    gradient_dbm(o, m, mu, p, 1.0f / (float)std::max<int64_t>(npos, 1), 1.0f / (float)std::max<int64_t>(nneg, 1));   // :293
    update_all_layers(m, o, (float)lr);                                            // :294
}

int32_t scdbm_dbm_pcd_step(scdbm_model* m, const scdbm_mat* x, scdbm_particles* p, scdbm_particles* mu, scdbm_opt* o,
                           double lr, int32_t nsteps, double eps, int32_t* mf_iters) {
    API_BEGIN
    REQUIRE(!m || !m->ctx->comm, "the context has a communicator: call scdbm_dbm_pcd_step_dp with the global sample counts");
    dbm_pcd_step_impl(m, x, p, mu, o, lr, nsteps, eps, 0, 0, mf_iters);
    API_END
}

int32_t scdbm_dbm_pcd_step_dp(scdbm_model* m, const scdbm_mat* x, scdbm_particles* p, scdbm_particles* mu, scdbm_opt* o,
                              double lr, int32_t nsteps, double eps, int64_t global_nsamples, int64_t global_nparticles,
                              int32_t* mf_iters) {
    API_BEGIN
    REQUIRE(global_nsamples > 0 && global_nparticles > 0, "global sample / particle counts must be positive");
    dbm_pcd_step_impl(m, x, p, mu, o, lr, nsteps, eps, global_nsamples, global_nparticles, mf_iters);
    API_END
}

// ------------------------------------------------------------ communicator
int32_t scdbm_comm_unique_id(uint8_t* id128) {
    API_BEGIN
    REQUIRE(id128, "NULL argument");
    std::string err;
    if (!nccl_api().load(err)) throw Err(SCDBM_EUNSUP, err);
    ncclUniqueId id;
    nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    API_END
}

int32_t scdbm_comm_init(scdbm_ctx* c, const uint8_t* id128, int32_t rank, int32_t nranks) {
    API_BEGIN
    REQUIRE(c && id128, "NULL argument");
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank must be in [0, nranks)");
    REQUIRE(!c->comm, "the context already has a communicator");
    std::string err;
    if (!nccl_api().load(err)) throw Err(SCDBM_EUNSUP, err);
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    nccl_check(nccl_api().CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
    c->rank = rank;
    c->nranks = nranks;
    CK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_grad, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    if (!c->d_red) CK(cudaMalloc(&c->d_red, 64 * sizeof(double)));
    API_END
}

int32_t scdbm_comm_destroy(scdbm_ctx* c) {
    API_BEGIN
    REQUIRE(c, "ctx is NULL");
    if (c->comm) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaStreamSynchronize(c->comm_stream));
        nccl_check(nccl_api().CommDestroy(c->comm), "ncclCommDestroy");
        c->comm = nullptr;
        cudaStreamDestroy(c->comm_stream);
        cudaEventDestroy(c->ev_grad);
        cudaEventDestroy(c->ev_comm);
        c->comm_stream = nullptr;
        c->rank = 0;
        c->nranks = 1;
    }
    API_END
}

int32_t scdbm_allreduce_gradient(scdbm_opt* o) {
    API_BEGIN
    REQUIRE(o, "optimizer is NULL");
    scdbm_ctx* c = o->model->ctx;
    if (c->comm) nccl_check(nccl_api().AllReduce(o->buf, o->buf, (size_t)o->n, ncclFloat32, ncclSum, c->comm, c->stream), "ncclAllReduce(gradient)");
    API_END
}

int32_t scdbm_comm_allreduce_f64(scdbm_ctx* c, double* values, int32_t n, int32_t op) {
    API_BEGIN
    REQUIRE(c && values, "NULL argument");
    REQUIRE(n >= 0 && n <= 64, "at most 64 values");
    REQUIRE(op == 0 || op == 1, "op must be 0 (sum) or 1 (max)");
    if (!c->comm || n == 0) return SCDBM_OK;
    CK(cudaMemcpyAsync(c->d_red, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    nccl_check(nccl_api().AllReduce(c->d_red, c->d_red, (size_t)n, ncclFloat64, op ? ncclMax : ncclSum, c->comm, c->stream), "ncclAllReduce(f64)");
    CK(cudaMemcpyAsync(values, c->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    API_END
}

int32_t scdbm_trainer_track_mean(scdbm_trainer* t, int32_t enable) {
    API_BEGIN
    REQUIRE(t, "trainer is NULL");
    t->track_mean = enable != 0;
    API_END
}

int32_t scdbm_trainer_estimated_mean(scdbm_trainer* t, scdbm_mat** borrowed) {
    API_BEGIN
    REQUIRE(t && borrowed, "NULL argument");
    REQUIRE(t->est, "no epoch has been trained with scdbm_trainer_track_mean enabled");
    *borrowed = mat_retain(t->est);
    API_END
}

int32_t scdbm_mle_theta(const scdbm_mat* x, const scdbm_mat* mu, double lambda, int32_t maxiter, double convtol, double* theta) {
    API_BEGIN
    REQUIRE(x && mu && theta, "NULL argument");
    REQUIRE(x->v.rows == mu->v.rows && x->v.cols == mu->v.cols, "x and mu must have the same shape");
    REQUIRE(x->v.rows > 0 && x->v.cols > 0 && maxiter >= 0, "empty input");
    scdbm_ctx* c = x->ctx;
    double* d = nullptr;
    CK(cudaMallocAsync(&d, x->v.cols * sizeof(double), c->stream));
    k_mle_theta<<<(unsigned)x->v.cols, 256, 0, c->stream>>>(x->v, mu->v, lambda, maxiter, convtol, d);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(theta, d, x->v.cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    cudaFreeAsync(d, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    CK(e);
    API_END
}

int32_t scdbm_sample_dropout(scdbm_mat* v, const scdbm_mat* x, const double* mid, const double* shape, int32_t nparams,
                             int32_t log1p_form, uint32_t tick, uint64_t row_offset) {
    API_BEGIN
    REQUIRE(v && x && mid && shape, "NULL argument");
    REQUIRE(v->v.rows == x->v.rows && v->v.cols == x->v.cols, "the data matrix must have the shape of the visible states");
    REQUIRE(nparams == 1 || nparams == v->v.cols, "dropout parameters: one value or one per visible unit");
    if (v->v.rows == 0) return SCDBM_OK;
    scdbm_ctx* c = v->ctx;
    std::vector<float> h(2 * nparams);
    for (int i = 0; i < nparams; ++i) { h[i] = (float)mid[i]; h[nparams + i] = (float)shape[i]; }
    float* d = nullptr;
    CK(cudaMallocAsync(&d, h.size() * sizeof(float), c->stream));
    CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    DrawAddr da{(uint32_t)c->seed, (uint32_t)(c->seed >> 32), tick, stream_id(0, P_DROPOUT), (uint32_t)row_offset, nullptr, 0};
    set_hiflags(c, v->v, true);
    k_dropout<<<grid1d(v->v.rows * ((v->v.cols + 3) / 4)), 256, 0, c->stream>>>(v->v, x->v, d, d + nparams, nparams, log1p_form, da);
    c->launches++;
    cudaFreeAsync(d, c->stream);
    CK(cudaStreamSynchronize(c->stream));   // `h` goes out of scope
    API_END
}

int32_t scdbm_gibbs_cond(scdbm_model* m, scdbm_particles* p, const uint8_t* mask, int64_t mask_rows, int64_t ldmask, int32_t nsteps) {
    API_BEGIN
    REQUIRE(m && p && mask, "NULL argument");
    REQUIRE(p->model == m && !p->real_valued, "particles belong to a different model or hold mean-field values");
    REQUIRE(mask_rows == 0 || (mask_rows == p->n && ldmask >= p->n), "mask must be a vector over the visible units or an (nparticles x nvisible) matrix");
    REQUIRE(nsteps >= 0, "nsteps must be non-negative");
    if (p->n == 0 || nsteps == 0) return SCDBM_OK;
    scdbm_ctx* c = m->ctx;
    const int V = m->L[0].V;
    const int nl = m->nlayers();
    // initial visible values (origvisibles = deepcopy(particles[1]))
    scdbm_mat* orig = mat_new(c, p->n, V, p->cur[0]->v.kind);
    int32_t* d_cols = nullptr;
    uint8_t* d_mask = nullptr;
    int ncols = 0;
    try {
        CK(launch_pdl(k_copy_rows, dim3((unsigned)p->n), dim3(128), 0, c->stream, p->cur[0]->v, orig->v, p->n));
        c->launches++;
        if (mask_rows == 0) {
            std::vector<int32_t> cols;
            for (int j = 0; j < V; ++j) if (mask[j]) cols.push_back(j);
            ncols = (int)cols.size();
            if (ncols) {
                CK(cudaMalloc(&d_cols, ncols * sizeof(int32_t)));
                CK(cudaMemcpyAsync(d_cols, cols.data(), ncols * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
                CK(cudaStreamSynchronize(c->stream));
            }
        } else {
            std::vector<uint8_t> rm((size_t)p->n * V);
            for (int j = 0; j < V; ++j)
                for (int64_t i = 0; i < p->n; ++i) rm[(size_t)i * V + j] = mask[(size_t)j * ldmask + i] ? 1 : 0;
            CK(cudaMalloc(&d_mask, rm.size()));
            CK(cudaMemcpyAsync(d_mask, rm.data(), rm.size(), cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        auto restore = [&](const MatView& dst) {
            if (d_mask) k_restore_mask<<<grid1d(p->n * V), 256, 0, c->stream>>>(dst, orig->v, d_mask);
            else if (ncols) k_restore_cols<<<grid1d(p->n * ncols), 256, 0, c->stream>>>(dst, orig->v, d_cols, ncols);
            c->launches++;
        };
        for (int step = 0; step < nsteps; ++step) {
            DrawCfg d;
            d.row_offset = (uint32_t)p->row_offset;
            if (m->L.size() == 1) {                         // src/gibbssampling.jl:1052-1064
                d.tick = (uint32_t)p->clock; d.stream = stream_id(0);
                op_visible(m, 0, p->cur[1]->v, p->cur[0]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
                restore(p->cur[0]->v);
                p->clock++;
                d.tick = (uint32_t)p->clock; d.stream = stream_id(1);
                op_hidden(m, 0, p->cur[0]->v, p->cur[1]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
                p->clock++;
            } else {                                        // :1066-1104
                d.tick = (uint32_t)p->clock;
                d.stream = stream_id(0);
                op_visible(m, 0, p->cur[1]->v, p->nxt[0]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
                restore(p->nxt[0]->v);
                for (int i = 1; i < nl - 1; ++i) {
                    d.stream = stream_id(i);
                    op_dbm_layer(m, i, &p->cur[i - 1]->v, nullptr, 0, p->cur[i + 1]->v, p->nxt[i]->v, SCDBM_OUT_SAMPLE, d, nullptr, nullptr);
                }
                d.stream = stream_id(nl - 1);
                op_hidden(m, nl - 2, p->cur[nl - 2]->v, p->nxt[nl - 1]->v, 1.0f, SCDBM_OUT_SAMPLE, d);
                std::swap(p->cur, p->nxt);
                p->clock++;
            }
        }
        CK(cudaStreamSynchronize(c->stream));
    } catch (...) {
        mat_free(orig); cudaFree(d_cols); cudaFree(d_mask);
        throw;
    }
    mat_free(orig); cudaFree(d_cols); cudaFree(d_mask);
    API_END
}

// --------------------------------------------------------------------- AIS
static void ais_ratio(scdbm_model* m, const SweepSrc* s, int ns, int64_t n, int64_t N, const float* bias0, const float* bias1,
                      float t1, float t2, double* logw) {
    EpiParams e = epi_base(m->ctx);
    e.mode = EPI_AIS_RATIO;
    e.bias0 = bias0;
    e.bias1 = bias1;
    e.t1 = t1;
    e.t2 = t2;
    e.logw = logw;
    run_sweep(m->ctx, s, ns, n, N, e);
}

int32_t scdbm_ais_run(scdbm_model* m, const double* temps, int32_t ntemps, int64_t n, int32_t burnin, uint64_t row_offset,
                      double* out) {
    API_BEGIN
    REQUIRE(m && temps && out, "NULL argument");
    REQUIRE(ntemps >= 2, "need at least two temperatures");
    REQUIRE(n > 0 && burnin >= 0, "nparticles must be positive, burnin non-negative");
    scdbm_ctx* c = m->ctx;
    double* d_logw = nullptr;
    CK(cudaMallocAsync(&d_logw, n * sizeof(double), c->stream));
    CK(cudaMemsetAsync(d_logw, 0, n * sizeof(double), c->stream));
    scdbm_particles* p = nullptr;
    if (scdbm_particles_create(m, n, row_offset, 0, &p) != SCDBM_OK) throw Err(SCDBM_ENOMEM, g_err);
    const int nl = m->nlayers();
    try {
        if (m->L.size() == 1) {
            // aislogimpweights(rbm) (evaluating.jl:18-46): h ~ Bernoulli(sigm(hidbias))
            Layer& L = m->L[0];
            k_init<<<grid1d(n * ((L.H + 3) / 4)), 256, 0, c->stream>>>(p->cur[1]->v, 1, L.b, nullptr, (uint32_t)c->seed,
                                                                        (uint32_t)(c->seed >> 32), 1u, (uint32_t)row_offset);
            c->launches++;
            p->clock = 1;
            for (int k = 1; k < ntemps; ++k) {
                SweepSrc s{&p->cur[1]->v, &L, false};
                ais_ratio(m, &s, 1, n, L.V, L.a, nullptr, (float)temps[k], (float)temps[k - 1], d_logw);
                if (scdbm_gibbs(m, p, burnin, temps[k], 1.0, 1.0) != SCDBM_OK) throw Err(SCDBM_ECUDA, g_err);
            }
        } else {
            if (scdbm_particles_init(p, 1) != SCDBM_OK) throw Err(SCDBM_ECUDA, g_err);
            for (int k = 1; k < ntemps; ++k) {
                const float t1 = (float)temps[k], t2 = (float)temps[k - 1];
                {   // visible layer summed out from h1 (evaluating.jl:195, :1356-1364)
                    SweepSrc s{&p->cur[1]->v, &m->L[0], false};
                    ais_ratio(m, &s, 1, n, m->L[0].V, m->L[0].a, nullptr, t1, t2, d_logw);
                }
                // even layers of the hidden DBM: absolute layers 2, 4, ... (evaluating.jl:265-290)
                for (int j = 2; j < nl; j += 2) {
                    SweepSrc s[2];
                    int ns = 0;
                    s[ns++] = SweepSrc{&p->cur[j - 1]->v, &m->L[j - 1], true};
                    const float* bias1 = nullptr;
                    if (j + 1 < nl) { s[ns++] = SweepSrc{&p->cur[j + 1]->v, &m->L[j], false}; bias1 = m->L[j].a; }
                    ais_ratio(m, s, ns, n, m->units(j), m->L[j - 1].b, bias1, t1, t2, d_logw);
                }
                gibbs_dbm_steps(m, p, burnin, t1);  // copyannealed! folded into the epilogue scale
            }
        }
        CK(cudaMemcpyAsync(out, d_logw, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } catch (...) {
        scdbm_particles_destroy(p);
        cudaFreeAsync(d_logw, c->stream);
        throw;
    }
    scdbm_particles_destroy(p);
    cudaFreeAsync(d_logw, c->stream);
    API_END
}

}  // extern "C"
