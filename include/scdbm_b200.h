/* libscdbm_b200 -- C ABI of the B200-native Gibbs / CD / PCD / mean-field / AIS
 * hot path of scDBM.jl.
 *
 * The reference (pure Julia) has no FFI; its seams are generic functions that
 * mutate caller-owned Matrix{Float64}.  Each entry point below names the
 * reference function (file:line under the reference checkout) it replaces.
 * A Julia `ccall` binding and the Python `ctypes` binding are shown in
 * INTEGRATION.md; the Python mirror lives in scdbm.jl_b200/.
 *
 * Conventions
 *  - every function returns int32 status: 0 ok, <0 error class (below);
 *    scdbm_last_error() returns a thread-local message.
 *  - opaque handles are library-owned and released by *_destroy.  Handles are
 *    reference-counted: a child keeps its parent alive (matrices, models ->
 *    context; particles, optimizers, trainers -> model), and every scdbm_mat*
 *    handed out by scdbm_particles_layer / scdbm_trainer_chainstate /
 *    scdbm_trainer_estimated_mean is a SHARED handle (one more reference on
 *    the same matrix) that the caller releases with scdbm_mat_destroy -- it
 *    stays valid after the particles / trainer object is destroyed.
 *  - host matrices are COLUMN-MAJOR Float64 with explicit leading dimension
 *    (a Julia Matrix{Float64} or a NumPy order='F' array passes its pointer
 *    unchanged); host buffers are never retained after return.
 *  - logical shapes follow the reference: states (nsamples x nunits), weights
 *    (nvisible x nhidden) (src/bmtypes.jl:17-21,85-92, src/gibbssampling.jl:1-9).
 *  - work is asynchronous on the context's stream; downloads and scdbm_sync
 *    synchronise.  One host thread per context.
 *  - randomness is counter-based Philox4x32-10 keyed by (seed; stream = layer |
 *    purpose; clock tick; global row; column), hence independent of launch
 *    geometry and of how chains are sharded over GPUs.
 */
#ifndef SCDBM_B200_H
#define SCDBM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCDBM_OK 0
#define SCDBM_EINVAL (-1)   /* bad argument / shape mismatch (Julia: error(...)/DimensionMismatch) */
#define SCDBM_ECUDA (-2)    /* CUDA runtime failure */
#define SCDBM_ENOMEM (-3)   /* device allocation failed */
#define SCDBM_EUNSUP (-4)   /* shape/engine combination not supported */

/* visible unit family of the first RBM (src/bmtypes.jl:17,85) */
#define SCDBM_VIS_BERNOULLI 0
#define SCDBM_VIS_NEGBIN 1

/* matrix kinds (device storage, see DESIGN.md "data layout") */
#define SCDBM_MAT_BINARY 0
#define SCDBM_MAT_COUNT 1
#define SCDBM_MAT_REAL 2

/* what a conditional operator returns */
#define SCDBM_OUT_INPUT 0      /* hiddeninput! / visibleinput!                */
#define SCDBM_OUT_POTENTIAL 1  /* hiddenpotential! / visiblepotential!        */
#define SCDBM_OUT_SAMPLE 2     /* samplehidden! / samplevisible!              */

/* engines (ctx option "engine") */
#define SCDBM_ENGINE_AUTO 0
#define SCDBM_ENGINE_SIMT 1    /* CUDA-core FFMA kernel (tiny / unaligned shapes) */
#define SCDBM_ENGINE_TC 2      /* tcgen05 + TMA + TMEM kernel                     */

typedef struct scdbm_ctx scdbm_ctx;
typedef struct scdbm_model scdbm_model;
typedef struct scdbm_mat scdbm_mat;
typedef struct scdbm_particles scdbm_particles;
typedef struct scdbm_opt scdbm_opt;
typedef struct scdbm_trainer scdbm_trainer;

const char* scdbm_last_error(void);
int32_t scdbm_version(void);

/* ---- context ------------------------------------------------------------ */
/* stream == NULL: the library creates its own stream.  Otherwise a
 * cudaStream_t owned by the caller (e.g. torch's current stream). */
int32_t scdbm_ctx_create(int32_t device, uint64_t seed, void* stream, scdbm_ctx** out);
int32_t scdbm_ctx_destroy(scdbm_ctx* ctx);
int32_t scdbm_sync(scdbm_ctx* ctx);
/* options: "engine" (0/1/2), "mu_inv_max" (NB inversion / gamma-Poisson switch),
 * "nb_pshift" (0 or 1e-6, quirk B5), "seed", "two_sm" (0/1: cta_group::2 sweep kernel
 * for wide Bernoulli / deterministic sweeps; same results),
 * "pdl" (0/1, process-wide: programmatic dependent launch of the library's kernels; same results) */
int32_t scdbm_ctx_set_option(scdbm_ctx* ctx, const char* name, double value);
int32_t scdbm_ctx_get_option(scdbm_ctx* ctx, const char* name, double* value);
/* CUDA-event timer on the context's stream, and a count of the library's own
 * kernel launches since the last reset (bench.py "gpu_launches"). */
int32_t scdbm_timer_start(scdbm_ctx* ctx);
int32_t scdbm_timer_stop(scdbm_ctx* ctx, float* elapsed_ms);
int32_t scdbm_launch_count(scdbm_ctx* ctx, int64_t* count, int32_t reset);

/* ---- model: RBM (nrbms == 1) or DBM = Vector{<:AbstractRBM} --------------
 * replaces the parameter structs BernoulliRBM / NegativeBinomialBernoulliRBM
 * (src/bmtypes.jl:17-21,85-92) and MultimodalDBM (:139).
 * units[0..nrbms] = layer sizes (visible, h1, ...). */
int32_t scdbm_model_create(scdbm_ctx* ctx, int32_t nrbms, const int32_t* units, int32_t visible_kind,
                           scdbm_model** out);
int32_t scdbm_model_destroy(scdbm_model* m);
/* W: (units[l] x units[l+1]) column-major, ldw >= units[l]; visbias, hidbias;
 * inversedispersion only for an NB layer 0 (else NULL). */
int32_t scdbm_model_set_layer(scdbm_model* m, int32_t l, const double* W, int64_t ldw, const double* visbias,
                              const double* hidbias, const double* inversedispersion);
int32_t scdbm_model_get_layer(scdbm_model* m, int32_t l, double* W, int64_t ldw, double* visbias, double* hidbias,
                              double* inversedispersion);
int32_t scdbm_model_copy_layer(scdbm_model* dst, int32_t dst_l, scdbm_model* src, int32_t src_l);
/* initrbm (src/rbmtraining.jl:332-374): NB: W = -U(0.0005,0.002), a = log(mean/(mean+r)), b = 0;
 * Bernoulli: W = randn/sqrt(V), a = logit(mean).  x supplies the column means. */
int32_t scdbm_model_init_layer(scdbm_model* m, int32_t l, const scdbm_mat* x, const double* inversedispersion,
                               uint64_t seed);

/* ---- state matrices (data, particles' layers, potentials) ---------------- */
int32_t scdbm_mat_create(scdbm_ctx* ctx, int64_t rows, int64_t cols, int32_t kind, scdbm_mat** out);
int32_t scdbm_mat_destroy(scdbm_mat* m);
int32_t scdbm_mat_shape(const scdbm_mat* m, int64_t* rows, int64_t* cols, int32_t* kind);
int32_t scdbm_mat_upload(scdbm_mat* m, const double* x, int64_t ld);        /* column-major fp64 in   */
int32_t scdbm_mat_download(const scdbm_mat* m, double* x, int64_t ld);      /* column-major fp64 out  */
int32_t scdbm_mat_download_u16(const scdbm_mat* m, uint16_t* x_rowmajor);   /* compact counts, row = sample */
/* Asynchronous form for generation pipelines (src/gibbssampling.jl:663-730 called in a loop): a uint16 snapshot of
 * the matrix is taken in stream order and drained to `x_rowmajor` (pinned host memory) by the copy engine on a
 * separate stream, so the next batch of chains runs while the previous one travels.  The host buffer is valid
 * after scdbm_wait_downloads(); one download is in flight per context, a second one queues behind it. */
int32_t scdbm_mat_download_u16_async(const scdbm_mat* m, uint16_t* x_rowmajor);
int32_t scdbm_wait_downloads(scdbm_ctx* ctx);

/* Sparse data path (scRNA-seq counts are 80-92 % zeros).  The reference takes dense
 * Matrix{Float64} data (src/rbmtraining.jl:89, src/dbmtraining.jl:108) and returns dense samples
 * (src/gibbssampling.jl:663-730); these entry points carry the same matrices compressed.
 *  upload_compressed: major = 0 CSR (ptr[rows+1], idx = column), 1 CSC (ptr[cols+1], idx = row: the arrays of a
 *    Julia SparseMatrixCSC after the 1 -> 0 index shift).  Entries are validated like scdbm_mat_upload.
 *  csr_prepare + download_csr: row-compressed download of a count / binary matrix; prepare returns nnz so the
 *    caller can allocate indices[nnz] (uint16 or int32, index_bytes = 2 | 4) and values[nnz].
 *  generate_counts: synthetic "10x-like" counts on the device, x_cg ~ NB(r_g, r_g/(s_c m_g + r_g)),
 *    s_c = exp(sigma z_c); Philox-addressed by the global cell index row_offset + c (oracle/synth.py twin). */
int32_t scdbm_mat_upload_compressed(scdbm_mat* m, int32_t major, const int64_t* ptr, const int32_t* idx, const float* val);
int32_t scdbm_mat_csr_prepare(scdbm_mat* m, int64_t* nnz);
int32_t scdbm_mat_download_csr(scdbm_mat* m, int64_t* indptr, void* indices, int32_t index_bytes, uint16_t* values);
int32_t scdbm_mat_generate_counts(scdbm_mat* m, const double* gene_mean, const double* r, double sizefactor_sigma,
                                  uint64_t seed, uint64_t row_offset);
int32_t scdbm_mat_copy(scdbm_mat* dst, const scdbm_mat* src);
/* per-column mean, variance (n-1 denominator, as Julia `var`) and zero fraction
 * (src/NegativeBinomialRBMPlots.jl:148-182,428-462,500-515) */
int32_t scdbm_mat_column_stats(const scdbm_mat* m, double* mean, double* var, double* zerofrac);

/* ---- single conditional operators (parity gates G1/G2) ------------------
 * `what` = SCDBM_OUT_*.  `uniforms`: optional host column-major fp64 matrix
 * with the shape of `out` that replaces the Philox uniforms of the primary
 * draw (gate "same uniforms"); `tick` is the Philox clock tick otherwise.
 *   hidden : hiddeninput!/hiddenpotential!/samplehidden!
 *            (src/gibbssampling.jl:207-212,271-276,327-335,594-610)
 *   visible: visibleinput!/visiblepotential!/samplevisible!
 *            (src/gibbssampling.jl:752-756,765-770,810-820,884-890,949-957,1013-1020)
 *   dbm_layer: intermediate DBM layer from both neighbours
 *            (src/gibbssampling.jl:165-170; src/dbmtraining.jl:197-207) */
int32_t scdbm_hidden(scdbm_model* m, int32_t l, const scdbm_mat* v, scdbm_mat* h_out, double factor, int32_t what,
                     uint32_t tick, const double* uniforms, int64_t ldu);
int32_t scdbm_visible(scdbm_model* m, int32_t l, const scdbm_mat* h, scdbm_mat* v_out, double factor, int32_t what,
                      uint32_t tick, const double* uniforms, int64_t ldu);
int32_t scdbm_dbm_layer(scdbm_model* m, int32_t layer, const scdbm_mat* below, const scdbm_mat* above, scdbm_mat* out,
                        int32_t what, uint32_t tick, const double* uniforms, int64_t ldu);

/* ---- particles and chains ------------------------------------------------
 * Particles (src/gibbssampling.jl:9); initparticles (:386-419, NB visible init
 * :512-531, hidden :421-453); row_offset = global index of particle 0 (chain
 * sharding over GPUs; results do not depend on the sharding). */
/* real_valued != 0: hidden layers hold probabilities (mean-field `mu`), not samples. */
int32_t scdbm_particles_create(scdbm_model* m, int64_t nparticles, uint64_t row_offset, int32_t real_valued,
                               scdbm_particles** out);
int32_t scdbm_particles_destroy(scdbm_particles* p);
int32_t scdbm_particles_init(scdbm_particles* p, int32_t biased);
/* shared handle of the matrix that holds layer `layer` NOW (release with scdbm_mat_destroy); a later
 * scdbm_gibbs may move the layer to the particles' second buffer -- fetch the handle again after sampling. */
int32_t scdbm_particles_layer(scdbm_particles* p, int32_t layer, scdbm_mat** shared);
int32_t scdbm_particles_clock(scdbm_particles* p, int64_t* clock, int64_t set_to /* <0: keep */);
/* gibbssample!(particles, dbm, nsteps) (src/gibbssampling.jl:153-182): synchronous
 * update of all layers; temperature scales all weight matrices (copyannealed!,
 * src/evaluating.jl:320-326,348-355).  For a 1-RBM model this is
 * gibbssample!(particles, rbm, nsteps, up, down) (src/gibbssampling.jl:62-70). */
int32_t scdbm_gibbs(scdbm_model* m, scdbm_particles* p, int32_t nsteps, double temperature, double upfactor,
                    double downfactor);
/* gibbssamplenegativebinomial!(x, particles, rbm, ...) (src/gibbssampling.jl:82-110):
 * p - 1e-6 form, hidden sampled from the OLD visible state. */
int32_t scdbm_gibbs_negbin_rbm(scdbm_model* m, scdbm_particles* p, int32_t nsteps, double upfactor, double downfactor);
/* last step of samples(...; samplelast=false): visible potential into `out`
 * (src/gibbssampling.jl:677-689) */
int32_t scdbm_particles_visible_potential(scdbm_model* m, scdbm_particles* p, scdbm_mat* out);

/* gibbssamplecond! / gibbssamplenegativebinomial_cond! / gibbssamplecondrow! (src/gibbssampling.jl:1042-1237):
 * Gibbs sampling with the masked visible entries reset to their initial values after every visible update.
 * mask_rows == 0: `mask` is a byte vector over the visible units; otherwise a column-major
 * (nparticles x nvisible) byte matrix with leading dimension ldmask. */
int32_t scdbm_gibbs_cond(scdbm_model* m, scdbm_particles* p, const uint8_t* mask, int64_t mask_rows, int64_t ldmask,
                         int32_t nsteps);
/* sampledropout (src/rbmtraining.jl:289-311), applied as in gibbssamplenegativebinomial!(...; zeroinflation=true)
 * (src/gibbssampling.jl:103-107,144-148): v .*= 1 - bernoulli(logistic(shape * (x' - mid))), x' = x (RBM form,
 * nparams == 1) or log1p(x) (DBM form, per-gene parameters, nparams == nvisible). */
int32_t scdbm_sample_dropout(scdbm_mat* v, const scdbm_mat* x, const double* mid, const double* shape, int32_t nparams,
                             int32_t log1p_form, uint32_t tick, uint64_t row_offset);

/* ---- mean-field (src/dbmtraining.jl:175-228) ------------------------------
 * mu: particles with nparticles == rows(x); layer 0 is set to x.  Runs until
 * the global max-abs change <= eps or maxiter iterations (maxiter <= 0: no cap).
 * x.W1 is hoisted out of the loop.  The stopping rule is evaluated on the device; iterations are
 * launched four at a time and those after convergence are no-ops, so the host synchronises once per group. */
int32_t scdbm_meanfield(scdbm_model* m, const scdbm_mat* x, double eps, int32_t maxiter, scdbm_particles* mu,
                        int32_t* niter, double* delta);

/* ---- learning --------------------------------------------------------------
 * scdbm_opt = LoglikelihoodOptimizer / StackedOptimizer gradient holder
 * (src/optimizers.jl:34-39,57-70,103). */
int32_t scdbm_opt_create(scdbm_model* m, scdbm_opt** out);
int32_t scdbm_opt_destroy(scdbm_opt* o);
/* computegradient!(opt, v, vmodel, h, hmodel, rbm) (src/optimizers.jl:207-214,263-310):
 * dW = v'h / rows(v) - vmodel'hmodel / nneg, da, db likewise from column means.
 * nrows_used > 0: only the first nrows_used rows of vmodel/hmodel enter (short last
 * minibatch of PCD, src/rbmtraining.jl:580-585); <= 0: all rows. */
int32_t scdbm_compute_gradient(scdbm_opt* o, scdbm_model* m, int32_t l, const scdbm_mat* v, const scdbm_mat* vmodel,
                               const scdbm_mat* h, const scdbm_mat* hmodel, int64_t nrows_used);
/* computegradient!(StackedOptimizer, mu, particles, dbm) (src/optimizers.jl:431-442).
 * pos_scale / neg_scale > 0 replace 1/rows(mu), 1/nparticles: a data-parallel rank passes the
 * GLOBAL 1/n+, 1/n- so that the allreduce(sum) of the per-rank blocks is the full-batch gradient. */
int32_t scdbm_compute_gradient_dbm(scdbm_opt* o, scdbm_model* m, scdbm_particles* mu, scdbm_particles* particles,
                                   double pos_scale, double neg_scale);
/* updateparameters!(bm, opt) (src/optimizers.jl:450-454,478-485,505-520); l < 0: all layers */
int32_t scdbm_update_parameters(scdbm_model* m, scdbm_opt* o, int32_t l, double learningrate);
int32_t scdbm_opt_get_gradient(scdbm_opt* o, int32_t l, double* W, int64_t ldw, double* visbias, double* hidbias);
/* the gradient block {dW_l, da_l, db_l for all l} as one contiguous device
 * fp32 buffer -- the payload of the data-parallel allreduce */
int32_t scdbm_opt_buffer(scdbm_opt* o, void** device_ptr, int64_t* nfloats);

/* trainrbm! state held by fitrbm across epochs (src/rbmtraining.jl:137-151): the
 * PCD chain state (`rand(batchsize, nhidden)`) and the draw clock. */
int32_t scdbm_trainer_create(scdbm_model* m, int32_t l, int32_t batchsize, int32_t pcd, scdbm_trainer** out);
int32_t scdbm_trainer_destroy(scdbm_trainer* t);
int32_t scdbm_trainer_clock(scdbm_trainer* t, int64_t* clock, int64_t set_to);
int32_t scdbm_trainer_chainstate(scdbm_trainer* t, scdbm_mat** shared);   /* release with scdbm_mat_destroy */
/* `estimatedmean` of trainrbm! (src/rbmtraining.jl:536,575-576): when enabled, each epoch records the model
 * means of its minibatches, in minibatch order, as a (nsamples x nvisible) matrix. */
int32_t scdbm_trainer_track_mean(scdbm_trainer* t, int32_t enable);
int32_t scdbm_trainer_estimated_mean(scdbm_trainer* t, scdbm_mat** shared);   /* release with scdbm_mat_destroy */
/* mle_for_θ (src/rbmtraining.jl:675-719), all genes at once: penalised Fisher scoring of the NB inverse
 * dispersion of gene g from (x[:,g], mu[:,g]); theta[nvisible] out. */
int32_t scdbm_mle_theta(const scdbm_mat* x, const scdbm_mat* mu, double lambda, int32_t maxiter, double convtol,
                        double* theta);
/* trainrbm!(rbm, x; ...) one epoch of minibatch CD-k / PCD
 * (src/rbmtraining.jl:498-591).  perm: the epoch's sample order (randombatchmasks,
 * :416-429), 0-based, length rows(x). */
int32_t scdbm_rbm_train_epoch(scdbm_trainer* t, const scdbm_mat* x, const int32_t* perm, double learningrate,
                              int32_t cdsteps, double upfactor, double downfactor);
/* traindbm!(dbm, x, particles, opt) one PCD step (src/dbmtraining.jl:287-297):
 * gibbssample!(particles, dbm, nsteps); meanfield; gradient; update. */
int32_t scdbm_dbm_pcd_step(scdbm_model* m, const scdbm_mat* x, scdbm_particles* particles, scdbm_particles* mu,
                           scdbm_opt* o, double learningrate, int32_t nsteps, double eps, int32_t* mf_iters);

/* ---- data-parallel exchange (NCCL over NVLink, libnccl.so.2 resolved at run time) ------------------
 * One process per GPU.  Rank 0 obtains a 128-byte id (scdbm_comm_unique_id), hands it to the other ranks by any
 * out-of-band means, and every rank calls scdbm_comm_init on its context.  With a communicator
 *  - scdbm_dbm_pcd_step_dp is traindbm!(dbm, x, particles, opt) (src/dbmtraining.jl:287-297) over row shards of x
 *    and shards of the particles: every rank forms its partial sums scaled by the GLOBAL 1/n+ and 1/n-
 *    (src/optimizers.jl:272-283), the gradient block {dW_l, da_l, db_l} is summed by allreduce on a second stream
 *    while later blocks are still being computed, and every rank applies the same update;
 *  - scdbm_meanfield / the step use the reference's GLOBAL stopping rule: one scalar allreduce(max) of the
 *    iteration's max |change| (src/dbmtraining.jl:194-225), so all ranks run the same number of iterations and the
 *    result does not depend on the sharding (every rank must call it, an empty shard included);
 *  - scdbm_allreduce_gradient sums the optimizer's whole block in place on the context's stream (for callers that
 *    drive computegradient! / updateparameters! themselves);
 *  - scdbm_comm_allreduce_f64 reduces <= 64 host scalars (op 0 sum, 1 max): logmeanexp over particle shards
 *    (src/evaluating.jl:935-943,979-984).
 * Generation and AIS need no collective: chains / particles are addressed by their global index (row_offset). */
int32_t scdbm_comm_unique_id(uint8_t* id128);
int32_t scdbm_comm_init(scdbm_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t nranks);
int32_t scdbm_comm_destroy(scdbm_ctx* ctx);
int32_t scdbm_allreduce_gradient(scdbm_opt* o);
int32_t scdbm_comm_allreduce_f64(scdbm_ctx* ctx, double* values, int32_t n, int32_t op);
int32_t scdbm_dbm_pcd_step_dp(scdbm_model* m, const scdbm_mat* x, scdbm_particles* particles, scdbm_particles* mu,
                              scdbm_opt* o, double learningrate, int32_t nsteps, double eps, int64_t global_nsamples,
                              int64_t global_nparticles, int32_t* mf_iters);

/* ---- AIS (src/evaluating.jl:18-46,164-206,265-355,1314-1364,1596-1648) -----
 * Returns nparticles log importance weights.  temperatures[0..ntemps-1]
 * ascending from 0 to 1. */
int32_t scdbm_ais_run(scdbm_model* m, const double* temperatures, int32_t ntemps, int64_t nparticles, int32_t burnin,
                      uint64_t row_offset, double* logimpweights);

#ifdef __cplusplus
}
#endif
#endif /* SCDBM_B200_H */
