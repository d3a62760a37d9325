# scDBM.jl facade over libscdbm_b200.so -- the reference-side binding.
#
# Mirrors the hot-path names of the reference's export list (src/scDBM.jl:17-64) and keeps its
# signatures; the bodies `ccall` the C ABI in include/scdbm_b200.h.  Julia is not installed in the
# build image, so this file has NOT been executed there; it is a mechanical mirror of the Python
# facade scdbm.jl_b200/api.py, which drives the same entry points in tests and benches.
module scDBM
using SparseArrays

export BernoulliRBM, NegativeBinomialBernoulliRBM, MultimodalDBM, Particles, TrainLayer,
       fitrbm, stackrbms, fitdbm, traindbm!, trainrbm!, gibbssample!, gibbssamplenegativebinomial!,
       samples, sampleparticles, initparticles, meanfield,
       hiddeninput, hiddenpotential, samplehidden, visibleinput, visiblepotential, samplevisible,
       aislogimpweights, logpartitionfunction, loglikelihoodoptimizer, computegradient!, updateparameters!

const LIB = joinpath(@__DIR__, "..", "scdbm.jl_b200", "libscdbm_b200.so")
const Particles = Array{Array{Float64,2},1}

abstract type AbstractRBM end
struct BernoulliRBM <: AbstractRBM                      # src/bmtypes.jl:17-21
   weights::Array{Float64,2}; visbias::Array{Float64,1}; hidbias::Array{Float64,1}
end
struct NegativeBinomialBernoulliRBM <: AbstractRBM      # src/bmtypes.jl:85-92
   weights::Array{Float64,2}; visbias::Array{Float64,1}; hidbias::Array{Float64,1}
   inversedispersion::AbstractArray; dropoutmid::Array{Float64,1}; dropoutshape::Array{Float64,1}
end
const MultimodalDBM = Vector{<:AbstractRBM}
const AbstractBM = Union{MultimodalDBM, AbstractRBM}

# ---------------------------------------------------------------- plumbing
const VIS_BERNOULLI, VIS_NEGBIN = Int32(0), Int32(1)
const MAT_BINARY, MAT_COUNT, MAT_REAL = Int32(0), Int32(1), Int32(2)
const OUT_INPUT, OUT_POTENTIAL, OUT_SAMPLE = Int32(0), Int32(1), Int32(2)

function check(status::Int32)
   status == 0 && return
   msg = unsafe_string(ccall((:scdbm_last_error, LIB), Cstring, ()))
   error("libscdbm_b200 error $status: $msg")
end

mutable struct Context
   h::Ptr{Cvoid}
   function Context(device::Integer = 0, seed::Integer = 0)
      r = Ref{Ptr{Cvoid}}(C_NULL)
      check(ccall((:scdbm_ctx_create, LIB), Int32, (Int32, UInt64, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, seed, C_NULL, r))
      c = new(r[]); finalizer(x -> ccall((:scdbm_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), c); c
   end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
defaultcontext() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context()); DEFAULT_CTX[])
seed!(s::Integer) = check(ccall((:scdbm_ctx_set_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Float64), defaultcontext().h, "seed", s))

rbms(bm::AbstractRBM) = AbstractRBM[bm]
rbms(bm::MultimodalDBM) = bm
viskind(rbm) = rbm isa NegativeBinomialBernoulliRBM ? VIS_NEGBIN : VIS_BERNOULLI

mutable struct DeviceModel
   h::Ptr{Cvoid}; units::Vector{Int32}; kind::Int32
end
function DeviceModel(bm::AbstractBM, ctx::Context = defaultcontext())
   rs = rbms(bm)
   units = Int32[size(rs[1].weights, 1); [size(r.weights, 2) for r in rs]]
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_model_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
               ctx.h, length(rs), units, viskind(rs[1]), ref))
   m = DeviceModel(ref[], units, viskind(rs[1]))
   finalizer(x -> ccall((:scdbm_model_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), m)
   for (l, r) in enumerate(rs)
      disp = r isa NegativeBinomialBernoulliRBM ? convert(Vector{Float64}, r.inversedispersion) : C_NULL
      GC.@preserve r disp check(ccall((:scdbm_model_set_layer, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            m.h, l - 1, r.weights, size(r.weights, 1), r.visbias, r.hidbias, disp))
   end
   m
end
function download!(bm::AbstractBM, m::DeviceModel)          # parameters back into the Julia structs
   for (l, r) in enumerate(rbms(bm))
      disp = r isa NegativeBinomialBernoulliRBM ? r.inversedispersion : C_NULL
      check(ccall((:scdbm_model_get_layer, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            m.h, l - 1, r.weights, size(r.weights, 1), r.visbias, r.hidbias, disp))
   end
   bm
end

mutable struct Mat
   h::Ptr{Cvoid}; rows::Int; cols::Int
end
function Mat(ctx::Context, rows::Integer, cols::Integer, kind::Int32)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_mat_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ref{Ptr{Cvoid}}), ctx.h, rows, cols, kind, ref))
   m = Mat(ref[], rows, cols); finalizer(x -> ccall((:scdbm_mat_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), m); m
end
function upload(ctx::Context, x::Matrix{Float64}, kind::Int32)
   m = Mat(ctx, size(x, 1), size(x, 2), kind)
   GC.@preserve x check(ccall((:scdbm_mat_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), m.h, x, max(size(x, 1), 1)))
   m
end
# sparse data (cells x genes SparseMatrixCSC): only the non-zeros cross PCIe
function upload(ctx::Context, x::SparseMatrixCSC{Float64,<:Integer}, kind::Int32)
   m = Mat(ctx, size(x, 1), size(x, 2), kind)
   ptr = Int64.(x.colptr .- 1); idx = Int32.(x.rowval .- 1); val = Float32.(x.nzval)
   GC.@preserve ptr idx val check(ccall((:scdbm_mat_upload_compressed, LIB), Int32,
      (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Float32}), m.h, Int32(1), ptr, idx, val))
   m
end
# generated cells as a genes x cells SparseMatrixCSC{UInt16} (= the row-compressed cells x genes matrix)
function downloadsparse(m::Mat)
   nnz = Ref{Int64}(0)
   check(ccall((:scdbm_mat_csr_prepare, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), m.h, nnz))
   ptr = Vector{Int64}(undef, m.rows + 1); idx = Vector{Int32}(undef, nnz[]); val = Vector{UInt16}(undef, nnz[])
   GC.@preserve ptr idx val check(ccall((:scdbm_mat_download_csr, LIB), Int32,
      (Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Int32, Ptr{UInt16}), m.h, ptr, idx, Int32(4), val))
   SparseMatrixCSC(m.cols, m.rows, ptr .+ 1, idx .+ Int32(1), val)
end
function download!(x::Matrix{Float64}, m::Mat)
   GC.@preserve x check(ccall((:scdbm_mat_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), m.h, x, max(size(x, 1), 1)))
   x
end
datakind(m::DeviceModel) = m.kind == VIS_NEGBIN ? MAT_COUNT : MAT_BINARY

# ------------------------------------------------- conditional operators
function conditional!(out::Matrix{Float64}, sym::Symbol, rbm::AbstractRBM, in::Matrix{Float64}, what::Int32,
      factor::Float64, inkind::Int32, outkind::Int32)
   ctx = defaultcontext(); m = DeviceModel(rbm, ctx)
   src = upload(ctx, in, inkind); dst = Mat(ctx, size(out, 1), size(out, 2), outkind)
   f = sym == :hidden ? (:scdbm_hidden, LIB) : (:scdbm_visible, LIB)
   if sym == :hidden
      check(ccall((:scdbm_hidden, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, UInt32, Ptr{Float64}, Int64),
                  m.h, 0, src.h, dst.h, factor, what, 0, C_NULL, 0))
   else
      check(ccall((:scdbm_visible, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, UInt32, Ptr{Float64}, Int64),
                  m.h, 0, src.h, dst.h, factor, what, 0, C_NULL, 0))
   end
   download!(out, dst)
end
# src/gibbssampling.jl:197-212,264-276,322-335,589-598
hiddeninput!(hh, rbm, vv) = conditional!(hh, :hidden, rbm, vv, OUT_INPUT, 1.0, rbm isa NegativeBinomialBernoulliRBM ? MAT_COUNT : MAT_REAL, MAT_REAL)
hiddenpotential!(hh, rbm, vv, factor::Float64 = 1.0) = conditional!(hh, :hidden, rbm, vv, OUT_POTENTIAL, factor, rbm isa NegativeBinomialBernoulliRBM ? MAT_COUNT : MAT_REAL, MAT_REAL)
samplehidden!(hh, rbm, vv, factor::Float64 = 1.0) = conditional!(hh, :hidden, rbm, vv, OUT_SAMPLE, factor, rbm isa NegativeBinomialBernoulliRBM ? MAT_COUNT : MAT_REAL, MAT_BINARY)
# src/gibbssampling.jl:747-756,872-890,945-957,1013-1020
visibleinput!(vv, rbm, hh) = conditional!(vv, :visible, rbm, hh, OUT_INPUT, 1.0, MAT_REAL, MAT_REAL)
visiblepotential!(vv, rbm, hh, factor::Float64 = 1.0) = conditional!(vv, :visible, rbm, hh, OUT_POTENTIAL, factor, MAT_REAL, MAT_REAL)
samplevisible!(vv, rbm, hh, factor::Float64 = 1.0) = conditional!(vv, :visible, rbm, hh, OUT_SAMPLE, factor, MAT_REAL, rbm isa NegativeBinomialBernoulliRBM ? MAT_COUNT : MAT_BINARY)
for f in (:hiddeninput, :hiddenpotential, :samplehidden)
   @eval $f(rbm::AbstractRBM, v::Matrix{Float64}, args...) = $(Symbol(f, :!))(Matrix{Float64}(undef, size(v, 1), size(rbm.weights, 2)), rbm, v, args...)
end
for f in (:visibleinput, :visiblepotential, :samplevisible)
   @eval $f(rbm::AbstractRBM, h::Matrix{Float64}, args...) = $(Symbol(f, :!))(Matrix{Float64}(undef, size(h, 1), size(rbm.weights, 1)), rbm, h, args...)
end

# ------------------------------------------------------- particles, chains
mutable struct DeviceParticles
   h::Ptr{Cvoid}; model::DeviceModel; n::Int
end
function DeviceParticles(m::DeviceModel, n::Integer; rowoffset::Integer = 0, realvalued::Bool = false)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_particles_create, LIB), Int32, (Ptr{Cvoid}, Int64, UInt64, Int32, Ref{Ptr{Cvoid}}), m.h, n, rowoffset, realvalued, ref))
   p = DeviceParticles(ref[], m, n); finalizer(x -> ccall((:scdbm_particles_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), p); p
end
function layer(p::DeviceParticles, i::Integer)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_particles_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), p.h, i - 1, ref))
   Mat(ref[], p.n, Int(p.model.units[i]))              # borrowed handle: never destroyed by Julia
end
function tohost(p::DeviceParticles)
   Particles([download!(Matrix{Float64}(undef, p.n, Int(u)), layer(p, i)) for (i, u) in enumerate(p.model.units)])
end

"src/gibbssampling.jl:386-419"
function initparticles(bm::AbstractBM, nparticles::Int; biased::Bool = false)
   p = DeviceParticles(DeviceModel(bm), nparticles)
   check(ccall((:scdbm_particles_init, LIB), Int32, (Ptr{Cvoid}, Int32), p.h, biased))
   tohost(p)
end

"src/gibbssampling.jl:62-70 (RBM), :153-182 (DBM)"
function gibbssample!(particles::Particles, bm::AbstractBM, nsteps::Int = 5, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0)
   m = DeviceModel(bm); p = DeviceParticles(m, size(particles[1], 1))
   for (i, x) in enumerate(particles)
      GC.@preserve x check(ccall((:scdbm_mat_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), layer(p, i).h, x, max(size(x, 1), 1)))
   end
   check(ccall((:scdbm_gibbs, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64, Float64), m.h, p.h, nsteps, 1.0, upfactor, downfactor))
   for (i, x) in enumerate(particles); download!(x, layer(p, i)); end
   particles
end

"src/gibbssampling.jl:82-110 (RBM form); the DBM form (:112-151) is gibbssample!"
function gibbssamplenegativebinomial!(x::Array{Float64,2}, particles::Particles, rbm::NegativeBinomialBernoulliRBM,
      nsteps::Int = 5, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0; zeroinflation::Bool = false)
   zeroinflation && error("zeroinflation is outside the accelerated hot path (DESIGN.md section 7)")
   m = DeviceModel(rbm); p = DeviceParticles(m, size(particles[1], 1))
   for (i, xx) in enumerate(particles)
      GC.@preserve xx check(ccall((:scdbm_mat_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), layer(p, i).h, xx, max(size(xx, 1), 1)))
   end
   check(ccall((:scdbm_gibbs_negbin_rbm, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64), m.h, p.h, nsteps, upfactor, downfactor))
   for (i, xx) in enumerate(particles); download!(xx, layer(p, i)); end
   particles
end

"src/gibbssampling.jl:646-650"
function sampleparticles(bm::AbstractBM, nparticles::Int, burnin::Int = 10)
   m = DeviceModel(bm); p = DeviceParticles(m, nparticles)
   check(ccall((:scdbm_particles_init, LIB), Int32, (Ptr{Cvoid}, Int32), p.h, false))
   check(ccall((:scdbm_gibbs, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64, Float64), m.h, p.h, burnin, 1.0, 1.0, 1.0))
   tohost(p)
end

"src/gibbssampling.jl:663-689"
function samples(bm::AbstractBM, nsamples::Int; burnin::Int = 50, samplelast::Bool = true)
   samplelast && return sampleparticles(bm, nsamples, burnin)[1]
   m = DeviceModel(bm); p = DeviceParticles(m, nsamples)
   check(ccall((:scdbm_particles_init, LIB), Int32, (Ptr{Cvoid}, Int32), p.h, false))
   check(ccall((:scdbm_gibbs, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64, Float64), m.h, p.h, burnin - 1, 1.0, 1.0, 1.0))
   out = Mat(defaultcontext(), nsamples, Int(m.units[1]), MAT_REAL)
   check(ccall((:scdbm_particles_visible_potential, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), m.h, p.h, out.h))
   download!(Matrix{Float64}(undef, nsamples, Int(m.units[1])), out)
end

"src/dbmtraining.jl:175-228"
function meanfield(dbm::MultimodalDBM, x::Array{Float64,2}, eps::Float64 = 0.001)
   ctx = defaultcontext(); m = DeviceModel(dbm, ctx); xm = upload(ctx, x, datakind(m))
   mu = DeviceParticles(m, size(x, 1); realvalued = true)
   check(ccall((:scdbm_meanfield, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}),
               m.h, xm.h, eps, 0, mu.h, C_NULL, C_NULL))
   tohost(mu)
end

# ----------------------------------------------------------------- learning
mutable struct LoglikelihoodOptimizer                      # src/optimizers.jl:34-39
   h::Ptr{Cvoid}; model::Union{Nothing,DeviceModel}; learningrate::Float64; sdlearningrate::Float64
end
loglikelihoodoptimizer(; learningrate::Float64 = 0.0, sdlearningrate::Float64 = 0.0) = LoglikelihoodOptimizer(C_NULL, nothing, learningrate, sdlearningrate)
function initialized(opt::LoglikelihoodOptimizer, m::DeviceModel)      # src/optimizers.jl:145-193
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_opt_create, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), m.h, ref))
   o = LoglikelihoodOptimizer(ref[], m, opt.learningrate, opt.sdlearningrate)
   finalizer(x -> ccall((:scdbm_opt_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), o); o
end
"src/optimizers.jl:207-214,263-310 (device matrices)"
computegradient!(opt::LoglikelihoodOptimizer, v::Mat, vmodel::Mat, h::Mat, hmodel::Mat, m::DeviceModel, layer::Integer = 1) =
   check(ccall((:scdbm_compute_gradient, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
               opt.h, m.h, layer - 1, v.h, vmodel.h, h.h, hmodel.h, 0))
"src/optimizers.jl:450-454,478-485,505-520"
updateparameters!(m::DeviceModel, opt::LoglikelihoodOptimizer) =
   check(ccall((:scdbm_update_parameters, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64), m.h, opt.h, -1, opt.learningrate))

"one epoch of CD/PCD, src/rbmtraining.jl:498-591"
function trainrbm!(m::DeviceModel, x::Mat, trainer::Ptr{Cvoid}; learningrate::Float64 = 0.005, cdsteps::Int = 1,
      upfactor::Float64 = 1.0, downfactor::Float64 = 1.0, perm::Vector{Int32} = Int32.(Random.randperm(x.rows) .- 1))
   GC.@preserve perm check(ccall((:scdbm_rbm_train_epoch, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Float64, Int32, Float64, Float64),
                                 trainer, x.h, perm, learningrate, cdsteps, upfactor, downfactor))
   m
end

"src/rbmtraining.jl:89-223 (fisherscoring = 0, zeroinflation = false on this path)"
function fitrbm(x::Matrix{Float64}; nhidden::Int = size(x, 2), inversedispersion::AbstractArray = ones(size(x, 2)) .* 0.5,
      epochs::Int = 10, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0, learningrate::Float64 = 0.005,
      learningrates::Vector{Float64} = fill(learningrate, epochs), pcd::Bool = true, cdsteps::Int = 1, batchsize::Int = 1,
      fisherscoring::Int = 0, zeroinflation::Bool = false, rbmtype::DataType = BernoulliRBM, monitoring::Function = (bm, epoch) -> nothing)
   (fisherscoring != 0 || zeroinflation) && error("fisherscoring / zeroinflation are outside the accelerated hot path")
   batchsize > size(x, 1) && error("Batchsize ($batchsize) exceeds number of samples ($(size(x,1))).")
   ctx = defaultcontext(); nb = rbmtype == NegativeBinomialBernoulliRBM
   xm = upload(ctx, x, nb ? MAT_COUNT : MAT_REAL)
   units = Int32[size(x, 2), nhidden]; ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_model_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), ctx.h, 1, units, nb ? VIS_NEGBIN : VIS_BERNOULLI, ref))
   m = DeviceModel(ref[], units, nb ? VIS_NEGBIN : VIS_BERNOULLI)
   disp = convert(Vector{Float64}, inversedispersion)
   check(ccall((:scdbm_model_init_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Float64}, UInt64), m.h, 0, xm.h, nb ? disp : C_NULL, 0))
   tr = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_trainer_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), m.h, 0, batchsize, pcd, tr))
   rbm = nb ? NegativeBinomialBernoulliRBM(zeros(size(x, 2), nhidden), zeros(size(x, 2)), zeros(nhidden), copy(disp), [0.5], [1.0]) :
              BernoulliRBM(zeros(size(x, 2), nhidden), zeros(size(x, 2)), zeros(nhidden))
   for epoch in 1:epochs
      trainrbm!(m, xm, tr[]; learningrate = learningrates[epoch], cdsteps = cdsteps, upfactor = upfactor, downfactor = downfactor)
      monitoring === nothing || monitoring(download!(rbm, m), epoch)
   end
   ccall((:scdbm_trainer_destroy, LIB), Int32, (Ptr{Cvoid},), tr[])
   download!(rbm, m)
end

"src/dbmtraining.jl:250-297"
function traindbm!(dbm::MultimodalDBM, x::Array{Float64,2}; epochs::Int = 10, nparticles::Int = 100, learningrate::Float64 = 0.005,
      learningrates::Vector{Float64} = learningrate * 11.0 ./ (10.0 .+ (1:epochs)), monitoring::Function = (bm, epoch) -> nothing)
   ctx = defaultcontext(); m = DeviceModel(dbm, ctx); xm = upload(ctx, x, datakind(m))
   p = DeviceParticles(m, nparticles); check(ccall((:scdbm_particles_init, LIB), Int32, (Ptr{Cvoid}, Int32), p.h, false))
   mu = DeviceParticles(m, size(x, 1); realvalued = true)
   opt = initialized(loglikelihoodoptimizer(), m)
   for epoch in 1:epochs
      check(ccall((:scdbm_dbm_pcd_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Float64, Ptr{Int32}),
                  m.h, xm.h, p.h, mu.h, opt.h, learningrates[epoch], 5, 0.001, C_NULL))
      monitoring(download!(dbm, m), epoch)
   end
   download!(dbm, m)
end

"src/evaluating.jl:18-46,164-206"
function aislogimpweights(bm::AbstractBM; ntemperatures::Int = 100, temperatures::Array{Float64,1} = collect(0:(1/ntemperatures):1),
      nparticles::Int = 100, burnin::Int = 5)
   m = DeviceModel(bm); out = zeros(nparticles)
   check(ccall((:scdbm_ais_run, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int32, UInt64, Ptr{Float64}),
               m.h, temperatures, length(temperatures), nparticles, burnin, 0, out))
   out
end

# stackrbms / fitdbm / logpartitionfunction: host orchestration over the calls above, exactly as in
# scdbm.jl_b200/api.py (stackrbms :228-284 factors; fitdbm :108-149; logpartitionfunction :967-992).
end # module
