# scDBM.jl facade over libscdbm_b200.so -- the reference-side binding.
#
# Mirrors the hot-path names of the reference's export list (src/scDBM.jl:17-64) and keeps their signatures; the
# bodies `ccall` the C ABI declared in include/scdbm_b200.h.
#
# STATUS: UNEXECUTED.  Julia is not installed in the build image or on the GPU boxes, so this file has never been
# run.  It is a mechanical mirror of the Python facade scdbm.jl_b200/api.py, which drives the same entry points in
# the tests and benches; tests/test_abi.py parses every `ccall` below and checks symbol, arity and C types against
# the prototypes the Python binding (and the header) use.
module scDBM
using Random
using SparseArrays

export BernoulliRBM, NegativeBinomialBernoulliRBM, MultimodalDBM, Particles, TrainLayer,
       fitrbm, stackrbms, fitdbm, traindbm!, trainrbm!, gibbssample!, gibbssamplenegativebinomial!,
       gibbssamplecond!, gibbssamplenegativebinomial_cond!, gibbssamplecondrow!,
       samples, sampleparticles, initparticles, meanfield,
       hiddeninput, hiddenpotential, samplehidden, visibleinput, visiblepotential, samplevisible,
       hiddeninput!, hiddenpotential!, samplehidden!, visibleinput!, visiblepotential!, samplevisible!,
       aislogimpweights, logmeanexp, logpartitionfunction, logpartitionfunctionzeroweights,
       loglikelihoodoptimizer, initialized, computegradient!, updateparameters!,
       estimatedropout, sampledropout!, mle_for_θ

const LIB = joinpath(@__DIR__, "..", "scdbm.jl_b200", "libscdbm_b200.so")
const Particles = Array{Array{Float64,2},1}

abstract type AbstractRBM end
struct BernoulliRBM <: AbstractRBM                      # src/bmtypes.jl:17-21
   weights::Array{Float64,2}; visbias::Array{Float64,1}; hidbias::Array{Float64,1}
end
mutable struct NegativeBinomialBernoulliRBM <: AbstractRBM      # src/bmtypes.jl:85-92
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
setoption!(ctx::Context, name::String, value::Real) =
   check(ccall((:scdbm_ctx_set_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Float64), ctx.h, name, value))
function getoption(ctx::Context, name::String)
   v = Ref{Float64}(0.0)
   check(ccall((:scdbm_ctx_get_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Ref{Float64}), ctx.h, name, v))
   v[]
end
seed!(s::Integer) = setoption!(defaultcontext(), "seed", s)
synchronize(ctx::Context = defaultcontext()) = check(ccall((:scdbm_sync, LIB), Int32, (Ptr{Cvoid},), ctx.h))

# ---- data-parallel communicator (NCCL inside the library; one Julia process per GPU, e.g. under MPI.jl / Distributed)
function communiqueid()
   id = Vector{UInt8}(undef, 128)
   check(ccall((:scdbm_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
   id
end
cominit!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
   check(ccall((:scdbm_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx.h, id, rank, nranks))
comdestroy!(ctx::Context) = check(ccall((:scdbm_comm_destroy, LIB), Int32, (Ptr{Cvoid},), ctx.h))
function allreduce!(ctx::Context, values::Vector{Float64}; op::Symbol = :sum)
   check(ccall((:scdbm_comm_allreduce_f64, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), ctx.h, values, length(values), op == :max ? 1 : 0))
   values
end

rbms(bm::AbstractRBM) = AbstractRBM[bm]
rbms(bm::MultimodalDBM) = bm
viskind(rbm) = rbm isa NegativeBinomialBernoulliRBM ? VIS_NEGBIN : VIS_BERNOULLI

mutable struct DeviceModel
   h::Ptr{Cvoid}; units::Vector{Int32}; kind::Int32; ctx::Context
end
function DeviceModel(ctx::Context, units::Vector{Int32}, kind::Int32)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_model_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
               ctx.h, length(units) - 1, units, kind, ref))
   m = DeviceModel(ref[], units, kind, ctx)
   finalizer(x -> ccall((:scdbm_model_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), m)
   m
end
function setlayer!(m::DeviceModel, l::Integer, r::AbstractRBM)
   disp = r isa NegativeBinomialBernoulliRBM ? convert(Vector{Float64}, r.inversedispersion) : C_NULL
   GC.@preserve r disp check(ccall((:scdbm_model_set_layer, LIB), Int32,
         (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
         m.h, l - 1, r.weights, size(r.weights, 1), r.visbias, r.hidbias, disp))
end
function DeviceModel(bm::AbstractBM, ctx::Context = defaultcontext())
   rs = rbms(bm)
   units = Int32[size(rs[1].weights, 1); [size(r.weights, 2) for r in rs]]
   m = DeviceModel(ctx, units, viskind(rs[1]))
   for (l, r) in enumerate(rs); setlayer!(m, l, r); end
   m
end
function download!(bm::AbstractBM, m::DeviceModel)          # parameters back into the Julia structs
   for (l, r) in enumerate(rbms(bm))
      disp = r isa NegativeBinomialBernoulliRBM ? convert(Vector{Float64}, r.inversedispersion) : C_NULL
      GC.@preserve r disp check(ccall((:scdbm_model_get_layer, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            m.h, l - 1, r.weights, size(r.weights, 1), r.visbias, r.hidbias, disp))
      r isa NegativeBinomialBernoulliRBM && (r.inversedispersion = disp)
   end
   bm
end
function hostmodel(m::DeviceModel)                           # fresh Julia structs holding the device parameters
   out = AbstractRBM[]
   for l in 1:length(m.units)-1
      V, H = Int(m.units[l]), Int(m.units[l+1])
      push!(out, (l == 1 && m.kind == VIS_NEGBIN) ?
            NegativeBinomialBernoulliRBM(zeros(V, H), zeros(V), zeros(H), ones(V), [0.5], [1.0]) : BernoulliRBM(zeros(V, H), zeros(V), zeros(H)))
   end
   bm = length(out) == 1 ? out[1] : out
   download!(bm, m)
end

# A device matrix.  `owner` keeps the object alive from which a shared handle was obtained (particles, trainer): the
# library counts the handle as one more reference, so it stays valid either way and is released by the finalizer.
mutable struct Mat
   h::Ptr{Cvoid}; rows::Int; cols::Int; owner::Any
end
function adopt(h::Ptr{Cvoid}, owner = nothing)
   r = Ref{Int64}(0); c = Ref{Int64}(0); k = Ref{Int32}(0)
   check(ccall((:scdbm_mat_shape, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int32}), h, r, c, k))
   m = Mat(h, Int(r[]), Int(c[]), owner)
   finalizer(x -> ccall((:scdbm_mat_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), m)
   m
end
function Mat(ctx::Context, rows::Integer, cols::Integer, kind::Int32)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_mat_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ref{Ptr{Cvoid}}), ctx.h, rows, cols, kind, ref))
   adopt(ref[], ctx)
end
function upload!(m::Mat, x::Matrix{Float64})
   GC.@preserve x check(ccall((:scdbm_mat_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), m.h, x, max(size(x, 1), 1)))
   m
end
upload(ctx::Context, x::Matrix{Float64}, kind::Int32) = upload!(Mat(ctx, size(x, 1), size(x, 2), kind), x)
# sparse data (cells x genes SparseMatrixCSC): only the non-zeros cross PCIe
function upload(ctx::Context, x::SparseMatrixCSC{Float64,<:Integer}, kind::Int32)
   m = Mat(ctx, size(x, 1), size(x, 2), kind)
   ptr = Int64.(x.colptr .- 1); idx = Int32.(x.rowval .- 1); val = Float32.(x.nzval)
   GC.@preserve ptr idx val check(ccall((:scdbm_mat_upload_compressed, LIB), Int32,
      (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Float32}), m.h, Int32(1), ptr, idx, val))
   m
end
upload(ctx::Context, x::Mat, kind::Int32) = x
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
tohost(m::Mat) = download!(Matrix{Float64}(undef, m.rows, m.cols), m)
# compact row-major uint16 counts (cells are the columns of the returned genes x cells matrix)
function downloadu16(m::Mat)
   out = Matrix{UInt16}(undef, m.cols, m.rows)
   GC.@preserve out check(ccall((:scdbm_mat_download_u16, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt16}), m.h, out))
   out
end
function columnstats(m::Mat)      # per-gene mean, variance, zero fraction (src/NegativeBinomialRBMPlots.jl:148-182)
   mean = zeros(m.cols); var = zeros(m.cols); zf = zeros(m.cols)
   check(ccall((:scdbm_mat_column_stats, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), m.h, mean, var, zf))
   mean, var, zf
end
datakind(m::DeviceModel) = m.kind == VIS_NEGBIN ? MAT_COUNT : MAT_BINARY

# ------------------------------------------------- conditional operators
function conditional!(out::Matrix{Float64}, sym::Symbol, rbm::AbstractRBM, in::Matrix{Float64}, what::Int32,
      factor::Float64, inkind::Int32, outkind::Int32)
   ctx = defaultcontext(); m = DeviceModel(rbm, ctx)
   src = upload(ctx, in, inkind); dst = Mat(ctx, size(out, 1), size(out, 2), outkind)
   if sym == :hidden
      check(ccall((:scdbm_hidden, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, UInt32, Ptr{Float64}, Int64),
                  m.h, 0, src.h, dst.h, factor, what, 0, C_NULL, 0))
   else
      check(ccall((:scdbm_visible, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, UInt32, Ptr{Float64}, Int64),
                  m.h, 0, src.h, dst.h, factor, what, 0, C_NULL, 0))
   end
   download!(out, dst)
end
viskindof(rbm) = rbm isa NegativeBinomialBernoulliRBM ? MAT_COUNT : MAT_REAL
# src/gibbssampling.jl:197-212,264-276,322-335,589-598
hiddeninput!(hh, rbm, vv) = conditional!(hh, :hidden, rbm, vv, OUT_INPUT, 1.0, viskindof(rbm), MAT_REAL)
hiddenpotential!(hh, rbm, vv, factor::Float64 = 1.0) = conditional!(hh, :hidden, rbm, vv, OUT_POTENTIAL, factor, viskindof(rbm), MAT_REAL)
samplehidden!(hh, rbm, vv, factor::Float64 = 1.0) = conditional!(hh, :hidden, rbm, vv, OUT_SAMPLE, factor, viskindof(rbm), MAT_BINARY)
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
init!(p::DeviceParticles; biased::Bool = false) = (check(ccall((:scdbm_particles_init, LIB), Int32, (Ptr{Cvoid}, Int32), p.h, biased)); p)
function layer(p::DeviceParticles, i::Integer)       # shared handle (one more reference), keeps `p` alive as well
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_particles_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), p.h, i - 1, ref))
   adopt(ref[], p)
end
function clock(p::DeviceParticles)
   c = Ref{Int64}(0)
   check(ccall((:scdbm_particles_clock, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Int64), p.h, c, -1))
   c[]
end
tohost(p::DeviceParticles) = Particles([tohost(layer(p, i)) for i in 1:length(p.model.units)])
function upload!(p::DeviceParticles, particles::Particles)
   for (i, x) in enumerate(particles); upload!(layer(p, i), x); end
   p
end
gibbs!(m::DeviceModel, p::DeviceParticles, nsteps::Integer; temperature::Float64 = 1.0, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0) =
   check(ccall((:scdbm_gibbs, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64, Float64), m.h, p.h, nsteps, temperature, upfactor, downfactor))

"src/gibbssampling.jl:386-419"
initparticles(bm::AbstractBM, nparticles::Int; biased::Bool = false) = tohost(init!(DeviceParticles(DeviceModel(bm), nparticles); biased = biased))

"src/gibbssampling.jl:62-70 (RBM), :153-182 (DBM)"
function gibbssample!(particles::Particles, bm::AbstractBM, nsteps::Int = 5, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0)
   m = DeviceModel(bm); p = upload!(DeviceParticles(m, size(particles[1], 1)), particles)
   gibbs!(m, p, nsteps; upfactor = upfactor, downfactor = downfactor)
   for (i, x) in enumerate(particles); download!(x, layer(p, i)); end
   particles
end

# sampledropout (src/rbmtraining.jl:289-311) applied to the visible states v given the data x
function sampledropout!(v::Mat, x::Mat, mid::Vector{Float64}, shape::Vector{Float64}; log1pform::Bool = false, tick::Integer = 0, rowoffset::Integer = 0)
   check(ccall((:scdbm_sample_dropout, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, UInt32, UInt64),
               v.h, x.h, mid, shape, length(mid), log1pform, tick, rowoffset))
   v
end

"src/gibbssampling.jl:82-110 (RBM form) and :112-151 (DBM form); `zeroinflation` applies `sampledropout` to the result"
function gibbssamplenegativebinomial!(x::Array{Float64,2}, particles::Particles, bm::AbstractBM,
      nsteps::Int = 5, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0; zeroinflation::Bool = true)
   ctx = defaultcontext(); m = DeviceModel(bm, ctx); p = upload!(DeviceParticles(m, size(particles[1], 1)), particles)
   if bm isa AbstractRBM
      check(ccall((:scdbm_gibbs_negbin_rbm, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Float64), m.h, p.h, nsteps, upfactor, downfactor))
   else
      gibbs!(m, p, nsteps)
   end
   first = rbms(bm)[1]
   if zeroinflation && first isa NegativeBinomialBernoulliRBM
      xm = upload(ctx, x, MAT_COUNT)
      dbmform = !(bm isa AbstractRBM)                         # :144-148 uses log1p(x) and per-gene parameters
      sampledropout!(layer(p, 1), xm, first.dropoutmid, first.dropoutshape; log1pform = dbmform && length(first.dropoutmid) > 1)
   end
   for (i, xx) in enumerate(particles); download!(xx, layer(p, i)); end
   particles
end

"gibbssamplecond! / gibbssamplenegativebinomial_cond! / gibbssamplecondrow! (src/gibbssampling.jl:1042-1237)"
function gibbssamplecond!(particles::Particles, bm::AbstractBM, fixedvars, nsteps::Int = 5)
   m = DeviceModel(bm); n = size(particles[1], 1); V = Int(m.units[1])
   p = upload!(DeviceParticles(m, n), particles)
   if fixedvars isa AbstractMatrix{Bool}
      mask = Matrix{UInt8}(fixedvars)
      GC.@preserve mask check(ccall((:scdbm_gibbs_cond, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int32), m.h, p.h, mask, n, n, nsteps))
   else
      mask = zeros(UInt8, V)
      fixedvars isa AbstractVector{Bool} ? (mask .= fixedvars) : (mask[collect(fixedvars)] .= 1)
      GC.@preserve mask check(ccall((:scdbm_gibbs_cond, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int32), m.h, p.h, mask, 0, 0, nsteps))
   end
   for (i, x) in enumerate(particles); download!(x, layer(p, i)); end
   particles
end
const gibbssamplenegativebinomial_cond! = gibbssamplecond!
const gibbssamplecondrow! = gibbssamplecond!

"src/gibbssampling.jl:646-650"
function sampleparticles(bm::AbstractBM, nparticles::Int, burnin::Int = 10)
   m = DeviceModel(bm); p = init!(DeviceParticles(m, nparticles))
   gibbs!(m, p, burnin)
   tohost(p)
end

"src/gibbssampling.jl:663-689"
function samples(bm::AbstractBM, nsamples::Int; burnin::Int = 50, samplelast::Bool = true)
   samplelast && return sampleparticles(bm, nsamples, burnin)[1]
   m = DeviceModel(bm); p = init!(DeviceParticles(m, nsamples))
   gibbs!(m, p, burnin - 1)
   out = Mat(defaultcontext(), nsamples, Int(m.units[1]), MAT_REAL)
   check(ccall((:scdbm_particles_visible_potential, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), m.h, p.h, out.h))
   tohost(out)
end

"src/dbmtraining.jl:175-228"
function meanfield(dbm::MultimodalDBM, x::Array{Float64,2}, eps::Float64 = 0.001)
   ctx = defaultcontext(); m = DeviceModel(dbm, ctx); xm = upload(ctx, x, datakind(m))
   mu = DeviceParticles(m, size(x, 1); realvalued = true)
   check(ccall((:scdbm_meanfield, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}),
               m.h, xm.h, eps, 0, mu.h, C_NULL, C_NULL))
   GC.@preserve xm tohost(mu)
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
"src/optimizers.jl:431-442 (StackedOptimizer); pos/neg scale > 0: the global 1/n+ and 1/n- of a data-parallel run"
computegradient!(opt::LoglikelihoodOptimizer, mu::DeviceParticles, particles::DeviceParticles, m::DeviceModel; posscale::Float64 = 0.0, negscale::Float64 = 0.0) =
   check(ccall((:scdbm_compute_gradient_dbm, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64),
               opt.h, m.h, mu.h, particles.h, posscale, negscale))
allreducegradient!(opt::LoglikelihoodOptimizer) = check(ccall((:scdbm_allreduce_gradient, LIB), Int32, (Ptr{Cvoid},), opt.h))
"src/optimizers.jl:450-454,478-485,505-520"
updateparameters!(m::DeviceModel, opt::LoglikelihoodOptimizer) =
   check(ccall((:scdbm_update_parameters, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64), m.h, opt.h, -1, opt.learningrate))

mutable struct RBMTrainer                                   # state fitrbm keeps across epochs (src/rbmtraining.jl:137-151)
   h::Ptr{Cvoid}; model::DeviceModel
end
function RBMTrainer(m::DeviceModel, batchsize::Integer, pcd::Bool; layer::Integer = 1)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_trainer_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), m.h, layer - 1, batchsize, pcd, ref))
   t = RBMTrainer(ref[], m); finalizer(x -> ccall((:scdbm_trainer_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), t); t
end
trackmean!(t::RBMTrainer, on::Bool = true) = check(ccall((:scdbm_trainer_track_mean, LIB), Int32, (Ptr{Cvoid}, Int32), t.h, on))
function estimatedmean(t::RBMTrainer)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_trainer_estimated_mean, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), t.h, ref))
   adopt(ref[], t)
end
function chainstate(t::RBMTrainer)
   ref = Ref{Ptr{Cvoid}}(C_NULL)
   check(ccall((:scdbm_trainer_chainstate, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), t.h, ref))
   adopt(ref[], t)
end

"one epoch of CD/PCD, src/rbmtraining.jl:498-591"
function trainrbm!(m::DeviceModel, x::Mat, trainer::RBMTrainer; learningrate::Float64 = 0.005, cdsteps::Int = 1,
      upfactor::Float64 = 1.0, downfactor::Float64 = 1.0, perm::Vector{Int32} = Int32.(Random.randperm(x.rows) .- 1))
   GC.@preserve perm check(ccall((:scdbm_rbm_train_epoch, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Float64, Int32, Float64, Float64),
                                 trainer.h, x.h, perm, learningrate, cdsteps, upfactor, downfactor))
   m
end

"`mle_for_θ` for every gene at once (src/rbmtraining.jl:675-719)"
function mle_for_θ(x::Mat, mu::Mat; lambda::Float64 = 100.0, maxiter::Int = 20, convtol::Float64 = 1e-6)
   theta = zeros(x.cols)
   check(ccall((:scdbm_mle_theta, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Float64, Ptr{Float64}), x.h, mu.h, lambda, maxiter, convtol, theta))
   theta
end

"`estimatedropout!` (src/rbmtraining.jl:271-287): binomial GLM of the per-gene zero fraction on log1p(per-gene mean)"
function estimatedropout(x::Mat)
   mean, _, zf = columnstats(x)
   X = hcat(ones(length(mean)), log1p.(mean)); beta = zeros(2)
   for _ in 1:100
      eta = X * beta; mu = 1.0 ./ (1.0 .+ exp.(-eta)); w = max.(mu .* (1.0 .- mu), 1e-12)
      z = eta .+ (zf .- mu) ./ w
      new = (X' * (w .* X)) \ (X' * (w .* z))
      done = maximum(abs.(new .- beta)) < 1e-12
      beta = new
      done && break
   end
   beta[1], beta[2]
end

"src/rbmtraining.jl:89-223.  NB RBMs: `fisherscoring != 0` re-estimates the inverse dispersion every 5th epoch on the device
(`estimatedispersion = \"gene\"`; the reference's other branches do not run, SURVEY B4), `zeroinflation` fits the dropout logistic."
function fitrbmdevice(xm::Mat, ctx::Context; nhidden::Int = xm.cols, inversedispersion::AbstractArray = ones(xm.cols) .* 0.5,
      epochs::Int = 10, upfactor::Float64 = 1.0, downfactor::Float64 = 1.0, learningrate::Float64 = 0.005,
      learningrates::Vector{Float64} = fill(learningrate, epochs), pcd::Bool = true, cdsteps::Int = 1, batchsize::Int = 1,
      fisherscoring::Int = 1, lambda::Float64 = 100.0, rbmtype::DataType = BernoulliRBM,
      startrbm::Union{Nothing,AbstractRBM} = nothing, monitoring::Function = (bm, epoch) -> nothing)
   length(learningrates) < epochs && error("Not enough `learningrates` (vector of length $(length(learningrates))) for training epochs ($epochs)")
   batchsize > xm.rows && error("Batchsize ($batchsize) exceeds number of samples ($(xm.rows)).")
   nb = rbmtype == NegativeBinomialBernoulliRBM
   if startrbm === nothing
      m = DeviceModel(ctx, Int32[xm.cols, nhidden], nb ? VIS_NEGBIN : VIS_BERNOULLI)
      disp = convert(Vector{Float64}, inversedispersion)
      GC.@preserve disp check(ccall((:scdbm_model_init_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Float64}, UInt64),
                                    m.h, 0, xm.h, nb ? pointer(disp) : Ptr{Float64}(C_NULL), UInt64(getoption(ctx, "seed"))))
   else
      m = DeviceModel(startrbm, ctx)
   end
   tr = RBMTrainer(m, batchsize, pcd)
   nb && fisherscoring != 0 && trackmean!(tr, true)
   for epoch in 1:epochs
      trainrbm!(m, xm, tr; learningrate = learningrates[epoch], cdsteps = cdsteps, upfactor = upfactor, downfactor = downfactor)
      monitoring(hostmodel(m), epoch)
      if nb && fisherscoring != 0 && epoch % 5 == 0                       # src/rbmtraining.jl:172-183
         theta = mle_for_θ(xm, estimatedmean(tr); lambda = lambda)
         GC.@preserve theta check(ccall((:scdbm_model_set_layer, LIB), Int32,
               (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), m.h, 0, C_NULL, 0, C_NULL, C_NULL, theta))
      end
   end
   m
end
function fitrbm(x::Union{Matrix{Float64},SparseMatrixCSC{Float64,<:Integer}}; rbmtype::DataType = BernoulliRBM, zeroinflation::Bool = true, kwargs...)
   ctx = defaultcontext(); nb = rbmtype == NegativeBinomialBernoulliRBM
   xm = upload(ctx, x, nb ? MAT_COUNT : MAT_REAL)
   rbm = hostmodel(fitrbmdevice(xm, ctx; rbmtype = rbmtype, kwargs...))
   if nb && zeroinflation                                                  # src/rbmtraining.jl:531-534
      mid, shape = estimatedropout(xm)
      rbm.dropoutmid = [mid]; rbm.dropoutshape = [shape]
   end
   rbm
end

"The hot-path subset of `TrainLayer` (src/rbmstacking.jl:9-94)"
Base.@kwdef struct TrainLayer
   nhidden::Int = -1
   rbmtype::DataType = BernoulliRBM
   epochs::Int = -1
   learningrate::Float64 = -Inf
   batchsize::Int = -1
   pcd::Bool = true
   cdsteps::Int = 1
   inversedispersion::AbstractArray = Float64[]
   startrbm::Union{Nothing,AbstractRBM} = nothing
end

"`stackrbms` (src/rbmstacking.jl:228-284; pre-training factors :247-275).  Returns the device model of the stack."
function stackrbmsdevice(xm::Mat, ctx::Context; nhiddens::Vector{Int} = Int[], epochs::Int = 10, predbm::Bool = false,
      learningrate::Float64 = 0.005, batchsize::Int = 1, trainlayers::Vector{TrainLayer} = TrainLayer[])
   isempty(trainlayers) && (trainlayers = [TrainLayer(nhidden = n) for n in (isempty(nhiddens) ? [xm.cols, xm.cols] : nhiddens)])
   baseseed = UInt64(getoption(ctx, "seed"))
   layers = DeviceModel[]
   upfactor, downfactor = predbm ? 2.0 : 1.0, 1.0
   hiddenval = xm
   for (i, tl) in enumerate(trainlayers)
      if i > 1
         prev = layers[i-1]
         hv = Mat(ctx, hiddenval.rows, Int(prev.units[2]), MAT_REAL)
         check(ccall((:scdbm_hidden, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, UInt32, Ptr{Float64}, Int64),
                     prev.h, 0, hiddenval.h, hv.h, upfactor, OUT_POTENTIAL, 0, C_NULL, 0))
         hiddenval = hv
         if predbm
            upfactor = downfactor = 2.0
            i == length(trainlayers) && (upfactor = 1.0)
         else
            upfactor = downfactor = 1.0
         end
      end
      setoption!(ctx, "seed", (baseseed + 1000003 * i) % (UInt64(1) << 52))
      push!(layers, fitrbmdevice(hiddenval, ctx; nhidden = tl.nhidden, epochs = tl.epochs >= 0 ? tl.epochs : epochs,
            upfactor = upfactor, downfactor = downfactor, learningrate = tl.learningrate >= 0 ? tl.learningrate : learningrate,
            pcd = tl.pcd, cdsteps = tl.cdsteps, batchsize = tl.batchsize > 0 ? tl.batchsize : batchsize, rbmtype = tl.rbmtype,
            inversedispersion = isempty(tl.inversedispersion) ? ones(hiddenval.cols) .* 0.5 : tl.inversedispersion,
            fisherscoring = 0, startrbm = tl.startrbm))
   end
   setoption!(ctx, "seed", baseseed)
   units = Int32[layers[1].units[1]; [l.units[2] for l in layers]]
   dbm = DeviceModel(ctx, units, layers[1].kind)
   for (i, l) in enumerate(layers)
      check(ccall((:scdbm_model_copy_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32), dbm.h, i - 1, l.h, 0))
   end
   dbm
end
function stackrbms(x::Union{Matrix{Float64},SparseMatrixCSC{Float64,<:Integer}}; trainlayers::Vector{TrainLayer} = TrainLayer[], kwargs...)
   ctx = defaultcontext()
   nb0 = !isempty(trainlayers) && trainlayers[1].rbmtype == NegativeBinomialBernoulliRBM
   hostmodel(stackrbmsdevice(upload(ctx, x, nb0 ? MAT_COUNT : MAT_REAL), ctx; trainlayers = trainlayers, kwargs...))
end

defaultfinetuninglearningrates(learningrate, epochs) = learningrate * 11.0 ./ (10.0 .+ (1:epochs))   # src/dbmtraining.jl:59-61

"`traindbm!` on device objects (src/dbmtraining.jl:250-297).  With a communicator on `m.ctx` the step is data-parallel:
`xm` / `p` are this rank's shards and `globalnsamples` / `globalnparticles` the totals over all ranks."
function traindbmdevice!(m::DeviceModel, xm::Mat, p::DeviceParticles; epochs::Int = 10, learningrate::Float64 = 0.005,
      learningrates::AbstractVector{Float64} = defaultfinetuninglearningrates(learningrate, epochs), nsteps::Int = 5, eps::Float64 = 0.001,
      globalnsamples::Integer = 0, globalnparticles::Integer = 0, monitoring::Function = (bm, epoch) -> nothing)
   length(learningrates) < epochs && error("Not enough `learningrates` for training epochs")
   mu = DeviceParticles(m, xm.rows; realvalued = true)
   opt = initialized(loglikelihoodoptimizer(), m)
   dp = getoption(m.ctx, "comm_size") > 1
   for epoch in 1:epochs
      if dp
         check(ccall((:scdbm_dbm_pcd_step_dp, LIB), Int32,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Float64, Int64, Int64, Ptr{Int32}),
               m.h, xm.h, p.h, mu.h, opt.h, learningrates[epoch], nsteps, eps, globalnsamples, globalnparticles, C_NULL))
      else
         check(ccall((:scdbm_dbm_pcd_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Float64, Ptr{Int32}),
                     m.h, xm.h, p.h, mu.h, opt.h, learningrates[epoch], nsteps, eps, C_NULL))
      end
      monitoring(hostmodel(m), epoch)
   end
   m
end
function traindbm!(dbm::MultimodalDBM, x::Union{Matrix{Float64},SparseMatrixCSC{Float64,<:Integer}}; nparticles::Int = 100, kwargs...)
   ctx = defaultcontext(); m = DeviceModel(dbm, ctx); xm = upload(ctx, x, datakind(m))
   traindbmdevice!(m, xm, init!(DeviceParticles(m, nparticles)); kwargs...)
   download!(dbm, m)
end

"`fitdbm` (src/dbmtraining.jl:108-149): stackrbms(predbm = true) followed by traindbm!"
function fitdbm(x::Union{Matrix{Float64},SparseMatrixCSC{Float64,<:Integer}}; nhiddens::Vector{Int} = Int[], epochs::Int = 10, nparticles::Int = 100,
      learningrate::Float64 = 0.005, learningrates::AbstractVector{Float64} = defaultfinetuninglearningrates(learningrate, epochs),
      learningratepretraining::Float64 = learningrate, epochspretraining::Int = epochs, batchsizepretraining::Int = 1,
      pretraining::Vector{TrainLayer} = TrainLayer[], monitoring::Function = (bm, epoch) -> nothing)
   ctx = defaultcontext()
   nb0 = !isempty(pretraining) && pretraining[1].rbmtype == NegativeBinomialBernoulliRBM
   xm = upload(ctx, x, nb0 ? MAT_COUNT : MAT_REAL)
   m = stackrbmsdevice(xm, ctx; nhiddens = nhiddens, epochs = epochspretraining, predbm = true, learningrate = learningratepretraining,
                       batchsize = batchsizepretraining, trainlayers = pretraining)
   traindbmdevice!(m, xm, init!(DeviceParticles(m, nparticles)); epochs = epochs, learningrates = learningrates, monitoring = monitoring)
   hostmodel(m)
end

# ------------------------------------------------------------------ AIS
"src/evaluating.jl:18-46,164-206; `rowoffset` = global index of the first particle of this shard"
function aislogimpweights(bm::AbstractBM; ntemperatures::Int = 100, temperatures::Array{Float64,1} = collect(0:(1/ntemperatures):1),
      nparticles::Int = 100, burnin::Int = 5, rowoffset::Integer = 0)
   m = DeviceModel(bm); out = zeros(nparticles)
   check(ccall((:scdbm_ais_run, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int32, UInt64, Ptr{Float64}),
               m.h, temperatures, length(temperatures), nparticles, burnin, rowoffset, out))
   out
end
"src/evaluating.jl:935-943"
function logmeanexp(v::AbstractVector{Float64})
   vmax = maximum(v)
   vmax + log(sum(exp.(v .- vmax))) - log(length(v))
end
softplus(x) = max(x, 0.0) + log1p(exp(-abs(x)))
"src/evaluating.jl:1001-1031,1055-1065 (O(units) scalar work: host)"
function logpartitionfunctionzeroweights(bm::AbstractBM)
   rs = rbms(bm); first = rs[1]
   vis = first isa NegativeBinomialBernoulliRBM ? sum(-first.inversedispersion .* log.(1.0 .- exp.(first.visbias))) : sum(softplus.(first.visbias))
   length(rs) == 1 && return vis + sum(softplus.(first.hidbias))
   tot = vis
   for i in 2:length(rs); tot += sum(softplus.(rs[i].visbias .+ rs[i-1].hidbias)); end
   tot + sum(softplus.(rs[end].hidbias))
end
"src/evaluating.jl:967-992"
logpartitionfunction(bm::AbstractBM; kwargs...) = logmeanexp(aislogimpweights(bm; kwargs...)) + logpartitionfunctionzeroweights(bm)
logpartitionfunction(bm::AbstractBM, logr::Float64) = logr + logpartitionfunctionzeroweights(bm)

end # module
