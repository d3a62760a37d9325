"""Host-side mirror of the reference's Julia API for the hot path
(export list src/scDBM.jl:17-64), written in Python because Julia is not
available in this image; `julia/scDBM.jl` is the same facade over `ccall`.

Naming: Julia's `f!` is `f_` here.  Matrices are NumPy Float64, logical shapes
as in the reference (samples x units, weights nvisible x nhidden).  Functions
take either host objects (NumPy arrays / RBM dataclasses: uploaded, processed
on the GPU, downloaded) or device handles (`DeviceModel`, `Mat`,
`DeviceParticles`: stay resident in HBM).
"""
import ctypes as C
import copy
from dataclasses import dataclass, field
import numpy as np

from . import _lib as L
from ._lib import lib, check

_D = L.D


def _f64(a, order="F"):
    return np.asarray(a, dtype=np.float64, order=order)


def _ptr(a):
    return a.ctypes.data_as(_D) if a is not None else None


# ------------------------------------------------------------------ context
class Context:
    """One GPU + stream + Philox seed (`Random.seed!` stands behind `seed`)."""

    def __init__(self, device=0, seed=0, stream=None):
        self.h = C.c_void_p()
        check(lib().scdbm_ctx_create(device, C.c_uint64(seed), stream, C.byref(self.h)))
        self.device = device

    def set(self, name, value):
        check(lib().scdbm_ctx_set_option(self.h, name.encode(), float(value)))

    def get(self, name):
        v = C.c_double()
        check(lib().scdbm_ctx_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def sync(self):
        check(lib().scdbm_sync(self.h))

    def wait_downloads(self):
        """Blocks until the host buffers of `Mat.to_host_u16(out, wait=False)` calls are filled."""
        check(lib().scdbm_wait_downloads(self.h))

    def timer_start(self):
        check(lib().scdbm_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().scdbm_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        n = C.c_int64()
        check(lib().scdbm_launch_count(self.h, C.byref(n), int(reset)))
        return n.value

    # ---- data-parallel communicator (NCCL inside the library)
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id; rank 0 creates it and hands it to the other ranks out of band."""
        buf = (C.c_uint8 * 128)()
        check(lib().scdbm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, nranks):
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of Context.comm_unique_id()")
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().scdbm_comm_init(self.h, buf, int(rank), int(nranks)))
        return self

    def comm_destroy(self):
        check(lib().scdbm_comm_destroy(self.h))

    @property
    def comm_size(self):
        return int(self.get("comm_size"))

    @property
    def comm_rank(self):
        return int(self.get("comm_rank"))

    def allreduce(self, values, op="sum"):
        """Sum / max of a few host scalars over the ranks (logmeanexp over particle shards, global sample counts)."""
        v = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        check(lib().scdbm_comm_allreduce_f64(self.h, _ptr(v), len(v), {"sum": 0, "max": 1}[op]))
        return v

    def close(self):
        if self.h:
            lib().scdbm_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_unique_id(make_id, rank, broadcast):
    """Rank 0 calls `make_id()`; `broadcast(obj_or_None) -> obj` carries it to every rank (torch.distributed
    `broadcast_object_list`, MPI bcast, a file ...).  Returns the 128-byte id on every rank."""
    uid = make_id() if rank == 0 else None
    uid = broadcast(uid)
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id exchange failed: expected 128 bytes on every rank")
    return bytes(uid)


def init_comm_from_torch(ctx):
    """Gives `ctx` a communicator spanning the ranks of an initialised `torch.distributed` process group (used only
    to carry the 128-byte id; the collectives of the hot path are the library's own NCCL calls)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    return ctx.comm_init(exchange_unique_id(Context.comm_unique_id, rank, bcast), rank, world)


def logmeanexp_sharded(w, ctx):
    """`logmeanexp` (src/evaluating.jl:935-943) of importance weights sharded over the ranks of `ctx`'s communicator."""
    w = np.asarray(w, float)
    wmax = float(ctx.allreduce([w.max() if len(w) else -np.inf], "max")[0])
    s, n = ctx.allreduce([np.exp(w - wmax).sum(), float(len(w))], "sum")
    return wmax + np.log(s) - np.log(n)


_default_ctx = None


def default_context(seed=None):
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0, 0 if seed is None else seed)
    elif seed is not None:
        _default_ctx.set("seed", seed)
    return _default_ctx


def seed_(seed):
    """`Random.seed!(seed)`"""
    return default_context(seed)


# -------------------------------------------------------------- host models
@dataclass
class BernoulliRBM:
    """src/bmtypes.jl:17-21"""
    weights: np.ndarray
    visbias: np.ndarray
    hidbias: np.ndarray
    kind = L.VIS_BERNOULLI


@dataclass
class NegativeBinomialBernoulliRBM:
    """src/bmtypes.jl:85-92"""
    weights: np.ndarray
    visbias: np.ndarray
    hidbias: np.ndarray
    inversedispersion: np.ndarray
    dropoutmid: np.ndarray = field(default_factory=lambda: np.array([0.5]))
    dropoutshape: np.ndarray = field(default_factory=lambda: np.array([1.0]))
    kind = L.VIS_NEGBIN


def _rbms(bm):
    return list(bm) if isinstance(bm, (list, tuple)) else [bm]


# ------------------------------------------------------------ device model
class DeviceModel:
    """RBM / DBM parameters resident on the GPU (fp32 master + bf16 operand
    planes)."""

    def __init__(self, ctx, units, visible_kind):
        self.ctx = ctx
        self.units = [int(u) for u in units]
        self.visible_kind = visible_kind
        self.h = C.c_void_p()
        arr = (C.c_int32 * len(self.units))(*self.units)
        check(lib().scdbm_model_create(ctx.h, len(self.units) - 1, arr, visible_kind, C.byref(self.h)))

    @property
    def nrbms(self):
        return len(self.units) - 1

    @classmethod
    def from_host(cls, bm, ctx=None):
        ctx = ctx or default_context()
        rbms = _rbms(bm)
        units = [rbms[0].weights.shape[0]] + [r.weights.shape[1] for r in rbms]
        for i in range(1, len(rbms)):
            if rbms[i].weights.shape[0] != units[i]:
                raise ValueError("DBM layer sizes are inconsistent")
            if rbms[i].kind != L.VIS_BERNOULLI:
                raise ValueError("only the first layer may have NB visible units")
        m = cls(ctx, units, rbms[0].kind)
        for l, r in enumerate(rbms):
            m.set_layer(l, r)
        return m

    def set_layer(self, l, rbm):
        W = _f64(rbm.weights)
        if W.shape != (self.units[l], self.units[l + 1]):
            raise ValueError("weights have the wrong shape")
        a, b = _f64(rbm.visbias), _f64(rbm.hidbias)
        if a.shape != (self.units[l],) or b.shape != (self.units[l + 1],):
            raise ValueError("bias vectors have the wrong length")
        r = _f64(rbm.inversedispersion) if rbm.kind == L.VIS_NEGBIN else None
        if r is not None and r.shape != (self.units[l],):
            raise ValueError("inversedispersion has the wrong length")
        check(lib().scdbm_model_set_layer(self.h, l, _ptr(W), W.shape[0], _ptr(a), _ptr(b), _ptr(r)))

    def get_layer(self, l):
        V, H = self.units[l], self.units[l + 1]
        W = np.zeros((V, H), order="F")
        a, b, r = np.zeros(V), np.zeros(H), np.zeros(V)
        check(lib().scdbm_model_get_layer(self.h, l, _ptr(W), V, _ptr(a), _ptr(b), _ptr(r)))
        if l == 0 and self.visible_kind == L.VIS_NEGBIN:
            return NegativeBinomialBernoulliRBM(W, a, b, r)
        return BernoulliRBM(W, a, b)

    def to_host(self):
        rbms = [self.get_layer(l) for l in range(self.nrbms)]
        return rbms if self.nrbms > 1 else rbms[0]

    def copy_layer_from(self, l, src, src_l):
        check(lib().scdbm_model_copy_layer(self.h, l, src.h, src_l))

    def init_layer(self, l, x, inversedispersion=None, seed=0):
        r = _f64(inversedispersion) if inversedispersion is not None else None
        check(lib().scdbm_model_init_layer(self.h, l, x.h, _ptr(r), C.c_uint64(seed)))

    def close(self):
        if self.h:
            lib().scdbm_model_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _model(bm, ctx=None):
    return bm if isinstance(bm, DeviceModel) else DeviceModel.from_host(bm, ctx)


# ------------------------------------------------------------------- Mat
class Mat:
    """Device state matrix (samples x units)."""

    def __init__(self, ctx, rows, cols, kind, handle=None, owns=True):
        self.ctx, self.rows, self.cols, self.kind, self.owns = ctx, int(rows), int(cols), kind, owns
        if handle is None:
            self.h = C.c_void_p()
            check(lib().scdbm_mat_create(ctx.h, self.rows, self.cols, kind, C.byref(self.h)))
        else:
            self.h = handle

    @classmethod
    def shared(cls, ctx, handle, owner=None):
        """Wraps a shared handle returned by `scdbm_particles_layer` / `scdbm_trainer_*`: the library counts
        it as one more reference on the matrix, released by `close()`; `owner` is kept alive as well."""
        r, c, k = C.c_int64(), C.c_int64(), C.c_int32()
        check(lib().scdbm_mat_shape(handle, C.byref(r), C.byref(c), C.byref(k)))
        m = cls(ctx, r.value, c.value, k.value, handle, owns=True)
        m._keepalive = owner
        return m

    @classmethod
    def from_host(cls, ctx, x, kind):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("expected a matrix (samples x units)")
        m = cls(ctx, x.shape[0], x.shape[1], kind)
        m.upload(x)
        return m

    @classmethod
    def from_compressed(cls, ctx, shape, ptr, idx, val, kind=L.MAT_COUNT, major="csr"):
        """Device matrix from CSR (`ptr` over rows, `idx` = columns) or CSC (`major="csc"`: `ptr` over columns,
        `idx` = rows -- the 0-based arrays of a Julia `SparseMatrixCSC`) arrays; only the non-zeros cross PCIe."""
        m = cls(ctx, shape[0], shape[1], kind)
        m.upload_compressed(ptr, idx, val, major)
        return m

    @classmethod
    def from_sparse(cls, ctx, x, kind=L.MAT_COUNT):
        """`x`: any scipy.sparse matrix (samples x units)."""
        if x.format == "csc":
            return cls.from_compressed(ctx, x.shape, x.indptr, x.indices, x.data, kind, "csc")
        x = x.tocsr()
        return cls.from_compressed(ctx, x.shape, x.indptr, x.indices, x.data, kind, "csr")

    def upload_compressed(self, ptr, idx, val, major="csr"):
        if major not in ("csr", "csc"):
            raise ValueError("major must be 'csr' or 'csc'")
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float32)
        nmajor = self.cols if major == "csc" else self.rows
        if ptr.shape != (nmajor + 1,) or idx.shape != val.shape or idx.ndim != 1 or (len(ptr) and ptr[-1] != len(idx)):
            raise ValueError("inconsistent compressed arrays")
        check(lib().scdbm_mat_upload_compressed(self.h, 1 if major == "csc" else 0, ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                idx.ctypes.data_as(C.POINTER(C.c_int32)), val.ctypes.data_as(C.POINTER(C.c_float))))

    def to_host_csr(self, index_dtype=None, buffers=None):
        """(indptr int64[rows+1], indices, values uint16) of a count / binary matrix, columns ascending within a
        row: the compact wire format for generated cells.  `index_dtype`: np.uint16 (default when it fits) or np.int32.
        `buffers=(indices_buf, values_buf)`: preallocated (pinned) 1-d arrays of sufficient capacity to fill instead
        of fresh pageable arrays; views of their first nnz entries are returned."""
        if index_dtype is None:
            index_dtype = np.uint16 if self.cols <= 65536 else np.int32
        index_dtype = np.dtype(index_dtype)
        if index_dtype not in (np.dtype(np.uint16), np.dtype(np.int32)):
            raise ValueError("index_dtype must be uint16 or int32")
        nnz = C.c_int64()
        check(lib().scdbm_mat_csr_prepare(self.h, C.byref(nnz)))
        indptr = np.zeros(self.rows + 1, dtype=np.int64)
        if buffers is None:
            indices = np.zeros(nnz.value, dtype=index_dtype)
            values = np.zeros(nnz.value, dtype=np.uint16)
        else:
            ib, vb = buffers
            if ib.dtype != index_dtype or vb.dtype != np.uint16 or ib.ndim != 1 or vb.ndim != 1 or min(len(ib), len(vb)) < nnz.value:
                raise ValueError(f"buffers too small or of the wrong dtype for {nnz.value} non-zeros")
            indices, values = ib[:nnz.value], vb[:nnz.value]
        check(lib().scdbm_mat_download_csr(self.h, indptr.ctypes.data_as(C.POINTER(C.c_int64)), indices.ctypes.data_as(C.c_void_p),
                                           index_dtype.itemsize, values.ctypes.data_as(C.POINTER(C.c_uint16))))
        return indptr, indices, values

    def generate_counts(self, gene_mean, r, sizefactor_sigma=0.3, seed=0, row_offset=0):
        """Fills a COUNT matrix with synthetic "10x-like" cells on the device: no host matrix exists."""
        gm, rr = _f64(gene_mean), _f64(r)
        if gm.shape != (self.cols,) or rr.shape != (self.cols,):
            raise ValueError("gene_mean and r must have one entry per column")
        check(lib().scdbm_mat_generate_counts(self.h, _ptr(gm), _ptr(rr), float(sizefactor_sigma), int(seed), int(row_offset)))
        return self

    def upload(self, x):
        x = _f64(x)
        if x.shape != (self.rows, self.cols):
            raise ValueError(f"shape {x.shape} does not match ({self.rows}, {self.cols})")
        check(lib().scdbm_mat_upload(self.h, _ptr(x), max(self.rows, 1)))

    def to_host(self):
        out = np.zeros((self.rows, self.cols), order="F")
        check(lib().scdbm_mat_download(self.h, _ptr(out), max(self.rows, 1)))
        return out

    def to_host_u16(self, out=None, wait=True):
        """Compact row-major uint16 counts.  `wait=False` (with a pinned `out`) snapshots the matrix in stream
        order and lets the copy engine drain it while later work runs; `Context.wait_downloads()` completes it."""
        if out is None:
            out = np.zeros((self.rows, self.cols), dtype=np.uint16)
        if out.shape != (self.rows, self.cols) or out.dtype != np.uint16 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint16 array of the matrix' shape")
        f = lib().scdbm_mat_download_u16 if wait else lib().scdbm_mat_download_u16_async
        check(f(self.h, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def column_stats(self):
        mean, var, zf = np.zeros(self.cols), np.zeros(self.cols), np.zeros(self.cols)
        check(lib().scdbm_mat_column_stats(self.h, _ptr(mean), _ptr(var), _ptr(zf)))
        return mean, var, zf

    def close(self):
        if self.owns and self.h:
            lib().scdbm_mat_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _data_kind(model, layer=0):
    if layer == 0:
        return L.MAT_COUNT if model.visible_kind == L.VIS_NEGBIN else L.MAT_BINARY
    return L.MAT_BINARY


def _is_sparse(x):
    return hasattr(x, "tocsr") and hasattr(x, "format")


def _as_mat(ctx, x, kind):
    if isinstance(x, Mat):
        return x
    return Mat.from_sparse(ctx, x, kind) if _is_sparse(x) else Mat.from_host(ctx, x, kind)


def _state_kind(x, default):
    """binary data may be passed for a REAL slot; pick the cheapest exact kind"""
    if isinstance(x, Mat):
        return x.kind
    x = np.asarray(x)
    if default == L.MAT_COUNT:
        return L.MAT_COUNT
    return L.MAT_BINARY if np.isin(x, (0.0, 1.0)).all() else L.MAT_REAL


# ---------------------------------------------- conditional operators (L1)
def _conditional(fn, bm, l, src, out_cols, what, factor, tick, u, src_kind, out_kind, ctx):
    m = _model(bm, ctx)
    s = _as_mat(m.ctx, src, src_kind)
    out = Mat(m.ctx, s.rows, out_cols, out_kind)
    uu = _f64(u) if u is not None else None
    if uu is not None and uu.shape != (s.rows, out_cols):
        raise ValueError("uniforms must have the shape of the output")
    check(fn(m.h, l, s.h, out.h, float(factor), what, int(tick) & 0xFFFFFFFF, _ptr(uu), max(s.rows, 1)))
    return out


def _hidden(bm, v, what, factor=1.0, layer=0, tick=0, u=None, device=False, ctx=None):
    m = _model(bm, ctx)
    kind = _state_kind(v, _data_kind(m, layer))
    okind = L.MAT_BINARY if what == L.OUT_SAMPLE else L.MAT_REAL
    out = _conditional(lib().scdbm_hidden, m, layer, v, m.units[layer + 1], what, factor, tick, u, kind, okind, ctx)
    return out if device else out.to_host()


def _visible(bm, h, what, factor=1.0, layer=0, tick=0, u=None, device=False, ctx=None):
    m = _model(bm, ctx)
    kind = _state_kind(h, L.MAT_BINARY)
    nb = layer == 0 and m.visible_kind == L.VIS_NEGBIN
    okind = (L.MAT_COUNT if nb else L.MAT_BINARY) if what == L.OUT_SAMPLE else L.MAT_REAL
    out = _conditional(lib().scdbm_visible, m, layer, h, m.units[layer], what, factor, tick, u, kind, okind, ctx)
    return out if device else out.to_host()


def hiddeninput(rbm, v, **kw):
    """src/gibbssampling.jl:189-276"""
    return _hidden(rbm, v, L.OUT_INPUT, **kw)


def hiddenpotential(rbm, v, factor=1.0, **kw):
    """src/gibbssampling.jl:314-335"""
    return _hidden(rbm, v, L.OUT_POTENTIAL, factor, **kw)


def samplehidden(rbm, v, factor=1.0, **kw):
    """src/gibbssampling.jl:575-610"""
    return _hidden(rbm, v, L.OUT_SAMPLE, factor, **kw)


def visibleinput(rbm, h, **kw):
    """src/gibbssampling.jl:860-890"""
    return _visible(rbm, h, L.OUT_INPUT, **kw)


def visiblepotential(rbm, h, factor=1.0, **kw):
    """src/gibbssampling.jl:930-1020"""
    return _visible(rbm, h, L.OUT_POTENTIAL, factor, **kw)


def samplevisible(rbm, h, factor=1.0, **kw):
    """src/gibbssampling.jl:733-756"""
    return _visible(rbm, h, L.OUT_SAMPLE, factor, **kw)


def dbmlayer(dbm, layer, below, above, what=L.OUT_POTENTIAL, tick=0, u=None, device=False, ctx=None):
    """Intermediate DBM layer from both neighbours (src/gibbssampling.jl:165-170)."""
    m = _model(dbm, ctx)
    lo = _as_mat(m.ctx, below, _state_kind(below, _data_kind(m, layer - 1)))
    hi = _as_mat(m.ctx, above, _state_kind(above, L.MAT_BINARY))
    out = Mat(m.ctx, lo.rows, m.units[layer], L.MAT_BINARY if what == L.OUT_SAMPLE else L.MAT_REAL)
    uu = _f64(u) if u is not None else None
    check(lib().scdbm_dbm_layer(m.h, layer, lo.h, hi.h, out.h, what, int(tick) & 0xFFFFFFFF, _ptr(uu), max(lo.rows, 1)))
    return out if device else out.to_host()


# ------------------------------------------------------------- particles
class DeviceParticles:
    """`Particles` (src/gibbssampling.jl:9) resident on the GPU."""

    def __init__(self, model, nparticles, row_offset=0, real_valued=False):
        self.model, self.n = model, int(nparticles)
        self.h = C.c_void_p()
        check(lib().scdbm_particles_create(model.h, self.n, C.c_uint64(row_offset), int(real_valued), C.byref(self.h)))

    def init(self, biased=False):
        check(lib().scdbm_particles_init(self.h, int(biased)))
        return self

    def layer(self, i):
        h = C.c_void_p()
        check(lib().scdbm_particles_layer(self.h, i, C.byref(h)))
        return Mat.shared(self.model.ctx, h, self)

    @property
    def clock(self):
        c = C.c_int64()
        check(lib().scdbm_particles_clock(self.h, C.byref(c), -1))
        return c.value

    @clock.setter
    def clock(self, v):
        check(lib().scdbm_particles_clock(self.h, None, int(v)))

    def to_host(self):
        return [self.layer(i).to_host() for i in range(len(self.model.units))]

    def upload(self, mats):
        for i, x in enumerate(mats):
            self.layer(i).upload(x)
        return self

    def close(self):
        if self.h:
            lib().scdbm_particles_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def initparticles(bm, nparticles, biased=False, row_offset=0, device=False, ctx=None):
    """src/gibbssampling.jl:386-419"""
    m = _model(bm, ctx)
    p = DeviceParticles(m, nparticles, row_offset).init(biased)
    return p if device else p.to_host()


def gibbssample_(particles, bm, nsteps=5, upfactor=1.0, downfactor=1.0, temperature=1.0, ctx=None):
    """`gibbssample!` (src/gibbssampling.jl:62-70 RBM, :153-182 DBM).  Host
    particles (list of arrays) are updated in place."""
    m = _model(bm, ctx)
    if isinstance(particles, DeviceParticles):
        check(lib().scdbm_gibbs(m.h, particles.h, nsteps, temperature, upfactor, downfactor))
        return particles
    p = DeviceParticles(m, particles[0].shape[0]).upload(particles)
    p.clock = getattr(particles, "clock", 1)
    check(lib().scdbm_gibbs(m.h, p.h, nsteps, temperature, upfactor, downfactor))
    for i, x in enumerate(p.to_host()):
        particles[i] = x
    if hasattr(particles, "clock"):
        particles.clock = p.clock
    return particles


def gibbssamplenegativebinomial_(particles, rbm, nsteps=5, upfactor=1.0, downfactor=1.0, ctx=None):
    """`gibbssamplenegativebinomial!` RBM form (src/gibbssampling.jl:82-110);
    the DBM form (:112-151) equals `gibbssample_` for a DBM."""
    m = _model(rbm, ctx)
    if m.nrbms > 1:
        return gibbssample_(particles, m, nsteps)
    if isinstance(particles, DeviceParticles):
        check(lib().scdbm_gibbs_negbin_rbm(m.h, particles.h, nsteps, upfactor, downfactor))
        return particles
    p = DeviceParticles(m, particles[0].shape[0]).upload(particles)
    p.clock = getattr(particles, "clock", 1)
    check(lib().scdbm_gibbs_negbin_rbm(m.h, p.h, nsteps, upfactor, downfactor))
    for i, x in enumerate(p.to_host()):
        particles[i] = x
    if hasattr(particles, "clock"):
        particles.clock = p.clock
    return particles


def gibbssamplecond_(particles, bm, fixedvars, nsteps=5, ctx=None):
    """`gibbssamplecond!` / `gibbssamplenegativebinomial_cond!` / `gibbssamplecondrow!`
    (src/gibbssampling.jl:1042-1237).  `fixedvars`: indices of clamped visible units, a boolean vector over
    the visible units, or a boolean (nparticles x nvisible) matrix."""
    m = _model(bm, ctx)
    host = not isinstance(particles, DeviceParticles)
    p = DeviceParticles(m, particles[0].shape[0]).upload(particles) if host else particles
    if host:
        p.clock = getattr(particles, "clock", 1)
    V = m.units[0]
    fv = np.asarray(fixedvars)
    if fv.dtype != bool:
        mask = np.zeros(V, dtype=bool)
        mask[fv.astype(np.int64)] = True
    else:
        mask = fv
    if mask.ndim == 1:
        if mask.shape != (V,):
            raise ValueError("mask must have one entry per visible unit")
        mm = np.ascontiguousarray(mask, dtype=np.uint8)
        check(lib().scdbm_gibbs_cond(m.h, p.h, mm.ctypes.data_as(C.POINTER(C.c_uint8)), 0, 0, nsteps))
    else:
        if mask.shape != (p.n, V):
            raise ValueError("mask must have the shape of the visible particles")
        mm = np.asfortranarray(mask, dtype=np.uint8)
        check(lib().scdbm_gibbs_cond(m.h, p.h, mm.ctypes.data_as(C.POINTER(C.c_uint8)), p.n, p.n, nsteps))
    if host:
        for i, x in enumerate(p.to_host()):
            particles[i] = x
        if hasattr(particles, "clock"):
            particles.clock = p.clock
        return particles
    return p


gibbssamplenegativebinomial_cond_ = gibbssamplecond_
gibbssamplecondrow_ = gibbssamplecond_


def estimatedropout(x, ctx=None):
    """`estimatedropout!` (src/rbmtraining.jl:271-287): binomial GLM of the per-gene zero fraction on
    log1p(per-gene mean); the per-gene statistics come from the device, the 2-parameter IRLS runs on the
    host.  Returns (dropoutmid, dropoutshape) = (intercept, slope)."""
    ctx = ctx or default_context()
    xm = _as_mat(ctx, x, L.MAT_COUNT)
    mean, _, zf = xm.column_stats()
    lc = np.log1p(mean)
    X = np.stack([np.ones_like(lc), lc], axis=1)
    beta = np.zeros(2)
    for _ in range(100):
        eta = X @ beta
        mu = 1.0 / (1.0 + np.exp(-eta))
        w = np.maximum(mu * (1 - mu), 1e-12)
        z = eta + (zf - mu) / w
        new = np.linalg.solve(X.T @ (w[:, None] * X), X.T @ (w * z))
        done = np.max(np.abs(new - beta)) < 1e-12
        beta = new
        if done:
            break
    return float(beta[0]), float(beta[1])


def sampledropout_(v, x, dropoutmid, dropoutshape, log1p_form=False, tick=0, row_offset=0, ctx=None):
    """Applies `sampledropout` (src/rbmtraining.jl:289-311) to the visible states in place:
    v .*= 1 - bernoulli(logistic(shape (x' - mid))) (src/gibbssampling.jl:103-107,144-148)."""
    ctx = ctx or (v.ctx if isinstance(v, Mat) else default_context())
    vm = _as_mat(ctx, v, L.MAT_COUNT)
    xm = _as_mat(ctx, x, L.MAT_COUNT)
    mid = np.atleast_1d(np.asarray(dropoutmid, dtype=np.float64))
    shp = np.atleast_1d(np.asarray(dropoutshape, dtype=np.float64))
    if mid.shape != shp.shape:
        raise ValueError("dropoutmid and dropoutshape must have the same length")
    check(lib().scdbm_sample_dropout(vm.h, xm.h, _ptr(mid), _ptr(shp), len(mid), int(log1p_form), int(tick) & 0xFFFFFFFF,
                                     C.c_uint64(row_offset)))
    return vm if isinstance(v, Mat) else vm.to_host()


def mle_for_theta(x, mu, lam=100.0, maxiter=20, convtol=1e-6, ctx=None):
    """`mle_for_θ` for every gene (src/rbmtraining.jl:675-719)."""
    ctx = ctx or (x.ctx if isinstance(x, Mat) else default_context())
    xm = _as_mat(ctx, x, L.MAT_COUNT)
    mm = mu if isinstance(mu, Mat) else Mat.from_host(ctx, mu, L.MAT_REAL)
    theta = np.zeros(xm.cols)
    check(lib().scdbm_mle_theta(xm.h, mm.h, float(lam), int(maxiter), float(convtol), _ptr(theta)))
    return theta


def sampleparticles(bm, nparticles, burnin=10, row_offset=0, device=False, ctx=None):
    """src/gibbssampling.jl:646-650"""
    m = _model(bm, ctx)
    p = DeviceParticles(m, nparticles, row_offset).init(False)
    check(lib().scdbm_gibbs(m.h, p.h, burnin, 1.0, 1.0, 1.0))
    return p if device else p.to_host()


def samples(bm, nsamples, burnin=50, samplelast=True, row_offset=0, device=False, dtype=np.float64, ctx=None):
    """src/gibbssampling.jl:663-689.  `dtype=np.uint16` returns the compact
    row-major count matrix instead of Float64."""
    m = _model(bm, ctx)
    if samplelast:
        p = sampleparticles(m, nsamples, burnin, row_offset, device=True)
        out = p.layer(0)
    else:
        p = sampleparticles(m, nsamples, burnin - 1, row_offset, device=True)
        out = Mat(m.ctx, nsamples, m.units[0], L.MAT_REAL)
        check(lib().scdbm_particles_visible_potential(m.h, p.h, out.h))
    if device:
        return out
    return out.to_host_u16() if dtype == np.uint16 else out.to_host()


# ------------------------------------------------------------- mean-field
def meanfield(dbm, x, eps=0.001, maxiter=0, device=False, return_info=False, ctx=None):
    """src/dbmtraining.jl:175-228"""
    m = _model(dbm, ctx)
    xm = _as_mat(m.ctx, x, _data_kind(m))
    mu = DeviceParticles(m, xm.rows, real_valued=True)
    it, delta = C.c_int32(), C.c_double()
    check(lib().scdbm_meanfield(m.h, xm.h, eps, maxiter, mu.h, C.byref(it), C.byref(delta)))
    mu._x = xm
    res = mu if device else mu.to_host()
    return (res, it.value, delta.value) if return_info else res


# ----------------------------------------------------------- optimizers
class LoglikelihoodOptimizer:
    """`LoglikelihoodOptimizer` / `StackedOptimizer` (src/optimizers.jl:34-39,
    57-70,103): holds the gradient on the device; implements the documented
    plug-in protocol `initialized`, `computegradient_`, `updateparameters_`
    (src/optimizers.jl:1-18)."""

    def __init__(self, learningrate=0.005, sdlearningrate=0.0):
        self.learningrate = learningrate
        self.sdlearningrate = sdlearningrate
        self.model = None
        self.h = C.c_void_p()

    def initialized(self, bm):
        o = LoglikelihoodOptimizer(self.learningrate, self.sdlearningrate)
        o.model = bm
        check(lib().scdbm_opt_create(bm.h, C.byref(o.h)))
        return o

    def computegradient_(self, v, vmodel, h, hmodel, rbm, layer=0, nrows_used=0):
        check(lib().scdbm_compute_gradient(self.h, rbm.h, layer, v.h, vmodel.h, h.h, hmodel.h, nrows_used))

    def computegradient_dbm_(self, mu, particles, dbm, pos_scale=0.0, neg_scale=0.0):
        check(lib().scdbm_compute_gradient_dbm(self.h, dbm.h, mu.h, particles.h, pos_scale, neg_scale))

    def updateparameters_(self, bm, layer=-1):
        check(lib().scdbm_update_parameters(bm.h, self.h, layer, self.learningrate))

    def gradient(self, layer=0):
        V, H = self.model.units[layer], self.model.units[layer + 1]
        W = np.zeros((V, H), order="F")
        a, b = np.zeros(V), np.zeros(H)
        check(lib().scdbm_opt_get_gradient(self.h, layer, _ptr(W), V, _ptr(a), _ptr(b)))
        return W, a, b

    def buffer(self):
        """(device pointer, number of floats) of the contiguous gradient block
        -- the data-parallel allreduce payload."""
        p, n = C.c_void_p(), C.c_int64()
        check(lib().scdbm_opt_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self):
        if self.h:
            lib().scdbm_opt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def loglikelihoodoptimizer(bm=None, learningrate=0.0, sdlearningrate=0.0):
    """src/optimizers.jl:57-70"""
    o = LoglikelihoodOptimizer(learningrate, sdlearningrate)
    return o.initialized(bm) if bm is not None else o


def computegradient_(optimizer, v, vmodel, h, hmodel, rbm):
    """src/optimizers.jl:207-214 with host matrices."""
    m = _model(rbm)
    if optimizer.model is not m:
        raise ValueError("optimizer was initialised for a different model")
    vk = _data_kind(m)
    mats = [_as_mat(m.ctx, v, _state_kind(v, vk)), _as_mat(m.ctx, vmodel, L.MAT_REAL if vk != L.MAT_COUNT else _state_kind(vmodel, L.MAT_REAL)),
            _as_mat(m.ctx, h, L.MAT_REAL), _as_mat(m.ctx, hmodel, L.MAT_REAL)]
    optimizer.computegradient_(*mats, m)


def updateparameters_(bm, optimizer):
    """src/optimizers.jl:450-454,478-485,505-510"""
    optimizer.updateparameters_(bm)


# ----------------------------------------------------------- samplers
@dataclass
class NoSampler:
    """src/samplers.jl:22-35"""


@dataclass
class GibbsSampler:
    """src/samplers.jl:4-6,47-62"""
    nsteps: int = 1


# ----------------------------------------------------------- RBM training
class RBMTrainer:
    """State `fitrbm` keeps across epochs (src/rbmtraining.jl:137-151)."""

    def __init__(self, model, batchsize, pcd=True, layer=0):
        self.model, self.batchsize, self.pcd, self.layer = model, batchsize, pcd, layer
        self.h = C.c_void_p()
        check(lib().scdbm_trainer_create(model.h, layer, batchsize, int(pcd), C.byref(self.h)))

    @property
    def clock(self):
        c = C.c_int64()
        check(lib().scdbm_trainer_clock(self.h, C.byref(c), -1))
        return c.value

    @clock.setter
    def clock(self, v):
        check(lib().scdbm_trainer_clock(self.h, None, int(v)))

    def chainstate(self):
        h = C.c_void_p()
        check(lib().scdbm_trainer_chainstate(self.h, C.byref(h)))
        return Mat.shared(self.model.ctx, h, self)

    def track_mean(self, enable=True):
        check(lib().scdbm_trainer_track_mean(self.h, int(enable)))

    def estimated_mean(self):
        h = C.c_void_p()
        check(lib().scdbm_trainer_estimated_mean(self.h, C.byref(h)))
        return Mat.shared(self.model.ctx, h, self)

    def close(self):
        if self.h:
            lib().scdbm_trainer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def philox_permutation(n, seed, tick):
    """`randperm(n)` stand-in shared with oracle.philox.Rng.permutation
    (host-side integer work; the library receives the permutation)."""
    from . import hostrng
    return hostrng.permutation(n, seed, tick)


def trainrbm_(model, x, trainer, learningrate=0.005, cdsteps=1, upfactor=1.0, downfactor=1.0, perm=None,
              sampler=None):
    """`trainrbm!` one epoch (src/rbmtraining.jl:498-591)."""
    if sampler is not None:
        cdsteps = 1 if isinstance(sampler, NoSampler) else sampler.nsteps + 1
    n = x.rows
    if trainer.batchsize > n:
        raise ValueError(f"Batchsize ({trainer.batchsize}) exceeds number of samples ({n}).")
    if perm is None:
        perm = philox_permutation(n, int(model.ctx.get("seed")), trainer.clock)
        trainer.clock = trainer.clock + 1
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    check(lib().scdbm_rbm_train_epoch(trainer.h, x.h, perm.ctypes.data_as(C.POINTER(C.c_int32)), learningrate, cdsteps,
                                      upfactor, downfactor))
    return model


def fitrbm(x, nhidden=None, epochs=10, upfactor=1.0, downfactor=1.0, learningrate=0.005, learningrates=None,
           pcd=True, cdsteps=1, batchsize=1, rbmtype=BernoulliRBM, inversedispersion=None, startrbm=None,
           monitoring=None, fisherscoring=0, zeroinflation=False, estimatedispersion="gene", lam=100.0,
           sampler=None, seed=None, device=False, ctx=None):
    """`fitrbm` (src/rbmtraining.jl:89-223).  For NB RBMs `fisherscoring != 0` re-estimates the per-gene
    inverse dispersion every 5th epoch by penalised Fisher scoring on the device (`mle_for_θ`, :172-183,
    estimatedispersion = "gene"; the reference's other two branches are broken -- SURVEY B4 -- and raise
    here), and `zeroinflation` fits the dropout logistic (`estimatedropout!`, :531-534).  Note that the
    reference's own defaults are fisherscoring = 1, zeroinflation = true; they are off by default here
    because the reference applies `mle_for_θ` to every RBM type and errors for non-NB ones."""
    ctx = ctx or default_context()
    if seed is not None:
        ctx.set("seed", seed)
    nb = rbmtype is NegativeBinomialBernoulliRBM or rbmtype == "nb"
    if fisherscoring != 0 and estimatedispersion != "gene":
        raise NotImplementedError('estimatedispersion must be "gene" (the other branches of the reference do not run)')
    if (fisherscoring != 0 or zeroinflation) and not nb:
        raise NotImplementedError("fisherscoring / zeroinflation apply to NegativeBinomialBernoulliRBM only")
    xm = _as_mat(ctx, x, L.MAT_COUNT if nb else L.MAT_REAL)
    nhidden = xm.cols if nhidden is None else nhidden
    if learningrates is None:
        learningrates = [learningrate] * epochs
    if len(learningrates) < epochs:
        raise ValueError("Not enough `learningrates` (vector of length %d) for training epochs (%d)" % (len(learningrates), epochs))
    if batchsize > xm.rows:
        raise ValueError(f"Batchsize ({batchsize}) exceeds number of samples ({xm.rows}).")
    if startrbm is not None:
        m = startrbm if isinstance(startrbm, DeviceModel) else DeviceModel.from_host(startrbm, ctx)
    else:
        m = DeviceModel(ctx, [xm.cols, nhidden], L.VIS_NEGBIN if nb else L.VIS_BERNOULLI)
        m.init_layer(0, xm, inversedispersion, seed=int(ctx.get("seed")))
    trainer = RBMTrainer(m, batchsize, pcd)
    if fisherscoring != 0:
        trainer.track_mean(True)
    for epoch in range(1, epochs + 1):
        trainrbm_(m, xm, trainer, learningrates[epoch - 1], cdsteps, upfactor, downfactor, sampler=sampler)
        if monitoring is not None:
            monitoring(m.to_host(), epoch)
        if fisherscoring != 0 and epoch % 5 == 0:
            theta = mle_for_theta(xm, trainer.estimated_mean(), lam, maxiter=20)
            check(lib().scdbm_model_set_layer(m.h, 0, None, 0, None, None, _ptr(theta)))
    trainer.close()
    if device:
        return m
    rbm = m.to_host()
    if zeroinflation:
        mid, shape = estimatedropout(xm)
        rbm.dropoutmid, rbm.dropoutshape = np.array([mid]), np.array([shape])
    return rbm


@dataclass
class TrainLayer:
    """The hot-path subset of `TrainLayer` (src/rbmstacking.jl:9-94)."""
    nhidden: int = -1
    rbmtype: type = BernoulliRBM
    epochs: int = -1
    learningrate: float = -np.inf
    batchsize: int = -1
    pcd: bool = True
    cdsteps: int = 1
    inversedispersion: np.ndarray = None
    startrbm: object = None


def stackrbms(x, nhiddens=None, epochs=10, predbm=False, learningrate=0.005, batchsize=1, trainlayers=None,
              seed=None, device=False, ctx=None):
    """`stackrbms` (src/rbmstacking.jl:228-284; factors :247-275)."""
    ctx = ctx or default_context()
    if seed is not None:
        ctx.set("seed", seed)
    base_seed = int(ctx.get("seed"))
    if not trainlayers:
        nb0 = False
        if nhiddens is None or len(nhiddens) == 0:
            ncol = x.cols if isinstance(x, Mat) else np.asarray(x).shape[1]
            nhiddens = [ncol] * 2
        trainlayers = [TrainLayer(nhidden=n) for n in nhiddens]
    nb0 = trainlayers[0].rbmtype is NegativeBinomialBernoulliRBM
    xm = _as_mat(ctx, x, L.MAT_COUNT if nb0 else L.MAT_REAL)
    layers = []
    upfactor, downfactor = (2.0 if predbm else 1.0), 1.0
    hiddenval = xm
    for i, tl in enumerate(trainlayers):
        if i > 0:
            hiddenval = _conditional(lib().scdbm_hidden, layers[i - 1], 0, hiddenval, layers[i - 1].units[1],
                                     L.OUT_POTENTIAL, upfactor, 0, None, None, L.MAT_REAL, ctx)
            if predbm:
                upfactor = downfactor = 2.0
                if i == len(trainlayers) - 1:
                    upfactor = 1.0
            else:
                upfactor = downfactor = 1.0
        ctx.set("seed", (base_seed + 1000003 * (i + 1)) % (1 << 52))
        layers.append(fitrbm(hiddenval, nhidden=tl.nhidden, epochs=tl.epochs if tl.epochs >= 0 else epochs,
                             upfactor=upfactor, downfactor=downfactor,
                             learningrate=tl.learningrate if tl.learningrate >= 0 else learningrate,
                             pcd=tl.pcd, cdsteps=tl.cdsteps, batchsize=tl.batchsize if tl.batchsize > 0 else batchsize,
                             rbmtype=tl.rbmtype, inversedispersion=tl.inversedispersion, startrbm=tl.startrbm,
                             device=True, ctx=ctx))
    ctx.set("seed", base_seed)
    units = [layers[0].units[0]] + [m.units[1] for m in layers]
    dbm = DeviceModel(ctx, units, layers[0].visible_kind)
    for i, m in enumerate(layers):
        dbm.copy_layer_from(i, m, 0)
    return dbm if device else dbm.to_host()


def defaultfinetuninglearningrates(learningrate, epochs):
    """src/dbmtraining.jl:59-61"""
    return learningrate * 11.0 / (10.0 + np.arange(1, epochs + 1))


def traindbm_(dbm, x, epochs=10, nparticles=100, learningrate=0.005, learningrates=None, monitoring=None,
              particles=None, nsteps=5, eps=0.001, global_nsamples=None, global_nparticles=None, device=False, ctx=None):
    """`traindbm!` (src/dbmtraining.jl:250-297).  When the model's context has a communicator (`Context.comm_init`),
    `x` and `particles` are this rank's shards and the step is data-parallel: partial sums scaled by the global counts,
    gradient block summed by NCCL inside the library, global mean-field stopping rule, identical update on every rank.
    `global_nsamples` / `global_nparticles` default to the sums of the shard sizes over the ranks."""
    m = _model(dbm, ctx)
    xm = _as_mat(m.ctx, x, _data_kind(m))
    if learningrates is None:
        learningrates = defaultfinetuninglearningrates(learningrate, epochs)
    if len(learningrates) < epochs:
        raise ValueError("Not enough `learningrates` for training epochs")
    p = particles if particles is not None else DeviceParticles(m, nparticles).init(False)
    mu = DeviceParticles(m, xm.rows, real_valued=True)
    opt = LoglikelihoodOptimizer().initialized(m)
    dp = m.ctx.comm_size > 1
    if dp and (global_nsamples is None or global_nparticles is None):
        tot = m.ctx.allreduce([xm.rows, p.n], "sum")          # once, outside the epoch loop
        global_nsamples = int(tot[0]) if global_nsamples is None else global_nsamples
        global_nparticles = int(tot[1]) if global_nparticles is None else global_nparticles
    for epoch in range(epochs):
        if dp:
            check(lib().scdbm_dbm_pcd_step_dp(m.h, xm.h, p.h, mu.h, opt.h, float(learningrates[epoch]), nsteps, eps,
                                              int(global_nsamples), int(global_nparticles), None))
        else:
            check(lib().scdbm_dbm_pcd_step(m.h, xm.h, p.h, mu.h, opt.h, float(learningrates[epoch]), nsteps, eps, None))
        if monitoring is not None:
            monitoring(m.to_host(), epoch + 1)
    opt.close()
    mu.close()
    return m if device else m.to_host()


def fitdbm(x, nhiddens=None, epochs=10, nparticles=100, learningrate=0.005, learningrates=None,
           learningratepretraining=None, epochspretraining=None, batchsizepretraining=1, pretraining=None,
           monitoring=None, seed=None, device=False, ctx=None):
    """`fitdbm` (src/dbmtraining.jl:108-149)."""
    ctx = ctx or default_context()
    if seed is not None:
        ctx.set("seed", seed)
    nb0 = bool(pretraining) and pretraining[0].rbmtype is NegativeBinomialBernoulliRBM
    xm = _as_mat(ctx, x, L.MAT_COUNT if nb0 else L.MAT_REAL)
    if not pretraining and not nhiddens:
        nhiddens = [xm.cols] * 2
    dbm = stackrbms(xm, nhiddens=nhiddens, epochs=epochs if epochspretraining is None else epochspretraining,
                    predbm=True, batchsize=batchsizepretraining,
                    learningrate=learningrate if learningratepretraining is None else learningratepretraining,
                    trainlayers=pretraining, device=True, ctx=ctx)
    xd = xm
    if not nb0 and xm.kind == L.MAT_REAL:
        xd = xm  # Bernoulli-visible DBMs train on the same matrix
    return traindbm_(dbm, xd, epochs=epochs, nparticles=nparticles, learningrate=learningrate,
                     learningrates=learningrates, monitoring=monitoring, device=device, ctx=ctx)


# ------------------------------------------------------------------ AIS
def aislogimpweights(bm, ntemperatures=100, temperatures=None, nparticles=100, burnin=5, row_offset=0, ctx=None):
    """src/evaluating.jl:18-46 (RBM), :164-206 (MultimodalDBM)."""
    m = _model(bm, ctx)
    if temperatures is None:
        temperatures = np.arange(ntemperatures + 1) / ntemperatures
    t = np.ascontiguousarray(temperatures, dtype=np.float64)
    out = np.zeros(nparticles)
    check(lib().scdbm_ais_run(m.h, _ptr(t), len(t), nparticles, burnin, C.c_uint64(row_offset), _ptr(out)))
    return out


def logmeanexp(v):
    """src/evaluating.jl:935-943"""
    v = np.asarray(v, float)
    vmax = v.max()
    return vmax + np.log(np.sum(np.exp(v - vmax))) - np.log(len(v))


def logpartitionfunctionzeroweights(bm):
    """src/evaluating.jl:1001-1031,1055-1065 (host: O(units) scalar work)."""
    rbms = _rbms(bm.to_host() if isinstance(bm, DeviceModel) else bm)
    first = rbms[0]
    if first.kind == L.VIS_NEGBIN:
        vis = np.sum(-first.inversedispersion * np.log(1.0 - np.exp(first.visbias)))
    else:
        vis = np.sum(np.logaddexp(0.0, first.visbias))
    if len(rbms) == 1:
        return vis + np.sum(np.logaddexp(0.0, first.hidbias))
    tot = vis
    for i in range(1, len(rbms)):
        tot += np.sum(np.logaddexp(0.0, rbms[i].visbias + rbms[i - 1].hidbias))
    return tot + np.sum(np.logaddexp(0.0, rbms[-1].hidbias))


def logpartitionfunction(bm, logr=None, **kw):
    """src/evaluating.jl:967-992"""
    if logr is None:
        logr = logmeanexp(aislogimpweights(bm, **kw))
    return logr + logpartitionfunctionzeroweights(bm)


def mostevenbatches(ntasks, nbatches):
    """src/misc.jl:157-169 -- the model for sharding chains over GPUs."""
    minparticles, nbigger = divmod(ntasks, nbatches)
    return [minparticles + 1] * nbigger + [minparticles] * (nbatches - nbigger)
