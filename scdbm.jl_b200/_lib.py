"""ctypes binding of libscdbm_b200.so (include/scdbm_b200.h).  The library is
the product; there is no Python/CPU fallback -- a missing library or a missing
GPU raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCDBM_B200_LIB: another build of the same library (kernel-tuning experiments under profiles/)
LIB_PATH = os.environ.get("SCDBM_B200_LIB") or os.path.join(_HERE, "libscdbm_b200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUP = 0, -1, -2, -3, -4
VIS_BERNOULLI, VIS_NEGBIN = 0, 1
MAT_BINARY, MAT_COUNT, MAT_REAL = 0, 1, 2
OUT_INPUT, OUT_POTENTIAL, OUT_SAMPLE = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2

P = C.c_void_p
PP = C.POINTER(C.c_void_p)
D = C.POINTER(C.c_double)
I32, I64, U32, U64, F64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double

# name -> argtypes; every function returns int32 unless listed in _RESTYPES
PROTOTYPES = {
    "scdbm_version": [],
    "scdbm_ctx_create": [I32, U64, P, PP],
    "scdbm_ctx_destroy": [P],
    "scdbm_sync": [P],
    "scdbm_ctx_set_option": [P, C.c_char_p, F64],
    "scdbm_ctx_get_option": [P, C.c_char_p, D],
    "scdbm_timer_start": [P],
    "scdbm_timer_stop": [P, C.POINTER(C.c_float)],
    "scdbm_launch_count": [P, C.POINTER(I64), I32],
    "scdbm_model_create": [P, I32, C.POINTER(I32), I32, PP],
    "scdbm_model_destroy": [P],
    "scdbm_model_set_layer": [P, I32, D, I64, D, D, D],
    "scdbm_model_get_layer": [P, I32, D, I64, D, D, D],
    "scdbm_model_copy_layer": [P, I32, P, I32],
    "scdbm_model_init_layer": [P, I32, P, D, U64],
    "scdbm_mat_create": [P, I64, I64, I32, PP],
    "scdbm_mat_destroy": [P],
    "scdbm_mat_shape": [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I32)],
    "scdbm_mat_upload": [P, D, I64],
    "scdbm_mat_download": [P, D, I64],
    "scdbm_mat_download_u16": [P, C.POINTER(C.c_uint16)],
    "scdbm_mat_download_u16_async": [P, C.POINTER(C.c_uint16)],
    "scdbm_wait_downloads": [P],
    "scdbm_mat_upload_compressed": [P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)],
    "scdbm_mat_csr_prepare": [P, C.POINTER(C.c_int64)],
    "scdbm_mat_download_csr": [P, C.POINTER(C.c_int64), C.c_void_p, C.c_int32, C.POINTER(C.c_uint16)],
    "scdbm_mat_generate_counts": [P, D, D, C.c_double, C.c_uint64, C.c_uint64],
    "scdbm_mat_copy": [P, P],
    "scdbm_mat_column_stats": [P, D, D, D],
    "scdbm_hidden": [P, I32, P, P, F64, I32, U32, D, I64],
    "scdbm_visible": [P, I32, P, P, F64, I32, U32, D, I64],
    "scdbm_dbm_layer": [P, I32, P, P, P, I32, U32, D, I64],
    "scdbm_particles_create": [P, I64, U64, I32, PP],
    "scdbm_particles_destroy": [P],
    "scdbm_particles_init": [P, I32],
    "scdbm_particles_layer": [P, I32, PP],
    "scdbm_particles_clock": [P, C.POINTER(I64), I64],
    "scdbm_gibbs": [P, P, I32, F64, F64, F64],
    "scdbm_gibbs_negbin_rbm": [P, P, I32, F64, F64],
    "scdbm_particles_visible_potential": [P, P, P],
    "scdbm_meanfield": [P, P, F64, I32, P, C.POINTER(I32), D],
    "scdbm_opt_create": [P, PP],
    "scdbm_opt_destroy": [P],
    "scdbm_compute_gradient": [P, P, I32, P, P, P, P, I64],
    "scdbm_compute_gradient_dbm": [P, P, P, P, F64, F64],
    "scdbm_update_parameters": [P, P, I32, F64],
    "scdbm_opt_get_gradient": [P, I32, D, I64, D, D],
    "scdbm_opt_buffer": [P, PP, C.POINTER(I64)],
    "scdbm_trainer_create": [P, I32, I32, I32, PP],
    "scdbm_trainer_destroy": [P],
    "scdbm_trainer_clock": [P, C.POINTER(I64), I64],
    "scdbm_trainer_chainstate": [P, PP],
    "scdbm_rbm_train_epoch": [P, P, C.POINTER(I32), F64, I32, F64, F64],
    "scdbm_dbm_pcd_step": [P, P, P, P, P, F64, I32, F64, C.POINTER(I32)],
    "scdbm_ais_run": [P, D, I32, I64, I32, U64, D],
    "scdbm_comm_unique_id": [C.POINTER(C.c_uint8)],
    "scdbm_comm_init": [P, C.POINTER(C.c_uint8), I32, I32],
    "scdbm_comm_destroy": [P],
    "scdbm_allreduce_gradient": [P],
    "scdbm_comm_allreduce_f64": [P, D, I32, I32],
    "scdbm_dbm_pcd_step_dp": [P, P, P, P, P, F64, I32, F64, I64, I64, C.POINTER(I32)],
    "scdbm_gibbs_cond": [P, P, C.POINTER(C.c_uint8), I64, I64, I32],
    "scdbm_sample_dropout": [P, P, D, D, I32, I32, U32, U64],
    "scdbm_trainer_track_mean": [P, I32],
    "scdbm_trainer_estimated_mean": [P, PP],
    "scdbm_mle_theta": [P, P, F64, I32, F64, D],
}


class ScdbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libscdbm_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the CUDA library is the product; there is no fallback)")
        L = C.CDLL(LIB_PATH)
        L.scdbm_last_error.restype = C.c_char_p
        L.scdbm_last_error.argtypes = []
        for name, args in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = I32
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise ScdbmError(status, lib().scdbm_last_error().decode("utf-8", "replace"))
