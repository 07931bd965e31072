"""ctypes binding of include/protograd_b200.h — the only way the host side reaches the GPU.

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
compute call on a box without a B200 raises `PGError`.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB_PATH") or os.path.join(HERE, "libprotograd_b200.so")   # PG_LIB_PATH: a debug build of the same library

PG_F32, PG_F64, PG_BF16, PG_F16, PG_I32 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU = 0, 1
ADD, SUB = 0, 1
X_PLUS_C, X_MINUS_C, C_MINUS_X, X_TIMES_C, X_OVER_C, C_OVER_X = range(6)
SCALAR_F64, SCALAR2_F64, MATH_FAST = 1, 2, 4


class PGError(RuntimeError):
    pass


class MlpDesc(ctypes.Structure):
    """struct pg_mlp_desc of include/protograd_b200.h"""
    _fields_ = [("n_layers", c_int32), ("opt", c_int32), ("batch", c_int64), ("dims", c_int64 * 5), ("x", c_void_p), ("labels", c_void_p),
                ("W", c_void_p * 4), ("b", c_void_p * 4), ("gW", c_void_p * 4), ("gb", c_void_p * 4), ("loss_out", c_void_p),
                ("loss_scale", c_double), ("arena_p", c_void_p), ("arena_g", c_void_p), ("arena_m", c_void_p), ("arena_v", c_void_p),
                ("arena_m_hat", c_void_p), ("arena_v_hat", c_void_p), ("arena_n", c_int64), ("stepsize", c_double), ("beta1", c_double),
                ("beta2", c_double), ("eps", c_double), ("it", c_int64), ("flags", c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or python protograd.jl_b200/build.py). There is no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)
lib.pg_last_error.restype = c_char_p

_P, _I, _L, _D, _Z = c_void_p, c_int, c_int64, c_double, c_size_t
_SIGS = {
    "pg_version": [],
    "pg_device_count": [POINTER(c_int)],
    "pg_init": [_I, POINTER(_P)],
    "pg_destroy": [_P],
    "pg_set_stream": [_P, _P],
    "pg_get_stream": [_P, POINTER(_P)],
    "pg_sync": [_P],
    "pg_launch_count": [_P, POINTER(c_int64)],
    "pg_reserve_workspace": [_P, _Z],
    "pg_set_sm_limit": [_P, _I],
    "pg_malloc": [_P, _Z, POINTER(_P)],
    "pg_free": [_P, _P],
    "pg_memset0": [_P, _P, _Z],
    "pg_copy_d2d": [_P, _P, _P, _Z],
    "pg_upload": [_P, _P, _P, _Z],
    "pg_download": [_P, _P, _P, _Z],
    "pg_download_async": [_P, _P, _P, _Z],
    "pg_host_alloc": [_Z, POINTER(_P)],
    "pg_host_free": [_P],
    "pg_event_create": [_P, POINTER(_P)],
    "pg_event_destroy": [_P, _P],
    "pg_event_record": [_P, _P, _I],
    "pg_stream_wait_event": [_P, _I, _P],
    "pg_event_sync": [_P, _P],
    "pg_upload_on": [_P, _I, _P, _P, _Z],
    "pg_graph_begin": [_P],
    "pg_graph_end": [_P, POINTER(_P)],
    "pg_graph_launch": [_P, _P],
    "pg_graph_destroy": [_P, _P],
    "pg_vec_binary": [_P, _I, _I, _P, _P, _P, _L],
    "pg_vec_scalar": [_P, _I, _I, _P, _P, _D, _I, _L],
    "pg_vec_axpy": [_P, _I, _P, _P, _D, _P, _I, _I, _L],
    "pg_vec_fill": [_P, _I, _P, _D, _L],
    "pg_vec_dot": [_P, _I, _P, _P, _L, POINTER(c_double)],
    "pg_vec_cast_bf16": [_P, _P, _P, _L],
    "pg_vec_cast_u8_bf16": [_P, _P, _P, _L, _D],
    "pg_opt_gd": [_P, _I, _P, _P, _L, _D, _I, _P],
    "pg_opt_heavyball": [_P, _I, _P, _P, _P, _L, _D, _D, _I, _P],
    "pg_opt_nesterov": [_P, _I, _P, _P, _P, _L, _D, _D, _I, _P],
    "pg_opt_adam": [_P, _I, _P, _P, _P, _P, _P, _P, _L, _D, _D, _D, _D, _L, _I, _P],
    "pg_opt_adam_dev": [_P, _I, _P, _P, _P, _P, _P, _P, _L, _D, _D, _D, _D, _P, _I, _P],
    "pg_opt_nesterov_dev": [_P, _I, _P, _P, _P, _L, _D, _P, _I, _P],
    "pg_linear_fwd": [_P, _I, _P, _P, _P, _P, _L, _L, _L, _I],
    "pg_linear_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L],
    "pg_tc_linear_fwd": [_P, _P, _P, _P, _P, _I, _L, _L, _L, _I],
    "pg_tc_linear_bwd": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _L, _L, _L],
    "pg_tc_linear_bwd_ex": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _L, _L, _L],
    "pg_tc_gemm": [_P, _I, _P, _P, _P, _I, _L, _L, _L],
    "pg_conv2d_fwd": [_P, _I, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I],
    "pg_conv2d_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I],
    "pg_tc_conv2d_supported": [_I, _I, _I, _I, _I, _I],
    "pg_tc_conv2d_fwd": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I],
    "pg_tc_conv2d_bwd": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I],
    "pg_conv2d_relu_pool2_supported": [_I, _I, _I, _I, _I, _I],
    "pg_conv2d_relu_pool2_fwd": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I],
    "pg_conv2d_relu_pool2_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I],
    "pg_maxpool2d_fwd": [_P, _I, _P, _P, _L, _I, _I, _I],
    "pg_maxpool2d_bwd": [_P, _I, _P, _P, _P, _L, _I, _I, _I],
    "pg_maxpool2d_relu_bwd": [_P, _I, _P, _P, _P, _L, _I, _I, _I, _I],
    "pg_meanpool2d_fwd": [_P, _I, _P, _P, _L, _I, _I, _I],
    "pg_meanpool2d_bwd": [_P, _I, _P, _P, _L, _I, _I, _I],
    "pg_dropout": [_P, _I, _P, _P, _L, _D, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64],
    "pg_relu_fwd": [_P, _I, _P, _P, _L],
    "pg_relu_bwd": [_P, _I, _P, _P, _P, _L],
    "pg_softmax_ce_fwd_bwd": [_P, _I, _P, _P, _P, _P, _P, _L, _I, _D],
    "pg_softmax_ce_idx_fwd_bwd": [_P, _I, _P, POINTER(c_int32), _P, _P, _P, _L, _I, _D],
    "pg_softmax_fwd": [_P, _I, _P, _P, _L, _I],
    "pg_softmax_bwd": [_P, _I, _P, _P, _P, _L, _I],
    "pg_cross_entropy_fwd_bwd": [_P, _I, _P, _P, _P, _P, _L, _I, _D],
    "pg_mse_fwd_bwd": [_P, _I, _P, _P, _P, _P, _L, _D],
    "pg_class_error": [_P, _I, _P, _P, _L, _I, POINTER(c_double)],
    "pg_gather_rows": [_P, _I, _P, _P, _P, _L, _L],
    "pg_onehot": [_P, _I, _P, _P, _L, _I],
    "pg_loss_read": [_P, _I, _P, POINTER(c_double)],
    "pg_dropout_ex": [_P, _I, _I, _P, _P, _L, _L, _L, _L, _D, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64],
    "pg_tc_linear_bwd2": [_P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _L, _L, _L],
    "pg_arena_create": [_P, _I, _I, POINTER(c_int64), POINTER(_P)],
    "pg_arena_wrap": [_P, _I, _I, POINTER(c_int64), _P, POINTER(_P)],
    "pg_arena_destroy": [_P],
    "pg_arena_like": [_P, _I, POINTER(_P)],
    "pg_arena_info": [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int64), POINTER(c_int64), POINTER(_P)],
    "pg_arena_leaf_ptr": [_P, _I, POINTER(_P), POINTER(c_int64)],
    "pg_arena_upload": [_P, _I, _P],
    "pg_arena_download": [_P, _I, _P],
    "pg_arena_get_scalar": [_P, _L, POINTER(c_double)],
    "pg_arena_set_scalar": [_P, _L, _D],
    "pg_vec_expr": [_P, _I, _P, POINTER(_P), _I, POINTER(c_int32), _I, POINTER(c_double), _I, _L, _I, POINTER(c_int64)],
    "pg_mlp_step": [_P, _P],
    "pg_plan_create": [_P, _P, _I, _I, _I, POINTER(_P)],
    "pg_plan_destroy": [_P],
    "pg_plan_node_flags": [_P, _I, POINTER(c_int)],
    "pg_plan_forward": [_P, _P, _P, ctypes.c_uint64, ctypes.c_uint64, _L, POINTER(_P), POINTER(c_int64)],
    "pg_plan_backward": [_P, _L, _P, _I, _P, _P],
    "pg_plan_value": [_P, _I, POINTER(_P), POINTER(c_int), POINTER(c_int64)],
    "pg_dp_unique_id": [_P],
    "pg_dp_init": [_P, _I, _I, _P, _I, POINTER(_P)],
    "pg_dp_destroy": [_P],
    "pg_dp_info": [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)],
    "pg_dp_allreduce": [_P, _P, _L, _I],
    "pg_dp_broadcast": [_P, _P, _L, _I, _I],
    "pg_dp_wait": [_P],
    "pg_dp_bucket_begin": [_P, _P, _L],
    "pg_dp_bucket_finish": [_P, POINTER(c_int), POINTER(c_int64), POINTER(c_int64), _I],
    "pg_dp_bucket_wait": [_P, _I],
    "pg_dp_p2p_enable": [_P, _P, _P, _P],
    "pg_dp_p2p_set_optimizer": [_P, _I, _P, _P, _D, _D, _D, _D, _L, _I],
    "pg_dp_p2p_shadow_half": [_P, POINTER(c_int)],
    "pg_dp_set_option": [_P, c_char_p, _I],
    "pg_dp_bucket_plan": [_I, _I, POINTER(c_int64), _L, POINTER(c_int), _I, POINTER(c_int64), POINTER(c_int64), _I, POINTER(c_int)],
}
_VOID = {"pg_dp_bucket_ready": [_P, _I]}

EXPORTS = tuple(_SIGS) + tuple(_VOID) + ("pg_last_error",)

for _name, _args in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here == header/library drift, which must be loud
    _f.argtypes = _args
    _f.restype = c_int


for _name, _args in _VOID.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = None


def check(rc):
    if rc != 0:
        raise PGError(f"protograd_b200 error {rc}: {lib.pg_last_error().decode(errors='replace')}")


_FN = {name: getattr(lib, name) for name in _SIGS}
_err = lib.pg_last_error


def call(name, *args):
    if _FN[name](*args):
        raise PGError(f"protograd_b200 error in {name}: {_err().decode(errors='replace')}")
