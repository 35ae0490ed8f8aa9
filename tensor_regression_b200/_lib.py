"""ctypes binding of libtrb200.so (include/tr_b200.h).  No torch types cross this boundary:
raw device pointers, sizes and a cudaStream_t."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TR_B200_LIB') or os.path.join(_HERE, 'libtrb200.so')   # env override: kernel-variant experiments

TR_F32, TR_F64 = 0, 1

SYMBOLS = [
    'tr_version', 'tr_create', 'tr_destroy', 'tr_last_error', 'tr_param_count', 'tr_gradsum_count',
    'tr_reserve', 'tr_forward_std', 'tr_forward_mn', 'tr_fwd_grad_std', 'tr_fwd_grad_mn',
    'tr_backward_std', 'tr_backward_mn', 'tr_finish_grad', 'tr_adam_step', 'tr_last_launch_info', 'tr_profile_enable',
    'tr_profile_read', 'tr_set_option', 'tr_lbfgs_direction', 'tr_lbfgs_point', 'tr_lbfgs_gtd',
    'tr_adam_step_groups', 'tr_allreduce', 'tr_comm_unique_id', 'tr_comm_create', 'tr_comm_destroy',
    'tr_upload', 'tr_upload_stats', 'tr_host_last_error',
    'tr_spec_create', 'tr_spec_fwd_grad', 'tr_spec_forward',
]


class TRError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise TRError(
            f'{LIB_PATH} not found: the CUDA extension is not built. Run '
            f'`python -c "import __graft_entry__ as g; g.build()"` or '
            f'`tensor_regression_b200/csrc/build.sh`. There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, u32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32, ctypes.c_double
    lib.tr_version.restype = i32
    lib.tr_create.argtypes = [ctypes.POINTER(vp), i32, i32, ctypes.POINTER(i64), i32, i32, i32]
    lib.tr_destroy.argtypes = [vp]
    lib.tr_last_error.argtypes = [vp]
    lib.tr_last_error.restype = ctypes.c_char_p
    lib.tr_param_count.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.tr_gradsum_count.argtypes = [vp, ctypes.POINTER(i64)]
    lib.tr_reserve.argtypes = [vp, i64]
    lib.tr_forward_std.argtypes = [vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp]
    lib.tr_forward_mn.argtypes = [vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp, vp]
    lib.tr_fwd_grad_std.argtypes = [vp, vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp, vp]
    lib.tr_fwd_grad_mn.argtypes = [vp, vp, vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp, vp]
    lib.tr_backward_std.argtypes = [vp, vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp]
    lib.tr_backward_mn.argtypes = [vp, vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp]
    lib.tr_finish_grad.argtypes = [vp, vp, dbl, dbl, vp, dbl, u32, dbl, dbl, vp, vp, vp]
    lib.tr_adam_step.argtypes = [vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl, dbl, vp]
    lib.tr_adam_step_groups.argtypes = [vp, vp, vp, vp, vp, vp, i64, ctypes.POINTER(dbl), i32, dbl, dbl, dbl, dbl, vp]
    lib.tr_allreduce.argtypes = [vp, vp, i64, vp, vp]
    lib.tr_comm_unique_id.argtypes = [vp]
    lib.tr_comm_create.argtypes = [ctypes.POINTER(vp), vp, i32, i32, i32]
    lib.tr_comm_destroy.argtypes = [vp]
    lib.tr_upload.argtypes = [vp, vp, ctypes.c_size_t, i32, i32, ctypes.c_size_t, vp]
    lib.tr_upload_stats.argtypes = [ctypes.POINTER(dbl)]
    lib.tr_host_last_error.argtypes = []
    lib.tr_host_last_error.restype = ctypes.c_char_p
    lib.tr_last_launch_info.argtypes = [vp, ctypes.POINTER(i64)]
    lib.tr_profile_enable.argtypes = [vp, i32]
    lib.tr_profile_read.argtypes = [vp, ctypes.POINTER(dbl)]
    lib.tr_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.tr_lbfgs_direction.argtypes = [vp, vp, vp, vp, dbl, i32, vp, vp, vp, i32, vp, vp]
    lib.tr_lbfgs_point.argtypes = [vp, vp, vp, dbl, vp, vp]
    lib.tr_lbfgs_gtd.argtypes = [vp, vp, vp, vp, vp]
    lib.tr_spec_create.argtypes = [ctypes.POINTER(vp), i32, i64, i64, i64, i32, i32, i32, i32]
    lib.tr_spec_fwd_grad.argtypes = [vp, vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp, vp]
    lib.tr_spec_forward.argtypes = [vp, vp, i64, vp, vp, u32, dbl, dbl, vp, vp, vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ('tr_last_error', 'tr_host_last_error'):
            fn.restype = i32
    return lib


lib = _load()
