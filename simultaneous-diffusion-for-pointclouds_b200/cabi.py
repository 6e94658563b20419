"""ctypes binding of libsdpc_b200.so (include/sdpc_b200.h).

The product path has no CPU fallback: `load()` raises if the shared library is missing
(run `python -c "import __graft_entry__ as g; g.build()"` at the repo root to compile it).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDPC_LIB") or os.path.join(_HERE, "libsdpc_b200.so")   # SDPC_LIB: A/B timing of two builds

ABI_VERSION = 2          # SDPC_ABI_VERSION of include/sdpc_b200.h
SDPC_VARIANT_POSE, SDPC_VARIANT_TRANSLATION = 0, 1
PREC_FP32, PREC_TF32, PREC_BF16, PREC_BF16X3, PREC_FP16 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp16": PREC_FP16}


class StepParams(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("group_size", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("big_rows", C.c_int32), ("variant", C.c_int32), ("share", C.c_int32), ("nan_to_num", C.c_int32),
        ("sky_filter", C.c_int32), ("tgt_first", C.c_int32), ("tgt_count", C.c_int32), ("scalar_div_recip", C.c_int32),
        ("key_shift_override", C.c_int32), ("winner_mode", C.c_int32),
        ("step_size", C.c_float), ("noise_scale", C.c_float), ("grad_ref", C.c_float), ("corr_coef", C.c_float),
        ("sigma_mod", C.c_float), ("min_depth_thr", C.c_float),
        ("allowance", C.c_double), ("h_min", C.c_double), ("dh", C.c_double),
        ("big_row_min", C.c_double), ("dv", C.c_double),
    ]


_BUF_FIELDS = ["x", "grad", "noise", "refer", "mask", "sky", "exist", "to_world", "from_world", "origins",
               "cos_az", "sin_az", "cos_el", "sin_el", "grad_likelihood", "new_images", "too_high",
               "dbg_row", "dbg_col", "dbg_valid", "dbg_cnt", "dbg_winner", "dbg_min_d"]


class StepBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _BUF_FIELDS]


class ScoreConfig(C.Structure):
    _fields_ = [("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("ngf", C.c_int32),
                ("num_classes", C.c_int32), ("precision", C.c_int32), ("max_views", C.c_int32),
                ("reserved", C.c_int32)]


class ProjectionParams(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("point_stride", C.c_int32), ("intensity_col", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("reserved", C.c_int32), ("origin", C.c_double * 3), ("h_min", C.c_double),
                ("dh", C.c_double), ("v_min", C.c_double), ("dv", C.c_double)]


# every symbol include/sdpc_b200.h declares: (name, restype, argtypes)
_P, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
SYMBOLS = [
    ("sdpc_abi_version", _I, []),
    ("sdpc_abi_struct_bytes", _SZ, [_I]),
    ("sdpc_last_error", C.c_char_p, []),
    ("sdpc_build_arch", C.c_char_p, []),
    ("sdpc_score_create", _I, [C.POINTER(ScoreConfig), C.POINTER(_P)]),
    ("sdpc_score_destroy", _I, [_P]),
    ("sdpc_score_param_count", _I, [_P]),
    ("sdpc_score_param_info", _I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(_I)]),
    ("sdpc_score_load_param", _I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I, _I, _P]),
    ("sdpc_score_finalize", _I, [_P, _P]),
    ("sdpc_score_workspace_bytes", _SZ, [_P, _I]),
    ("sdpc_score_forward", _I, [_P, _P, _P, _P, _I, _P, _SZ, _P]),
    ("sdpc_score_read_tap", _I, [_P, C.c_char_p, _P, _SZ, _I, C.POINTER(_I), _P]),
    ("sdpc_score_last_launch_count", _I, [_P]),
    ("sdpc_score_flops_per_view", C.c_double, [_P]),
    ("sdpc_score_set_profiling", _I, [_P, _I]),
    ("sdpc_score_profile_collect", _I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I)]),
    ("sdpc_step_workspace_bytes", _SZ, [_I, _I, _I, _I]),
    ("sdpc_step_workspace_init", _I, [_P, _SZ, _I, _I, _I, _I, _P]),
    ("sdpc_langevin_update", _I, [C.POINTER(StepParams), C.POINTER(StepBuffers), _P, _SZ, _P]),
    ("sdpc_step_merge_max", _I, [_P, _P, _I, _P]),
    ("sdpc_step_read_max", _I, [_P, _P, _P]),
    ("sdpc_shard_slot_floats", _SZ, [_I, _I, _I]),
    ("sdpc_shard_pack", _I, [_P, _P, _P, _I, _I, _I, _P]),
    ("sdpc_shard_unpack", _I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    ("sdpc_crossview_share", _I, [C.POINTER(StepParams), C.POINTER(StepBuffers), _P, _SZ, _P]),
    ("sdpc_step_kernel_launches", _I, [C.POINTER(StepParams), C.POINTER(StepBuffers)]),
    ("sdpc_langevin_reproject_step", _I, [C.POINTER(StepParams), C.POINTER(StepBuffers), _P, _SZ, _P]),
    ("sdpc_projection_workspace_bytes", _SZ, [_I, _I]),
    ("sdpc_pointcloud_to_range_image", _I, [C.POINTER(ProjectionParams), _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    ("sdpc_points_workspace_bytes", _SZ, [_I, _I, _I]),
    ("sdpc_range_image_to_points", _I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    ("sdpc_depth_intensity_errors", _I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    ("sdpc_transform_scan", _I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P]),
    ("sdpc_range_image_postprocess", _I, [_P, _P, _P, _P, _I, _I, C.c_double, _P, _P, _P, _P]),
    ("sdpc_langevin_reproject_step_host", _I, [C.POINTER(StepParams), C.POINTER(StepBuffers), _P, _P, _P, _SZ, _P, _P, _P, _P,
                                               _P, _SZ, _P]),
]

_lib = None


class SdpcError(RuntimeError):
    pass


def load(path=None):
    """dlopen the C-ABI library and type every symbol; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SdpcError(f"{path} not found: the CUDA extension is not built (no CPU fallback exists). "
                        f"Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.sdpc_abi_version() != ABI_VERSION:
        raise SdpcError(f"{path}: ABI version {lib.sdpc_abi_version()} != {ABI_VERSION} of this binding (stale build: run "
                        f"`python -c 'import __graft_entry__ as g; g.build()'`)")
    for which, st in enumerate((StepParams, StepBuffers, ScoreConfig, ProjectionParams)):
        if lib.sdpc_abi_struct_bytes(which) != C.sizeof(st):
            raise SdpcError(f"{path}: sizeof({st.__name__}) is {lib.sdpc_abi_struct_bytes(which)} in the library, "
                            f"{C.sizeof(st)} in this binding")
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(lib, status, what):
    if status != 0:
        raise SdpcError(f"{what} failed with status {status}: {lib.sdpc_last_error().decode()}")
