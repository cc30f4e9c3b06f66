"""ctypes binding of include/trl_b200.h (the C ABI). Fails loudly if the CUDA library is missing."""
import ctypes as C
import os

from . import build as _build

_dp = C.POINTER(C.c_double)


class ObsCfg(C.Structure):
    _fields_ = [("obs_kind", C.c_int), ("sampler", C.c_int), ("num_samples", C.c_int), ("grid_res", C.c_int),
                ("view_min", C.c_double), ("view_max", C.c_double), ("num_phase_bins", C.c_int)]


class StepArgs(C.Structure):
    _fields_ = [("dt", C.c_double),
                ("pose", C.c_void_p), ("vel", C.c_void_p), ("tar_pose", C.c_void_p), ("tar_vel", C.c_void_p),
                ("joint_active", C.c_void_p), ("kin_time", C.c_void_p), ("origin_pos", C.c_void_p),
                ("origin_rot", C.c_void_p), ("body_pos", C.c_void_p), ("body_vel", C.c_void_p),
                ("phase", C.c_void_p), ("flip", C.c_void_p),
                ("out_tau", C.c_void_p), ("out_kin_pose", C.c_void_p), ("out_kin_vel", C.c_void_p),
                ("out_kin_body_pos", C.c_void_p), ("out_state", C.c_void_p)]


class StepOpts(C.Structure):
    """trl_step_opts (include/trl_b200.h)."""
    _fields_ = [("out_reward", C.c_void_p), ("fallen", C.c_void_p), ("out_state_f32", C.c_void_p)]


TERRAIN_NUM_PARAMS = 40


class TerrainCfg(C.Structure):
    """trl_terrain_cfg (include/trl_b200.h)."""
    _fields_ = [("type", C.c_int), ("pad", C.c_int), ("params", C.c_double * TERRAIN_NUM_PARAMS),
                ("ground_width", C.c_double), ("world_scale", C.c_double), ("spacing_x", C.c_double),
                ("spacing_z", C.c_double)]


# every symbol include/trl_b200.h declares: (name, restype, argtypes)
_VP = C.c_void_p
SYMBOLS = [
    ("trl_model_load", _VP, [C.c_char_p, C.c_char_p, _VP]),
    ("trl_model_create", _VP, [C.c_int, _VP, _VP, _VP, C.c_int, _VP, C.c_int, _VP]),
    ("trl_model_dims", C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("trl_model_get_matrices", C.c_int, [_VP, _VP, _VP, _VP]),
    ("trl_model_get_motion", C.c_int, [_VP, _VP, C.POINTER(C.c_int)]),
    ("trl_model_free", None, [_VP]),
    ("trl_arg_file_get", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    ("trl_state_read", C.c_int, [_VP, C.c_char_p, _VP, _VP]),
    ("trl_state_write", C.c_int, [_VP, C.c_char_p, _VP, _VP]),
    ("trl_batch_create", _VP, [_VP, C.c_int, C.c_int, C.POINTER(ObsCfg)]),
    ("trl_batch_destroy", None, [_VP]),
    ("trl_batch_set_pointer_mode", C.c_int, [_VP, C.c_int]),
    ("trl_batch_set_stream", C.c_int, [_VP, _VP]),
    ("trl_batch_sync", C.c_int, [_VP]),
    ("trl_poli_state_size", C.c_int, [_VP]),
    ("trl_batch_num_envs", C.c_int, [_VP]),
    ("trl_batch_launch_count", C.c_longlong, [_VP]),
    ("trl_batch_set_host_async", C.c_int, [_VP, C.c_int]),
    ("trl_debug_trace_enable", C.c_int, [C.c_int]),
    ("trl_debug_trace_read", C.c_int, [C.POINTER(C.c_ulonglong)]),
    ("trl_imp_pd_update_control_force", C.c_int, [_VP, C.c_double] + [_VP] * 10),
    ("trl_exp_pd_update_control_force", C.c_int, [_VP, C.c_double] + [_VP] * 8),
    ("trl_kin_tree_fk", C.c_int, [_VP, _VP, _VP, _VP]),
    ("trl_rbd_extras", C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    ("trl_rbd_end_effector_jacobian", C.c_int, [_VP, _VP, _VP, C.c_int, _VP]),
    ("trl_apply_control_forces", C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    ("trl_kin_char_pose", C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    ("trl_ground_set_heightfields", C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, C.c_longlong, _VP, _VP, _VP, _VP, _VP, _VP]),
    ("trl_ground_sample_height", C.c_int, [_VP, _VP, C.c_int, _VP, _VP, _VP]),
    ("trl_terrain_default_cfg", C.c_int, [C.POINTER(TerrainCfg)]),
    ("trl_terrain_load", C.c_int, [C.c_char_p, C.c_double, C.POINTER(TerrainCfg)]),
    ("trl_ground_generate", C.c_int, [_VP, C.POINTER(TerrainCfg), C.c_int, _VP, _VP, _VP, _VP]),
    ("trl_ground_update", C.c_int, [_VP, _VP, _VP, _VP]),
    ("trl_ground_read_field", C.c_longlong, [_VP, C.c_int, _VP, C.c_longlong, _VP, _VP, _VP, _VP, _VP]),
    ("trl_build_poli_state", C.c_int, [_VP] * 12),
    ("trl_step_fused", C.c_int, [_VP, C.POINTER(StepArgs)]),
    ("trl_step_fused_ex", C.c_int, [_VP, C.POINTER(StepArgs), C.POINTER(StepOpts)]),
    ("trl_imitate_calc_reward", C.c_int, [_VP] * 9),
    ("trl_batch_get_status", C.c_int, [_VP, _VP]),
    ("trl_device_fp64_probe", C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    ("trl_last_error", C.c_char_p, []),
]

_lib = None


def lib():
    """Load terrainrlsim_b200/lib/libtrl_b200.so (building it with nvcc if it is stale or missing)."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if _build.is_stale():
            if _build.nvcc_path() is None and not os.path.exists(path):
                raise RuntimeError("libtrl_b200.so is missing and nvcc is not available; "
                                   "terrainrlsim_b200 has no CPU fallback")
            if _build.nvcc_path() is not None:
                path = _build.build()
        path = os.environ.get("TRL_LIB", path)  # tuning hook: load an alternative build of the same ABI
        l = C.CDLL(path)
        for name, res, args in SYMBOLS:
            f = getattr(l, name)  # AttributeError here = the library does not export the ABI
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def last_error():
    return lib().trl_last_error().decode("utf-8", "replace")


class TrlError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise TrlError("%s failed (%d): %s" % (what, rc, last_error()))
