"""ctypes binding of libilrl_b200.so (the C ABI of include/ilrl.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("ILRL_SO") or os.path.join(HERE, "libilrl_b200.so")  # ILRL_SO: development builds only

SYMBOLS = ["ilrl_create", "ilrl_destroy", "ilrl_last_error", "ilrl_load_clip", "ilrl_set_clip_ids", "ilrl_reset",
           "ilrl_step", "ilrl_step_sequence", "ilrl_step_host", "ilrl_step_host_async", "ilrl_wait", "ilrl_wait_step_host_async", "ilrl_serve_begin", "ilrl_serve_post", "ilrl_serve_wait", "ilrl_serve_step", "ilrl_serve_end", "ilrl_step_pull", "ilrl_pull", "ilrl_set_config", "ilrl_high_step", "ilrl_high_readout", "ilrl_get_state", "ilrl_set_state",
           "ilrl_set_heightfield", "ilrl_set_self_collision", "ilrl_set_forced_reset_noise", "ilrl_get_joint_target", "ilrl_set_joint_target",
           "ilrl_set_forced_target_deg", "ilrl_get_grouping", "ilrl_step_no_physics", "ilrl_physics_only", "ilrl_endpoint_score",
           "ilrl_stats", "ilrl_gae", "ilrl_episode_columns", "ilrl_gae_decisions", "ilrl_policy_blob_bytes", "ilrl_policy_pack", "ilrl_policy_step",
           "ilrl_launch_count", "ilrl_kernel_timing"]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("num_envs", C.c_int32), ("mode", C.c_int32), ("auto_reset", C.c_int32),
                ("seed", C.c_uint64), ("skip_frame", C.c_int32), ("max_timestep", C.c_int32),
                ("step_per_level", C.c_int32), ("env_id_base", C.c_int32)]


class IlrlError(RuntimeError):
    pass


_lib = None
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise IlrlError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback." % SO)
    L = C.CDLL(SO)
    L.ilrl_create.argtypes = [C.POINTER(Config), C.POINTER(_vp)]
    L.ilrl_destroy.argtypes = [_vp]
    L.ilrl_destroy.restype = None
    L.ilrl_last_error.argtypes = [_vp]
    L.ilrl_last_error.restype = C.c_char_p
    L.ilrl_load_clip.argtypes = [_vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_int32]
    L.ilrl_set_clip_ids.argtypes = [_vp, _vp]
    L.ilrl_reset.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_step.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_step_host.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_step_host_async.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_wait.argtypes = [_vp, C.c_int32]
    L.ilrl_serve_begin.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp]
    L.ilrl_serve_post.argtypes = [_vp, C.c_int32, _vp]
    L.ilrl_serve_wait.argtypes = [_vp, C.c_int32]
    L.ilrl_serve_step.argtypes = [_vp, _vp]
    L.ilrl_serve_end.argtypes = [_vp]
    L.ilrl_step_pull.argtypes = [_vp, _vp, C.c_int32, _vp, _vp]
    L.ilrl_pull.argtypes = [_vp, _vp, _vp, _vp]
    L.ilrl_wait_step_host_async.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_set_config.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32]
    L.ilrl_step_no_physics.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_get_grouping.argtypes = [_vp, _vp, _vp, _vp]
    L.ilrl_step_sequence.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ilrl_high_step.argtypes = [_vp, _vp, _vp, _vp]
    L.ilrl_high_readout.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.ilrl_get_state.argtypes = [_vp, _vp, _vp, _vp]
    L.ilrl_set_state.argtypes = [_vp, _vp, _vp, _vp]
    L.ilrl_set_forced_target_deg.argtypes = [_vp, _vp]
    L.ilrl_set_heightfield.argtypes = [_vp, _vp, C.c_int32, C.c_int32, C.c_float]
    L.ilrl_set_self_collision.argtypes = [_vp, C.c_int32]
    L.ilrl_set_forced_reset_noise.argtypes = [_vp, _vp]
    L.ilrl_get_joint_target.argtypes = [_vp, _vp, _vp]
    L.ilrl_set_joint_target.argtypes = [_vp, _vp, _vp]
    L.ilrl_physics_only.argtypes = [_vp, _vp, _vp]
    L.ilrl_endpoint_score.argtypes = [_vp, _vp, _vp]
    L.ilrl_stats.argtypes = [_vp, _vp, _vp]
    L.ilrl_gae.argtypes = [_vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, C.c_int32, C.c_int32, _vp]
    L.ilrl_episode_columns.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int64, _vp]
    L.ilrl_gae_decisions.argtypes = [_vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp]
    L.ilrl_policy_blob_bytes.restype = C.c_int64
    L.ilrl_policy_blob_bytes.argtypes = []
    L.ilrl_policy_pack.argtypes = [_vp] * 13 + [C.c_int32, C.c_int32, _vp, _vp]
    L.ilrl_policy_step.argtypes = [_vp] * 7 + [C.c_int32, C.c_int32, C.c_int32, _vp]
    L.ilrl_launch_count.argtypes = [_vp]
    L.ilrl_launch_count.restype = C.c_int64
    L.ilrl_kernel_timing.argtypes = [_vp, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    for s in SYMBOLS:
        if getattr(L, s).restype is C.c_int:
            getattr(L, s).restype = C.c_int
    _lib = L
    return L
