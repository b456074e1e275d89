"""ctypes binding of the C ABI in include/pih.h (csrc/libpih_b200.so).

There is NO CPU fallback: if the CUDA library has not been built, importing this module raises, and if no CUDA
device is present pih_create fails.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``peg_in_hole_gym_b200/csrc/build.sh``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIH_LIB", os.path.join(_HERE, "csrc", "libpih_b200.so"))  # PIH_LIB: perf experiments only

PIH_OK, PIH_ERR_INVALID, PIH_ERR_CUDA, PIH_ERR_STATE = 0, -1, -2, -3
TASK_IDS = {"peg-in-hole": 0, "random-fly": 1, "peg-in-hole-scripted": 2}
ARM_IDS = {"panda": 0, "ur5": 1}
FLY_OBJECTS = {"banana": 0, "amicelli": 1}
STATE_F4 = 12


class PihConfig(C.Structure):
    """Mirror of ``pih_config`` (include/pih.h)."""
    _fields_ = [("task", C.c_int32), ("arm", C.c_int32), ("fly_object", C.c_int32), ("num_envs", C.c_int32),
                ("sub_num", C.c_int32), ("env_id_base", C.c_int32), ("offset", C.c_float * 3), ("dt", C.c_float), ("solver_residual", C.c_float),
                ("solver_iters", C.c_int32), ("ik_iters", C.c_int32), ("max_steps", C.c_int32),
                ("auto_reset", C.c_int32), ("action_frame_env", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("kernel_family", C.c_int32), ("host_path", C.c_int32), ("peg_model", C.c_int32)]


class NativeLibraryMissing(ImportError):
    pass


_lib = None


def lib():
    """Load libpih_b200.so; raises loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.pih_abi_version.restype = C.c_int
    L.pih_state_words.restype = C.c_int
    L.pih_state_words_for.restype = C.c_int
    L.pih_state_words_for.argtypes = [C.POINTER(PihConfig)]
    L.pih_action_dim.argtypes = [C.c_int]
    L.pih_obs_dim.restype = C.c_int
    L.pih_create.argtypes = [C.POINTER(PihConfig), vp, C.POINTER(vp)]
    L.pih_destroy.argtypes = [vp]
    L.pih_destroy.restype = None
    L.pih_last_error.argtypes = [vp]
    L.pih_last_error.restype = C.c_char_p
    L.pih_env_offset.argtypes = [C.POINTER(PihConfig), i32, C.POINTER(C.c_float)]
    L.pih_reset.argtypes = [vp, vp, vp, vp]
    L.pih_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pih_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pih_rollout.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.pih_host_bind.argtypes = [vp, vp, C.c_uint64]
    L.pih_host_unbind.argtypes = [vp, vp]
    L.pih_set_stats.argtypes = [vp, vp]
    L.pih_ik_targets.argtypes = [vp, vp, vp, vp]
    L.pih_ee_pose.argtypes = [vp, vp, vp, vp]
    L.pih_render.argtypes = [vp, i32, i32, vp, vp]
    L.pih_grasp_labels.argtypes = [vp, i32, i32, vp, vp, vp]
    L.pih_render_grasp.argtypes = [vp, i32, i32, vp, vp]
    L.pih_step_signal.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.pih_step_signal_wait.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32, vp, i32, C.c_uint32, C.c_uint32, vp, vp]
    L.pih_step_publish_prev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32, vp, i32, C.c_uint32, C.c_uint32, vp, vp]
    L.pih_peer_alloc.argtypes = [i32, C.c_uint64, C.POINTER(vp), C.c_char_p]
    L.pih_peer_open.argtypes = [i32, C.c_char_p, C.POINTER(vp)]
    L.pih_peer_release.argtypes = [vp, i32]
    L.pih_enable_peer.argtypes = [vp, i32]
    L.pih_signal_flag.argtypes = [vp, vp, C.c_uint32, vp]
    L.pih_wait_flags.argtypes = [vp, vp, i32, C.c_uint32, C.c_uint32, vp, vp]
    L.pih_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    for name in ("pih_create", "pih_env_offset", "pih_reset", "pih_step", "pih_step_host", "pih_rollout",
                 "pih_ik_targets", "pih_ee_pose", "pih_measure_fp32_peak", "pih_action_dim", "pih_render", "pih_grasp_labels",
                 "pih_step_signal", "pih_peer_alloc", "pih_peer_open", "pih_peer_release", "pih_enable_peer", "pih_signal_flag", "pih_wait_flags",
                 "pih_host_bind", "pih_host_unbind", "pih_set_stats", "pih_step_signal_wait", "pih_render_grasp", "pih_step_publish_prev"):
        getattr(L, name).restype = C.c_int
    if L.pih_abi_version() != 2:
        raise ImportError("libpih_b200.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = ["pih_abi_version", "pih_state_words", "pih_state_words_for", "pih_action_dim", "pih_obs_dim", "pih_create", "pih_destroy",
           "pih_last_error", "pih_env_offset", "pih_reset", "pih_step", "pih_step_host", "pih_rollout",
           "pih_ik_targets", "pih_ee_pose", "pih_render", "pih_grasp_labels", "pih_step_signal", "pih_peer_alloc", "pih_peer_open", "pih_peer_release",
           "pih_enable_peer", "pih_signal_flag",
           "pih_wait_flags", "pih_measure_fp32_peak", "pih_host_bind", "pih_host_unbind", "pih_set_stats", "pih_step_signal_wait", "pih_render_grasp", "pih_step_publish_prev"]
KERNEL_FAMILIES = {"auto": 0, "ws": 1, "fused": 2, "split": 3, "ws2": 4}
PEG_MODELS = {"rigid": 0, "pipe": 1}      # pih_config.peg_model: one rigid segment / the reference's 25-link pipe.urdf chain


def make_config(num_envs, task="peg-in-hole", arm="panda", fly_object="banana", sub_num=None, env_id_base=0,
                offset=(0.0, 0.0, 0.0), dt=1.0 / 240.0, solver_residual=1e-7, solver_iters=50, ik_iters=20, max_steps=0, auto_reset=False,
                action_frame_env=False, device=0, seed=0, kernel_family="auto", host_path=0, peg_model="rigid"):
    c = PihConfig()
    c.peg_model = PEG_MODELS[peg_model] if isinstance(peg_model, str) else int(peg_model)
    c.kernel_family = KERNEL_FAMILIES[kernel_family] if isinstance(kernel_family, str) else int(kernel_family)
    c.host_path = int(host_path)
    c.task = TASK_IDS[task] if isinstance(task, str) else int(task)
    c.arm = ARM_IDS[arm] if isinstance(arm, str) else int(arm)
    c.fly_object = FLY_OBJECTS[fly_object.lower()] if isinstance(fly_object, str) else int(fly_object)
    c.num_envs = int(num_envs)
    c.sub_num = int(num_envs if sub_num is None else sub_num)
    c.env_id_base = int(env_id_base)
    c.offset[:] = [float(v) for v in offset]
    c.dt = float(dt)
    c.solver_residual = float(solver_residual)
    c.solver_iters, c.ik_iters, c.max_steps = int(solver_iters), int(ik_iters), int(max_steps)
    c.auto_reset, c.action_frame_env, c.device, c.seed = int(bool(auto_reset)), int(bool(action_frame_env)), int(device), int(seed)
    return c


class PihError(RuntimeError):
    pass


def check(rc, handle=None, what=""):
    if rc == PIH_OK:
        return
    msg = lib().pih_last_error(handle).decode() if handle else ""
    if rc == PIH_ERR_STATE:
        # the reference raises AttributeError when step() is called before reset() (base_env_mp.py:42)
        raise AttributeError(f"{what}: step() before reset() ({msg})")
    kind = {PIH_ERR_INVALID: "invalid argument", PIH_ERR_CUDA: "CUDA error / no device"}.get(rc, f"error {rc}")
    raise PihError(f"{what}: {kind} {msg}")
