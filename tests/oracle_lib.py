"""ctypes binding of the CPU oracle (oracle/libpih_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libpih_oracle.so")


class PihConfig(C.Structure):
    """Mirror of include/pih.h: pih_config."""
    _fields_ = [("task", C.c_int32), ("arm", C.c_int32), ("fly_object", C.c_int32), ("num_envs", C.c_int32),
                ("sub_num", C.c_int32), ("env_id_base", C.c_int32), ("offset", C.c_float * 3), ("dt", C.c_float), ("solver_residual", C.c_float),
                ("solver_iters", C.c_int32), ("ik_iters", C.c_int32), ("max_steps", C.c_int32),
                ("auto_reset", C.c_int32), ("action_frame_env", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("kernel_family", C.c_int32), ("host_path", C.c_int32), ("peg_model", C.c_int32)]


def make_config(num_envs, task=0, arm=0, fly_object=0, sub_num=None, env_id_base=0, offset=(0.0, 0.0, 0.0),
                dt=1.0 / 240.0, solver_residual=1e-7, solver_iters=50, ik_iters=20, max_steps=0, auto_reset=0, action_frame_env=0,
                device=0, seed=0, peg_model=0):
    c = PihConfig()
    c.peg_model = peg_model
    c.task, c.arm, c.fly_object, c.num_envs = task, arm, fly_object, num_envs
    c.sub_num = num_envs if sub_num is None else sub_num
    c.env_id_base = env_id_base
    c.offset[:] = [float(v) for v in offset]
    c.dt = dt
    c.solver_residual = solver_residual
    c.solver_iters, c.ik_iters, c.max_steps, c.auto_reset = solver_iters, ik_iters, max_steps, auto_reset
    c.action_frame_env, c.device, c.seed = action_frame_env, device, seed
    return c


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.exists("/root/reference") or True:
            try:
                build_oracle()
            except Exception:
                if not os.path.exists(LIB):
                    raise
        L = C.CDLL(LIB)
        L.pih_oracle_create.restype = C.c_void_p
        L.pih_oracle_create.argtypes = [C.POINTER(PihConfig), C.c_int]
        L.pih_oracle_destroy.argtypes = [C.c_void_p]
        for name in ("pih_oracle_reset", "pih_oracle_step", "pih_oracle_rollout", "pih_oracle_get_state",
                     "pih_oracle_set_state", "pih_oracle_rollout_range"):
            getattr(L, name).restype = C.c_int
        L.pih_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pih_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.pih_oracle_rollout.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.pih_oracle_rollout_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.pih_oracle_advance_global_step.argtypes = [C.c_void_p, C.c_int]
        L.pih_oracle_state_words.argtypes = [C.POINTER(PihConfig)]
        L.pih_oracle_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.pih_oracle_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.pih_oracle_env_offset.argtypes = [C.POINTER(PihConfig), C.c_int, C.c_void_p]
        L.pih_oracle_fk.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pih_oracle_ik.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.pih_oracle_aba.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.pih_oracle_model_nb.argtypes = [C.c_int]
        L.pih_oracle_sincos.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.pih_oracle_atan2_q1.restype = C.c_float
        L.pih_oracle_atan2_q1.argtypes = [C.c_float, C.c_float]
        L.pih_oracle_rng4.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.pih_oracle_philox.argtypes = [C.c_void_p] * 3
        L.pih_oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pih_oracle_grasp_labels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


STATE_WORDS = 48
PS_POS, PS_QUAT, PS_VEL, PS_ANG, PS_GRASP, PS_QJ, PS_QDJ = 64, 67, 71, 74, 77, 80, 104      # flexible pipe block (peg_model=1)
S_Q, S_STEP, S_EPISODE, S_FLAGS, S_QD, S_PPOS, S_PQUAT, S_PVEL, S_PANG = 0, 9, 10, 11, 12, 24, 27, 32, 35
S_FSM_STATE, S_FSM_T, S_RANDY, S_MTGT, S_GRASP_ANG = 21, 22, 23, 40, 47


class Oracle:
    """Batched CPU oracle with the BaseEnvMp reset/step semantics."""

    def __init__(self, cfg, use_double=False):
        self.cfg, self.n = cfg, cfg.num_envs
        self.A = 4 if cfg.arm == 0 else 6
        self.h = lib().pih_oracle_create(C.byref(cfg), int(use_double))
        self.obs = np.zeros((self.n, 5), np.float32)
        self.reward = np.zeros(self.n, np.float32)
        self.done = np.zeros(self.n, np.uint8)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:     # (module globals may already be gone at interpreter exit)
            try:
                _lib.pih_oracle_destroy(self.h)
            except Exception:
                pass
            self.h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        if mask is None:
            self.reward[:] = 0
            self.done[:] = 0
        lib().pih_oracle_reset(self.h, _p(m), _p(self.obs))
        return self.obs.copy()

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32).reshape(self.n, self.A)
        rc = lib().pih_oracle_step(self.h, _p(a), _p(self.obs), _p(self.reward), _p(self.done))
        if rc != 0:
            raise RuntimeError(f"oracle step rc={rc}")
        return self.obs.copy(), self.reward.copy(), self.done.copy()

    def rollout(self, steps):
        cnt = np.zeros(2, np.uint64)
        rc = lib().pih_oracle_rollout(self.h, steps, _p(self.obs), _p(self.reward), _p(self.done), _p(cnt))
        if rc != 0:
            raise RuntimeError(f"oracle rollout rc={rc}")
        return self.obs.copy(), self.reward.copy(), self.done.copy(), cnt

    def render(self, width=300, height=300, lo=0, hi=None):
        hi = self.n if hi is None else hi
        img = np.zeros((hi - lo, height, width, 4), np.float32)
        lib().pih_oracle_render(self.h, width, height, lo, hi, _p(img))
        return img

    def render_grasp(self, width=300, height=300, lo=0, hi=None):
        hi = self.n if hi is None else hi
        img = np.zeros((hi - lo, height, width, 4), np.float32)
        lib().pih_oracle_render_grasp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib().pih_oracle_render_grasp(self.h, width, height, lo, hi, _p(img))
        return img

    def grasp_labels(self, width=300, height=300, lo=0, hi=None):
        hi = self.n if hi is None else hi
        lab = np.zeros((hi - lo, 4, height, width), np.float32)
        par = np.zeros((hi - lo, 5), np.float32)
        lib().pih_oracle_grasp_labels(self.h, width, height, lo, hi, _p(lab), _p(par))
        return lab, par

    @property
    def state_words(self):
        return int(lib().pih_oracle_state_words(C.byref(self.cfg)))

    def get_state(self):
        w = np.zeros((self.state_words // 4, self.n, 4), np.float32)
        lib().pih_oracle_get_state(self.h, _p(w))
        return w

    def set_state(self, w):
        w = np.ascontiguousarray(w, np.float32)
        assert w.shape == (self.state_words // 4, self.n, 4)
        lib().pih_oracle_set_state(self.h, _p(w))


def state_field(w, word, count=1):
    """View words [word, word+count) of the [slot][env][4] state buffer as an (N, count) array."""
    n = w.shape[1]
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    return flat[:, word:word + count]


def state_int(w, word):
    return state_field(w, word, 1).copy().view(np.int32)[:, 0]


def fk(arm, q, use_double=True):
    q = np.ascontiguousarray(q, np.float64)
    pos, rot = np.zeros(3), np.zeros(9)
    lib().pih_oracle_fk(arm, int(use_double), _p(q), _p(pos), _p(rot))
    return pos, rot.reshape(3, 3)


def ik(arm, q, tpos, tquat, iters=20, use_double=True):
    q = np.ascontiguousarray(q, np.float64)
    n = 7 if arm == 0 else 6
    out = np.zeros(n)
    lib().pih_oracle_ik(arm, int(use_double), iters, _p(q), _p(np.ascontiguousarray(tpos, np.float64)),
                        _p(np.ascontiguousarray(tquat, np.float64)), _p(out))
    return out


def aba(arm, q, qd, tau, g=9.8):
    nb = lib().pih_oracle_model_nb(arm)
    qdd, minv = np.zeros(nb), np.zeros((nb, nb))
    lib().pih_oracle_aba(arm, _p(np.ascontiguousarray(q, np.float64)), _p(np.ascontiguousarray(qd, np.float64)),
                         _p(np.ascontiguousarray(tau, np.float64)), g, _p(qdd), _p(minv))
    return qdd, minv


def pipe_dyn(bquat, bvel, bang, qj, qdj, dt=1.0 / 240.0, g=9.8):
    """(vstar[29], col[29, 29]) of the flexible pipe from the double oracle (base twist in body coordinates)."""
    vstar, col = np.zeros(29), np.zeros((29, 29))
    L = lib()
    L.pih_oracle_pipe_dyn.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    arrs = [np.ascontiguousarray(a, np.float64) for a in (bquat, bvel, bang, qj, qdj)]
    L.pih_oracle_pipe_dyn(*[_p(a) for a in arrs], dt, g, _p(vstar), _p(col))
    return vstar, col


def pipe_fk(bpos, bquat, qj):
    pos, rot = np.zeros((24, 3)), np.zeros((24, 9))
    L = lib()
    L.pih_oracle_pipe_fk.argtypes = [C.c_void_p] * 5
    arrs = [np.ascontiguousarray(a, np.float64) for a in (bpos, bquat, qj)]
    L.pih_oracle_pipe_fk(*[_p(a) for a in arrs], _p(pos), _p(rot))
    return pos, rot.reshape(24, 3, 3)
