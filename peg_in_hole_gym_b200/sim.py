"""BatchedSim - thin torch-side owner of the HBM-resident env state around the C ABI (include/pih.h).

This is the layer that replaces L4+L3+L1+L0 of the reference: the multiprocessing fan-out
(envs/base_env_mp.py:27-51), the per-sub-env Python loops (envs/base_env.py:60-71) and the PyBullet calls.  PyTorch
is used for device memory and streams only; all arithmetic runs in the hand-written sm_100a kernels.
"""
import ctypes as C

import torch

from . import _native as N


class HostStepper:
    """Owner of one set of host buffers bound to a simulator (pih_host_bind / pih_host_unbind): the lifetime contract of
    the zero-copy host path is this object's lifetime."""

    def __init__(self, sim, action_nps, obs_np, reward_np, done_np):
        self._sim, self._keep = sim, (list(action_nps), obs_np, reward_np, done_np)
        self._a = [C.c_void_p(a.ctypes.data) for a in action_nps]
        self._o, self._r, self._d = (C.c_void_p(x.ctypes.data) for x in (obs_np, reward_np, done_np))
        self._fn, self._h, self._st = sim._L.pih_step_host, sim._h, C.c_void_p(sim._stream())
        self._bound = []
        spans = sorted((a.ctypes.data, a.nbytes) for a in list(action_nps) + [obs_np, reward_np, done_np])
        merged = []
        for addr, nb in spans:                                  # views of one page-locked block bind as one range
            if merged and merged[-1][0] + merged[-1][1] == addr:
                merged[-1][1] += nb
            else:
                merged.append([addr, nb])
        for addr, nb in merged:
            if sim._L.pih_host_bind(sim._h, addr, nb) == 0:      # pageable buffers stay unbound (queried per call)
                self._bound.append(addr)

    def __call__(self, i):
        rc = self._fn(self._h, self._a[i], self._o, self._r, self._d, self._st)
        if rc:
            N.check(rc, self._h, "pih_step_host")

    def close(self):
        if self._bound and getattr(self._sim, "_h", None):
            for p in self._bound:
                self._sim._L.pih_host_unbind(self._sim._h, p)
        self._bound = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedSim:
    """num_envs sub-environments resident on one GPU.  State lives in a torch-owned float4-SoA tensor
    ``state[12, N, 4]`` (``[32, N, 4]`` with ``peg_model='pipe'``; layout: csrc/pih_constants.h); outputs in ``obs[N,5]``, ``reward[N]``, ``done[N]``."""

    def __init__(self, num_envs, task="peg-in-hole", arm="panda", fly_object="banana", sub_num=None, env_id_base=0,
                 offset=(0.0, 0.0, 0.0), dt=1.0 / 240.0, solver_residual=1e-7, solver_iters=50, ik_iters=20, max_steps=0, auto_reset=False,
                 action_frame_env=False, device=None, seed=0, kernel_family="auto", host_path=0, peg_model="rigid"):
        L = N.lib()
        if not torch.cuda.is_available():
            raise N.PihError("peg_in_hole_gym_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.cfg = N.make_config(num_envs, task, arm, fly_object, sub_num, env_id_base, offset, dt, solver_residual,
                                 solver_iters, ik_iters, max_steps, auto_reset, action_frame_env, self.device.index, seed,
                                 kernel_family, host_path, peg_model)
        self.n = int(num_envs)
        self.action_dim = L.pih_action_dim(self.cfg.arm)
        self.obs_dim = L.pih_obs_dim()
        with torch.cuda.device(self.device):
            self.state_f4 = L.pih_state_words_for(C.byref(self.cfg)) // 4       # 12 slots, 32 with the flexible pipe
            self.state = torch.zeros(self.state_f4, self.n, 4, dtype=torch.float32, device=self.device)
            self.obs = torch.zeros(self.n, self.obs_dim, dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
            self.counters = torch.zeros(2, dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
            h = C.c_void_p()
            N.check(L.pih_create(C.byref(self.cfg), self.state.data_ptr(), C.byref(h)), None, "pih_create")
        self._h = h
        self._L = L

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.pih_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- reference API equivalents
    def reset(self, mask=None):
        """BaseEnv.reset (envs/base_env.py:84-94).  mask: optional uint8/bool tensor [N]; None resets every env."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            self.reward[m.bool()] = 0
            self.done[m.bool()] = 0
        else:
            self.reward.zero_()
            self.done.zero_()
        N.check(self._L.pih_reset(self._h, m.data_ptr() if m is not None else None, self.obs.data_ptr(), self._stream()),
                self._h, "pih_reset")
        return self.obs

    def step(self, action):
        """BaseEnv.step (envs/base_env.py:60-75) for all envs in one launch.  action: float32 CUDA tensor [N, A]."""
        a = action
        if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()):
            a = torch.as_tensor(a, dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
        if a.numel() != self.n * self.action_dim:
            raise ValueError(f"action has {a.numel()} elements, expected {self.n}x{self.action_dim}")
        N.check(self._L.pih_step(self._h, a.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                                 self._stream()), self._h, "pih_step")
        return self.obs, self.reward, self.done

    def step_host(self, action_np, obs_np, reward_np, done_np):
        """Host-buffer plugin call (pih_step_host): numpy in, numpy out, H2D/D2H inside."""
        rc = self._L.pih_step_host(self._h, action_np.ctypes.data, obs_np.ctypes.data, reward_np.ctypes.data,
                                   done_np.ctypes.data, self._stream())
        if rc:
            N.check(rc, self._h, "pih_step_host")

    def host_stepper(self, action_nps, obs_np, reward_np, done_np):
        """Binds pih_step_host to fixed host buffers once and returns a callable ``step(i)`` which steps with
        ``action_nps[i]``: the per-call Python work is one ctypes call.  With page-locked buffers
        (``pinned_host_buffers``) the library runs its zero-copy path; the stepper keeps the buffers alive and bound
        (pih_host_bind) for exactly as long as it lives."""
        return HostStepper(self, action_nps, obs_np, reward_np, done_np)

    def pinned_host_buffers(self, n_action_buffers=1):
        """Page-locked numpy views for the host-buffer call: `n_action_buffers` action arrays [N, A] and ONE block laid
        out [obs | reward | done] so that the results come back in a single DMA without staging copies."""
        import numpy as np
        block = torch.empty(n_action_buffers, self.n, self.action_dim, dtype=torch.float32).pin_memory()   # one allocation
        acts = [block[i] for i in range(n_action_buffers)]
        nbytes = self.n * (self.obs_dim * 4 + 4 + 1)
        self._pinned_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self._pinned_act = acts
        raw = self._pinned_out.numpy()
        o_end, r_end = self.n * self.obs_dim * 4, self.n * self.obs_dim * 4 + self.n * 4
        obs = raw[:o_end].view(np.float32).reshape(self.n, self.obs_dim)
        rew = raw[o_end:r_end].view(np.float32)
        done = raw[r_end:]
        return [a.numpy() for a in acts], obs, rew, done

    def rollout(self, steps, count=True):
        """`steps` env.steps in one launch with in-kernel U[-1,1] actions (action_space.sample(), README.md:48)."""
        N.check(self._L.pih_rollout(self._h, int(steps), self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                                    self.counters.data_ptr() if count else None, self._stream()), self._h, "pih_rollout")
        return self.obs, self.reward, self.done

    def stats(self, enable=True):
        """Execution counters (pih_set_stats): returns the zeroed int64[8] tensor the kernels add to; ``enable=False`` detaches."""
        if not enable:
            self._L.pih_set_stats(self._h, None)
            return None
        self._stats = torch.zeros(8, dtype=torch.int64, device=self.device)
        self._L.pih_set_stats(self._h, self._stats.data_ptr())
        return self._stats

    def ik_targets(self, action):
        a = torch.as_tensor(action, dtype=torch.float32).to(self.device).contiguous()
        out = torch.zeros(self.n, 9, dtype=torch.float32, device=self.device)
        N.check(self._L.pih_ik_targets(self._h, a.data_ptr(), out.data_ptr(), self._stream()), self._h, "pih_ik_targets")
        return out

    def ee_pose(self):
        pos = torch.zeros(self.n, 3, dtype=torch.float32, device=self.device)
        rot = torch.zeros(self.n, 9, dtype=torch.float32, device=self.device)
        N.check(self._L.pih_ee_pose(self._h, pos.data_ptr(), rot.data_ptr(), self._stream()), self._h, "pih_ee_pose")
        return pos, rot.view(self.n, 3, 3)

    def render(self, width=300, height=300, out=None):
        """PegInHole.render (peg_in_hole.py:276-304) for all envs: [N, H, W, 4] = (depth, r, g, b)."""
        if out is None:
            out = torch.empty(self.n, height, width, 4, dtype=torch.float32, device=self.device)
        N.check(self._L.pih_render(self._h, int(width), int(height), out.data_ptr(), self._stream()), self._h, "pih_render")
        return out

    def render_grasp(self, width=300, height=300, out=None):
        """Scripted task: the camera on the scene latched when the state machine entered state 2 - the reference's
        ``grasp_img`` (peg_in_hole.py:65-66), the image its grasp labels belong to."""
        if out is None:
            out = torch.empty(self.n, height, width, 4, dtype=torch.float32, device=self.device)
        N.check(self._L.pih_render_grasp(self._h, int(width), int(height), out.data_ptr(), self._stream()), self._h, "pih_render_grasp")
        return out

    def grasp_labels(self, width=300, height=300):
        """Label images of the shipped task's random_grasp (peg_in_hole.py:41-45,65-98): ``(labels[N, 4, H, W], params[N, 5])``
        = ([pos_img, sin_img, cos_img, wid_img], [x, y, angle in degrees, width, length])."""
        labels = torch.empty(self.n, 4, height, width, dtype=torch.float32, device=self.device)
        params = torch.empty(self.n, 5, dtype=torch.float32, device=self.device)
        N.check(self._L.pih_grasp_labels(self._h, int(width), int(height), labels.data_ptr(), params.data_ptr(), self._stream()),
                self._h, "pih_grasp_labels")
        return labels, params

    # -- state views (SoA words -> [N, k]); copies, for tests / checkpointing
    def words(self, first, count=1):
        flat = self.state.permute(1, 0, 2).reshape(self.n, -1)
        return flat[:, first:first + count]

    def int_word(self, word):
        return self.words(word, 1).contiguous().view(torch.int32)[:, 0]

    def env_offsets_device(self):
        """World offsets [N, 3] of all envs as a device tensor (BaseEnv._create_env, base_env.py:35-55), computed with
        torch ops from the closed form - env_offsets() loops over envs on the host, too slow for 10^5+ envs."""
        c = self.cfg
        sub = c.sub_num if c.sub_num > 0 else self.n
        e = (torch.arange(self.n, device=self.device, dtype=torch.int64) + int(c.env_id_base)) % sub
        ox, oy, oz = float(c.offset[0]), float(c.offset[1]), float(c.offset[2])
        ef = e.to(torch.float32)
        if ox == 0.0 or oy == 0.0:
            return torch.stack([ox * ef, oy * ef, oz * ef], 1)
        s = 0
        while s * s < sub:
            s += 1
        i, j = (e // s).to(torch.float32), (e % s).to(torch.float32)
        return torch.stack([ox * i, oy * j, torch.full_like(i, oz)], 1)

    def env_offsets(self):
        out = torch.zeros(self.n, 3)
        buf = (C.c_float * 3)()
        for i in range(self.n):
            self._L.pih_env_offset(C.byref(self.cfg), i, buf)
            out[i] = torch.tensor(list(buf))
        return out
