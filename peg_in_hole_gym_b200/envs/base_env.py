"""BaseEnv - counterpart of peg_in_hole_gym/envs/base_env.py:13-99 (one pybullet client, `task_num` sub-envs).

The per-sub-env Python loops and the single stepSimulation (base_env.py:60-71) become ONE kernel launch over all
sub-envs; env placement follows _create_env (base_env.py:35-55); done envs are frozen until reset (base_env.py:61-71)
unless the NEW auto_reset flag is set."""
import numpy as np
import torch

from .. import gym_shim as gym
from ..sim import BatchedSim
from .peg_in_hole import PegInHole, PegInHolePipe, PegInHoleScripted, PegInHoleScriptedPipe, PegInHoleUR5
from .random_fly import RandomFly
from .utils import MultiAgentActionSpace, MultiAgentObservationSpace, test_mode

TASK_LIST = {
    'peg-in-hole': PegInHole,      # base_env.py:9-11
    'peg-in-hole-ur5': PegInHoleUR5,
    'peg-in-hole-scripted': PegInHoleScripted,   # the scripted grasp FSM the reference ships (peg_in_hole.py:39-225)
    'peg-in-hole-pipe': PegInHolePipe,           # action-driven, the reference's 25-link pipe.urdf as the peg
    'peg-in-hole-scripted-pipe': PegInHoleScriptedPipe,   # the shipped task on the shipped object
    'random-fly': RandomFly,       # README.md:38 (source missing from the snapshot; NEW semantics)
}


class LazyInfos:
    """`infos` of the reference is a list of dicts (base_env.py:90); materialising N dicts (or even reading the flags
    word back) per step would cap throughput, so nothing is computed until an element is
    accessed: then the flags word is read once and dicts are built on demand.  The word is SNAPSHOT on the device at the
    step that produced it (one [N] int32 copy, no D2H), so infos kept from step t still describe step t after later
    steps, resets or auto-resets - the reference's infos are per-step snapshots too.  Indexes like the reference's
    nested lists.  `success` is true for the step in which an episode ended in success, also under auto_reset (where the
    env has already restarted: PIH_FLAG_PREV_SUCCESS); `fault` reports the non-finite guard."""

    def __init__(self, shape, flags_dev=None, flags=None):
        self._shape, self._dev, self._flags = tuple(shape), flags_dev, flags

    def _load(self):
        if self._flags is None:
            self._flags = self._dev.cpu().numpy()
        return self._flags

    def __len__(self):
        return self._shape[0]

    def __getitem__(self, i):
        flags = self._load()
        if len(self._shape) > 1:
            n = int(np.prod(self._shape[1:]))
            return LazyInfos(self._shape[1:], flags=flags[i * n:(i + 1) * n])
        f = int(flags[i])
        return {"success": bool(f & (2 | 8)), "contact_pairs": (f >> 8) & 0xff, "fault": bool(f & 4)}

    def tolist(self):
        return [self[i].tolist() if len(self._shape) > 1 else self[i] for i in range(self._shape[0])]


def _to_action_tensor(action, n, adim, device):
    """Accept the reference's nested lists of arrays, a numpy array or a torch tensor -> float32 CUDA [n, A]."""
    if isinstance(action, torch.Tensor):
        a = action
    else:
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(action, dtype=np.float32)))
    a = a.to(device=device, dtype=torch.float32, non_blocking=True).reshape(n, adim)
    return a.contiguous()


class BaseEnv(gym.Env):
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, client, task='peg-in-hole', task_num=1, offset=[0, 0, 0], args=None, is_test=False, *,
                 arm=None, seed=0, max_steps=0, auto_reset=False, action_frame='world', solver_iters=50, ik_iters=20,
                 device=None, as_lists=False, peg_model=None, _mp_num=1, _env_id_base=0):
        assert task in TASK_LIST, "Please regisiter your custom env first!"
        assert (task_num == 1 or (task_num > 1 and list(offset) != [0, 0, 0])), "Offset is in valid."
        self.task, self.task_num, self.offset, self.args = task, task_num, offset, args
        self.sub_env = TASK_LIST[self.task]
        self.p = None
        self.as_lists = as_lists
        self._mp = _mp_num
        self.plugin = self.sub_env(None, offset, args)
        if arm is not None:
            self.plugin.arm = arm
        if peg_model is not None:
            self.plugin.peg_model = peg_model
        cfg = dict(task=self.plugin.native_task, arm=self.plugin.arm, dt=1.0 / 240.0)
        cfg.update(self.plugin.sim_config())
        cfg["arm"] = self.plugin.arm
        self.sim = BatchedSim(self._mp * task_num, sub_num=task_num, env_id_base=_env_id_base, offset=offset, seed=seed,
                              max_steps=max_steps, auto_reset=auto_reset, action_frame_env=(action_frame == 'env'),
                              solver_iters=solver_iters, ik_iters=ik_iters, device=device, **cfg)
        self.p = self.sim
        self.plugin.bind(self.sim)
        self.sub_envs = [self.plugin]  # one batched plugin object stands for all sub-envs
        space_a, space_o = self.sub_env.action_space, self.sub_env.observation_space
        self.action_space = MultiAgentActionSpace([space_a for _ in range(task_num)])
        self.observation_space = MultiAgentObservationSpace([space_o for _ in range(task_num)])
        self.is_test = is_test
        self._was_reset = False

    # shapes of the returned containers: [sub] here, [mp][sub] in BaseEnvMp
    def _shape(self):
        return (self.task_num,)

    def _pack(self, obs, reward, done):
        shp = self._shape()
        obs, reward, done = obs.view(*shp, -1), reward.view(*shp), done.view(*shp).bool()
        infos = LazyInfos(shp, flags_dev=self.sim.state[2, :, 3].clone().view(torch.int32))   # word 11 = flags
        if self.as_lists:
            return obs.cpu().numpy().tolist(), reward.cpu().numpy().tolist(), done.cpu().numpy().tolist(), infos.tolist()
        return obs, reward, done, infos

    def step(self, action):
        if not self._was_reset:
            raise AttributeError("step() called before reset()")  # reference: no self.dones yet (base_env.py:62)
        a = _to_action_tensor(action, self.sim.n, self.sim.action_dim, self.sim.device)
        a = self.plugin.apply_action(a)
        self.plugin.native_step(a)
        obs, reward, done, _ = self.plugin.get_info()
        if self.is_test:
            test_mode('r', self.reset)
        return self._pack(obs, reward, done)

    def render(self, mode='rgb_array', width=300, height=300):
        img = self.plugin.render(mode, width, height)
        return img.view(*self._shape(), height, width, 4)

    def reset(self, hard_reset=False):
        self.plugin.reset(hard_reset)
        obs = self.sim.reset()
        self._was_reset = True
        obs = obs.view(*self._shape(), -1)
        return obs.cpu().numpy().tolist() if self.as_lists else obs

    def close(self):
        self.sim.close()
