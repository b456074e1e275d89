"""Batched task-plugin base, the counterpart of peg_in_hole_gym/envs/meta_env.py:8-42.

In the reference one MetaEnv instance is ONE sub-environment inside a pybullet world.  Here one instance describes
ALL sub-environments of a simulator: the physics (apply_action's IK + motors, stepSimulation, the built-in
observation / reward / done) runs inside the CUDA step kernel selected by ``native_task`` / ``arm``; the hooks
operate on whole-batch tensors.  Hook names and order are the reference's:

    apply_action(action)  -> called before the native step with the [N, A] action tensor; returns the tensor to apply
    get_info()            -> called after the native step; returns (obs[N,5], reward[N], done[N], info)
    reset(hard_reset)     -> _load_models() on hard reset, always _reset_internals()

Register a custom task by inserting the class into ``TASK_LIST`` (envs/base_env.py:9-11).
"""
import numpy as np

from ..gym_shim import spaces
from . import utils as U


class MetaEnv(object):
    action_space = spaces.Box(np.array([-1]), np.array([1]))
    observation_space = spaces.Box(np.array([-1]), np.array([1]))
    native_task = "peg-in-hole"   # which in-kernel task the plugin rides on
    arm = "panda"                 # 'panda' (envs/utils.py:30-38) or 'ur5' (envs/utils.py:40-48)

    def __init__(self, client, offset, args=[]):
        self.p = client           # BatchedSim (set by the env once the state exists), the reference's pybullet client
        self.offset = np.array(offset)
        self.args = args
        self._load_models()
        self._reset_internals()
        self.done = False

    # ---- configuration consumed by BaseEnv / BaseEnvMp when they create the simulator
    def sim_config(self):
        return {}

    def bind(self, sim):
        self.p = sim

    # ---- reference hooks (meta_env.py:19-42)
    def _load_models(self):
        pass

    def _reset_internals(self):
        pass

    def apply_action(self, action):
        return action

    def native_step(self, action):
        """One engine step of every sub-env (BaseEnv.step's single stepSimulation, base_env.py:64)."""
        self.p.step(action)

    def get_info(self):
        return self.p.obs, self.p.reward, self.p.done, None

    def reset(self, hard_reset=False):
        if hard_reset:
            self._load_models()
        self._reset_internals()
        self.done = False

    def render(self, mode="rgb_array", width=300, height=300):
        """peg_in_hole.py:276-304: end-effector camera, [N, 300, 300, 4] = (depth, r, g, b)."""
        return self.p.render(width, height)

    def test_mode(self, test_key, func):
        return None

    # ---- arm helpers as methods (meta_env.py:52-117); the broken `vel_constraint` signature of the reference
    #      (meta_env.py:107) is fixed by delegating to utils
    def init_table(self, table_pos=[0.0, -0.5, -1.3], table_orn=None, scale=2, flags=0):
        self.table_id = U.init_table(self.p, table_pos, flags)

    def init_panda(self, panda_pos, panda_orn, table_pos, flags=0):
        self.panda_id, self.table_id = U.init_panda(self.p, panda_pos, panda_orn, table_pos, flags)

    def init_ur(self, ur_pos, ur_orn, table_pos, flags=0):
        self.ur_id, self.table_id = U.init_ur(self.p, ur_pos, ur_orn, table_pos, flags)

    def reset_panda(self, panda_orn):
        U.reset_panda(self.p, 0, panda_orn)

    def reset_ur(self, ur_orn):
        U.reset_ur(self.p, 0, ur_orn)

    def panda_execute(self, action, pandaEndEffectorIndex=11, pandaNumDofs=7, dv=2 / 240.):
        return U.panda_execute(self.p, 0, action, pandaEndEffectorIndex, pandaNumDofs, dv)

    def ur_execute(self, action, urEndEffectorIndex=7, urNumDofs=6, dv=2 / 240.):
        return U.ur_execute(self.p, 0, action, urEndEffectorIndex, urNumDofs, dv)

    @staticmethod
    def vel_constraint(cur, tar, dv):
        return U.vel_constraint(cur, tar, dv)
