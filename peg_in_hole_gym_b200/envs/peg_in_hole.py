"""'peg-in-hole' task plugin - counterpart of peg_in_hole_gym/envs/peg_in_hole.py.

Spaces are the declared ones (peg_in_hole.py:12-13): action Box(-1,1,(4,)) = end-effector xyz + finger,
observation Box(-1,1,(5,)) = [finger1, finger2, EE x, y, z].  The per-step path is the one the reference's helpers
define (panda_execute, envs/utils.py:60-68) with the success reward of peg_in_hole.py:114-117; reset reproduces the
distributions of peg_in_hole.py:230-251.  The scripted grasp FSM (peg_in_hole.py:39-225) and the camera
(peg_in_hole.py:276-304) are SURVEY.md section 8f 'next' rows, not part of env.step."""
import numpy as np

from ..gym_shim import spaces
from .meta_env import MetaEnv


class PegInHole(MetaEnv):
    action_space = spaces.Box(np.array([-1] * 4), np.array([1] * 4))
    observation_space = spaces.Box(np.array([-1] * 5), np.array([1] * 5))
    native_task = "peg-in-hole"
    arm = "panda"

    def __init__(self, client, offset=[0, 0, 0], args=None):
        self.pandaUid = 0
        self.pandaEndEffectorIndex = 11   # peg_in_hole.py:20
        self.pandaNumDofs = 7             # peg_in_hole.py:21
        super().__init__(client, offset, args)

    def sim_config(self):
        return dict(task="peg-in-hole", arm=self.arm, dt=1.0 / 240.0)  # timeStep peg_in_hole.py:254


class PegInHoleUR5(PegInHole):
    """ur5 arm variant (envs/utils.py:40-48,70-82): 6-D action = absolute position + Euler orientation (NEW Box)."""
    action_space = spaces.Box(np.array([-1] * 6), np.array([1] * 6))
    arm = "ur5"
