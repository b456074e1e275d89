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
    peg_model = "rigid"      # 'rigid': one end segment of pipe.urdf; 'pipe': the whole 25-link chain (peg_in_hole.py:240-245)

    def __init__(self, client, offset=[0, 0, 0], args=None):
        self.pandaUid = 0
        self.pandaEndEffectorIndex = 11   # peg_in_hole.py:20
        self.pandaNumDofs = 7             # peg_in_hole.py:21
        super().__init__(client, offset, args)

    def sim_config(self):
        return dict(task="peg-in-hole", arm=self.arm, dt=1.0 / 240.0, peg_model=self.peg_model)  # timeStep peg_in_hole.py:254


class PegInHolePipe(PegInHole):
    """The action-driven task on the object the reference loads: pipe.urdf, 25 links / 23 continuous joints with random
    bends, grasp link 0 or 23, success on that link (peg_in_hole.py:239-245,266,114-117).  Warp-per-env kernels."""
    peg_model = "pipe"


class PegInHoleUR5(PegInHole):
    """ur5 arm variant (envs/utils.py:40-48,70-82): 6-D action = absolute position + Euler orientation (NEW Box)."""
    action_space = spaces.Box(np.array([-1] * 6), np.array([1] * 6))
    arm = "ur5"


class PegInHoleScripted(PegInHole):
    """The task exactly as the reference ships it (peg_in_hole.py:30-37): ``apply_action`` is a no-op and one
    ``step()`` runs the whole scripted grasp -> lift -> insert -> release episode (10 timed states, 2226 engine
    steps of 1/240 s) and returns the success reward of peg_in_hole.py:114-117.  The state machine runs inside the
    rollout kernel; the observation is the 5-vector; the camera image and the grasp label images of peg_in_hole.py:66-98
    come from ``render()`` and ``random_grasp_labels()``.
    The grasp constraint of peg_in_hole.py:99-104 is built as its evident intent (grasp point pinned to the end effector
    while the state is 4..6; csrc/pih_constants.h PIH_PAIR_GRASP).  `PegInHoleScriptedPipe` runs it on the 25-link pipe."""
    native_task = "peg-in-hole-scripted"
    EPISODE_STEPS = 2226

    def sim_config(self):
        return dict(task="peg-in-hole-scripted", arm="panda", dt=1.0 / 240.0, peg_model=self.peg_model)

    def render(self, mode="rgb_array", width=300, height=300):
        """The observation of the shipped task: ``grasp_img``, the camera image taken when the state machine entered
        state 2 (peg_in_hole.py:33-37,65-66) - the scene the grasp labels describe, not the scene at episode end."""
        return self.p.render_grasp(width, height)

    def apply_action(self, action):      # peg_in_hole.py:30-31
        return action

    def native_step(self, action):       # get_info -> random_grasp(): `while not done` loop (peg_in_hole.py:52-112)
        self.p.rollout(self.EPISODE_STEPS, count=False)

    def grasp_image(self, width=300, height=300):
        return self.p.render_grasp(width, height)

    def random_grasp_labels(self, width=300, height=300):
        """What the reference's ``random_grasp`` returns beside the reward (peg_in_hole.py:119):
        ``[[pos_img, sin_img, cos_img, wid_img], [x, y, angle_deg, width, length]]`` for every env, as tensors
        ``labels[N, 4, H, W]`` and ``params[N, 5]`` (the rectangle was captured when the state machine entered state 2)."""
        return self.p.grasp_labels(width, height)


class PegInHoleScriptedPipe(PegInHoleScripted):
    """The shipped task on the shipped object: scripted grasp -> insert episode with the 25-link pipe."""
    peg_model = "pipe"
