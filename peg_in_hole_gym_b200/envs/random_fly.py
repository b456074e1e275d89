"""'random-fly' task plugin.  The README (README.md:31-41) and BASELINE config 4 name this task, but its source is
NOT in the reference snapshot; only the assets (assets/urdf/banana.urdf,
Amicelli_800_tex.urdf) and random_pos_in_panda_space (envs/utils.py:97-107) survive.  Semantics here are NEW and
documented in DESIGN.md section 7: a 1 kg frictionless free body is thrown through the panda workspace, the arm is
driven by panda_execute, success = object within 5 cm of the end effector, episode ends on table impact.
``args = [object_name, timestep]`` as in README.md:38."""
import numpy as np

from ..gym_shim import spaces
from .meta_env import MetaEnv


class RandomFly(MetaEnv):
    action_space = spaces.Box(np.array([-1] * 4), np.array([1] * 4))
    observation_space = spaces.Box(np.array([-1] * 5), np.array([1] * 5))
    native_task = "random-fly"
    arm = "panda"

    def sim_config(self):
        args = self.args or ["Banana", 1.0 / 240.0]
        obj = str(args[0]).lower() if len(args) > 0 else "banana"
        if obj.startswith("amicelli"):
            obj = "amicelli"
        dt = float(args[1]) if len(args) > 1 else 1.0 / 240.0
        return dict(task="random-fly", arm=self.arm, fly_object=obj, dt=dt)
