"""peg_in_hole_gym_b200 - B200-native drop-in for peg-in-hole-gym's vectorised env.step.

Registers the reference's ids (peg_in_hole_gym/__init__.py:3-11) with the bundled gym shim and, when importable,
with a real gym / gymnasium:

    import peg_in_hole_gym_b200 as gym_b200
    env = gym_b200.make('peg-in-hole-mp-v0', client=gym_b200.DIRECT, task='peg-in-hole', mp_num=4, sub_num=1024,
                        offset=[2., 3., 0.], args=None)
"""
from .gym_shim import DIRECT, GUI, make, register, spaces  # noqa: F401

register(id='peg-in-hole-v0', entry_point='peg_in_hole_gym_b200.envs.base_env:BaseEnv')
register(id='peg-in-hole-mp-v0', entry_point='peg_in_hole_gym_b200.envs.base_env_mp:BaseEnvMp')

__all__ = ["make", "register", "spaces", "DIRECT", "GUI"]
