"""Batched counterparts of peg_in_hole_gym/envs/utils.py (reference lines cited per function).

Robot/scene loading has no run-time equivalent here - the URDF constants are baked into csrc/pih_model_data.h -
so init_* only validate arguments and return ids; the arithmetic helpers keep the reference semantics."""
import math
import random

import numpy as np

from ..gym_shim import spaces


def test_mode(test_key, func):
    """envs/utils.py:11-17 - GUI hot-key hook; there is no GUI in the batched simulator, so this is a no-op."""
    return None


def data_normalize(data, normalize_range):
    """envs/utils.py:19-22 - in-place (x - lo) / (hi - lo) per component; returns the same object."""
    for i, (lo, hi) in enumerate(normalize_range[:len(data)]):
        data[i] = (data[i] - lo) / (hi - lo)
    return data


def init_table(client, table_pos, flags=0):
    """envs/utils.py:24-28 - the table is the static plane z = offset_z - 0.05 inside the kernels."""
    return 1


def init_panda(client, panda_pos, panda_orn, table_pos, flags=0):
    """envs/utils.py:30-38 - writes the joint angles (zero velocity) for every env of the simulator `client`."""
    reset_panda(client, 0, panda_orn)
    return 0, init_table(client, table_pos, flags)


def init_ur(client, ur_pos, ur_orn, table_pos, flags=0):
    """envs/utils.py:40-48."""
    reset_ur(client, 0, ur_orn)
    return 0, init_table(client, table_pos, flags)


def _write_joints(sim, values, first):
    import torch
    vals = torch.as_tensor(list(values), dtype=torch.float32, device=sim.device)
    flat = sim.state.permute(1, 0, 2).reshape(sim.n, -1).clone()
    k = min(len(vals), 9 - first)
    flat[:, first:first + k] = vals[:k]
    flat[:, 12 + first:12 + first + k] = 0.0  # resetJointState zeroes the velocity [3P]
    sim.state.copy_(flat.reshape(sim.n, -1, 4).permute(1, 0, 2))


def reset_panda(client, panda_id, panda_orn):
    """envs/utils.py:51-53.  The reference writes 9 numbers to joint indices 0..8; 7 and 8 are fixed joints in
    pybullet's panda, so only the first 7 reach a movable joint."""
    if hasattr(client, "state"):
        _write_joints(client, list(panda_orn)[:7], 0)


def reset_ur(client, ur_id, ur_orn):
    """envs/utils.py:55-57."""
    if hasattr(client, "state"):
        _write_joints(client, list(ur_orn)[:6], 0)


def panda_execute(client, panda_id, action, pandaEndEffectorIndex=11, pandaNumDofs=7, dv=2 / 240.):
    """envs/utils.py:60-68 - returns the joint position targets [N, 9] the motors are given (FK -> vel_constraint
    -> DLS IK on the GPU).  The motors themselves are applied inside BatchedSim.step."""
    return client.ik_targets(action)


def ur_execute(client, ur_id, action, urEndEffectorIndex=7, urNumDofs=6, dv=2 / 240.):
    """envs/utils.py:70-82."""
    return client.ik_targets(action)


def vel_constraint(cur, tar, dv):
    """envs/utils.py:85-95 - move from `cur` toward the absolute target `tar`, at most `dv` per axis."""
    return [c + max(-dv, min(dv, t - c)) for c, t in zip(cur, tar)]


def random_pos_in_panda_space():
    """envs/utils.py:97-107 - point on the radius-0.7 sphere centred (0, 0, 0.2), above the panda: x uniform,
    |y| = sqrt(U(0, r^2 - x^2)) - U(0, 0.4) with a random sign, resampled until it lies inside the disc."""
    radius = 0.7
    r2 = radius * radius
    while True:
        x = random.uniform(-radius, radius)
        y = (math.sqrt(random.uniform(0, r2 - x * x)) - random.uniform(0, 0.4)) * random.choice([-1, 1])
        rest = r2 - x * x - y * y
        if rest >= 0:
            return np.array([x, y, math.sqrt(rest) + 0.2])


class _AgentSpaces(list):
    """Common part of the four containers of envs/utils.py:132-209: a list of per-agent spaces that also exposes
    `.shape/.high/.low` of the first child and samples every child.  `_as_list=False` reproduces the reference's MP
    variants, which skip list.__init__ and therefore have len() == 0."""
    _child_type = spaces.space.Space
    _as_list = True

    def __init__(self, children):
        children = list(children)
        assert all(isinstance(c, self._child_type) for c in children)
        if self._as_list:
            super().__init__(children)
        self._children = children
        first = children[0]
        self.shape, self.high, self.low = first.shape, first.high, first.low

    def sample(self):
        return [c.sample() for c in self._children]

    def contains(self, items):
        return all(c.contains(x) for c, x in zip(self._children, items))


class MultiAgentObservationSpace(_AgentSpaces):
    """envs/utils.py:132-153."""

    @property
    def _agents_observation_space(self):
        return self._children


class MultiAgentActionSpace(_AgentSpaces):
    """envs/utils.py:156-170 (+ NEW sample_array: one [sub, A] float32 array instead of `sub` Python objects)."""

    @property
    def _agents_action_space(self):
        return self._children

    def sample_array(self):
        s0 = self._children[0]
        return s0._rng.uniform(s0.low, s0.high, (len(self._children),) + s0.shape).astype(np.float32)


class MPMultiAgentObservationSpace(MultiAgentObservationSpace):
    """envs/utils.py:173-193."""
    _child_type = MultiAgentObservationSpace
    _as_list = False


class MPMultiAgentActionSpace(MultiAgentActionSpace):
    """envs/utils.py:196-209 (+ NEW sample_array: one [mp, sub, A] float32 array)."""
    _child_type = MultiAgentActionSpace
    _as_list = False

    def sample_array(self):
        return np.stack([c.sample_array() for c in self._children])
