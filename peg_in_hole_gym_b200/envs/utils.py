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
    """envs/utils.py:19-22 - in-place (x-lo)/(hi-lo) per component."""
    for i in range(len(data)):
        data[i] = (data[i] - normalize_range[i][0]) / (normalize_range[i][1] - normalize_range[i][0])
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
    """envs/utils.py:85-95 - per-axis clamp of the step toward an absolute target."""
    res = []
    for i in range(len(tar)):
        diff = tar[i] - cur[i]
        if abs(diff) > dv:
            re = cur[i] + (dv if diff > 0 else -dv)
        else:
            re = cur[i] + diff
        res.append(re)
    return res


def random_pos_in_panda_space():
    """envs/utils.py:97-107 - rejection-sample a point of the radius-0.7 sphere cap above the panda."""
    x = y = z = 1
    length = 0.7
    while (length * length - x * x - y * y) < 0:
        x = random.uniform(-length, length)
        y = (math.sqrt(random.uniform(0, length * length - x * x)) - random.uniform(0, 0.4)) * random.choice([-1, 1])
    z = math.sqrt(length * length - x * x - y * y) + 0.2
    return np.array([x, y, z])


class MultiAgentObservationSpace(list):
    """envs/utils.py:132-153."""

    def __init__(self, agents_observation_space):
        for x in agents_observation_space:
            assert isinstance(x, spaces.space.Space)
        super().__init__(agents_observation_space)
        self._agents_observation_space = agents_observation_space
        self.shape = agents_observation_space[0].shape
        self.high = agents_observation_space[0].high
        self.low = agents_observation_space[0].low

    def sample(self):
        return [s.sample() for s in self._agents_observation_space]

    def contains(self, obs):
        for space, ob in zip(self._agents_observation_space, obs):
            if not space.contains(ob):
                return False
        return True


class MultiAgentActionSpace(list):
    """envs/utils.py:156-170."""

    def __init__(self, agents_action_space):
        for x in agents_action_space:
            assert isinstance(x, spaces.space.Space)
        super().__init__(agents_action_space)
        self._agents_action_space = agents_action_space
        self.shape = agents_action_space[0].shape
        self.high = agents_action_space[0].high
        self.low = agents_action_space[0].low

    def sample(self):
        return [s.sample() for s in self._agents_action_space]

    def sample_array(self):
        """NEW: one [sub, A] float32 array (avoids sub_num Python objects per step)."""
        s0 = self._agents_action_space[0]
        return s0._rng.uniform(s0.low, s0.high, (len(self._agents_action_space),) + s0.shape).astype(np.float32)


class MPMultiAgentObservationSpace(MultiAgentObservationSpace):
    """envs/utils.py:173-193 - like the reference it does NOT call list.__init__, so len() == 0 (Appendix A.9)."""

    def __init__(self, agents_observation_space):
        for x in agents_observation_space:
            assert isinstance(x, MultiAgentObservationSpace)
        self._agents_observation_space = agents_observation_space
        self.shape = agents_observation_space[0].shape
        self.high = agents_observation_space[0].high
        self.low = agents_observation_space[0].low


class MPMultiAgentActionSpace(MultiAgentActionSpace):
    """envs/utils.py:196-209."""

    def __init__(self, agents_action_space):
        for x in agents_action_space:
            assert isinstance(x, MultiAgentActionSpace)
        self._agents_action_space = agents_action_space
        self.shape = agents_action_space[0].shape
        self.high = agents_action_space[0].high
        self.low = agents_action_space[0].low

    def sample_array(self):
        """NEW: one [mp, sub, A] float32 array."""
        return np.stack([s.sample_array() for s in self._agents_action_space])
