"""BaseEnvMp - counterpart of peg_in_hole_gym/envs/base_env_mp.py:7-87.

The reference forks `mp_num` worker processes, each owning one pybullet client with `sub_num` sub-envs, and talks to
them through pickled [opcode, payload] messages over Queue(1) pairs (base_env_mp.py:27-51,75-87).  Here all
mp_num x sub_num sub-envs of this rank are stepped by ONE CUDA kernel launch; the opcodes survive as constants and
map to C-ABI calls (RESET/HARD_RESET -> pih_reset, STEP -> pih_step, CLOSE -> pih_destroy).  With torch.distributed
initialised, the mp axis is sharded over ranks (parallel.shard_clients) and the containers hold this rank's block."""
from .. import gym_shim as gym
from ..parallel import shard_clients, world
from .base_env import BaseEnv
from .utils import MPMultiAgentActionSpace, MPMultiAgentObservationSpace


class BaseEnvMp(BaseEnv):
    CLOSE = 0
    RESET = 1
    STEP = 2
    RENDER = 3
    HARD_RESET = 4

    def __init__(self, client, task, mp_num=1, sub_num=1, offset=[0, 0, 0], args=None, is_test=False, *, shard=True,
                 **kwargs):
        if client == gym.GUI:
            print("Warning: GUI mode is not supported in this env.")  # base_env_mp.py:16-19
            print("Auto switch to DIRECT mode.")
            client = gym.DIRECT
        self.mp_num_global = mp_num
        rank, ws = world() if shard else (0, 1)
        lo, hi = shard_clients(mp_num, rank, ws)
        self.mp_lo, self.mp_hi = lo, hi
        self.mp_num = hi - lo          # clients owned by this rank
        self.sub_num = sub_num
        if self.mp_num <= 0:
            raise ValueError(f"mp_num={mp_num} clients cannot be sharded over {ws} ranks: rank {rank} would own none "
                             "(use mp_num >= world size, or shard=False for a full copy per rank)")
        super().__init__(client, task, sub_num, offset, args, is_test, _mp_num=self.mp_num,
                         _env_id_base=lo * sub_num, **kwargs)
        self.clients = [self] * self.mp_num
        base_a, base_o = self.action_space, self.observation_space
        self.action_space = MPMultiAgentActionSpace([base_a for _ in range(self.mp_num)])
        self.observation_space = MPMultiAgentObservationSpace([base_o for _ in range(self.mp_num)])

    def _shape(self):
        return (self.mp_num, self.sub_num)

    # render(mode): the reference only posts RENDER to the workers and discards the images (base_env_mp.py:66-68); here
    # the inherited BaseEnv.render returns them as a [mp, sub, H, W, 4] tensor.
