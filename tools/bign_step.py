import os, sys
sys.path.insert(0, os.getcwd())
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
n = 262144
sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
for i in range(30):
    a = torch.rand(n, 4, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    sim.step(a)
torch.cuda.synchronize()
