"""Dev tool: the scripted task in its lift phase (state 4: grasp constraint + finger contacts on every env) for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
sim = BatchedSim(4096, task="peg-in-hole-scripted", sub_num=4096, offset=(2.0, 3.0, 0.0), seed=0, device=0)
sim.reset()
sim.rollout(1340, count=False)
for _ in range(6):
    sim.rollout(10, count=False)
torch.cuda.synchronize()
print("ok")
