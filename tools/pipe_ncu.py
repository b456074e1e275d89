"""Dev tool: a short pipe rollout for ncu (k_pipe_step / k_pipe_arm)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim
import torch
sim = BatchedSim(4096, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
sim.rollout(60, count=False)
torch.cuda.synchronize()
print("ok")
