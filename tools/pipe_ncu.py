"""Dev tool: a short pipe rollout for ncu (k_pipe_step / k_pipe_arm).  usage: pipe_ncu.py [envs] [steps]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sim = BatchedSim(n, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
sim.rollout(steps, count=False)
torch.cuda.synchronize()
print("ok")
