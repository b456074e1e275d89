"""Dev tool for `ncu --metrics gpu__time_duration.sum`: pipe steps at a given batch and solver_iters (no early exit)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim
import torch
n, iters = int(sys.argv[1]), int(sys.argv[2])
sim = BatchedSim(n, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0, solver_iters=iters, solver_residual=0.0)
sim.reset()
sim.rollout(16, count=False)
torch.cuda.synchronize()
print("ok")
