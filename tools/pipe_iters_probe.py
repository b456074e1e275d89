"""Dev tool: where the pipe step's time goes - fixed part vs per-solver-iteration part, single wave (740 envs)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim

for n in (740, 4096):
    for iters in (0, 10, 50):
        sim = BatchedSim(n, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0, solver_iters=iters,
                         solver_residual=0.0)
        sim.reset()
        sim.rollout(10, count=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        sim.rollout(20, count=False)
        e1.record()
        torch.cuda.synchronize()
        print(f"n={n} solver_iters={iters:3d}: {1e3 * e0.elapsed_time(e1) / 20:8.1f} us/step", flush=True)
        sim.close()
