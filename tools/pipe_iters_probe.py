"""Dev tool: where the pipe step's time goes - fixed part vs per-solver-iteration part at several occupancies
(148 envs = one warp per SM, 592 = one per scheduler, 2368 = 16 per SM = full residency)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim

for n in (148, 592, 2368, 4096):
    res = []
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
        res.append(1e3 * e0.elapsed_time(e1) / 20)
        sim.close()
    print(f"n={n:5d}: solver_iters 0/10/50 -> {res[0]:7.1f} {res[1]:7.1f} {res[2]:7.1f} us/step; per sweep {(res[2] - res[0]) / 50:5.2f} us (falling phase: 34 motor/limit rows)", flush=True)
