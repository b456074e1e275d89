"""Dev tool (VERDICT r1 item 5 evidence): what bounds the 4096-env step.  The small-batch kernel runs apply_action (IK)
in warp 0 and the dynamics prologue in warp 1 of a CTA, then the solver.  Sweeping ik_iters / solver_iters (no early exit)
separates the three: step time vs IK iterations is flat until the IK warp becomes the longer one."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim

n = 4096
g = torch.Generator(device="cuda"); g.manual_seed(0)


def measure(ik, pgs, residual):
    sim = BatchedSim(n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0, ik_iters=ik, solver_iters=pgs, solver_residual=residual)
    sim.reset()
    off = torch.zeros(n, 4, device="cuda"); off[:, :3] = sim.env_offsets_device()
    acts = [torch.rand(n, 4, device="cuda", generator=g) * 2 - 1 + off for _ in range(32)]
    for i in range(60):
        sim.step(acts[i % 32])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(400):
        sim.step(acts[i % 32])
    e1.record()
    torch.cuda.synchronize()
    sim.close()
    return 1e3 * e0.elapsed_time(e1) / 400


print("ik_iters  solver_iters  us/step   (4096 panda envs, k_step_ws, solver_residual 0 = no early exit unless stated)")
for ik in (0, 5, 10, 15, 20):
    print(f"{ik:8d}  {50:12d}  {measure(ik, 50, 0.0):7.2f}", flush=True)
for pgs in (0, 10, 25):
    print(f"{20:8d}  {pgs:12d}  {measure(20, pgs, 0.0):7.2f}", flush=True)
print(f"{20:8d}  {'50 (1e-7)':>12s}  {measure(20, 50, 1e-7):7.2f}   <- the configuration of the bench", flush=True)
print(f"{0:8d}  {0:12d}  {measure(0, 0, 0.0):7.2f}   <- FK, ABA, M^-1, narrow phase, row set-up, integration, get_info only", flush=True)
