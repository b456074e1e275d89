"""Timing probe of the flexible-pipe kernels (dev tool): microseconds per env.step at several batch sizes, the share of
k_pipe_arm / k_pipe_step, executed solver iterations and contact counts."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim

for n in (1024, 4096, 16384):
    sim = BatchedSim(n, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
    sim.reset()
    sim.rollout(20, count=False)
    for phase, pre in (("falling", 0), ("settling", 60), ("resting", 100)):
        if pre:
            sim.rollout(pre, count=False)
        st = sim.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        sim.rollout(20, count=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        v = st.cpu().numpy().astype(float)
        sim.stats(False)
        print(f"n={n:6d} {phase:9s} {1e3 * ms / 20:9.1f} us/step  {n * 20 / ms * 1e3:12.3e} env-steps/s  contacts {v[4] / v[0]:5.1f}  "
              f"pgs iters {v[3] / v[0]:5.1f}  rows*iters {v[2] / v[0]:8.0f}", flush=True)
    sim.close()
