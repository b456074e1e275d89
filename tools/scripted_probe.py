#!/usr/bin/env python3
"""Throughput of the scripted grasp->insert task (the task the reference ships: one step() = one 2226-engine-step
episode, peg_in_hole.py:39-119) - dev tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]
for n in sizes:
    sim = BatchedSim(n, task="peg-in-hole-scripted", sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, device=0)
    sim.reset()
    sim.rollout(50, count=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    blocks = []
    for b0 in range(0, 2226 - 50, 272):
        nsteps = min(272, 2226 - 50 - b0)
        torch.cuda.synchronize(); e0.record()
        sim.rollout(nsteps, count=False)
        e1.record(); torch.cuda.synchronize()
        blocks.append((b0 + 50, e0.elapsed_time(e1) / nsteps * 1e3))
    tot_ms = sum(us * min(272, 2226 - 50 - (b - 50)) for b, us in blocks) / 1e3
    succ = float((sim.int_word(11) & 2).ne(0).float().mean())
    print(f"n={n:6d}: " + " ".join(f"{b}:{us:.0f}us" for b, us in blocks))
    print(f"          episode (2176 steps) {tot_ms:.1f} ms -> {n / tot_ms * 1e3:.3e} episodes/s, {n * 2176 / tot_ms / 1e3:.1f} M env-steps/s, success {succ:.2f}")
    sim.close()
