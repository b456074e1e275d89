#!/usr/bin/env python3
"""Device time per step as a function of the episode phase on ONE batch (dev tool)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200 import _native  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = _native.lib()
sim = BatchedSim(n, offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
ring = []
for _ in range(64):
    a = torch.rand(n, 4, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    ring.append(a.contiguous())
for i in range(200):
    sim.step(ring[(i * 7) % 64])
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
for blk in range(16):
    torch.cuda.synchronize()
    ev0.record()
    for i in range(25):
        sim.step(ring[((blk * 25 + i) * 7) % 64])
    ev1.record()
    torch.cuda.synchronize()
    out.append(1e3 * ev0.elapsed_time(ev1) / 25)
print("us/step per 25-step block (episode = 200 steps):", " ".join(f"{v:.1f}" for v in out))
