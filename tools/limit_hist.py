#!/usr/bin/env python3
"""Dev tool: which arm joints sit at a limit, per warp of 32 envs, over the headline workload (4096 panda envs, i.i.d.
actions, 200-step episodes).  The solver picks its sweep instantiation per warp from that mask (none / elbow only /
anything else -> general sweep), and a step lasts as long as its slowest CTA."""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

LO = np.array([-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671])
HI = np.array([2.9671, 1.8326, 2.9671, 0.0, 2.9671, 3.8223, 2.9671])
n = 4096
sim = BatchedSim(n, arm="panda", task="peg-in-hole", offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
hist = collections.Counter()
steps_with = collections.Counter()
for t in range(400):
    a = torch.rand(n, 4, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    sim.step(a.contiguous())
    q = sim.state.cpu().numpy().transpose(1, 0, 2).reshape(n, -1)[:, 0:7]
    at = ((q - LO) < 0.01) | ((HI - q) < 0.01)                      # [n, 7] (approximation of the kernel's row rule)
    masks = at.reshape(n // 32, 32, 7).any(1)                         # per warp
    keys = [tuple(np.nonzero(m)[0].tolist()) for m in masks]
    hist.update(keys)
    kinds = set("none" if k == () else ("elbow" if k == (3,) else "general") for k in keys)
    steps_with.update(kinds)
tot = sum(hist.values())
print("per-warp masks of arm joints at a limit (fraction of warp-steps):")
for k, v in hist.most_common(12):
    print(f"  {k}: {v / tot:.4f}")
print("fraction of steps in which some warp needs: ", {k: v / 400 for k, v in steps_with.items()})
