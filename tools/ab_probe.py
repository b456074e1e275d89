#!/usr/bin/env python3
"""A/B timing of the step kernel for experiment builds (dev tool): PIH_LIB=tools/_build/libpih_X.so python tools/ab_probe.py
Back-to-back launches (no sync inside, so programmatic dependent launch overlaps as in bench.py), i.i.d. actions, 4096
panda envs with 200-step episodes + auto-reset; prints the CUDA-event mean per step over whole episodes and a checksum
of the final state (identical builds must print identical checksums)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sim = BatchedSim(n, arm="panda", task="peg-in-hole", offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = []
for _ in range(200):
    a = torch.rand(n, 4, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    acts.append(a.contiguous())
for a in acts:
    sim.step(a)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for _ in range(reps):
    ev0.record()
    for a in acts:
        sim.step(a)
    ev1.record()
    torch.cuda.synchronize()
    res.append(1e3 * ev0.elapsed_time(ev1) / len(acts))
chk = int(sim.state.view(torch.int32).to(torch.int64).sum().item())
print(f"{os.environ.get('PIH_LIB', 'default')}: n={n} us/step " + " ".join(f"{r:.2f}" for r in res) + f"  min {min(res):.2f}  checksum {chk}")
