#!/usr/bin/env python3
"""Large batches: fused k_step vs split k_apply_action + k_step_dyn (dev tool): us/step of per-step launches, i.i.d. actions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

arm = sys.argv[1] if len(sys.argv) > 1 else "panda"
adim = 4 if arm == "panda" else 6
for n in [int(v) for v in os.environ.get('SIZES', '24576,32768,65536,131072,262144').split(',')]:
    res = []
    for split in ("0", "1"):
        os.environ["PIH_SPLIT"] = split
        sim = BatchedSim(n, arm=arm, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
        sim.reset()
        offs = sim.env_offsets().cuda()
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        acts = []
        for _ in range(60):
            a = torch.rand(n, adim, device="cuda", generator=g) * 2 - 1
            a[:, :3] += offs
            acts.append(a)
        for i in range(20):
            sim.step(acts[i])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(40):
            sim.step(acts[20 + i])
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 40 * 1e3)
        sim.close()
    print(f"{arm} n={n:7d}  fused {res[0]:8.1f} us  split {res[1]:8.1f} us  -> {n / min(res):6.1f} M env-steps/s", flush=True)
