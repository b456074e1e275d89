#!/usr/bin/env python3
"""Long-horizon agreement check (dev tool): GPU rollout vs the CPU oracle over many auto-reset episodes, bit for bit."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for arm, task, n in (("panda", "peg-in-hole", 96), ("ur5", "peg-in-hole", 64), ("panda", "random-fly", 64)):
    kw = dict(offset=(2.0, 3.0, 0.0), seed=77, max_steps=200, auto_reset=True)
    dt = 1.0 / 120.0 if task == "random-fly" else 1.0 / 240.0
    sim = BatchedSim(n, arm=arm, task=task, dt=dt, device=0, **kw)
    orc = ol.Oracle(ol.make_config(n, arm={"panda": 0, "ur5": 1}[arm], task={"peg-in-hole": 0, "random-fly": 1}[task], dt=dt,
                                   offset=kw["offset"], seed=77, max_steps=200, auto_reset=1))
    sim.reset(); orc.reset()
    t0 = time.time()
    done_steps = 0
    while done_steps < steps:
        k = min(2500, steps - done_steps)
        sim.rollout(k, count=False); orc.rollout(k)
        done_steps += k
        sg, sc = sim.state.cpu().numpy(), orc.get_state()
        same = np.array_equal(sg.view(np.int32), sc.view(np.int32))
        finite = np.isfinite(sg).all()
        if not (same and finite):
            bad = np.argwhere(sg.view(np.int32) != sc.view(np.int32))
            print(f"{arm} {task}: MISMATCH after {done_steps} steps at {bad[:5].tolist()} finite={finite}")
            break
    else:
        ep = int(sim.int_word(10).float().mean().item())
        print(f"{arm} {task}: {n} envs x {steps} steps bit-identical to the oracle ({ep} episodes per env, {time.time() - t0:.0f} s)")
    sim.close()
