#!/usr/bin/env python3
"""e2e timing of the host-buffer step (dev tool): pih_step_host (zero-copy, one launch + sync per step) against the
resident step server (pih_server_step), wall clock per step, fresh i.i.d. action buffer every step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.sim import BatchedSim, HostStepper  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for persistent in (False, True, False, True):
    sim = BatchedSim(n, arm="panda", task="peg-in-hole", offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
    sim.reset()
    offs = sim.env_offsets().numpy()
    nb = 64
    pas, pobs, prew, pdone = sim.pinned_host_buffers(nb)
    rng = np.random.default_rng(1)
    for a in pas:
        a[...] = rng.uniform(-1, 1, a.shape).astype(np.float32)
        a[:, :3] += offs
    hs = HostStepper(sim, pas, pobs, prew, pdone, persistent=persistent)
    for t in range(200):
        hs(t % nb)
    t0 = time.perf_counter()
    for t in range(steps):
        hs(t % nb)
    dt = time.perf_counter() - t0
    hs.close()
    chk = int(sim.state.view(torch.int32).to(torch.int64).sum().item())
    print(f"persistent={persistent}: {1e6 * dt / steps:.2f} us/step  {n * steps / dt / 1e6:.1f} M env-steps/s  checksum {chk} obs-sum {float(pobs.sum()):.4f}")
