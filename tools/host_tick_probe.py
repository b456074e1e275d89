#!/usr/bin/env python3
"""Where the end-to-end step spends its time (dev tool; -DPIH_PHASE_TICKS build): pih_step_host on page-locked buffers
(zero-copy: the action warp reads its rows from host memory over PCIe) with the in-kernel phase clocks of CTA 0 and the
per-CTA durations, against the wall clock of the call."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200 import _native  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim, HostStepper  # noqa: E402

n = 4096
L = _native.lib()
sim = BatchedSim(n, arm="panda", task="peg-in-hole", offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
sim.reset()
offs = sim.env_offsets().numpy()
NB = 450
pas, pobs, prew, pdone = sim.pinned_host_buffers(NB)   # a fresh i.i.d. batch every step (a periodic sequence drives the arms into the table)
rng = np.random.default_rng(1)
for a in pas:
    a[...] = rng.uniform(-1, 1, a.shape).astype(np.float32)
    a[:, :3] += offs
hs = HostStepper(sim, pas, pobs, prew, pdone)
buf = (C.c_longlong * 16)()
cta = (C.c_ulonglong * 2048)()
rows, walls, ctas = [], [], []
for t in range(400):
    t0 = time.perf_counter()
    hs(t)
    walls.append(time.perf_counter() - t0)
    L.pih_debug_ticks(buf); rows.append(list(buf))
    L.pih_debug_cta_ns(cta); ctas.append(np.array(list(cta)[:256], dtype=np.float64).reshape(-1, 2))
r = np.array(rows[50:], dtype=np.float64)
m = lambda a, b: float(np.median(r[:, b] - r[:, a])) / 1965.0
c = np.stack(ctas[50:]); dur = (c[:, :, 1] - c[:, :, 0]) / 1e3
span = (c[:, :, 1].max(1) - c[:, :, 0].min(1)) / 1e3
print(f"zero-copy host step: wall {1e6 * np.median(walls[50:]):.1f} us | CTA0: IK {m(0, 1):.1f}  D pre-barrier {m(2, 3):.1f}  wait+setup {m(3, 4):.1f}  PGS {m(4, 5):.1f}  tail {m(5, 7):.1f}  stores {m(7, 8):.1f}  total {m(2, 8):.1f} | "
      f"per-CTA (state loaded -> step done) median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {np.median(dur.max(1)):.1f}; first start -> last end {np.median(span):.1f}")
