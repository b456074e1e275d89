#!/usr/bin/env python3
"""Phase timing of k_rollout_ws (last step of a multi-step launch) in CTA 0 (dev tool; -DPIH_PHASE_TICKS build)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200 import _native  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

L = _native.lib()
sim = BatchedSim(4096, arm="panda", task="peg-in-hole", offset=(2.0, 3.0, 0.0), max_steps=200, auto_reset=True, seed=0, device=0)
sim.reset()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
buf = (C.c_longlong * 16)()
rows = []
cta = (C.c_ulonglong * 2048)()
ctas = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for t in range(60):
    ev0.record(); sim.rollout(K, count=False); ev1.record(); torch.cuda.synchronize()
    tot += ev0.elapsed_time(ev1)
    L.pih_debug_ticks(buf); rows.append(list(buf))
    L.pih_debug_cta_ns(cta); ctas.append(np.array(list(cta)[:256], dtype=np.float64).reshape(-1, 2))
r = np.array(rows[10:], dtype=np.float64)
m = lambda a, b: float(np.median(r[:, b] - r[:, a])) / 1965.0
print(f"rollout: {1e3 * tot / (60 * K):.1f} us/step ({K}-step launches) | IK {m(0, 1):.1f} | D: go -> ABA {m(2, 12):.1f} own cols {m(12, 13):.1f} -> ready {m(13, 3):.1f} | setup {m(3, 4):.1f} PGS {m(4, 5):.1f} tail {m(5, 8):.1f} | C: front {m(9, 10):.1f} cols {m(10, 11):.1f} | D go->end {m(2, 8):.1f}")
c = np.stack(ctas[10:]); dur = (c[:, :, 1] - c[:, :, 0]) / 1e3
print(f"per-CTA last-step duration by globaltimer: CTA0 {np.median(dur[:, 0]):.1f} us, median {np.median(dur):.1f}, p90 {np.percentile(dur, 90):.1f}, max-per-launch median {np.median(dur.max(1)):.1f}; clock64 span CTA0 {m(2, 8):.1f} us at 1965 MHz -> SM clock {1965.0 * m(2, 8) / np.median(dur[:, 0]):.0f} MHz")
print("slowest CTAs (mean over launches):", np.argsort(-dur.mean(0))[:10].tolist(), np.sort(-dur.mean(0))[:10].round(1).tolist())
