#!/usr/bin/env python3
"""Phase timing of the resident step server in CTA 0 (dev tool; experiment build with -DPIH_PHASE_TICKS)."""
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
pas, pobs, prew, pdone = sim.pinned_host_buffers(16)
rng = np.random.default_rng(1)
for a in pas:
    a[...] = rng.uniform(-1, 1, a.shape).astype(np.float32)
    a[:, :3] += offs
hs = HostStepper(sim, pas, pobs, prew, pdone, persistent=True, idle_timeout_ms=2000)
buf = (C.c_longlong * 16)()
stamps = (C.c_ulonglong * 2)()
L.pih_debug_server_stamps.argtypes = [C.c_void_p, C.c_void_p]
rows, walls, spans = [], [], []
cta = (C.c_ulonglong * 2048)()
ctas = []
for t in range(300):
    t0 = time.perf_counter()
    hs(t % 16)
    walls.append(time.perf_counter() - t0)
    L.pih_debug_server_stamps(sim._h, stamps)
    spans.append((stamps[1] - stamps[0]) / 1e3)
    # (cudaMemcpyFromSymbol on the default stream while the server kernel is resident on its non-blocking stream)
    L.pih_debug_ticks(buf)
    rows.append(list(buf))
    L.pih_debug_cta_ns(cta); ctas.append(np.array(list(cta)[:256], dtype=np.float64).reshape(-1, 2))
hs.close()
r = np.array(rows[50:], dtype=np.float64)
m = lambda a, b: float(np.median(r[:, b] - r[:, a])) / 1965.0
print(f"wall per step median {1e6 * np.median(walls[50:]):.1f} us; device: step seen by CTA 0 -> completion published by the last CTA median {np.median(spans[50:]):.1f} us")
print(f"CTA0: poll start -> seq seen {m(14, 15):.1f} | seq seen -> IK warp go {m(15, 0):.1f} | IK {m(0, 1):.1f} | D: go -> ABA {m(2, 12):.1f} own cols {m(12, 13):.1f} -> barrier-ready {m(13, 3):.1f} "
      f"| setup {m(3, 4):.1f} PGS {m(4, 5):.1f} tail {m(5, 7):.1f} outputs+fence+count {m(7, 8):.1f} | C: front {m(9, 10):.1f} cols {m(10, 11):.1f} | D go->end {m(2, 8):.1f} | end -> next poll start {float(np.median(r[1:, 14] - r[:-1, 8])) / 1965.0:.1f}")
c = np.stack(ctas[50:]); dur = (c[:, :, 1] - c[:, :, 0]) / 1e3
print(f"per-CTA step duration by globaltimer: CTA0 {np.median(dur[:, 0]):.1f} us, median {np.median(dur):.1f}, p90 {np.percentile(dur, 90):.1f}, max-per-step median {np.median(dur.max(1)):.1f}; start skew (max-min of step starts) median {np.median((c[:, :, 0].max(1) - c[:, :, 0].min(1)) / 1e3):.1f} us; clock64 span CTA0 {m(2, 8):.1f}")
print("slowest CTAs:", np.argsort(-dur.mean(0))[:10].tolist(), np.sort(-dur.mean(0))[:10].round(1).tolist())
