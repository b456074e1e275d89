#!/usr/bin/env python3
"""Phase timing of the warp-specialised step kernel in a HOT run (dev tool; needs an experiment build:
tools/fast_build.sh ticks -DPIH_PHASE_TICKS, then PIH_LIB=tools/_build/libpih_ticks.so python tools/tick_probe.py).
CTA 0 records the SM clock at the phase boundaries; medians over many steps, in microseconds at 1965 MHz."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200 import _native  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
arm = sys.argv[3] if len(sys.argv) > 3 else "panda"
task = sys.argv[4] if len(sys.argv) > 4 else "peg-in-hole"
preroll = int(sys.argv[5]) if len(sys.argv) > 5 else 0
adim = 4 if arm == "panda" else 6
L = _native.lib()
sim = BatchedSim(n, arm=arm, task=task, offset=(2.0, 3.0, 0.0), max_steps=0 if task.endswith("scripted") else 200,
                 auto_reset=not task.endswith("scripted"), seed=0, device=0)
sim.reset()
if preroll:
    sim.rollout(preroll, count=False)
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = []
for _ in range(steps):
    a = torch.rand(n, adim, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    acts.append(a.contiguous())
buf = (C.c_longlong * 16)()
cta = (C.c_ulonglong * 2048)()
rows, ctas = [], []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for t in range(steps):
    ev0.record()
    sim.step(acts[t])
    ev1.record()
    torch.cuda.synchronize()
    tot += ev0.elapsed_time(ev1)
    L.pih_debug_ticks(buf)
    rows.append(list(buf)[:16])
    L.pih_debug_cta_ns(cta)
    ctas.append(np.array(list(cta)[:2 * ((n + 31) // 32)], dtype=np.float64).reshape(-1, 2))
    if t in (180, 190):
        d = (ctas[-1][:, 1] - ctas[-1][:, 0]) / 1e3
        flags = sim.int_word(11).cpu().numpy()
        q = sim.state.cpu().numpy().transpose(1, 0, 2).reshape(n, -1)[:, 0:9]
        order = np.argsort(-d)[:6]
        print(f"step {t}: per-CTA us: median {np.median(d):.1f} max {d.max():.1f}; slowest CTAs:")
        for c_ in order:
            fl = flags[32 * c_:32 * c_ + 32] >> 8
            qq = q[32 * c_:32 * c_ + 32]
            print(f"   cta {c_:4d} {d[c_]:6.1f} us  pair masks {sorted(set(fl.tolist()))}  min/max q: "
                  + " ".join(f"{qq[:, j].min():.2f}/{qq[:, j].max():.2f}" for j in range(6)))
r = np.array(rows[50:], dtype=np.float64)
mhz = 1965.0


def us(a, b):
    return float(np.median(r[:, b] - r[:, a])) / mhz


print(f"n={n}: event mean {1e3 * tot / steps:.1f} us/step (with a sync after every step)")
print(f"warp0  FK+apply_action(IK)         {us(0, 1):6.1f} us")
print(f"warp1  FK+ABA+Minv+collide+rows    {us(2, 3):6.1f} us")
print(f"warp1    of which FK+ABA+publish     {us(2, 12):6.1f} us, own M^-1 columns + free-body setup {us(12, 13):6.1f} us, waits + narrow phase + rows {us(13, 3):6.1f} us")
print(f"warp2  start -> contact summary    {us(9, 10):6.1f} us, -> its M^-1 columns done {us(9, 11):6.1f} us (after warp 1's start: {us(2, 11):6.1f} us)")
print(f"warp1  barrier+motor/limit setup   {us(3, 4):6.1f} us")
print(f"warp1  PGS                         {us(4, 5):6.1f} us")
print(f"warp1  integrate+FK+get_info       {us(5, 7):6.1f} us")
print(f"warp1  store state/obs/reward     {us(7, 8):6.1f} us")
print(f"warp1  total (CTA 0)               {us(2, 8):6.1f} us")
c = np.stack(ctas[50:])                      # [steps, ctas, 2] ns
dur = (c[:, :, 1] - c[:, :, 0]) / 1e3
span = (c[:, :, 1].max(1) - c[:, :, 0].min(1)) / 1e3
print(f"per-CTA duration (globaltimer): median {np.median(dur):.1f} us, p90 {np.percentile(dur, 90):.1f}, max-per-step median {np.median(dur.max(1)):.1f}; "
      f"first-start to last-end median {np.median(span):.1f} us; start skew median {np.median((c[:, :, 0].max(1) - c[:, :, 0].min(1)) / 1e3):.1f} us")

mx = dur.max(1)
print(f"slowest CTA per step: mean {mx.mean():.1f} us, median {np.median(mx):.1f}, p90 {np.percentile(mx, 90):.1f}; steps whose slowest CTA takes > 1.15 x the median CTA: "
      f"{float((mx > 1.15 * np.median(dur)).mean()):.2f}; CTAs over that bar per step: {float((dur > 1.15 * np.median(dur)).sum(1).mean()):.2f} of {dur.shape[1]}")
# per 25-step block of the episode (all envs of the batch share the phase): where does the time go?
print("block  IK    prologue  setup  PGS   tail   total(us)")
for b0 in range(0, steps, 25):
    rr = np.array(rows[b0:b0 + 25], dtype=np.float64)
    m = lambda a, b: float(np.median(rr[:, b] - rr[:, a])) / mhz
    print(f"{b0:5d} {m(0, 1):5.1f} {m(2, 3):8.1f} {m(3, 4):6.1f} {m(4, 5):6.1f} {m(5, 7):5.1f} {m(2, 7):7.1f}")
