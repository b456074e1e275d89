#!/usr/bin/env python3
"""Where the end-to-end step time goes (dev tool): wall clock per step for
  dev+sync : device-resident actions/outputs, stream sync after every step
  zc-in    : actions read from mapped host memory by the kernel, outputs stay on the device
  zc-all   : pih_step_host zero-copy (actions in, obs|reward|done out, host memory)
  staged   : pih_step_host with PIH_NO_ZERO_COPY=1 semantics (H2D + step + D2H), run in a second process"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200 import _native  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402

n, steps = 4096, 2000
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
h_acts, obs_h, rew_h, done_h = sim.pinned_host_buffers(64)
for b in range(64):
    h_acts[b][...] = ring[b].cpu().numpy()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, name):
    for i in range(50):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:10s} {1e6 * dt / steps:7.1f} us/step  {n * steps / dt / 1e6:6.1f} M env-steps/s")


def dev_nosync(i):
    L.pih_step(sim._h, C.c_void_p(ring[(i * 7) % 64].data_ptr()), C.c_void_p(sim.obs.data_ptr()), C.c_void_p(sim.reward.data_ptr()),
               C.c_void_p(sim.done.data_ptr()), st)


def dev_sync(i):
    dev_nosync(i)
    torch.cuda.current_stream().synchronize()


a_ptrs = [C.c_void_p(a.ctypes.data) for a in h_acts]


def zc_in(i):
    L.pih_step(sim._h, a_ptrs[(i * 7) % 64], C.c_void_p(sim.obs.data_ptr()), C.c_void_p(sim.reward.data_ptr()), C.c_void_p(sim.done.data_ptr()), st)
    torch.cuda.current_stream().synchronize()


host_step = sim.host_stepper(h_acts, obs_h, rew_h, done_h)


def zc_all(i):
    host_step((i * 7) % 64)


def noop(i):
    L.pih_action_dim(0)


timeit(noop, "ctypes-noop")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for blk in range(4):
    torch.cuda.synchronize()
    ev0.record()
    t0 = time.perf_counter()
    for i in range(500):
        dev_nosync(i)
    t1 = time.perf_counter()
    ev1.record()
    torch.cuda.synchronize()
    print(f"dev block {blk}: device {1e3 * ev0.elapsed_time(ev1) / 500:.1f} us/step, host enqueue {1e6 * (t1 - t0) / 500:.1f} us/step")
timeit(dev_nosync, "dev")
timeit(dev_sync, "dev+sync")
timeit(zc_in, "zc-in")
timeit(zc_all, "zc-all" if not os.environ.get("PIH_NO_ZERO_COPY") else "staged")
