#!/usr/bin/env python3
"""GPU perf exploration (not part of the bench contract): kernel time vs N, block size, rollout chunking."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from peg_in_hole_gym_b200.sim import BatchedSim


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    arm = sys.argv[1] if len(sys.argv) > 1 else "panda"
    for n in (4096, 16384, 65536, 262144):
        for block in (32, 64, 128):
            os.environ["PIH_BLOCK"] = str(block)
            sim = BatchedSim(n, arm=arm, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
            sim.reset()
            adim = sim.action_dim
            offs = sim.env_offsets().cuda() if n <= 16384 else None
            acts = []
            for _ in range(8):
                a = torch.rand(n, adim, device="cuda") * 2 - 1
                if offs is not None:
                    a[:, :3] += offs
                acts.append(a)
            if offs is None:
                sim.close()
                sim = BatchedSim(n, arm=arm, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0,
                                 action_frame_env=True)
                sim.reset()
            i = [0]

            def step():
                sim.step(acts[i[0] % 8]); i[0] += 1
            for _ in range(250):
                step()
            ms = timed(step, 100)
            r1 = timed(lambda: sim.rollout(1, count=False), 50)
            r10 = timed(lambda: sim.rollout(10, count=False), 10) / 10
            r100 = timed(lambda: sim.rollout(100, count=False), 2) / 100
            print(f"arm={arm} n={n:7d} block={block:3d} step {ms*1e3:8.1f} us  {n/ms/1e3:8.2f} M env-steps/s | rollout/step: x1 {r1*1e3:8.1f} us, x10 {r10*1e3:8.1f} us, x100 {r100*1e3:8.1f} us",
                  flush=True)
            sim.close()


if __name__ == "__main__":
    main()
