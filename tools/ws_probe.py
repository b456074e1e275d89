import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
for n in (4096, 8192, 16384, 32768, 65536):
    for ws in ("0", "1"):
        os.environ["PIH_WS"] = ws
        sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
        sim.reset()
        offs = sim.env_offsets().cuda() if n <= 16384 else torch.zeros(n, 3, device="cuda")
        acts = []
        for _ in range(4):
            a = torch.rand(n, 4, device="cuda") * 2 - 1
            a[:, :3] += offs
            acts.append(a.contiguous())
        for i in range(200):
            sim.step(acts[i % 4])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(200):
            sim.step(acts[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        e0.record(); sim.rollout(200, count=False); e1.record(); torch.cuda.synchronize()
        mr = e0.elapsed_time(e1) / 200
        print(f"n={n:6d} ws={ws}  step {ms*1e3:8.1f} us ({n/ms/1e3:7.1f} M/s)   rollout {mr*1e3:8.1f} us/step ({n/mr/1e3:7.1f} M/s)", flush=True)
        sim.close()
