import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
for n in (4096, 65536, 262144):
    for block in (64, 128):
        os.environ["PIH_BLOCK"] = str(block)
        sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0, action_frame_env=True)
        sim.reset()
        sim.rollout(200, count=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        sim.rollout(200, count=False)          # exactly one full episode
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        print(f"lib={os.path.basename(os.environ.get('PIH_LIB','default'))} n={n:7d} block={block:3d}  {ms*1e3:8.1f} us/step  {n/ms/1e3:8.2f} M env-steps/s", flush=True)
        sim.close()
