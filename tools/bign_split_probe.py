import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for split in ("0", "1"):
    os.environ["PIH_SPLIT"] = split
    sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
    sim.reset()
    sim.rollout(200, count=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for blk in range(4):
        torch.cuda.synchronize(); e0.record()
        sim.rollout(50, count=False)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 50 * 1e3)
    print(f"n={n} split={split}: us/step per 50-step quarter of the episode: " + " ".join(f"{v:.0f}" for v in out) + f"  mean {sum(out) / 4:.0f}", flush=True)
    sim.close()
