import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
sim = BatchedSim(4096, sub_num=4096, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
def t(steps, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): sim.rollout(steps, count=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * steps) * 1e3
sim.rollout(10)
for steps, reps in [(10, 20), (100, 4), (200, 2), (400, 1), (1000, 1), (100, 4), (2000, 1), (10, 20)]:
    print(f"rollout({steps}) x{reps}: {t(steps, reps):8.1f} us/step", flush=True)
