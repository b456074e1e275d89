import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
n = 4096
sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
sim.rollout(200)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
for t in range(200):
    a = torch.rand(n, 4, device="cuda", generator=g) * 2 - 1
    a[:, :3] += offs
    torch.cuda.synchronize()
    e0.record(); sim.step(a); e1.record(); torch.cuda.synchronize()
    fl = sim.int_word(11)
    q = sim.words(0, 9); qd = sim.words(12, 9)
    out.append((t, e0.elapsed_time(e1) * 1e3, float(((fl >> 8) & 1).float().mean()), float(((fl >> 8) & 16).bool().float().mean()),
                float(qd[:, :7].abs().max()), float(sim.words(35, 3).abs().max())))
for r in out:
    if r[0] < 60 or r[0] % 10 == 0:
        print("t=%3d  %7.1f us  peg-table %.2f arm-table %.2f  max|qd| %8.2f max|w_peg| %8.2f" % r)
import statistics
print("mean us/step over the episode:", statistics.mean(r[1] for r in out))
