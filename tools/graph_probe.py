import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset()
offs = sim.env_offsets().cuda()
acts = []
for _ in range(8):
    a = torch.rand(n, 4, device="cuda") * 2 - 1
    a[:, :3] += offs
    acts.append(a.contiguous())
for i in range(300):
    sim.step(acts[i % 8])
torch.cuda.synchronize()
def ev():
    return torch.cuda.Event(enable_timing=True)
# 1. python loop, one event pair around all
e0, e1 = ev(), ev()
t0 = time.perf_counter(); e0.record()
for i in range(200):
    sim.step(acts[i % 8])
e1.record(); t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"python loop: {e0.elapsed_time(e1)/200*1e3:.1f} us/step on device; CPU issue time {t_issue/200*1e6:.1f} us/step")
# 2. CUDA graph of 8 steps
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3):
        sim.step(acts[i % 8])
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(8):
        sim.step(acts[i % 8])
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(25):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"cuda graph (8 steps/graph): {e0.elapsed_time(e1)/200*1e3:.1f} us/step")
# 3. per-kernel events inside python loop
evs = [(ev(), ev()) for _ in range(100)]
for i in range(100):
    evs[i][0].record(); sim.step(acts[i % 8]); evs[i][1].record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in evs)
print(f"per-step events: median {ts[50]*1e3:.1f} us, min {ts[0]*1e3:.1f} us, max {ts[-1]*1e3:.1f} us")
e0.record(); sim.rollout(200, count=False); e1.record(); torch.cuda.synchronize()
print(f"rollout(200): {e0.elapsed_time(e1)/200*1e3:.1f} us/step")
