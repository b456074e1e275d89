import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.sim import BatchedSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset(); sim.rollout(100)
out = torch.empty(n, 300, 300, 4, device="cuda")
for _ in range(3):
    sim.render(out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    sim.render(out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = out.numel() * 4 / 1e9
print(f"render {n} envs x 300x300x4: {ms:.3f} ms  -> {gb / (ms * 1e-3):.0f} GB/s written ({gb:.2f} GB), {n / (ms * 1e-3):.3e} frames/s")
# same scene with the fingers fully open (their boxes leave the field of view): what the store path does when almost
# every warp takes the background path
st = sim.state.clone()
flat = st.permute(1, 0, 2).reshape(n, -1)
flat[:, 7] = 0.04; flat[:, 8] = 0.04
sim.state.copy_(flat.reshape(n, -1, 4).permute(1, 0, 2))
for _ in range(3):
    sim.render(out=out)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    sim.render(out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"fingers open: {ms:.3f} ms  -> {gb / (ms * 1e-3):.0f} GB/s written, {n / (ms * 1e-3):.3e} frames/s")
