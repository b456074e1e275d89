"""Small workload for compute-sanitizer (one tool per gpurun call): every kernel family once, tiny batches.
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.parallel import PeerLearnerBuffer
from peg_in_hole_gym_b200.sim import BatchedSim

kw = dict(offset=(2.0, 3.0, 0.0), seed=1, max_steps=6, auto_reset=True, device=0)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for fam in ("ws", "fused", "split"):
    for arm, task in (("panda", "peg-in-hole"), ("ur5", "peg-in-hole"), ("panda", "random-fly")):
        n = 70
        sim = BatchedSim(n, arm=arm, task=task, kernel_family=fam, **kw)
        sim.reset()
        off = torch.zeros(n, sim.action_dim, device="cuda"); off[:, :3] = sim.env_offsets_device()
        for t in range(8):
            sim.step(torch.rand(n, sim.action_dim, device="cuda", generator=g) * 2 - 1 + off)
        sim.rollout(5)
        torch.cuda.synchronize()
        sim.close()
    print("family", fam, "ok", flush=True)
# scripted task (state machine in apply_action), camera, labels
sim = BatchedSim(40, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=2, device=0)
sim.reset(); sim.rollout(12); sim.render(32, 24); sim.grasp_labels(32, 24)
torch.cuda.synchronize(); sim.close()
# fused gather epilogue + folded wait: two members in one process
n = 64
kw2 = dict(offset=(2.0, 3.0, 0.0), max_steps=12, auto_reset=True, seed=9, sub_num=2 * n, device=0)
sims = [BatchedSim(n, env_id_base=r * n, **kw2) for r in range(2)]
for s_ in sims:
    s_.reset()
buf = torch.zeros(PeerLearnerBuffer.total_bytes(n, 5, 2), dtype=torch.uint8, device="cuda:0")
mem = [PeerLearnerBuffer(sims[r], rank=r, world_size=2, buf=buf) for r in range(2)]
for t in range(6):
    a = [torch.rand(n, 4, device="cuda", generator=g) * 2 - 1 for _ in range(2)]
    mem[1].step(a[1]); mem[0].step(a[0], wait_prev=True)
    mem[0].step_no = mem[1].step_no
mem[0].wait(); mem[0].check()
print("peer epilogue ok", flush=True)
# flexible pipe: warp-per-env kernels
sim = BatchedSim(10, peg_model="pipe", offset=(2.0, 3.0, 0.0), seed=3, max_steps=40, auto_reset=True, device=0)
sim.reset(); sim.rollout(50)
off = torch.zeros(10, 4, device="cuda"); off[:, :3] = sim.env_offsets_device()
sim.step(torch.rand(10, 4, device="cuda", generator=g) * 2 - 1 + off)
torch.cuda.synchronize(); sim.close()
print("pipe ok", flush=True)
print("SANITIZE_TARGET_DONE", flush=True)
