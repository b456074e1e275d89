"""Dev tool: cost of the learner-buffer delivery at a large batch on ONE GPU (world size 1: copy + flag are local), to
separate the launch / copy overhead from NVLink effects."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peg_in_hole_gym_b200.parallel import PeerLearnerBuffer
from peg_in_hole_gym_b200.sim import BatchedSim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
sim = BatchedSim(n, sub_num=n, offset=(2.0, 3.0, 0.0), seed=0, max_steps=200, auto_reset=True, device=0)
sim.reset(); sim.rollout(60, count=False)
state0 = sim.state.clone()
off = torch.zeros(n, 4, device="cuda"); off[:, :3] = sim.env_offsets_device()
g = torch.Generator(device="cuda"); g.manual_seed(0)
ring = [(torch.rand(n, 4, device="cuda", generator=g) * 2 - 1 + off).contiguous() for _ in range(6)]
buf = torch.zeros(PeerLearnerBuffer.total_bytes(n, 5, 1), dtype=torch.uint8, device="cuda:0")
peer = PeerLearnerBuffer(sim, rank=0, world_size=1, buf=buf)


def timed(fn, k=24):
    sim.state.copy_(state0)          # every mode steps the same states
    for i in range(6):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(k):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / k


print("none               ", round(timed(lambda i: sim.step(ring[i % 6])), 1), "us/step")
print("peer (no wait)     ", round(timed(lambda i: peer.step(ring[i % 6])), 1), "us/step")
print("peer (folded wait) ", round(timed(lambda i: peer.step(ring[i % 6], wait_prev=True)), 1), "us/step")
print("none again         ", round(timed(lambda i: sim.step(ring[i % 6])), 1), "us/step")
