"""Worker of test_peer_write_gather_across_processes: two processes (gloo rendezvous), BOTH on cuda:0, share rank 0's
learner buffer through CUDA IPC exactly as two ranks on two GPUs would (tests/ only; launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peg_in_hole_gym_b200.parallel import PeerLearnerBuffer  # noqa: E402
from peg_in_hole_gym_b200.sim import BatchedSim  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    counts = [128, 96][:ws] if ws <= 2 else [128] * ws      # UNEVEN shards: the layout must come from the agreed counts
    base = [sum(counts[:r]) for r in range(ws)]
    n = counts[rank]
    kw = dict(offset=(2.0, 3.0, 0.0), max_steps=9, auto_reset=True, seed=4, sub_num=sum(counts), device=0)
    sim = BatchedSim(n, env_id_base=base[rank], **kw)
    sim.reset()
    peer = PeerLearnerBuffer(sim)
    assert peer.counts == counts
    refs = [BatchedSim(counts[r], env_id_base=base[r], **kw) for r in range(ws)] if rank == 0 else []
    for s_ in refs:
        s_.reset()
    offs = [torch.cat([refs[r].env_offsets().cuda(), torch.zeros(counts[r], 1, device="cuda")], 1) for r in range(ws)] if rank == 0 else None
    my_off = torch.cat([sim.env_offsets().cuda(), torch.zeros(n, 1, device="cuda")], 1)
    ok = True
    for t in range(25):
        acts = [torch.from_numpy(np.random.default_rng(1000 * t + r).uniform(-1, 1, (counts[r], 4)).astype(np.float32)).cuda() for r in range(ws)]
        peer.step(acts[rank] + my_off)
        # Both ranks share ONE GPU here, so rank 0 must not spin on a flag that another process's kernel has yet to
        # write (kernels of different processes are time-sliced, not concurrent: B200_PROFILING.md).  Every rank
        # finishes its step first; rank 0's wait kernel then finds all flags set.  On a multi-GPU node the wait overlaps
        # the peers' kernels (bench.py --gather peer).
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            rows = peer.wait()
            for r in range(ws):
                refs[r].step(acts[r] + offs[r])
                torch.cuda.synchronize()
                ok = ok and torch.equal(rows[r][0], refs[r].obs) and torch.equal(rows[r][1], refs[r].reward) and torch.equal(rows[r][2], refs[r].done)
        torch.cuda.synchronize()
        dist.barrier()          # lockstep only because rank 0 re-uses the buffer slice every step in this test
    # deferred publication: the steps only store their rows, every rank's next launch (or flush) announces them
    for t in range(25, 33):
        acts = [torch.from_numpy(np.random.default_rng(1000 * t + r).uniform(-1, 1, (counts[r], 4)).astype(np.float32)).cuda() for r in range(ws)]
        peer.step(acts[rank] + my_off, deferred=True)
        torch.cuda.synchronize()
        dist.barrier()
        peer.flush()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            rows = peer.wait()
            for r in range(ws):
                refs[r].step(acts[r] + offs[r])
                torch.cuda.synchronize()
                ok = ok and torch.equal(rows[r][0], refs[r].obs) and torch.equal(rows[r][1], refs[r].reward) and torch.equal(rows[r][2], refs[r].done)
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        peer.check()
        print("PEER_GATHER_OK" if ok else "PEER_GATHER_MISMATCH", flush=True)
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
