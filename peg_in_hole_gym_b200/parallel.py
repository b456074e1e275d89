"""Multi-GPU plumbing: one process per GPU, envs sharded with NO data-path collective.

`mp_num` of the reference (one pybullet client process each, envs/base_env_mp.py:27-35) maps to ranks: every rank
owns a contiguous block of clients, hence a contiguous range of global env ids.  RNG is keyed by global env id, so
results do not depend on the number of ranks.  The only collective is the OPTIONAL gather of the packed
(obs5, reward, done) rows into a rank-0 learner buffer (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import os

import torch


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_clients(mp_num, rank, world_size):
    """Contiguous block of reference clients owned by `rank`: [lo, hi).  Remainder goes to the low ranks."""
    if world_size <= 1:
        return 0, mp_num
    base, rem = divmod(mp_num, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_rows(obs, reward, done):
    """[n,5] obs + [n] reward + [n] done -> one [n,7] float32 block (28 B/env)."""
    return torch.cat([obs, reward.unsqueeze(1), done.to(obs.dtype).unsqueeze(1)], dim=1).contiguous()


def unpack_rows(rows):
    return rows[:, :5], rows[:, 5], rows[:, 6] != 0


def gather_rows_to_rank0(rows, counts=None):
    """Gather every rank's [n_r,7] block on rank 0, in rank order.  Returns the [sum n_r,7] tensor on rank 0 and
    None elsewhere.  Uses all_gather on padded blocks when shard sizes differ (NCCL has no ragged gather)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    ws, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        c = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        cl = [torch.zeros_like(c) for _ in range(ws)]
        dist.all_gather(cl, c)
        counts = [int(v.item()) for v in cl]
    mx = max(counts)
    if rows.shape[0] < mx:
        rows = torch.cat([rows, rows.new_zeros(mx - rows.shape[0], rows.shape[1])], 0)
    if all(v == mx for v in counts):
        out = rows.new_empty(ws * mx, rows.shape[1]) if rank == 0 else None
        gl = list(out.chunk(ws, 0)) if rank == 0 else None
        dist.gather(rows, gl, dst=0)
        return out
    buf = [torch.empty_like(rows) for _ in range(ws)]
    dist.all_gather(buf, rows)
    if rank != 0:
        return None
    return torch.cat([b[:n] for b, n in zip(buf, counts)], 0)


class _DevMem:
    """Raw device allocation exposed to torch through __cuda_array_interface__ (zero-copy view, no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerLearnerBuffer:
    """Gather to a rank-0 learner buffer WITHOUT a collective (preferred form).

    Rank 0 owns one device allocation ``[world][obs n*5 f32 | reward n f32 | done n u8]`` plus one flag word per rank.
    Every other rank opens its CUDA IPC handle ON ITS OWN GPU, which maps rank 0's HBM into that rank's address space as
    peer memory (NVLink / NVSwitch), and hands its slice to ``pih_step`` as the output pointers: the step kernel itself
    stores the rows into rank 0's HBM.  ``pih_signal_flag`` then publishes the step number, and rank 0's ``wait`` (a
    bounded device-side wait in its own stream) returns the views of all ranks' rows.  Nothing is packed, copied or
    reduced, and no host synchronisation is involved.  (There is no back-pressure: a learner that reads asynchronously
    should alternate two buffers.)

    ``rank`` / ``world_size`` / ``buf`` may be given explicitly to build several members inside ONE process on one GPU
    (used by the single-GPU test); by default they come from torch.distributed and the buffer is shared over IPC."""

    ALIGN = 256

    def __init__(self, sim, rank=None, world_size=None, buf=None, counts=None):
        import ctypes as C
        from . import _native as N
        self.sim = sim
        r, ws = world()
        self.rank = r if rank is None else int(rank)
        self.world_size = ws if world_size is None else int(world_size)
        od = sim.obs_dim
        self.obs_dim = od
        # Every rank's slice is laid out from the AGREED per-rank env counts (shard_clients hands out uneven blocks when
        # mp_num % world_size != 0): one all_gather at construction, prefix offsets from it.  A layout computed from the
        # local count alone would make ranks disagree about where slices and flag words live.
        if counts is None:
            if buf is None and self.world_size > 1:
                import torch.distributed as dist
                counts = [None] * self.world_size
                dist.all_gather_object(counts, int(sim.n))
            else:
                counts = [sim.n] * self.world_size
        self.counts = [int(c) for c in counts]
        if len(self.counts) != self.world_size or self.counts[self.rank] != sim.n:
            raise ValueError(f"learner buffer: per-rank env counts {self.counts} do not match this rank's {sim.n} envs")
        if min(self.counts) <= 0:
            raise ValueError(f"learner buffer: a rank owns no envs ({self.counts}); use mp_num >= world_size")
        self.n = sim.n
        self.offsets, total = self.layout(self.counts, od)
        self.flag_off = total - self.ALIGN
        self._raw, self._owner = None, False
        if buf is not None:                                   # members of one process share a torch tensor
            self.buf, self.base = buf, buf.data_ptr()
            if buf.numel() < total:
                raise ValueError("learner buffer too small")
        else:
            import torch.distributed as dist
            L = sim._L
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            if self.rank == 0:
                N.check(L.pih_peer_alloc(sim.device.index, total, C.byref(ptr), handle), None, "pih_peer_alloc")
                self._owner = True
            obj = [handle.raw if self.rank == 0 else None]
            if self.world_size > 1:
                dist.broadcast_object_list(obj, src=0)
                if self.rank != 0:
                    N.check(L.pih_peer_open(sim.device.index, obj[0], C.byref(ptr)), None, "pih_peer_open")
            self._raw = self.base = int(ptr.value)
            # torch views exist on the owner only (the other ranks need nothing but the addresses)
            self.buf = torch.as_tensor(_DevMem(self.base, total), device=sim.device) if self.rank == 0 else None
            if self.world_size > 1:
                dist.barrier()
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=sim.device)
        self.step_no = 0
        self._rows = None
        self._my_ptrs = self._ptrs(self.rank) + (self._flag_ptr(self.rank),)

    def close(self):
        if self._raw:
            self.buf = None
            torch.cuda.synchronize(self.sim.device)
            self.sim._L.pih_peer_release(self._raw, int(self._owner))
            self._raw = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def layout(cls, counts, obs_dim):
        """(slice offsets per rank, total bytes): slices [obs n_r*od f32 | reward n_r f32 | done n_r u8] rounded up to ALIGN,
        then one ALIGN-sized block of flag words."""
        offs, pos = [], 0
        for n in counts:
            offs.append(pos)
            raw = n * obs_dim * 4 + n * 4 + n
            pos += (raw + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN
        return offs, pos + cls.ALIGN

    @classmethod
    def total_bytes(cls, n, obs_dim, world_size):
        return cls.layout([n] * world_size, obs_dim)[1]

    def _views(self, r):
        n, od = self.counts[r], self.obs_dim
        base = self.buf[self.offsets[r]:]
        obs = base[:n * od * 4].view(torch.float32).view(n, od)
        rew = base[n * od * 4:n * od * 4 + n * 4].view(torch.float32)
        done = base[n * od * 4 + n * 4:n * od * 4 + n * 4 + n]
        return obs, rew, done

    def _ptrs(self, r):
        n, od = self.counts[r], self.obs_dim
        b = self.base + self.offsets[r]
        return b, b + n * od * 4, b + n * od * 4 + n * 4

    def _flag_ptr(self, r):
        return self.base + self.flag_off + 4 * r

    def step(self, action, sim=None, wait_prev=False, timeout_ms=2000, deferred=False):
        """pih_step_signal: this rank's slice of the learner buffer as the outputs; the kernel publishes the step number.
        ``sim``: another BatchedSim of the same size on this rank (several env batches may feed one learner buffer).
        ``wait_prev`` (owner rank): the kernel first waits until every rank has delivered the PREVIOUS step's rows
        (pih_step_signal_wait) - the consumer's wait folded into its next launch, no wait kernel of its own; rows of
        step t are then complete for everything stream-ordered after step t+1 (alternate two buffers to read them
        while the peers already write step t+1).
        ``deferred`` (pih_step_publish_prev): the step only stores its rows; the NEXT launch announces them with one
        system-scope fence at its start, so no CTA fences on its way out (~6 us per step at 4096 envs).  The flags then
        arrive one launch later - exactly when ``wait_prev`` looks for them; call ``flush()`` on every rank before an
        explicit ``wait()`` or at the end of a rollout."""
        from . import _native as N
        sim = self.sim if sim is None else sim
        po, pr, pd, pf = self._my_ptrs
        self.step_no += 1
        if deferred:
            waiting = wait_prev and self.step_no > 1
            rc = sim._L.pih_step_publish_prev(sim._h, action.data_ptr(), po, pr, pd, pf, self.step_no - 1,
                                             (self.base + self.flag_off) if waiting else None, self.world_size if waiting else 0,
                                             self.step_no - 1, int(timeout_ms), self.timed_out.data_ptr() if waiting else None, sim._stream())
            self._pending = True
        elif wait_prev and self.step_no > 1:
            rc = sim._L.pih_step_signal_wait(sim._h, action.data_ptr(), po, pr, pd, pf, self.step_no, self.base + self.flag_off,
                                            self.world_size, self.step_no - 1, int(timeout_ms), self.timed_out.data_ptr(), sim._stream())
        else:
            rc = sim._L.pih_step_signal(sim._h, action.data_ptr(), po, pr, pd, pf, self.step_no, sim._stream())
        if rc:
            N.check(rc, sim._h, "pih_step_signal")

    def flush(self):
        """Deferred mode: announce the rows of the last step issued on this rank (pih_signal_flag)."""
        from . import _native as N
        if getattr(self, "_pending", False):
            N.check(self.sim._L.pih_signal_flag(self.sim._h, self._my_ptrs[3], self.step_no, self.sim._stream()), self.sim._h, "pih_signal_flag")
            self._pending = False

    def wait(self, timeout_ms=2000):
        """Rank 0: enqueue the wait for every rank's flag of the current step; returns [(obs, reward, done)] per rank
        (views of the learner buffer, valid for work enqueued after this call on the same stream)."""
        from . import _native as N
        sim = self.sim
        self.flush()      # (deferred mode: this rank's own last step; the peers flush theirs)
        rc = sim._L.pih_wait_flags(sim._h, self.base + self.flag_off, self.world_size, self.step_no, int(timeout_ms),
                                   self.timed_out.data_ptr(), sim._stream())
        if rc:
            N.check(rc, sim._h, "pih_wait_flags")
        if self._rows is None:     # the views never change: build them once (24 tensor slices cost more than a step)
            self._rows = [self._views(r) for r in range(self.world_size)]
        return self._rows

    def check(self):
        """Host-side: raises if a device-side wait since the previous check() gave up (synchronises the stream); the rows
        returned by that wait() are not valid then."""
        v = int(self.timed_out.item())
        self.timed_out.zero_()          # re-armed: the next check() reports only waits that gave up after this one
        if v:
            raise RuntimeError(f"peer gather: rank {v - 1} did not publish step {self.step_no} in time")
