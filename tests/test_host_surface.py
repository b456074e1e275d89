"""Host-side logic and the C-ABI surface, without a GPU: gym shim, spaces, registry, helpers, sharding, and that
libpih_b200.so loads and exports every symbol include/pih.h declares (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pih.h")).read()
    declared = set(re.findall(r"\b(pih_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = os.path.join(ROOT, "peg_in_hole_gym_b200", "csrc", "libpih_b200.so")
    if not os.path.exists(lib):                      # fresh checkout: the .so is git-ignored, nvcc cross-compiles on CPU
        import __graft_entry__
        __graft_entry__.build_cuda()
    assert os.path.exists(lib), "build the CUDA library first (__graft_entry__.build())"
    L = C.CDLL(lib)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/pih.h but not exported"
    from peg_in_hole_gym_b200 import _native
    assert set(_native.EXPORTS) == declared
    LL = _native.lib()
    assert LL.pih_abi_version() == 2 and LL.pih_state_words() == 48 and LL.pih_obs_dim() == 5
    assert LL.pih_action_dim(0) == 4 and LL.pih_action_dim(1) == 6 and LL.pih_action_dim(7) < 0


def test_host_env_offset_matches_oracle_layout():
    import oracle_lib as ol
    from peg_in_hole_gym_b200 import _native
    L = _native.lib()
    for offset, sub in [((2.0, 3.0, 0.0), 10), ((0.0, 3.0, 1.0), 5), ((2.0, 3.0, 0.5), 4096)]:
        cfg = _native.make_config(sub * 2, sub_num=sub, offset=offset)
        ocfg = ol.make_config(sub * 2, sub_num=sub, offset=offset)
        a, b = (C.c_float * 3)(), (C.c_float * 3)()
        for i in range(0, sub * 2, max(1, sub // 20)):
            assert L.pih_env_offset(C.byref(cfg), i, a) == 0
            ol.lib().pih_oracle_env_offset(C.byref(ocfg), i, b)
            assert list(a) == list(b)


def test_create_fails_loudly_without_gpu_or_with_bad_config():
    import torch
    from peg_in_hole_gym_b200 import _native
    L = _native.lib()
    h = C.c_void_p()
    bad = _native.make_config(0)
    assert L.pih_create(C.byref(bad), C.c_void_p(16), C.byref(h)) == _native.PIH_ERR_INVALID
    bad = _native.make_config(4, sub_num=4, offset=(0, 0, 0))           # reference assert base_env.py:17
    assert L.pih_create(C.byref(bad), C.c_void_p(16), C.byref(h)) == _native.PIH_ERR_INVALID
    if not torch.cuda.is_available():
        ok = _native.make_config(4, sub_num=4, offset=(1, 0, 0))
        assert L.pih_create(C.byref(ok), C.c_void_p(16), C.byref(h)) == _native.PIH_ERR_CUDA
        import peg_in_hole_gym_b200 as g
        with pytest.raises(_native.PihError):
            g.make('peg-in-hole-mp-v0', client=g.DIRECT, task='peg-in-hole', mp_num=2, sub_num=2, offset=[2., 3., 0.])


def test_registry_and_task_list_and_asserts():
    import peg_in_hole_gym_b200 as g
    from peg_in_hole_gym_b200 import gym_shim
    from peg_in_hole_gym_b200.envs import base_env
    assert 'peg-in-hole-v0' in gym_shim._REGISTRY and 'peg-in-hole-mp-v0' in gym_shim._REGISTRY
    assert set(base_env.TASK_LIST) >= {'peg-in-hole', 'random-fly'}
    with pytest.raises(AssertionError):
        base_env.BaseEnv(g.DIRECT, task='no-such-task')
    with pytest.raises(AssertionError):
        base_env.BaseEnv(g.DIRECT, task='peg-in-hole', task_num=4, offset=[0, 0, 0])
    with pytest.raises(KeyError):
        g.make('nope-v0')


def test_spaces_match_reference_declarations():
    from peg_in_hole_gym_b200.envs.peg_in_hole import PegInHole, PegInHoleUR5
    from peg_in_hole_gym_b200.envs.utils import (MPMultiAgentActionSpace, MPMultiAgentObservationSpace,
                                                  MultiAgentActionSpace, MultiAgentObservationSpace)
    assert PegInHole.action_space.shape == (4,) and PegInHole.observation_space.shape == (5,)   # peg_in_hole.py:12-13
    assert PegInHoleUR5.action_space.shape == (6,)
    assert (PegInHole.action_space.low == -1).all() and (PegInHole.action_space.high == 1).all()
    sub = MultiAgentActionSpace([PegInHole.action_space] * 3)
    assert len(sub) == 3 and sub.shape == (4,) and len(sub.sample()) == 3 and sub.sample()[0].shape == (4,)
    mp = MPMultiAgentActionSpace([sub, sub])
    assert len(mp) == 0                                     # the reference skips list.__init__ (utils.py:201)
    s = mp.sample()
    assert len(s) == 2 and len(s[0]) == 3 and s[0][0].dtype == np.float32
    assert mp.sample_array().shape == (2, 3, 4)
    osub = MultiAgentObservationSpace([PegInHole.observation_space] * 3)
    omp = MPMultiAgentObservationSpace([osub, osub])
    assert omp.contains(omp.sample()) and not osub.contains([np.full(5, 2.0)] * 3)
    assert omp.shape == (5,) and (omp.high == 1).all()


def test_reference_helpers():
    from peg_in_hole_gym_b200.envs import utils as U
    assert U.vel_constraint([0, 0, 0], [1, -1, 0.001], 2 / 240.) == [2 / 240., -2 / 240., 0.001]      # utils.py:85-95
    assert U.data_normalize([1.0, 5.0], [[0, 2], [0, 10]]) == [0.5, 0.5]                                # utils.py:19-22
    for _ in range(200):                                                                               # utils.py:97-107
        p = U.random_pos_in_panda_space()
        assert abs(p[0] ** 2 + p[1] ** 2 + (p[2] - 0.2) ** 2 - 0.49) < 1e-9 and p[2] >= 0.2
    from peg_in_hole_gym_b200.envs.meta_env import MetaEnv
    assert MetaEnv.vel_constraint([0.0], [1.0], 0.5) == [0.5]                                          # fixes meta_env.py:107


def test_shard_clients_partition():
    from peg_in_hole_gym_b200.parallel import shard_clients
    for mp, ws in [(8, 1), (8, 2), (8, 8), (10, 4), (3, 8)]:
        got = [shard_clients(mp, r, ws) for r in range(ws)]
        assert got[0][0] == 0 and got[-1][1] == mp
        assert all(got[i][1] == got[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1


def test_peer_learner_buffer_layout_is_agreed_from_all_ranks_counts():
    """uneven shards (mp_num % world_size != 0) must not make ranks disagree about slice / flag offsets: the
    layout is a pure function of the all-gathered per-rank counts."""
    from peg_in_hole_gym_b200.parallel import PeerLearnerBuffer as PLB, shard_clients
    counts = [(b - a) * 6 for a, b in (shard_clients(5, r, 2) for r in range(2))]         # 18 and 12 envs
    offs, total = PLB.layout(counts, 5)
    assert counts == [18, 12] and offs[0] == 0 and offs[1] % PLB.ALIGN == 0
    assert offs[1] >= 18 * 25 and total - PLB.ALIGN >= offs[1] + 12 * 25                   # slices do not overlap each other or the flags
    assert PLB.total_bytes(160, 5, 2) == PLB.layout([160, 160], 5)[1]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    from peg_in_hole_gym_b200.parallel import gather_rows_to_rank0, pack_rows, shard_clients, unpack_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mp_num, sub = 5, 6                                       # uneven split: ranks own 3 and 2 clients
    lo, hi = shard_clients(mp_num, rank, world)
    n = (hi - lo) * sub
    cfg = ol.make_config(n, sub_num=sub, env_id_base=lo * sub, offset=(2.0, 3.0, 0.0), seed=17, max_steps=9, auto_reset=1)
    o = ol.Oracle(cfg)                                       # the oracle stands in for the GPU engine on CPU
    o.reset()
    o.rollout(25)
    rows = pack_rows(torch.from_numpy(o.obs), torch.from_numpy(o.reward), torch.from_numpy(o.done.astype(np.float32)))
    allrows = gather_rows_to_rank0(rows)
    if rank == 0:
        obs, rew, done = unpack_rows(allrows)
        q.put((obs.numpy(), rew.numpy(), done.numpy()))
    dist.destroy_process_group()


def test_world_size_2_sharding_and_gather_matches_single_process():
    """N>1 host path on CPU (gloo): shards keyed by global env id reproduce the single-process result."""
    import torch.multiprocessing as mp
    import oracle_lib as ol
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    obs, rew, done = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref = ol.Oracle(ol.make_config(30, sub_num=6, offset=(2.0, 3.0, 0.0), seed=17, max_steps=9, auto_reset=1))
    ref.reset()
    ref.rollout(25)
    assert np.array_equal(obs, ref.obs) and np.array_equal(rew, ref.reward) and np.array_equal(done, ref.done != 0)
