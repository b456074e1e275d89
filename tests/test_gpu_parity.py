"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): discrete outputs (done flags, contact-pair presence, reset indices,
episode/step counters) EXACT over a 200-step horizon; joint positions / EE / peg pose / reward within 1e-3 m,
1e-2 rad per step - in practice the float oracle build and the kernels agree bit for bit, which is asserted where
stated.  Drift of fp32 (GPU) against the fp64 oracle build is reported and bounded separately."""
import os

import numpy as np
import pytest

import oracle_lib as ol

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_POS, TOL_ANG = 1e-3, 1e-2   # north_star tolerances (m, rad)


def _sim(n, **kw):
    from peg_in_hole_gym_b200.sim import BatchedSim
    return BatchedSim(n, device=0, **kw)


def _pair(n, arm="panda", task="peg-in-hole", use_double=False, **kw):
    sim = _sim(n, arm=arm, task=task, **kw)
    okw = dict(kw)
    okw.pop("kernel_family", None); okw.pop("host_path", None)
    okw["auto_reset"] = int(okw.get("auto_reset", False))
    okw["action_frame_env"] = int(okw.get("action_frame_env", False))
    if "fly_object" in okw:
        okw["fly_object"] = {"banana": 0, "amicelli": 1}[okw["fly_object"]]
    orc = ol.Oracle(ol.make_config(n, task={"peg-in-hole": 0, "random-fly": 1}[task], arm={"panda": 0, "ur5": 1}[arm], **okw),
                    use_double=use_double)
    return sim, orc


def _assert_state_equal(sim, orc, exact=True):
    sg, sc = sim.state.cpu().numpy(), orc.get_state()
    ig, ic = [ol.state_int(sg, w) for w in (9, 10, 11)], [ol.state_int(sc, w) for w in (9, 10, 11)]
    for a, b, name in zip(ig, ic, ("step", "episode", "flags")):
        assert np.array_equal(a, b), f"{name} differs in {np.flatnonzero(a != b)[:8]}"
    assert np.isfinite(sg).all() and np.isfinite(sc).all(), "non-finite state"
    if exact:
        bad = np.argwhere(sg != sc)
        assert bad.size == 0, f"state words differ at {bad[:8]} max|d|={np.abs(sg - sc).max()}"
    else:
        assert np.allclose(sg, sc, atol=1e-5)


def _run_compare(sim, orc, actions, exact=True, reset=True):
    n = sim.n
    if reset:
        og, oc = sim.reset().cpu().numpy(), orc.reset()
        assert np.array_equal(og, oc)
    for t, a in enumerate(actions):
        g_obs, g_rew, g_done = sim.step(torch.from_numpy(a).cuda())
        c_obs, c_rew, c_done = orc.step(a)
        assert np.array_equal(g_done.cpu().numpy(), c_done), f"done differs at step {t}"
        assert np.array_equal(g_rew.cpu().numpy(), c_rew), f"reward differs at step {t}"
        if exact:
            assert np.array_equal(g_obs.cpu().numpy(), c_obs), f"obs differs at step {t}"
        else:
            assert np.allclose(g_obs.cpu().numpy(), c_obs, atol=TOL_POS)
    _assert_state_equal(sim, orc, exact)


def _random_actions(n, adim, steps, offs, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(steps):
        a = rng.uniform(-1, 1, (n, adim)).astype(np.float32)
        a[:, :3] += offs
        out.append(a)
    return out


def test_panda_random_actions_200_steps_bit_exact():
    n = 256
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=11)
    acts = _random_actions(n, 4, 200, sim.env_offsets().numpy(), 1)
    _run_compare(sim, orc, acts)
    flags = sim.int_word(11).cpu().numpy()
    assert ((flags >> 8) & 1).any(), "expected peg-table contacts after 200 steps"


def test_panda_autoreset_time_limit_and_frozen_envs():
    n = 128
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=3, max_steps=25, auto_reset=True)
    acts = _random_actions(n, 4, 80, sim.env_offsets().numpy(), 2)
    _run_compare(sim, orc, acts)
    assert (sim.int_word(10).cpu().numpy() == 3).all()      # 80 steps / 25 -> 3 auto resets, episode index 3
    # reference behaviour: done envs are frozen and keep their outputs (base_env.py:61-71)
    sim2, orc2 = _pair(n, offset=(2.0, 3.0, 0.0), seed=3, max_steps=10, auto_reset=False)
    acts = _random_actions(n, 4, 30, sim2.env_offsets().numpy(), 4)
    _run_compare(sim2, orc2, acts)
    assert sim2.done.cpu().numpy().all() and (sim2.int_word(9).cpu().numpy() == 10).all()


def _scripted_reach_actions(n, steps, peg_xy, offs, close_at, grip):
    """drive the gripper over the peg, descend onto it, close: exercises finger<->peg, arm<->table contacts"""
    out = []
    for t in range(steps):
        a = np.zeros((n, 4), np.float32)
        a[:, 0:2] = peg_xy
        z = 0.25 if t < 60 else -0.02 if t < 150 else -0.1
        a[:, 2] = z
        a[:, 3] = 0.04 if t < close_at else grip
        a[:, :3] += offs
        out.append(a)
    return out


def test_panda_scripted_grasp_contacts_bit_exact():
    n = 64
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=21)
    sim.reset(); orc.reset()
    peg = ol.state_field(orc.get_state(), ol.S_PPOS, 3)[:, :2]
    acts = _scripted_reach_actions(n, 260, peg, sim.env_offsets().numpy(), 170, 0.0)
    _run_compare(sim, orc, acts, reset=False)
    masks = (sim.int_word(11).cpu().numpy() >> 8) & 0xff
    seen = np.bitwise_or.reduce(masks)
    assert seen & 1, "peg-table contact expected"
    assert seen & 16, "arm-table contact expected (gripper pushed into the table)"
    assert seen & (4 | 8), "finger-peg contact expected"


@pytest.mark.parametrize("arm", ["panda", "ur5"])
def test_mixed_contact_types_inside_one_warp_bit_exact(arm):
    """The small-batch kernels choose their solver path per WARP (free-body island before the barrier, arm-only rows,
    general rows when some lane couples its islands).  Interleave envs that grasp the peg (coupled), envs that press
    the arm onto the table away from the peg (arm rows only) and envs that move randomly, so that every warp holds all
    three kinds at once; everything must still equal the oracle bit for bit."""
    n = 96
    sim, orc = _pair(n, arm=arm, offset=(2.0, 3.0, 0.0), seed=31)
    sim.reset(); orc.reset()
    offs = sim.env_offsets().numpy()
    peg = ol.state_field(orc.get_state(), ol.S_PPOS, 3)[:, :2]
    adim = 4 if arm == "panda" else 6
    rng = np.random.default_rng(17)
    kind = np.arange(n) % 3                      # 0 grasp, 1 press on the table, 2 random
    acts = []
    for t in range(240):
        a = rng.uniform(-1, 1, (n, adim)).astype(np.float32)
        z = 0.25 if t < 60 else -0.02 if t < 150 else -0.1
        g = kind == 0
        a[g, 0:2] = peg[g]; a[g, 2] = z
        p_ = kind == 1
        a[p_, 0] = 0.45; a[p_, 1] = 0.25; a[p_, 2] = -0.2 if t > 30 else 0.2
        if arm == "panda":
            a[g, 3] = 0.04 if t < 170 else 0.0
            a[p_, 3] = 0.02
        else:
            a[g | p_, 3:6] = np.array([0.0, 3.14159, 0.0], np.float32)
        a[:, :3] += offs
        acts.append(a)
    _run_compare(sim, orc, acts, reset=False)
    masks = (sim.int_word(11).cpu().numpy() >> 8) & 0xff
    assert (masks[kind == 1] & 16).any(), "arm-table contact expected in the pressing envs"
    if arm == "panda":
        assert (masks[kind == 0] & (4 | 8)).any(), "finger-peg contact expected in the grasping envs"


def test_peg_dropped_into_hole_region_contacts_bit_exact():
    n = 64
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=8)
    sim.reset(); orc.reset()
    w = orc.get_state()
    rng = np.random.default_rng(5)
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    # place the peg just above / inside the annular hole with random tilt and small velocities
    flat[:, ol.S_PPOS:ol.S_PPOS + 3] = np.array([0.5, -0.2, 0.2]) + rng.uniform(-0.03, 0.03, (n, 3)) + np.array([0, 0, 0.03])
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    flat[:, ol.S_PQUAT:ol.S_PQUAT + 4] = q
    flat[:, ol.S_PVEL:ol.S_PVEL + 3] = rng.uniform(-0.2, 0.2, (n, 3))
    w2 = np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32)
    orc.set_state(w2)
    sim.state.copy_(torch.from_numpy(w2))
    acts = _random_actions(n, 4, 120, sim.env_offsets().numpy(), 6)
    for t, a in enumerate(acts):
        g = sim.step(torch.from_numpy(a).cuda())
        c = orc.step(a)
        assert np.array_equal(g[2].cpu().numpy(), c[2]) and np.array_equal(g[1].cpu().numpy(), c[1]), t
        if t == 10:
            assert ((sim.int_word(11).cpu().numpy() >> 8) & 2).any(), "peg-hole contacts expected"
    _assert_state_equal(sim, orc)
    assert np.isfinite(sim.state.cpu().numpy()).all()


def test_ur5_random_actions_bit_exact():
    n = 128
    sim, orc = _pair(n, arm="ur5", offset=(2.0, 3.0, 0.0), seed=4, max_steps=60, auto_reset=True)
    acts = _random_actions(n, 6, 150, sim.env_offsets().numpy(), 7)
    _run_compare(sim, orc, acts)


def test_random_fly_bit_exact():
    n = 128
    sim, orc = _pair(n, task="random-fly", fly_object="banana", dt=1.0 / 120.0, offset=(2.0, 3.0, 0.0), seed=9, max_steps=100,
                     auto_reset=True)
    acts = _random_actions(n, 4, 160, sim.env_offsets().numpy(), 8)
    _run_compare(sim, orc, acts)
    assert (sim.int_word(10).cpu().numpy() >= 1).all()     # every object reached the table or timed out at least once


def test_rollout_kernel_matches_oracle_rollout():
    n = 200   # not a multiple of the block size: exercises the tail warp
    for frame_env in (False, True):
        sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=77, max_steps=40, auto_reset=True, action_frame_env=frame_env)
        sim.reset(); orc.reset()
        for chunk in (1, 7, 50, 3):
            sim.rollout(chunk)
            orc.rollout(chunk)
        _assert_state_equal(sim, orc)
        torch.cuda.synchronize()
        assert int(sim.counters[0].item()) == n * 61
        assert np.array_equal(sim.obs.cpu().numpy(), orc.obs)


def test_apply_action_ik_targets_and_ee_pose():
    n = 64
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=1)
    sim.reset(); orc.reset()
    offs = sim.env_offsets().numpy()
    pos, rot = sim.ee_pose()
    p_ref, R_ref = ol.fk(0, [0, -0.215, -np.pi / 3, -2.57, 0, 2.356, 2.356, 0, 0], use_double=True)
    assert np.allclose(pos.cpu().numpy() - offs, p_ref, atol=2e-6)
    assert np.allclose(rot.cpu().numpy()[0], R_ref, atol=2e-6)
    a = np.zeros((n, 4), np.float32)
    a[:, :3] = p_ref + np.array([0.1, -0.1, 0.05]) + offs
    qt = sim.ik_targets(torch.from_numpy(a)).cpu().numpy()
    # panda_execute: target = current + clamp(dv) per axis, orientation Euler [0,-pi,0]  (envs/utils.py:60-68)
    tgt = p_ref + np.array([1, -1, 1]) * (2 / 240.0)
    want = ol.ik(0, [0, -0.215, -np.pi / 3, -2.57, 0, 2.356, 2.356, 0, 0], tgt, [0, -1, 0, 0], 20, use_double=True)
    assert np.allclose(qt[0, :7], want, atol=5e-4)
    assert np.allclose(qt[:, 7:], 0.0)


def test_step_host_path_equals_device_path():
    n = 96
    sim_a, _ = _pair(n, offset=(2.0, 3.0, 0.0), seed=2)
    sim_b, _ = _pair(n, offset=(2.0, 3.0, 0.0), seed=2)
    sim_a.reset(); sim_b.reset()
    acts = _random_actions(n, 4, 10, sim_a.env_offsets().numpy(), 3)
    obs, rew, done = np.zeros((n, 5), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
    for a in acts:
        sim_a.step(torch.from_numpy(a).cuda())
        sim_b.step_host(a, obs, rew, done)
    assert np.array_equal(sim_a.obs.cpu().numpy(), obs)
    assert np.array_equal(sim_a.state.cpu().numpy(), sim_b.state.cpu().numpy())
    # page-locked caller buffers: DMA straight from / into the caller's memory (one D2H for obs|reward|done)
    (pa,), pobs, prew, pdone = sim_b.pinned_host_buffers(1)
    for a in acts[:5]:
        sim_a.step(torch.from_numpy(a).cuda())
        pa[...] = a
        sim_b.step_host(pa, pobs, prew, pdone)
    assert np.array_equal(sim_a.obs.cpu().numpy(), pobs) and np.array_equal(sim_a.reward.cpu().numpy(), prew)
    assert np.array_equal(sim_a.done.cpu().numpy(), pdone)


@pytest.mark.parametrize("family", ["fused", "ws", "split"])
def test_zero_copy_host_path_with_frozen_envs_and_fresh_buffers(family):
    """pih_step_host on page-locked buffers lets the kernel read actions from / write results to host memory.  Done
    envs are frozen (base_env.py:61-71): their outputs are re-emitted from the frozen state, so handing in FRESH output
    buffers every step still returns the last outputs.  Checked against the oracle for every kernel family.  The buffers
    are NOT bound (pih_host_bind), so the library re-queries the driver for their device aliases on every call - a
    stale alias of a freed buffer can never be used."""
    n = 100
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=21, max_steps=6, auto_reset=False, kernel_family=family)
    sim.reset(); orc.reset()
    acts = _random_actions(n, 4, 12, sim.env_offsets().numpy(), 5)
    for t, a in enumerate(acts):
        (pa,), pobs, prew, pdone = sim.pinned_host_buffers(1)        # fresh page-locked buffers every step
        pobs[...] = -7.0; prew[...] = -7.0; pdone[...] = 9
        pa[...] = a
        sim.step_host(pa, pobs, prew, pdone)
        oc, rc, dc = orc.step(a)
        assert np.array_equal(pobs, oc) and np.array_equal(prew, rc) and np.array_equal(pdone, dc), t
    assert pdone.all()
    _assert_state_equal(sim, orc)


def test_grasp_label_images_match_oracle():
    """SURVEY 8f rank 3, second half: the label images of random_grasp (peg_in_hole.py:65-98) - CUDA == oracle, exactly,
    before the capture, after it, and for hand-set angles on a non-square image."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 40
    sim = BatchedSim(n, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=3, device=0)
    orc = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    sim.reset(); orc.reset()
    for steps in (0, 545, 700):
        if steps:
            sim.rollout(steps); orc.rollout(steps)
        gl, gp = sim.grasp_labels(300, 300)
        cl, cp = orc.grasp_labels(300, 300)
        assert np.array_equal(gl.cpu().numpy(), cl) and np.array_equal(gp.cpu().numpy(), cp), steps
    _assert_state_equal(sim, orc)
    w = orc.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    flat[:, ol.S_GRASP_ANG] = np.linspace(-3.1, 3.1, n).astype(np.float32)
    w2 = np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2))
    orc.set_state(w2); sim.state.copy_(torch.from_numpy(w2))
    gl, gp = sim.grasp_labels(80, 64)
    cl, cp = orc.grasp_labels(80, 64)
    assert np.array_equal(gl.cpu().numpy(), cl) and np.array_equal(gp.cpu().numpy(), cp)
    import peg_in_hole_gym_b200 as gym
    env = gym.make('peg-in-hole-mp-v0', client=gym.DIRECT, task='peg-in-hole-scripted', mp_num=1, sub_num=4, offset=[2., 3., 0.], shard=False)
    env.reset()
    env.step(env.action_space.sample_array())
    labels, params = env.plugin.random_grasp_labels()
    assert tuple(labels.shape) == (4, 4, 300, 300) and tuple(params.shape) == (4, 5) and bool((labels[:, 0] == 50).any())
    env.close()


def test_peer_write_gather_on_one_gpu_matches_plain_outputs():
    """SURVEY 8e preferred form: the step kernel writes its rows into the learner buffer itself and publishes a flag;
    the owner waits on the flags on the device.  Exercised here with two members inside one process on one GPU (across
    processes the buffer is mapped through CUDA IPC; bench.py --gather peer runs that on >= 2 GPUs)."""
    from peg_in_hole_gym_b200.parallel import PeerLearnerBuffer
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 160
    kw = dict(offset=(2.0, 3.0, 0.0), max_steps=12, auto_reset=True, seed=9, sub_num=2 * n)
    sims = [BatchedSim(n, env_id_base=r * n, device=0, **kw) for r in range(2)]
    refs = [BatchedSim(n, env_id_base=r * n, device=0, **kw) for r in range(2)]
    for s_ in sims + refs:
        s_.reset()
    buf = torch.zeros(PeerLearnerBuffer.total_bytes(n, 5, 2), dtype=torch.uint8, device="cuda:0")
    members = [PeerLearnerBuffer(sims[r], rank=r, world_size=2, buf=buf) for r in range(2)]
    rng = np.random.default_rng(3)
    for t in range(30):
        acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() + torch.cat(
            [sims[r].env_offsets().cuda(), torch.zeros(n, 1, device="cuda")], 1) for r in range(2)]
        members[1].step(acts[1])                      # the "remote" rank first, then the owner
        members[0].step(acts[0])
        members[0].step_no = members[1].step_no
        rows = members[0].wait()
        for r in range(2):
            refs[r].step(acts[r])
            assert torch.equal(rows[r][0], refs[r].obs) and torch.equal(rows[r][1], refs[r].reward)
            assert torch.equal(rows[r][2], refs[r].done)
    members[0].check()
    # deferred publication (pih_step_publish_prev): a step only stores its rows, the NEXT launch announces them with one
    # fence at its start; the owner's folded wait looks for the previous step's flags, flush() announces the last step
    flags = buf[members[0].flag_off:members[0].flag_off + 8].view(torch.int32)
    for t in range(12):
        acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() + torch.cat(
            [sims[r].env_offsets().cuda(), torch.zeros(n, 1, device="cuda")], 1) for r in range(2)]
        members[1].step(acts[1], deferred=True)
        members[0].step_no = members[1].step_no - 1
        members[0].step(acts[0], deferred=True, wait_prev=True)
        torch.cuda.synchronize()
        assert flags.tolist() == [members[0].step_no - 1] * 2, (t, flags.tolist())      # announced so far: the previous step
        for r in range(2):
            refs[r].step(acts[r])
    for m in members:
        m.flush()
    rows = members[0].wait()
    torch.cuda.synchronize()
    assert flags.tolist() == [members[0].step_no] * 2
    for r in range(2):
        assert torch.equal(rows[r][0], refs[r].obs) and torch.equal(rows[r][1], refs[r].reward) and torch.equal(rows[r][2], refs[r].done)
    members[0].check()
    # a flag that never arrives: the bounded wait gives up and reports which one
    members[0].step_no += 1
    members[0].wait(timeout_ms=20)
    with pytest.raises(RuntimeError):
        members[0].check()


def test_peer_write_gather_across_processes():
    """Two rank processes share rank 0's learner buffer through CUDA IPC (pih_peer_alloc / pih_peer_open) and the step
    kernels of both write into it; rank 0 checks every gathered row against locally recomputed steps.  Both ranks sit on
    cuda:0 here (the driver's GPU tier has one GPU); on a multi-GPU node the same code maps the buffer as NVLink peer
    memory (bench.py --gather peer)."""
    import subprocess
    import sys
    import socket
    here = os.path.dirname(os.path.abspath(__file__))
    with socket.socket() as sk:                      # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(here, "peer_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert "PEER_GATHER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_fp32_gpu_vs_fp64_oracle_drift_within_tolerance():
    """north_star tolerance: <=1e-3 m / 1e-2 rad per step; drift over the 200-step horizon is reported."""
    n = 128
    sim, orc = _pair(n, use_double=True, offset=(2.0, 3.0, 0.0), seed=13)
    sim.reset(); orc.reset()
    acts = _random_actions(n, 4, 200, sim.env_offsets().numpy(), 12)
    worst_q = worst_p = 0.0
    for t, a in enumerate(acts):
        g_obs, g_rew, g_done = sim.step(torch.from_numpy(a).cuda())
        c_obs, c_rew, c_done = orc.step(a)
        sg, sc = sim.state.cpu().numpy(), orc.get_state()
        dq = np.abs(ol.state_field(sg, ol.S_Q, 9) - ol.state_field(sc, ol.S_Q, 9)).max()
        dp = np.abs(ol.state_field(sg, ol.S_PPOS, 3) - ol.state_field(sc, ol.S_PPOS, 3)).max()
        de = np.abs(g_obs.cpu().numpy() - c_obs).max()
        worst_q, worst_p = max(worst_q, dq), max(worst_p, dp, de)
        assert np.array_equal(g_done.cpu().numpy(), c_done)
    print(f"fp32 GPU vs fp64 oracle over 200 steps: max |dq| = {worst_q:.3e} rad, max |dpos| = {worst_p:.3e} m")
    assert worst_q < TOL_ANG and worst_p < TOL_POS


def test_full_size_properties_4096_envs():
    """BASELINE config 2 size: determinism, shard-independence (RNG keyed by global env id) and invariants."""
    n = 4096
    kw = dict(offset=(2.0, 3.0, 0.0), seed=123, max_steps=200, auto_reset=True)
    a = _sim(n, **kw); a.reset(); a.rollout(210)
    b = _sim(n, **kw); b.reset(); b.rollout(100); b.rollout(110)
    assert torch.equal(a.state, b.state)
    # shard independence: envs [1024, 2048) simulated alone (as rank 1 of 4 would) give the same trajectories
    c = _sim(1024, sub_num=n, env_id_base=1024, **kw); c.reset(); c.rollout(210)
    assert torch.equal(a.state[:, 1024:2048], c.state)
    st = a.state.cpu().numpy()
    assert np.isfinite(st).all()
    quat = ol.state_field(st, ol.S_PQUAT, 4)
    assert np.allclose(np.linalg.norm(quat, axis=1), 1.0, atol=1e-5)
    assert (ol.state_field(st, ol.S_PPOS, 3)[:, 2] > -0.06).all()          # nothing fell through the table
    q = ol.state_field(st, ol.S_Q, 9)
    assert (q[:, 7:] > -1e-3).all() and (q[:, 7:] < 0.041).all()           # finger limits hold
    assert (ol.state_int(st, 10) == 1).all() and (ol.state_int(st, 9) == 10).all()
    assert int(a.counters[0].item()) == n * 210


@pytest.mark.parametrize("case", ["panda_peg", "ur5_peg", "panda_fly", "panda_pipe"])
def test_cuda_path_reproduces_golden_fixtures(case):
    """Committed fixtures (tests/golden/, generated by gen_golden.py from the oracle): CUDA path must match them."""
    import os
    from golden.gen_golden import CASES, N, STEPS
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    kw = dict(CASES[case])
    task = {0: "peg-in-hole", 1: "random-fly"}[kw.pop("task")]
    arm = {0: "panda", 1: "ur5"}[kw.pop("arm")]
    if "fly_object" in kw:
        kw["fly_object"] = {0: "banana", 1: "amicelli"}[kw["fly_object"]]
    kw["auto_reset"] = bool(kw["auto_reset"])
    if "peg_model" in kw:
        kw["peg_model"] = {0: "rigid", 1: "pipe"}[kw["peg_model"]]
    sim = _sim(N, task=task, arm=arm, **kw)
    assert np.array_equal(sim.reset().cpu().numpy(), g["obs0"])
    for t in range(STEPS):
        obs, r, d = sim.step(torch.from_numpy(g["actions"][t]).cuda())
        assert np.array_equal(d.cpu().numpy(), g["done"][t]) and np.array_equal(r.cpu().numpy(), g["reward"][t]), t
        assert np.array_equal(obs.cpu().numpy(), g["obs"][t]), t
    assert np.array_equal(sim.state.cpu().numpy(), g["state"])


def test_gym_surface_end_to_end():
    """README usage loop (README.md:33-50) through the registered id, nested-list and tensor containers."""
    import peg_in_hole_gym_b200 as gym
    env = gym.make('peg-in-hole-mp-v0', client=gym.GUI, task='peg-in-hole', mp_num=2, sub_num=3, offset=[2., 3., 0.],
                   args=None, is_test=True, max_steps=5, seed=1, shard=False)
    with pytest.raises(AttributeError):
        env.step(env.action_space.sample())
    obs = env.reset()
    assert tuple(obs.shape) == (2, 3, 5)
    for t in range(5):
        obs, reward, done, info = env.step(env.action_space.sample())      # nested [mp][sub] lists of Box samples
    assert tuple(obs.shape) == (2, 3, 5) and tuple(reward.shape) == (2, 3) and done.dtype == torch.bool
    assert bool(done.all()) and info[1][2]["success"] is False
    frozen = obs.clone()
    obs2, _, _, _ = env.step(env.action_space.sample_array())
    assert torch.equal(obs2, frozen)                                      # done envs keep their outputs
    env.close()
    env = gym.make('peg-in-hole-v0', client=gym.DIRECT, task='random-fly', task_num=4, offset=[2., 0., 0.],
                   args=['Banana', 1 / 120.], as_lists=True)
    obs = env.reset()
    o, r, d, i = env.step(env.action_space.sample())
    assert isinstance(o, list) and len(o) == 4 and len(o[0]) == 5 and isinstance(i[0], dict)
    env.close()


def test_scripted_grasp_fsm_bit_exact_and_through_gym_surface():
    """the scripted state machine, 2226 engine steps, CUDA vs oracle bit for bit."""
    n = 32
    sim, orc = _pair(n, task="peg-in-hole", offset=(2.0, 3.0, 0.0), seed=3)
    sim.close()
    from peg_in_hole_gym_b200.sim import BatchedSim
    sim = BatchedSim(n, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=3, device=0)
    orc = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    assert np.array_equal(sim.reset().cpu().numpy(), orc.reset())
    for chunk in (500, 800, 926):
        sim.rollout(chunk)
        orc.rollout(chunk)
        _assert_state_equal(sim, orc)
    assert sim.done.cpu().numpy().all() and np.array_equal(sim.reward.cpu().numpy(), orc.reward)
    assert ((sim.int_word(11).cpu().numpy() >> 8) & 0xff).any()
    import peg_in_hole_gym_b200 as gym
    env = gym.make('peg-in-hole-mp-v0', client=gym.DIRECT, task='peg-in-hole-scripted', mp_num=2, sub_num=16, offset=[2., 3., 0.],
                   seed=3, shard=False)
    env.reset()
    obs, reward, done, info = env.step(env.action_space.sample_array())      # one step() = one scripted episode
    assert bool(done.all()) and np.array_equal(reward.cpu().numpy().ravel(), orc.reward)
    env.close()


@pytest.mark.parametrize("family", ["ws", "ws2", "fused"])
def test_scripted_episode_bit_exact_in_every_kernel_family(family):
    """The scripted episode through its grasp and lift phases (grasp constraint + finger contacts in every env: the
    coupled contact rows, which the scripted small-batch kernels solve from two alternating sets of coefficients with
    M^-1 in shared memory, and the contacts warp's narrow-phase skip) in the three-warp, the two-warp and the fused
    kernel, against the oracle bit for bit.  100 envs: padding lanes in the last CTA."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 100
    sim = BatchedSim(n, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=11, device=0, kernel_family=family)
    orc = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=11))
    assert np.array_equal(sim.reset().cpu().numpy(), orc.reset())
    for chunk in (1100, 300, 400, 426):                  # approach + close | grasp on | lift | carry, insert, release
        sim.rollout(chunk)
        orc.rollout(chunk)
        _assert_state_equal(sim, orc)
    assert sim.done.cpu().numpy().all() and np.array_equal(sim.reward.cpu().numpy(), orc.reward)
    sim.close()


def test_camera_render_matches_oracle_and_full_size():
    """PegInHole.render as a batched ray caster; CUDA image == oracle image, then 300x300."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 6
    sim = BatchedSim(n, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=3, device=0)
    orc = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    sim.reset(); orc.reset()
    for steps in (560, 600, 900):                            # hovering over the peg / grasping / at the hole
        sim.rollout(steps); orc.rollout(steps)
        g = sim.render(80, 64).cpu().numpy()
        c = orc.render(80, 64)
        assert np.array_equal(g[..., 1:], c[..., 1:]), "rgb differs"
        assert np.array_equal(g[..., 0], c[..., 0]), "depth differs"
    img = sim.render()                                       # reference resolution (peg_in_hole.py:24-26)
    assert tuple(img.shape) == (n, 300, 300, 4)
    assert float(img[..., 0].min()) > 0.0 and float(img[..., 0].max()) <= 1.0
    import peg_in_hole_gym_b200 as gym
    env = gym.make('peg-in-hole-mp-v0', client=gym.DIRECT, task='peg-in-hole', mp_num=2, sub_num=2, offset=[2., 3., 0.], shard=False)
    env.reset()
    assert tuple(env.render('rgb_array').shape) == (2, 2, 300, 300, 4)
    env.close()


def test_grasp_image_is_the_scene_latched_on_entering_state_2():
    """the reference returns `grasp_img`, rendered when the state machine enters state 2, together with the
    grasp labels (peg_in_hole.py:33-37,65-66).  The scene is latched into the state at that step; pih_render_grasp casts it.
    CUDA == oracle exactly; the image taken at the end of the episode (arm parked) is a different one."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 6
    sim = BatchedSim(n, task="peg-in-hole-scripted", offset=(2.0, 3.0, 0.0), seed=3, device=0)
    orc = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    sim.reset(); orc.reset()
    sim.rollout(300); orc.rollout(300)                        # state 1: nothing captured yet -> the present scene
    assert np.array_equal(sim.render_grasp(80, 64).cpu().numpy(), sim.render(80, 64).cpu().numpy())
    sim.rollout(241); orc.rollout(241)                        # step 541 = first step of state 2 (59 + 481 + 1)
    at_capture = sim.render(80, 64).cpu().numpy()             # scene before that step's physics = what was latched ...
    _assert_state_equal(sim, orc)
    sim.rollout(1685); orc.rollout(1685)                      # ... to the end of the episode
    _assert_state_equal(sim, orc)
    g, c = sim.render_grasp(80, 64).cpu().numpy(), orc.render_grasp(80, 64)
    assert np.array_equal(g, c)
    assert not np.array_equal(g, sim.render(80, 64).cpu().numpy()), "the latched image must not be the end-of-episode scene"
    # the latch holds the pre-step scene of step 541; one engine step later the rendered scene has barely moved
    assert np.abs(g[..., 0] - at_capture[..., 0]).mean() < 5e-3
    import peg_in_hole_gym_b200 as gym
    env = gym.make('peg-in-hole-mp-v0', client=gym.DIRECT, task='peg-in-hole-scripted', mp_num=1, sub_num=n, offset=[2., 3., 0.], seed=3, shard=False)
    env.reset()
    env.step(env.action_space.sample_array())
    img = env.render('rgb_array', 80, 64)
    assert np.array_equal(img.cpu().numpy().reshape(n, 64, 80, 4), g)
    env.close()


@pytest.mark.parametrize("arm,task,fly", [("panda", "peg-in-hole", "banana"), ("ur5", "peg-in-hole", "banana"),
                                          ("panda", "random-fly", "amicelli")])
def test_randomised_states_fuzz_bit_exact(arm, task, fly):
    """Fuzz: random (even unphysical) joint states, velocities and free-body poses injected into BOTH engines, then
    40 steps of random actions - exercises limits, deep penetrations, table/hole/finger contacts, velocity clamps."""
    n = 192
    dt = 1.0 / 120.0 if task == "random-fly" else 1.0 / 240.0
    sim, orc = _pair(n, arm=arm, task=task, fly_object=fly, dt=dt, offset=(2.0, 3.0, 0.0), seed=31, max_steps=25, auto_reset=True)
    sim.reset(); orc.reset()
    rng = np.random.default_rng(17)
    w = orc.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    nq = 9 if arm == "panda" else 6
    flat[:, ol.S_Q:ol.S_Q + nq] += rng.normal(0, 0.6, (n, nq)).astype(np.float32)
    if arm == "panda":
        flat[:, ol.S_Q + 7:ol.S_Q + 9] = rng.uniform(-0.01, 0.05, (n, 2))
    flat[:, ol.S_QD:ol.S_QD + nq] = rng.normal(0, 2.0, (n, nq))
    # free body near the gripper / table / hole, random orientation and twist
    ee = orc.obs[:, 2:5] - sim.env_offsets().numpy()
    where = rng.integers(0, 3, n)
    pos = np.where(where[:, None] == 0, ee + rng.uniform(-0.04, 0.04, (n, 3)),
                   np.where(where[:, None] == 1, np.array([0.5, -0.2, 0.2]) + rng.uniform(-0.04, 0.04, (n, 3)),
                            np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.7, -0.3, n), rng.uniform(-0.07, 0.05, n)], 1)))
    flat[:, ol.S_PPOS:ol.S_PPOS + 3] = pos
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    flat[:, ol.S_PQUAT:ol.S_PQUAT + 4] = q
    flat[:, ol.S_PVEL:ol.S_PVEL + 3] = rng.normal(0, 1.0, (n, 3))
    flat[:, ol.S_PANG:ol.S_PANG + 3] = rng.normal(0, 10.0, (n, 3))
    w2 = np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32)
    orc.set_state(w2)
    sim.state.copy_(torch.from_numpy(w2))
    acts = _random_actions(n, 4 if arm == "panda" else 6, 40, sim.env_offsets().numpy(), 5)
    seen = 0
    for t, a in enumerate(acts):
        g = sim.step(torch.from_numpy(a).cuda())
        c = orc.step(a)
        assert np.array_equal(g[2].cpu().numpy(), c[2]) and np.array_equal(g[1].cpu().numpy(), c[1]), t
        seen |= int(np.bitwise_or.reduce(sim.int_word(11).cpu().numpy() >> 8))
    _assert_state_equal(sim, orc)
    assert seen & 1, "table contacts expected"


@pytest.mark.parametrize("arm,task", [("panda", "peg-in-hole"), ("ur5", "peg-in-hole"), ("panda", "random-fly"),
                                      ("panda", "peg-in-hole-scripted")])
def test_all_kernel_families_agree_with_the_oracle(arm, task):
    """Kernel families that compute one env.step: the warp-specialised small-batch kernel k_step_ws with three warps per
    32 envs (action / dynamics / contacts; "ws") and with two ("ws2", batches of 9473..18944 envs; pih_rollout on both =
    chained launches of the same kernel with in-kernel actions),
    the fused thread-per-env kernels (k_step / k_rollout) and the split pair k_apply_action + k_step_dyn that batches
    >= 49152 envs (BASELINE config 5) run.  pih_config.kernel_family forces each of them at a test-sized batch; all
    must give the oracle's bits, for pih_step with EXTERNAL actions and for pih_rollout.  (The scripted task has no
    split form: its state machine lives in apply_action.)"""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n = 100                                              # not a multiple of 32: padding lanes in the last CTA
    kw = dict(task=task, arm=arm, offset=(2.0, 3.0, 0.0), seed=6, max_steps=0 if "scripted" in task else 30,
              auto_reset="scripted" not in task, dt=1.0 / 120.0 if task == "random-fly" else 1.0 / 240.0)
    families = ["ws", "ws2", "fused"] + ([] if "scripted" in task else ["split"])
    sims = {f: BatchedSim(n, device=0, kernel_family=f, **kw) for f in families}
    for sm in sims.values():
        sm.reset()
    orc = ol.Oracle(ol.make_config(n, task={"peg-in-hole": 0, "random-fly": 1, "peg-in-hole-scripted": 2}[task],
                                   arm={"panda": 0, "ur5": 1}[arm], offset=kw["offset"], seed=6, max_steps=kw["max_steps"],
                                   auto_reset=int(kw["auto_reset"]), dt=kw["dt"]))
    orc.reset()
    acts = _random_actions(n, 4 if arm == "panda" else 6, 45, sims["ws"].env_offsets().numpy(), 9)
    for t, a in enumerate(acts):                         # per-step launches with external actions
        c = orc.step(a)
        for f, sm in sims.items():
            g = sm.step(torch.from_numpy(a).cuda())
            assert np.array_equal(g[0].cpu().numpy(), c[0]) and np.array_equal(g[1].cpu().numpy(), c[1]), (f, t)
            assert np.array_equal(g[2].cpu().numpy(), c[2]), (f, t)
    for sm in sims.values():
        _assert_state_equal(sm, orc)
    for sm in sims.values():                             # then the rollout kernels
        sm.rollout(70)
    orc.rollout(70)
    for f, sm in sims.items():
        _assert_state_equal(sm, orc)
        assert np.array_equal(sm.obs.cpu().numpy(), orc.obs), f


def test_config5_size_panda_peg_split_kernels():
    """BASELINE config 5's per-GPU shape: panda peg-in-hole at 65536 envs on one GPU, where pih_create picks the split
    pair k_apply_action<panda, peg> + k_step_dyn<panda, peg> by itself.  pih_step with external actions and pih_rollout,
    determinism across launch partitions, invariants, and an oracle check on a slice with the same global env ids."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    n, m = 65536, 128
    kw = dict(offset=(2.0, 3.0, 0.0), seed=55, max_steps=40, auto_reset=True)
    a = BatchedSim(n, device=0, **kw); a.reset()
    orc = ol.Oracle(ol.make_config(m, sub_num=n, offset=kw["offset"], seed=55, max_steps=40, auto_reset=1)); orc.reset()
    offs = torch.zeros(n, 4, device="cuda")
    offs[:, :3] = a.env_offsets_device()
    assert torch.equal(offs[:300, :3].cpu(), BatchedSim(300, sub_num=n, device=0, **kw).env_offsets())
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for t in range(50):                                  # crosses one auto-reset
        act = (torch.rand(n, 4, device="cuda", generator=g) * 2 - 1) + offs
        go, gr, gd = a.step(act)
        co, cr, cd = orc.step(act[:m].cpu().numpy())
        assert np.array_equal(go[:m].cpu().numpy(), co) and np.array_equal(gr[:m].cpu().numpy(), cr), t
        assert np.array_equal(gd[:m].cpu().numpy(), cd), t
    a.rollout(60); orc.rollout(60)
    st = a.state.cpu().numpy()
    assert np.array_equal(st[:, :m], orc.get_state())
    assert np.isfinite(st).all()
    assert np.allclose(np.linalg.norm(ol.state_field(st, ol.S_PQUAT, 4), axis=1), 1.0, atol=1e-5)
    assert (ol.state_int(st, 10) == 2).all() and int(a.counters[0].item()) == n * 60
    b = BatchedSim(n, device=0, kernel_family="fused", **kw); b.reset()     # the fused kernel on the same batch
    g.manual_seed(5)
    for t in range(50):
        b.step((torch.rand(n, 4, device="cuda", generator=g) * 2 - 1) + offs)
    b.rollout(25); b.rollout(35)
    assert torch.equal(a.state, b.state)


def test_nonfinite_state_trips_the_fault_guard_and_resets():
    """a non-finite word in an env's state must not live on in HBM.  NaN / Inf are injected into
    different state words of some envs of BOTH engines; after one step those envs are re-initialised (new episode),
    reported done with reward 0 and carry PIH_FLAG_FAULT; every other env is untouched; afterwards the engines keep
    agreeing bit for bit.  Exercised for every kernel family, with and without auto_reset."""
    n = 96
    for family, auto in (("ws", True), ("fused", False), ("split", True)):
        sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=14, max_steps=50, auto_reset=auto, kernel_family=family)
        acts = _random_actions(n, 4, 30, sim.env_offsets().numpy(), 3)
        _run_compare(sim, orc, acts[:10])
        w = orc.get_state()
        flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
        bad = np.arange(3, n, 7)
        words = [ol.S_Q + 2, ol.S_QD + 5, ol.S_PPOS + 1, ol.S_PQUAT + 3, ol.S_PVEL, ol.S_PANG + 2]
        for k, e in enumerate(bad):
            wd = words[k % len(words)]
            flat[e, wd] = np.nan if (k % 2 == 0 or wd == ol.S_QD + 5) else np.inf   # (an infinite joint RATE alone is clamped to 100)
        w2 = np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32)
        orc.set_state(w2); sim.state.copy_(torch.from_numpy(w2))
        stats = sim.stats()
        go, gr, gd = sim.step(torch.from_numpy(acts[10]).cuda())
        co, cr, cd = orc.step(acts[10])
        assert np.array_equal(gd.cpu().numpy(), cd) and np.array_equal(gr.cpu().numpy(), cr)
        assert np.array_equal(go.cpu().numpy(), co)
        flags = sim.int_word(11).cpu().numpy()
        assert (flags[bad] & 4).all() and gd.cpu().numpy()[bad].all() and (gr.cpu().numpy()[bad] == 0).all()
        ok = np.setdiff1d(np.arange(n), bad)
        assert not (flags[ok] & 4).any()
        assert int(stats[6].item()) == len(bad) and int(stats[0].item()) == n
        assert np.isfinite(sim.state.cpu().numpy()).all()
        _assert_state_equal(sim, orc)
        sim.stats(False)
        _run_compare(sim, orc, acts[11:], reset=False)


def test_execution_counters():
    """pih_set_stats: the counters behind the bench line's executed-iteration figures."""
    n = 256
    sim, orc = _pair(n, offset=(2.0, 3.0, 0.0), seed=2, max_steps=200, auto_reset=True)
    sim.reset()
    sim.rollout(40, count=False)
    st = sim.stats()
    sim.rollout(25, count=False)
    torch.cuda.synchronize()
    v = st.cpu().numpy()
    assert v[0] == n * 25
    assert 0 < v[1] <= 20 * v[0] and v[3] >= v[0] and v[3] <= 50 * v[0] and v[7] <= 50
    assert v[2] >= 9 * v[0]                          # at least the 9 motor rows once per env-step
    assert v[6] == 0


@pytest.mark.parametrize("n,arm,task,fly,dt", [(16384, "ur5", "peg-in-hole", "banana", 1.0 / 240.0),      # BASELINE config 3
                                                (65536, "panda", "random-fly", "banana", 1.0 / 120.0)])    # BASELINE config 4
def test_full_size_properties_other_configs(n, arm, task, fly, dt):
    """Full BASELINE sizes of configs 3 and 4 (classic kernels): determinism, split-rollout equivalence, invariants,
    and an oracle spot check on a slice of the envs."""
    from peg_in_hole_gym_b200.sim import BatchedSim
    kw = dict(arm=arm, task=task, fly_object=fly, dt=dt, offset=(2.0, 3.0, 0.0), seed=77, max_steps=60, auto_reset=True)
    a = BatchedSim(n, device=0, **kw); a.reset(); a.rollout(90)
    b = BatchedSim(n, device=0, **kw); b.reset(); b.rollout(40); b.rollout(50)
    assert torch.equal(a.state, b.state)
    st = a.state.cpu().numpy()
    assert np.isfinite(st).all()
    assert np.allclose(np.linalg.norm(ol.state_field(st, ol.S_PQUAT, 4), axis=1), 1.0, atol=1e-5)
    assert (ol.state_int(st, 10) >= 1).all() and int(a.counters[0].item()) == n * 90
    m = 96                                                   # oracle on the first 96 envs (same global ids)
    orc = ol.Oracle(ol.make_config(m, sub_num=n, arm={"panda": 0, "ur5": 1}[arm], task={"peg-in-hole": 0, "random-fly": 1}[task],
                                   dt=dt, offset=(2.0, 3.0, 0.0), seed=77, max_steps=60, auto_reset=1))
    orc.reset(); orc.rollout(90)
    assert np.array_equal(st[:, :m], orc.get_state())
