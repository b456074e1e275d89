"""Closed-form physics and episode-semantics checks of the CPU oracle (CPU only)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

DT = 1.0 / 240.0


def _set(o, n, **fields):
    w = o.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    for word, vals in fields.values():
        vals = np.asarray(vals, np.float32)
        flat[:, word:word + vals.shape[-1]] = vals
    o.set_state(np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32))


def _hold_action(o):
    # action = current EE position keeps the arm (nearly) still; fingers open
    a = np.zeros((o.n, 4), np.float32)
    a[:, :3] = o.obs[:, 2:5]
    a[:, 3] = 0.04
    return a


@pytest.mark.parametrize("dbl", [False, True])
def test_free_fall_matches_semi_implicit_euler(dbl):
    o = ol.Oracle(ol.make_config(1, seed=1), use_double=dbl)
    o.reset()
    z0 = ol.state_field(o.get_state(), ol.S_PPOS, 3)[0, 2]
    for k in range(1, 31):
        o.step(_hold_action(o))
        z = ol.state_field(o.get_state(), ol.S_PPOS, 3)[0, 2]
        assert abs(z - (z0 - 9.8 * DT * DT * k * (k + 1) / 2)) < 2e-6


def test_peg_comes_to_rest_on_table_without_penetration():
    o = ol.Oracle(ol.make_config(4, offset=(2.0, 3.0, 0.0), seed=2))
    o.reset()
    for _ in range(200):
        o.step(_hold_action(o))
    w = o.get_state()
    z = ol.state_field(w, ol.S_PPOS, 3)[:, 2]
    assert np.allclose(z, -0.05 + 0.01, atol=2e-4)                       # table top + peg radius, slop << 1 mm
    assert np.abs(ol.state_field(w, ol.S_PVEL, 3)).max() < 1e-3
    assert ((ol.state_int(w, ol.S_FLAGS) >> 8) & 1).all()               # peg-table pair present


def test_spinning_peg_rolls_without_slipping():
    o = ol.Oracle(ol.make_config(1, seed=1), use_double=True)
    o.reset()
    _set(o, 1, pos=(ol.S_PPOS, [[0.1, -0.5, -0.04]]), ang=(ol.S_PANG, [[0.0, 50.0, 0.0]]))
    for _ in range(20):
        o.step(_hold_action(o))
    w = o.get_state()
    v, om = ol.state_field(w, ol.S_PVEL, 3)[0], ol.state_field(w, ol.S_PANG, 3)[0]
    assert abs(v[0] - om[1] * 0.01) < 1e-4 and abs(om[1]) > 1.0           # v = omega x r at the contact


def test_motor_reaches_ik_target_in_one_step_kp1():
    """POSITION_CONTROL with positionGains=1 (utils.py:68): the joints land on the IK result in exactly one step."""
    o = ol.Oracle(ol.make_config(1, seed=1), use_double=True)
    o.reset()
    for _ in range(30):                                                    # let the initial re-orientation settle
        o.step(_hold_action(o))
    q0 = ol.state_field(o.get_state(), ol.S_Q, 9)[0].astype(np.float64)
    p0 = o.obs[0, 2:5].astype(np.float64)
    a = _hold_action(o)
    a[0, :3] = p0 + np.array([0.3, 0.0, 0.0])                              # far target: vel_constraint clamps to dv
    p_cur, _ = ol.fk(0, q0, True)
    tgt = p_cur + np.array([2 / 240.0, 0.0, 0.0])
    tgt[1:] = p_cur[1:] + np.clip(a[0, 1:3] - p_cur[1:], -2 / 240.0, 2 / 240.0)
    want = ol.ik(0, q0, tgt, [0, -1, 0, 0], 20, True)
    o.step(a)
    q1 = ol.state_field(o.get_state(), ol.S_Q, 9)[0]
    assert np.allclose(q1[:7], want, atol=2e-5)
    moved = o.obs[0, 2:5] - p0
    # BussIK's lambda=0.6 leaves a residual after 20 iterations (PyBullet's default IK is known to be loose)
    assert 0.5 * 2 / 240.0 < moved[0] <= 2 / 240.0 + 1e-5


def test_finger_limits_hold_for_out_of_range_targets():
    o = ol.Oracle(ol.make_config(16, offset=(2.0, 3.0, 0.0), seed=3))
    o.reset()
    rng = np.random.default_rng(0)
    for _ in range(60):
        a = _hold_action(o)
        a[:, 3] = rng.uniform(-1, 1, o.n)
        o.step(a)
        q = ol.state_field(o.get_state(), ol.S_Q, 9)[:, 7:]
        assert q.min() > -1e-4 and q.max() < 0.0401


def test_time_limit_done_frozen_then_reset():
    cfg = ol.make_config(3, offset=(2.0, 0.0, 0.0), seed=5, max_steps=7)
    o = ol.Oracle(cfg)
    with pytest.raises(RuntimeError):
        o.step(np.zeros((3, 4), np.float32))                               # step before reset is an error
    o.reset()
    for t in range(7):
        obs, r, d = o.step(_hold_action(o))
        assert d.all() == (t == 6)
    frozen = obs.copy()
    w = o.get_state()
    a = _hold_action(o); a[:, 0] += 0.5
    obs2, r2, d2 = o.step(a)                                               # done envs are not stepped (base_env.py:61-71)
    assert np.array_equal(obs2, frozen) and d2.all() and np.array_equal(o.get_state(), w)
    o.reset()
    assert not o.done.any() and (ol.state_int(o.get_state(), ol.S_EPISODE) == 1).all()
    # masked reset only touches the selected envs
    for _ in range(7):
        o.step(_hold_action(o))
    o.reset(mask=np.array([1, 0, 1], np.uint8))
    assert list(ol.state_int(o.get_state(), ol.S_EPISODE)) == [2, 1, 2]


def test_auto_reset_reset_distribution_and_determinism():
    n = 512
    cfg = ol.make_config(n, offset=(2.0, 3.0, 0.0), seed=9, max_steps=5, auto_reset=1)
    a, b = ol.Oracle(cfg), ol.Oracle(cfg)
    a.reset(); b.reset()
    a.rollout(12); b.rollout(5); b.rollout(7)
    assert np.array_equal(a.get_state(), b.get_state())
    w = a.get_state()
    assert (ol.state_int(w, ol.S_EPISODE) == 2).all() and (ol.state_int(w, ol.S_STEP) == 2).all()
    a.reset()
    p = ol.state_field(a.get_state(), ol.S_PPOS, 3)
    assert p[:, 0].min() >= -0.2 and p[:, 0].max() <= 0.2 and p[:, 1].min() >= -0.6 and p[:, 1].max() <= -0.4
    assert np.allclose(p[:, 2], 0.11) and p[:, 0].std() > 0.08            # peg_in_hole.py:239
    q = ol.state_field(a.get_state(), ol.S_Q, 9)
    assert np.allclose(q[0], [0, -0.215, -np.pi / 3, -2.57, 0, 2.356, 2.356, 0, 0], atol=1e-6)  # peg_in_hole.py:233


def test_env_offsets_follow_create_env():
    """BaseEnv._create_env (base_env.py:35-55): line layout when offset.x or offset.y is 0, else ceil(sqrt(n)) grid."""
    def ref(offset, n):
        import math
        out = []
        if offset[0] == 0 or offset[1] == 0:
            return [np.array(offset) * i for i in range(n)]
        s = int(math.ceil(math.sqrt(n)))
        for i in range(s):
            for j in range(s):
                out.append(np.array([offset[0] * i, offset[1] * j, offset[2]]))
                if len(out) >= n:
                    return out
    for offset, n in [((2.0, 3.0, 0.0), 10), ((2.0, 3.0, 0.5), 16), ((0.0, 3.0, 1.0), 5), ((2.0, 0.0, 0.0), 7), ((2.0, 3.0, 0.0), 4096)]:
        cfg = ol.make_config(n, offset=offset)
        want = ref(offset, n)
        buf = (C.c_float * 3)()
        for i in range(0, n, max(1, n // 50)):
            ol.lib().pih_oracle_env_offset(C.byref(cfg), i, buf)
            assert np.allclose(list(buf), want[i])


def test_success_reward_when_peg_is_at_the_hole():
    o = ol.Oracle(ol.make_config(2, offset=(2.0, 0.0, 0.0), seed=1))
    o.reset()
    _set(o, 2, pos=(ol.S_PPOS, [[0.5, -0.2, 0.24], [0.5, -0.2, 0.40]]))   # env 0 within 5 cm of the hole, env 1 not
    obs, r, d = o.step(_hold_action(o))
    assert r[0] == 1.0 and d[0] == 1 and r[1] == 0.0 and d[1] == 0           # peg_in_hole.py:114-117


@pytest.mark.parametrize("case", ["panda_peg", "ur5_peg", "panda_fly", "panda_pipe"])
def test_oracle_reproduces_golden_fixtures(case):
    from golden.gen_golden import CASES, N, STEPS
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    o = ol.Oracle(ol.make_config(N, **CASES[case]))
    assert np.array_equal(o.reset(), g["obs0"])
    for t in range(STEPS):
        obs, r, d = o.step(g["actions"][t])
        assert np.array_equal(d, g["done"][t]) and np.array_equal(r, g["reward"][t]), t
        assert np.allclose(obs, g["obs"][t], atol=1e-6), t
    assert np.allclose(o.get_state(), g["state"], atol=1e-6)


def test_scripted_grasp_fsm_runs_the_reference_schedule():
    """peg_in_hole.py:206-212,263: 10 timed states, 2226 engine steps; the episode ends in state 9; the scripted
    pick-and-insert succeeds for part of the envs (friction grasp of the rigid peg stand-in)."""
    n = 12
    o = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3), use_double=True)
    o.reset()
    a = np.zeros((n, 4), np.float32)
    counts = np.zeros(10, int)
    seen_hole = False
    for t in range(2300):
        obs, r, d = o.step(a)
        w = o.get_state()
        if d.all():
            break
        counts[ol.state_int(w, ol.S_FSM_STATE)[0]] += 1
        seen_hole = seen_hole or bool(((ol.state_int(w, ol.S_FLAGS) >> 8) & 2).any())
    assert t + 1 == 2226 and list(counts) == [59, 481, 481, 241, 361, 361, 121, 60, 60, 0]
    assert np.isfinite(w).all() and seen_hole
    assert r.mean() >= 0.25                           # some pegs end within 5 cm of the hole (peg_in_hole.py:114-117)
    assert np.allclose(obs[:, 2:5] - np.array([[2.0 * (i // 4), 3.0 * (i % 4), 0.0] for i in range(n)]), [0.2, -0.6, 0.4], atol=0.02)


def test_camera_render_sees_table_and_peg():
    """PegInHole.render (peg_in_hole.py:276-304): camera at the end effector looking down; depth is the OpenGL
    depth-buffer encoding for near 0.001 / far 1000; the peg below the gripper shows up as a bright bar."""
    o = ol.Oracle(ol.make_config(2, task=2, offset=(2.0, 0.0, 0.0), seed=3))
    o.reset()
    a = np.zeros((2, 4), np.float32)
    for _ in range(560):
        o.step(a)                                            # FSM state 2: gripper hovering over the peg
    img = o.render(60, 60)
    assert img.shape == (2, 60, 60, 4)
    depth, rgb = img[..., 0], img[..., 1:]
    assert 0.99 < depth.min() <= depth.max() <= 1.0
    ee_z = o.obs[0, 4]
    z_table = 0.001 * 1000.0 / (1000.0 - depth[0, 0, 0] * (1000.0 - 0.001))     # invert the depth encoding
    assert abs(z_table - (ee_z + 0.05)) < 2e-3               # corner pixel sees the table at z = -0.05
    assert (rgb[0, 30, 30] > rgb[0, 0, 0]).all()             # peg (white) in the image centre, table (brown) at the corner
    assert np.array_equal(rgb, np.round(rgb)) and rgb.min() >= 0 and rgb.max() <= 255


def test_ur5_holds_its_pose_and_fly_object_is_ballistic():
    """ur5: POSITION_CONTROL kp 0.03 / 300 N m (utils.py:82) holds the arm against gravity when the target is the
    current pose; random-fly: before any contact the thrown object follows the semi-implicit-Euler parabola."""
    from scipy.spatial.transform import Rotation as Rot
    o = ol.Oracle(ol.make_config(1, arm=1, seed=1), use_double=True)
    o.reset()
    q0 = ol.state_field(o.get_state(), ol.S_Q, 9)[0, :6].astype(np.float64)
    p, R = ol.fk(1, q0, True)
    eul = Rot.from_matrix(R).as_euler("xyz")                 # getQuaternionFromEuler: fixed-axis XYZ
    a = np.array([[p[0], p[1], p[2], eul[0], eul[1], eul[2]]], np.float32)
    for _ in range(240):
        o.step(a)
    assert np.abs(ol.state_field(o.get_state(), ol.S_Q, 9)[0, :6] - q0).max() < 1e-4
    o = ol.Oracle(ol.make_config(4, task=1, dt=1.0 / 120.0, seed=2, offset=(2.0, 0.0, 0.0)), use_double=True)
    o.reset()
    w0 = o.get_state()
    p0, v0 = ol.state_field(w0, ol.S_PPOS, 3).astype(np.float64), ol.state_field(w0, ol.S_PVEL, 3).astype(np.float64)
    act = np.zeros((4, 4), np.float32)
    act[:, :3] = o.obs[:, 2:5]
    k, dt = 20, 1.0 / 120.0
    for _ in range(k):
        o.step(act)
    want = p0 + v0 * dt * k
    want[:, 2] -= 9.8 * dt * dt * k * (k + 1) / 2
    assert np.abs(ol.state_field(o.get_state(), ol.S_PPOS, 3) - want).max() < 1e-6


def test_grasp_label_images_follow_random_grasp():
    """peg_in_hole.py:41-45,65-98: empty labels (cos image = 1) until the state machine enters state 2; then a rotated
    rectangle of 0.1 x 0.2 normalised half extents = 30 x 60 pixels on 300 x 300, angle = atan2 of the rotated
    random_vector (the peg still lies along its spawn axis: +-90 degrees), pos 50, sin/cos of twice the angle, width."""
    n = 6
    o = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    o.reset()
    lab, par = o.grasp_labels(300, 300)
    assert not lab[:, [0, 1, 3]].any() and (lab[:, 2] == 1.0).all() and not par.any()
    o.rollout(540)                                               # states 0, 1 last 60 + 480 engine steps
    assert (ol.state_int(o.get_state(), ol.S_FSM_STATE) == 1).all()
    assert (ol.state_field(o.get_state(), ol.S_GRASP_ANG)[:, 0] == -100.0).all()
    o.rollout(2)
    assert (ol.state_int(o.get_state(), ol.S_FSM_STATE) == 2).all()
    lab, par = o.grasp_labels(300, 300)
    ang = ol.state_field(o.get_state(), ol.S_GRASP_ANG)[:, 0]
    ry = ol.state_field(o.get_state(), ol.S_RANDY)[:, 0]
    assert np.allclose(np.abs(ang), np.pi / 2, atol=2e-2) and (np.sign(ang) == np.sign(ry)).all()
    assert np.allclose(par[:, 2], ang / np.pi * 180, atol=1e-3) and np.allclose(par[:, 3], 60, atol=1e-3) and np.allclose(par[:, 4], 30, atol=1e-3)
    inside = lab[:, 0] == 50.0
    assert ((lab[:, 0] == 0.0) | inside).all()
    area = inside.sum(axis=(1, 2))
    assert ((area >= 1700) & (area <= 1950)).all(), area          # 60 x 30 rectangle, closed boundary
    assert inside[:, 150, 150].all() and not inside[:, 10, 10].any()
    for i in range(n):
        assert np.allclose(lab[i, 1][inside[i]], np.sin(2 * ang[i]), atol=1e-5) and not lab[i, 1][~inside[i]].any()
        assert np.allclose(lab[i, 2][inside[i]], np.cos(2 * ang[i]), atol=1e-5) and (lab[i, 2][~inside[i]] == 1.0).all()
        assert np.allclose(lab[i, 3][inside[i]], 60, atol=1e-3) and not lab[i, 3][~inside[i]].any()
    # a non-square image and a rotated rectangle: symmetric about the centre, rows <-> second vertex coordinate
    w = o.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    flat[:, ol.S_GRASP_ANG] = np.linspace(-3.0, 3.0, n)
    o.set_state(flat.reshape(n, -1, 4).transpose(1, 0, 2))
    lab, par = o.grasp_labels(80, 64)
    ins = lab[:, 0] == 50.0
    assert ins.any(axis=(1, 2)).all()
    assert (ins[:, 1:, 1:] == ins[:, 1:, 1:][:, ::-1, ::-1]).mean() > 0.98   # point symmetry about (W/2, H/2)


@pytest.mark.parametrize("dbl", [False, True])
def test_frictionless_impact_keeps_tangential_velocity_and_pushes_out_with_erp(dbl):
    """banana.urdf:5-6 has lateral_friction 0: hitting the table must leave the tangential velocity untouched, and the
    normal velocity after the solve is the push-out bias ERP * penetration / dt of a penetrating contact (Bullet's
    contact ERP 0.2 [3P]); a flat box on four symmetric corner contacts picks up no spin.  Closed forms, exact to rounding."""
    dt = 1.0 / 120.0
    o = ol.Oracle(ol.make_config(1, task=1, fly_object=0, dt=dt, seed=1), use_double=dbl)
    o.reset()
    hz, table = 0.0187, -0.05
    _set(o, 1, pos=(ol.S_PPOS, [[0.3, -0.3, table + hz + 0.004]]), vel=(ol.S_PVEL, [[1.5, -0.7, -1.0]]),
         ang=(ol.S_PANG, [[0.0, 0.0, 0.0]]), quat=(ol.S_PQUAT, [[0.0, 0.0, 0.0, 1.0]]))
    a = np.zeros((1, 4), np.float32); a[:, :3] = o.obs[:, 2:5]; a[:, 3] = 0.04
    o.step(a)                                   # 4 mm above the table: outside the 2 mm margin, pure free fall
    w = o.get_state()
    v1 = ol.state_field(w, ol.S_PVEL, 3)[0]
    z1 = ol.state_field(w, ol.S_PPOS, 3)[0, 2]
    assert np.allclose(v1, [1.5, -0.7, -1.0 - 9.8 * dt], atol=1e-6) and not (ol.state_int(w, ol.S_FLAGS)[0] >> 8)
    pen = table - (z1 - hz)
    assert pen > 0.004                          # it tunnelled: the next step sees a penetrating contact
    _, r, d = o.step(a)
    w = o.get_state()
    v2, om = ol.state_field(w, ol.S_PVEL, 3)[0], ol.state_field(w, ol.S_PANG, 3)[0]
    assert (ol.state_int(w, ol.S_FLAGS)[0] >> 8) & 1 and d[0] == 1          # table impact ends the random-fly episode
    assert abs(v2[0] - 1.5) < 1e-6 and abs(v2[1] + 0.7) < 1e-6              # mu = 0: no tangential impulse at all
    assert abs(v2[2] - 0.2 * pen / dt) < 2e-4                                # ERP push-out velocity
    assert np.abs(om).max() < 1e-3
