"""Pins the CPU oracle (oracle/) before it is trusted as the parity checker (CPU only, no GPU).
PyBullet parity is UNPINNED (not installable here; the reference has no golden vectors): what is pinned is
(a) Random123 Philox4x32-10 known answers, (b) the UR5 forward-kinematics known answers derived from the in-tree
ur5.urdf, (c) agreement of FK / ABA / M^-1 with an independent textbook numpy model,
(d) IK fixed-point / convergence properties, (e) closed-form physics."""
import numpy as np
import pytest

import oracle_lib as ol
from model_tables import Arm

PANDA, UR5 = 0, 1


def test_philox_known_answers():
    import ctypes as C
    L = ol.lib()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        L.pih_oracle_philox(ol._p(c), ol._p(k), ol._p(o))
        assert tuple(int(v) for v in o) == want


def test_uniform_range_and_determinism():
    L = ol.lib()
    u = np.zeros(4, np.float32)
    v = np.zeros(4, np.float32)
    allu = []
    for e in range(200):
        L.pih_oracle_rng4(7, e, 3, 0, ol._p(u))
        L.pih_oracle_rng4(7, e, 3, 0, ol._p(v))
        assert (u == v).all() and (u >= 0).all() and (u < 1).all()
        allu.append(u.copy())
    allu = np.concatenate(allu)
    assert abs(allu.mean() - 0.5) < 0.05


def test_det_sincos_atan_accuracy():
    import ctypes as C
    L = ol.lib()
    s, c = C.c_float(), C.c_float()
    xs = np.concatenate([np.linspace(-12, 12, 4001), [0.0, 1e-8, -1e-8, np.pi / 2, np.pi, -np.pi]])
    err = 0.0
    for x in xs:
        L.pih_oracle_sincos(C.c_float(x), C.byref(s), C.byref(c))
        xf = float(np.float32(x))
        err = max(err, abs(s.value - np.sin(xf)), abs(c.value - np.cos(xf)))
    assert err < 3e-7
    err = 0.0
    for y, x in np.random.default_rng(0).uniform(0, 1, (2000, 2)):
        err = max(err, abs(L.pih_oracle_atan2_q1(C.c_float(y), C.c_float(x)) - np.arctan2(np.float32(y), np.float32(x))))
    assert err < 4e-7


# SURVEY.md Appendix C.1 [DERIVED from peg_in_hole_gym/envs/assets/urdf/ur5.urdf]
UR5_KAT = [((0, 0, 0, 0, 0, 0), (-0.814442, -0.190153, 0.070686),
            ((-0.001593, -0.999572, 0.029199), (-0.999999, 0.001592, -0.000047), (0, -0.0292, -0.999574))),
           ((.1, .2, .3, .4, .5, .6), (-0.679057, -0.249298, -0.193684),
            ((-0.199313, -0.020081, 0.97973), (-0.901526, 0.395628, -0.175295), (-0.384089, -0.918191, -0.096957))),
           ((0, -1.57, 1.57, -1.57, -1.57, 0), (-0.497143, -0.108424, 0.517412),
            ((0.029994, -0.001617, -0.999549), (-0.000844, -0.999998, 0.001592), (-0.99955, 0.000796, -0.029996)))]


@pytest.mark.parametrize("dbl", [True, False])
def test_ur5_fk_known_answers(dbl):
    for q, pos, rot in UR5_KAT:
        p, R = ol.fk(UR5, q, use_double=dbl)
        assert np.allclose(p, pos, atol=3e-6)
        assert np.allclose(R, np.array(rot), atol=3e-6)


@pytest.mark.parametrize("arm,name", [(PANDA, "PANDA"), (UR5, "UR5")])
def test_fk_matches_independent_numpy(arm, name):
    A = Arm(name)
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.uniform(-1.5, 1.5, A.nb)
        if arm == PANDA:
            q[7:] = rng.uniform(0, 0.04, 2)
        p, R = ol.fk(arm, q, True)
        p2, R2 = A.ee(q)
        assert np.allclose(p, p2, atol=1e-12) and np.allclose(R, R2, atol=1e-12)
        pf, Rf = ol.fk(arm, q, False)
        assert np.allclose(pf, p2, atol=2e-6) and np.allclose(Rf, R2, atol=2e-6)


@pytest.mark.parametrize("arm,name", [(PANDA, "PANDA"), (UR5, "UR5")])
def test_ik_fixed_point_and_convergence(arm, name):
    A = Arm(name)
    rng = np.random.default_rng(2)
    from scipy.spatial.transform import Rotation as Rot
    for _ in range(10):
        q = rng.uniform(-1.0, 1.0, A.nb)
        if arm == PANDA:
            q[3] = -1.5 + 0.5 * rng.uniform(-1, 1)
            q[7:] = 0.0
        p, R = A.ee(q)
        quat = Rot.from_matrix(R).as_quat()
        # target == current pose: IK must not move (residual < threshold on entry)
        out = ol.ik(arm, q, p, quat, 20, True)
        assert np.allclose(out, q[:A.narm], atol=1e-9)
        # perturbed target: every extra iteration reduces the pose error; 20 iterations get within a few mm
        q2 = q.copy()
        q2[:A.narm] += rng.uniform(-0.05, 0.05, A.narm)
        p2, R2 = A.ee(q2)
        quat2 = Rot.from_matrix(R2).as_quat()
        errs = []
        for it in (1, 3, 8, 20, 60):
            o = ol.ik(arm, q, p2, quat2, it, True)
            qq = q.copy()
            qq[:A.narm] = o
            pe, Re = A.ee(qq)
            errs.append(np.linalg.norm(pe - p2) + np.linalg.norm(Rot.from_matrix(R2 @ Re.T).as_rotvec()))
        assert all(errs[i + 1] <= errs[i] + 1e-12 for i in range(len(errs) - 1)), errs
        assert errs[-1] < 0.25 * errs[0]
        # float build agrees with the double build
        of = ol.ik(arm, q, p2, quat2, 20, False)
        od = ol.ik(arm, q, p2, quat2, 20, True)
        assert np.allclose(of, od, atol=2e-4)


@pytest.mark.parametrize("arm,name", [(PANDA, "PANDA"), (UR5, "UR5")])
def test_ik_matches_independent_damped_least_squares(arm, name):
    """The IK restated independently (numpy + scipy rotations, geometric Jacobian from the independent FK of
    tests/model_tables.py; shares no code with the oracle): dtheta = J^T (J J^T + lambda^2 I)^-1 e with lambda = 0.6,
    e = [p* - p ; rotation vector of R* R^T], largest |dtheta| clamped to 45 degrees, a fixed number of iterations
    seeded at the current joints - the algorithm DESIGN.md section 2.3 states.  The double oracle follows it to 1e-9 rad
    after 1, 5 and 20 iterations, for small and for out-of-reach steps (where the clamp is active)."""
    from scipy.spatial.transform import Rotation as Rot
    A = Arm(name)
    rng = np.random.default_rng(12)

    def ik_np(q0, p_t, R_t, iters):
        q = q0.copy()
        for _ in range(iters):
            Rw, pw = A.fk(q)
            pe, Re = A.ee(q)
            e = np.concatenate([p_t - pe, Rot.from_matrix(R_t @ Re.T).as_rotvec()])
            if e @ e < 1e-8:
                break
            J = np.zeros((6, A.narm))
            for j in range(A.narm):
                z = Rw[j][:, 2]
                J[:3, j], J[3:, j] = np.cross(z, pe - pw[j]), z
            d = J.T @ np.linalg.solve(J @ J.T + 0.36 * np.eye(6), e)
            m = np.abs(d).max()
            if m > np.pi / 4:
                d *= (np.pi / 4) / m
            q[:A.narm] += d
        return q[:A.narm]

    for trial in range(12):
        q = rng.uniform(-1.0, 1.0, A.nb)
        if arm == PANDA:
            q[3] = -1.5 + 0.5 * rng.uniform(-1, 1)
            q[7:] = 0.0
        scale = 0.05 if trial < 6 else 0.8            # second half: far targets, the 45 degree clamp engages
        q2 = q.copy()
        q2[:A.narm] += rng.uniform(-scale, scale, A.narm)
        p_t, R_t = A.ee(q2)
        if trial >= 9:
            p_t = p_t + rng.uniform(-1.0, 1.0, 3)      # out of reach
        quat = Rot.from_matrix(R_t).as_quat()
        for iters in (1, 5, 20):
            want = ik_np(q, p_t, R_t, iters)
            got = ol.ik(arm, q, p_t, quat, iters, True)
            assert np.allclose(got, want, atol=1e-9), (trial, iters, np.abs(got - want).max())


@pytest.mark.parametrize("arm,name", [(PANDA, "PANDA"), (UR5, "UR5")])
def test_aba_and_minv_match_lagrangian_model(arm, name):
    A = Arm(name)
    rng = np.random.default_rng(3)
    for trial in range(6):
        q = rng.uniform(-1.2, 1.2, A.nb)
        qd = rng.uniform(-1.0, 1.0, A.nb) if trial else np.zeros(A.nb)
        if arm == PANDA:
            q[7:] = rng.uniform(0, 0.04, 2)
            qd[7:] *= 0.05
        tau = rng.uniform(-5, 5, A.nb)
        qdd, minv = ol.aba(arm, q, qd, tau)
        M = A.mass_matrix(q)
        assert np.allclose(minv @ M, np.eye(A.nb), atol=1e-8)
        want = np.linalg.solve(M, tau - A.bias(q, qd))
        assert np.allclose(qdd, want, rtol=1e-5, atol=2e-5 * max(1.0, np.abs(want).max())), (qdd, want)


def test_gravity_compensation_gives_zero_acceleration():
    A = Arm("PANDA")
    q = np.array([0, -0.215, -np.pi / 3, -2.57, 0, 2.356, 2.356, 0.01, 0.02])
    g = A.bias(q, np.zeros(9))
    qdd, _ = ol.aba(PANDA, q, np.zeros(9), g)
    assert np.abs(qdd).max() < 1e-6
