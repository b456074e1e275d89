"""Pins the SOLVER half of the oracle independently: for env-steps with motors, joint limits,
contacts with friction and the grasp constraint, the assembled velocity-level problem (rows, bounds, targets, M^-1, v*)
is taken from the oracle through a test hook and checked in numpy, sharing no code with it:

  1. the oracle's impulses satisfy the mixed-LCP conditions at convergence:  lo <= lambda <= hi (friction: |lambda_t| <=
     mu lambda_n),  w = J v+ - tv  is 0 for a free row, >= 0 at the lower bound, <= 0 at the upper bound, and
     v+ = v* + M^-1 J^T lambda;
  2. a DENSE solve of the same problem - the box QP  min 1/2 l^T A l + l^T (J v* - tv),  A = J M^-1 J^T,  with the
     friction bounds frozen at the oracle's normal impulses, solved by scipy's bounded-variable least squares on the
     dense matrix - gives the same post-step velocities (v+ is unique although lambda need not be)."""
import ctypes as C

import numpy as np

import oracle_lib as ol

MAXR, NV = 3 * 9 + 3 * 8, 9 + 6


class Capture(C.Structure):
    _fields_ = [("nrows", C.c_int), ("nv", C.c_int), ("J", C.c_double * NV * MAXR), ("tv", C.c_double * MAXR), ("lo", C.c_double * MAXR),
                ("hi", C.c_double * MAXR), ("kind", C.c_int * MAXR), ("lam", C.c_double * MAXR), ("Minv", C.c_double * NV * NV),
                ("vstar", C.c_double * NV), ("vplus", C.c_double * NV), ("iters", C.c_int)]


def capture(o, i, action):
    L = ol.lib()
    assert L.pih_oracle_capture_size() == C.sizeof(Capture)
    cap = Capture()
    a = np.ascontiguousarray(action, np.float32)
    L.pih_oracle_capture_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert L.pih_oracle_capture_step(C.c_void_p(o.h), i, a.ctypes.data_as(C.c_void_p), C.byref(cap)) == 0
    R, nv = cap.nrows, cap.nv
    g = lambda f, shape: np.ctypeslib.as_array(getattr(cap, f)).reshape(shape).copy()
    return dict(R=R, nv=nv, J=g("J", (MAXR, NV))[:R, :nv], tv=g("tv", MAXR)[:R], lo=g("lo", MAXR)[:R], hi=g("hi", MAXR)[:R],
                kind=g("kind", MAXR)[:R], lam=g("lam", MAXR)[:R], Minv=g("Minv", (NV, NV))[:nv, :nv], vstar=g("vstar", NV)[:nv],
                vplus=g("vplus", NV)[:nv], iters=cap.iters)


def bounds(p, lam):
    lo, hi = p["lo"].copy(), p["hi"].copy()
    for r in np.flatnonzero(p["kind"] == 1):
        lim = p["hi"][r] * lam[int(p["lo"][r])]          # friction row: hi holds mu, lo the index of its normal row
        lo[r], hi[r] = -lim, lim
    return lo, np.minimum(hi, 1e30)


def dense_velocity_solve(p, lo, hi):
    """Dense solve of the step's velocity-level problem with scipy's bounded-variable least squares (BVLS, an active-set
    method on the dense matrix).  With M^-1 = C C^T and B = C^T J^T:  1/2 l^T J M^-1 J^T l + l^T (J v* - tv)
    = 1/2 |B l + C^-1 (v* - u)|^2 + const  where u is any velocity with J u = tv (least-squares: redundant rows)."""
    from scipy.optimize import lsq_linear
    J, Minv = p["J"], p["Minv"]
    Cc = np.linalg.cholesky(0.5 * (Minv + Minv.T))
    u = np.linalg.lstsq(J, p["tv"], rcond=None)[0]
    B = Cc.T @ J.T
    rhs = -np.linalg.solve(Cc, p["vstar"] - u)
    hi_ = np.where(hi > 1e29, np.inf, hi)
    fixed = hi_ - lo < 1e-300                                 # (a friction row under a zero normal impulse)
    hi_ = np.where(fixed, lo + 1e-300, hi_)
    res = lsq_linear(B, rhs, bounds=(lo, hi_), method="bvls", tol=1e-15, max_iter=2000)
    consistent = np.abs(J @ u - p["tv"]).max() < 1e-9        # tv in the range of J: the reformulation is exact
    return p["vstar"] + Minv @ J.T @ res.x, consistent


def check_problem(p, tol_v):
    """Returns (rows, friction rows, checked).  A set of velocity constraints can be redundant or even contradictory (two
    contact points of one rigid segment whose rows are linearly dependent but ask for different push-out velocities):
    Gauss-Seidel then settles on velocities while the impulses drift in the null space, and no LCP solution exists to
    compare with.  The conditions are therefore asserted when the rows strictly inside their bounds are linearly
    independent (the well-posed case); the bound checks and v+ = v* + M^-1 J^T lambda are asserted always."""
    J, Minv, lam = p["J"], p["Minv"], p["lam"]
    vplus = p["vstar"] + Minv @ J.T @ lam
    assert np.abs(vplus - p["vplus"]).max() < 1e-5 * max(1.0, np.abs(vplus).max())            # v+ = v* + M^-1 J^T lambda
    lo, hi = bounds(p, lam)
    scale = np.maximum(1.0, np.abs(lam))
    assert (lam >= lo - 1e-6 * scale).all() and (lam <= hi + 1e-6 * scale).all()
    w = J @ vplus - p["tv"]
    at_lo, at_hi = lam <= lo + 1e-7 * scale, lam >= hi - 1e-7 * scale
    free = ~at_lo & ~at_hi
    well_posed = True
    if free.any():
        sv = np.linalg.svd(J[free] @ np.linalg.cholesky(0.5 * (Minv + Minv.T)), compute_uv=False)
        well_posed = sv.min() > 1e-6 * sv.max()
    if not well_posed:
        return len(lam), int((p["kind"] == 1).sum()), False
    assert np.abs(w[free]).max(initial=0.0) < tol_v                                            # free rows reach their target velocity
    assert (w[at_lo & ~at_hi] > -tol_v).all() and (w[at_hi & ~at_lo] < tol_v).all()           # complementarity at the bounds
    v_dense, consistent = dense_velocity_solve(p, lo, hi)
    if consistent:
        assert np.abs(v_dense - vplus).max() < 50 * tol_v, np.abs(v_dense - vplus).max()      # dense re-solve: same velocities
    return len(lam), int((p["kind"] == 1).sum()), True


def _pair_of_oracles(n, task, seed, iters=20000, **kw):
    """Two double-precision oracles that differ by ONE Gauss-Seidel sweep (no early exit): where their impulses agree the
    iteration has reached its fixed point.  (It need not: with the peg's friction coefficient of 100 x 0.5 a sliding
    contact can make the sweeps cycle between 'contact on' and 'contact off' - a property of projected Gauss-Seidel on
    a pyramid-friction problem, which Bullet shares; such env-steps are skipped, and counted.)"""
    mk = lambda it: ol.Oracle(ol.make_config(n, task=task, offset=(2.0, 3.0, 0.0), seed=seed, solver_iters=it, solver_residual=0.0, **kw), use_double=True)
    return mk(iters), mk(iters + 1)


def _checked_capture(oa, ob, i, action, tol_v):
    pa, pb = capture(oa, i, action), capture(ob, i, action)
    scale = np.maximum(1.0, np.abs(pa["lam"]))
    if pa["R"] != pb["R"] or (np.abs(pa["lam"] - pb["lam"]) > 1e-9 * scale).any():
        return pa, False                                        # no fixed point reached (cycling sweep)
    return pa, check_problem(pa, tol_v)[2]


def _fuzz_state(o, rng, lying_frac=0.5, fly=False):
    n = o.n
    w = o.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    flat[:, ol.S_Q:ol.S_Q + 7] += rng.normal(0, 0.5, (n, 7)).astype(np.float32)     # some joints land on their limits
    flat[:, ol.S_Q + 7:ol.S_Q + 9] = rng.uniform(0.0, 0.04, (n, 2))
    flat[:, ol.S_QD:ol.S_QD + 9] = rng.normal(0, 1.0, (n, 9))
    k = int(n * lying_frac)
    q = rng.normal(size=(n, 4))
    q[:k] *= np.array([0.05, 0.05, 1, 1])                                            # lying flat (several contact points) ...
    q /= np.linalg.norm(q, axis=1, keepdims=True)                                    # ... or tilted (one end / corner touches)
    flat[:, ol.S_PQUAT:ol.S_PQUAT + 4] = q
    flat[:, ol.S_PPOS + 2] = rng.uniform(-0.045, -0.03, n) if not fly else rng.uniform(-0.04, -0.02, n)
    flat[:, ol.S_PPOS:ol.S_PPOS + 2] = rng.uniform(-0.3, 0.3, (n, 2)) + np.array([0.0, -0.5])
    flat[:, ol.S_PVEL:ol.S_PVEL + 3] = rng.normal(0, 0.3, (n, 3))
    flat[:, ol.S_PANG:ol.S_PANG + 3] = rng.normal(0, 1.0, (n, 3))
    return np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32)


def _offsets(o):
    offs = np.zeros((o.n, 3), np.float32)
    for i in range(o.n):
        b = (C.c_float * 3)(); ol.lib().pih_oracle_env_offset(C.byref(o.cfg), i, b); offs[i] = list(b)
    return offs


def test_pgs_fixed_point_is_the_lcp_solution_motors_limits_and_friction_contacts():
    n = 48
    oa, ob = _pair_of_oracles(n, 0, 5)
    oa.reset(); ob.reset()
    rng = np.random.default_rng(2)
    st = _fuzz_state(oa, rng)
    oa.set_state(st); ob.set_state(st)
    offs = _offsets(oa)
    checked = fric = lim = 0
    for i in range(n):
        a = rng.uniform(-1, 1, 4).astype(np.float32); a[:3] += offs[i]
        p, ok = _checked_capture(oa, ob, i, a, 2e-6)
        if ok:
            checked += 1
            fric += int(((p["kind"] == 1) & (np.abs(p["lam"]) > 0)).sum())
            lim += int(((p["hi"] > 1e29) & (np.abs(p["J"][:, :9]).sum(1) == 1) & (np.abs(p["J"][:, 9:]).sum(1) == 0)).sum())
    assert checked >= 6 and fric >= 4 and lim >= 10, (checked, fric, lim)


def test_pgs_fixed_point_is_the_lcp_solution_frictionless_box_with_many_contacts():
    """random-fly object (mu = 0, banana.urdf:6): up to eight corner contacts of one box on the table plus the arm rows."""
    n = 32
    oa, ob = _pair_of_oracles(n, 1, 7, dt=1.0 / 120.0)
    oa.reset(); ob.reset()
    rng = np.random.default_rng(4)
    st = _fuzz_state(oa, rng, lying_frac=0.7, fly=True)
    oa.set_state(st); ob.set_state(st)
    offs = _offsets(oa)
    checked = contacts = 0
    for i in range(n):
        a = rng.uniform(-1, 1, 4).astype(np.float32); a[:3] += offs[i]
        p, ok = _checked_capture(oa, ob, i, a, 2e-6)
        if ok:
            checked += 1
            contacts += int(((p["hi"] > 1e29) & (np.abs(p["J"][:, 9:]).sum(1) > 0)).sum())
    assert checked >= 6 and contacts >= 8, (checked, contacts)


def test_pgs_fixed_point_is_the_lcp_solution_grasp_constraint_rows():
    """Scripted episode, lift phase (state 4): the grasp constraint's three bilateral rows couple the arm's joints and the
    peg's six degrees of freedom.  The fingers are opened in the injected state so that the (redundant, mu = 100)
    finger<->peg contacts do not make the sweep cycle: what remains - motors, finger limits, the coupled rows - is a
    well-posed problem whose fixed point must be the LCP solution."""
    n = 16
    oa, ob = _pair_of_oracles(n, 2, 3, iters=60000)
    quick = ol.Oracle(ol.make_config(n, task=2, offset=(2.0, 3.0, 0.0), seed=3))
    quick.reset(); quick.rollout(1400)                                              # state 4 (lift), default solver settings
    w = quick.get_state()
    flat = np.ascontiguousarray(w.transpose(1, 0, 2)).reshape(n, -1)
    assert (flat[:, ol.S_FSM_STATE].view(np.int32) == 4).all()
    flat[:, ol.S_Q + 7:ol.S_Q + 9] = 0.04
    flat[:, ol.S_PPOS + 2] = np.maximum(flat[:, ol.S_PPOS + 2], 0.05)               # clear of the table
    st = np.ascontiguousarray(flat.reshape(n, -1, 4).transpose(1, 0, 2)).astype(np.float32)
    oa.reset(); ob.reset()
    oa.set_state(st); ob.set_state(st)
    coupled = checked = 0
    for i in range(n):
        p, ok = _checked_capture(oa, ob, i, np.zeros(4, np.float32), 5e-5)
        checked += int(ok)
        coupled += int(((np.abs(p["J"][:, :9]).sum(1) > 0) & (np.abs(p["J"][:, 9:]).sum(1) > 0)).sum())
    assert coupled >= 3 * n and checked >= n // 2, (coupled, checked)
