"""OPTIONAL parity against the real reference engine.  PyBullet is not installable in this image (no wheel, no
network), so this test auto-skips; it is kept ready for an environment that provides `pybullet` + `pybullet_data`
(SURVEY.md section 8c): it drives panda_execute (envs/utils.py:60-68) + stepSimulation in p.DIRECT with the same
actions as the CPU oracle and reports the per-step deviation of joint positions and end-effector pose against the
north-star tolerance (1e-3 m / 1e-2 rad).  Because the Panda constants and contact geometry are restated from
memory, a failure here is information for DESIGN.md section 8, not a regression of the CUDA path."""
import math

import numpy as np
import pytest

pybullet = pytest.importorskip("pybullet")
pybullet_data = pytest.importorskip("pybullet_data")

import oracle_lib as ol  # noqa: E402


def test_panda_execute_trajectory_against_pybullet():
    p = pybullet
    cid = p.connect(p.DIRECT)
    try:
        p.setAdditionalSearchPath(pybullet_data.getDataPath())
        p.setGravity(0, 0, -9.8)
        panda = p.loadURDF("franka_panda/panda.urdf", basePosition=[0, 0, 0],
                           baseOrientation=p.getQuaternionFromEuler([0, 0, -math.pi / 2]), useFixedBase=True)
        rest = [0, -0.215, -math.pi / 3, -2.57, 0, 2.356, 2.356, 0.08, 0.08]
        for i, v in enumerate(rest):
            p.resetJointState(panda, i, v)
        o = ol.Oracle(ol.make_config(1, seed=0), use_double=True)
        o.reset()
        rng = np.random.default_rng(0)
        worst_q = worst_p = 0.0
        for t in range(200):
            a = rng.uniform(-1, 1, 4).astype(np.float32)
            # reference: panda_execute (envs/utils.py:60-68)
            orn = p.getQuaternionFromEuler([0., -math.pi, 0.])
            cur = p.getLinkState(panda, 11)[0]
            new = [cur[i] + max(-2 / 240., min(2 / 240., a[i] - cur[i])) for i in range(3)]
            jp = p.calculateInverseKinematics(panda, 11, new, orn)[0:7]
            p.setJointMotorControlArray(panda, list(range(7)) + [9, 10], p.POSITION_CONTROL, list(jp) + 2 * [float(a[3])],
                                        positionGains=[1] * 9)
            p.stepSimulation()
            obs, _, _ = o.step(a[None])
            q_ref = np.array([p.getJointState(panda, j)[0] for j in list(range(7)) + [9, 10]])
            q_ours = ol.state_field(o.get_state(), ol.S_Q, 9)[0]
            ee_ref = np.array(p.getLinkState(panda, 11)[0])
            worst_q = max(worst_q, np.abs(q_ref[:7] - q_ours[:7]).max())
            worst_p = max(worst_p, np.abs(ee_ref - obs[0, 2:5]).max())
        print(f"vs PyBullet over 200 steps: max |dq| = {worst_q:.3e} rad, max |dEE| = {worst_p:.3e} m")
        assert worst_q < 1e-2 and worst_p < 1e-3
    finally:
        p.disconnect(cid)
