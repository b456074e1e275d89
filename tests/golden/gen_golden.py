#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the CPU oracle (float build).  The reference itself cannot run here (PyBullet
is not installable), so these are ORACLE-generated regression fixtures, not reference outputs: they pin the oracle
(and through it the CUDA path) against accidental change.  Usage: python tests/golden/gen_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = {
    "panda_peg": dict(task=0, arm=0, offset=(2.0, 3.0, 0.0), seed=42, max_steps=80, auto_reset=1),
    "ur5_peg": dict(task=0, arm=1, offset=(2.0, 3.0, 0.0), seed=43, max_steps=80, auto_reset=1),
    "panda_fly": dict(task=1, arm=0, fly_object=0, dt=1.0 / 120.0, offset=(2.0, 3.0, 0.0), seed=44, max_steps=100, auto_reset=1),
    "panda_pipe": dict(task=0, arm=0, peg_model=1, offset=(2.0, 3.0, 0.0), seed=45, max_steps=90, auto_reset=1),   # the 25-link pipe.urdf chain
}
N, STEPS = 8, 200


def actions(case, n, steps, offs):
    rng = np.random.default_rng(abs(hash(case)) % 1000 if False else sum(map(ord, case)))
    adim = 6 if "ur5" in case else 4
    out = rng.uniform(-1, 1, (steps, n, adim)).astype(np.float32)
    out[:, :, :3] += offs
    return out


def run(case):
    import ctypes as C
    cfg = ol.make_config(N, **CASES[case])
    o = ol.Oracle(cfg)
    offs = np.zeros((N, 3), np.float32)
    for i in range(N):
        buf = (C.c_float * 3)()
        ol.lib().pih_oracle_env_offset(C.byref(cfg), i, buf)
        offs[i] = list(buf)
    acts = actions(case, N, STEPS, offs)
    obs0 = o.reset()
    obs, rew, done = [], [], []
    for t in range(STEPS):
        a, b, c = o.step(acts[t])
        obs.append(a); rew.append(b); done.append(c)
    return dict(actions=acts, obs0=obs0, obs=np.stack(obs), reward=np.stack(rew), done=np.stack(done), state=o.get_state())


if __name__ == "__main__":
    for case in CASES:
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **run(case))
        print("wrote", case)
