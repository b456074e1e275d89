"""Independent numpy reading of the generated model tables (csrc/pih_model_data.h) + a textbook rigid-body
dynamics model (mass matrix from body Jacobians, Lagrangian bias forces by finite differences).  Used to pin the
oracle's FK / ABA / M^-1 against arithmetic that shares no code with it."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "peg_in_hole_gym_b200", "csrc",
                   "pih_model_data.h")


def load(prefix):
    txt = open(HDR).read()
    out = {}
    for m in re.finditer(r"#define PIH_%s_(\w+) (-?\d+)" % prefix, txt):
        out[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"(?:static const|PIH_TABLE) (int|double) PIH_%s_(\w+)((?:\[\d+\])+) = (.*?);" % prefix, txt, re.S):
        dims = [int(d) for d in re.findall(r"\[(\d+)\]", m.group(3))]
        vals = [float(v) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", m.group(4))]
        a = np.array(vals).reshape(dims)
        out[m.group(2)] = a.astype(int) if m.group(1) == "int" else a
    return out


class Arm:
    def __init__(self, prefix):
        t = load(prefix)
        self.t = t
        self.nb, self.narm, self.ee_body = t["NB"], t["NARM"], t["EE_BODY"]
        self.parent, self.jtype = t["PARENT"], t["JTYPE"]
        self.R0 = t["R0"].reshape(-1, 3, 3)
        self.t0 = t["T0"]
        self.mass, self.h = t["MASS"], t["H"]
        self.Irot = t["IROT"].reshape(-1, 3, 3)
        self.com = self.h / self.mass[:, None]
        self.Ic = np.array([self.Irot[i] - self.mass[i] * (self.com[i] @ self.com[i] * np.eye(3) - np.outer(self.com[i], self.com[i]))
                            for i in range(self.nb)])
        self.ee_R, self.ee_t = t["EE_R"].reshape(3, 3), t["EE_T"]

    def fk(self, q):
        Rw, pw = [None] * self.nb, [None] * self.nb
        for i in range(self.nb):
            c, s = np.cos(q[i]), np.sin(q[i])
            if self.jtype[i] == 0:
                Rl, tl = self.R0[i] @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]), self.t0[i]
            else:
                Rl, tl = self.R0[i], self.t0[i] + self.R0[i][:, 2] * q[i]
            p = self.parent[i]
            if p < 0:
                Rw[i], pw[i] = Rl, tl
            else:
                Rw[i], pw[i] = Rw[p] @ Rl, pw[p] + Rw[p] @ tl
        return Rw, pw

    def ee(self, q):
        Rw, pw = self.fk(q)
        e = self.ee_body
        return pw[e] + Rw[e] @ self.ee_t, Rw[e] @ self.ee_R

    def mass_matrix(self, q):
        Rw, pw = self.fk(q)
        M = np.zeros((self.nb, self.nb))
        for i in range(self.nb):
            cw = pw[i] + Rw[i] @ self.com[i]
            Jv, Jw = np.zeros((3, self.nb)), np.zeros((3, self.nb))
            j = i
            while j >= 0:
                z = Rw[j][:, 2]
                if self.jtype[j] == 0:
                    Jv[:, j], Jw[:, j] = np.cross(z, cw - pw[j]), z
                else:
                    Jv[:, j] = z
                j = self.parent[j]
            Iw = Rw[i] @ self.Ic[i] @ Rw[i].T
            M += self.mass[i] * Jv.T @ Jv + Jw.T @ Iw @ Jw
        return M

    def potential(self, q, g=9.8):
        Rw, pw = self.fk(q)
        return sum(self.mass[i] * g * (pw[i] + Rw[i] @ self.com[i])[2] for i in range(self.nb))

    def bias(self, q, qd, g=9.8, eps=1e-6):
        """C(q,qd) qd + g(q) from the Lagrangian, by central differences of M(q) and V(q)."""
        n = self.nb
        dM = np.zeros((n, n, n))
        dV = np.zeros(n)
        for k in range(n):
            e = np.zeros(n)
            e[k] = eps
            dM[k] = (self.mass_matrix(q + e) - self.mass_matrix(q - e)) / (2 * eps)
            dV[k] = (self.potential(q + e, g) - self.potential(q - e, g)) / (2 * eps)
        Mdot = np.tensordot(qd, dM, axes=(0, 0))
        return Mdot @ qd - 0.5 * np.array([qd @ dM[k] @ qd for k in range(n)]) + dV
