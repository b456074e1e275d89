#!/usr/bin/env python3
"""Build-time model-table generator (runs in the dev container only; its OUTPUT is committed).

Reads the in-tree reference URDFs for the UR5 (peg_in_hole_gym/envs/assets/urdf/ur5.urdf:11-218,534-539
and the collision STLs next to it) and a hand-restated Franka Panda table (pybullet_data is not
on disk; see SURVEY.md Appendix C.4 - [3P], UNVERIFIED), and writes
``peg_in_hole_gym_b200/csrc/pih_model_data.h``: plain C arrays of doubles consumed by BOTH the CPU
oracle (oracle/) and the CUDA kernels (csrc/).  The runtime never parses XML.

Conventions of the emitted tables:
  * every moving joint becomes one *body*; fixed-joint children are merged into their parent body;
  * each body frame has its origin at the joint origin and its **z axis along the joint axis**
    (URDF link frames with x/y axes are re-expressed offline), so the joint motion subspace is
    always e_z (revolute: angular z, prismatic: linear z);
  * body 0's placement already contains the arm base pose in the env frame
    (panda: yaw -pi/2, envs/utils.py:32-34; ur5: identity + world_joint z=0.1, ur5.urdf:534-539);
  * link inertia follows Bullet's default URDF loading (inertia tensors in the file are ignored and a
    box inertia of the collision AABB (+1 mm margin) is used).
"""
import math
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("PIH_REFERENCE", "/root/reference")
ASSETS = os.path.join(REF, "peg_in_hole_gym", "envs", "assets")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "peg_in_hole_gym_b200", "csrc",
                   "pih_model_data.h")
MARGIN = 0.001  # pybullet URDF default collision margin [3P]


def rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def snap(a):
    a = np.array(a, dtype=np.float64)
    a[np.abs(a) < 1e-12] = 0.0
    for v in (1.0, -1.0):
        a[np.abs(a - v) < 1e-12] = v
    return a


def axis_to_Q(axis):
    """Rotation Q with Q e_z = axis (coordinate axes only)."""
    a = snap(axis)
    table = {
        (0, 0, 1): np.eye(3),
        (0, 0, -1): np.diag([1.0, -1.0, -1.0]),
        (0, 1, 0): rpy_to_R(-math.pi / 2, 0, 0),
        (0, -1, 0): rpy_to_R(math.pi / 2, 0, 0),
        (1, 0, 0): rpy_to_R(0, math.pi / 2, 0),
        (-1, 0, 0): rpy_to_R(0, -math.pi / 2, 0),
    }
    key = tuple(int(v) for v in a)
    assert key in table, f"non-coordinate joint axis {axis}"
    return snap(table[key])


def box_inertia(mass, half):
    l = 2.0 * (np.asarray(half, dtype=np.float64) + MARGIN)
    return mass / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])


def stl_half_extents(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * n == len(data):
        v = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                          count=n, offset=84)["v"].reshape(-1, 3)
    else:  # ascii
        v = np.array([[float(t) for t in line.split()[1:4]] for line in data.decode().splitlines()
                      if line.strip().startswith("vertex")])
    return (v.max(0) - v.min(0)) * 0.5


class Link:
    def __init__(self, name, mass=0.0, com=(0, 0, 0), half=None, idiag=None):
        self.name, self.mass, self.com = name, float(mass), np.array(com, dtype=np.float64)
        self.idiag = np.array(idiag if idiag is not None else (box_inertia(mass, half) if half is not None
                                                                else (0, 0, 0)), dtype=np.float64)


class Joint:
    def __init__(self, name, jtype, parent, child, xyz, rpy, axis=(0, 0, 1), lo=0.0, hi=0.0, effort=0.0,
                 damping=0.0):
        self.name, self.jtype, self.parent, self.child = name, jtype, parent, child
        self.xyz, self.rpy = np.array(xyz, dtype=np.float64), tuple(rpy)
        self.axis, self.lo, self.hi, self.effort, self.damping = np.array(axis, dtype=np.float64), lo, hi, effort, damping


def parse_ur5():
    """ur5.urdf (reference, in tree).  Everything commented out in the file is invisible to ElementTree."""
    root = ET.parse(os.path.join(ASSETS, "urdf", "ur5.urdf")).getroot()
    links, joints = {}, []
    for l in root.findall("link"):
        name = l.get("name")
        ine = l.find("inertial")
        mass, com = 0.0, (0, 0, 0)
        if ine is not None:
            mass = float(ine.find("mass").get("value"))
            o = ine.find("origin")
            if o is not None:
                com = [float(t) for t in o.get("xyz", "0 0 0").split()]
        half = None
        col = l.find("collision")
        if col is not None:
            g = col.find("geometry")
            if g.find("mesh") is not None:
                fn = os.path.normpath(os.path.join(ASSETS, "urdf", g.find("mesh").get("filename")))
                half = stl_half_extents(fn)
            elif g.find("box") is not None:
                half = np.array([float(t) for t in g.find("box").get("size").split()]) * 0.5
        links[name] = Link(name, mass, com, half if half is not None else (0, 0, 0))
    for j in root.findall("joint"):
        o = j.find("origin")
        xyz = [float(t) for t in o.get("xyz", "0 0 0").split()]
        rpy = [float(t) for t in o.get("rpy", "0 0 0").split()]
        ax = j.find("axis")
        axis = [float(t) for t in ax.get("xyz").split()] if ax is not None else (0, 0, 1)
        lim, dyn = j.find("limit"), j.find("dynamics")
        joints.append(Joint(j.get("name"), j.get("type"), j.find("parent").get("link"), j.find("child").get("link"),
                            xyz, rpy, axis,
                            float(lim.get("lower")) if lim is not None else 0.0,
                            float(lim.get("upper")) if lim is not None else 0.0,
                            float(lim.get("effort")) if lim is not None else 0.0,
                            float(dyn.get("damping")) if dyn is not None else 0.0))
    return links, joints, "world"


def panda_tables():
    """Franka Panda as shipped in pybullet_data/franka_panda/panda.urdf.
    [3P] RESTATED FROM MEMORY - NOT ON DISK - UNVERIFIED.
    Link AABB half extents (for Bullet-style box inertias) are coarse estimates of the collision hulls."""
    L = {}
    def link(name, mass, com, half):
        L[name] = Link(name, mass, com, half)
    link("panda_link0", 0.0, (0, 0, 0), (0.1, 0.1, 0.07))
    link("panda_link1", 2.74, (0, -0.04, -0.05), (0.055, 0.09, 0.13))
    link("panda_link2", 2.74, (0, -0.04, 0.06), (0.055, 0.13, 0.09))
    link("panda_link3", 2.38, (0.01, 0.01, -0.05), (0.09, 0.065, 0.1))
    link("panda_link4", 2.38, (-0.03, 0.03, 0.02), (0.09, 0.1, 0.065))
    link("panda_link5", 2.74, (0, 0.04, -0.12), (0.055, 0.1, 0.17))
    link("panda_link6", 1.55, (0.04, 0, 0), (0.09, 0.06, 0.065))
    link("panda_link7", 0.54, (0, 0, 0.08), (0.05, 0.05, 0.06))
    link("panda_link8", 0.0, (0, 0, 0), (0, 0, 0))
    link("panda_hand", 0.73, (0, 0, 0.04), (0.032, 0.102, 0.045))
    link("panda_leftfinger", 0.1, (0, 0.01, 0.02), (0.011, 0.013, 0.027))
    link("panda_rightfinger", 0.1, (0, -0.01, 0.02), (0.011, 0.013, 0.027))
    link("panda_grasptarget", 0.0, (0, 0, 0), (0, 0, 0))
    hp = math.pi / 2
    J = [
        Joint("panda_joint1", "revolute", "panda_link0", "panda_link1", (0, 0, 0.333), (0, 0, 0), (0, 0, 1), -2.9671, 2.9671, 87),
        Joint("panda_joint2", "revolute", "panda_link1", "panda_link2", (0, 0, 0), (-hp, 0, 0), (0, 0, 1), -1.8326, 1.8326, 87),
        Joint("panda_joint3", "revolute", "panda_link2", "panda_link3", (0, -0.316, 0), (hp, 0, 0), (0, 0, 1), -2.9671, 2.9671, 87),
        Joint("panda_joint4", "revolute", "panda_link3", "panda_link4", (0.0825, 0, 0), (hp, 0, 0), (0, 0, 1), -3.1416, 0.0, 87),
        Joint("panda_joint5", "revolute", "panda_link4", "panda_link5", (-0.0825, 0.384, 0), (-hp, 0, 0), (0, 0, 1), -2.9671, 2.9671, 12),
        Joint("panda_joint6", "revolute", "panda_link5", "panda_link6", (0, 0, 0), (hp, 0, 0), (0, 0, 1), -0.0873, 3.8223, 12),
        Joint("panda_joint7", "revolute", "panda_link6", "panda_link7", (0.088, 0, 0), (hp, 0, 0), (0, 0, 1), -2.9671, 2.9671, 12),
        Joint("panda_joint8", "fixed", "panda_link7", "panda_link8", (0, 0, 0.107), (0, 0, 0)),
        Joint("panda_hand_joint", "fixed", "panda_link8", "panda_hand", (0, 0, 0), (0, 0, -0.785398163397)),
        Joint("panda_finger_joint1", "prismatic", "panda_hand", "panda_leftfinger", (0, 0, 0.0584), (0, 0, 0), (0, 1, 0), 0.0, 0.04, 20),
        Joint("panda_finger_joint2", "prismatic", "panda_hand", "panda_rightfinger", (0, 0, 0.0584), (0, 0, 0), (0, -1, 0), 0.0, 0.04, 20),
        Joint("panda_grasptarget_hand", "fixed", "panda_hand", "panda_grasptarget", (0, 0, 0.105), (0, 0, 0)),
    ]
    return L, J, "panda_link0"


def build(links, joints, root, base_R, base_t, ee_link, spheres, boxes):
    """Flatten to moving bodies with z-axis joint frames.  spheres/boxes: (link, centre, radius|half)."""
    children = {}
    for j in joints:
        children.setdefault(j.parent, []).append(j)
    bodies = []          # dicts
    link_owner = {}      # link name -> (body index, R_link_in_body, t_link_in_body)

    def visit(link_name, body, R_lb, t_lb):
        link_owner[link_name] = (body, R_lb, t_lb)
        for j in children.get(link_name, []):
            Ro = snap(rpy_to_R(*j.rpy))
            if j.jtype == "fixed":
                visit(j.child, body, snap(R_lb @ Ro), R_lb @ j.xyz + t_lb)
            else:
                Q = axis_to_Q(j.axis)
                b = dict(parent=body, jtype=0 if j.jtype in ("revolute", "continuous") else 1,
                         R0=snap(R_lb @ Ro @ Q), t0=snap(R_lb @ j.xyz + t_lb), lo=j.lo, hi=j.hi,
                         effort=j.effort, damping=j.damping, name=j.name, parts=[])
                bodies.append(b)
                visit(j.child, len(bodies) - 1, snap(Q.T), np.zeros(3))

    visit(root, -1, snap(base_R), np.array(base_t, dtype=np.float64))
    # inertials
    for name, (b, R, t) in link_owner.items():
        l = links[name]
        if b >= 0 and l.mass > 0:
            bodies[b]["parts"].append((l.mass, R @ l.com + t, R @ np.diag(l.idiag) @ R.T))
    for b in bodies:
        m = sum(p[0] for p in b["parts"])
        c = sum(p[0] * p[1] for p in b["parts"]) / m
        Ic = np.zeros((3, 3))
        for (mi, ci, Ii) in b["parts"]:
            d = ci - c
            Ic += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
        b["mass"], b["com"], b["Ic"] = m, c, Ic
        b["Irot"] = snap(Ic + m * (c @ c * np.eye(3) - np.outer(c, c)))
        b["h"] = snap(m * c)
    eb, eR, et = link_owner[ee_link]
    out_s = []
    for (ln, c, r) in spheres:
        b, R, t = link_owner[ln]
        out_s.append((b, snap(R @ np.array(c, dtype=np.float64) + t), r))
    out_b = []
    for (ln, c, h) in boxes:
        b, R, t = link_owner[ln]
        out_b.append((b, snap(R @ np.array(c, dtype=np.float64) + t), snap(R), np.array(h, dtype=np.float64)))
    return dict(bodies=bodies, ee_body=eb, ee_R=snap(eR), ee_t=snap(et), spheres=out_s, boxes=out_b)


def fmt(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    return "{" + ", ".join(repr(float(v)) for v in a) + "}"


def emit_arm(f, P, m, n_arm, rest, kp, maxforce, max_sph, max_box):
    nb = len(m["bodies"])
    B = m["bodies"]
    f.write(f"/* ---- {P}: {nb} moving bodies: " + ", ".join(b["name"] for b in B) + " ---- */\n")
    f.write(f"#define PIH_{P}_NB {nb}\n#define PIH_{P}_NARM {n_arm}\n#define PIH_{P}_EE_BODY {m['ee_body']}\n")
    f.write(f"#define PIH_{P}_NSPH {len(m['spheres'])}\n#define PIH_{P}_NBOX {len(m['boxes'])}\n")
    f.write(f"#define PIH_{P}_PARENT_LIST " + ", ".join(str(b["parent"]) for b in B) + "\n")
    f.write(f"#define PIH_{P}_JTYPE_LIST " + ", ".join(str(b["jtype"]) for b in B) + "\n")
    f.write(f"PIH_TABLE int PIH_{P}_PARENT[{nb}] = {{" + ", ".join(str(b["parent"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE int PIH_{P}_JTYPE[{nb}] = {{" + ", ".join(str(b["jtype"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_R0[{nb}][9] = {{" + ",\n  ".join(fmt(b["R0"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_T0[{nb}][3] = {{" + ", ".join(fmt(b["t0"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_MASS[{nb}] = " + fmt([b["mass"] for b in B]) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_H[{nb}][3] = {{" + ", ".join(fmt(b["h"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_IROT[{nb}][9] = {{" + ",\n  ".join(fmt(b["Irot"]) for b in B) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_LO[{nb}] = " + fmt([b["lo"] for b in B]) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_HI[{nb}] = " + fmt([b["hi"] for b in B]) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_DAMP[{nb}] = " + fmt([b["damping"] for b in B]) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_MAXF[{nb}] = " + fmt(maxforce) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_KP[{nb}] = " + fmt(kp) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_REST[{nb}] = " + fmt(rest) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_EE_R[9] = " + fmt(m["ee_R"]) + ";\n")
    f.write(f"PIH_TABLE double PIH_{P}_EE_T[3] = " + fmt(m["ee_t"]) + ";\n")
    sp = m["spheres"] + [(0, np.zeros(3), 0.0)] * (max_sph - len(m["spheres"]))
    f.write(f"#define PIH_{P}_SPH_BODY_LIST " + ", ".join(str(s[0]) for s in sp) + "\n")
    f.write(f"PIH_TABLE int PIH_{P}_SPH_BODY[{max_sph}] = {{" + ", ".join(str(s[0]) for s in sp) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_SPH_C[{max_sph}][3] = {{" + ", ".join(fmt(s[1]) for s in sp) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_SPH_R[{max_sph}] = " + fmt([s[2] for s in sp]) + ";\n")
    bx = m["boxes"] + [(0, np.zeros(3), np.eye(3), np.zeros(3))] * (max_box - len(m["boxes"]))
    f.write(f"#define PIH_{P}_BOX_BODY_LIST " + ", ".join(str(s[0]) for s in bx) + "\n")
    f.write(f"PIH_TABLE int PIH_{P}_BOX_BODY[{max_box}] = {{" + ", ".join(str(s[0]) for s in bx) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_BOX_C[{max_box}][3] = {{" + ", ".join(fmt(s[1]) for s in bx) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_BOX_R[{max_box}][9] = {{" + ",\n  ".join(fmt(s[2]) for s in bx) + "};\n")
    f.write(f"PIH_TABLE double PIH_{P}_BOX_H[{max_box}][3] = {{" + ", ".join(fmt(s[3]) for s in bx) + "};\n\n")


def main():
    MAX_SPH, MAX_BOX = 8, 2
    # ---- panda (envs/utils.py:30-38: base yaw -pi/2 at `offset`; rest pose envs/peg_in_hole.py:233) ----
    pl, pj, proot = panda_tables()
    p_spheres = [("panda_leftfinger", (0, 0.012, 0.045), 0.012), ("panda_rightfinger", (0, -0.012, 0.045), 0.012),
                 ("panda_hand", (0, 0.06, 0.03), 0.035), ("panda_hand", (0, -0.06, 0.03), 0.035),
                 ("panda_link7", (0, 0, 0.06), 0.05), ("panda_link6", (0.088, 0, 0), 0.06),
                 ("panda_link5", (0, 0.04, -0.1), 0.06)]
    p_boxes = [("panda_leftfinger", (0, 0.01, 0.027), (0.011, 0.01, 0.027)),
               ("panda_rightfinger", (0, -0.01, 0.027), (0.011, 0.01, 0.027))]
    pm = build(pl, pj, proot, rpy_to_R(0, 0, -math.pi / 2), (0, 0, 0), "panda_grasptarget", p_spheres, p_boxes)
    # rest pose: 9 numbers written to joint indices 0..8; 7,8 are FIXED joints in pybullet's panda, so the
    # fingers (indices 9,10) stay at 0.
    p_rest = [0, -0.215, -math.pi / 3, -2.57, 0, 2.356, 2.356, 0.0, 0.0]
    # POSITION_CONTROL with positionGains=[1]*9, default force (envs/utils.py:68); pybullet default max force [3P]
    p_kp = [1.0] * 9
    p_maxf = [100000.0] * 9
    # ---- ur5 (envs/utils.py:40-48: identity base at `ur_pos`; gains 0.03, forces = URDF effort, :82) ----
    ul, uj, uroot = parse_ur5()
    u_spheres = [("ee_link", (0, 0, 0), 0.02), ("wrist_3_link", (0, 0.06, 0), 0.04), ("wrist_2_link", (0, 0, 0.09), 0.045),
                 ("wrist_1_link", (0, 0.09, 0), 0.045), ("forearm_link", (0, 0, 0.39), 0.055),
                 ("forearm_link", (0, 0, 0.0), 0.06)]
    um = build(ul, uj, uroot, np.eye(3), (0, 0, 0), "ee_link", u_spheres, [])
    u_rest = [0.0, -1.57, 1.57, -1.57, -1.57, 0.0]   # NEW (no UR rest pose in tree)
    u_kp = [0.03] * 6
    u_maxf = [b["effort"] for b in um["bodies"]]
    with open(OUT, "w") as f:
        f.write("/* GENERATED by tools/gen_model.py - do not edit.  Model tables (DATA, not code) shared by the CPU oracle\n"
                " * and the CUDA kernels.  Sources: reference ur5.urdf (in tree) and a restated Franka Panda table\n"
                " * ([3P] pybullet_data, from memory, UNVERIFIED).  Frames: body origin at joint origin, z = joint axis. */\n"
                "#ifndef PIH_MODEL_DATA_H\n#define PIH_MODEL_DATA_H\n\n")
        f.write(f"#define PIH_MAX_SPH {MAX_SPH}\n#define PIH_MAX_BOX {MAX_BOX}\n"
                "/* C++: constexpr so that kernels can dispatch on structural zeros at compile time */\n"
                "#ifdef __cplusplus\n#define PIH_TABLE static constexpr\n#else\n#define PIH_TABLE static const\n#endif\n\n")
        emit_arm(f, "PANDA", pm, 7, p_rest, p_kp, p_maxf, MAX_SPH, MAX_BOX)
        emit_arm(f, "UR5", um, 6, u_rest, u_kp, u_maxf, MAX_SPH, MAX_BOX)
        f.write("#endif\n")
    # KAT: UR5 FK vectors of SURVEY.md Appendix C.1 (URDF convention), checked against the emitted tables
    def fk(m, q):
        Rw, pw = {-1: np.eye(3)}, {-1: np.zeros(3)}
        for i, b in enumerate(m["bodies"]):
            c, s = math.cos(q[i]), math.sin(q[i])
            Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
            if b["jtype"] == 0:
                Rl, tl = b["R0"] @ Rz, b["t0"]
            else:
                Rl, tl = b["R0"], b["t0"] + b["R0"][:, 2] * q[i]
            Rw[i] = Rw[b["parent"]] @ Rl
            pw[i] = pw[b["parent"]] + Rw[b["parent"]] @ tl
        e = m["ee_body"]
        return pw[e] + Rw[e] @ m["ee_t"], Rw[e] @ m["ee_R"]
    for q, want in [((0, 0, 0, 0, 0, 0), (-0.814442, -0.190153, 0.070686)),
                    ((.1, .2, .3, .4, .5, .6), (-0.679057, -0.249298, -0.193684)),
                    ((0, -1.57, 1.57, -1.57, -1.57, 0), (-0.497143, -0.108424, 0.517412))]:
        p, R = fk(um, q)
        assert np.allclose(p, want, atol=2e-6), (q, p, want)
    p, R = fk(pm, p_rest)
    print("ur5 FK KATs ok; panda rest EE (env frame):", p, "\nR=", R)
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
