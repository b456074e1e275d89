"""Independent numpy model of the reference's 25-link pipe (assets/urdf/pipe.urdf, globalScaling 0.01) for pinning the
oracle's floating-base ABA.  TEST INFRASTRUCTURE.  Shares no code with oracle/: 6x6 spatial matrices, the mass matrix
from body Jacobians (M = sum J_b^T I_b J_b) and the bias force from a recursive Newton-Euler pass; constants are
restated here from the URDF (joint origins pipe.urdf:39-43,76-83; masses :16,54,87; inertial origin of link1 :53;
inertia = box of the collision AABB + 1 mm about the inertial origin, Bullet's default URDF loading [3P])."""
import numpy as np

NB, ND = 24, 29
R_, HL, PITCH, Y1 = 0.01, 0.03, 0.055, 0.03
M_LINK, M_LINK0, COM1 = 0.0111, 0.00111, 0.015


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)


def box_inertia(m):
    lx, ly = 2 * (R_ + 0.001), 2 * (HL + 0.001)
    return np.diag([m / 12 * (ly * ly + lx * lx), m / 12 * (lx * lx + lx * lx), m / 12 * (ly * ly + lx * lx)])


def spatial_inertia(m, c, Ic):
    """6x6 inertia on (omega; v) about the frame origin for mass m at c with rotational inertia Ic about c."""
    C = skew(c)
    top = np.hstack([Ic + m * C @ C.T, m * C])
    bot = np.hstack([m * C.T, m * np.eye(3)])
    return np.vstack([top, bot])


def body_inertias():
    I0 = spatial_inertia(M_LINK0, np.zeros(3), box_inertia(M_LINK0)) + spatial_inertia(M_LINK, np.array([0, Y1 + COM1, 0.0]), box_inertia(M_LINK))
    return [I0] + [spatial_inertia(M_LINK, np.zeros(3), box_inertia(M_LINK)) for _ in range(1, NB)]


def axis_of(b):
    return 2 if b % 2 == 1 else 0          # pipe_joint(b+1): axis z for even joint numbers (odd b), x otherwise


def joint_rot(b, q):
    c, s = np.cos(q), np.sin(q)
    if axis_of(b) == 2:
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])


def joint_offset(b):
    return np.array([0.0, Y1 + PITCH if b == 1 else PITCH, 0.0])


def motion_xf(R, r):
    """parent -> child coordinates of a motion vector: omega' = R^T omega, v' = R^T (v + omega x r)."""
    X = np.zeros((6, 6))
    X[:3, :3] = R.T
    X[3:, 3:] = R.T
    X[3:, :3] = -R.T @ skew(r)
    return X


def crossm(v):
    X = np.zeros((6, 6))
    X[:3, :3] = skew(v[:3]); X[3:, 3:] = skew(v[:3]); X[3:, :3] = skew(v[3:])
    return X


def mass_matrix_and_bias(Rb, qj, nu, g):
    """nu = (omega_b, v_b, qdot[23]) with the base twist in BODY coordinates.  Returns M (29x29) and the generalized bias
    h (Coriolis / centrifugal + gravity) such that M nudot + h = 0 for the unforced chain."""
    I = body_inertias()
    X = [None] * NB
    S = [None] * NB
    J = [np.zeros((6, ND)) for _ in range(NB)]
    J[0][:, :6] = np.eye(6)
    for b in range(1, NB):
        X[b] = motion_xf(joint_rot(b, qj[b - 1]), joint_offset(b))
        S[b] = np.zeros(6); S[b][axis_of(b)] = 1.0
        J[b] = X[b] @ J[b - 1]
        J[b][:, 5 + b] += S[b]
    M = sum(J[b].T @ I[b] @ J[b] for b in range(NB))
    # recursive Newton-Euler with zero generalized acceleration; gravity as a base acceleration of +g along world z
    v = [None] * NB
    a = [None] * NB
    f = [None] * NB
    v[0] = nu[:6].copy()
    a[0] = np.concatenate([np.zeros(3), Rb.T @ np.array([0.0, 0.0, g])])
    for b in range(1, NB):
        vj = S[b] * nu[5 + b]
        v[b] = X[b] @ v[b - 1] + vj
        a[b] = X[b] @ a[b - 1] + crossm(v[b]) @ vj
    for b in range(NB):
        f[b] = I[b] @ a[b] - crossm(v[b]).T @ (I[b] @ v[b])
    h = np.zeros(ND)
    for b in range(NB - 1, 0, -1):
        h[5 + b] = S[b] @ f[b]
        f[b - 1] = f[b - 1] + X[b].T @ f[b]
    h[:6] = f[0]
    return M, h


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def fk(bpos, Rb, qj):
    pos, rot = [np.asarray(bpos, float)], [Rb]
    for b in range(1, NB):
        pos.append(pos[b - 1] + rot[b - 1] @ joint_offset(b))
        rot.append(rot[b - 1] @ joint_rot(b, qj[b - 1]))
    return np.array(pos), np.array(rot)
