/* CPU ORACLE, flexible pipe - TEST INFRASTRUCTURE ONLY (included by pih_oracle_impl.h, once per REAL type).
 *
 * PARITY UNPINNED vs PyBullet (see pih_oracle_impl.h).  Restates, for the reference's 25-link "peg"
 *   peg_in_hole_gym/envs/assets/urdf/pipe.urdf                 (links, joints, masses, friction)
 *   peg_in_hole_gym/envs/peg_in_hole.py:239-245                (spawn pose, URDF_USE_SELF_COLLISION, random bends)
 *   peg_in_hole_gym/envs/peg_in_hole.py:266, 58, 114-117       (grasp link, its pose, success test on it)
 * the published algorithms PyBullet runs on it [3P]: Featherstone ABA for a floating-base chain
 * (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof), unit-impulse responses from the same
 * factorisation (calcAccelerationDeltasMultiDof), the default velocity motor of every movable joint (target 0, max
 * impulse 1 per step, PhysicsServerCommandProcessor::createJointMotors), and sequential impulses over motors, limits and
 * contacts (btMultiBodyConstraintSolver).  Deliberate deviations: no self-collision between segments,
 * analytic cylinders / annulus instead of hull / trimesh manifolds, <= PIH_PIPE_SEG_MAXC contact points per segment,
 * arm and pipe solved as ONE island (Bullet batches small islands into one solver call as well).
 *
 * Numerics: the CUDA kernel maps the 29 degrees of freedom to the lanes of a warp.  Sums over degrees of freedom are
 * therefore DEFINED as the 32-leaf butterfly of the warp shuffle (FN(bfly32)); everything else is a fixed chain of
 * explicit FMAs, so that the float build and the kernel agree bit for bit. */

#define PNB PIH_PIPE_NBODY
#define PND PIH_PIPE_NDOF

typedef struct {
  real bpos[3], bquat[4], bvel[3], bang[3];
  real qj[PNB], qdj[PNB]; /* [0] unused */
  int32_t grasp_idx;
} FN(pipestate);
#define pipestate FN(pipestate)

/* constants of the chain in REAL precision */
typedef struct {
  real Ix, Iy;          /* link inertia about its frame origin: Ixx = Izz, Iyy (box of the AABB + 1 mm, B.1) */
  real m;               /* link mass */
  real I0x, I0y, I0z;   /* base = link0 + link1 about the link0 origin */
  real m0, h0y;         /* base mass, first moment (0, h0y, 0) */
} FN(pipemodel);
#define pipemodel FN(pipemodel)

static void FN(pipemodel_init)(pipemodel* P) {
  const double lx = 2 * (PIH_PIPE_R + 0.001), ly = 2 * (PIH_PIPE_HL + 0.001);
  const double m = PIH_PIPE_LINK_MASS, m0 = PIH_PIPE_LINK0_MASS;
  const double ix = m / 12.0 * (ly * ly + lx * lx), iy = m / 12.0 * (lx * lx + lx * lx);
  const double ix0 = m0 / 12.0 * (ly * ly + lx * lx), iy0 = m0 / 12.0 * (lx * lx + lx * lx);
  const double c1 = PIH_PIPE_LINK1_Y + PIH_PIPE_LINK1_COM_Y;   /* link1 COM in the base frame */
  P->Ix = RC(ix); P->Iy = RC(iy); P->m = RC(m);
  P->I0x = RC(ix0 + ix + m * c1 * c1); P->I0y = RC(iy0 + iy); P->I0z = RC(ix0 + ix + m * c1 * c1);
  P->m0 = RC(m0 + m); P->h0y = RC(m * c1);
}

#define PIPE_AXIS(b) (((b) & 1) ? 2 : 0)                         /* joint of body b: z for odd b, x for even b */
#define PIPE_LEN(b) (((b) == 1) ? RC(PIH_PIPE_LINK1_Y + PIH_PIPE_PITCH) : RC(PIH_PIPE_PITCH))

/* sum of 32 leaves in the order of a warp xor-butterfly (offsets 16, 8, 4, 2, 1); every lane of the warp ends with
 * this value */
static inline real FN(bfly32)(const real* p) {
  real s[32], t[32];
  for (int i = 0; i < 32; ++i) s[i] = p[i];
  for (int off = 16; off >= 1; off >>= 1) {
    for (int i = 0; i < 32; ++i) t[i] = s[i] + s[i ^ off];
    for (int i = 0; i < 32; ++i) s[i] = t[i];
  }
  return s[0];
}

/* R^T u for the joint rotation about z / x by angle with (s, c) */
static inline v3 FN(pipe_rotT)(int axis, real s, real c, v3 u) {
  if (axis == 2) return FN(v3_make)(FMA(s, u.y, c * u.x), FMA(c, u.y, -(s * u.x)), u.z);
  return FN(v3_make)(u.x, FMA(s, u.z, c * u.y), FMA(c, u.z, -(s * u.y)));
}
/* R u */
static inline v3 FN(pipe_rot)(int axis, real s, real c, v3 u) {
  if (axis == 2) return FN(v3_make)(FMA(-s, u.y, c * u.x), FMA(c, u.y, s * u.x), u.z);
  return FN(v3_make)(u.x, FMA(-s, u.z, c * u.y), FMA(c, u.z, s * u.y));
}
static inline real FN(v3_get)(v3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
static inline void FN(v3_addk)(v3* v, int k, real d) { if (k == 0) v->x = v->x + d; else if (k == 1) v->y = v->y + d; else v->z = v->z + d; }

typedef struct {
  m3 Rw[PNB]; v3 pw[PNB];
  real sn[PNB], cs[PNB];
} FN(pipefk);
#define pipefk FN(pipefk)

static void FN(pipe_fk)(const pipestate* PS, pipefk* K) {
  K->Rw[0] = FN(quat_to_m3)(PS->bquat);
  K->pw[0] = FN(v3_make)(PS->bpos[0], PS->bpos[1], PS->bpos[2]);
  K->sn[0] = RC(0); K->cs[0] = RC(1);
  for (int b = 1; b < PNB; ++b) {
    const m3* P = &K->Rw[b - 1];
    const real L = PIPE_LEN(b);
    K->pw[b] = FN(v3_make)(FMA(L, P->m[0][1], K->pw[b - 1].x), FMA(L, P->m[1][1], K->pw[b - 1].y), FMA(L, P->m[2][1], K->pw[b - 1].z));
    real s, c;
    FN(det_sincos)(PS->qj[b], &s, &c);
    K->sn[b] = s; K->cs[b] = c;
    m3* R = &K->Rw[b];
    for (int i = 0; i < 3; ++i) {
      if (PIPE_AXIS(b) == 2) {
        R->m[i][0] = FMA(s, P->m[i][1], c * P->m[i][0]);
        R->m[i][1] = FMA(c, P->m[i][1], -(s * P->m[i][0]));
        R->m[i][2] = P->m[i][2];
      } else {
        R->m[i][0] = P->m[i][0];
        R->m[i][1] = FMA(s, P->m[i][2], c * P->m[i][1]);
        R->m[i][2] = FMA(c, P->m[i][2], -(s * P->m[i][1]));
      }
    }
  }
}

/* pose of the grasp link's centre-of-mass frame = getLinkState(objectUid, grasp_joint_idx)[0:2] (peg_in_hole.py:58,115):
 * index 0 = pipe_link1 (inertial origin (0,1.5,0), pipe.urdf:53), index 23 = pipe_link24 (inertial origin = link frame) */
static void FN(pipe_grasp_pose)(const pipestate* PS, const pipefk* K, real pos[3], real quat[4]) {
  if (PS->grasp_idx == 0) {
    const real cy = RC(PIH_PIPE_LINK1_Y + PIH_PIPE_LINK1_COM_Y);
    pos[0] = FMA(cy, K->Rw[0].m[0][1], K->pw[0].x); pos[1] = FMA(cy, K->Rw[0].m[1][1], K->pw[0].y); pos[2] = FMA(cy, K->Rw[0].m[2][1], K->pw[0].z);
    FN(m3_to_quat)(&K->Rw[0], quat);
  } else {
    pos[0] = K->pw[PNB - 1].x; pos[1] = K->pw[PNB - 1].y; pos[2] = K->pw[PNB - 1].z;
    FN(m3_to_quat)(&K->Rw[PNB - 1], quat);
  }
}

/* dense joint rotation (child -> parent) */
static inline m3 FN(pipe_Rl)(int axis, real s, real c) {
  m3 R;
  if (axis == 2) { m3 T = {{{c, -s, RC(0)}, {s, c, RC(0)}, {RC(0), RC(0), RC(1)}}}; R = T; }
  else { m3 T = {{{RC(1), RC(0), RC(0)}, {RC(0), c, -s}, {RC(0), s, c}}}; R = T; }
  return R;
}

/* motion transform parent -> child for a joint at r = (0, L, 0) */
static inline sv FN(pipe_motion_to_child)(int axis, real s, real c, real L, sv m) {
  v3 t = FN(v3_make)(FMA(-m.a.z, L, m.l.x), m.l.y, FMA(m.a.x, L, m.l.z));
  sv o;
  o.a = FN(pipe_rotT)(axis, s, c, m.a);
  o.l = FN(pipe_rotT)(axis, s, c, t);
  return o;
}
/* force transform child -> parent: f_p = R f, n_p = R n + r x f_p with r = (0, L, 0) */
static inline sv FN(pipe_force_to_parent)(int axis, real s, real c, real L, sv f) {
  sv o;
  o.l = FN(pipe_rot)(axis, s, c, f.l);
  v3 n = FN(pipe_rot)(axis, s, c, f.a);
  o.a = FN(v3_make)(FMA(L, o.l.z, n.x), n.y, FMA(-L, o.l.x, n.z));
  return o;
}

/* solve K x = rhs for the symmetric positive definite 6x6 base inertia, given its factorisation K = L D L^T */
typedef struct { real L[6][6], invd[6]; } FN(ldl6);
#define ldl6 FN(ldl6)
static void FN(ldl6_factor)(const m3* A, const m3* B, const m3* Cm, ldl6* Fz) {
  real Kl[6][6], W[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      if (i < 3) Kl[i][j] = A->m[i][j];
      else if (j < 3) Kl[i][j] = B->m[j][i - 3];
      else Kl[i][j] = Cm->m[i - 3][j - 3];
    }
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < i; ++j) {
      real s = Kl[i][j];
      for (int kk = 0; kk < j; ++kk) s = FMA(-W[i][kk], Fz->L[j][kk], s);
      W[i][j] = s;
      Fz->L[i][j] = s * Fz->invd[j];
    }
    real d = Kl[i][i];
    for (int kk = 0; kk < i; ++kk) d = FMA(-W[i][kk], Fz->L[i][kk], d);
    Fz->invd[i] = RC(1) / d;
  }
}
static void FN(ldl6_solve)(const ldl6* Fz, const real* rhs, real* x) {
  real y[6];
  for (int i = 0; i < 6; ++i) {
    real s = rhs[i];
    for (int kk = 0; kk < i; ++kk) s = FMA(-Fz->L[i][kk], y[kk], s);
    y[i] = s;
  }
  for (int i = 5; i >= 0; --i) {
    real s = y[i] * Fz->invd[i];
    for (int kk = i + 1; kk < 6; ++kk) s = FMA(-Fz->L[kk][i], y[kk], s);
    y[i] = s;
  }
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

typedef struct {
  sv U[PNB], UD[PNB];
  real Dinv[PNB], u[PNB];
  sv cb[PNB];
  ldl6 base;
  real vstar[PND];        /* unconstrained velocities after dt: base (body coordinates: angular, linear), joints */
  real col[PND][PND];     /* col[i][j] = velocity change of dof j for a unit impulse on dof i */
} FN(pipedyn);
#define pipedyn FN(pipedyn)

/* ABA of the floating chain: unconstrained velocities and the unit-impulse responses of all 29 degrees of freedom */
static void FN(pipe_dynamics)(const pipemodel* PM, const pipestate* PS, const pipefk* K, real dt, real g, pipedyn* D) {
  sv vel[PNB], pA[PNB];
  m3 IA[PNB], IB[PNB], IC[PNB];
  /* (B) velocities in body coordinates, velocity-product accelerations, bias forces */
  {
    v3 ww = FN(v3_make)(PS->bang[0], PS->bang[1], PS->bang[2]), vw = FN(v3_make)(PS->bvel[0], PS->bvel[1], PS->bvel[2]);
    vel[0].a = FN(m3T_v)(&K->Rw[0], ww);
    vel[0].l = FN(m3T_v)(&K->Rw[0], vw);
    D->cb[0].a = FN(v3_make)(0, 0, 0); D->cb[0].l = D->cb[0].a;
    v3 h = FN(v3_make)(RC(0), PM->h0y, RC(0));
    v3 w = vel[0].a, v = vel[0].l;
    v3 La = FN(v3_add)(FN(v3_make)(PM->I0x * w.x, PM->I0y * w.y, PM->I0z * w.z), FN(v3_cross)(h, v));
    v3 hw = FN(v3_cross)(h, w);
    v3 Ll = FN(v3_make)(FMA(PM->m0, v.x, -hw.x), FMA(PM->m0, v.y, -hw.y), FMA(PM->m0, v.z, -hw.z));
    pA[0].a = FN(v3_add)(FN(v3_cross)(w, La), FN(v3_cross)(v, Ll));
    pA[0].l = FN(v3_cross)(w, Ll);
    m3 A0 = {{{PM->I0x, RC(0), RC(0)}, {RC(0), PM->I0y, RC(0)}, {RC(0), RC(0), PM->I0z}}};
    m3 B0 = {{{RC(0), -h.z, h.y}, {h.z, RC(0), -h.x}, {-h.y, h.x, RC(0)}}};
    m3 C0 = {{{PM->m0, RC(0), RC(0)}, {RC(0), PM->m0, RC(0)}, {RC(0), RC(0), PM->m0}}};
    IA[0] = A0; IB[0] = B0; IC[0] = C0;
  }
  for (int b = 1; b < PNB; ++b) {
    const int ax = PIPE_AXIS(b);
    sv vp = FN(pipe_motion_to_child)(ax, K->sn[b], K->cs[b], PIPE_LEN(b), vel[b - 1]);
    const real w = PS->qdj[b];
    if (ax == 2) { D->cb[b].a = FN(v3_make)(vp.a.y * w, -(vp.a.x * w), RC(0)); D->cb[b].l = FN(v3_make)(vp.l.y * w, -(vp.l.x * w), RC(0)); }
    else { D->cb[b].a = FN(v3_make)(RC(0), vp.a.z * w, -(vp.a.y * w)); D->cb[b].l = FN(v3_make)(RC(0), vp.l.z * w, -(vp.l.y * w)); }
    sv v = vp;
    FN(v3_addk)(&v.a, ax, w);
    vel[b] = v;
    v3 Iw = FN(v3_make)(PM->Ix * v.a.x, PM->Iy * v.a.y, PM->Ix * v.a.z);
    v3 mv = FN(v3_make)(PM->m * v.l.x, PM->m * v.l.y, PM->m * v.l.z);
    pA[b].a = FN(v3_cross)(v.a, Iw);
    pA[b].l = FN(v3_cross)(v.a, mv);
    m3 A0 = {{{PM->Ix, RC(0), RC(0)}, {RC(0), PM->Iy, RC(0)}, {RC(0), RC(0), PM->Ix}}};
    m3 Z = {{{RC(0), RC(0), RC(0)}, {RC(0), RC(0), RC(0)}, {RC(0), RC(0), RC(0)}}};
    m3 C0 = {{{PM->m, RC(0), RC(0)}, {RC(0), PM->m, RC(0)}, {RC(0), RC(0), PM->m}}};
    IA[b] = A0; IB[b] = Z; IC[b] = C0;
  }
  /* (C) inward pass: articulated inertias and bias forces (no joint torques: pipe.urdf has no damping, tau = 0) */
  for (int b = PNB - 1; b >= 1; --b) {
    const int ax = PIPE_AXIS(b);
    sv U;
    U.a = FN(m3_col)(&IA[b], ax); U.l = FN(m3_row)(&IB[b], ax);
    const real Dd = IA[b].m[ax][ax];
    const real u = -FN(v3_get)(pA[b].a, ax);
    const real Dinv = RC(1) / Dd;
    sv UD;
    UD.a = FN(v3_scale)(U.a, Dinv); UD.l = FN(v3_scale)(U.l, Dinv);
    D->U[b] = U; D->UD[b] = UD; D->Dinv[b] = Dinv; D->u[b] = u;
    real uda[3] = {UD.a.x, UD.a.y, UD.a.z}, udl[3] = {UD.l.x, UD.l.y, UD.l.z};
    real ua[3] = {U.a.x, U.a.y, U.a.z}, ul[3] = {U.l.x, U.l.y, U.l.z};
    m3 A1, B1, C1;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        A1.m[r][c] = FMA(-uda[r], ua[c], IA[b].m[r][c]);
        B1.m[r][c] = FMA(-uda[r], ul[c], IB[b].m[r][c]);
        C1.m[r][c] = FMA(-udl[r], ul[c], IC[b].m[r][c]);
      }
    sv pa;
    v3 t1 = FN(v3_add)(FN(m3_v)(&A1, D->cb[b].a), FN(m3_v)(&B1, D->cb[b].l));
    v3 t2 = FN(v3_add)(FN(m3T_v)(&B1, D->cb[b].a), FN(m3_v)(&C1, D->cb[b].l));
    pa.a = FN(v3_axpy)(u, UD.a, FN(v3_add)(pA[b].a, t1));
    pa.l = FN(v3_axpy)(u, UD.l, FN(v3_add)(pA[b].l, t2));
    const m3 Rl = FN(pipe_Rl)(ax, K->sn[b], K->cs[b]);
    m3 T = FN(m3_mul)(&Rl, &A1); m3 Ar = FN(m3_mulT)(&T, &Rl);
    T = FN(m3_mul)(&Rl, &B1); m3 Br = FN(m3_mulT)(&T, &Rl);
    T = FN(m3_mul)(&Rl, &C1); m3 Cr = FN(m3_mulT)(&T, &Rl);
    v3 r = FN(v3_make)(RC(0), PIPE_LEN(b), RC(0));
    m3 Bp, Ap;
    for (int c = 0; c < 3; ++c) {
      v3 cc = FN(v3_cross)(r, FN(m3_col)(&Cr, c));
      Bp.m[0][c] = Br.m[0][c] + cc.x; Bp.m[1][c] = Br.m[1][c] + cc.y; Bp.m[2][c] = Br.m[2][c] + cc.z;
    }
    for (int c = 0; c < 3; ++c) {
      v3 cc = FN(v3_cross)(r, FN(m3_row)(&Br, c));
      Ap.m[0][c] = Ar.m[0][c] + cc.x; Ap.m[1][c] = Ar.m[1][c] + cc.y; Ap.m[2][c] = Ar.m[2][c] + cc.z;
    }
    for (int rr = 0; rr < 3; ++rr) {
      v3 cc = FN(v3_cross)(FN(m3_row)(&Bp, rr), r);
      Ap.m[rr][0] = Ap.m[rr][0] - cc.x; Ap.m[rr][1] = Ap.m[rr][1] - cc.y; Ap.m[rr][2] = Ap.m[rr][2] - cc.z;
    }
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) {
        IA[b - 1].m[rr][c] = IA[b - 1].m[rr][c] + Ap.m[rr][c];
        IB[b - 1].m[rr][c] = IB[b - 1].m[rr][c] + Bp.m[rr][c];
        IC[b - 1].m[rr][c] = IC[b - 1].m[rr][c] + Cr.m[rr][c];
      }
    sv fp = FN(pipe_force_to_parent)(ax, K->sn[b], K->cs[b], PIPE_LEN(b), pa);
    pA[b - 1].a = FN(v3_add)(pA[b - 1].a, fp.a);
    pA[b - 1].l = FN(v3_add)(pA[b - 1].l, fp.l);
  }
  /* (D) base: K a0 = -pA0 (gravity enters as a uniform field afterwards: the joint accelerations do not see it) */
  FN(ldl6_factor)(&IA[0], &IB[0], &IC[0], &D->base);
  real rhs[6] = {-pA[0].a.x, -pA[0].a.y, -pA[0].a.z, -pA[0].l.x, -pA[0].l.y, -pA[0].l.z}, a0[6];
  FN(ldl6_solve)(&D->base, rhs, a0);
  /* (E) outward pass */
  {
    const m3* R0 = &K->Rw[0];
    real gb[3] = {-(g * R0->m[2][0]), -(g * R0->m[2][1]), -(g * R0->m[2][2])};
    D->vstar[0] = FMA(dt, a0[0], vel[0].a.x); D->vstar[1] = FMA(dt, a0[1], vel[0].a.y); D->vstar[2] = FMA(dt, a0[2], vel[0].a.z);
    D->vstar[3] = FMA(dt, a0[3] + gb[0], vel[0].l.x); D->vstar[4] = FMA(dt, a0[4] + gb[1], vel[0].l.y); D->vstar[5] = FMA(dt, a0[5] + gb[2], vel[0].l.z);
    sv acc;
    acc.a = FN(v3_make)(a0[0], a0[1], a0[2]); acc.l = FN(v3_make)(a0[3], a0[4], a0[5]);
    for (int b = 1; b < PNB; ++b) {
      const int ax = PIPE_AXIS(b);
      sv a = FN(pipe_motion_to_child)(ax, K->sn[b], K->cs[b], PIPE_LEN(b), acc);
      a.a = FN(v3_add)(a.a, D->cb[b].a);
      a.l = FN(v3_add)(a.l, D->cb[b].l);
      real ua = FN(v3_dot)(D->U[b].a, a.a) + FN(v3_dot)(D->U[b].l, a.l);
      real qdd = D->Dinv[b] * (D->u[b] - ua);
      FN(v3_addk)(&a.a, ax, qdd);
      acc = a;
      D->vstar[5 + b] = FMA(dt, qdd, PS->qdj[b]);
    }
  }
  /* (F) unit-impulse responses: dof i = 0..5 base (angular, linear; body coordinates), 5 + b = joint of body b */
  for (int i = 0; i < PND; ++i) {
    sv ph[PNB];
    real uh[PNB];
    for (int b = 0; b < PNB; ++b) { ph[b].a = FN(v3_make)(0, 0, 0); ph[b].l = ph[b].a; uh[b] = RC(0); }
    for (int b = PNB - 1; b >= 1; --b) {
      const int ax = PIPE_AXIS(b);
      real sp = FN(v3_get)(ph[b].a, ax);
      uh[b] = ((5 + b == i) ? RC(1) : RC(0)) - sp;
      sv pa;
      pa.a = FN(v3_axpy)(uh[b], D->UD[b].a, ph[b].a);
      pa.l = FN(v3_axpy)(uh[b], D->UD[b].l, ph[b].l);
      sv fp = FN(pipe_force_to_parent)(ax, K->sn[b], K->cs[b], PIPE_LEN(b), pa);
      ph[b - 1].a = FN(v3_add)(ph[b - 1].a, fp.a);
      ph[b - 1].l = FN(v3_add)(ph[b - 1].l, fp.l);
    }
    real r6[6] = {ph[0].a.x, ph[0].a.y, ph[0].a.z, ph[0].l.x, ph[0].l.y, ph[0].l.z}, x6[6];
    for (int j = 0; j < 6; ++j) r6[j] = ((j == i) ? RC(1) : RC(0)) - r6[j];
    FN(ldl6_solve)(&D->base, r6, x6);
    for (int j = 0; j < 6; ++j) D->col[i][j] = x6[j];
    sv acc;
    acc.a = FN(v3_make)(x6[0], x6[1], x6[2]); acc.l = FN(v3_make)(x6[3], x6[4], x6[5]);
    for (int b = 1; b < PNB; ++b) {
      const int ax = PIPE_AXIS(b);
      sv a = FN(pipe_motion_to_child)(ax, K->sn[b], K->cs[b], PIPE_LEN(b), acc);
      real ua = FN(v3_dot)(D->U[b].a, a.a) + FN(v3_dot)(D->U[b].l, a.l);
      real d = D->Dinv[b] * (uh[b] - ua);
      FN(v3_addk)(&a.a, ax, d);
      acc = a;
      D->col[i][5 + b] = d;
    }
  }
}

/* ------------------------------------------------------------------ contacts of the pipe task */
typedef struct {
  int pair;
  int seg;         /* >= 0: pipe segment involved (normal points INTO the segment) */
  int arm_body;    /* >= 0: arm body involved (normal points away from it when seg >= 0, else into it) */
  v3 p, n;
  real dist, mu;
} FN(pcontact);
#define pcontact FN(pcontact)

#define PIPE_SEG_BODY(s) ((s) < 2 ? 0 : (s) - 1)
#define PIPE_SEG_CY(s) ((s) == 1 ? RC(PIH_PIPE_LINK1_Y + PIH_PIPE_HL) : RC(PIH_PIPE_HL))   /* centre of the cylinder on the body's y axis */
#define PIPE_SEG_MU(s) (((s) < 2 || (s) > 22) ? RC(PIH_PIPE_MU_END) : RC(PIH_MU_DEFAULT))

/* keep the `cap` deepest candidates of a slot list; ties keep the earlier candidate (same rule as contact_push) */
static inline void FN(pcontact_push)(pcontact* slot, int* n, int cap, const pcontact* c) {
  if (!(c->dist < RC(PIH_CONTACT_MARGIN))) return;
  if (*n < cap) { slot[(*n)++] = *c; return; }
  int w = 0;
  for (int i = 1; i < cap; ++i) if (slot[i].dist > slot[w].dist) w = i;
  if (c->dist < slot[w].dist) slot[w] = *c;
}

/* signed distance / outward normal (env frame) of segment with centre cw, rotation R (axis = column 1) at point x */
static inline real FN(seg_sdf)(const m3* R, v3 cw, v3 x, v3* n_env) {
  v3 l = FN(m3T_v)(R, FN(v3_sub)(x, cw));
  real rho = SQRT(FMA(l.z, l.z, l.x * l.x));
  real da = rho - RC(PIH_PIPE_R), db = ABS(l.y) - RC(PIH_PIPE_HL);
  real na, nb_;
  real d = FN(sd_rect)(da, db, RC(1), (l.y < RC(0)) ? RC(-1) : RC(1), &na, &nb_);
  real ir = (rho > RC(1e-12)) ? RC(1) / rho : RC(0);
  v3 nl = FN(v3_make)(na * l.x * ir, nb_, na * l.z * ir);
  *n_env = FN(m3_v)(R, nl);
  return d;
}

/* candidates of segment s against the table, the hole, the finger boxes and the arm spheres, in that order */
static void FN(pipe_collide_segment)(const model* M, const fkres* F, const pipefk* K, int s, pcontact* slot, int* n) {
  static const double cs8[8][2] = {{1, 0}, {0.7071067811865476, 0.7071067811865476}, {0, 1}, {-0.7071067811865476, 0.7071067811865476},
                                   {-1, 0}, {-0.7071067811865476, -0.7071067811865476}, {0, -1}, {0.7071067811865476, -0.7071067811865476}};
  const int body = PIPE_SEG_BODY(s);
  const m3* R = &K->Rw[body];
  const real cy = PIPE_SEG_CY(s);
  const v3 cw = FN(v3_make)(FMA(cy, R->m[0][1], K->pw[body].x), FMA(cy, R->m[1][1], K->pw[body].y), FMA(cy, R->m[2][1], K->pw[body].z));
  const real mu_seg = PIPE_SEG_MU(s);
  const real hl = RC(PIH_PIPE_HL), rad = RC(PIH_PIPE_R);
  pcontact c;
  c.seg = s;
  *n = 0;
  /* (1) table plane */
  if (cw.z - RC(PIH_TABLE_Z) < RC(0.05)) {
    v3 ax = FN(m3_col)(R, 1);
    real naz = ax.z;
    v3 nperp = FN(v3_make)(-(naz * ax.x), -(naz * ax.y), FMA(-naz, ax.z, RC(1)));
    real l2 = FN(v3_dot)(nperp, nperp);
    c.pair = PIH_PAIR_PEG_TABLE; c.arm_body = -1; c.n = FN(v3_make)(0, 0, 1); c.mu = mu_seg * RC(PIH_MU_DEFAULT);
    if (l2 > RC(0.0025)) {
      real il = RC(1) / SQRT(l2);
      v3 t = FN(v3_scale)(nperp, il);
      for (int k = 0; k < 2; ++k) {
        real sgn = k ? RC(-1) : RC(1);
        v3 cc = FN(v3_axpy)(sgn * hl, ax, cw);
        c.p = FN(v3_axpy)(-rad, t, cc);
        c.dist = c.p.z - RC(PIH_TABLE_Z);
        FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
      }
    } else {
      real sgn = (naz > RC(0)) ? RC(-1) : RC(1);
      v3 cc = FN(v3_axpy)(sgn * hl, ax, cw);
      v3 ex = FN(m3_col)(R, 0), ez = FN(m3_col)(R, 2);
      for (int k = 0; k < 4; ++k) {
        c.p = FN(v3_axpy)((k & 1) ? -rad : rad, (k & 2) ? ez : ex, cc);
        c.dist = c.p.z - RC(PIH_TABLE_Z);
        FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
      }
    }
  }
  /* (2) annular hole */
  {
    v3 dh = FN(v3_sub)(cw, FN(v3_make)(RC(PIH_HOLE_X), RC(PIH_HOLE_Y), RC(PIH_HOLE_Z)));
    const real reach = RC(PIH_PIPE_HL + PIH_PIPE_R + PIH_HOLE_ROUT + PIH_HOLE_HL + PIH_CONTACT_MARGIN);
    if (FN(v3_dot)(dh, dh) < reach * reach) {
      c.pair = PIH_PAIR_PEG_HOLE; c.arm_body = -1; c.mu = mu_seg * RC(PIH_MU_DEFAULT);
      for (int k = 0; k < 16; ++k) {
        real sy = (k < 8) ? RC(1) : RC(-1);
        v3 pl = FN(v3_make)(rad * RC(cs8[k & 7][0]), sy * hl, rad * RC(cs8[k & 7][1]));
        v3 pt = FN(m3_v_add)(R, pl, cw);
        v3 nn;
        c.dist = FN(hole_sdf)(pt, &nn);
        c.p = pt; c.n = nn;
        FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
      }
      for (int k = 0; k < 32; ++k) {
        real rr = (k < 16) ? RC(PIH_HOLE_RIN) : RC(PIH_HOLE_ROUT);
        real sx = (k & 8) ? RC(-PIH_HOLE_HL) : RC(PIH_HOLE_HL);
        v3 pt = FN(v3_make)(RC(PIH_HOLE_X) + sx, FMA(rr, RC(cs8[k & 7][0]), RC(PIH_HOLE_Y)), FMA(rr, RC(cs8[k & 7][1]), RC(PIH_HOLE_Z)));
        v3 nn;
        c.dist = FN(seg_sdf)(R, cw, pt, &nn);
        c.p = pt; c.n = FN(v3_make)(-nn.x, -nn.y, -nn.z);
        FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
      }
    }
  }
  /* (3) finger boxes: five spheres on the segment axis against each box */
  for (int b = 0; b < M->nbox; ++b) {
    int abody = M->box_body[b];
    v3 cb = FN(m3_v_add)(&F->Rw[abody], M->box_c[b], F->pw[abody]);
    v3 dd = FN(v3_sub)(cw, cb);
    real reach = hl + rad + M->box_h[b].z + M->box_h[b].x + RC(PIH_CONTACT_MARGIN);
    if (!(FN(v3_dot)(dd, dd) < reach * reach)) continue;
    m3 Rb = FN(m3_mul)(&F->Rw[abody], &M->box_R[b]);
    for (int k = 0; k < 5; ++k) {
      v3 pl = FN(v3_make)(RC(0), (hl - rad) * RC(0.5) * (real)(k - 2), RC(0));
      v3 pt = FN(m3_v_add)(R, pl, cw);
      v3 l = FN(m3T_v)(&Rb, FN(v3_sub)(pt, cb));
      real dx = ABS(l.x) - M->box_h[b].x, dy = ABS(l.y) - M->box_h[b].y, dz = ABS(l.z) - M->box_h[b].z;
      real sx = (l.x < RC(0)) ? RC(-1) : RC(1), sy = (l.y < RC(0)) ? RC(-1) : RC(1), sz = (l.z < RC(0)) ? RC(-1) : RC(1);
      real px = dx > RC(0) ? dx : RC(0), py = dy > RC(0) ? dy : RC(0), pz = dz > RC(0) ? dz : RC(0);
      real o2 = FMA(pz, pz, FMA(py, py, px * px));
      v3 nl;
      real d;
      if (o2 > RC(0)) { d = SQRT(o2); real il = RC(1) / d; nl = FN(v3_make)(sx * px * il, sy * py * il, sz * pz * il); }
      else if (dx > dy && dx > dz) { d = dx; nl = FN(v3_make)(sx, 0, 0); }
      else if (dy > dz) { d = dy; nl = FN(v3_make)(0, sy, 0); }
      else { d = dz; nl = FN(v3_make)(0, 0, sz); }
      v3 nn = FN(m3_v)(&Rb, nl);
      c.pair = b ? PIH_PAIR_PEG_FINGER2 : PIH_PAIR_PEG_FINGER1; c.arm_body = abody;
      c.p = FN(v3_axpy)(-rad, nn, pt); c.n = nn; c.dist = d - rad; c.mu = mu_seg * RC(PIH_MU_FINGER);
      FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
    }
  }
  /* (3b) arm collision spheres of the arm chain (hand, wrist, forearm) against the segment */
  for (int q = 0; q < M->nsph; ++q) {
    int abody = M->sph_body[q];
    if (abody >= M->narm) continue;
    v3 sw = FN(m3_v_add)(&F->Rw[abody], M->sph_c[q], F->pw[abody]);
    v3 dd = FN(v3_sub)(sw, cw);
    real reach = hl + rad + M->sph_r[q] + RC(PIH_CONTACT_MARGIN);
    if (!(FN(v3_dot)(dd, dd) < reach * reach)) continue;
    v3 nn;
    real d = FN(seg_sdf)(R, cw, sw, &nn);
    c.pair = PIH_PAIR_FB_ARM; c.arm_body = abody;
    c.n = FN(v3_make)(-nn.x, -nn.y, -nn.z);
    c.p = FN(v3_axpy)(-M->sph_r[q], nn, sw); c.dist = d - M->sph_r[q]; c.mu = mu_seg * RC(PIH_MU_DEFAULT);
    FN(pcontact_push)(slot, n, PIH_PIPE_SEG_MAXC, &c);
  }
}

/* one contact point = three constraint rows (normal, two tangents) over the 29 pipe + nb arm degrees of freedom */
typedef struct {
  int has_arm, has_pipe;
  real Jp[3][32], MJp[3][32];      /* lanes 29..31 are zero */
  real Ja[3][NBM], MJa[3][NBM];
  real dinv[3], diag[3], b[3], lam[3], mu;
} FN(prow);
#define prow FN(prow)

/* One env.step with the flexible pipe as the peg.  Same contract as step_env. */
static int FN(step_env_pipe)(const model* M, const pipemodel* PM, const simcfg* C, envstate* S, pipestate* PS, const real* action,
                             const real off[3], uint32_t gid, real* obs, real* reward, uint8_t* done);
static void FN(reset_env_pipe)(const model* M, const simcfg* C, envstate* S, pipestate* PS, uint32_t gid);

static void FN(reset_env_pipe)(const model* M, const simcfg* C, envstate* S, pipestate* PS, uint32_t gid) {
  FN(reset_env)(M, C, S, gid);   /* arm rest pose, S->ppos = spawn point (peg_in_hole.py:239), FSM variables */
  float u[4];
  pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_RESET, u);
  PS->bpos[0] = S->ppos[0]; PS->bpos[1] = S->ppos[1]; PS->bpos[2] = S->ppos[2];
  PS->bquat[0] = PS->bquat[1] = PS->bquat[2] = RC(0); PS->bquat[3] = RC(1);
  for (int k = 0; k < 3; ++k) { PS->bvel[k] = RC(0); PS->bang[k] = RC(0); }
  PS->grasp_idx = (u[2] < 0.5f) ? 0 : 23;                                    /* random.choice([0, 23]), peg_in_hole.py:266 */
  /* random.sample(range(24), random.randint(5, 24)) joints <- random.uniform(0, pi/3) (peg_in_hole.py:244-245): a uniform
   * k-subset = the k smallest of 24 i.i.d. keys; joint index 0 is the fixed joint (resetJointState on it is a no-op) */
  float v[4], key[24], ang[24];
  pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_PIPE, v);
  int kcount = 5 + (int)(v[0] * 20.0f);
  if (kcount > 24) kcount = 24;
  for (int q = 0; q < 6; ++q) {
    pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_PIPE + 1 + q, &key[4 * q]);
    pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_PIPE + 7 + q, &ang[4 * q]);
  }
  PS->qj[0] = RC(0); PS->qdj[0] = RC(0);
  for (int b = 1; b < PNB; ++b) {
    int rank = 0;
    for (int j = 0; j < 24; ++j) rank += (key[j] < key[b] || (key[j] == key[b] && j < b)) ? 1 : 0;
    PS->qj[b] = (rank < kcount) ? (real)(ang[b] * (float)PIH_PIPE_BEND_MAX) : RC(0);
    PS->qdj[b] = RC(0);
  }
  pipefk K;
  FN(pipe_fk)(PS, &K);
  FN(pipe_grasp_pose)(PS, &K, S->ppos, S->pquat);
}

static int FN(step_env_pipe)(const model* M, const pipemodel* PM, const simcfg* C, envstate* S, pipestate* PS, const real* action,
                             const real off[3], uint32_t gid, real* obs, real* reward, uint8_t* done) {
  if (S->flags & PIH_FLAG_DONE) return 0;
  const int nb = M->nb;
  const real dt = C->dt, idt = RC(1) / dt;
  /* ---- arm: apply_action and unconstrained dynamics, exactly as in step_env */
  fkres F;
  FN(fk)(M, S->q, nb, &F);
  real qstar[NBM], kpc[NBM], mfc[NBM];
  int fsm_done = 0;
  FN(apply_action)(M, C, S, &F, action, off, qstar, kpc, mfc, &fsm_done);
  real tau[NBM];
  for (int i = 0; i < nb; ++i) tau[i] = -(M->damp[i] * S->qd[i]);
  abares AB;
  FN(aba)(M, &F, S->qd, tau, RC(PIH_GRAVITY), &AB);
  real Minv[NBM][NBM];
  for (int j = 0; j < nb; ++j) {
    real col[NBM];
    FN(minv_column)(M, &F, &AB, j, col);
    for (int i = 0; i < nb; ++i) Minv[i][j] = col[i];
  }
  real vs[NBM];
  for (int i = 0; i < nb; ++i) vs[i] = FMA(dt, AB.qdd[i], S->qd[i]);
  /* ---- pipe: kinematics, ABA, unit-impulse responses */
  pipefk K;
  FN(pipe_fk)(PS, &K);
  pipedyn DyL;
  pipedyn* Dy = &DyL;
  FN(pipe_dynamics)(PM, PS, &K, dt, RC(PIH_GRAVITY), Dy);
  /* ---- arm motor / limit rows (as in step_env) */
  real m_dinv[NBM], m_b[NBM], m_lam[NBM], m_max[NBM];
  int l_on[2][NBM];
  real l_b[2][NBM], l_lam[2][NBM];
  for (int j = 0; j < nb; ++j) {
    real qt = qstar[j];
    qt = (qt < M->lo[j]) ? M->lo[j] : ((qt > M->hi[j]) ? M->hi[j] : qt);
    real vstar = kpc[j] * (qt - S->q[j]) * idt;
    m_dinv[j] = RC(1) / Minv[j][j];
    m_b[j] = (vstar - vs[j]) * m_dinv[j];
    m_lam[j] = RC(0);
    m_max[j] = mfc[j] * dt;
    real glo = S->q[j] - M->lo[j], ghi = M->hi[j] - S->q[j];
    real vdn = (-vstar > -vs[j]) ? -vstar : -vs[j];
    vdn = (vdn > RC(0)) ? vdn : RC(0);
    real vup = (vstar > vs[j]) ? vstar : vs[j];
    vup = (vup > RC(0)) ? vup : RC(0);
    l_on[0][j] = glo < FMA(vdn, dt, RC(0.005));
    l_on[1][j] = ghi < FMA(vup, dt, RC(0.005));
    l_lam[0][j] = l_lam[1][j] = RC(0);
    real tlo = (glo > RC(0)) ? -(glo * idt) : -(RC(PIH_ERP) * glo * idt);
    real thi = (ghi > RC(0)) ? -(ghi * idt) : -(RC(PIH_ERP) * ghi * idt);
    l_b[0][j] = (tlo - vs[j]) * m_dinv[j];
    l_b[1][j] = (thi + vs[j]) * m_dinv[j];
  }
  /* ---- pipe joint motors: default velocity motors, target 0, |impulse| <= 1 */
  real pm_dinv[PNB], pm_b[PNB], pm_lam[PNB];
  for (int b = 1; b < PNB; ++b) {
    pm_dinv[b] = RC(1) / Dy->col[5 + b][5 + b];
    pm_b[b] = (RC(0) - Dy->vstar[5 + b]) * pm_dinv[b];
    pm_lam[b] = RC(0);
  }
  /* ---- contacts.  Order = solver row order: every segment's deeper point (segment order), then the segments' second
   *      points, at most PIH_PIPE_SEG_TOTAL in all; then the arm-table contacts.  At most PIH_PIPE_ARMROWS of them may
   *      involve the arm (later ones are dropped). */
  pcontact con[PIH_PIPE_MAXC];
  int ncon = 0, narmc = 0, pairmask = 0;
  if (FN(grasp_active)(C, S->fsm_state)) {
    /* grasp constraint of the scripted task (peg_in_hole.py:99-104; see PIH_PAIR_GRASP): the grasp link's point
     * random_vector is pinned to the end-effector origin - first row set of the step, bilateral, along the env axes */
    v3 pee; m3 Ree;
    FN(ee_pose)(M, &F, &pee, &Ree);
    m3 Rg = FN(quat_to_m3)(S->pquat);
    v3 pc = FN(m3_v_add)(&Rg, FN(v3_make)(RC(0), S->rand_y, RC(0)), FN(v3_make)(S->ppos[0], S->ppos[1], S->ppos[2]));
    pcontact c;
    c.pair = PIH_PAIR_GRASP; c.seg = (PS->grasp_idx == 0) ? 1 : PIH_PIPE_NSEG - 1; c.arm_body = M->ee_body;
    c.p = pee; c.n = FN(v3_sub)(pc, pee); c.dist = RC(-1e30); c.mu = RC(-1);
    con[ncon++] = c;
    narmc += 1;
  }
  const int nweld = ncon;
  {
    pcontact slots[PIH_PIPE_NSEG][PIH_PIPE_SEG_MAXC];
    int cnt[PIH_PIPE_NSEG];
    for (int s = 0; s < PIH_PIPE_NSEG; ++s) {
      FN(pipe_collide_segment)(M, &F, &K, s, slots[s], &cnt[s]);
      if (cnt[s] == 2 && slots[s][1].dist < slots[s][0].dist) { pcontact t = slots[s][0]; slots[s][0] = slots[s][1]; slots[s][1] = t; }
    }
    for (int k = 0; k < PIH_PIPE_SEG_MAXC; ++k)
      for (int s = 0; s < PIH_PIPE_NSEG; ++s) {
        if (k >= cnt[s] || ncon - nweld >= PIH_PIPE_SEG_TOTAL) continue;
        if (slots[s][k].arm_body >= 0) { if (narmc >= PIH_PIPE_ARMROWS) continue; narmc += 1; }
        con[ncon++] = slots[s][k];
      }
  }
  {
    pcontact slot[PIH_PIPE_ARM_MAXC];
    int n = 0;
    for (int q = 0; q < M->nsph; ++q) {
      int body = M->sph_body[q];
      v3 cw = FN(m3_v_add)(&F.Rw[body], M->sph_c[q], F.pw[body]);
      pcontact c;
      c.pair = PIH_PAIR_ARM_TABLE; c.seg = -1; c.arm_body = body; c.n = FN(v3_make)(0, 0, 1);
      c.p = FN(v3_make)(cw.x, cw.y, cw.z - M->sph_r[q]);
      c.dist = cw.z - M->sph_r[q] - RC(PIH_TABLE_Z); c.mu = RC(PIH_MU_DEFAULT) * RC(PIH_MU_DEFAULT);
      FN(pcontact_push)(slot, &n, PIH_PIPE_ARM_MAXC, &c);
    }
    for (int k = 0; k < n; ++k) {
      if (narmc >= PIH_PIPE_ARMROWS) continue;
      narmc += 1;
      con[ncon++] = slot[k];
    }
  }
  /* ---- contact rows */
  prow* rows = (prow*)malloc(sizeof(prow) * (size_t)(ncon > 0 ? ncon : 1));
  for (int k = 0; k < ncon; ++k) {
    const pcontact* c = &con[k];
    prow* r = &rows[k];
    memset(r, 0, sizeof(*r));
    pairmask |= c->pair;
    r->has_pipe = c->seg >= 0; r->has_arm = c->arm_body >= 0; r->mu = c->mu;
    v3 dir[3];
    const int weld = c->pair == PIH_PAIR_GRASP;
    if (weld) { dir[0] = FN(v3_make)(1, 0, 0); dir[1] = FN(v3_make)(0, 1, 0); dir[2] = FN(v3_make)(0, 0, 1); }
    else { dir[0] = c->n; FN(plane_space)(c->n, &dir[1], &dir[2]); }
    for (int d = 0; d < 3; ++d) {
      if (r->has_pipe) {
        const int L = PIPE_SEG_BODY(c->seg);
        v3 rb = FN(v3_sub)(c->p, K.pw[0]);
        v3 ja = FN(m3T_v)(&K.Rw[0], FN(v3_cross)(rb, dir[d]));
        v3 jl = FN(m3T_v)(&K.Rw[0], dir[d]);
        r->Jp[d][0] = ja.x; r->Jp[d][1] = ja.y; r->Jp[d][2] = ja.z; r->Jp[d][3] = jl.x; r->Jp[d][4] = jl.y; r->Jp[d][5] = jl.z;
        for (int b = 1; b <= L; ++b) {
          v3 axw = FN(m3_col)(&K.Rw[b], PIPE_AXIS(b));
          r->Jp[d][5 + b] = FN(v3_dot)(dir[d], FN(v3_cross)(axw, FN(v3_sub)(c->p, K.pw[b])));
        }
        for (int j = 0; j < PND; ++j) {
          real s = RC(0);
          for (int i = 0; i < PND; ++i) s = FMA(Dy->col[j][i], r->Jp[d][i], s);
          r->MJp[d][j] = s;
        }
      }
      if (r->has_arm) {
        real sg = r->has_pipe ? RC(-1) : RC(1);
        for (int j = 0; j < nb; ++j) {
          real v = RC(0);
          if ((M->anc[c->arm_body] >> j) & 1u) {
            v3 z = FN(m3_col)(&F.Rw[j], 2);
            if (M->jtype[j] == 0) v = FN(v3_dot)(dir[d], FN(v3_cross)(z, FN(v3_sub)(c->p, F.pw[j])));
            else v = FN(v3_dot)(dir[d], z);
            v = sg * v;
          }
          r->Ja[d][j] = v;
        }
        for (int i = 0; i < nb; ++i) {
          real s = RC(0);
          for (int j = 0; j < nb; ++j) s = FMA(Minv[i][j], r->Ja[d][j], s);
          r->MJa[d][i] = s;
        }
      }
      /* diag = J M^-1 J^T and the unconstrained relative velocity J v*: butterfly sums over the lanes */
      real pd[32], pv[32];
      for (int i = 0; i < 32; ++i) {
        real a = RC(0), b = RC(0);
        if (r->has_pipe && i < PND) { a = r->Jp[d][i] * r->MJp[d][i]; b = r->Jp[d][i] * Dy->vstar[i]; }
        if (r->has_arm && i < nb) {
          if (r->has_pipe) { a = FMA(r->Ja[d][i], r->MJa[d][i], a); b = FMA(r->Ja[d][i], vs[i], b); }
          else { a = r->Ja[d][i] * r->MJa[d][i]; b = r->Ja[d][i] * vs[i]; }
        }
        pd[i] = a; pv[i] = b;
      }
      real diag = FN(bfly32)(pd), vrel = FN(bfly32)(pv);
      r->dinv[d] = RC(1) / diag;
      r->diag[d] = diag;
      real tv = RC(0);
      if (weld) tv = -(RC(PIH_ERP) * (d == 0 ? c->n.x : (d == 1 ? c->n.y : c->n.z)) * idt);
      else if (d == 0) tv = (c->dist > RC(0)) ? -(c->dist * idt) : -(RC(PIH_ERP) * c->dist * idt);
      r->b[d] = (tv - vrel) * r->dinv[d];
      r->lam[d] = RC(0);
    }
  }
  /* ---- projected Gauss-Seidel over ONE island: arm motors, arm limits, pipe motors, contact normals, contact friction */
  real dva[NBM], dvp[32];
  for (int j = 0; j < NBM; ++j) dva[j] = RC(0);
  for (int j = 0; j < 32; ++j) dvp[j] = RC(0);
  int iters = 0;
  for (int it = 0; it < C->solver_iters; ++it) {
    real res = RC(0);
#define PIH_RES(dl_, diag_) do { real rv_ = (dl_) * (diag_); rv_ = rv_ * rv_; if (rv_ > res) res = rv_; } while (0)
#define PIH_PROJECT(dl_, lam_, lo_, hi_) do { real dlo_ = (lo_) - (lam_), dhi_ = (hi_) - (lam_); \
      dl_ = ((dl_) < dlo_) ? dlo_ : (dl_); dl_ = ((dl_) > dhi_) ? dhi_ : (dl_); lam_ = (lam_) + (dl_); } while (0)
#define PIH_PROJECT_LO(dl_, lam_) do { real dlo_ = -(lam_); dl_ = ((dl_) < dlo_) ? dlo_ : (dl_); lam_ = (lam_) + (dl_); } while (0)
    for (int j = 0; j < nb; ++j) {
      real dl = FMA(-dva[j], m_dinv[j], m_b[j]);
      PIH_PROJECT(dl, m_lam[j], -m_max[j], m_max[j]);
      PIH_RES(dl, Minv[j][j]);
      for (int i = 0; i < nb; ++i) dva[i] = FMA(Minv[i][j], dl, dva[i]);
    }
    for (int j = 0; j < nb; ++j)
      for (int side = 0; side < 2; ++side) {
        if (!l_on[side][j]) continue;
        real sg = side ? RC(-1) : RC(1);
        real dl = FMA(-(sg * dva[j]), m_dinv[j], l_b[side][j]);
        PIH_PROJECT_LO(dl, l_lam[side][j]);
        dl = dl * sg;
        PIH_RES(dl, Minv[j][j]);
        for (int i = 0; i < nb; ++i) dva[i] = FMA(Minv[i][j], dl, dva[i]);
      }
    for (int b = 1; b < PNB; ++b) {
      const int c = 5 + b;
      real dl = FMA(-dvp[c], pm_dinv[b], pm_b[b]);
      PIH_PROJECT(dl, pm_lam[b], RC(-PIH_PIPE_MOTOR_IMPULSE), RC(PIH_PIPE_MOTOR_IMPULSE));
      PIH_RES(dl, Dy->col[c][c]);
      for (int i = 0; i < PND; ++i) dvp[i] = FMA(Dy->col[i][c], dl, dvp[i]);
    }
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < ncon; ++k) {
        prow* r = &rows[k];
        for (int d = (pass ? 1 : 0); d < (pass ? 3 : 1); ++d) {
          real pp[32];
          for (int i = 0; i < 32; ++i) {
            real a = RC(0);
            if (r->has_pipe) a = r->Jp[d][i] * dvp[i];
            if (r->has_arm && i < nb) a = r->has_pipe ? FMA(r->Ja[d][i], dva[i], a) : r->Ja[d][i] * dva[i];
            pp[i] = a;
          }
          real rel = FN(bfly32)(pp);
          real dl = FMA(-rel, r->dinv[d], r->b[d]);
          if (r->mu < RC(0)) {                     /* grasp constraint: bilateral, |impulse| <= maxForce dt */
            const real wl = RC(PIH_GRASP_MAX_FORCE) * dt;
            PIH_PROJECT(dl, r->lam[d], -wl, wl);
          } else if (d == 0) PIH_PROJECT_LO(dl, r->lam[d]);
          else {
            real lim = r->mu * r->lam[0];
            PIH_PROJECT(dl, r->lam[d], -lim, lim);
          }
          PIH_RES(dl, r->diag[d]);
          if (r->has_pipe) for (int i = 0; i < PND; ++i) dvp[i] = FMA(r->MJp[d][i], dl, dvp[i]);
          if (r->has_arm) for (int j = 0; j < nb; ++j) dva[j] = FMA(r->MJa[d][j], dl, dva[j]);
        }
      }
#undef PIH_RES
#undef PIH_PROJECT
#undef PIH_PROJECT_LO
    iters += 1;
    if (res <= C->solver_residual) break;
  }
  pih_dbg_pgs_iters += (unsigned long long)iters;
  pih_dbg_pgs_steps += 1;
  free(rows);
  /* ---- integrate */
  for (int j = 0; j < nb; ++j) {
    real v = vs[j] + dva[j];
    v = (v > RC(PIH_MAX_COORD_VEL)) ? RC(PIH_MAX_COORD_VEL) : ((v < RC(-PIH_MAX_COORD_VEL)) ? RC(-PIH_MAX_COORD_VEL) : v);
    S->qd[j] = v;
    S->q[j] = FMA(dt, S->qd[j], S->q[j]);
  }
  for (int b = 1; b < PNB; ++b) {
    real v = Dy->vstar[5 + b] + dvp[5 + b];
    v = (v > RC(PIH_MAX_COORD_VEL)) ? RC(PIH_MAX_COORD_VEL) : ((v < RC(-PIH_MAX_COORD_VEL)) ? RC(-PIH_MAX_COORD_VEL) : v);
    PS->qdj[b] = v;
    PS->qj[b] = FMA(dt, v, PS->qj[b]);
  }
  {
    v3 wb = FN(v3_make)(Dy->vstar[0] + dvp[0], Dy->vstar[1] + dvp[1], Dy->vstar[2] + dvp[2]);
    v3 vb = FN(v3_make)(Dy->vstar[3] + dvp[3], Dy->vstar[4] + dvp[4], Dy->vstar[5] + dvp[5]);
    v3 fw = FN(clamp_norm)(FN(m3_v)(&K.Rw[0], wb));
    v3 fv = FN(clamp_norm)(FN(m3_v)(&K.Rw[0], vb));
    PS->bvel[0] = fv.x; PS->bvel[1] = fv.y; PS->bvel[2] = fv.z;
    PS->bang[0] = fw.x; PS->bang[1] = fw.y; PS->bang[2] = fw.z;
    PS->bpos[0] = FMA(dt, fv.x, PS->bpos[0]); PS->bpos[1] = FMA(dt, fv.y, PS->bpos[1]); PS->bpos[2] = FMA(dt, fv.z, PS->bpos[2]);
    real w2 = FN(v3_dot)(fw, fw);
    real wn = SQRT(w2);
    real ha = RC(0.5) * dt * wn;
    real sh, ch, kk;
    FN(det_sincos)(ha, &sh, &ch);
    if (wn > RC(1e-6)) kk = sh / wn; else kk = RC(0.5) * dt;
    real dq[4] = {fw.x * kk, fw.y * kk, fw.z * kk, ch}, nq[4];
    FN(quat_mul)(dq, PS->bquat, nq);
    real n2 = FMA(nq[3], nq[3], FMA(nq[2], nq[2], FMA(nq[1], nq[1], nq[0] * nq[0])));
    real in = RC(1) / SQRT(n2);
    for (int i = 0; i < 4; ++i) PS->bquat[i] = nq[i] * in;
  }
  S->step += 1;
  /* ---- grasp link pose of the new state (what the next step's state machine and the success test read) */
  {
    pipefk K2;
    FN(pipe_fk)(PS, &K2);
    FN(pipe_grasp_pose)(PS, &K2, S->ppos, S->pquat);
  }
  /* ---- fault guard */
  {
    int finite = 1;
    for (int j = 0; j < nb; ++j) finite = finite && isfinite(S->q[j]) && isfinite(S->qd[j]);
    for (int k = 0; k < 3; ++k) finite = finite && isfinite(PS->bpos[k]) && isfinite(PS->bvel[k]) && isfinite(PS->bang[k]) && isfinite(S->ppos[k]);
    for (int k = 0; k < 4; ++k) finite = finite && isfinite(PS->bquat[k]) && isfinite(S->pquat[k]);
    for (int b = 1; b < PNB; ++b) finite = finite && isfinite(PS->qj[b]) && isfinite(PS->qdj[b]);
    if (!finite) {
      S->episode += 1;
      FN(reset_env_pipe)(M, C, S, PS, gid);
      S->flags = PIH_FLAG_FAULT | (C->auto_reset ? 0 : PIH_FLAG_DONE);
      *reward = RC(0);
      *done = 1;
      FN(observe)(M, C, S, off, obs);
      return ncon;
    }
  }
  /* ---- get_info: success on the grasp link (peg_in_hole.py:114-117) */
  v3 dd = FN(v3_make)(S->ppos[0] - RC(PIH_HOLE_X), S->ppos[1] - RC(PIH_HOLE_Y), S->ppos[2] - RC(PIH_HOLE_Z));
  int success = FN(v3_dot)(dd, dd) < RC(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
  real rew = success ? RC(1) : RC(0);
  int dn = success || (C->max_steps > 0 && S->step >= C->max_steps);
  if (C->task == PIH_TASK_SCRIPTED) dn = fsm_done || (C->max_steps > 0 && S->step >= C->max_steps);
  S->flags = (dn ? PIH_FLAG_DONE : 0) | (success ? PIH_FLAG_SUCCESS : 0) | (pairmask << PIH_PAIR_SHIFT);
  *reward = rew;
  *done = (uint8_t)dn;
  if (dn && C->auto_reset) {
    int keep = S->flags & ~PIH_FLAG_DONE;
    S->episode += 1;
    FN(reset_env_pipe)(M, C, S, PS, gid);
    S->flags = (keep & ~0xff) | (success ? PIH_FLAG_PREV_SUCCESS : 0);
  }
  FN(observe)(M, C, S, off, obs);
  return ncon;
}

#undef PNB
#undef PND
#undef pipestate
#undef pipemodel
#undef pipefk
#undef ldl6
#undef pipedyn
#undef pcontact
#undef prow
