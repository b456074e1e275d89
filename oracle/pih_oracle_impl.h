/* CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path (peg_in_hole_gym_b200/);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY UNPINNED vs PyBullet: the reference's arithmetic lives in the un-vendored, unpinned `pybullet`
 * C++ extension (requirements.txt:2, setup.py:19), which is not installed here and has no golden vectors in
 * the reference.  This file is a plain-C restatement of the *published algorithms*
 * behind the reference's call sites for one env.step, anchored on those call sites:
 *   BaseEnv.step                  peg_in_hole_gym/envs/base_env.py:60-75
 *   panda_execute / ur_execute    peg_in_hole_gym/envs/utils.py:60-82
 *   vel_constraint                peg_in_hole_gym/envs/utils.py:85-95
 *   getLinkState (FK)             call sites utils.py:62,80
 *   calculateInverseKinematics    call sites utils.py:67,79          (BussIK damped least squares [3P])
 *   setJointMotorControlArray     call sites utils.py:68,82          (btMultiBodyJointMotor rows [3P])
 *   stepSimulation                base_env.py:64                     (ABA + PGS(50) + semi-implicit Euler [3P])
 *   success / reward              peg_in_hole_gym/envs/peg_in_hole.py:114-117
 *   reset distributions           peg_in_hole_gym/envs/peg_in_hole.py:230-267
 * It is pinned against: the UR5 FK known answers derived from the in-tree URDF,
 * Random123 Philox4x32-10 KATs, and closed-form physics checks (tests/test_oracle_*.py).
 *
 * Included twice by pih_oracle.c: REAL=float (bit-reproducible mode: every multiply-add that is fused is an
 * explicit FMA(); compile with -ffp-contract=off) and REAL=double (drift reference).  The CUDA kernels follow the
 * same operation order so that the float build is bit-comparable with the GPU. */

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#ifndef REAL
#error "define REAL, SUFFIX and REAL_IS_DOUBLE before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef REAL FN(real);
#define real FN(real)

#if REAL_IS_DOUBLE
#define FMA(a, b, c) fma((a), (b), (c))
#define SQRT(a) sqrt(a)
#define RC(x) ((real)(x))
#else
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(a) sqrtf(a)
#define RC(x) ((real)(x))
#endif
#define ABS(a) ((a) < RC(0) ? -(a) : (a))

/* ------------------------------------------------------------------ deterministic elementary functions */
#ifndef PIH_ORACLE_SINCOSF
#define PIH_ORACLE_SINCOSF
static inline void pih_sincosf_det(float x, float* s, float* c) {
  float kf = rintf(x * 0.636619772367581343f);
  int k = (int)kf;
  float r = fmaf(kf, -1.57079637050628662109375f, x);
  r = fmaf(kf, 4.37113900018624283e-8f, r);
  float z = r * r;
  float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
  float sn = fmaf(ps * z, r, r);
  float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
  float cs = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  switch (k & 3) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}
#endif
/* float mode: Cody-Waite reduction + cephes-style minimax polynomials written with explicit FMA so the CUDA
 * kernel reproduces them bit for bit; double mode: libm (truth). */
static inline void FN(det_sincos)(real x, real* s, real* c) {
#if REAL_IS_DOUBLE
  *s = sin(x);
  *c = cos(x);
#else
  pih_sincosf_det(x, s, c);
#endif
}

/* atan2(y, x) for y >= 0, x >= 0 (first quadrant only - all that the IK orientation error needs) */
static inline real FN(det_atan2_q1)(real y, real x) {
#if REAL_IS_DOUBLE
  return atan2(y, x);
#else
  int swap = y > x;
  float num = swap ? x : y, den = swap ? y : x;
  if (den == 0.0f) return 0.0f;
  float t = num / den; /* in [0,1] */
  float off = 0.0f;
  if (t > 0.4142135679721832f) { /* tan(pi/8) */
    off = 0.785398185253143310546875f;
    t = (t - 1.0f) / (t + 1.0f);
  }
  float z = t * t;
  float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
  float a = off + fmaf(p * z, t, t);
  return swap ? (1.57079637050628662109375f - a) : a;
#endif
}

/* full-range atan2 from the first-quadrant kernel */
static inline real FN(det_atan2)(real y, real x) {
#if REAL_IS_DOUBLE
  return atan2(y, x);
#else
  float a = FN(det_atan2_q1)(y < 0.0f ? -y : y, x < 0.0f ? -x : x);
  if (x < 0.0f) a = 3.1415927410125732421875f - a;
  return (y < 0.0f) ? -a : a;
#endif
}

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) */
#ifndef PIH_ORACLE_PHILOX
#define PIH_ORACLE_PHILOX
static inline void pih_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* counter = (global env id, index, stream, 0), key = 64-bit seed; uniform in [0,1): top 24 bits */
static inline void pih_rng4(uint64_t seed, uint32_t env, uint32_t index, uint32_t stream, float u[4]) {
  uint32_t ctr[4] = {env, index, stream, 0u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  pih_philox4x32_10(ctr, key, o);
  for (int i = 0; i < 4; ++i) u[i] = (float)(o[i] >> 8) * 5.9604644775390625e-8f;
}
#endif

/* ------------------------------------------------------------------ small linear algebra (canonical op order) */
typedef struct { real x, y, z; } FN(v3);
#define v3 FN(v3)
typedef struct { real m[3][3]; } FN(m3);
#define m3 FN(m3)

static inline v3 FN(v3_make)(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static inline v3 FN(v3_add)(v3 a, v3 b) { return FN(v3_make)(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 FN(v3_sub)(v3 a, v3 b) { return FN(v3_make)(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 FN(v3_scale)(v3 a, real s) { return FN(v3_make)(a.x * s, a.y * s, a.z * s); }
static inline real FN(v3_dot)(v3 a, v3 b) { return FMA(a.z, b.z, FMA(a.y, b.y, a.x * b.x)); }
static inline v3 FN(v3_cross)(v3 a, v3 b) {
  return FN(v3_make)(FMA(a.y, b.z, -(a.z * b.y)), FMA(a.z, b.x, -(a.x * b.z)), FMA(a.x, b.y, -(a.y * b.x)));
}
/* a + s*b */
static inline v3 FN(v3_axpy)(real s, v3 b, v3 a) { return FN(v3_make)(FMA(s, b.x, a.x), FMA(s, b.y, a.y), FMA(s, b.z, a.z)); }
static inline v3 FN(m3_v)(const m3* M, v3 v) {
  return FN(v3_make)(FMA(M->m[0][2], v.z, FMA(M->m[0][1], v.y, M->m[0][0] * v.x)),
                     FMA(M->m[1][2], v.z, FMA(M->m[1][1], v.y, M->m[1][0] * v.x)),
                     FMA(M->m[2][2], v.z, FMA(M->m[2][1], v.y, M->m[2][0] * v.x)));
}
static inline v3 FN(m3T_v)(const m3* M, v3 v) {
  return FN(v3_make)(FMA(M->m[2][0], v.z, FMA(M->m[1][0], v.y, M->m[0][0] * v.x)),
                     FMA(M->m[2][1], v.z, FMA(M->m[1][1], v.y, M->m[0][1] * v.x)),
                     FMA(M->m[2][2], v.z, FMA(M->m[1][2], v.y, M->m[0][2] * v.x)));
}
/* a + M v */
static inline v3 FN(m3_v_add)(const m3* M, v3 v, v3 a) {
  return FN(v3_make)(FMA(M->m[0][2], v.z, FMA(M->m[0][1], v.y, FMA(M->m[0][0], v.x, a.x))),
                     FMA(M->m[1][2], v.z, FMA(M->m[1][1], v.y, FMA(M->m[1][0], v.x, a.y))),
                     FMA(M->m[2][2], v.z, FMA(M->m[2][1], v.y, FMA(M->m[2][0], v.x, a.z))));
}
static inline m3 FN(m3_mul)(const m3* A, const m3* B) {
  m3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = FMA(A->m[i][2], B->m[2][j], FMA(A->m[i][1], B->m[1][j], A->m[i][0] * B->m[0][j]));
  return R;
}
static inline m3 FN(m3_mulT)(const m3* A, const m3* B) { /* A B^T */
  m3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = FMA(A->m[i][2], B->m[j][2], FMA(A->m[i][1], B->m[j][1], A->m[i][0] * B->m[j][0]));
  return R;
}
static inline v3 FN(m3_col)(const m3* M, int c) { return FN(v3_make)(M->m[0][c], M->m[1][c], M->m[2][c]); }
static inline v3 FN(m3_row)(const m3* M, int r) { return FN(v3_make)(M->m[r][0], M->m[r][1], M->m[r][2]); }

/* quaternion (x,y,z,w) -> rotation matrix */
static inline m3 FN(quat_to_m3)(const real q[4]) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  real xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  m3 R;
  R.m[0][0] = FMA(RC(-2), yy + zz, RC(1)); R.m[0][1] = RC(2) * (xy - wz);           R.m[0][2] = RC(2) * (xz + wy);
  R.m[1][0] = RC(2) * (xy + wz);           R.m[1][1] = FMA(RC(-2), xx + zz, RC(1)); R.m[1][2] = RC(2) * (yz - wx);
  R.m[2][0] = RC(2) * (xz - wy);           R.m[2][1] = RC(2) * (yz + wx);           R.m[2][2] = FMA(RC(-2), xx + yy, RC(1));
  return R;
}
/* rotation matrix -> quaternion (x,y,z,w), Shepperd's method */
static inline void FN(m3_to_quat)(const m3* R, real q[4]) {
  real m00 = R->m[0][0], m11 = R->m[1][1], m22 = R->m[2][2];
  real tr = m00 + m11 + m22;
  if (tr > RC(0)) {
    real s = SQRT(tr + RC(1)) * RC(2);
    real is = RC(1) / s;
    q[3] = RC(0.25) * s;
    q[0] = (R->m[2][1] - R->m[1][2]) * is;
    q[1] = (R->m[0][2] - R->m[2][0]) * is;
    q[2] = (R->m[1][0] - R->m[0][1]) * is;
  } else if (m00 > m11 && m00 > m22) {
    real s = SQRT(RC(1) + m00 - m11 - m22) * RC(2);
    real is = RC(1) / s;
    q[3] = (R->m[2][1] - R->m[1][2]) * is;
    q[0] = RC(0.25) * s;
    q[1] = (R->m[0][1] + R->m[1][0]) * is;
    q[2] = (R->m[0][2] + R->m[2][0]) * is;
  } else if (m11 > m22) {
    real s = SQRT(RC(1) + m11 - m00 - m22) * RC(2);
    real is = RC(1) / s;
    q[3] = (R->m[0][2] - R->m[2][0]) * is;
    q[0] = (R->m[0][1] + R->m[1][0]) * is;
    q[1] = RC(0.25) * s;
    q[2] = (R->m[1][2] + R->m[2][1]) * is;
  } else {
    real s = SQRT(RC(1) + m22 - m00 - m11) * RC(2);
    real is = RC(1) / s;
    q[3] = (R->m[1][0] - R->m[0][1]) * is;
    q[0] = (R->m[0][2] + R->m[2][0]) * is;
    q[1] = (R->m[1][2] + R->m[2][1]) * is;
    q[2] = RC(0.25) * s;
  }
}
/* a (x) b, Hamilton product, (x,y,z,w) */
static inline void FN(quat_mul)(const real a[4], const real b[4], real r[4]) {
  real ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[0] = FMA(aw, bx, FMA(ax, bw, FMA(ay, bz, -(az * by))));
  r[1] = FMA(aw, by, FMA(ay, bw, FMA(az, bx, -(ax * bz))));
  r[2] = FMA(aw, bz, FMA(az, bw, FMA(ax, by, -(ay * bx))));
  r[3] = FMA(aw, bw, -FMA(az, bz, FMA(ay, by, ax * bx)));
}
/* getQuaternionFromEuler([r,p,y]) = qz(y) qy(p) qx(r)   */
static inline void FN(euler_to_quat)(real roll, real pitch, real yaw, real q[4]) {
  real sr, cr, sp, cp, sy, cy;
  FN(det_sincos)(roll * RC(0.5), &sr, &cr);
  FN(det_sincos)(pitch * RC(0.5), &sp, &cp);
  FN(det_sincos)(yaw * RC(0.5), &sy, &cy);
  q[0] = FMA(sr * cp, cy, -(cr * sp * sy));
  q[1] = FMA(cr * sp, cy, sr * cp * sy);
  q[2] = FMA(cr * cp, sy, -(sr * sp * cy));
  q[3] = FMA(cr * cp, cy, sr * sp * sy);
}

/* ------------------------------------------------------------------ model */
#define NBM PIH_NB_MAX
typedef struct {
  int nb, narm, ee_body, nsph, nbox;
  int parent[NBM], jtype[NBM];
  unsigned anc[NBM]; /* bit j set: body j is an ancestor of (or is) body i */
  m3 R0[NBM];
  v3 t0[NBM], h[NBM];
  m3 Irot[NBM];
  real mass[NBM], lo[NBM], hi[NBM], damp[NBM], maxf[NBM], kp[NBM], rest[NBM];
  m3 ee_R;
  v3 ee_t;
  int sph_body[PIH_MAX_SPH];
  v3 sph_c[PIH_MAX_SPH];
  real sph_r[PIH_MAX_SPH];
  int box_body[PIH_MAX_BOX];
  v3 box_c[PIH_MAX_BOX], box_h[PIH_MAX_BOX];
  m3 box_R[PIH_MAX_BOX];
  /* free body ("peg" or fly object) */
  int fb_shape; /* 0 cylinder (axis local y), 1 box */
  real fb_mass, fb_mu;
  v3 fb_half;   /* cylinder: (r, half length, r); box: half extents */
  v3 fb_iinv;   /* inverse body-frame principal inertia (Bullet box-of-AABB, margin 1 mm) */
  v3 fb_idiag;
} FN(model);
#define model FN(model)

static void FN(m3_from)(m3* M, const double* a) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M->m[i][j] = RC(a[3 * i + j]);
}
static v3 FN(v3_from)(const double* a) { return FN(v3_make)(RC(a[0]), RC(a[1]), RC(a[2])); }

#define LOAD_ARM(M, P)                                                                                   \
  do {                                                                                                   \
    (M)->nb = PIH_##P##_NB; (M)->narm = PIH_##P##_NARM; (M)->ee_body = PIH_##P##_EE_BODY;                 \
    (M)->nsph = PIH_##P##_NSPH; (M)->nbox = PIH_##P##_NBOX;                                              \
    for (int i = 0; i < PIH_##P##_NB; ++i) {                                                             \
      (M)->parent[i] = PIH_##P##_PARENT[i]; (M)->jtype[i] = PIH_##P##_JTYPE[i];                           \
      FN(m3_from)(&(M)->R0[i], PIH_##P##_R0[i]); (M)->t0[i] = FN(v3_from)(PIH_##P##_T0[i]);               \
      (M)->h[i] = FN(v3_from)(PIH_##P##_H[i]); FN(m3_from)(&(M)->Irot[i], PIH_##P##_IROT[i]);            \
      (M)->mass[i] = RC(PIH_##P##_MASS[i]); (M)->lo[i] = RC(PIH_##P##_LO[i]); (M)->hi[i] = RC(PIH_##P##_HI[i]); \
      (M)->damp[i] = RC(PIH_##P##_DAMP[i]); (M)->maxf[i] = RC(PIH_##P##_MAXF[i]);                         \
      (M)->kp[i] = RC(PIH_##P##_KP[i]); (M)->rest[i] = RC(PIH_##P##_REST[i]);                             \
    }                                                                                                    \
    FN(m3_from)(&(M)->ee_R, PIH_##P##_EE_R); (M)->ee_t = FN(v3_from)(PIH_##P##_EE_T);                     \
    for (int i = 0; i < PIH_MAX_SPH; ++i) {                                                              \
      (M)->sph_body[i] = PIH_##P##_SPH_BODY[i]; (M)->sph_c[i] = FN(v3_from)(PIH_##P##_SPH_C[i]);          \
      (M)->sph_r[i] = RC(PIH_##P##_SPH_R[i]);                                                            \
    }                                                                                                    \
    for (int i = 0; i < PIH_MAX_BOX; ++i) {                                                              \
      (M)->box_body[i] = PIH_##P##_BOX_BODY[i]; (M)->box_c[i] = FN(v3_from)(PIH_##P##_BOX_C[i]);          \
      (M)->box_h[i] = FN(v3_from)(PIH_##P##_BOX_H[i]); FN(m3_from)(&(M)->box_R[i], PIH_##P##_BOX_R[i]);   \
    }                                                                                                    \
  } while (0)

static void FN(model_init)(model* M, int arm, int task, int fly_object) {
  memset(M, 0, sizeof(*M));
  if (arm == PIH_ARM_PANDA) LOAD_ARM(M, PANDA); else LOAD_ARM(M, UR5);
  for (int i = 0; i < M->nb; ++i) {
    unsigned a = 1u << i;
    for (int p = M->parent[i]; p >= 0; p = M->parent[p]) a |= 1u << p;
    M->anc[i] = a;
  }
  double hx, hy, hz, mass, mu;
  if (task == PIH_TASK_RANDOM_FLY) {
    M->fb_shape = 1;
    if (fly_object == 0) { hx = PIH_BANANA_HX; hy = PIH_BANANA_HY; hz = PIH_BANANA_HZ; }
    else { hx = PIH_AMICELLI_HX; hy = PIH_AMICELLI_HY; hz = PIH_AMICELLI_HZ; }
    mass = PIH_FLY_MASS; mu = PIH_FLY_MU;
  } else {
    M->fb_shape = 0; hx = PIH_PEG_R; hy = PIH_PEG_HL; hz = PIH_PEG_R; mass = PIH_PEG_MASS; mu = PIH_PEG_MU;
  }
  /* Bullet default URDF loading: box inertia of the collision AABB + 1 mm margin */
  double lx = 2 * (hx + 0.001), ly = 2 * (hy + 0.001), lz = 2 * (hz + 0.001);
  double ix = mass / 12.0 * (ly * ly + lz * lz), iy = mass / 12.0 * (lx * lx + lz * lz), iz = mass / 12.0 * (lx * lx + ly * ly);
  M->fb_mass = RC(mass); M->fb_mu = RC(mu);
  M->fb_half = FN(v3_make)(RC(hx), RC(hy), RC(hz));
  M->fb_idiag = FN(v3_make)(RC(ix), RC(iy), RC(iz));
  M->fb_iinv = FN(v3_make)(RC(1.0 / ix), RC(1.0 / iy), RC(1.0 / iz));
}

/* ------------------------------------------------------------------ per-env state */
typedef struct {
  real q[NBM], qd[NBM];
  real ppos[3], pquat[4], pvel[3], pang[3];
  int32_t step, episode, flags;
  int32_t fsm_state, fsm_t;   /* scripted task */
  real rand_y, mtgt[7];
  real grasp_ang;             /* captured on entering FSM state 2 (peg_in_hole.py:65-72) */
  real latch[16];             /* scene of the grasp image, latched at the same moment: q[9], ppos[3], pquat[4] (PIH_S_LATCH_*) */
} FN(envstate);
#define envstate FN(envstate)

typedef struct {
  int task, arm, fly_object;
  real dt, solver_residual;
  int solver_iters, ik_iters, max_steps, auto_reset, action_frame_env;
  uint64_t seed;
} FN(simcfg);
#define simcfg FN(simcfg)

/* ------------------------------------------------------------------ forward kinematics (getLinkState) */
typedef struct {
  m3 Rl[NBM];  /* child->parent rotation  R0 * Rz(q)            */
  v3 tl[NBM];  /* child origin in parent coordinates            */
  m3 Rw[NBM];  /* body->env rotation                            */
  v3 pw[NBM];  /* body origin in env frame                      */
} FN(fkres);
#define fkres FN(fkres)

static inline void FN(joint_local)(const model* M, int i, real q, m3* Rl, v3* tl) {
  const m3* R0 = &M->R0[i];
  if (M->jtype[i] == 0) {
    real s, c;
    FN(det_sincos)(q, &s, &c);
    for (int r = 0; r < 3; ++r) {
      Rl->m[r][0] = FMA(R0->m[r][1], s, R0->m[r][0] * c);
      Rl->m[r][1] = FMA(R0->m[r][1], c, -(R0->m[r][0] * s));
      Rl->m[r][2] = R0->m[r][2];
    }
    *tl = M->t0[i];
  } else {
    *Rl = *R0;
    *tl = FN(v3_make)(FMA(R0->m[0][2], q, M->t0[i].x), FMA(R0->m[1][2], q, M->t0[i].y), FMA(R0->m[2][2], q, M->t0[i].z));
  }
}

static void FN(fk)(const model* M, const real* q, int nbodies, fkres* F) {
  for (int i = 0; i < nbodies; ++i) {
    FN(joint_local)(M, i, q[i], &F->Rl[i], &F->tl[i]);
    int p = M->parent[i];
    if (p < 0) { F->Rw[i] = F->Rl[i]; F->pw[i] = F->tl[i]; }
    else { F->Rw[i] = FN(m3_mul)(&F->Rw[p], &F->Rl[i]); F->pw[i] = FN(m3_v_add)(&F->Rw[p], F->tl[i], F->pw[p]); }
  }
}
static inline void FN(ee_pose)(const model* M, const fkres* F, v3* pee, m3* Ree) {
  int e = M->ee_body;
  *Ree = FN(m3_mul)(&F->Rw[e], &M->ee_R);
  *pee = FN(m3_v_add)(&F->Rw[e], M->ee_t, F->pw[e]);
}

/* ------------------------------------------------------------------ damped-least-squares IK
 * calculateInverseKinematics(body, eeLink, pos, orn) with PyBullet defaults [3P]: seeded at the current joint
 * state, <= ik_iters iterations, dtheta = J^T (J J^T + lambda^2 I)^-1 e, per-iteration clamp 45 deg,
 * early exit when |e| < 1e-4.  Only the narm chain joints get Jacobian columns (finger columns are zero). */
static void FN(ik_dls)(const model* M, const simcfg* C, const fkres* F0, const real* q0, v3 tpos, const real tq[4], real* qout) {
  const int n = M->narm;
  fkres F = *F0;
  real q[NBM];
  for (int i = 0; i < M->nb; ++i) q[i] = q0[i];
  for (int it = 0; it < C->ik_iters; ++it) {
    if (it > 0) FN(fk)(M, q, n, &F);
    v3 pee; m3 Ree;
    FN(ee_pose)(M, &F, &pee, &Ree);
    real e[6];
    v3 ep = FN(v3_sub)(tpos, pee);
    real qe[4], qc[4], qd[4];
    FN(m3_to_quat)(&Ree, qe);
    qc[0] = -qe[0]; qc[1] = -qe[1]; qc[2] = -qe[2]; qc[3] = qe[3];
    FN(quat_mul)(tq, qc, qd);
    if (qd[3] < RC(0)) { qd[0] = -qd[0]; qd[1] = -qd[1]; qd[2] = -qd[2]; qd[3] = -qd[3]; }
    real sn2 = FMA(qd[2], qd[2], FMA(qd[1], qd[1], qd[0] * qd[0]));
    real sn = SQRT(sn2);
    real k = RC(2);
    if (sn > RC(1e-9)) k = RC(2) * FN(det_atan2_q1)(sn, qd[3]) / sn;
    e[0] = ep.x; e[1] = ep.y; e[2] = ep.z; e[3] = k * qd[0]; e[4] = k * qd[1]; e[5] = k * qd[2];
    real en2 = RC(0);
    for (int a = 0; a < 6; ++a) en2 = FMA(e[a], e[a], en2);
    if (en2 < RC(PIH_IK_RESIDUAL * PIH_IK_RESIDUAL)) break;
    /* geometric Jacobian columns: revolute joint j about its body z axis through pw[j] */
    real J[6][NBM];
    for (int j = 0; j < n; ++j) {
      v3 z = FN(m3_col)(&F.Rw[j], 2);
      v3 d = FN(v3_sub)(pee, F.pw[j]);
      v3 jv = FN(v3_cross)(z, d);
      J[0][j] = jv.x; J[1][j] = jv.y; J[2][j] = jv.z; J[3][j] = z.x; J[4][j] = z.y; J[5][j] = z.z;
    }
    /* A = J J^T + lambda^2 I (upper triangle computed, mirrored); A = L D L^T (unit L; W = L D), then
     * L z = e, y = z / d, L^T x = y.  (Square-root-free form of the 6x6 SPD solve.) */
    real A[6][6], L[6][6], W[6][6], invd[6], y[6];
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) {
        real s = (a == b) ? RC(PIH_IK_LAMBDA * PIH_IK_LAMBDA) : RC(0);
        for (int j = 0; j < n; ++j) s = FMA(J[a][j], J[b][j], s);
        A[a][b] = s; A[b][a] = s;
      }
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < i; ++j) {
        real s = A[i][j];
        for (int kk = 0; kk < j; ++kk) s = FMA(-W[i][kk], L[j][kk], s);
        W[i][j] = s;
        L[i][j] = s * invd[j];
      }
      real d = A[i][i];
      for (int kk = 0; kk < i; ++kk) d = FMA(-W[i][kk], L[i][kk], d);
      invd[i] = RC(1) / d;
    }
    for (int i = 0; i < 6; ++i) {
      real s = e[i];
      for (int kk = 0; kk < i; ++kk) s = FMA(-L[i][kk], y[kk], s);
      y[i] = s;
    }
    for (int i = 5; i >= 0; --i) {
      real s = y[i] * invd[i];
      for (int kk = i + 1; kk < 6; ++kk) s = FMA(-L[kk][i], y[kk], s);
      y[i] = s;
    }
    real dth[NBM], mx = RC(0);
    for (int j = 0; j < n; ++j) {
      real s = J[0][j] * y[0];
      for (int a = 1; a < 6; ++a) s = FMA(J[a][j], y[a], s);
      dth[j] = s;
      real as = ABS(s);
      if (as > mx) mx = as;
    }
    if (mx > RC(PIH_IK_MAXANGLE)) {
      real sc = RC(PIH_IK_MAXANGLE) / mx;
      for (int j = 0; j < n; ++j) dth[j] = dth[j] * sc;
    }
    for (int j = 0; j < n; ++j) q[j] = q[j] + dth[j];
  }
  for (int j = 0; j < n; ++j) qout[j] = q[j];
}

/* ------------------------------------------------------------------ articulated-body algorithm
 * Featherstone ABA in body coordinates; joint motion subspace is e_z (angular for revolute, linear for
 * prismatic).  Spatial inertia [[A,B],[B^T,C]] on (omega; v).  Gravity enters as a base acceleration +g z.
 * Restates stepSimulation's computeAccelerationsArticulatedBodyAlgorithmMultiDof [3P]. */
typedef struct { v3 a, l; } FN(sv); /* spatial vector: angular, linear */
#define sv FN(sv)
typedef struct {
  sv U[NBM];      /* I^A S                               */
  sv UD[NBM];     /* U / D                               */
  real Dinv[NBM];
  real u[NBM];
  real qdd[NBM];
} FN(abares);
#define abares FN(abares)

/* parent-frame force from child-frame force: f_p = Rl f ; n_p = Rl n + r x f_p */
static inline sv FN(xf_force_to_parent)(const m3* Rl, v3 r, sv f) {
  sv o;
  o.l = FN(m3_v)(Rl, f.l);
  v3 n = FN(m3_v)(Rl, f.a);
  o.a = FN(v3_add)(n, FN(v3_cross)(r, o.l));
  return o;
}
/* child-frame motion from parent-frame motion: w' = Rl^T w ; v' = Rl^T (v + w x r) */
static inline sv FN(xf_motion_to_child)(const m3* Rl, v3 r, sv m) {
  sv o;
  o.a = FN(m3T_v)(Rl, m.a);
  o.l = FN(m3T_v)(Rl, FN(v3_add)(m.l, FN(v3_cross)(m.a, r)));
  return o;
}

static void FN(aba)(const model* M, const fkres* F, const real* qd, const real* tau, real g, abares* R) {
  const int nb = M->nb;
  sv vel[NBM], cb[NBM], pA[NBM];
  m3 IA[NBM], IB[NBM], IC[NBM];
  for (int i = 0; i < nb; ++i) {
    int p = M->parent[i];
    sv vp;
    if (p < 0) { vp.a = FN(v3_make)(0, 0, 0); vp.l = vp.a; vp = FN(xf_motion_to_child)(&F->Rl[i], F->tl[i], vp); }
    else vp = FN(xf_motion_to_child)(&F->Rl[i], F->tl[i], vel[p]);
    sv v = vp;
    real w = qd[i];
    if (M->jtype[i] == 0) {
      v.a.z = vp.a.z + w;
      cb[i].a = FN(v3_make)(vp.a.y * w, -(vp.a.x * w), RC(0));
      cb[i].l = FN(v3_make)(vp.l.y * w, -(vp.l.x * w), RC(0));
    } else {
      v.l.z = vp.l.z + w;
      cb[i].a = FN(v3_make)(0, 0, 0);
      cb[i].l = FN(v3_make)(vp.a.y * w, -(vp.a.x * w), RC(0));
    }
    vel[i] = v;
    /* momentum L = I v ; bias force p = v x* L */
    v3 La = FN(v3_add)(FN(m3_v)(&M->Irot[i], v.a), FN(v3_cross)(M->h[i], v.l));
    v3 hw = FN(v3_cross)(M->h[i], v.a);
    v3 Ll = FN(v3_make)(FMA(M->mass[i], v.l.x, -hw.x), FMA(M->mass[i], v.l.y, -hw.y), FMA(M->mass[i], v.l.z, -hw.z));
    pA[i].a = FN(v3_add)(FN(v3_cross)(v.a, La), FN(v3_cross)(v.l, Ll));
    pA[i].l = FN(v3_cross)(v.a, Ll);
    IA[i] = M->Irot[i];
    v3 h = M->h[i];
    m3 B = {{{RC(0), -h.z, h.y}, {h.z, RC(0), -h.x}, {-h.y, h.x, RC(0)}}};
    IB[i] = B;
    m3 Cm = {{{M->mass[i], RC(0), RC(0)}, {RC(0), M->mass[i], RC(0)}, {RC(0), RC(0), M->mass[i]}}};
    IC[i] = Cm;
  }
  for (int i = nb - 1; i >= 0; --i) {
    sv U;
    real D, u;
    if (M->jtype[i] == 0) {
      U.a = FN(m3_col)(&IA[i], 2); U.l = FN(m3_row)(&IB[i], 2); D = IA[i].m[2][2]; u = tau[i] - pA[i].a.z;
    } else {
      U.a = FN(m3_col)(&IB[i], 2); U.l = FN(m3_col)(&IC[i], 2); D = IC[i].m[2][2]; u = tau[i] - pA[i].l.z;
    }
    real Dinv = RC(1) / D;
    sv UD;
    UD.a = FN(v3_scale)(U.a, Dinv);
    UD.l = FN(v3_scale)(U.l, Dinv);
    R->U[i] = U; R->UD[i] = UD; R->Dinv[i] = Dinv; R->u[i] = u;
    int p = M->parent[i];
    if (p < 0) continue;
    real uda[3] = {UD.a.x, UD.a.y, UD.a.z}, udl[3] = {UD.l.x, UD.l.y, UD.l.z};
    real ua[3] = {U.a.x, U.a.y, U.a.z}, ul[3] = {U.l.x, U.l.y, U.l.z};
    m3 A1, B1, C1;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        A1.m[r][c] = FMA(-uda[r], ua[c], IA[i].m[r][c]);
        B1.m[r][c] = FMA(-uda[r], ul[c], IB[i].m[r][c]);
        C1.m[r][c] = FMA(-udl[r], ul[c], IC[i].m[r][c]);
      }
    /* pa = p + Ia c + UD u */
    sv pa;
    v3 t1 = FN(v3_add)(FN(m3_v)(&A1, cb[i].a), FN(m3_v)(&B1, cb[i].l));
    v3 t2 = FN(v3_add)(FN(m3T_v)(&B1, cb[i].a), FN(m3_v)(&C1, cb[i].l));
    pa.a = FN(v3_axpy)(u, UD.a, FN(v3_add)(pA[i].a, t1));
    pa.l = FN(v3_axpy)(u, UD.l, FN(v3_add)(pA[i].l, t2));
    /* rotate blocks into parent axes: M' = Rl M Rl^T */
    const m3* Rl = &F->Rl[i];
    m3 T = FN(m3_mul)(Rl, &A1); m3 Ar = FN(m3_mulT)(&T, Rl);
    T = FN(m3_mul)(Rl, &B1); m3 Br = FN(m3_mulT)(&T, Rl);
    T = FN(m3_mul)(Rl, &C1); m3 Cr = FN(m3_mulT)(&T, Rl);
    /* shift by r: Cp = Cr ; Bp = Br + [r]x Cr ; Ap = Ar + [r]x Br^T - Bp [r]x */
    v3 r = F->tl[i];
    m3 Bp, Ap;
    for (int c = 0; c < 3; ++c) {
      v3 cc = FN(v3_cross)(r, FN(m3_col)(&Cr, c));
      Bp.m[0][c] = Br.m[0][c] + cc.x; Bp.m[1][c] = Br.m[1][c] + cc.y; Bp.m[2][c] = Br.m[2][c] + cc.z;
    }
    for (int c = 0; c < 3; ++c) { /* column c of [r]x Br^T = r x (row c of Br) */
      v3 cc = FN(v3_cross)(r, FN(m3_row)(&Br, c));
      Ap.m[0][c] = Ar.m[0][c] + cc.x; Ap.m[1][c] = Ar.m[1][c] + cc.y; Ap.m[2][c] = Ar.m[2][c] + cc.z;
    }
    for (int rr = 0; rr < 3; ++rr) { /* row rr of Bp [r]x = (row rr of Bp) x r */
      v3 cc = FN(v3_cross)(FN(m3_row)(&Bp, rr), r);
      Ap.m[rr][0] = Ap.m[rr][0] - cc.x; Ap.m[rr][1] = Ap.m[rr][1] - cc.y; Ap.m[rr][2] = Ap.m[rr][2] - cc.z;
    }
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) {
        IA[p].m[rr][c] = IA[p].m[rr][c] + Ap.m[rr][c];
        IB[p].m[rr][c] = IB[p].m[rr][c] + Bp.m[rr][c];
        IC[p].m[rr][c] = IC[p].m[rr][c] + Cr.m[rr][c];
      }
    sv fp = FN(xf_force_to_parent)(Rl, r, pa);
    pA[p].a = FN(v3_add)(pA[p].a, fp.a);
    pA[p].l = FN(v3_add)(pA[p].l, fp.l);
  }
  sv acc[NBM];
  for (int i = 0; i < nb; ++i) {
    int p = M->parent[i];
    sv ap;
    if (p < 0) { ap.a = FN(v3_make)(0, 0, 0); ap.l = FN(v3_make)(RC(0), RC(0), g); }
    else ap = acc[p];
    sv a = FN(xf_motion_to_child)(&F->Rl[i], F->tl[i], ap);
    a.a = FN(v3_add)(a.a, cb[i].a);
    a.l = FN(v3_add)(a.l, cb[i].l);
    real ua = FN(v3_dot)(R->U[i].a, a.a) + FN(v3_dot)(R->U[i].l, a.l);
    real qdd = R->Dinv[i] * (R->u[i] - ua);
    if (M->jtype[i] == 0) a.a.z = a.a.z + qdd; else a.l.z = a.l.z + qdd;
    acc[i] = a;
    R->qdd[i] = qdd;
  }
}

/* Column j of M^-1: response of all joint velocities to a unit impulse on joint j, reusing the ABA
 * factorisation (Bullet: calcAccelerationDeltasMultiDof [3P]). */
static void FN(minv_column)(const model* M, const fkres* F, const abares* R, int j, real* col) {
  const int nb = M->nb;
  sv ph[NBM];
  real uh[NBM];
  for (int i = 0; i < nb; ++i) { ph[i].a = FN(v3_make)(0, 0, 0); ph[i].l = ph[i].a; uh[i] = RC(0); }
  for (int i = j; i >= 0; --i) {
    if (!((M->anc[j] >> i) & 1u)) continue;
    real sp = (M->jtype[i] == 0) ? ph[i].a.z : ph[i].l.z;
    uh[i] = ((i == j) ? RC(1) : RC(0)) - sp;
    int p = M->parent[i];
    if (p < 0) continue;
    sv pa;
    pa.a = FN(v3_axpy)(uh[i], R->UD[i].a, ph[i].a);
    pa.l = FN(v3_axpy)(uh[i], R->UD[i].l, ph[i].l);
    sv fp = FN(xf_force_to_parent)(&F->Rl[i], F->tl[i], pa);
    ph[p].a = FN(v3_add)(ph[p].a, fp.a);
    ph[p].l = FN(v3_add)(ph[p].l, fp.l);
  }
  sv acc[NBM];
  for (int i = 0; i < nb; ++i) {
    int p = M->parent[i];
    sv a;
    if (p < 0) { a.a = FN(v3_make)(0, 0, 0); a.l = a.a; }
    else a = FN(xf_motion_to_child)(&F->Rl[i], F->tl[i], acc[p]);
    real ua = FN(v3_dot)(R->U[i].a, a.a) + FN(v3_dot)(R->U[i].l, a.l);
    real d = R->Dinv[i] * (uh[i] - ua);
    if (M->jtype[i] == 0) a.a.z = a.a.z + d; else a.l.z = a.l.z + d;
    acc[i] = a;
    col[i] = d;
  }
}

/* ------------------------------------------------------------------ contacts */
typedef struct {
  int pair;        /* PIH_PAIR_* bit */
  int has_fb;      /* free body involved (normal points INTO the free body) */
  int arm_body;    /* >=0: arm body involved (normal points AWAY from the arm body when has_fb, else into it) */
  v3 p;            /* contact point (env frame) */
  v3 n;            /* unit normal */
  real dist;       /* signed distance (negative = penetration) */
  real mu;
} FN(contact);
#define contact FN(contact)

typedef struct { contact c[PIH_MAXC]; int n; } FN(contactset);
#define contactset FN(contactset)

/* keep the PIH_MAXC deepest candidates; ties keep the earlier candidate */
static inline void FN(contact_push)(contactset* S, const contact* c) {
  if (!(c->dist < RC(PIH_CONTACT_MARGIN))) return;
  if (S->n < PIH_MAXC) { S->c[S->n++] = *c; return; }
  int w = 0;
  for (int i = 1; i < PIH_MAXC; ++i) if (S->c[i].dist > S->c[w].dist) w = i;
  if (c->dist < S->c[w].dist) S->c[w] = *c;
}

/* signed distance + outward 2-D normal of the rectangle |a-ac|<=ha, |b|<=hb in the (a,b) half plane:
 * shared by the annular hole (a = radial, b = axial) and the capped cylinder peg (ac = 0 ... handled by caller). */
static inline real FN(sd_rect)(real da, real db, real sa, real sb, real* na, real* nb_) {
  /* da = |a-ac|-ha, db = |b|-hb ; sa, sb = signs of (a-ac), b */
  if (da > RC(0) && db > RC(0)) {
    real l = SQRT(FMA(da, da, db * db));
    real il = RC(1) / l;
    *na = sa * da * il; *nb_ = sb * db * il;
    return l;
  }
  if (da > db) { *na = sa; *nb_ = RC(0); return da; }
  *na = RC(0); *nb_ = sb; return db;
}

/* distance / outward normal of the free body's surface at env-frame point x (body pose R,p) */
static inline real FN(fb_sdf)(const model* M, const m3* R, v3 p, v3 x, v3* n_env) {
  v3 l = FN(m3T_v)(R, FN(v3_sub)(x, p));
  v3 nl;
  real d;
  if (M->fb_shape == 0) { /* capped cylinder, axis local y */
    real rho = SQRT(FMA(l.z, l.z, l.x * l.x));
    real da = rho - M->fb_half.x, db = ABS(l.y) - M->fb_half.y;
    real na, nb_;
    d = FN(sd_rect)(da, db, RC(1), (l.y < RC(0)) ? RC(-1) : RC(1), &na, &nb_);
    real ir = (rho > RC(1e-12)) ? RC(1) / rho : RC(0);
    nl = FN(v3_make)(na * l.x * ir, nb_, na * l.z * ir);
  } else { /* box */
    real dx = ABS(l.x) - M->fb_half.x, dy = ABS(l.y) - M->fb_half.y, dz = ABS(l.z) - M->fb_half.z;
    real sx = (l.x < RC(0)) ? RC(-1) : RC(1), sy = (l.y < RC(0)) ? RC(-1) : RC(1), sz = (l.z < RC(0)) ? RC(-1) : RC(1);
    real px = dx > RC(0) ? dx : RC(0), py = dy > RC(0) ? dy : RC(0), pz = dz > RC(0) ? dz : RC(0);
    real o2 = FMA(pz, pz, FMA(py, py, px * px));
    if (o2 > RC(0)) {
      d = SQRT(o2);
      real il = RC(1) / d;
      nl = FN(v3_make)(sx * px * il, sy * py * il, sz * pz * il);
    } else if (dx > dy && dx > dz) { d = dx; nl = FN(v3_make)(sx, 0, 0); }
    else if (dy > dz) { d = dy; nl = FN(v3_make)(0, sy, 0); }
    else { d = dz; nl = FN(v3_make)(0, 0, sz); }
  }
  *n_env = FN(m3_v)(R, nl);
  return d;
}

/* distance / outward normal of the static annular hole at env-frame point x (axis = env x) */
static inline real FN(hole_sdf)(v3 x, v3* n) {
  real s = x.x - RC(PIH_HOLE_X), y = x.y - RC(PIH_HOLE_Y), z = x.z - RC(PIH_HOLE_Z);
  real rho = SQRT(FMA(z, z, y * y));
  const real rmid = RC(0.5 * (PIH_HOLE_RIN + PIH_HOLE_ROUT)), hw = RC(0.5 * (PIH_HOLE_ROUT - PIH_HOLE_RIN));
  real a = rho - rmid;
  real da = ABS(a) - hw, db = ABS(s) - RC(PIH_HOLE_HL);
  real na, nb_;
  real d = FN(sd_rect)(da, db, (a < RC(0)) ? RC(-1) : RC(1), (s < RC(0)) ? RC(-1) : RC(1), &na, &nb_);
  real ir = (rho > RC(1e-12)) ? RC(1) / rho : RC(0);
  *n = FN(v3_make)(nb_, na * y * ir, na * z * ir);
  return d;
}

/* the k-th surface sample point of the free body in its own frame (16 points):
 * cylinder: 8 rim points on each cap (k<8: +y cap); box: 8 corners (k<8), then unused */
static inline int FN(fb_sample)(const model* M, int k, v3* pl) {
  static const double cs[8][2] = {{1, 0}, {0.7071067811865476, 0.7071067811865476}, {0, 1}, {-0.7071067811865476, 0.7071067811865476},
                                  {-1, 0}, {-0.7071067811865476, -0.7071067811865476}, {0, -1}, {0.7071067811865476, -0.7071067811865476}};
  if (M->fb_shape == 0) {
    int kk = k & 7;
    real sy = (k < 8) ? RC(1) : RC(-1);
    *pl = FN(v3_make)(M->fb_half.x * RC(cs[kk][0]), sy * M->fb_half.y, M->fb_half.z * RC(cs[kk][1]));
    return 1;
  }
  if (k >= 8) return 0;
  *pl = FN(v3_make)((k & 1) ? M->fb_half.x : -M->fb_half.x, (k & 2) ? M->fb_half.y : -M->fb_half.y, (k & 4) ? M->fb_half.z : -M->fb_half.z);
  return 1;
}

/* FSM state of THIS step (update_state runs before the constraint is created / removed and before stepSimulation,
 * peg_in_hole.py:55,99-108): the grasp constraint exists while the state is 4, 5 or 6 */
static inline int FN(grasp_active)(const simcfg* C, int fsm_state_now) {
  return C->task == PIH_TASK_SCRIPTED && fsm_state_now >= PIH_FSM_GRASP_FIRST && fsm_state_now <= PIH_FSM_GRASP_LAST;
}

static void FN(collide)(const model* M, const simcfg* C, const envstate* S, const fkres* F, const m3* Rfb, contactset* CS) {
  CS->n = 0;
  v3 pfb = FN(v3_make)(S->ppos[0], S->ppos[1], S->ppos[2]);
  contact c;
  /* (0) grasp constraint of the scripted task: a pseudo contact whose three rows are bilateral and run along the env axes.
   *     p = end-effector origin (the pin point on both bodies), n carries the position error child point - parent point. */
  if (FN(grasp_active)(C, S->fsm_state)) {
    v3 pee; m3 Ree;
    FN(ee_pose)(M, F, &pee, &Ree);
    v3 pc = FN(m3_v_add)(Rfb, FN(v3_make)(RC(0), S->rand_y, RC(0)), pfb);
    c.pair = PIH_PAIR_GRASP; c.has_fb = 1; c.arm_body = M->ee_body; c.p = pee; c.n = FN(v3_sub)(pc, pee);
    c.dist = RC(-1e30); c.mu = RC(-1);
    FN(contact_push)(CS, &c);
  }
  /* (1) free body vs table plane z = PIH_TABLE_Z */
  if (M->fb_shape == 0) {
    v3 ax = FN(m3_col)(Rfb, 1);
    real naz = ax.z; /* n . a with n = +z */
    v3 nperp = FN(v3_make)(-(naz * ax.x), -(naz * ax.y), FMA(-naz, ax.z, RC(1)));
    real l2 = FN(v3_dot)(nperp, nperp);
    if (l2 > RC(0.0025)) { /* lying / tilted: lowest rim point of each cap */
      real il = RC(1) / SQRT(l2);
      v3 t = FN(v3_scale)(nperp, il);
      for (int k = 0; k < 2; ++k) {
        real sgn = k ? RC(-1) : RC(1);
        v3 cc = FN(v3_axpy)(sgn * M->fb_half.y, ax, pfb);
        v3 pt = FN(v3_axpy)(-M->fb_half.x, t, cc);
        c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = FN(v3_make)(0, 0, 1);
        c.dist = pt.z - RC(PIH_TABLE_Z); c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
        FN(contact_push)(CS, &c);
      }
    } else { /* standing on a cap: four rim points of the lower cap */
      real sgn = (naz > RC(0)) ? RC(-1) : RC(1);
      v3 cc = FN(v3_axpy)(sgn * M->fb_half.y, ax, pfb);
      v3 ex = FN(m3_col)(Rfb, 0), ez = FN(m3_col)(Rfb, 2);
      for (int k = 0; k < 4; ++k) {
        v3 pt = FN(v3_axpy)((k & 1) ? -M->fb_half.x : M->fb_half.x, (k & 2) ? ez : ex, cc);
        c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = FN(v3_make)(0, 0, 1);
        c.dist = pt.z - RC(PIH_TABLE_Z); c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
        FN(contact_push)(CS, &c);
      }
    }
  } else {
    for (int k = 0; k < 8; ++k) {
      v3 pl;
      FN(fb_sample)(M, k, &pl);
      v3 pt = FN(m3_v_add)(Rfb, pl, pfb);
      c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = FN(v3_make)(0, 0, 1);
      c.dist = pt.z - RC(PIH_TABLE_Z); c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
      FN(contact_push)(CS, &c);
    }
  }
  /* (2) peg vs annular hole (peg-in-hole task only): peg rim points against the hole solid, hole rim points
   *     against the peg solid; broadphase on centre distance */
  if (C->task != PIH_TASK_RANDOM_FLY) {
    v3 hc = FN(v3_make)(RC(PIH_HOLE_X), RC(PIH_HOLE_Y), RC(PIH_HOLE_Z));
    v3 dh = FN(v3_sub)(pfb, hc);
    const real reach = RC(PIH_PEG_HL + PIH_PEG_R + PIH_HOLE_ROUT + PIH_HOLE_HL + PIH_CONTACT_MARGIN);
    if (FN(v3_dot)(dh, dh) < reach * reach) {
      for (int k = 0; k < 16; ++k) {
        v3 pl, nn;
        FN(fb_sample)(M, k, &pl);
        v3 pt = FN(m3_v_add)(Rfb, pl, pfb);
        real d = FN(hole_sdf)(pt, &nn);
        c.pair = PIH_PAIR_PEG_HOLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = nn; c.dist = d;
        c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
        FN(contact_push)(CS, &c);
      }
      static const double cs[8][2] = {{1, 0}, {0.7071067811865476, 0.7071067811865476}, {0, 1}, {-0.7071067811865476, 0.7071067811865476},
                                      {-1, 0}, {-0.7071067811865476, -0.7071067811865476}, {0, -1}, {0.7071067811865476, -0.7071067811865476}};
      for (int k = 0; k < 32; ++k) { /* inner rims (k<16) then outer rims, both faces */
        real rr = (k < 16) ? RC(PIH_HOLE_RIN) : RC(PIH_HOLE_ROUT);
        real sx = (k & 8) ? RC(-PIH_HOLE_HL) : RC(PIH_HOLE_HL);
        v3 pt = FN(v3_make)(RC(PIH_HOLE_X) + sx, FMA(rr, RC(cs[k & 7][0]), RC(PIH_HOLE_Y)), FMA(rr, RC(cs[k & 7][1]), RC(PIH_HOLE_Z)));
        v3 nn;
        real d = FN(fb_sdf)(M, Rfb, pfb, pt, &nn);
        c.pair = PIH_PAIR_PEG_HOLE; c.has_fb = 1; c.arm_body = -1; c.p = pt;
        c.n = FN(v3_make)(-nn.x, -nn.y, -nn.z); /* normal must point into the peg */
        c.dist = d; c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
        FN(contact_push)(CS, &c);
      }
    }
  }
  /* (3) free body vs finger boxes (panda): capsule-chain spheres on the peg axis / box corners of the fly object
   *     against each finger box */
  for (int b = 0; b < M->nbox; ++b) {
    int body = M->box_body[b];
    m3 Rb = FN(m3_mul)(&F->Rw[body], &M->box_R[b]);
    v3 cb = FN(m3_v_add)(&F->Rw[body], M->box_c[b], F->pw[body]);
    v3 dd = FN(v3_sub)(pfb, cb);
    real reach = M->fb_half.y + M->fb_half.x + M->box_h[b].z + M->box_h[b].x + RC(PIH_CONTACT_MARGIN);
    if (!(FN(v3_dot)(dd, dd) < reach * reach)) continue;
    int npts = (M->fb_shape == 0) ? 5 : 8;
    for (int k = 0; k < npts; ++k) {
      v3 pl;
      real rad;
      if (M->fb_shape == 0) { pl = FN(v3_make)(RC(0), (M->fb_half.y - M->fb_half.x) * RC(0.5) * (real)(k - 2), RC(0)); rad = M->fb_half.x; }
      else { FN(fb_sample)(M, k, &pl); rad = RC(0); }
      v3 pt = FN(m3_v_add)(Rfb, pl, pfb);
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
      c.pair = b ? PIH_PAIR_PEG_FINGER2 : PIH_PAIR_PEG_FINGER1; c.has_fb = 1; c.arm_body = body;
      c.p = FN(v3_axpy)(-rad, nn, pt); c.n = nn; c.dist = d - rad; c.mu = M->fb_mu * RC(PIH_MU_FINGER);
      FN(contact_push)(CS, &c);
    }
  }
  /* (3b) free body vs arm collision spheres (hand, wrist, forearm): sphere centre against the free body's SDF.
   *      Finger-tip spheres are skipped on the panda: the finger boxes of (3) already cover them. */
  for (int s = 0; s < M->nsph; ++s) {
    int body = M->sph_body[s];
    if (body >= M->narm) continue;
    v3 cw = FN(m3_v_add)(&F->Rw[body], M->sph_c[s], F->pw[body]);
    v3 dd = FN(v3_sub)(cw, pfb);
    real reach = M->fb_half.x + M->fb_half.y + M->fb_half.z + M->sph_r[s] + RC(PIH_CONTACT_MARGIN);
    if (!(FN(v3_dot)(dd, dd) < reach * reach)) continue;
    v3 nn;
    real d = FN(fb_sdf)(M, Rfb, pfb, cw, &nn);
    c.pair = PIH_PAIR_FB_ARM; c.has_fb = 1; c.arm_body = body;
    c.n = FN(v3_make)(-nn.x, -nn.y, -nn.z);                  /* from the arm sphere into the free body */
    c.p = FN(v3_axpy)(-M->sph_r[s], nn, cw); c.dist = d - M->sph_r[s]; c.mu = M->fb_mu * RC(PIH_MU_DEFAULT);
    FN(contact_push)(CS, &c);
  }
  /* (4) arm collision spheres vs table plane */
  for (int s = 0; s < M->nsph; ++s) {
    int body = M->sph_body[s];
    v3 cw = FN(m3_v_add)(&F->Rw[body], M->sph_c[s], F->pw[body]);
    c.pair = PIH_PAIR_ARM_TABLE; c.has_fb = 0; c.arm_body = body; c.n = FN(v3_make)(0, 0, 1);
    c.p = FN(v3_make)(cw.x, cw.y, cw.z - M->sph_r[s]);
    c.dist = cw.z - M->sph_r[s] - RC(PIH_TABLE_Z); c.mu = RC(PIH_MU_DEFAULT) * RC(PIH_MU_DEFAULT);
    FN(contact_push)(CS, &c);
  }
}

/* btPlaneSpace1 [3P]: two tangents orthogonal to n */
static inline void FN(plane_space)(v3 n, v3* t1, v3* t2) {
  if (ABS(n.z) > RC(0.7071067811865476)) {
    real a = FMA(n.z, n.z, n.y * n.y);
    real k = RC(1) / SQRT(a);
    *t1 = FN(v3_make)(RC(0), -(n.z * k), n.y * k);
    *t2 = FN(v3_make)(a * k, -(n.x * t1->z), n.x * t1->y);
  } else {
    real a = FMA(n.y, n.y, n.x * n.x);
    real k = RC(1) / SQRT(a);
    *t1 = FN(v3_make)(-(n.y * k), n.x * k, RC(0));
    *t2 = FN(v3_make)(-(n.z * t1->y), n.z * t1->x, a * k);
  }
}

/* clamp |v| to PIH_MAX_COORD_VEL (btMultiBody m_maxCoordinateVelocity [3P]) */
static inline v3 FN(clamp_norm)(v3 v) {
  real n2 = FN(v3_dot)(v, v);
  if (n2 > RC(PIH_MAX_COORD_VEL * PIH_MAX_COORD_VEL)) return FN(v3_scale)(v, RC(PIH_MAX_COORD_VEL) / SQRT(n2));
  return v;
}

/* ------------------------------------------------------------------ test hook: the assembled velocity-level problem
 * of one env.step (rows, bounds, targets, M^-1, v*) and the solver's answer, for an INDEPENDENT check of the solution
 * (tests/test_oracle_lcp.py verifies the complementarity conditions and re-solves the box QP densely in numpy). */
#ifndef PIH_ORACLE_CAPTURE
#define PIH_ORACLE_CAPTURE
#define PIH_CAP_MAXR (3 * PIH_NB_MAX + 3 * PIH_MAXC)
#define PIH_CAP_NV (PIH_NB_MAX + 6)
typedef struct {
  int nrows, nv;
  double J[PIH_CAP_MAXR][PIH_CAP_NV];   /* generalized velocity = (arm joint rates[nb], free-body linear[3], angular[3]) */
  double tv[PIH_CAP_MAXR];              /* target row velocity */
  double lo[PIH_CAP_MAXR], hi[PIH_CAP_MAXR];   /* impulse bounds; friction rows: hi = mu, lo = index of their normal row, kind = 1 */
  int kind[PIH_CAP_MAXR];               /* 0 box-bounded, 1 friction (|lambda| <= mu * lambda[normal]) */
  double lam[PIH_CAP_MAXR];
  double Minv[PIH_CAP_NV][PIH_CAP_NV];
  double vstar[PIH_CAP_NV], vplus[PIH_CAP_NV];
  int iters;
} pih_capture;
extern pih_capture* pih_dbg_capture;
#endif

/* ------------------------------------------------------------------ one env.step */
typedef struct {
  int has_fb, arm_body;
  v3 dir[3];        /* n, t1, t2 */
  v3 jang[3];       /* r x dir               (free body)            */
  v3 mang[3];       /* Iw^-1 (r x dir)                              */
  real ja[3][NBM];  /* arm Jacobian rows of v_arm(p) . dir          */
  real ma[3][NBM];  /* Minv ja^T                                    */
  real dinv[3], diag[3], b[3], lam[3], mu;
} FN(crow);
#define crow FN(crow)

static void FN(reset_env)(const model* M, const simcfg* C, envstate* S, uint32_t gid) {
  /* envs/peg_in_hole.py:233-245 reset distributions; Python's global MT19937 stream cannot be matched, so the
   * DISTRIBUTIONS are reproduced with Philox keyed (seed; env id, episode) - NEW */
  float u[4];
  pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_RESET, u);
  for (int i = 0; i < NBM; ++i) { S->q[i] = (i < M->nb) ? M->rest[i] : RC(0); S->qd[i] = RC(0); }
  if (C->task != PIH_TASK_RANDOM_FLY) {
    S->ppos[0] = (real)fmaf(u[0], 0.4f, -0.2f);   /* random.uniform(-0.2, 0.2)  peg_in_hole.py:239 */
    S->ppos[1] = (real)fmaf(u[1], 0.2f, -0.6f);   /* random.uniform(-0.4,-0.6)  peg_in_hole.py:239 */
    S->ppos[2] = RC(PIH_PEG_SPAWN_Z);
    S->pvel[0] = S->pvel[1] = S->pvel[2] = RC(0);
  } else {
    /* random-fly (NEW, source missing from the reference snapshot): the object is thrown from
     * 1.2 m out at a random azimuth so that it passes through random_pos_in_panda_space() (envs/utils.py:97-107:
     * x ~ U(-.7,.7), y = (sqrt(U(0,.49-x^2)) - U(0,.4)) * choice(+-1), rejected while .49-x^2-y^2 < 0, z on the
     * sphere + 0.2; the rejection loop is bounded by PIH_FLY_MAX_ATTEMPTS, NEW) after 0.5 s */
    float x = 0.0f, y = 0.0f, zz = 0.49f;
    for (int k = 0; k < PIH_FLY_MAX_ATTEMPTS; ++k) {   /* the `while` loop of utils.py:101-104, attempt k = Philox stream FLY + k */
      float v[4];
      pih_rng4(C->seed, gid, (uint32_t)S->episode, PIH_STREAM_FLY + k, v);
      const float xk = fmaf(v[0], 1.4f, -0.7f);                         /* random.uniform(-length, length) */
      const float a2 = fmaf(-xk, xk, 0.49f);
      const float yk = (sqrtf(v[1] * (a2 > 0.0f ? a2 : 0.0f)) - v[2] * 0.4f) * ((v[3] < 0.5f) ? -1.0f : 1.0f);
      const float rem = fmaf(-yk, yk, a2);
      if (!(rem < 0.0f)) { x = xk; y = yk; zz = rem; break; }
    }
    float z = sqrtf(zz > 0.0f ? zz : 0.0f) + 0.2f;
    float ca, sa;
    pih_sincosf_det(u[3] * 6.2831855f, &sa, &ca);
    const float T = 0.5f, D = 1.2f;
    float sx = x + D * ca, sy = y + D * sa, sz = z + 0.3f;
    S->ppos[0] = (real)sx; S->ppos[1] = (real)sy; S->ppos[2] = (real)sz;
    S->pvel[0] = (real)((x - sx) / T); S->pvel[1] = (real)((y - sy) / T);
    S->pvel[2] = (real)((z - sz) / T + 0.5f * 9.8f * T);
  }
  S->pquat[0] = S->pquat[1] = S->pquat[2] = RC(0); S->pquat[3] = RC(1);
  S->pang[0] = S->pang[1] = S->pang[2] = RC(0);
  S->step = 0;
  S->flags = 0;
  /* scripted task: cur_state 0, random_vector = [0, U(-0.03, 0.03), 0] (peg_in_hole.py:255-267) */
  S->fsm_state = 0; S->fsm_t = 0;
  S->rand_y = (real)fmaf(u[3], 0.06f, -0.03f);
  S->grasp_ang = RC(PIH_LABEL_NONE);
  for (int i = 0; i < 7; ++i) S->mtgt[i] = (i < M->narm) ? M->rest[i] : RC(0);
}

static void FN(observe)(const model* M, const simcfg* C, const envstate* S, const real off[3], real* obs) {
  /* observation_space Box(5): [finger1, finger2, EE x, EE y, EE z]  (envs/peg_in_hole.py:13); EE = getLinkState(ee)[0]
   * in WORLD coordinates = env frame + offset.  ur5 has no fingers: slots 0,1 = 0 (NEW). */
  fkres F;
  FN(fk)(M, S->q, M->nb, &F);
  v3 pee; m3 Ree;
  FN(ee_pose)(M, &F, &pee, &Ree);
  (void)C;
  obs[0] = (M->nb > M->narm) ? S->q[M->narm] : RC(0);
  obs[1] = (M->nb > M->narm + 1) ? S->q[M->narm + 1] : RC(0);
  obs[2] = pee.x + off[0]; obs[3] = pee.y + off[1]; obs[4] = pee.z + off[2];
}


/* ------------------------------------------------------------------ camera (PegInHole.render, peg_in_hole.py:276-304)
 * Ray caster over the same analytic primitives the contact code uses: table plane, free body (capped cylinder or
 * box), annular hole, finger boxes.  TinyRenderer itself is [3P] and not restated; what is kept is the camera model
 * (computeViewMatrix eye=EE, target=EE-(0,0,10), up=(0,1,0); computeProjectionMatrixFOV 60/1/0.001/1000), the
 * 300x300x4 (depth, r, g, b) layout and the OpenGL depth-buffer encoding.  Colours / flat Lambert shading are NEW. */
typedef struct {
  v3 eye;
  m3 Rfb; v3 pfb;
  m3 Rbox[PIH_MAX_BOX]; v3 cbox[PIH_MAX_BOX];
} FN(scene);
#define scene FN(scene)

static void FN(build_scene)(const model* M, const envstate* S, scene* SC) {
  fkres F;
  FN(fk)(M, S->q, M->nb, &F);
  m3 Ree;
  FN(ee_pose)(M, &F, &SC->eye, &Ree);
  SC->Rfb = FN(quat_to_m3)(S->pquat);
  SC->pfb = FN(v3_make)(S->ppos[0], S->ppos[1], S->ppos[2]);
  for (int b = 0; b < PIH_MAX_BOX; ++b) {
    int body = (b < M->nbox) ? M->box_body[b] : 0;
    SC->Rbox[b] = FN(m3_mul)(&F.Rw[body], &M->box_R[b]);
    SC->cbox[b] = FN(m3_v_add)(&F.Rw[body], M->box_c[b], F.pw[body]);
  }
}

/* nearest hit of the ray o + s d (s > smin) with an axis-aligned box of half extents h (box frame) */
static inline int FN(ray_box)(v3 o, v3 d, v3 h, real smin, real* s_hit, v3* n) {
  real oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, hh[3] = {h.x, h.y, h.z};
  real s0 = smin, s1 = RC(1e30);
  int ax0 = -1; real sg0 = RC(0);
  for (int a = 0; a < 3; ++a) {
    if (ABS(dd[a]) < RC(1e-12)) { if (ABS(oo[a]) > hh[a]) return 0; continue; }
    real inv = RC(1) / dd[a];
    real ta = (-hh[a] - oo[a]) * inv, tb = (hh[a] - oo[a]) * inv;
    real sgn = RC(-1);
    if (ta > tb) { real t = ta; ta = tb; tb = t; sgn = RC(1); }
    if (ta > s0) { s0 = ta; ax0 = a; sg0 = sgn; }
    if (tb < s1) s1 = tb;
  }
  if (s0 > s1 || ax0 < 0) return 0;
  *s_hit = s0;
  *n = FN(v3_make)(ax0 == 0 ? sg0 : RC(0), ax0 == 1 ? sg0 : RC(0), ax0 == 2 ? sg0 : RC(0));
  return 1;
}

/* nearest hit with a capped cylinder of radius r, half length hl whose axis is the frame's `axis` coordinate
 * (1 = y for the peg).  (u, v) are the two radial coordinates. */
static inline int FN(ray_cyl_y)(v3 o, v3 d, real r, real hl, real smin, real* s_hit, v3* n) {
  int hit = 0; real best = RC(1e30);
  real a = FMA(d.z, d.z, d.x * d.x), b = FMA(o.z, d.z, o.x * d.x), c = FMA(o.z, o.z, o.x * o.x) - r * r;
  if (a > RC(1e-12)) {
    real disc = FMA(b, b, -(a * c));
    if (disc >= RC(0)) {
      real s = (-b - SQRT(disc)) / a;
      real y = FMA(s, d.y, o.y);
      if (s > smin && ABS(y) <= hl) { best = s; hit = 1; real ir = RC(1) / r; *n = FN(v3_make)(FMA(s, d.x, o.x) * ir, RC(0), FMA(s, d.z, o.z) * ir); }
    }
  }
  if (ABS(d.y) > RC(1e-12)) {
    real inv = RC(1) / d.y;
    for (int k = 0; k < 2; ++k) {
      real yc = k ? -hl : hl;
      real s = (yc - o.y) * inv;
      real x = FMA(s, d.x, o.x), z = FMA(s, d.z, o.z);
      if (s > smin && s < best && FMA(z, z, x * x) <= r * r) { best = s; hit = 1; *n = FN(v3_make)(RC(0), k ? RC(-1) : RC(1), RC(0)); }
    }
  }
  *s_hit = best;
  return hit;
}

/* nearest hit with the annular hole: tube r_in <= rho <= r_out, |x| <= hl, axis x, centre at the origin of (o, d) */
static inline int FN(ray_tube_x)(v3 o, v3 d, real smin, real* s_hit, v3* n) {
  const real rin = RC(PIH_HOLE_RIN), rout = RC(PIH_HOLE_ROUT), hl = RC(PIH_HOLE_HL);
  real a = FMA(d.z, d.z, d.y * d.y), b = FMA(o.z, d.z, o.y * d.y), q2 = FMA(o.z, o.z, o.y * o.y);
  /* slab interval */
  real sa = smin, sb = RC(1e30); int slab_first = 0; real slab_sign = RC(0);
  if (ABS(d.x) > RC(1e-12)) {
    real inv = RC(1) / d.x;
    real ta = (-hl - o.x) * inv, tb = (hl - o.x) * inv; real sg = RC(-1);
    if (ta > tb) { real t = ta; ta = tb; tb = t; sg = RC(1); }
    if (ta > sa) { sa = ta; slab_first = 1; slab_sign = sg; }
    sb = tb;
  } else if (ABS(o.x) > hl) return 0;
  if (!(a > RC(1e-12))) return 0;   /* ray parallel to the axis: looking straight through or at a cap edge-on */
  real disco = FMA(b, b, -(a * (q2 - rout * rout)));
  if (disco < RC(0)) return 0;
  real sq = SQRT(disco), ia = RC(1) / a;
  real so1 = (-b - sq) * ia, so2 = (-b + sq) * ia;
  int first_is_slab = slab_first;
  real s0 = sa, s1 = (sb < so2) ? sb : so2;
  if (so1 > s0) { s0 = so1; first_is_slab = 0; }
  if (s0 > s1) return 0;
  real disci = FMA(b, b, -(a * (q2 - rin * rin)));
  real si1 = RC(1e30), si2 = RC(-1e30);
  if (disci > RC(0)) { real sqi = SQRT(disci); si1 = (-b - sqi) * ia; si2 = (-b + sqi) * ia; }
  if (s0 < si1 || s0 >= si2) {              /* entry point is in the solid */
    if (first_is_slab && s0 == smin) return 0;
    *s_hit = s0;
    if (first_is_slab) *n = FN(v3_make)(slab_sign, RC(0), RC(0));
    else { real ir = RC(1) / rout; *n = FN(v3_make)(RC(0), FMA(s0, d.y, o.y) * ir, FMA(s0, d.z, o.z) * ir); }
    return (s0 > smin);
  }
  if (si2 < s1 && si2 > smin) {             /* entered through the bore: hit the inner wall from inside */
    real ir = RC(1) / rin;
    *s_hit = si2;
    *n = FN(v3_make)(RC(0), -(FMA(si2, d.y, o.y) * ir), -(FMA(si2, d.z, o.z) * ir));
    return 1;
  }
  return 0;
}

static void FN(render_pixel)(const model* M, const simcfg* C, const scene* SC, int row, int col, int W, int H, float out[4]) {
  const real th = RC(PIH_CAM_TAN_HALF_FOV);
  real ndx = ((real)col + RC(0.5)) / (real)W * RC(2) - RC(1);
  real ndy = RC(1) - ((real)row + RC(0.5)) / (real)H * RC(2);
  v3 d = FN(v3_make)(ndx * th, ndy * th, RC(-1));     /* eye-space depth == ray parameter */
  v3 o = SC->eye;
  const real smin = RC(PIH_CAM_NEAR);
  real best = RC(1e30); v3 nb = FN(v3_make)(0, 0, 1); int obj = -1;
  { real s = o.z - RC(PIH_TABLE_Z); if (s > smin) { best = s; obj = 0; } }
  { /* free body */
    v3 ol = FN(m3T_v)(&SC->Rfb, FN(v3_sub)(o, SC->pfb)), dl = FN(m3T_v)(&SC->Rfb, d);
    real s; v3 n; int h;
    if (M->fb_shape == 0) h = FN(ray_cyl_y)(ol, dl, M->fb_half.x, M->fb_half.y, smin, &s, &n);
    else h = FN(ray_box)(ol, dl, M->fb_half, smin, &s, &n);
    if (h && s < best) { best = s; obj = 1; nb = FN(m3_v)(&SC->Rfb, n); }
  }
  if (C->task != PIH_TASK_RANDOM_FLY) {
    v3 oh = FN(v3_make)(o.x - RC(PIH_HOLE_X), o.y - RC(PIH_HOLE_Y), o.z - RC(PIH_HOLE_Z));
    real s; v3 n;
    if (FN(ray_tube_x)(oh, d, smin, &s, &n) && s < best) { best = s; obj = 2; nb = n; }
  }
  for (int b = 0; b < M->nbox; ++b) {
    v3 ol = FN(m3T_v)(&SC->Rbox[b], FN(v3_sub)(o, SC->cbox[b])), dl = FN(m3T_v)(&SC->Rbox[b], d);
    real s; v3 n;
    if (FN(ray_box)(ol, dl, M->box_h[b], smin, &s, &n) && s < best) { best = s; obj = 3; nb = FN(m3_v)(&SC->Rbox[b], n); }
  }
  const real fn = RC(PIH_CAM_FAR), nn = RC(PIH_CAM_NEAR);
  real depth = RC(1);
  real base[3] = {RC(255), RC(255), RC(255)};
  real shade = RC(1);
  if (obj >= 0) {
    depth = FMA(RC(0.5), (fn + nn) / (fn - nn) - (RC(2) * fn * nn) / ((fn - nn) * best), RC(0.5));
    static const double cols[4][3] = {{153, 128, 102}, {232, 232, 232}, {232, 232, 232}, {204, 204, 204}};
    base[0] = RC(cols[obj][0]); base[1] = RC(cols[obj][1]); base[2] = RC(cols[obj][2]);
    const v3 l = FN(v3_make)(RC(0.2672612419124244), RC(0.1781741612749496), RC(0.9449111825230680));
    real nl = FN(v3_dot)(nb, l);
    nl = (nl < RC(0)) ? RC(0) : nl;
    shade = FMA(RC(0.65), nl, RC(0.35));
  }
  out[0] = (float)depth;
  for (int k = 0; k < 3; ++k) {
    real v = base[k] * shade + RC(0.5);
    out[1 + k] = (float)(real)(int)v;
  }
}

/* apply_action: panda_execute / ur_execute (envs/utils.py:60-82), or the scripted state machine of the shipped task
 * (peg_in_hole.py:39-212): joint position targets, gains and force caps of this step.  F = FK of S->q. */
static void FN(apply_action)(const model* M, const simcfg* C, envstate* S, const fkres* Fp, const real* action, const real off[3],
                             real* qstar, real* kpc, real* mfc, int* fsm_done_out) {
  const int nb = M->nb;
  const fkres F = *Fp;
  v3 pee; m3 Ree;
  FN(ee_pose)(M, &F, &pee, &Ree);
  v3 tgt;
  real tq[4];
  int fsm_done = 0;
  for (int j = 0; j < nb; ++j) { kpc[j] = M->kp[j]; mfc[j] = M->maxf[j]; }
  if (C->task == PIH_TASK_SCRIPTED) {
    /* The task the reference ships: apply_action is `pass` (peg_in_hole.py:30-31) and the arm is driven by the
     * scripted grasp->insert state machine of random_grasp/grasp_process (peg_in_hole.py:39-212).  One engine step
     * here = one pass of its `while not done` loop body: update_state, grasp target, grasp_process, stepSimulation. */
    static const int fsm_steps[PIH_FSM_NSTATES] = {PIH_FSM_STEPS_LIST};
    const int entered = (S->fsm_t == fsm_steps[S->fsm_state]);
    if (entered) {                                                   /* update_state, peg_in_hole.py:206-212 */
      S->fsm_state += 1; S->fsm_t = 0;
      if (S->fsm_state >= PIH_FSM_NSTATES) S->fsm_state = 0;
    }
    S->fsm_t += 1;
    const int st = S->fsm_state;
    /* grasp point = COM of the grasp link + R(q) random_vector; yaw of that link (peg_in_hole.py:58-62,125,215-225) */
    m3 Rp = FN(quat_to_m3)(S->pquat);
    /* on entering state 2 the reference captures the grasp image and its label rectangle; the rectangle's angle is
     * atan2(rotate_vector[1], rotate_vector[0]) with rotate_vector = R(q) [0, rand_y, 0] (peg_in_hole.py:58,65-72) */
    if (entered && st == 2) {
      S->grasp_ang = FN(det_atan2)(S->rand_y * Rp.m[1][1], S->rand_y * Rp.m[0][1]);
      for (int k = 0; k < 9; ++k) S->latch[k] = S->q[k];             /* self.grasp_img = self.render(), peg_in_hole.py:65-66 */
      for (int k = 0; k < 3; ++k) S->latch[9 + k] = S->ppos[k];
      for (int k = 0; k < 4; ++k) S->latch[12 + k] = S->pquat[k];
    }
    v3 tp = FN(m3_v_add)(&Rp, FN(v3_make)(RC(0), S->rand_y, RC(0)), FN(v3_make)(S->ppos[0], S->ppos[1], S->ppos[2]));
    real qx = S->pquat[0], qy = S->pquat[1], qz = S->pquat[2], qw = S->pquat[3];
    real yaw = FN(det_atan2)(RC(2) * FMA(qx, qy, qw * qz), FMA(qw, qw, qx * qx) - FMA(qy, qy, qz * qz));
    real cur[3] = {pee.x, pee.y, pee.z}, tar[3] = {tp.x, tp.y, tp.z}, res[3];
    int arm_cmd = 0, clamp = 1;
    real yaw_cmd = RC(0), zoff = RC(0);
    if (st == 1 || st == 2) { arm_cmd = 1; yaw_cmd = RC(1.5707963267948966) + yaw; zoff = (st == 1) ? RC(0.05) : RC(-0.01); }
    else if (st == 4 || st == 5 || st == 6) {
      arm_cmd = 1; yaw_cmd = RC(-3.141592653589793);
      tar[0] = RC(PIH_HOLE_X) - ((st == 4) ? RC(0.2) : ((st == 5) ? RC(0.04) : RC(0)));
      tar[1] = RC(PIH_HOLE_Y); tar[2] = RC(PIH_HOLE_Z);
      clamp = (st != 6);
    } else if (st == 8) { arm_cmd = 1; yaw_cmd = RC(1.5707963267948966); tar[0] = RC(0.2); tar[1] = RC(-0.6); tar[2] = RC(0.4); clamp = 0; }
    else if (st == 9) fsm_done = 1;
    if (arm_cmd) {
      for (int i = 0; i < 3; ++i) {                                   /* vel_constraint(cur, target, self.dv) */
        real diff = tar[i] - cur[i];
        if (clamp && ABS(diff) > RC(PIH_FSM_DV)) res[i] = cur[i] + ((diff > RC(0)) ? RC(PIH_FSM_DV) : RC(-PIH_FSM_DV));
        else res[i] = cur[i] + diff;
      }
      tgt = FN(v3_make)(res[0], res[1], res[2] + zoff);
      FN(euler_to_quat)(RC(0), RC(-3.141592653589793), yaw_cmd, tq);
      real qik[NBM];
      FN(ik_dls)(M, C, &F, S->q, tgt, tq, qik);
      for (int j = 0; j < M->narm && j < 7; ++j) S->mtgt[j] = qik[j];
    }
    for (int j = 0; j < M->narm; ++j) { qstar[j] = S->mtgt[j < 7 ? j : 0]; kpc[j] = RC(PIH_FSM_KP); mfc[j] = RC(PIH_FSM_ARM_FORCE); }
    int closed = (st >= 3 && st <= 6);
    for (int j = M->narm; j < nb; ++j) {
      qstar[j] = closed ? RC(PIH_FSM_FINGER_CLOSED) : RC(PIH_FSM_FINGER_OPEN);
      kpc[j] = RC(PIH_FSM_KP);
      mfc[j] = closed ? RC(PIH_FSM_FINGER_FORCE_HI) : RC(PIH_FSM_FINGER_FORCE_LO);
    }
  } else {
    real ax = action[0], ay = action[1], az = action[2];
    if (!C->action_frame_env) { ax = ax - off[0]; ay = ay - off[1]; az = az - off[2]; }
    if (C->arm == PIH_ARM_PANDA) {
      /* vel_constraint(cur, action[:3], dv) envs/utils.py:85-95 */
      real cur[3] = {pee.x, pee.y, pee.z}, tar[3] = {ax, ay, az}, res[3];
      for (int i = 0; i < 3; ++i) {
        real diff = tar[i] - cur[i];
        if (ABS(diff) > RC(PIH_DV)) res[i] = cur[i] + ((diff > RC(0)) ? RC(PIH_DV) : RC(-PIH_DV));
        else res[i] = cur[i] + diff;
      }
      tgt = FN(v3_make)(res[0], res[1], res[2]);
      tq[0] = RC(0); tq[1] = RC(-1); tq[2] = RC(0); tq[3] = RC(0); /* getQuaternionFromEuler([0,-pi,0]) utils.py:61 */
    } else {
      tgt = FN(v3_make)(ax, ay, az);                                /* absolute, no clamp: utils.py:71 */
      FN(euler_to_quat)(action[3], action[4], action[5], tq);      /* utils.py:72 */
    }
    FN(ik_dls)(M, C, &F, S->q, tgt, tq, qstar);
    for (int j = M->narm; j < nb; ++j) qstar[j] = action[3];         /* fingers = action[3]: utils.py:66,68 */
  }
  *fsm_done_out = fsm_done;
}

/* returns contact count; outputs written only if the env was stepped (done envs are frozen: base_env.py:61-71) */
static int FN(step_env)(const model* M, const simcfg* C, envstate* S, const real* action, const real off[3], uint32_t gid,
                        real* obs, real* reward, uint8_t* done) {
  if (S->flags & PIH_FLAG_DONE) return 0;
  const int nb = M->nb;
  const real dt = C->dt;
  fkres F;
  FN(fk)(M, S->q, nb, &F);
  real qstar[NBM], kpc[NBM], mfc[NBM];
  int fsm_done = 0;
  FN(apply_action)(M, C, S, &F, action, off, qstar, kpc, mfc, &fsm_done);
  /* ---- stepSimulation: unconstrained accelerations ---- */
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
  /* free body: gravity + gyroscopic term in the body frame */
  m3 Rfb = FN(quat_to_m3)(S->pquat);
  v3 fv = FN(v3_make)(S->pvel[0], S->pvel[1], FMA(dt, RC(-PIH_GRAVITY), S->pvel[2]));
  v3 fw = FN(v3_make)(S->pang[0], S->pang[1], S->pang[2]);
  {
    v3 wb = FN(m3T_v)(&Rfb, fw);
    v3 Iw = FN(v3_make)(M->fb_idiag.x * wb.x, M->fb_idiag.y * wb.y, M->fb_idiag.z * wb.z);
    v3 g = FN(v3_cross)(wb, Iw);
    v3 al = FN(v3_make)(-(g.x * M->fb_iinv.x), -(g.y * M->fb_iinv.y), -(g.z * M->fb_iinv.z));
    fw = FN(clamp_norm)(FN(v3_axpy)(dt, FN(m3_v)(&Rfb, al), fw));
  }
  /* world inverse inertia  Rfb diag(iinv) Rfb^T */
  m3 Iinv;
  {
    m3 T;
    for (int r = 0; r < 3; ++r) { T.m[r][0] = Rfb.m[r][0] * M->fb_iinv.x; T.m[r][1] = Rfb.m[r][1] * M->fb_iinv.y; T.m[r][2] = Rfb.m[r][2] * M->fb_iinv.z; }
    Iinv = FN(m3_mulT)(&T, &Rfb);
  }
  const real im = RC(1) / M->fb_mass;
  /* ---- constraint rows ---- */
  /* motors: POSITION_CONTROL, v* = kp (q*-q)/dt (kd=1, target velocity 0), |impulse| <= maxForce dt */
  real m_dinv[NBM], m_b[NBM], m_lam[NBM], m_max[NBM];
  int l_on[2][NBM];            /* [0] lower, [1] upper */
  real l_b[2][NBM], l_lam[2][NBM];
  const real idt = RC(1) / dt;
  for (int j = 0; j < nb; ++j) {
    /* position targets outside the URDF limits are unreachable: the motor aims at the
     * limit instead of fighting the limit row with its 1e5 N force cap (NEW, keeps PGS well conditioned) */
    real qt = qstar[j];
    qt = (qt < M->lo[j]) ? M->lo[j] : ((qt > M->hi[j]) ? M->hi[j] : qt);
    real vstar = kpc[j] * (qt - S->q[j]) * idt;
    m_dinv[j] = RC(1) / Minv[j][j];
    m_b[j] = (vstar - vs[j]) * m_dinv[j];
    m_lam[j] = RC(0);
    m_max[j] = mfc[j] * dt;
    /* joint limits (btMultiBodyJointLimitConstraint [3P]): a row exists for each limit that can be reached this step,
     * judged by the velocities that may drive the joint towards it (motor target v*, unconstrained velocity vs):
     *   lower: gap < max(-v*, -vs, 0) dt + 0.005      upper: gap < max(v*, vs, 0) dt + 0.005
     * inside the limit the joint may close the gap, outside it is pushed back with ERP */
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
  /* contacts */
  contactset CS;
  FN(collide)(M, C, S, &F, &Rfb, &CS);
  crow rows[PIH_MAXC];
  int pairmask = 0;
  v3 pfb = FN(v3_make)(S->ppos[0], S->ppos[1], S->ppos[2]);
  for (int k = 0; k < CS.n; ++k) {
    const contact* c = &CS.c[k];
    crow* r = &rows[k];
    pairmask |= c->pair;
    r->has_fb = c->has_fb; r->arm_body = c->arm_body; r->mu = c->mu;
    const int weld = c->pair == PIH_PAIR_GRASP;
    if (weld) { r->dir[0] = FN(v3_make)(1, 0, 0); r->dir[1] = FN(v3_make)(0, 1, 0); r->dir[2] = FN(v3_make)(0, 0, 1); }
    else { r->dir[0] = c->n; FN(plane_space)(c->n, &r->dir[1], &r->dir[2]); }
    v3 rr = FN(v3_sub)(c->p, pfb);
    for (int d = 0; d < 3; ++d) {
      real diag = RC(0), vrel = RC(0);
      if (c->has_fb) {
        r->jang[d] = FN(v3_cross)(rr, r->dir[d]);
        r->mang[d] = FN(m3_v)(&Iinv, r->jang[d]);
        diag = im + FN(v3_dot)(r->jang[d], r->mang[d]);
        vrel = FN(v3_dot)(r->dir[d], fv) + FN(v3_dot)(r->jang[d], fw);
      }
      if (c->arm_body >= 0) {
        /* arm Jacobian of the point velocity along dir; sign: relative velocity = v_fb - v_arm when has_fb */
        real sg = c->has_fb ? RC(-1) : RC(1);
        for (int j = 0; j < nb; ++j) {
          real v = RC(0);
          if ((M->anc[c->arm_body] >> j) & 1u) {
            v3 z = FN(m3_col)(&F.Rw[j], 2);
            if (M->jtype[j] == 0) v = FN(v3_dot)(r->dir[d], FN(v3_cross)(z, FN(v3_sub)(c->p, F.pw[j])));
            else v = FN(v3_dot)(r->dir[d], z);
            v = sg * v;
          }
          r->ja[d][j] = v;
        }
        real da = RC(0), va = RC(0);
        for (int i = 0; i < nb; ++i) {
          real s = RC(0);
          for (int j = 0; j < nb; ++j) s = FMA(Minv[i][j], r->ja[d][j], s);
          r->ma[d][i] = s;
        }
        for (int j = 0; j < nb; ++j) { da = FMA(r->ja[d][j], r->ma[d][j], da); va = FMA(r->ja[d][j], vs[j], va); }
        diag = diag + da;
        vrel = vrel + va;
      }
      r->dinv[d] = RC(1) / diag;
      r->diag[d] = diag;
      real tv = RC(0);
      if (weld) tv = -(RC(PIH_ERP) * (d == 0 ? c->n.x : (d == 1 ? c->n.y : c->n.z)) * idt);
      else if (d == 0) tv = (c->dist > RC(0)) ? -(c->dist * idt) : -(RC(PIH_ERP) * c->dist * idt);
      r->b[d] = (tv - vrel) * r->dinv[d];
      r->lam[d] = RC(0);
    }
  }
  /* ---- projected Gauss-Seidel (sequential impulses), Bullet row order: motors, limits, normals, friction ---- */
  real dva[NBM];
  for (int j = 0; j < nb; ++j) dva[j] = RC(0);
  v3 dvl = FN(v3_make)(0, 0, 0), dvw = dvl;
  /* Early exit: PyBullet's solverResidualThreshold (default 1e-7) on the largest squared velocity change
   * (deltaImpulse / jacDiagABInv)^2 of any row in the iteration [3P].  Like Bullet's simulation islands, bodies that
   * share no constraint are solved independently: the arm island (motors, limits, arm-table contacts) and the
   * free-body island (free body vs static geometry) each stop on their own residual; a finger<->free-body contact
   * merges them into one island. */
  int coupled = 0;
  for (int k = 0; k < CS.n; ++k) if (rows[k].has_fb && rows[k].arm_body >= 0) coupled = 1;
  int arm_on = 1, fb_on = 1, it_done_ = 0;
  for (int it = 0; it < C->solver_iters; ++it) {
    real res_a = RC(0), res_f = RC(0);
    it_done_ = it + 1;
#define PIH_RES(acc_, dl_, diag_) do { real rv_ = (dl_) * (diag_); rv_ = rv_ * rv_; if (rv_ > acc_) acc_ = rv_; } while (0)
    /* Accumulated-impulse projection of one row (btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric [3P]:
     * sum = applied + delta, clamped to [lower, upper]).  Stated on the delta: delta is clamped to
     * [lower - applied, upper - applied] and then accumulated - the same projection, with the shortest dependent chain
     * (delta -> min/max -> velocity update); applied may sit one rounding off a bound, which the next sweep corrects. */
#define PIH_PROJECT(dl_, lam_, lo_, hi_) do { real dlo_ = (lo_) - (lam_), dhi_ = (hi_) - (lam_); \
      dl_ = ((dl_) < dlo_) ? dlo_ : (dl_); dl_ = ((dl_) > dhi_) ? dhi_ : (dl_); lam_ = (lam_) + (dl_); } while (0)
#define PIH_PROJECT_LO(dl_, lam_) do { real dlo_ = -(lam_); dl_ = ((dl_) < dlo_) ? dlo_ : (dl_); lam_ = (lam_) + (dl_); } while (0)
    if (arm_on) {
      for (int j = 0; j < nb; ++j) {
        real dl = FMA(-dva[j], m_dinv[j], m_b[j]);
        PIH_PROJECT(dl, m_lam[j], -m_max[j], m_max[j]);
        PIH_RES(res_a, dl, Minv[j][j]);
        for (int i = 0; i < nb; ++i) dva[i] = FMA(Minv[i][j], dl, dva[i]);
      }
      for (int j = 0; j < nb; ++j)
        for (int side = 0; side < 2; ++side) {
          if (!l_on[side][j]) continue;
          real sg = side ? RC(-1) : RC(1);
          real dl = FMA(-(sg * dva[j]), m_dinv[j], l_b[side][j]);
          PIH_PROJECT_LO(dl, l_lam[side][j]);
          dl = dl * sg;
          PIH_RES(res_a, dl, Minv[j][j]);
          for (int i = 0; i < nb; ++i) dva[i] = FMA(Minv[i][j], dl, dva[i]);
        }
    }
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < CS.n; ++k) {
        crow* r = &rows[k];
        const int is_fb = r->has_fb && r->arm_body < 0;
        if (!coupled && !(is_fb ? fb_on : arm_on)) continue;
        for (int d = (pass ? 1 : 0); d < (pass ? 3 : 1); ++d) {
          real rel = RC(0);
          if (r->has_fb) rel = FN(v3_dot)(r->dir[d], dvl) + FN(v3_dot)(r->jang[d], dvw);
          if (r->arm_body >= 0) {
            /* J_arm . dva as two interleaved partial sums (even / odd joints) */
            real s0 = RC(0), s1 = RC(0);
            for (int j = 0; j < nb; j += 2) s0 = FMA(r->ja[d][j], dva[j], s0);
            for (int j = 1; j < nb; j += 2) s1 = FMA(r->ja[d][j], dva[j], s1);
            rel = r->has_fb ? rel + (s0 + s1) : (s0 + s1);
          }
          real dl = FMA(-rel, r->dinv[d], r->b[d]);
          if (r->mu < RC(0)) {                     /* grasp constraint: bilateral, |impulse| <= maxForce dt */
            const real wl = RC(PIH_GRASP_MAX_FORCE) * dt;
            PIH_PROJECT(dl, r->lam[d], -wl, wl);
          } else if (d == 0) PIH_PROJECT_LO(dl, r->lam[d]);
          else {
            real lim = r->mu * r->lam[0];
            PIH_PROJECT(dl, r->lam[d], -lim, lim);
          }
          if (is_fb) PIH_RES(res_f, dl, r->diag[d]); else PIH_RES(res_a, dl, r->diag[d]);
          if (r->has_fb) { dvl = FN(v3_axpy)(dl * im, r->dir[d], dvl); dvw = FN(v3_axpy)(dl, r->mang[d], dvw); }
          if (r->arm_body >= 0) for (int j = 0; j < nb; ++j) dva[j] = FMA(r->ma[d][j], dl, dva[j]);
        }
      }
#undef PIH_RES
#undef PIH_PROJECT
#undef PIH_PROJECT_LO
    pih_dbg_pgs_iters += 1;
    if (coupled) {
      if (((res_a > res_f) ? res_a : res_f) <= C->solver_residual) break;
    } else {
      if (res_a <= C->solver_residual) arm_on = 0;
      if (res_f <= C->solver_residual) fb_on = 0;
      if (!arm_on && !fb_on) break;
    }
  }
  pih_dbg_pgs_steps += 1;
  if (pih_dbg_capture) {
    pih_capture* Q = pih_dbg_capture;
    memset(Q, 0, sizeof(*Q));
    const int nv = nb + 6;
    Q->nv = nv;
    int R = 0;
    for (int j = 0; j < nb; ++j) {             /* motors: v_j = v*_target, |lambda| <= maxForce dt */
      real qt = qstar[j];
      qt = (qt < M->lo[j]) ? M->lo[j] : ((qt > M->hi[j]) ? M->hi[j] : qt);
      Q->J[R][j] = 1.0; Q->tv[R] = (double)(kpc[j] * (qt - S->q[j]) * idt); Q->lo[R] = -(double)m_max[j]; Q->hi[R] = (double)m_max[j];
      Q->lam[R] = (double)m_lam[j]; R++;
    }
    for (int j = 0; j < nb; ++j)
      for (int side = 0; side < 2; ++side) {
        if (!l_on[side][j]) continue;
        real glo = S->q[j] - M->lo[j], ghi = M->hi[j] - S->q[j];
        real g_ = side ? ghi : glo;
        Q->J[R][j] = side ? -1.0 : 1.0;
        Q->tv[R] = (double)((g_ > RC(0)) ? -(g_ * idt) : -(RC(PIH_ERP) * g_ * idt));
        Q->lo[R] = 0.0; Q->hi[R] = 1e300; Q->lam[R] = (double)l_lam[side][j]; R++;
      }
    int nrow0[PIH_MAXC];
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < CS.n; ++k) {
        crow* r = &rows[k];
        for (int d = (pass ? 1 : 0); d < (pass ? 3 : 1); ++d) {
          if (r->arm_body >= 0) for (int j = 0; j < nb; ++j) Q->J[R][j] = (double)r->ja[d][j];
          if (r->has_fb) {
            Q->J[R][nb + 0] = (double)r->dir[d].x; Q->J[R][nb + 1] = (double)r->dir[d].y; Q->J[R][nb + 2] = (double)r->dir[d].z;
            Q->J[R][nb + 3] = (double)r->jang[d].x; Q->J[R][nb + 4] = (double)r->jang[d].y; Q->J[R][nb + 5] = (double)r->jang[d].z;
          }
          Q->tv[R] = (double)(r->b[d] * r->diag[d]) + 0.0;   /* b = (tv - J v*) / diag  ->  tv = b diag + J v* (added below) */
          if (r->mu < RC(0)) { Q->lo[R] = -(double)(RC(PIH_GRASP_MAX_FORCE) * dt); Q->hi[R] = -Q->lo[R]; }
          else if (d == 0) { Q->lo[R] = 0.0; Q->hi[R] = 1e300; nrow0[k] = R; }
          else { Q->kind[R] = 1; Q->lo[R] = (double)nrow0[k]; Q->hi[R] = (double)r->mu; }
          Q->lam[R] = (double)r->lam[d];
          R++;
        }
      }
    Q->nrows = R;
    for (int i = 0; i < nb; ++i) {
      Q->vstar[i] = (double)vs[i]; Q->vplus[i] = (double)(vs[i] + dva[i]);
      for (int j = 0; j < nb; ++j) Q->Minv[i][j] = (double)Minv[i][j];
    }
    real fvv[3] = {fv.x, fv.y, fv.z}, fww[3] = {fw.x, fw.y, fw.z}, dl3[3] = {dvl.x, dvl.y, dvl.z}, dw3[3] = {dvw.x, dvw.y, dvw.z};
    for (int k = 0; k < 3; ++k) {
      Q->vstar[nb + k] = (double)fvv[k]; Q->vstar[nb + 3 + k] = (double)fww[k];
      Q->vplus[nb + k] = (double)(fvv[k] + dl3[k]); Q->vplus[nb + 3 + k] = (double)(fww[k] + dw3[k]);
      Q->Minv[nb + k][nb + k] = (double)im;
      for (int c2 = 0; c2 < 3; ++c2) Q->Minv[nb + 3 + k][nb + 3 + c2] = (double)Iinv.m[k][c2];
    }
    /* contact / constraint rows (the last 3 per contact) stored b = (tv - J v*) / diag: recover tv */
    for (int r2 = R - 3 * CS.n; r2 < R; ++r2) {
      double jv = 0.0;
      for (int c2 = 0; c2 < nv; ++c2) jv += Q->J[r2][c2] * Q->vstar[c2];
      Q->tv[r2] += jv;
    }
    Q->iters = it_done_;
  }
  /* ---- integrate (semi-implicit Euler) ---- */
  for (int j = 0; j < nb; ++j) {
    real v = vs[j] + dva[j];
    v = (v > RC(PIH_MAX_COORD_VEL)) ? RC(PIH_MAX_COORD_VEL) : ((v < RC(-PIH_MAX_COORD_VEL)) ? RC(-PIH_MAX_COORD_VEL) : v);
    S->qd[j] = v;
    S->q[j] = FMA(dt, S->qd[j], S->q[j]);
  }
  fv = FN(clamp_norm)(FN(v3_add)(fv, dvl));
  fw = FN(clamp_norm)(FN(v3_add)(fw, dvw));
  S->pvel[0] = fv.x; S->pvel[1] = fv.y; S->pvel[2] = fv.z;
  S->pang[0] = fw.x; S->pang[1] = fw.y; S->pang[2] = fw.z;
  S->ppos[0] = FMA(dt, fv.x, S->ppos[0]); S->ppos[1] = FMA(dt, fv.y, S->ppos[1]); S->ppos[2] = FMA(dt, fv.z, S->ppos[2]);
  {
    /* orientation: q <- exp(0.5 dt w) (x) q, renormalised (btTransformUtil::integrateTransform [3P]) */
    real w2 = FN(v3_dot)(fw, fw);
    real wn = SQRT(w2);
    real ha = RC(0.5) * dt * wn;
    real sh, ch, k;
    FN(det_sincos)(ha, &sh, &ch);
    if (wn > RC(1e-6)) k = sh / wn; else k = RC(0.5) * dt;
    real dq[4] = {fw.x * k, fw.y * k, fw.z * k, ch}, nq[4];
    FN(quat_mul)(dq, S->pquat, nq);
    real n2 = FMA(nq[3], nq[3], FMA(nq[2], nq[2], FMA(nq[1], nq[1], nq[0] * nq[0])));
    real in = RC(1) / SQRT(n2);
    for (int i = 0; i < 4; ++i) S->pquat[i] = nq[i] * in;
  }
  S->step += 1;
  /* ---- fault guard (NEW): non-finite state -> re-initialise the env (new episode), report done
   * with reward 0, flag PIH_FLAG_FAULT; without auto_reset the env then stays frozen at the fresh state until reset() */
  {
    int finite = 1;
    for (int j = 0; j < nb; ++j) finite = finite && isfinite(S->q[j]) && isfinite(S->qd[j]);
    for (int k = 0; k < 3; ++k) finite = finite && isfinite(S->ppos[k]) && isfinite(S->pvel[k]) && isfinite(S->pang[k]);
    for (int k = 0; k < 4; ++k) finite = finite && isfinite(S->pquat[k]);
    if (!finite) {
      S->episode += 1;
      FN(reset_env)(M, C, S, gid);
      S->flags = PIH_FLAG_FAULT | (C->auto_reset ? 0 : PIH_FLAG_DONE);
      *reward = RC(0);
      *done = 1;
      FN(observe)(M, C, S, off, obs);
      return CS.n;
    }
  }
  /* ---- get_info: observation, reward, done (envs/peg_in_hole.py:13,114-117; time limit NEW) ---- */
  real rew = RC(0);
  int success = 0;
  fkres F2;
  v3 pee2; m3 Ree2;
  FN(fk)(M, S->q, nb, &F2);
  FN(ee_pose)(M, &F2, &pee2, &Ree2);
  if (C->task != PIH_TASK_RANDOM_FLY) {
    v3 d = FN(v3_make)(S->ppos[0] - RC(PIH_HOLE_X), S->ppos[1] - RC(PIH_HOLE_Y), S->ppos[2] - RC(PIH_HOLE_Z));
    success = FN(v3_dot)(d, d) < RC(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
    rew = success ? RC(1) : RC(0);
  } else {
    /* random-fly (NEW): success = object within 5 cm of the end effector; episode also ends on table impact */
    v3 d = FN(v3_make)(S->ppos[0] - pee2.x, S->ppos[1] - pee2.y, S->ppos[2] - pee2.z);
    success = FN(v3_dot)(d, d) < RC(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
    rew = success ? RC(1) : RC(0);
  }
  int dn = success || (C->max_steps > 0 && S->step >= C->max_steps);
  if (C->task == PIH_TASK_RANDOM_FLY && (pairmask & PIH_PAIR_PEG_TABLE)) dn = 1;
  /* scripted task: the episode ends when the FSM reaches state 9 (peg_in_hole.py:203-204); the reward is the success
   * test evaluated then (peg_in_hole.py:114-117) */
  if (C->task == PIH_TASK_SCRIPTED) dn = fsm_done || (C->max_steps > 0 && S->step >= C->max_steps);
  S->flags = (dn ? PIH_FLAG_DONE : 0) | (success ? PIH_FLAG_SUCCESS : 0) | (pairmask << PIH_PAIR_SHIFT);
  *reward = rew;
  *done = (uint8_t)dn;
  if (dn && C->auto_reset) {
    /* NEW: per-env auto-reset; the observation returned is the first one of the new episode */
    int keep = S->flags & ~PIH_FLAG_DONE;
    S->episode += 1;
    FN(reset_env)(M, C, S, gid);
    S->flags = (keep & ~0xff) | (success ? PIH_FLAG_PREV_SUCCESS : 0); /* contact mask of the terminal step stays visible; done/success cleared */
    FN(observe)(M, C, S, off, obs);
  } else {
    obs[0] = (nb > M->narm) ? S->q[M->narm] : RC(0);
    obs[1] = (nb > M->narm + 1) ? S->q[M->narm + 1] : RC(0);
    obs[2] = pee2.x + off[0]; obs[3] = pee2.y + off[1]; obs[4] = pee2.z + off[2];
  }
  return CS.n;
}

#include "pih_pipe_impl.h"

#undef real
#undef v3
#undef m3
#undef sv
#undef model
#undef envstate
#undef simcfg
#undef fkres
#undef abares
#undef contact
#undef contactset
#undef crow
#undef scene
#undef FMA
#undef SQRT
#undef RC
#undef ABS
#undef FN
#undef CAT
#undef CAT_
#undef NBM
