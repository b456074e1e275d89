// pih_pipe.cuh - the reference's 25-link flexible pipe as the peg: one WARP per sub-environment.
//
// Reference: peg_in_hole_gym/envs/assets/urdf/pipe.urdf (25 links, 1 fixed + 23 continuous joints), loaded with
// globalScaling 0.01 and random bends at peg_in_hole.py:239-245; grasp link chosen at :266, its pose read at :58 and the
// success test on it at :114-117.  PyBullet steps it as a floating-base btMultiBody [3P]: Featherstone ABA, unit-impulse
// responses from the same factorisation, a default velocity motor (target 0, max impulse 1) on every movable joint, and
// sequential impulses over motors, limits and contacts.
//
// Mapping.  A step of this task is two launches:
//   k_pipe_arm   thread per env: the arm half that the rigid-peg kernels already have - FK, apply_action (vel_constraint,
//                DLS IK or the scripted state machine), ABA, M^-1 - handed over through a 1 KB per-env scratch block;
//   k_pipe_step  warp per env: the 29 degrees of freedom of the chain (6 base + 23 joints) live on the lanes of the warp.
//                Lane i owns dof i: its unconstrained velocity, its column of M^-1 (its own unit-impulse response, 29
//                registers), its motor row.  Lane s owns collision segment s (25 cylinders).  Serial recurrences over the
//                chain (FK, the articulated-inertia pass) are run by all lanes redundantly from shared memory; sums over
//                degrees of freedom are xor-butterflies (fixed order, restated by the oracle: FN(bfly32)).  The step is
//                the latency of one warp's dependent instruction stream, so what matters is how many warps an SM holds:
//                body loops are rolled (one z-joint and one x-joint body per pass, small code), M^-1 lives in shared
//                memory (3.6 KB), and the constraint rows (Jacobian + M^-1 J^T, 384 B each, read once per solver sweep)
//                live in a per-env global scratch block that stays in L1/L2 - 13 KB of shared memory and <= 128
//                registers per warp-env, 16 resident envs per SM.
// Bit-comparable with oracle/pih_pipe_impl.h (float build): same chains of explicit FMAs, file compiled with -fmad=false.
#pragma once
#include "pih_step.cuh"

namespace pih {

// ---- per-env scratch block written by k_pipe_arm, read by k_pipe_step (floats, AoS: the reader is one warp per env)
#define PIH_PA_WORDS 256
#define PA_QSTAR 0
#define PA_KP 9
#define PA_MAXF 18
#define PA_VS 27
#define PA_MINV 36          // [i * 9 + j]
#define PA_FSM_DONE 117
#define PA_FSM_STATE 118
#define PA_FSM_T 119
#define PA_MTGT 120         // 7
#define PA_GRASP_ANG 127
#define PA_ZJ 128           // joint axes in the env frame [j * 3 + k]
#define PA_PW 155           // joint origins
#define PA_BOXC 182         // finger box centres [b * 3 + k]
#define PA_BOXR 188         // finger box rotations [b * 9 + 3 r + c]
#define PA_SPH 206          // arm collision sphere centres [q * 3 + k]
#define PA_ACTIVE 227       // 1: the env is stepped in this launch, 0: frozen (done, waiting for reset)
#define PA_IK_N 228         // statistics: IK iterations of this step

constexpr int PNB = PIH_PIPE_NBODY, PND = PIH_PIPE_NDOF, PNSEG = PIH_PIPE_NSEG;
// resident warp-envs per SM the step kernel is compiled for (register budget 65536 / (32 * this))
#ifndef PIH_PIPE_MIN_BLOCKS
#define PIH_PIPE_MIN_BLOCKS 16
#endif
// warp-envs per CTA of the step kernel
#ifndef PIH_PIPE_WARPS
#define PIH_PIPE_WARPS 1
#endif

// ---- shared-memory layout of k_pipe_step (floats)
constexpr int SM_SW = 0;                                  // state words [128]
constexpr int SM_PA = SM_SW + 4 * PIH_PIPE_STATE_F4;      // arm scratch [256]
constexpr int SM_SN = SM_PA + PIH_PA_WORDS;               // sin / cos of the joint angles [24] each
constexpr int SM_CS = SM_SN + PNB;
constexpr int SM_RW = SM_CS + PNB;                        // body rotations [24][9]
constexpr int SM_PW = SM_RW + PNB * 9;                    // body origins [24][3]
constexpr int SM_CB = SM_PW + PNB * 3;                    // velocity-product accelerations [24][6]
constexpr int SM_PB = SM_CB + PNB * 6;                    // bias forces [24][6]
constexpr int SM_U = SM_PB + PNB * 6;                     // U = I^A S [24][6]
constexpr int SM_UD = SM_U + PNB * 6;                     // U / D [24][6]
constexpr int SM_DINV = SM_UD + PNB * 6;                  // [24]
constexpr int SM_UU = SM_DINV + PNB;                      // u [24]
constexpr int SM_CON = SM_UU + PNB;                       // contacts [PIH_PIPE_MAXC][12]: p, n, dist, mu, pair, seg, arm_body, arm slot
constexpr int SM_ROWS = SM_CON + PIH_PIPE_MAXC * 12;      // row scalars [3 * MAXC][4]: dinv, diag, b, lam
constexpr int SM_MCOL = SM_ROWS + 3 * PIH_PIPE_MAXC * 4;  // M^-1 of the chain: [c][lane] = velocity change of dof c for a unit impulse on dof `lane`
constexpr int SM_JTMP = SM_MCOL + PND * 32;               // the Jacobian row being built [32] (+ arm part [16])
constexpr int SM_TOTAL = SM_JTMP + 48;
constexpr int SM_UH = SM_CON;                             // unit-impulse pass: uh[b][lane] (dead before the contacts are written)
static_assert(PNB * 32 <= PIH_PIPE_MAXC * 12 + 3 * PIH_PIPE_MAXC * 4, "uh aliases the contact + row-scalar region");
constexpr int kPipeSmemBytes = SM_TOTAL * 4;
// constraint rows in the per-env global scratch block: [row][96] floats = J[32] | M^-1 J^T[32] | arm J[16] | arm M^-1 J^T[16]
constexpr int kRowWords = 96;
constexpr int kRowBufWords = 3 * PIH_PIPE_MAXC * kRowWords;

__host__ __device__ constexpr int pipe_axis(int b) { return (b & 1) ? 2 : 0; }
PIH_D float pipe_len(int b) { return (b == 1) ? (float)(PIH_PIPE_LINK1_Y + PIH_PIPE_PITCH) : (float)PIH_PIPE_PITCH; }
__host__ __device__ constexpr int seg_body(int s) { return s < 2 ? 0 : s - 1; }
PIH_D float seg_cy(int s) { return (s == 1) ? (float)(PIH_PIPE_LINK1_Y + PIH_PIPE_HL) : (float)PIH_PIPE_HL; }
PIH_D float seg_mu(int s) { return (s < 2 || s > 22) ? (float)PIH_PIPE_MU_END : (float)PIH_MU_DEFAULT; }

// chain constants (oracle: pipemodel_init; doubles folded at compile time, rounded to float once)
struct PipeModel { float Ix, Iy, m, I0x, I0y, I0z, m0, h0y; };
PIH_D PipeModel pipe_model() {
  constexpr double lx = 2 * (PIH_PIPE_R + 0.001), ly = 2 * (PIH_PIPE_HL + 0.001);
  constexpr double m = PIH_PIPE_LINK_MASS, m0 = PIH_PIPE_LINK0_MASS;
  constexpr double ix = m / 12.0 * (ly * ly + lx * lx), iy = m / 12.0 * (lx * lx + lx * lx);
  constexpr double ix0 = m0 / 12.0 * (ly * ly + lx * lx), iy0 = m0 / 12.0 * (lx * lx + lx * lx);
  constexpr double c1 = PIH_PIPE_LINK1_Y + PIH_PIPE_LINK1_COM_Y;
  PipeModel P;
  P.Ix = (float)ix; P.Iy = (float)iy; P.m = (float)m;
  P.I0x = (float)(ix0 + ix + m * c1 * c1); P.I0y = (float)(iy0 + iy); P.I0z = (float)(ix0 + ix + m * c1 * c1);
  P.m0 = (float)(m0 + m); P.h0y = (float)(m * c1);
  return P;
}

// sum over the 32 lanes in xor-butterfly order (oracle: bfly32); every lane gets the same bits
PIH_D float bfly32(float p) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, off);
  return p;
}

template <int AX> PIH_D v3 pipe_rotT(float s, float c, v3 u) {
  if (AX == 2) return mk(fmaf(s, u.y, c * u.x), fmaf(c, u.y, -(s * u.x)), u.z);
  return mk(u.x, fmaf(s, u.z, c * u.y), fmaf(c, u.z, -(s * u.y)));
}
template <int AX> PIH_D v3 pipe_rot(float s, float c, v3 u) {
  if (AX == 2) return mk(fmaf(-s, u.y, c * u.x), fmaf(c, u.y, s * u.x), u.z);
  return mk(u.x, fmaf(-s, u.z, c * u.y), fmaf(c, u.z, s * u.y));
}
template <int AX> PIH_D sv pipe_motion_to_child(float s, float c, float L, sv m) {
  v3 t = mk(fmaf(-m.a.z, L, m.l.x), m.l.y, fmaf(m.a.x, L, m.l.z));
  sv o;
  o.a = pipe_rotT<AX>(s, c, m.a);
  o.l = pipe_rotT<AX>(s, c, t);
  return o;
}
template <int AX> PIH_D sv pipe_force_to_parent(float s, float c, float L, sv f) {
  sv o;
  o.l = pipe_rot<AX>(s, c, f.l);
  v3 n = pipe_rot<AX>(s, c, f.a);
  o.a = mk(fmaf(L, o.l.z, n.x), n.y, fmaf(-L, o.l.x, n.z));
  return o;
}
template <int AX> PIH_D float vget(v3 v) { return AX == 0 ? v.x : (AX == 1 ? v.y : v.z); }
template <int AX> PIH_D void vaddk(v3& v, float d) { if (AX == 0) v.x = v.x + d; else if (AX == 1) v.y = v.y + d; else v.z = v.z + d; }
PIH_D float vgetr(v3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

PIH_D v3 lds3(const float* p) { return mk(p[0], p[1], p[2]); }
PIH_D m3 ldsm(const float* p) {
  m3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i / 3][i % 3] = p[i];
  return R;
}
PIH_D void sts3(float* p, v3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// 6x6 LDL^T of the base's articulated inertia (oracle: ldl6_factor / ldl6_solve)
struct Ldl6 { float L[6][6], invd[6]; };
PIH_D void ldl6_factor(const m3& A, const m3& B, const m3& Cm, Ldl6& F) {
  float Kl[6][6], W[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      if (i < 3) Kl[i][j] = A.m[i][j];
      else if (j < 3) Kl[i][j] = B.m[j][i - 3];
      else Kl[i][j] = Cm.m[i - 3][j - 3];
    }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) {
      float s = Kl[i][j];
#pragma unroll
      for (int kk = 0; kk < j; ++kk) s = fmaf(-W[i][kk], F.L[j][kk], s);
      W[i][j] = s;
      F.L[i][j] = s * F.invd[j];
    }
    float d = Kl[i][i];
#pragma unroll
    for (int kk = 0; kk < i; ++kk) d = fmaf(-W[i][kk], F.L[i][kk], d);
    F.invd[i] = 1.0f / d;
  }
}
PIH_D void ldl6_solve(const Ldl6& F, const float* rhs, float* x) {
  float y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float s = rhs[i];
#pragma unroll
    for (int kk = 0; kk < i; ++kk) s = fmaf(-F.L[i][kk], y[kk], s);
    y[i] = s;
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float s = y[i] * F.invd[i];
#pragma unroll
    for (int kk = i + 1; kk < 6; ++kk) s = fmaf(-F.L[kk][i], y[kk], s);
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// pipe block of the state as the warp sees it (every lane holds the base; joint b lives on lane 5 + b)
struct PipeBase { float pos[3], quat[4], vel[3], ang[3]; int grasp_idx; };

// Chain passes are ROLLED: bodies alternate z joints (odd b) and x joints (even b), so a pass is a loop over the pairs
// (b, b + 1) with one instantiation of the body per axis - twenty times less code than the unrolled chain, which is what a
// lone warp per scheduler needs (instruction fetch was 37 % of the stall samples of the unrolled kernel).
template <class F> PIH_D void chain_forward(F&& body) {      // b = 1 .. 23
#pragma unroll 1
  for (int b = 1; b < PNB; b += 2) {
    body(std::integral_constant<int, 2>{}, b);
    if (b + 1 < PNB) body(std::integral_constant<int, 0>{}, b + 1);
  }
}
template <class F> PIH_D void chain_backward(F&& body) {     // b = 23 .. 1
#pragma unroll 1
  for (int b = PNB - 1; b >= 1; b -= 2) {
    body(std::integral_constant<int, 2>{}, b);
    if (b - 1 >= 1) body(std::integral_constant<int, 0>{}, b - 1);
  }
}
static_assert((PNB - 1) % 2 == 1, "the last body of the chain hangs on a z joint");

// FK of the chain into shared memory: Rw[b], pw[b], sn[b], cs[b].  All lanes run the recurrence (identical values).
PIH_D void pipe_fk(float* sm, const PipeBase& B, int lane, float qj_lane) {
  {
    // lane 5 + b owns joint b: its sine / cosine go to shared memory for the serial pass
    float s, c;
    det_sincos(qj_lane, s, c);
    const int b = lane - 5;
    if (b >= 1 && b < PNB) { sm[SM_SN + b] = s; sm[SM_CS + b] = c; }
  }
  __syncwarp();
  m3 R = quat_to_m3(B.quat);
  v3 p = mk(B.pos[0], B.pos[1], B.pos[2]);
#pragma unroll
  for (int i = 0; i < 9; ++i) sm[SM_RW + i] = R.m[i / 3][i % 3];
  sts3(&sm[SM_PW], p);
  chain_forward([&](auto ax_, int b) {
    constexpr int ax = decltype(ax_)::value;
    const float L = pipe_len(b);
    p = mk(fmaf(L, R.m[0][1], p.x), fmaf(L, R.m[1][1], p.y), fmaf(L, R.m[2][1], p.z));
    const float s = sm[SM_SN + b], c = sm[SM_CS + b];
    m3 N;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (ax == 2) {
        N.m[i][0] = fmaf(s, R.m[i][1], c * R.m[i][0]);
        N.m[i][1] = fmaf(c, R.m[i][1], -(s * R.m[i][0]));
        N.m[i][2] = R.m[i][2];
      } else {
        N.m[i][0] = R.m[i][0];
        N.m[i][1] = fmaf(s, R.m[i][2], c * R.m[i][1]);
        N.m[i][2] = fmaf(c, R.m[i][2], -(s * R.m[i][1]));
      }
    }
    R = N;
#pragma unroll
    for (int i = 0; i < 9; ++i) sm[SM_RW + 9 * b + i] = R.m[i / 3][i % 3];
    sts3(&sm[SM_PW + 3 * b], p);
  });
  __syncwarp();
}

// getLinkState(objectUid, grasp_joint_idx)[0:2] (oracle: pipe_grasp_pose)
PIH_D void pipe_grasp_pose(const float* sm, int grasp_idx, float (&pos)[3], float (&quat)[4]) {
  if (grasp_idx == 0) {
    const float cy = (float)(PIH_PIPE_LINK1_Y + PIH_PIPE_LINK1_COM_Y);
    m3 R = ldsm(&sm[SM_RW]);
    pos[0] = fmaf(cy, R.m[0][1], sm[SM_PW]); pos[1] = fmaf(cy, R.m[1][1], sm[SM_PW + 1]); pos[2] = fmaf(cy, R.m[2][1], sm[SM_PW + 2]);
    m3_to_quat(R, quat);
  } else {
    m3 R = ldsm(&sm[SM_RW + 9 * (PNB - 1)]);
    pos[0] = sm[SM_PW + 3 * (PNB - 1)]; pos[1] = sm[SM_PW + 3 * (PNB - 1) + 1]; pos[2] = sm[SM_PW + 3 * (PNB - 1) + 2];
    m3_to_quat(R, quat);
  }
}

PIH_D m3 pipe_Rl(int axis, float s, float c) {
  m3 R;
  if (axis == 2) {
    R.m[0][0] = c; R.m[0][1] = -s; R.m[0][2] = 0.0f; R.m[1][0] = s; R.m[1][1] = c; R.m[1][2] = 0.0f; R.m[2][0] = 0.0f; R.m[2][1] = 0.0f; R.m[2][2] = 1.0f;
  } else {
    R.m[0][0] = 1.0f; R.m[0][1] = 0.0f; R.m[0][2] = 0.0f; R.m[1][0] = 0.0f; R.m[1][1] = c; R.m[1][2] = -s; R.m[2][0] = 0.0f; R.m[2][1] = s; R.m[2][2] = c;
  }
  return R;
}

// ABA of the floating chain (oracle: pipe_dynamics).  Serial passes by all lanes (rolled); results: lane i gets vstar of
// dof i, and shared memory gets M^-1: SM_MCOL[c][i] = velocity change of dof c for a unit impulse on dof i (lane i's own
// response pass).  The per-body U, U/D, 1/D, u stay in shared memory.
PIH_D void pipe_dynamics(float* sm, const PipeBase& B, int lane, float qdj_lane, float dt, float g, float& vstar_lane) {
  const PipeModel PM = pipe_model();
  // joint rates to shared memory (SM_UU until the inward pass overwrites it with u)
  {
    const int b = lane - 5;
    if (b >= 1 && b < PNB) sm[SM_UU + b] = qdj_lane;
  }
  __syncwarp();
  const m3 R0 = ldsm(&sm[SM_RW]);
  // (B) velocities, velocity-product accelerations, bias forces
  sv vel;
  vel.a = mTv(R0, mk(B.ang[0], B.ang[1], B.ang[2]));
  vel.l = mTv(R0, mk(B.vel[0], B.vel[1], B.vel[2]));
  const sv vel0 = vel;
  sv pA0;
  {
    v3 h = mk(0.0f, PM.h0y, 0.0f);
    v3 w = vel.a, v = vel.l;
    v3 La = vadd(mk(PM.I0x * w.x, PM.I0y * w.y, PM.I0z * w.z), vcross(h, v));
    v3 hw = vcross(h, w);
    v3 Ll = mk(fmaf(PM.m0, v.x, -hw.x), fmaf(PM.m0, v.y, -hw.y), fmaf(PM.m0, v.z, -hw.z));
    pA0.a = vadd(vcross(w, La), vcross(v, Ll));
    pA0.l = vcross(w, Ll);
  }
  chain_forward([&](auto ax_, int b) {
    constexpr int ax = decltype(ax_)::value;
    sv vp = pipe_motion_to_child<ax>(sm[SM_SN + b], sm[SM_CS + b], pipe_len(b), vel);
    const float w = sm[SM_UU + b];
    sv cb;
    if (ax == 2) { cb.a = mk(vp.a.y * w, -(vp.a.x * w), 0.0f); cb.l = mk(vp.l.y * w, -(vp.l.x * w), 0.0f); }
    else { cb.a = mk(0.0f, vp.a.z * w, -(vp.a.y * w)); cb.l = mk(0.0f, vp.l.z * w, -(vp.l.y * w)); }
    sv v = vp;
    vaddk<ax>(v.a, w);
    vel = v;
    v3 Iw = mk(PM.Ix * v.a.x, PM.Iy * v.a.y, PM.Ix * v.a.z);
    v3 mv_ = mk(PM.m * v.l.x, PM.m * v.l.y, PM.m * v.l.z);
    v3 pa = vcross(v.a, Iw), pl = vcross(v.a, mv_);
    sts3(&sm[SM_CB + 6 * b], cb.a); sts3(&sm[SM_CB + 6 * b + 3], cb.l);
    sts3(&sm[SM_PB + 6 * b], pa); sts3(&sm[SM_PB + 6 * b + 3], pl);
  });
  // (C) inward pass.  The articulated inertia / bias handed up by the child is carried in registers; body 23 (the tip)
  //     has none: `have` stays false for it.
  m3 accA, accB, accC;
  sv accp;
  accp.a = mk(0.0f, 0.0f, 0.0f); accp.l = accp.a;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { accA.m[r][c] = 0.0f; accB.m[r][c] = 0.0f; accC.m[r][c] = 0.0f; }
  bool have = false;
  chain_backward([&](auto ax_, int b) {
    constexpr int ax = decltype(ax_)::value;
    m3 IA, IB, IC;
    sv pAb;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { IA.m[r][c] = 0.0f; IB.m[r][c] = 0.0f; IC.m[r][c] = (r == c) ? PM.m : 0.0f; }
    IA.m[0][0] = PM.Ix; IA.m[1][1] = PM.Iy; IA.m[2][2] = PM.Ix;
    pAb.a = lds3(&sm[SM_PB + 6 * b]); pAb.l = lds3(&sm[SM_PB + 6 * b + 3]);
    if (have) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { IA.m[r][c] = IA.m[r][c] + accA.m[r][c]; IB.m[r][c] = IB.m[r][c] + accB.m[r][c]; IC.m[r][c] = IC.m[r][c] + accC.m[r][c]; }
      pAb.a = vadd(pAb.a, accp.a); pAb.l = vadd(pAb.l, accp.l);
    }
    sv U;
    U.a = mcol(IA, ax); U.l = mrow(IB, ax);
    const float Dd = IA.m[ax][ax];
    const float u = -vget<ax>(pAb.a);
    const float Dinv = 1.0f / Dd;
    sv UD;
    UD.a = vscale(U.a, Dinv); UD.l = vscale(U.l, Dinv);
    sts3(&sm[SM_U + 6 * b], U.a); sts3(&sm[SM_U + 6 * b + 3], U.l);
    sts3(&sm[SM_UD + 6 * b], UD.a); sts3(&sm[SM_UD + 6 * b + 3], UD.l);
    sm[SM_DINV + b] = Dinv; sm[SM_UU + b] = u;
    float uda[3] = {UD.a.x, UD.a.y, UD.a.z}, udl[3] = {UD.l.x, UD.l.y, UD.l.z};
    float ua[3] = {U.a.x, U.a.y, U.a.z}, ul[3] = {U.l.x, U.l.y, U.l.z};
    m3 A1, B1, C1;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A1.m[r][c] = fmaf(-uda[r], ua[c], IA.m[r][c]);
        B1.m[r][c] = fmaf(-uda[r], ul[c], IB.m[r][c]);
        C1.m[r][c] = fmaf(-udl[r], ul[c], IC.m[r][c]);
      }
    sv cb;
    cb.a = lds3(&sm[SM_CB + 6 * b]); cb.l = lds3(&sm[SM_CB + 6 * b + 3]);
    sv pa;
    v3 t1 = vadd(mv(A1, cb.a), mv(B1, cb.l));
    v3 t2 = vadd(mTv(B1, cb.a), mv(C1, cb.l));
    pa.a = vaxpy(u, UD.a, vadd(pAb.a, t1));
    pa.l = vaxpy(u, UD.l, vadd(pAb.l, t2));
    const float sn = sm[SM_SN + b], cs = sm[SM_CS + b];
    const m3 Rl = pipe_Rl(ax, sn, cs);
    m3 Tm = mmul(Rl, A1); m3 Ar = mmulT(Tm, Rl);
    Tm = mmul(Rl, B1); m3 Br = mmulT(Tm, Rl);
    Tm = mmul(Rl, C1); m3 Cr = mmulT(Tm, Rl);
    v3 r = mk(0.0f, pipe_len(b), 0.0f);
    m3 Bp, Ap;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v3 cc = vcross(r, mcol(Cr, c));
      Bp.m[0][c] = Br.m[0][c] + cc.x; Bp.m[1][c] = Br.m[1][c] + cc.y; Bp.m[2][c] = Br.m[2][c] + cc.z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v3 cc = vcross(r, mrow(Br, c));
      Ap.m[0][c] = Ar.m[0][c] + cc.x; Ap.m[1][c] = Ar.m[1][c] + cc.y; Ap.m[2][c] = Ar.m[2][c] + cc.z;
    }
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      v3 cc = vcross(mrow(Bp, rr), r);
      Ap.m[rr][0] = Ap.m[rr][0] - cc.x; Ap.m[rr][1] = Ap.m[rr][1] - cc.y; Ap.m[rr][2] = Ap.m[rr][2] - cc.z;
    }
    accA = Ap; accB = Bp; accC = Cr;
    accp = pipe_force_to_parent<ax>(sn, cs, pipe_len(b), pa);
    have = true;
  });
  // (D) base
  m3 IA0, IB0, IC0;
  {
    const float hy = PM.h0y;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { IA0.m[r][c] = 0.0f; IB0.m[r][c] = 0.0f; IC0.m[r][c] = (r == c) ? PM.m0 : 0.0f; }
    IA0.m[0][0] = PM.I0x; IA0.m[1][1] = PM.I0y; IA0.m[2][2] = PM.I0z;
    // [h]x with h = (0, hy, 0)  (the negations of zero are written as the oracle writes them)
    IB0.m[0][1] = -0.0f; IB0.m[0][2] = hy; IB0.m[1][0] = 0.0f; IB0.m[1][2] = -0.0f; IB0.m[2][0] = -hy; IB0.m[2][1] = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { IA0.m[r][c] = IA0.m[r][c] + accA.m[r][c]; IB0.m[r][c] = IB0.m[r][c] + accB.m[r][c]; IC0.m[r][c] = IC0.m[r][c] + accC.m[r][c]; }
    pA0.a = vadd(pA0.a, accp.a); pA0.l = vadd(pA0.l, accp.l);
  }
  Ldl6 Fz;
  ldl6_factor(IA0, IB0, IC0, Fz);
  float a0[6];
  {
    float rhs[6] = {-pA0.a.x, -pA0.a.y, -pA0.a.z, -pA0.l.x, -pA0.l.y, -pA0.l.z};
    ldl6_solve(Fz, rhs, a0);
  }
  // (E) outward pass
  {
    float gb[3] = {-(g * R0.m[2][0]), -(g * R0.m[2][1]), -(g * R0.m[2][2])};
    float vsb[6];
    vsb[0] = fmaf(dt, a0[0], vel0.a.x); vsb[1] = fmaf(dt, a0[1], vel0.a.y); vsb[2] = fmaf(dt, a0[2], vel0.a.z);
    vsb[3] = fmaf(dt, a0[3] + gb[0], vel0.l.x); vsb[4] = fmaf(dt, a0[4] + gb[1], vel0.l.y); vsb[5] = fmaf(dt, a0[5] + gb[2], vel0.l.z);
    vstar_lane = 0.0f;
#pragma unroll
    for (int k = 0; k < 6; ++k) vstar_lane = (lane == k) ? vsb[k] : vstar_lane;
    sv acc;
    acc.a = mk(a0[0], a0[1], a0[2]); acc.l = mk(a0[3], a0[4], a0[5]);
    chain_forward([&](auto ax_, int b) {
      constexpr int ax = decltype(ax_)::value;
      sv a = pipe_motion_to_child<ax>(sm[SM_SN + b], sm[SM_CS + b], pipe_len(b), acc);
      a.a = vadd(a.a, lds3(&sm[SM_CB + 6 * b]));
      a.l = vadd(a.l, lds3(&sm[SM_CB + 6 * b + 3]));
      const float ua = vdot(lds3(&sm[SM_U + 6 * b]), a.a) + vdot(lds3(&sm[SM_U + 6 * b + 3]), a.l);
      const float qdd = sm[SM_DINV + b] * (sm[SM_UU + b] - ua);
      vaddk<ax>(a.a, qdd);
      acc = a;
      const float vsj = fmaf(dt, qdd, qdj_lane);       // (only lane 5 + b holds joint b's rate: the others discard this)
      vstar_lane = (lane == 5 + b) ? vsj : vstar_lane;
    });
  }
  // (F) unit-impulse response of dof `lane` (lanes 29..31: no impulse, all zeros) -> column `lane` of SM_MCOL
  {
    sv ph;
    ph.a = mk(0.0f, 0.0f, 0.0f); ph.l = ph.a;
    chain_backward([&](auto ax_, int b) {
      constexpr int ax = decltype(ax_)::value;
      const float sp = vget<ax>(ph.a);
      const float uh = ((5 + b == lane) ? 1.0f : 0.0f) - sp;
      sm[SM_UH + 32 * b + lane] = uh;
      sv pa;
      pa.a = vaxpy(uh, lds3(&sm[SM_UD + 6 * b]), ph.a);
      pa.l = vaxpy(uh, lds3(&sm[SM_UD + 6 * b + 3]), ph.l);
      // the oracle adds the transformed force to the parent's (zero) accumulator: 0 + x keeps x bit for bit
      sv fp = pipe_force_to_parent<ax>(sm[SM_SN + b], sm[SM_CS + b], pipe_len(b), pa);
      ph.a = vadd(mk(0.0f, 0.0f, 0.0f), fp.a);
      ph.l = vadd(mk(0.0f, 0.0f, 0.0f), fp.l);
    });
    float r6[6] = {ph.a.x, ph.a.y, ph.a.z, ph.l.x, ph.l.y, ph.l.z}, x6[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) r6[j] = ((j == lane) ? 1.0f : 0.0f) - r6[j];
    ldl6_solve(Fz, r6, x6);
#pragma unroll
    for (int j = 0; j < 6; ++j) sm[SM_MCOL + 32 * j + lane] = x6[j];
    sv acc;
    acc.a = mk(x6[0], x6[1], x6[2]); acc.l = mk(x6[3], x6[4], x6[5]);
    chain_forward([&](auto ax_, int b) {
      constexpr int ax = decltype(ax_)::value;
      sv a = pipe_motion_to_child<ax>(sm[SM_SN + b], sm[SM_CS + b], pipe_len(b), acc);
      const float ua = vdot(lds3(&sm[SM_U + 6 * b]), a.a) + vdot(lds3(&sm[SM_U + 6 * b + 3]), a.l);
      const float d = sm[SM_DINV + b] * (sm[SM_UH + 32 * b + lane] - ua);
      vaddk<ax>(a.a, d);
      acc = a;
      sm[SM_MCOL + 32 * (5 + b) + lane] = d;
    });
  }
  __syncwarp();
}

// ---------------------------------------------------------------- contacts
struct PContact { int pair, seg, arm_body; v3 p, n; float dist, mu; };

PIH_D void pcontact_push(PContact (&slot)[PIH_PIPE_SEG_MAXC], int& n, const PContact& c) {
  static_assert(PIH_PIPE_SEG_MAXC == 2, "two slots per segment");
  if (!(c.dist < (float)PIH_CONTACT_MARGIN)) return;
  if (n < 2) { if (n == 0) slot[0] = c; else slot[1] = c; n += 1; return; }
  const int w = (slot[1].dist > slot[0].dist) ? 1 : 0;
  if (w == 0) { if (c.dist < slot[0].dist) slot[0] = c; }
  else { if (c.dist < slot[1].dist) slot[1] = c; }
}

PIH_D float seg_sdf(const m3& R, v3 cw, v3 x, v3& n_env) {
  v3 l = mTv(R, vsub(x, cw));
  float rho = sqrtf(fmaf(l.z, l.z, l.x * l.x));
  float da = rho - (float)PIH_PIPE_R, db = fabsf(l.y) - (float)PIH_PIPE_HL;
  float na, nb_;
  float d = sd_rect(da, db, 1.0f, (l.y < 0.0f) ? -1.0f : 1.0f, na, nb_);
  float ir = (rho > 1e-12f) ? 1.0f / rho : 0.0f;
  v3 nl = mk(na * l.x * ir, nb_, na * l.z * ir);
  n_env = mv(R, nl);
  return d;
}

// candidates of segment s (oracle: pipe_collide_segment).  Arm geometry comes from the scratch block in shared memory.
template <int ARM> PIH_D void pipe_collide_segment(const float* sm, int s, PContact (&slot)[PIH_PIPE_SEG_MAXC], int& n) {
  const ArmTables& T = c_arm[ARM];
  const float* pa = sm + SM_PA;
  const int body = seg_body(s);
  const m3 R = ldsm(&sm[SM_RW + 9 * body]);
  const v3 pw = lds3(&sm[SM_PW + 3 * body]);
  const float cy = seg_cy(s);
  const v3 cw = mk(fmaf(cy, R.m[0][1], pw.x), fmaf(cy, R.m[1][1], pw.y), fmaf(cy, R.m[2][1], pw.z));
  const float mu_s = seg_mu(s);
  const float hl = (float)PIH_PIPE_HL, rad = (float)PIH_PIPE_R;
  PContact c;
  c.seg = s;
  n = 0;
  // (1) table plane
  if (cw.z - (float)PIH_TABLE_Z < 0.05f) {
    v3 ax = mcol(R, 1);
    float naz = ax.z;
    v3 nperp = mk(-(naz * ax.x), -(naz * ax.y), fmaf(-naz, ax.z, 1.0f));
    float l2 = vdot(nperp, nperp);
    c.pair = PIH_PAIR_PEG_TABLE; c.arm_body = -1; c.n = mk(0.0f, 0.0f, 1.0f); c.mu = mu_s * (float)PIH_MU_DEFAULT;
    if (l2 > 0.0025f) {
      float il = 1.0f / sqrtf(l2);
      v3 t = vscale(nperp, il);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float sgn = k ? -1.0f : 1.0f;
        v3 cc = vaxpy(sgn * hl, ax, cw);
        c.p = vaxpy(-rad, t, cc);
        c.dist = c.p.z - (float)PIH_TABLE_Z;
        pcontact_push(slot, n, c);
      }
    } else {
      float sgn = (naz > 0.0f) ? -1.0f : 1.0f;
      v3 cc = vaxpy(sgn * hl, ax, cw);
      v3 ex = mcol(R, 0), ez = mcol(R, 2);
      for (int k = 0; k < 4; ++k) {
        c.p = vaxpy((k & 1) ? -rad : rad, (k & 2) ? ez : ex, cc);
        c.dist = c.p.z - (float)PIH_TABLE_Z;
        pcontact_push(slot, n, c);
      }
    }
  }
  // (2) annular hole
  {
    v3 dh = vsub(cw, mk((float)PIH_HOLE_X, (float)PIH_HOLE_Y, (float)PIH_HOLE_Z));
    const float reach = (float)(PIH_PIPE_HL + PIH_PIPE_R + PIH_HOLE_ROUT + PIH_HOLE_HL + PIH_CONTACT_MARGIN);
    if (vdot(dh, dh) < reach * reach) {
      c.pair = PIH_PAIR_PEG_HOLE; c.arm_body = -1; c.mu = mu_s * (float)PIH_MU_DEFAULT;
      for (int k = 0; k < 16; ++k) {
        float cc, ss;
        circ8(k, cc, ss);
        float sy = (k < 8) ? 1.0f : -1.0f;
        v3 pl = mk(rad * cc, sy * hl, rad * ss);
        v3 pt = mv_add(R, pl, cw);
        v3 nn;
        c.dist = hole_sdf(pt, nn);
        c.p = pt; c.n = nn;
        pcontact_push(slot, n, c);
      }
      for (int k = 0; k < 32; ++k) {
        float rr = (k < 16) ? (float)PIH_HOLE_RIN : (float)PIH_HOLE_ROUT;
        float sx = (k & 8) ? (float)(-PIH_HOLE_HL) : (float)PIH_HOLE_HL;
        float cc, ss;
        circ8(k, cc, ss);
        v3 pt = mk((float)PIH_HOLE_X + sx, fmaf(rr, cc, (float)PIH_HOLE_Y), fmaf(rr, ss, (float)PIH_HOLE_Z));
        v3 nn;
        c.dist = seg_sdf(R, cw, pt, nn);
        c.p = pt; c.n = mk(-nn.x, -nn.y, -nn.z);
        pcontact_push(slot, n, c);
      }
    }
  }
  // (3) finger boxes
#pragma unroll
  for (int b = 0; b < ArmGeom<ARM>::NBOX; ++b) {
    const int abody = box_body_of<ARM>(b);
    v3 cb = lds3(&pa[PA_BOXC + 3 * b]);
    v3 dd = vsub(cw, cb);
    float reach = hl + rad + T.box_h[b][2] + T.box_h[b][0] + (float)PIH_CONTACT_MARGIN;
    if (!(vdot(dd, dd) < reach * reach)) continue;
    m3 Rb = ldsm(&pa[PA_BOXR + 9 * b]);
    for (int k = 0; k < 5; ++k) {
      v3 pl = mk(0.0f, (hl - rad) * 0.5f * (float)(k - 2), 0.0f);
      v3 pt = mv_add(R, pl, cw);
      v3 l = mTv(Rb, vsub(pt, cb));
      v3 nl;
      float d = sd_box(l, ld3(T.box_h[b]), nl);
      v3 nn = mv(Rb, nl);
      c.pair = b ? PIH_PAIR_PEG_FINGER2 : PIH_PAIR_PEG_FINGER1; c.arm_body = abody;
      c.p = vaxpy(-rad, nn, pt); c.n = nn; c.dist = d - rad; c.mu = mu_s * (float)PIH_MU_FINGER;
      pcontact_push(slot, n, c);
    }
  }
  // (3b) arm collision spheres of the arm chain
  sfor<0, ArmGeom<ARM>::NSPH>([&](auto q_) {
    constexpr int q = decltype(q_)::value;
    constexpr int abody = sph_body_of<ARM>(q);
    if constexpr (abody < ArmDims<ARM>::NARM) {
      v3 sw = lds3(&pa[PA_SPH + 3 * q]);
      v3 dd = vsub(sw, cw);
      float reach = hl + rad + T.sph_r[q] + (float)PIH_CONTACT_MARGIN;
      if (vdot(dd, dd) < reach * reach) {
        v3 nn;
        float d = seg_sdf(R, cw, sw, nn);
        c.pair = PIH_PAIR_FB_ARM; c.arm_body = abody;
        c.n = mk(-nn.x, -nn.y, -nn.z);
        c.p = vaxpy(-T.sph_r[q], nn, sw); c.dist = d - T.sph_r[q]; c.mu = mu_s * (float)PIH_MU_DEFAULT;
        pcontact_push(slot, n, c);
      }
    }
  });
}

PIH_D void contact_store(float* sm, int k, const PContact& c, int armslot) {
  float* d = &sm[SM_CON + 12 * k];
  d[0] = c.p.x; d[1] = c.p.y; d[2] = c.p.z; d[3] = c.n.x; d[4] = c.n.y; d[5] = c.n.z; d[6] = c.dist; d[7] = c.mu;
  d[8] = __int_as_float(c.pair); d[9] = __int_as_float(c.seg); d[10] = __int_as_float(c.arm_body); d[11] = __int_as_float(armslot);
}

// PegInHole.reset for the pipe (oracle: reset_env_pipe).  Every lane computes the whole reset (rare, warp-uniform branch).
template <int ARM, int TASK> PIH_D void reset_env_pipe(const SimParams& P, EnvState& S, PipeBase& B, float* sm, int lane, float& qj_lane, float& qdj_lane,
                                                       unsigned gid) {
  reset_env<ARM, TASK>(P, S, gid);
  float u[4];
  rng4(P, gid, (unsigned)S.episode, PIH_STREAM_RESET, u);
  B.pos[0] = S.ppos[0]; B.pos[1] = S.ppos[1]; B.pos[2] = S.ppos[2];
  B.quat[0] = B.quat[1] = B.quat[2] = 0.0f; B.quat[3] = 1.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { B.vel[k] = 0.0f; B.ang[k] = 0.0f; }
  B.grasp_idx = (u[2] < 0.5f) ? 0 : 23;
  float v[4];
  rng4(P, gid, (unsigned)S.episode, PIH_STREAM_PIPE, v);
  int kcount = 5 + (int)(v[0] * 20.0f);
  if (kcount > 24) kcount = 24;
  // keys / angles of the 24 joint indices: lane j < 24 holds those of joint index j (= body j)
  float key = 0.0f, ang = 0.0f;
  {
    const int j = lane < 24 ? lane : 0;
    float kk[4], aa[4];
    rng4(P, gid, (unsigned)S.episode, PIH_STREAM_PIPE + 1 + (j >> 2), kk);
    rng4(P, gid, (unsigned)S.episode, PIH_STREAM_PIPE + 7 + (j >> 2), aa);
#pragma unroll
    for (int t = 0; t < 4; ++t) { key = ((j & 3) == t) ? kk[t] : key; ang = ((j & 3) == t) ? aa[t] : ang; }
  }
  int rank = 0;
  for (int j = 0; j < 24; ++j) {
    const float kj = __shfl_sync(0xffffffffu, key, j);
    rank += (kj < key || (kj == key && j < lane)) ? 1 : 0;
  }
  const float qb = (rank < kcount) ? ang * (float)PIH_PIPE_BEND_MAX : 0.0f;     // value of joint index `lane`
  // joint b lives on lane 5 + b
  const float qsh = __shfl_sync(0xffffffffu, qb, (lane >= 6 && lane < 29) ? lane - 5 : 0);
  qj_lane = (lane >= 6 && lane < 29) ? qsh : 0.0f;
  qdj_lane = 0.0f;
  pipe_fk(sm, B, lane, qj_lane);
  pipe_grasp_pose(sm, B.grasp_idx, S.ppos, S.pquat);
}

// ---------------------------------------------------------------- kernels
// thread per env: the arm half of the step (see the header of this file)
template <int ARM, int TASK> __global__ void __launch_bounds__(32, 1) k_pipe_arm(SimParams P, const float* __restrict__ action, float* __restrict__ scratch) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int A = ArmDims<ARM>::ADIM, nb = ArmDims<ARM>::NB;
  const ArmTables& T = c_arm[ARM];
  float* pa = scratch + (size_t)e * PIH_PA_WORDS;
  EnvState S;
  load_state(P, e, S);
  if (S.flags & PIH_FLAG_DONE) { pa[PA_ACTIVE] = 0.0f; return; }
  const v3 off = env_offset(P, e);
  float a[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (action) load_action<A>(action, e, a);
  else {   // synthetic rollouts: action_space.sample() ~ U[-1,1]^A from Philox (seed; env, global step)
    const unsigned gid = (unsigned)(P.env_id_base + e);
    float u0[4];
    rng4(P, gid, P.global_step, PIH_STREAM_ACTION, u0);
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = fmaf(u0[k], 2.0f, -1.0f);
    if (A > 4) {
      rng4(P, gid, P.global_step, PIH_STREAM_ACTION + 1, u0);
#pragma unroll
      for (int k = 4; k < A; ++k) a[k] = fmaf(u0[k - 4], 2.0f, -1.0f);
    }
    if (!P.action_frame_env) { a[0] += off.x; a[1] += off.y; a[2] += off.z; }
  }
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qs[nb], kpc[nb], mfc[nb];
  int ik_n;
  const bool fsm_done = apply_action<ARM, TASK>(P, S, F, a, off, qs, kpc, mfc, ik_n, e);
  float tau[nb];
#pragma unroll
  for (int i = 0; i < nb; ++i) tau[i] = -(T.damp[i] * S.qd[i]);
  ABA<ARM> AB;
  aba<ARM>(F, S.qd, tau, (float)PIH_GRAVITY, AB);
  for (int j = 0; j < nb; ++j) {
    float col[nb];
    minv_column<ARM>(F, AB, j, col);
#pragma unroll
    for (int i = 0; i < nb; ++i) pa[PA_MINV + 9 * i + j] = col[i];
  }
#pragma unroll
  for (int j = 0; j < nb; ++j) {
    pa[PA_QSTAR + j] = qs[j]; pa[PA_KP + j] = kpc[j]; pa[PA_MAXF + j] = mfc[j];
    pa[PA_VS + j] = fmaf(P.dt, AB.qdd[j], S.qd[j]);
    v3 z = mcol(F.Rw[j], 2);
    sts3(&pa[PA_ZJ + 3 * j], z);
    sts3(&pa[PA_PW + 3 * j], F.pw[j]);
  }
  pa[PA_FSM_DONE] = fsm_done ? 1.0f : 0.0f;
  pa[PA_FSM_STATE] = __int_as_float(S.fsm_state); pa[PA_FSM_T] = __int_as_float(S.fsm_t);
#pragma unroll
  for (int k = 0; k < 7; ++k) pa[PA_MTGT + k] = S.mtgt[k];
  pa[PA_GRASP_ANG] = S.grasp_ang;
#pragma unroll
  for (int b = 0; b < ArmGeom<ARM>::NBOX; ++b) {
    const int body = box_body_of<ARM>(b);
    v3 cb = mv_add(F.Rw[body], ld3(T.box_c[b]), F.pw[body]);
    m3 Rb = mmul(F.Rw[body], ldm(T.box_R[b]));
    sts3(&pa[PA_BOXC + 3 * b], cb);
#pragma unroll
    for (int i = 0; i < 9; ++i) pa[PA_BOXR + 9 * b + i] = Rb.m[i / 3][i % 3];
  }
#pragma unroll
  for (int q = 0; q < ArmGeom<ARM>::NSPH; ++q) {
    const int body = sph_body_of<ARM>(q);
    v3 cw = mv_add(F.Rw[body], ld3(T.sph_c[q]), F.pw[body]);
    sts3(&pa[PA_SPH + 3 * q], cw);
  }
  pa[PA_ACTIVE] = 1.0f;
  pa[PA_IK_N] = __int_as_float(ik_n);
}

// EnvState <-> the 48 base words in shared memory (word map of pih_constants.h)
PIH_D void envstate_from_words(const float* w, EnvState& S) {
#pragma unroll
  for (int k = 0; k < PIH_NB_MAX; ++k) { S.q[k] = w[PIH_S_Q + k]; S.qd[k] = w[PIH_S_QD + k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { S.ppos[k] = w[PIH_S_PPOS + k]; S.pvel[k] = w[PIH_S_PVEL + k]; S.pang[k] = w[PIH_S_PANG + k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) S.pquat[k] = w[PIH_S_PQUAT + k];
  S.step = __float_as_int(w[PIH_S_STEP]); S.episode = __float_as_int(w[PIH_S_EPISODE]); S.flags = __float_as_int(w[PIH_S_FLAGS]);
  S.fsm_state = __float_as_int(w[PIH_S_FSM_STATE]); S.fsm_t = __float_as_int(w[PIH_S_FSM_T]); S.rand_y = w[PIH_S_RANDY]; S.grasp_ang = w[PIH_S_GRASP_ANG];
#pragma unroll
  for (int k = 0; k < 7; ++k) S.mtgt[k] = w[PIH_S_MTGT + k];
}
PIH_D void envstate_to_words(float* w, const EnvState& S) {
#pragma unroll
  for (int k = 0; k < PIH_NB_MAX; ++k) { w[PIH_S_Q + k] = S.q[k]; w[PIH_S_QD + k] = S.qd[k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { w[PIH_S_PPOS + k] = S.ppos[k]; w[PIH_S_PVEL + k] = S.pvel[k]; w[PIH_S_PANG + k] = S.pang[k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) w[PIH_S_PQUAT + k] = S.pquat[k];
  w[PIH_S_STEP] = __int_as_float(S.step); w[PIH_S_EPISODE] = __int_as_float(S.episode); w[PIH_S_FLAGS] = __int_as_float(S.flags);
  w[PIH_S_FSM_STATE] = __int_as_float(S.fsm_state); w[PIH_S_FSM_T] = __int_as_float(S.fsm_t); w[PIH_S_RANDY] = S.rand_y; w[PIH_S_GRASP_ANG] = S.grasp_ang;
#pragma unroll
  for (int k = 0; k < 7; ++k) w[PIH_S_MTGT + k] = S.mtgt[k];
}
PIH_D void pipe_from_words(const float* w, int lane, PipeBase& B, float& qj_lane, float& qdj_lane) {
#pragma unroll
  for (int k = 0; k < 3; ++k) { B.pos[k] = w[PIH_PS_POS + k]; B.vel[k] = w[PIH_PS_VEL + k]; B.ang[k] = w[PIH_PS_ANG + k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) B.quat[k] = w[PIH_PS_QUAT + k];
  B.grasp_idx = __float_as_int(w[PIH_PS_GRASP]);
  const int b = lane - 5;
  const bool has = b >= 1 && b < PNB;
  qj_lane = has ? w[PIH_PS_QJ + (has ? b - 1 : 0)] : 0.0f;
  qdj_lane = has ? w[PIH_PS_QDJ + (has ? b - 1 : 0)] : 0.0f;
}
PIH_D void pipe_to_words(float* w, int lane, const PipeBase& B, float qj_lane, float qdj_lane) {
#pragma unroll
  for (int k = 0; k < 3; ++k) { w[PIH_PS_POS + k] = B.pos[k]; w[PIH_PS_VEL + k] = B.vel[k]; w[PIH_PS_ANG + k] = B.ang[k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) w[PIH_PS_QUAT + k] = B.quat[k];
  w[PIH_PS_GRASP] = __int_as_float(B.grasp_idx);
  const int b = lane - 5;
  if (b >= 1 && b < PNB) { w[PIH_PS_QJ + b - 1] = qj_lane; w[PIH_PS_QDJ + b - 1] = qdj_lane; }
}

// warp per env: first reset of the pipe task (k_reset analogue) - lane-cooperative because of the pipe block
template <int ARM, int TASK> __global__ void __launch_bounds__(32) k_pipe_reset(SimParams P, const unsigned char* __restrict__ mask, float* __restrict__ obs) {
  extern __shared__ float sm[];
  const int e = blockIdx.x, lane = threadIdx.x;
  if (mask && !mask[e]) return;
  float* w = sm + SM_SW;
  {
    const float4 v = P.state[(size_t)lane * P.n + e];
    w[4 * lane] = v.x; w[4 * lane + 1] = v.y; w[4 * lane + 2] = v.z; w[4 * lane + 3] = v.w;
  }
  __syncwarp();
  EnvState S;
  envstate_from_words(w, S);
  PipeBase B;
  float qj, qdj;
  pipe_from_words(w, lane, B, qj, qdj);
  S.episode += 1;
  reset_env_pipe<ARM, TASK>(P, S, B, sm, lane, qj, qdj, (unsigned)(P.env_id_base + e));
  __syncwarp();
  for (int k = lane; k < 4 * PIH_PIPE_STATE_F4; k += 32) if (k < PIH_S_LATCH_Q || k >= PIH_PS_BASE) w[k] = 0.0f;   // (the latched grasp scene stays)
  __syncwarp();
  envstate_to_words(w, S);
  pipe_to_words(w, lane, B, qj, qdj);
  __syncwarp();
  P.state[(size_t)lane * P.n + e] = make_float4(w[4 * lane], w[4 * lane + 1], w[4 * lane + 2], w[4 * lane + 3]);
  if (obs) {
    float o[5];
    observe<ARM>(S, env_offset(P, e), o);
    if (lane < 5) obs[5 * (size_t)e + lane] = o[lane < 5 ? lane : 0];
  }
}

// warp per env: everything of the step after the arm half
template <int ARM, int TASK> __global__ void __launch_bounds__(32 * PIH_PIPE_WARPS, PIH_PIPE_MIN_BLOCKS / PIH_PIPE_WARPS) k_pipe_step(SimParams P, const float* __restrict__ scratch,
                                                                                               float* __restrict__ rowbuf, float* __restrict__ obs,
                                                                                               float* __restrict__ reward, unsigned char* __restrict__ done,
                                                                                               unsigned long long* __restrict__ counters) {
  extern __shared__ float sm_all[];
  constexpr int nb = ArmDims<ARM>::NB;
  const ArmTables& T = c_arm[ARM];
  // PIH_PIPE_WARPS envs per CTA, one warp each, each with its own SM_TOTAL floats of shared memory; the warps never
  // synchronise with one another (only __syncwarp inside a warp)
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * PIH_PIPE_WARPS + (threadIdx.x >> 5);
  if (e >= P.n) return;
  float* sm = sm_all + (threadIdx.x >> 5) * SM_TOTAL;
  const unsigned FULL = 0xffffffffu;
  float* w = sm + SM_SW;
  float* pa = sm + SM_PA;
  float* rows_g = rowbuf + (size_t)e * kRowBufWords;     // this env's constraint rows (global scratch, L1/L2 resident)
  // ---- load: lane l fetches state slot l (one float4) and two float4 of the scratch block
  {
    const float4 v = P.state[(size_t)lane * P.n + e];
    w[4 * lane] = v.x; w[4 * lane + 1] = v.y; w[4 * lane + 2] = v.z; w[4 * lane + 3] = v.w;
    const float4* sp = reinterpret_cast<const float4*>(scratch + (size_t)e * PIH_PA_WORDS);
#pragma unroll
    for (int k = 0; k < PIH_PA_WORDS / 128; ++k) {
      const float4 q = __ldg(&sp[32 * k + lane]);
      float* d = &pa[4 * (32 * k + lane)];
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    }
  }
  __syncwarp();
  PipeBase B;
  float qj, qdj;
  pipe_from_words(w, lane, B, qj, qdj);
  const v3 off = env_offset(P, e);
  const unsigned gid = (unsigned)(P.env_id_base + e);
  float o[5], rew;
  unsigned char dn;
  if (pa[PA_ACTIVE] == 0.0f) {
    // frozen env (done, not reset): outputs re-emitted from the frozen state (base_env.py:61-71)
    EnvState S;
    envstate_from_words(w, S);
    observe<ARM>(S, off, o);
    rew = (S.flags & PIH_FLAG_SUCCESS) ? 1.0f : 0.0f;
    dn = 1;
  } else {
    const float dt = P.dt, idt = P.idt;
    const bool fsm_done = pa[PA_FSM_DONE] != 0.0f;
    const int fsm_now = (TASK == PIH_TASK_SCRIPTED) ? __float_as_int(pa[PA_FSM_STATE]) : 0;
    // ---- pipe kinematics and dynamics
    pipe_fk(sm, B, lane, qj);
    float vstar;
    pipe_dynamics(sm, B, lane, qdj, dt, (float)PIH_GRAVITY, vstar);
    // ---- arm rows: lane j < nb owns joint j (its row of M^-1, motor and limit rows)
    const int ja = lane < nb ? lane : 0;
    float MinvA[nb];
#pragma unroll
    for (int j = 0; j < nb; ++j) MinvA[j] = pa[PA_MINV + 9 * ja + j];
    const float vsa = pa[PA_VS + ja];
    float m_dinv, m_b, m_max, m_lam = 0.0f, l_b0, l_b1, l_lam0 = 0.0f, l_lam1 = 0.0f, m_diag;
    bool l_on0, l_on1;
    {
      float qt = pa[PA_QSTAR + ja];
      const float lo = T.lo[ja], hi = T.hi[ja];
      const float qa = w[PIH_S_Q + ja];
      qt = (qt < lo) ? lo : ((qt > hi) ? hi : qt);
      const float vst = pa[PA_KP + ja] * (qt - qa) * idt;
      const float dg = pa[PA_MINV + 10 * ja];
      m_diag = dg;
      m_dinv = 1.0f / dg;
      m_b = (vst - vsa) * m_dinv;
      m_max = pa[PA_MAXF + ja] * dt;
      const float glo = qa - lo, ghi = hi - qa;
      float vdn = (-vst > -vsa) ? -vst : -vsa;
      vdn = (vdn > 0.0f) ? vdn : 0.0f;
      float vup = (vst > vsa) ? vst : vsa;
      vup = (vup > 0.0f) ? vup : 0.0f;
      l_on0 = (lane < nb) && (glo < fmaf(vdn, dt, 0.005f));
      l_on1 = (lane < nb) && (ghi < fmaf(vup, dt, 0.005f));
      const float tlo = (glo > 0.0f) ? -(glo * idt) : -((float)PIH_ERP * glo * idt);
      const float thi = (ghi > 0.0f) ? -(ghi * idt) : -((float)PIH_ERP * ghi * idt);
      l_b0 = (tlo - vsa) * m_dinv;
      l_b1 = (thi + vsa) * m_dinv;
    }
    const unsigned lim_mask0 = __ballot_sync(FULL, l_on0), lim_mask1 = __ballot_sync(FULL, l_on1);
    // ---- pipe joint motors: lane c = 5 + b owns the motor of joint b
    const float pm_diag = (lane < PND) ? sm[SM_MCOL + 33 * lane] : 1.0f;      // M^-1[lane][lane]
    const float pm_dinv = 1.0f / pm_diag;
    const float pm_b = (0.0f - vstar) * pm_dinv;
    float pm_lam = 0.0f;
    // ---- contacts: lane s owns segment s; then the slot-major compaction the oracle defines.  (SM_UH, which aliases the
    //      contact list, is dead: pipe_dynamics ended with a __syncwarp.)
    int ncon = 0, narmc = 0, pairmask = 0;
    if (TASK == PIH_TASK_SCRIPTED && grasp_active(TASK, fsm_now)) {
      // grasp constraint of the scripted task (peg_in_hole.py:99-104): first row set, bilateral, along the env axes
      EnvState S;
      envstate_from_words(w, S);
      FK<ARM> Fa;
      fk<ARM, nb>(S.q, Fa);
      v3 pee; m3 Ree;
      ee_pose<ARM>(Fa, pee, Ree);
      m3 Rg = quat_to_m3(S.pquat);
      v3 pc = mv_add(Rg, mk(0.0f, S.rand_y, 0.0f), mk(S.ppos[0], S.ppos[1], S.ppos[2]));
      PContact c;
      c.pair = PIH_PAIR_GRASP; c.seg = (B.grasp_idx == 0) ? 1 : PNSEG - 1; c.arm_body = ArmDims<ARM>::EE;
      c.p = pee; c.n = vsub(pc, pee); c.dist = -1e30f; c.mu = -1.0f;
      contact_store(sm, 0, c, 0);
      ncon = 1; narmc = 1;
    }
    const int nweld = ncon;
    {
      PContact slot[PIH_PIPE_SEG_MAXC];
      int cnt = 0;
      if (lane < PNSEG) {
        pipe_collide_segment<ARM>(sm, lane, slot, cnt);
        if (cnt == 2 && slot[1].dist < slot[0].dist) { PContact t = slot[0]; slot[0] = slot[1]; slot[1] = t; }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PIH_PIPE_SEG_MAXC; ++k) {
        const bool have = lane < PNSEG && k < cnt;
        const bool arm = have && slot[k].arm_body >= 0;
        // sequential admission in segment order (caps on the total and on the arm-involving contacts)
        const unsigned mh = __ballot_sync(FULL, have), ma = __ballot_sync(FULL, arm);
        int my_idx = -1, my_arm = -1;
        for (int s = 0; s < PNSEG; ++s) {
          if (!((mh >> s) & 1u) || ncon - nweld >= PIH_PIPE_SEG_TOTAL) continue;
          const bool a = (ma >> s) & 1u;
          if (a && narmc >= PIH_PIPE_ARMROWS) continue;
          if (s == lane) { my_idx = ncon; my_arm = a ? narmc : -1; }
          ncon += 1;
          narmc += a ? 1 : 0;
        }
        if (my_idx >= 0) contact_store(sm, my_idx, slot[k], my_arm);
      }
      // arm collision spheres against the table: every lane runs the (tiny) deepest-four selection
      PContact as[PIH_PIPE_ARM_MAXC];
      int an = 0;
      sfor<0, ArmGeom<ARM>::NSPH>([&](auto q_) {
        constexpr int q = decltype(q_)::value;
        v3 cw = lds3(&pa[PA_SPH + 3 * q]);
        PContact c;
        c.pair = PIH_PAIR_ARM_TABLE; c.seg = -1; c.arm_body = sph_body_of<ARM>(q); c.n = mk(0.0f, 0.0f, 1.0f);
        c.p = mk(cw.x, cw.y, cw.z - T.sph_r[q]);
        c.dist = cw.z - T.sph_r[q] - (float)PIH_TABLE_Z; c.mu = (float)PIH_MU_DEFAULT * (float)PIH_MU_DEFAULT;
        if (c.dist < (float)PIH_CONTACT_MARGIN) {
          if (an < PIH_PIPE_ARM_MAXC) {
#pragma unroll
            for (int t = 0; t < PIH_PIPE_ARM_MAXC; ++t) if (t == an) as[t] = c;
            an += 1;
          } else {
            int wi = 0;
#pragma unroll
            for (int t = 1; t < PIH_PIPE_ARM_MAXC; ++t) if (as[t].dist > as[wi].dist) wi = t;
            float wd = as[0].dist;
#pragma unroll
            for (int t = 1; t < PIH_PIPE_ARM_MAXC; ++t) wd = (t == wi) ? as[t].dist : wd;
            if (c.dist < wd) {
#pragma unroll
              for (int t = 0; t < PIH_PIPE_ARM_MAXC; ++t) if (t == wi) as[t] = c;
            }
          }
        }
      });
#pragma unroll
      for (int t = 0; t < PIH_PIPE_ARM_MAXC; ++t) {
        if (t < an && narmc < PIH_PIPE_ARMROWS) {
          contact_store(sm, ncon, as[t], narmc);     // (all lanes store the same values)
          ncon += 1; narmc += 1;
        }
      }
    }
    __syncwarp();
    // ---- contact rows: for every contact and direction, lane i computes entry i of the Jacobian row, then of M^-1 J^T
    {
      // per-lane kinematic data of "its" dof: joint b = lane - 5 (axis and origin in the env frame), arm joint j = lane
      const int jb = (lane >= 6 && lane < PND) ? lane - 5 : 1;
      const int jax = (jb & 1) ? 2 : 0;                  // joint axis of body jb: column jax of its rotation
      const v3 axw = mk(sm[SM_RW + 9 * jb + jax], sm[SM_RW + 9 * jb + 3 + jax], sm[SM_RW + 9 * jb + 6 + jax]);
      const v3 ob = lds3(&sm[SM_PW + 3 * jb]);
      const m3 R0 = ldsm(&sm[SM_RW]);
      const v3 p0 = lds3(&sm[SM_PW]);
      const v3 zja = lds3(&pa[PA_ZJ + 3 * ja]), pja = lds3(&pa[PA_PW + 3 * ja]);
      const int jt_a = T.jtype[ja];
      float* jt = &sm[SM_JTMP];
#pragma unroll 1
      for (int k = 0; k < ncon; ++k) {
        const float* cd = &sm[SM_CON + 12 * k];
        const v3 cp = lds3(cd), cn = lds3(cd + 3);
        const float cdist = cd[6];
        const int cpair = __float_as_int(cd[8]), cseg = __float_as_int(cd[9]), carm = __float_as_int(cd[10]);
        pairmask |= cpair;
        const bool has_pipe = cseg >= 0, has_arm = carm >= 0;
        v3 dir[3];
        const bool weld = TASK == PIH_TASK_SCRIPTED && cpair == PIH_PAIR_GRASP;
        if (weld) { dir[0] = mk(1.0f, 0.0f, 0.0f); dir[1] = mk(0.0f, 1.0f, 0.0f); dir[2] = mk(0.0f, 0.0f, 1.0f); }
        else { dir[0] = cn; plane_space(cn, dir[1], dir[2]); }
        const int Lb = has_pipe ? seg_body(cseg) : 0;
        const unsigned anc = has_arm ? T.anc[carm] : 0u;
#pragma unroll 1
        for (int d = 0; d < 3; ++d) {
          const int r = 3 * k + d;
          const v3 dd = (d == 0) ? dir[0] : ((d == 1) ? dir[1] : dir[2]);
          float* rg = rows_g + (size_t)r * kRowWords;
          float Jp = 0.0f, MJp = 0.0f, Ja = 0.0f, MJa = 0.0f;
          if (has_pipe) {
            const v3 rb = vsub(cp, p0);
            const v3 jang = mTv(R0, vcross(rb, dd));
            const v3 jlin = mTv(R0, dd);
            const float jj = vdot(dd, vcross(axw, vsub(cp, ob)));
            Jp = (lane < 3) ? vgetr(jang, lane) : ((lane < 6) ? vgetr(jlin, lane - 3) : ((lane < PND && lane - 5 <= Lb) ? jj : 0.0f));
            jt[lane] = Jp;
            __syncwarp();
            // entries beyond the contact's body are structural zeros: fma(x, 0, s) == s, so the sum stops at 5 + Lb
            float s = 0.0f;
            const int iend = 6 + Lb;
#pragma unroll 2
            for (int i = 0; i < iend; ++i) s = fmaf(sm[SM_MCOL + 32 * i + lane], jt[i], s);
            MJp = (lane < PND) ? s : 0.0f;
            rg[lane] = Jp; rg[32 + lane] = MJp;
            __syncwarp();
          }
          if (has_arm) {
            const float sg = has_pipe ? -1.0f : 1.0f;
            float v = 0.0f;
            if (lane < nb && ((anc >> lane) & 1u)) {
              if (jt_a == 0) v = vdot(dd, vcross(zja, vsub(cp, pja)));
              else v = vdot(dd, zja);
              v = sg * v;
            }
            Ja = v;
            if (lane < 16) jt[32 + lane] = Ja;
            __syncwarp();
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < nb; ++j) s = fmaf(MinvA[j], jt[32 + j], s);
            MJa = (lane < nb) ? s : 0.0f;
            if (lane < 16) { rg[64 + lane] = Ja; rg[80 + lane] = MJa; }
            __syncwarp();
          }
          float pd = 0.0f, pv = 0.0f;
          if (has_pipe && lane < PND) { pd = Jp * MJp; pv = Jp * vstar; }
          if (has_arm && lane < nb) {
            if (has_pipe) { pd = fmaf(Ja, MJa, pd); pv = fmaf(Ja, vsa, pv); }
            else { pd = Ja * MJa; pv = Ja * vsa; }
          }
          const float diag = bfly32(pd), vrel = bfly32(pv);
          const float dinv = 1.0f / diag;
          float tv = 0.0f;
          if (weld) tv = -((float)PIH_ERP * (d == 0 ? cn.x : (d == 1 ? cn.y : cn.z)) * idt);
          else if (d == 0) tv = (cdist > 0.0f) ? -(cdist * idt) : -((float)PIH_ERP * cdist * idt);
          float* rs = &sm[SM_ROWS + 4 * r];
          rs[0] = dinv; rs[1] = diag; rs[2] = (tv - vrel) * dinv; rs[3] = 0.0f;      // (every lane writes the same values)
        }
      }
    }
    __syncwarp();
    // ---- projected Gauss-Seidel, one island: arm motors, arm limits, pipe motors, contact normals, contact friction
    float dva = 0.0f, dvp = 0.0f;
    int st_it = 0;
    // which contacts involve the pipe / the arm, as bit masks in registers (no shared-memory look-up on the row's critical path)
    unsigned long long mpipe = 0ull, marm = 0ull;
    static_assert(PIH_PIPE_MAXC <= 64, "contact masks are 64-bit");
    for (int k = 0; k < ncon; ++k) {
      mpipe |= (unsigned long long)(__float_as_int(sm[SM_CON + 12 * k + 9]) >= 0) << k;
      marm |= (unsigned long long)(__float_as_int(sm[SM_CON + 12 * k + 10]) >= 0) << k;
    }
    const int nrow = 3 * ncon;
    // Row t of a sweep: the normals of all contacts first (t < ncon), then the two friction rows of each contact.  Its
    // coefficients live in the global row block; they are fetched ONE ROW AHEAD so that the L2 latency of row t + 1
    // hides behind the butterfly of row t (the shared-memory carve-out leaves little L1 for them).
    struct RowRegs { float j, mj, ja, mja; };
    auto fetch = [&](int t, RowRegs& q) {
      q.j = 0.0f; q.mj = 0.0f; q.ja = 0.0f; q.mja = 0.0f;
      if (t < nrow) {
        const int u = t - ncon;
        const int k = (t < ncon) ? t : (u >> 1), d = (t < ncon) ? 0 : 1 + (u & 1);
        const float* rg = rows_g + (size_t)(3 * k + d) * kRowWords;
        if ((mpipe >> k) & 1ull) { q.j = rg[lane]; q.mj = rg[32 + lane]; }
        if (((marm >> k) & 1ull) && lane < nb) { q.ja = rg[64 + lane]; q.mja = rg[80 + lane]; }
      }
    };
#pragma unroll 1
    for (int it = 0; it < P.solver_iters; ++it) {
      float res = 0.0f;
      RowRegs nxt;
      fetch(0, nxt);                  // in flight while the motor rows are solved
      // Motor and limit rows.  Every lane evaluates "its" row formula at every row step, but only the owner's value is
      // used (broadcast by the shuffle).  A lane owns each of its rows ONCE per sweep, so its accumulated impulse, the
      // clamp interval derived from it and its residual are handled outside the row steps: a row step is FFMA, two
      // FMNMX, a select, the shuffle and the velocity update - the dependent chain of the sweep.  (fminf / fmaxf against
      // the oracle's ternaries: same values; the rigid-peg kernels rely on the same equivalence.)
      {
        const float dlo = -m_max - m_lam, dhi = m_max - m_lam;
        float dl_own = 0.0f;
#pragma unroll
        for (int j = 0; j < nb; ++j) {
          float dl = fminf(fmaxf(fmaf(-dva, m_dinv, m_b), dlo), dhi);
          dl_own = (lane == j) ? dl : dl_own;
          dl = __shfl_sync(FULL, dl, j);
          dva = fmaf(MinvA[j], dl, dva);
        }
        m_lam = m_lam + dl_own;                     // (lanes >= nb own nothing: their values are never read)
        float rv = dl_own * m_diag;
        rv = rv * rv;
        res = (lane < nb && rv > res) ? rv : res;
      }
      if (lim_mask0 | lim_mask1) {
        float dl0_own = 0.0f, dl1_own = 0.0f;
#pragma unroll
        for (int j = 0; j < nb; ++j) {
          if ((lim_mask0 >> j) & 1u) {
            float dl = fmaxf(fmaf(-dva, m_dinv, l_b0), -l_lam0);
            dl0_own = (lane == j) ? dl : dl0_own;
            dl = __shfl_sync(FULL, dl, j);
            dva = fmaf(MinvA[j], dl, dva);
          }
          if ((lim_mask1 >> j) & 1u) {
            float dl = fmaxf(fmaf(-(-1.0f * dva), m_dinv, l_b1), -l_lam1);
            dl1_own = (lane == j) ? dl : dl1_own;
            dl = dl * -1.0f;
            dl = __shfl_sync(FULL, dl, j);
            dva = fmaf(MinvA[j], dl, dva);
          }
        }
        l_lam0 = l_lam0 + dl0_own; l_lam1 = l_lam1 + dl1_own;
        float r0 = dl0_own * m_diag, r1 = (dl1_own * -1.0f) * m_diag;
        r0 = r0 * r0; r1 = r1 * r1;
        res = (l_on0 && r0 > res) ? r0 : res;
        res = (l_on1 && r1 > res) ? r1 : res;
      }
      // pipe joint motors: lane c = 5 + b owns joint b; its column of M^-1 comes from shared memory with an immediate offset
      {
        const float dlo = (float)(-PIH_PIPE_MOTOR_IMPULSE) - pm_lam, dhi = (float)PIH_PIPE_MOTOR_IMPULSE - pm_lam;
        float dl_own = 0.0f;
        const float* mcol = &sm[SM_MCOL + lane];
#pragma unroll
        for (int c = 6; c < PND; ++c) {
          float dl = fminf(fmaxf(fmaf(-dvp, pm_dinv, pm_b), dlo), dhi);
          dl_own = (lane == c) ? dl : dl_own;
          dl = __shfl_sync(FULL, dl, c);
          dvp = fmaf(mcol[32 * c], dl, dvp);
        }
        pm_lam = pm_lam + dl_own;
        float rv = dl_own * pm_diag;
        rv = rv * rv;
        res = (lane >= 6 && lane < PND && rv > res) ? rv : res;
      }
      // contacts: normals, then friction
#pragma unroll 1
      for (int t = 0; t < nrow; ++t) {
        const RowRegs cur = nxt;
        fetch(t + 1, nxt);
        const int u = t - ncon;
        const int k = (t < ncon) ? t : (u >> 1), d = (t < ncon) ? 0 : 1 + (u & 1);
        const bool has_pipe = (mpipe >> k) & 1ull, has_arm = (marm >> k) & 1ull;
        const int r = 3 * k + d;
        float* rs = &sm[SM_ROWS + 4 * r];
        const float rdinv = rs[0], rdiag = rs[1], rb = rs[2], lam = rs[3];
        const float mu = sm[SM_CON + 12 * k + 7], lam0 = sm[SM_ROWS + 12 * k + 3];
        float p = 0.0f;
        if (has_pipe) p = cur.j * dvp;
        if (has_arm && lane < nb) p = has_pipe ? fmaf(cur.ja, dva, p) : cur.ja * dva;
        const float rel = bfly32(p);
        float dl = fmaf(-rel, rdinv, rb);
        if (TASK == PIH_TASK_SCRIPTED && mu < 0.0f) {   // grasp constraint: bilateral, |impulse| <= maxForce dt
          const float wl = (float)PIH_GRASP_MAX_FORCE * dt;
          const float dlo = -wl - lam, dhi = wl - lam;
          dl = (dl < dlo) ? dlo : dl;
          dl = (dl > dhi) ? dhi : dl;
        } else if (d == 0) {
          const float dlo = -lam;
          dl = (dl < dlo) ? dlo : dl;
        } else {
          const float lim = mu * lam0;
          const float dlo = -lim - lam, dhi = lim - lam;
          dl = (dl < dlo) ? dlo : dl;
          dl = (dl > dhi) ? dhi : dl;
        }
        rs[3] = lam + dl;      // (every lane writes the same value and later reads back its own write)
        float rv = dl * rdiag;
        rv = rv * rv;
        res = (rv > res) ? rv : res;
        if (has_pipe) dvp = fmaf(cur.mj, dl, dvp);
        if (has_arm && lane < nb) dva = fmaf(cur.mja, dl, dva);
      }
      st_it += 1;
      // residual of the iteration: max over the lanes (non-negative floats order like their bit patterns)
      const unsigned rmax = __reduce_max_sync(FULL, __float_as_uint(res));
      if (__uint_as_float(rmax) <= P.solver_residual) break;
    }
    if (P.stats) {
      if (lane == 0) {
        atomicAdd(&P.stats[PIH_STAT_ENV_STEPS], 1ull);
        atomicAdd(&P.stats[PIH_STAT_IK_ITERS], (unsigned long long)__float_as_int(pa[PA_IK_N]));
        atomicAdd(&P.stats[PIH_STAT_PGS_ROW_ITERS], (unsigned long long)(st_it * (nb + __popc(lim_mask0) + __popc(lim_mask1) + PIH_PIPE_NJ + 3 * ncon)));
        atomicAdd(&P.stats[PIH_STAT_PGS_ITERS], (unsigned long long)st_it);
        atomicAdd(&P.stats[PIH_STAT_CONTACTS], (unsigned long long)ncon);
        atomicAdd(&P.stats[PIH_STAT_LIMIT_ROWS], (unsigned long long)(__popc(lim_mask0) + __popc(lim_mask1)));
        atomicMax(&P.stats[PIH_STAT_PGS_ITERS_MAX], (unsigned long long)st_it);
      }
    }
    // ---- integrate.  Arm: every lane integrates all joints (the full EnvState is identical in all lanes from here on).
    EnvState S;
    envstate_from_words(w, S);
    if (TASK == PIH_TASK_SCRIPTED) {
      S.fsm_state = __float_as_int(pa[PA_FSM_STATE]); S.fsm_t = __float_as_int(pa[PA_FSM_T]);
#pragma unroll
      for (int k = 0; k < 7; ++k) S.mtgt[k] = pa[PA_MTGT + k];
      S.grasp_ang = pa[PA_GRASP_ANG];
    }
#pragma unroll
    for (int j = 0; j < nb; ++j) {
      const float vsj = __shfl_sync(FULL, vsa, j), dvj = __shfl_sync(FULL, dva, j);
      float v = vsj + dvj;
      v = (v > (float)PIH_MAX_COORD_VEL) ? (float)PIH_MAX_COORD_VEL : ((v < (float)(-PIH_MAX_COORD_VEL)) ? (float)(-PIH_MAX_COORD_VEL) : v);
      S.qd[j] = v;
      S.q[j] = fmaf(dt, S.qd[j], S.q[j]);
    }
    {
      const float vnew = vstar + dvp;
      // joints: lane 5 + b
      float vj = vnew;
      vj = (vj > (float)PIH_MAX_COORD_VEL) ? (float)PIH_MAX_COORD_VEL : ((vj < (float)(-PIH_MAX_COORD_VEL)) ? (float)(-PIH_MAX_COORD_VEL) : vj);
      if (lane >= 6 && lane < PND) { qdj = vj; qj = fmaf(dt, vj, qj); }
      // base: twist in body coordinates from lanes 0..5
      float t6[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) t6[k] = __shfl_sync(FULL, vnew, k);
      const m3 R0b = ldsm(&sm[SM_RW]);
      v3 fw = clamp_norm(mv(R0b, mk(t6[0], t6[1], t6[2])));
      v3 fv = clamp_norm(mv(R0b, mk(t6[3], t6[4], t6[5])));
      B.vel[0] = fv.x; B.vel[1] = fv.y; B.vel[2] = fv.z;
      B.ang[0] = fw.x; B.ang[1] = fw.y; B.ang[2] = fw.z;
      B.pos[0] = fmaf(dt, fv.x, B.pos[0]); B.pos[1] = fmaf(dt, fv.y, B.pos[1]); B.pos[2] = fmaf(dt, fv.z, B.pos[2]);
      float w2 = vdot(fw, fw);
      float wn = sqrtf(w2);
      float ha = 0.5f * dt * wn;
      float sh, ch, kk;
      det_sincos(ha, sh, ch);
      if (wn > 1e-6f) kk = sh / wn; else kk = 0.5f * dt;
      float dq[4] = {fw.x * kk, fw.y * kk, fw.z * kk, ch}, nq[4];
      quat_mul(dq, B.quat, nq);
      float n2 = fmaf(nq[3], nq[3], fmaf(nq[2], nq[2], fmaf(nq[1], nq[1], nq[0] * nq[0])));
      float in = 1.0f / sqrtf(n2);
#pragma unroll
      for (int i = 0; i < 4; ++i) B.quat[i] = nq[i] * in;
    }
    S.step += 1;
    // ---- grasp link pose of the new state
    __syncwarp();
    pipe_fk(sm, B, lane, qj);
    pipe_grasp_pose(sm, B.grasp_idx, S.ppos, S.pquat);
    // ---- fault guard
    bool finite = true;
#pragma unroll
    for (int j = 0; j < nb; ++j) finite = finite && finite_word(S.q[j]) && finite_word(S.qd[j]);
#pragma unroll
    for (int k = 0; k < 3; ++k) finite = finite && finite_word(B.pos[k]) && finite_word(B.vel[k]) && finite_word(B.ang[k]) && finite_word(S.ppos[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) finite = finite && finite_word(B.quat[k]) && finite_word(S.pquat[k]);
    finite = finite && finite_word(qj) && finite_word(qdj);
    finite = __all_sync(FULL, finite);
    if (!finite) {
      if (P.stats && lane == 0) atomicAdd(&P.stats[PIH_STAT_FAULTS], 1ull);
      S.episode += 1;
      reset_env_pipe<ARM, TASK>(P, S, B, sm, lane, qj, qdj, gid);
      S.flags = PIH_FLAG_FAULT | (P.auto_reset ? 0 : PIH_FLAG_DONE);
      rew = 0.0f;
      dn = 1;
    } else {
      v3 dd = mk(S.ppos[0] - (float)PIH_HOLE_X, S.ppos[1] - (float)PIH_HOLE_Y, S.ppos[2] - (float)PIH_HOLE_Z);
      const int success = vdot(dd, dd) < (float)(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
      rew = success ? 1.0f : 0.0f;
      int d_ = success || (P.max_steps > 0 && S.step >= P.max_steps);
      if (TASK == PIH_TASK_SCRIPTED) d_ = fsm_done || (P.max_steps > 0 && S.step >= P.max_steps);
      S.flags = (d_ ? PIH_FLAG_DONE : 0) | (success ? PIH_FLAG_SUCCESS : 0) | (pairmask << PIH_PAIR_SHIFT);
      dn = (unsigned char)d_;
      if (d_ && P.auto_reset) {
        const int keep = S.flags & ~PIH_FLAG_DONE;
        S.episode += 1;
        reset_env_pipe<ARM, TASK>(P, S, B, sm, lane, qj, qdj, gid);
        S.flags = (keep & ~0xff) | (success ? PIH_FLAG_PREV_SUCCESS : 0);
      }
    }
    observe<ARM>(S, off, o);
    // ---- store: pack the words in shared memory, lane l writes slot l
    __syncwarp();
    for (int k = lane; k < 4 * PIH_PIPE_STATE_F4; k += 32) if (k < PIH_S_LATCH_Q || k >= PIH_PS_BASE) w[k] = 0.0f;   // (the latched grasp scene stays)
    __syncwarp();
    envstate_to_words(w, S);
    pipe_to_words(w, lane, B, qj, qdj);
    __syncwarp();
    P.state[(size_t)lane * P.n + e] = make_float4(w[4 * lane], w[4 * lane + 1], w[4 * lane + 2], w[4 * lane + 3]);
    if (counters && lane == 0) { atomicAdd(&counters[0], 1ull); atomicAdd(&counters[1], (unsigned long long)dn); }
  }
  if (obs && lane < 5) obs[5 * (size_t)e + lane] = o[lane < 5 ? lane : 0];
  if (lane == 0) {
    if (reward) reward[e] = rew;
    if (done) done[e] = dn;
  }
}

}  // namespace pih
