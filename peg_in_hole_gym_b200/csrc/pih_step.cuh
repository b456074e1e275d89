// pih_step.cuh - contacts, constraint rows, projected Gauss-Seidel, reset / observe and the fused per-env step.
// Reference call sites: BaseEnv.step (envs/base_env.py:60-75); stepSimulation [3P];
// PegInHole.reset distributions (envs/peg_in_hole.py:230-267); success test (envs/peg_in_hole.py:114-117).
#pragma once
#include "pih_device.cuh"

namespace pih {

// unroll factor of the PGS motor / limit row loops (1 = rolled: small loop body that stays in the L0 I-cache)
#ifndef PIH_PGS_UNROLL
#define PIH_PGS_UNROLL 9
#endif
constexpr int kPgsUnroll = PIH_PGS_UNROLL;

template <int ARM> __host__ __device__ constexpr int sph_body_of(int s) {
  constexpr int tp[] = {PIH_PANDA_SPH_BODY_LIST};
  constexpr int tu[] = {PIH_UR5_SPH_BODY_LIST};
  return ARM == PIH_ARM_PANDA ? tp[s] : tu[s];
}
template <int ARM> __host__ __device__ constexpr int box_body_of(int s) {
  constexpr int tp[] = {PIH_PANDA_BOX_BODY_LIST};
  constexpr int tu[] = {PIH_UR5_BOX_BODY_LIST};
  return ARM == PIH_ARM_PANDA ? tp[s] : tu[s];
}
template <int ARM> struct ArmGeom;
template <> struct ArmGeom<PIH_ARM_PANDA> { static constexpr int NSPH = PIH_PANDA_NSPH, NBOX = PIH_PANDA_NBOX; };
template <> struct ArmGeom<PIH_ARM_UR5> { static constexpr int NSPH = PIH_UR5_NSPH, NBOX = PIH_UR5_NBOX; };

struct Contact {
  int pair, has_fb, arm_body;
  v3 p, n;
  float dist, mu;
};
struct ContactSet {
  Contact c[PIH_MAXC];
  int n;
};

// keep the PIH_MAXC deepest candidates; ties keep the earlier candidate.  The margin test is inlined at every
// candidate site; the (rare) insertion is one out-of-line copy so that it does not bloat the instruction stream.
__device__ __noinline__ void contact_insert(ContactSet& S, const Contact& c) {
  if (S.n < PIH_MAXC) { S.c[S.n++] = c; return; }
  int w = 0;
  for (int i = 1; i < PIH_MAXC; ++i)
    if (S.c[i].dist > S.c[w].dist) w = i;
  if (c.dist < S.c[w].dist) S.c[w] = c;
}
PIH_D void contact_push(ContactSet& S, const Contact& c) {
  if (c.dist < (float)PIH_CONTACT_MARGIN) contact_insert(S, c);
}

PIH_D float sd_rect(float da, float db, float sa, float sb, float& na, float& nb_) {
  if (da > 0.0f && db > 0.0f) {
    float l = sqrtf(fmaf(da, da, db * db));
    float il = 1.0f / l;
    na = sa * da * il; nb_ = sb * db * il;
    return l;
  }
  if (da > db) { na = sa; nb_ = 0.0f; return da; }
  na = 0.0f; nb_ = sb;
  return db;
}

// signed distance + outward normal (local frame) of an axis-aligned box with half extents h at local point l
PIH_D float sd_box(v3 l, v3 h, v3& nl) {
  float dx = fabsf(l.x) - h.x, dy = fabsf(l.y) - h.y, dz = fabsf(l.z) - h.z;
  float sx = (l.x < 0.0f) ? -1.0f : 1.0f, sy = (l.y < 0.0f) ? -1.0f : 1.0f, sz = (l.z < 0.0f) ? -1.0f : 1.0f;
  float px = dx > 0.0f ? dx : 0.0f, py = dy > 0.0f ? dy : 0.0f, pz = dz > 0.0f ? dz : 0.0f;
  float o2 = fmaf(pz, pz, fmaf(py, py, px * px));
  float d;
  if (o2 > 0.0f) {
    d = sqrtf(o2);
    float il = 1.0f / d;
    nl = mk(sx * px * il, sy * py * il, sz * pz * il);
  } else if (dx > dy && dx > dz) { d = dx; nl = mk(sx, 0.0f, 0.0f); }
  else if (dy > dz) { d = dy; nl = mk(0.0f, sy, 0.0f); }
  else { d = dz; nl = mk(0.0f, 0.0f, sz); }
  return d;
}

template <int TASK> PIH_D float fb_sdf(const SimParams& P, const m3& R, v3 p, v3 x, v3& n_env) {
  v3 l = mTv(R, vsub(x, p));
  v3 nl;
  float d;
  if ((TASK != PIH_TASK_RANDOM_FLY)) {
    float rho = sqrtf(fmaf(l.z, l.z, l.x * l.x));
    float da = rho - P.fb_half[0], db = fabsf(l.y) - P.fb_half[1];
    float na, nb_;
    d = sd_rect(da, db, 1.0f, (l.y < 0.0f) ? -1.0f : 1.0f, na, nb_);
    float ir = (rho > 1e-12f) ? 1.0f / rho : 0.0f;
    nl = mk(na * l.x * ir, nb_, na * l.z * ir);
  } else {
    d = sd_box(l, mk(P.fb_half[0], P.fb_half[1], P.fb_half[2]), nl);
  }
  n_env = mv(R, nl);
  return d;
}

PIH_D float hole_sdf(v3 x, v3& n) {
  float s = x.x - (float)PIH_HOLE_X, y = x.y - (float)PIH_HOLE_Y, z = x.z - (float)PIH_HOLE_Z;
  float rho = sqrtf(fmaf(z, z, y * y));
  const float rmid = (float)(0.5 * (PIH_HOLE_RIN + PIH_HOLE_ROUT)), hw = (float)(0.5 * (PIH_HOLE_ROUT - PIH_HOLE_RIN));
  float a = rho - rmid;
  float da = fabsf(a) - hw, db = fabsf(s) - (float)PIH_HOLE_HL;
  float na, nb_;
  float d = sd_rect(da, db, (a < 0.0f) ? -1.0f : 1.0f, (s < 0.0f) ? -1.0f : 1.0f, na, nb_);
  float ir = (rho > 1e-12f) ? 1.0f / rho : 0.0f;
  n = mk(nb_, na * y * ir, na * z * ir);
  return d;
}

// unit circle at 45 degree steps (same doubles as the oracle, rounded to float at compile time)
PIH_D void circ8(int k, float& c, float& s) {
  const float h = (float)0.7071067811865476;
  const int kk = k & 7;
  c = (kk == 0) ? 1.0f : (kk == 4) ? -1.0f : (kk == 2 || kk == 6) ? 0.0f : (kk == 1 || kk == 7) ? h : -h;
  s = (kk == 2) ? 1.0f : (kk == 6) ? -1.0f : (kk == 0 || kk == 4) ? 0.0f : (kk == 1 || kk == 3) ? h : -h;
}

template <int TASK> PIH_D bool fb_sample(const SimParams& P, int k, v3& pl) {
  if ((TASK != PIH_TASK_RANDOM_FLY)) {
    float c, s;
    circ8(k, c, s);
    float sy = (k < 8) ? 1.0f : -1.0f;
    pl = mk(P.fb_half[0] * c, sy * P.fb_half[1], P.fb_half[2] * s);
    return true;
  }
  if (k >= 8) return false;
  pl = mk((k & 1) ? P.fb_half[0] : -P.fb_half[0], (k & 2) ? P.fb_half[1] : -P.fb_half[1], (k & 4) ? P.fb_half[2] : -P.fb_half[2]);
  return true;
}

template <int ARM, int TASK> PIH_D void collide(const SimParams& P, const EnvState& S, const FK<ARM>& F, const m3& Rfb, ContactSet& CS) {
  const ArmTables& T = c_arm[ARM];
  CS.n = 0;
  v3 pfb = mk(S.ppos[0], S.ppos[1], S.ppos[2]);
  Contact c;
  const float mu_static = P.fb_mu * (float)PIH_MU_DEFAULT;
  // (1) free body vs table plane
  if (pfb.z - (float)PIH_TABLE_Z < 0.25f) {  // cheap cull: nothing of the body is near the table (exact no-op otherwise)
    if ((TASK != PIH_TASK_RANDOM_FLY)) {
      v3 ax = mcol(Rfb, 1);
      float naz = ax.z;
      v3 nperp = mk(-(naz * ax.x), -(naz * ax.y), fmaf(-naz, ax.z, 1.0f));
      float l2 = vdot(nperp, nperp);
      if (l2 > 0.0025f) {
        float il = 1.0f / sqrtf(l2);
        v3 t = vscale(nperp, il);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float sgn = k ? -1.0f : 1.0f;
          v3 cc = vaxpy(sgn * P.fb_half[1], ax, pfb);
          v3 pt = vaxpy(-P.fb_half[0], t, cc);
          c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = mk(0.0f, 0.0f, 1.0f);
          c.dist = pt.z - (float)PIH_TABLE_Z; c.mu = mu_static;
          contact_push(CS, c);
        }
      } else {
        float sgn = (naz > 0.0f) ? -1.0f : 1.0f;
        v3 cc = vaxpy(sgn * P.fb_half[1], ax, pfb);
        v3 ex = mcol(Rfb, 0), ez = mcol(Rfb, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v3 pt = vaxpy((k & 1) ? -P.fb_half[0] : P.fb_half[0], (k & 2) ? ez : ex, cc);
          c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = mk(0.0f, 0.0f, 1.0f);
          c.dist = pt.z - (float)PIH_TABLE_Z; c.mu = mu_static;
          contact_push(CS, c);
        }
      }
    } else {
      for (int k = 0; k < 8; ++k) {
        v3 pl;
        fb_sample<TASK>(P, k, pl);
        v3 pt = mv_add(Rfb, pl, pfb);
        c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = mk(0.0f, 0.0f, 1.0f);
        c.dist = pt.z - (float)PIH_TABLE_Z; c.mu = mu_static;
        contact_push(CS, c);
      }
    }
  }
  // (2) peg vs annular hole
  if (TASK != PIH_TASK_RANDOM_FLY) {
    v3 dh = vsub(pfb, mk((float)PIH_HOLE_X, (float)PIH_HOLE_Y, (float)PIH_HOLE_Z));
    const float reach = (float)(PIH_PEG_HL + PIH_PEG_R + PIH_HOLE_ROUT + PIH_HOLE_HL + PIH_CONTACT_MARGIN);
    if (vdot(dh, dh) < reach * reach) {
      for (int k = 0; k < 16; ++k) {
        v3 pl, nn;
        fb_sample<TASK>(P, k, pl);
        v3 pt = mv_add(Rfb, pl, pfb);
        float d = hole_sdf(pt, nn);
        c.pair = PIH_PAIR_PEG_HOLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = nn; c.dist = d; c.mu = mu_static;
        contact_push(CS, c);
      }
      for (int k = 0; k < 32; ++k) {
        float rr = (k < 16) ? (float)PIH_HOLE_RIN : (float)PIH_HOLE_ROUT;
        float sx = (k & 8) ? (float)(-PIH_HOLE_HL) : (float)PIH_HOLE_HL;
        float cc, ss;
        circ8(k, cc, ss);
        v3 pt = mk((float)PIH_HOLE_X + sx, fmaf(rr, cc, (float)PIH_HOLE_Y), fmaf(rr, ss, (float)PIH_HOLE_Z));
        v3 nn;
        float d = fb_sdf<TASK>(P, Rfb, pfb, pt, nn);
        c.pair = PIH_PAIR_PEG_HOLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = mk(-nn.x, -nn.y, -nn.z); c.dist = d; c.mu = mu_static;
        contact_push(CS, c);
      }
    }
  }
  // (3) free body vs finger boxes
#pragma unroll
  for (int b = 0; b < ArmGeom<ARM>::NBOX; ++b) {
    constexpr int dummy = 0; (void)dummy;
    const int body = box_body_of<ARM>(b);
    v3 cb = mv_add(F.Rw[body], ld3(T.box_c[b]), F.pw[body]);
    v3 dd = vsub(pfb, cb);
    float reach = P.fb_half[1] + P.fb_half[0] + T.box_h[b][2] + T.box_h[b][0] + (float)PIH_CONTACT_MARGIN;
    if (!(vdot(dd, dd) < reach * reach)) continue;
    m3 Rb = mmul(F.Rw[body], ldm(T.box_R[b]));
    int npts = ((TASK != PIH_TASK_RANDOM_FLY)) ? 5 : 8;
    for (int k = 0; k < npts; ++k) {
      v3 pl;
      float rad;
      if ((TASK != PIH_TASK_RANDOM_FLY)) { pl = mk(0.0f, (P.fb_half[1] - P.fb_half[0]) * 0.5f * (float)(k - 2), 0.0f); rad = P.fb_half[0]; }
      else { fb_sample<TASK>(P, k, pl); rad = 0.0f; }
      v3 pt = mv_add(Rfb, pl, pfb);
      v3 l = mTv(Rb, vsub(pt, cb));
      v3 nl;
      float d = sd_box(l, ld3(T.box_h[b]), nl);
      v3 nn = mv(Rb, nl);
      c.pair = b ? PIH_PAIR_PEG_FINGER2 : PIH_PAIR_PEG_FINGER1; c.has_fb = 1; c.arm_body = body;
      c.p = vaxpy(-rad, nn, pt); c.n = nn; c.dist = d - rad; c.mu = P.fb_mu * (float)PIH_MU_FINGER;
      contact_push(CS, c);
    }
  }
  // (3b) free body vs arm collision spheres (hand, wrist, forearm); finger-tip spheres are covered by the boxes of (3)
  sfor<0, ArmGeom<ARM>::NSPH>([&](auto s_) {
    constexpr int s = decltype(s_)::value;
    constexpr int body = sph_body_of<ARM>(s);
    if constexpr (body < ArmDims<ARM>::NARM) {
      v3 cw = mv_add(F.Rw[body], ld3(T.sph_c[s]), F.pw[body]);
      v3 dd = vsub(cw, pfb);
      float reach = P.fb_half[0] + P.fb_half[1] + P.fb_half[2] + T.sph_r[s] + (float)PIH_CONTACT_MARGIN;
      if (vdot(dd, dd) < reach * reach) {
        v3 nn;
        float d = fb_sdf<TASK>(P, Rfb, pfb, cw, nn);
        c.pair = PIH_PAIR_FB_ARM; c.has_fb = 1; c.arm_body = body;
        c.n = mk(-nn.x, -nn.y, -nn.z);
        c.p = vaxpy(-T.sph_r[s], nn, cw); c.dist = d - T.sph_r[s]; c.mu = P.fb_mu * (float)PIH_MU_DEFAULT;
        contact_push(CS, c);
      }
    }
  });
  // (4) arm collision spheres vs table plane
#pragma unroll
  for (int s = 0; s < ArmGeom<ARM>::NSPH; ++s) {
    const int body = sph_body_of<ARM>(s);
    v3 cw = mv_add(F.Rw[body], ld3(T.sph_c[s]), F.pw[body]);
    c.pair = PIH_PAIR_ARM_TABLE; c.has_fb = 0; c.arm_body = body; c.n = mk(0.0f, 0.0f, 1.0f);
    c.p = mk(cw.x, cw.y, cw.z - T.sph_r[s]);
    c.dist = cw.z - T.sph_r[s] - (float)PIH_TABLE_Z; c.mu = (float)PIH_MU_DEFAULT * (float)PIH_MU_DEFAULT;
    contact_push(CS, c);
  }
}

PIH_D void plane_space(v3 n, v3& t1, v3& t2) {
  if (fabsf(n.z) > (float)0.7071067811865476) {
    float a = fmaf(n.z, n.z, n.y * n.y);
    float k = 1.0f / sqrtf(a);
    t1 = mk(0.0f, -(n.z * k), n.y * k);
    t2 = mk(a * k, -(n.x * t1.z), n.x * t1.y);
  } else {
    float a = fmaf(n.y, n.y, n.x * n.x);
    float k = 1.0f / sqrtf(a);
    t1 = mk(-(n.y * k), n.x * k, 0.0f);
    t2 = mk(-(n.z * t1.y), n.z * t1.x, a * k);
  }
}

// clamp |v| to PIH_MAX_COORD_VEL (btMultiBody m_maxCoordinateVelocity [3P])
PIH_D v3 clamp_norm(v3 v) {
  float n2 = vdot(v, v);
  if (n2 > (float)(PIH_MAX_COORD_VEL * PIH_MAX_COORD_VEL)) return vscale(v, (float)PIH_MAX_COORD_VEL / sqrtf(n2));
  return v;
}

template <int ARM> struct CRow {
  int has_fb, arm_body;
  v3 dir[3], jang[3], mang[3];
  float ja[3][ArmDims<ARM>::NB], ma[3][ArmDims<ARM>::NB];
  float dinv[3], diag[3], b[3], lam[3], mu;
};

template <int ARM> PIH_D unsigned anc_rt(int body) {
  unsigned a = 0;
#pragma unroll
  for (int i = 0; i < ArmDims<ARM>::NB; ++i)
    if (body == i) a = anc_of<ARM>(i);
  return a;
}

// PegInHole.reset distributions (envs/peg_in_hole.py:233-245) with Philox (seed; env id, episode)
template <int ARM, int TASK> PIH_D void reset_env(const SimParams& P, EnvState& S, unsigned gid) {
  const ArmTables& T = c_arm[ARM];
  float u[4];
  rng4(P, gid, (unsigned)S.episode, PIH_STREAM_RESET, u);
#pragma unroll
  for (int i = 0; i < PIH_NB_MAX; ++i) { S.q[i] = (i < ArmDims<ARM>::NB) ? T.rest[i < ArmDims<ARM>::NB ? i : 0] : 0.0f; S.qd[i] = 0.0f; }
  if (TASK != PIH_TASK_RANDOM_FLY) {
    S.ppos[0] = fmaf(u[0], 0.4f, -0.2f);
    S.ppos[1] = fmaf(u[1], 0.2f, -0.6f);
    S.ppos[2] = (float)PIH_PEG_SPAWN_Z;
    S.pvel[0] = S.pvel[1] = S.pvel[2] = 0.0f;
  } else {
    float x = fmaf(u[0], 1.4f, -0.7f);
    float y2 = u[1] * (0.49f - x * x);
    y2 = y2 > 0.0f ? y2 : 0.0f;
    float y = sqrtf(y2) * ((u[2] < 0.5f) ? -1.0f : 1.0f);
    float zz = 0.49f - x * x - y2;
    float z = sqrtf(zz > 0.0f ? zz : 0.0f) + 0.2f;
    float ca, sa;
    det_sincos(u[3] * 6.2831855f, sa, ca);
    const float Tf = 0.5f, D = 1.2f;
    float sx = x + D * ca, sy = y + D * sa, sz = z + 0.3f;
    S.ppos[0] = sx; S.ppos[1] = sy; S.ppos[2] = sz;
    S.pvel[0] = (x - sx) / Tf; S.pvel[1] = (y - sy) / Tf;
    S.pvel[2] = (z - sz) / Tf + 0.5f * 9.8f * Tf;
  }
  S.pquat[0] = S.pquat[1] = S.pquat[2] = 0.0f; S.pquat[3] = 1.0f;
  S.pang[0] = S.pang[1] = S.pang[2] = 0.0f;
  S.step = 0;
  S.flags = 0;
  S.fsm_state = 0; S.fsm_t = 0;
  S.rand_y = fmaf(u[3], 0.06f, -0.03f);
#pragma unroll
  for (int i = 0; i < 7; ++i) S.mtgt[i] = (i < ArmDims<ARM>::NARM) ? T.rest[i < ArmDims<ARM>::NARM ? i : 0] : 0.0f;
}

// observation Box(5): [finger1, finger2, EE x, EE y, EE z] in world coordinates (envs/peg_in_hole.py:13)
template <int ARM> PIH_D void observe(const EnvState& S, v3 off, float (&obs)[5]) {
  constexpr int nb = ArmDims<ARM>::NB, na = ArmDims<ARM>::NARM;
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  v3 pee; m3 Ree;
  ee_pose<ARM>(F, pee, Ree);
  obs[0] = (nb > na) ? S.q[na < PIH_NB_MAX ? na : 0] : 0.0f;
  obs[1] = (nb > na + 1) ? S.q[(na + 1) < PIH_NB_MAX ? na + 1 : 0] : 0.0f;
  obs[2] = pee.x + off.x; obs[3] = pee.y + off.y; obs[4] = pee.z + off.z;
}

// apply_action: panda_execute / ur_execute (envs/utils.py:60-82) -> joint position targets, gains and force caps.
// Scripted task: apply_action is `pass` (peg_in_hole.py:30-31); the command comes from the grasp state machine
// (peg_in_hole.py:39-212).  Returns true when the FSM has reached its final state.
template <int ARM, int TASK> PIH_D bool apply_action(const SimParams& P, EnvState& S, const FK<ARM>& F, const float* action, v3 off, float* qstar,
                                           float* kpc, float* mfc) {
  constexpr int nb = ArmDims<ARM>::NB, na = ArmDims<ARM>::NARM;
  const ArmTables& T = c_arm[ARM];
  v3 pee; m3 Ree;
  ee_pose<ARM>(F, pee, Ree);
#pragma unroll
  for (int j = 0; j < nb; ++j) { kpc[j] = T.kp[j]; mfc[j] = T.maxf[j]; }
  v3 tgt;
  float tq[4];
  if (TASK == PIH_TASK_SCRIPTED) {
    bool fsm_done = false;
    const int fsm_steps[PIH_FSM_NSTATES] = {PIH_FSM_STEPS_LIST};
    int lim = 0;
#pragma unroll
    for (int k = 0; k < PIH_FSM_NSTATES; ++k) lim = (S.fsm_state == k) ? fsm_steps[k] : lim;
    if (S.fsm_t == lim) {
      S.fsm_state += 1; S.fsm_t = 0;
      if (S.fsm_state >= PIH_FSM_NSTATES) S.fsm_state = 0;
    }
    S.fsm_t += 1;
    const int st = S.fsm_state;
    m3 Rp = quat_to_m3(S.pquat);
    v3 tp = mv_add(Rp, mk(0.0f, S.rand_y, 0.0f), mk(S.ppos[0], S.ppos[1], S.ppos[2]));
    float qx = S.pquat[0], qy = S.pquat[1], qz = S.pquat[2], qw = S.pquat[3];
    float yaw = det_atan2(2.0f * fmaf(qx, qy, qw * qz), fmaf(qw, qw, qx * qx) - fmaf(qy, qy, qz * qz));
    float cur[3] = {pee.x, pee.y, pee.z}, tar[3] = {tp.x, tp.y, tp.z}, res[3];
    bool arm_cmd = false, clamp = true;
    float yaw_cmd = 0.0f, zoff = 0.0f;
    if (st == 1 || st == 2) { arm_cmd = true; yaw_cmd = (float)1.5707963267948966 + yaw; zoff = (st == 1) ? 0.05f : -0.01f; }
    else if (st == 4 || st == 5 || st == 6) {
      arm_cmd = true; yaw_cmd = (float)(-3.141592653589793);
      tar[0] = (float)PIH_HOLE_X - ((st == 4) ? 0.2f : ((st == 5) ? 0.04f : 0.0f));
      tar[1] = (float)PIH_HOLE_Y; tar[2] = (float)PIH_HOLE_Z;
      clamp = (st != 6);
    } else if (st == 8) { arm_cmd = true; yaw_cmd = (float)1.5707963267948966; tar[0] = 0.2f; tar[1] = -0.6f; tar[2] = 0.4f; clamp = false; }
    else if (st == 9) fsm_done = true;
    if (arm_cmd) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float diff = tar[i] - cur[i];
        if (clamp && fabsf(diff) > (float)PIH_FSM_DV) res[i] = cur[i] + ((diff > 0.0f) ? (float)PIH_FSM_DV : (float)(-PIH_FSM_DV));
        else res[i] = cur[i] + diff;
      }
      tgt = mk(res[0], res[1], res[2] + zoff);
      euler_to_quat(0.0f, (float)(-3.141592653589793), yaw_cmd, tq);
      float qik[nb];
      ik_dls<ARM>(P, F, S.q, tgt, tq, qik);
#pragma unroll
      for (int j = 0; j < na && j < 7; ++j) S.mtgt[j] = qik[j];
    }
#pragma unroll
    for (int j = 0; j < na; ++j) { qstar[j] = S.mtgt[j < 7 ? j : 0]; kpc[j] = (float)PIH_FSM_KP; mfc[j] = (float)PIH_FSM_ARM_FORCE; }
    const bool closed = (st >= 3 && st <= 6);
#pragma unroll
    for (int j = na; j < nb; ++j) {
      qstar[j] = closed ? (float)PIH_FSM_FINGER_CLOSED : (float)PIH_FSM_FINGER_OPEN;
      kpc[j] = (float)PIH_FSM_KP;
      mfc[j] = closed ? (float)PIH_FSM_FINGER_FORCE_HI : (float)PIH_FSM_FINGER_FORCE_LO;
    }
    return fsm_done;
  }
  float ax = action[0], ay = action[1], az = action[2];
  if (!P.action_frame_env) { ax = ax - off.x; ay = ay - off.y; az = az - off.z; }
  if (ARM == PIH_ARM_PANDA) {
    float cur[3] = {pee.x, pee.y, pee.z}, tar[3] = {ax, ay, az}, res[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // vel_constraint, envs/utils.py:85-95
      float diff = tar[i] - cur[i];
      if (fabsf(diff) > (float)PIH_DV) res[i] = cur[i] + ((diff > 0.0f) ? (float)PIH_DV : (float)(-PIH_DV));
      else res[i] = cur[i] + diff;
    }
    tgt = mk(res[0], res[1], res[2]);
    tq[0] = 0.0f; tq[1] = -1.0f; tq[2] = 0.0f; tq[3] = 0.0f;
  } else {
    tgt = mk(ax, ay, az);
    euler_to_quat(action[3], action[4], action[5], tq);
  }
  ik_dls<ARM>(P, F, S.q, tgt, tq, qstar);
#pragma unroll
  for (int j = na; j < nb; ++j) qstar[j] = action[3];
  return false;
}

// One env.step for one env.  Returns false if the env was frozen (done and not reset): outputs untouched.
template <int ARM, int TASK> __device__ __noinline__ bool step_env(const SimParams& P, EnvState& S, const float* action, v3 off, unsigned gid,
                                                         float (&obs)[5], float& reward, unsigned char& done) {
  if (S.flags & PIH_FLAG_DONE) return false;
  constexpr int nb = ArmDims<ARM>::NB, na = ArmDims<ARM>::NARM;
  const ArmTables& T = c_arm[ARM];
  const float dt = P.dt, idt = P.idt;
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qstar[nb], kpc[nb], mfc[nb];
  const bool fsm_done = apply_action<ARM, TASK>(P, S, F, action, off, qstar, kpc, mfc);
  // ---- unconstrained accelerations
  float tau[nb];
#pragma unroll
  for (int i = 0; i < nb; ++i) tau[i] = -(T.damp[i] * S.qd[i]);
  ABA<ARM> AB;
  aba<ARM>(F, S.qd, tau, (float)PIH_GRAVITY, AB);
  float Minv[nb][nb];
  for (int j = 0; j < nb; ++j) {
    float col[nb];
    minv_column<ARM>(F, AB, j, col);
#pragma unroll
    for (int i = 0; i < nb; ++i) Minv[i][j] = col[i];
  }
  float vs[nb];
#pragma unroll
  for (int i = 0; i < nb; ++i) vs[i] = fmaf(dt, AB.qdd[i], S.qd[i]);
  m3 Rfb = quat_to_m3(S.pquat);
  v3 fv = mk(S.pvel[0], S.pvel[1], fmaf(dt, (float)(-PIH_GRAVITY), S.pvel[2]));
  v3 fw = mk(S.pang[0], S.pang[1], S.pang[2]);
  {
    v3 wb = mTv(Rfb, fw);
    v3 Iw = mk(P.fb_idiag[0] * wb.x, P.fb_idiag[1] * wb.y, P.fb_idiag[2] * wb.z);
    v3 g = vcross(wb, Iw);
    v3 al = mk(-(g.x * P.fb_iinv[0]), -(g.y * P.fb_iinv[1]), -(g.z * P.fb_iinv[2]));
    fw = clamp_norm(vaxpy(dt, mv(Rfb, al), fw));
  }
  m3 Iinv;
  {
    m3 Tm;
#pragma unroll
    for (int r = 0; r < 3; ++r) { Tm.m[r][0] = Rfb.m[r][0] * P.fb_iinv[0]; Tm.m[r][1] = Rfb.m[r][1] * P.fb_iinv[1]; Tm.m[r][2] = Rfb.m[r][2] * P.fb_iinv[2]; }
    Iinv = mmulT(Tm, Rfb);
  }
  const float im = P.fb_im;
  // ---- motor + limit rows
  float m_dinv[nb], m_b[nb], m_lam[nb], m_max[nb], l_b[2][nb], l_lam[2][nb];
  unsigned l_on = 0u;  // bit 2j: lower limit row of joint j present, bit 2j+1: upper
#pragma unroll
  for (int j = 0; j < nb; ++j) {
    // targets outside the URDF limits are unreachable: aim at the limit instead of fighting the limit row
    float qt = qstar[j];
    qt = (qt < T.lo[j]) ? T.lo[j] : ((qt > T.hi[j]) ? T.hi[j] : qt);
    float vstar = kpc[j] * (qt - S.q[j]) * idt;
    m_dinv[j] = 1.0f / Minv[j][j];
    m_b[j] = (vstar - vs[j]) * m_dinv[j];
    m_lam[j] = 0.0f;
    m_max[j] = mfc[j] * dt;
    // a limit row exists only if the joint can reach that limit this step (see oracle for the rule)
    float glo = S.q[j] - T.lo[j], ghi = T.hi[j] - S.q[j];
    float vdn = fmaxf(fmaxf(-vstar, -vs[j]), 0.0f), vup = fmaxf(fmaxf(vstar, vs[j]), 0.0f);
    l_on |= (glo < fmaf(vdn, dt, 0.005f) ? 1u : 0u) << (2 * j);
    l_on |= (ghi < fmaf(vup, dt, 0.005f) ? 1u : 0u) << (2 * j + 1);
    l_lam[0][j] = l_lam[1][j] = 0.0f;
    float tlo = (glo > 0.0f) ? -(glo * idt) : -((float)PIH_ERP * glo * idt);
    float thi = (ghi > 0.0f) ? -(ghi * idt) : -((float)PIH_ERP * ghi * idt);
    l_b[0][j] = (tlo - vs[j]) * m_dinv[j];
    l_b[1][j] = (thi + vs[j]) * m_dinv[j];
  }
  // ---- contacts
  ContactSet CS;
  collide<ARM, TASK>(P, S, F, Rfb, CS);
  CRow<ARM> rows[PIH_MAXC];
  int pairmask = 0;
  bool arm_contacts = false, coupled = false;
  v3 pfb = mk(S.ppos[0], S.ppos[1], S.ppos[2]);
  for (int k = 0; k < CS.n; ++k) {
    const Contact& c = CS.c[k];
    CRow<ARM>& r = rows[k];
    pairmask |= c.pair;
    arm_contacts = arm_contacts || (c.arm_body >= 0);
    coupled = coupled || (c.has_fb && c.arm_body >= 0);
    r.has_fb = c.has_fb; r.arm_body = c.arm_body; r.mu = c.mu;
    r.dir[0] = c.n;
    plane_space(c.n, r.dir[1], r.dir[2]);
    v3 rr = vsub(c.p, pfb);
    unsigned anc = (c.arm_body >= 0) ? anc_rt<ARM>(c.arm_body) : 0u;
    for (int d = 0; d < 3; ++d) {
      float diag = 0.0f, vrel = 0.0f;
      if (c.has_fb) {
        r.jang[d] = vcross(rr, r.dir[d]);
        r.mang[d] = mv(Iinv, r.jang[d]);
        diag = im + vdot(r.jang[d], r.mang[d]);
        vrel = vdot(r.dir[d], fv) + vdot(r.jang[d], fw);
      }
      if (c.arm_body >= 0) {
        float sg = c.has_fb ? -1.0f : 1.0f;
        float jrow[nb];
#pragma unroll
        for (int j = 0; j < nb; ++j) {
          float v = 0.0f;
          if ((anc >> j) & 1u) {
            v3 z = mcol(F.Rw[j], 2);
            if (jtype_of<ARM>(j) == 0) v = vdot(r.dir[d], vcross(z, vsub(c.p, F.pw[j])));
            else v = vdot(r.dir[d], z);
            v = sg * v;
          }
          jrow[j] = v;
          r.ja[d][j] = v;
        }
        float da = 0.0f, va = 0.0f;
        float mrow_[nb];
#pragma unroll
        for (int i = 0; i < nb; ++i) {
          float s = 0.0f;
#pragma unroll
          for (int j = 0; j < nb; ++j) s = fmaf(Minv[i][j], jrow[j], s);
          mrow_[i] = s;
          r.ma[d][i] = s;
        }
#pragma unroll
        for (int j = 0; j < nb; ++j) { da = fmaf(jrow[j], mrow_[j], da); va = fmaf(jrow[j], vs[j], va); }
        diag = diag + da;
        vrel = vrel + va;
      }
      r.dinv[d] = 1.0f / diag;
      r.diag[d] = diag;
      float tv = 0.0f;
      if (d == 0) tv = (c.dist > 0.0f) ? -(c.dist * idt) : -((float)PIH_ERP * c.dist * idt);
      r.b[d] = (tv - vrel) * r.dinv[d];
      r.lam[d] = 0.0f;
    }
  }
  // ---- projected Gauss-Seidel: motors, limits, contact normals, contact friction.  Early exit on PyBullet's
  //      solverResidualThreshold: largest squared velocity change (deltaImpulse / jacDiagABInv)^2 of the iteration,
  //      evaluated per simulation island (arm island / free-body island; merged by a finger<->free-body contact).
  float dva[nb];
#pragma unroll
  for (int j = 0; j < nb; ++j) dva[j] = 0.0f;
  v3 dvl = mk(0.0f, 0.0f, 0.0f), dvw = dvl;
  const int ncon = CS.n;
  bool arm_on = true, fb_on = true;
  for (int it = 0; it < P.solver_iters; ++it) {
    float res_a = 0.0f, res_f = 0.0f;
    if (arm_on) {
#pragma unroll kPgsUnroll
      for (int j = 0; j < nb; ++j) {
        float dl = fmaf(-dva[j], m_dinv[j], m_b[j]);
        float s = m_lam[j] + dl;
        s = fminf(fmaxf(s, -m_max[j]), m_max[j]);
        dl = s - m_lam[j];
        m_lam[j] = s;
        float rv = dl * Minv[j][j];
        res_a = fmaxf(res_a, rv * rv);
#pragma unroll
        for (int i = 0; i < nb; ++i) dva[i] = fmaf(Minv[i][j], dl, dva[i]);
      }
      if (l_on) {  // uniform-ish skip: most steps only the finger limit rows exist
#pragma unroll kPgsUnroll
        for (int j = 0; j < nb; ++j) {
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!((l_on >> (2 * j + side)) & 1u)) continue;
            float sg = side ? -1.0f : 1.0f;
            float dl = fmaf(-(sg * dva[j]), m_dinv[j], l_b[side][j]);
            float s = l_lam[side][j] + dl;
            s = fmaxf(s, 0.0f);
            dl = (s - l_lam[side][j]) * sg;
            l_lam[side][j] = s;
            float rv = dl * Minv[j][j];
            res_a = fmaxf(res_a, rv * rv);
#pragma unroll
            for (int i = 0; i < nb; ++i) dva[i] = fmaf(Minv[i][j], dl, dva[i]);
          }
        }
      }
    }
    if (!arm_contacts) {
      // common case: every contact is free body vs static geometry -> no per-row type branches, no arm coupling
      if (fb_on) {
        for (int k = 0; k < ncon; ++k) {
          CRow<ARM>& r = rows[k];
          float rel = vdot(r.dir[0], dvl) + vdot(r.jang[0], dvw);
          float dl = fmaf(-rel, r.dinv[0], r.b[0]);
          float s = fmaxf(r.lam[0] + dl, 0.0f);
          dl = s - r.lam[0];
          r.lam[0] = s;
          float rv = dl * r.diag[0];
          res_f = fmaxf(res_f, rv * rv);
          dvl = vaxpy(dl * im, r.dir[0], dvl);
          dvw = vaxpy(dl, r.mang[0], dvw);
        }
        for (int k = 0; k < ncon; ++k) {
          CRow<ARM>& r = rows[k];
          const float lim = r.mu * r.lam[0];
#pragma unroll
          for (int d = 1; d < 3; ++d) {
            float rel = vdot(r.dir[d], dvl) + vdot(r.jang[d], dvw);
            float dl = fmaf(-rel, r.dinv[d], r.b[d]);
            float s = fminf(fmaxf(r.lam[d] + dl, -lim), lim);
            dl = s - r.lam[d];
            r.lam[d] = s;
            float rv = dl * r.diag[d];
            res_f = fmaxf(res_f, rv * rv);
            dvl = vaxpy(dl * im, r.dir[d], dvl);
            dvw = vaxpy(dl, r.mang[d], dvw);
          }
        }
      }
    } else {
      for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < ncon; ++k) {
          CRow<ARM>& r = rows[k];
          const bool is_fb = r.has_fb && r.arm_body < 0;
          if (!coupled && !(is_fb ? fb_on : arm_on)) continue;
          for (int d = (pass ? 1 : 0); d < (pass ? 3 : 1); ++d) {
            float rel = 0.0f;
            if (r.has_fb) rel = vdot(r.dir[d], dvl) + vdot(r.jang[d], dvw);
            if (r.arm_body >= 0) {
              float ra = 0.0f;
#pragma unroll
              for (int j = 0; j < nb; ++j) ra = fmaf(r.ja[d][j], dva[j], ra);
              rel = rel + ra;
            }
            float dl = fmaf(-rel, r.dinv[d], r.b[d]);
            float s = r.lam[d] + dl;
            if (d == 0) s = fmaxf(s, 0.0f);
            else {
              float lim = r.mu * r.lam[0];
              s = fminf(fmaxf(s, -lim), lim);
            }
            dl = s - r.lam[d];
            r.lam[d] = s;
            float rv = dl * r.diag[d];
            if (is_fb) res_f = fmaxf(res_f, rv * rv); else res_a = fmaxf(res_a, rv * rv);
            if (r.has_fb) { dvl = vaxpy(dl * im, r.dir[d], dvl); dvw = vaxpy(dl, r.mang[d], dvw); }
            if (r.arm_body >= 0) {
#pragma unroll
              for (int j = 0; j < nb; ++j) dva[j] = fmaf(r.ma[d][j], dl, dva[j]);
            }
          }
        }
    }
    if (coupled) {
      if (fmaxf(res_a, res_f) <= P.solver_residual) break;
    } else {
      if (res_a <= P.solver_residual) arm_on = false;
      if (res_f <= P.solver_residual) fb_on = false;
      if (!arm_on && !fb_on) break;
    }
  }
  // ---- integrate
#pragma unroll
  for (int j = 0; j < nb; ++j) {
    float v = vs[j] + dva[j];
    v = (v > (float)PIH_MAX_COORD_VEL) ? (float)PIH_MAX_COORD_VEL : ((v < (float)(-PIH_MAX_COORD_VEL)) ? (float)(-PIH_MAX_COORD_VEL) : v);
    S.qd[j] = v;
    S.q[j] = fmaf(dt, S.qd[j], S.q[j]);
  }
  fv = clamp_norm(vadd(fv, dvl));
  fw = clamp_norm(vadd(fw, dvw));
  S.pvel[0] = fv.x; S.pvel[1] = fv.y; S.pvel[2] = fv.z;
  S.pang[0] = fw.x; S.pang[1] = fw.y; S.pang[2] = fw.z;
  S.ppos[0] = fmaf(dt, fv.x, S.ppos[0]); S.ppos[1] = fmaf(dt, fv.y, S.ppos[1]); S.ppos[2] = fmaf(dt, fv.z, S.ppos[2]);
  {
    float w2 = vdot(fw, fw);
    float wn = sqrtf(w2);
    float ha = 0.5f * dt * wn;
    float sh, ch, k;
    det_sincos(ha, sh, ch);
    if (wn > 1e-6f) k = sh / wn; else k = 0.5f * dt;
    float dq[4] = {fw.x * k, fw.y * k, fw.z * k, ch}, nq[4];
    quat_mul(dq, S.pquat, nq);
    float n2 = fmaf(nq[3], nq[3], fmaf(nq[2], nq[2], fmaf(nq[1], nq[1], nq[0] * nq[0])));
    float in = 1.0f / sqrtf(n2);
#pragma unroll
    for (int i = 0; i < 4; ++i) S.pquat[i] = nq[i] * in;
  }
  S.step += 1;
  // ---- get_info
  float rew = 0.0f;
  int success = 0;
  FK<ARM> F2;
  v3 pee2; m3 Ree2;
  fk<ARM, nb>(S.q, F2);
  ee_pose<ARM>(F2, pee2, Ree2);
  if (TASK != PIH_TASK_RANDOM_FLY) {
    v3 d = mk(S.ppos[0] - (float)PIH_HOLE_X, S.ppos[1] - (float)PIH_HOLE_Y, S.ppos[2] - (float)PIH_HOLE_Z);
    success = vdot(d, d) < (float)(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
  } else {
    v3 d = mk(S.ppos[0] - pee2.x, S.ppos[1] - pee2.y, S.ppos[2] - pee2.z);
    success = vdot(d, d) < (float)(PIH_SUCCESS_RADIUS * PIH_SUCCESS_RADIUS);
  }
  rew = success ? 1.0f : 0.0f;
  int dn = success || (P.max_steps > 0 && S.step >= P.max_steps);
  if (TASK == PIH_TASK_RANDOM_FLY && (pairmask & PIH_PAIR_PEG_TABLE)) dn = 1;
  if (TASK == PIH_TASK_SCRIPTED) dn = fsm_done || (P.max_steps > 0 && S.step >= P.max_steps);
  S.flags = (dn ? PIH_FLAG_DONE : 0) | (success ? PIH_FLAG_SUCCESS : 0) | (pairmask << PIH_PAIR_SHIFT);
  reward = rew;
  done = (unsigned char)dn;
  if (dn && P.auto_reset) {
    int keep = S.flags & ~PIH_FLAG_DONE;
    S.episode += 1;
    reset_env<ARM, TASK>(P, S, gid);
    S.flags = keep & ~0xff;
    observe<ARM>(S, off, obs);
  } else {
    obs[0] = (nb > na) ? S.q[na < PIH_NB_MAX ? na : 0] : 0.0f;
    obs[1] = (nb > na + 1) ? S.q[(na + 1) < PIH_NB_MAX ? na + 1 : 0] : 0.0f;
    obs[2] = pee2.x + off.x; obs[3] = pee2.y + off.y; obs[4] = pee2.z + off.z;
  }
  return true;
}

}  // namespace pih
