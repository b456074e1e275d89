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
// One copy per <arm, task, kernel family> (TAG): with a single out-of-line function shared by the kernels of every
// instantiation ptxas gives it one calling convention for all of them, and the register allocation of the callers' hot
// loops suffers (solver sweep 250 instead of 236 instructions as soon as a second <arm, task> was instantiated in the
// same module).
template <int TAG> static __device__ __noinline__ void contact_insert(ContactSet& S, const Contact& c) {
  if (S.n < PIH_MAXC) { S.c[S.n++] = c; return; }
  int w = 0;
  for (int i = 1; i < PIH_MAXC; ++i)
    if (S.c[i].dist > S.c[w].dist) w = i;
  if (c.dist < S.c[w].dist) S.c[w] = c;
}
template <int TAG> PIH_D void contact_push(ContactSet& S, const Contact& c) {
  if (c.dist < (float)PIH_CONTACT_MARGIN) contact_insert<TAG>(S, c);
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

// FSM state of THIS step as the dynamics sees it (update_state runs before the constraint is created / removed and before
// stepSimulation, peg_in_hole.py:55,99-108).  `advanced`: S already holds the updated state (apply_action ran in this thread).
template <int TASK> PIH_D int fsm_state_now(const EnvState& S, bool advanced) {
  if (TASK != PIH_TASK_SCRIPTED) return 0;
  if (advanced) return S.fsm_state;
  const int fsm_steps[PIH_FSM_NSTATES] = {PIH_FSM_STEPS_LIST};
  int lim = 0;
#pragma unroll
  for (int k = 0; k < PIH_FSM_NSTATES; ++k) lim = (S.fsm_state == k) ? fsm_steps[k] : lim;
  int st = S.fsm_state;
  if (S.fsm_t == lim) { st += 1; if (st >= PIH_FSM_NSTATES) st = 0; }
  return st;
}
PIH_D bool grasp_active(int task, int fsm_now) { return task == PIH_TASK_SCRIPTED && fsm_now >= PIH_FSM_GRASP_FIRST && fsm_now <= PIH_FSM_GRASP_LAST; }

// KID: which kernel family the call is compiled into (0 classic, 2 / 3 small-batch kernels): see contact_insert
template <int ARM, int TASK, int KID> PIH_D void collide(const SimParams& P, const EnvState& S, const FK<ARM>& F, const m3& Rfb, int fsm_now, ContactSet& CS) {
  const ArmTables& T = c_arm[ARM];
  CS.n = 0;
  v3 pfb = mk(S.ppos[0], S.ppos[1], S.ppos[2]);
  Contact c;
  // (0) grasp constraint of the scripted task (peg_in_hole.py:99-104; PIH_PAIR_GRASP in pih_constants.h): a pseudo contact
  //     whose three rows are bilateral and run along the env axes; p = end-effector origin, n carries the position error
  if (TASK == PIH_TASK_SCRIPTED && grasp_active(TASK, fsm_now)) {
    v3 pee; m3 Ree;
    ee_pose<ARM>(F, pee, Ree);
    v3 pc = mv_add(Rfb, mk(0.0f, S.rand_y, 0.0f), pfb);
    c.pair = PIH_PAIR_GRASP; c.has_fb = 1; c.arm_body = ArmDims<ARM>::EE; c.p = pee; c.n = vsub(pc, pee);
    c.dist = -1e30f; c.mu = -1.0f;
    contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
  }
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
          contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
          contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
        }
      }
    } else {
      for (int k = 0; k < 8; ++k) {
        v3 pl;
        fb_sample<TASK>(P, k, pl);
        v3 pt = mv_add(Rfb, pl, pfb);
        c.pair = PIH_PAIR_PEG_TABLE; c.has_fb = 1; c.arm_body = -1; c.p = pt; c.n = mk(0.0f, 0.0f, 1.0f);
        c.dist = pt.z - (float)PIH_TABLE_Z; c.mu = mu_static;
        contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
        contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
        contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
      contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
        contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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
    contact_push<ARM * 8 + TASK + 64 * KID>(CS, c);
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

// ---- contact rows of one PGS sweep of a warp in which some env couples its islands (finger <-> free body): every lane
// runs the same straight-line body per row - free-body part, arm part, projection, both updates - with the row's island
// flag folded into the interval, so the 32 lanes never diverge whatever mix of contact types they hold; values equal
// the oracle's per-type formulas (rows are zero-filled: the absent terms are exact zeros).  The coefficients of row
// i + 1 are fetched while row i is being solved (a lone warp otherwise waits for an L1 round trip or three in front of
// every row); used where the caller can spare two 35-word sets of registers (pih_step_body.inc: kPrefetchRows).
template <int ARM> struct GRowsIO {
  float dva[ArmDims<ARM>::NB];
  v3 dvl, dvw;
  float res_a, res_f, im, dt;
  int ncon;
  bool arm_on, fb_on;
};
template <int ARM> struct GRow {
  v3 dir, jang, mang;
  float ja[ArmDims<ARM>::NB], ma[ArmDims<ARM>::NB], dinv, b, diag, lam, lam0, mu;
  int arm_body;
};
template <int ARM> PIH_D void grow_load(const CRow<ARM>* rows, int k, int d, GRow<ARM>& R) {
  const CRow<ARM>& r = rows[k];
  R.dir = r.dir[d]; R.jang = r.jang[d]; R.mang = r.mang[d];
#pragma unroll
  for (int j = 0; j < ArmDims<ARM>::NB; ++j) { R.ja[j] = r.ja[d][j]; R.ma[j] = r.ma[d][j]; }
  R.dinv = r.dinv[d]; R.b = r.b[d]; R.diag = r.diag[d]; R.lam = r.lam[d]; R.lam0 = r.lam[0]; R.mu = r.mu;
  R.arm_body = r.arm_body;
}
template <int ARM, int TASK> PIH_D void general_rows_sweep(CRow<ARM>* rows, GRowsIO<ARM>& io) {
  constexpr int nb = ArmDims<ARM>::NB;
  const int ncon = io.ncon;
  if (ncon == 0) return;
  float dva[nb];
#pragma unroll
  for (int j = 0; j < nb; ++j) dva[j] = io.dva[j];
  v3 dvl = io.dvl, dvw = io.dvw;
  float res_a = io.res_a, res_f = io.res_f;
  const float im = io.im, dt = io.dt;
  const bool arm_on = io.arm_on, fb_on = io.fb_on;
  auto solve = [&](const GRow<ARM>& R, float& lam_out, auto normal_) {
    constexpr bool NORMAL = decltype(normal_)::value;
    const bool is_fb = R.arm_body < 0;
    const bool on = is_fb ? fb_on : arm_on;
    const float relfb = vdot(R.dir, dvl) + vdot(R.jang, dvw);
    float s0 = 0.0f, s1 = 0.0f;   // J_arm . dva as two interleaved partial sums (even / odd joints), as in the oracle
#pragma unroll
    for (int j = 0; j < nb; j += 2) s0 = fmaf(R.ja[j], dva[j], s0);
#pragma unroll
    for (int j = 1; j < nb; j += 2) s1 = fmaf(R.ja[j], dva[j], s1);
    const float rel = relfb + (s0 + s1);
    float dl = fmaf(-rel, R.dinv, R.b);
    const float lam = R.lam;
    float dlo, dhi;
    if (NORMAL) { dlo = -lam; dhi = __int_as_float(0x7f800000); }
    else { const float lim = R.mu * R.lam0; dlo = -lim - lam; dhi = lim - lam; }
    if (TASK == PIH_TASK_SCRIPTED && R.mu < 0.0f) {   // grasp constraint: bilateral, |impulse| <= maxForce dt
      const float wl = (float)PIH_GRASP_MAX_FORCE * dt;
      dlo = -wl - lam; dhi = wl - lam;
    }
    dl = fminf(fmaxf(dl, on ? dlo : 0.0f), on ? dhi : 0.0f);
    lam_out = lam + dl;
    const float rv = fabsf(dl * R.diag);   // (residuals are collected as |dv| and squared once per sweep by the caller)
    res_f = fmaxf(res_f, is_fb ? rv : 0.0f);
    res_a = fmaxf(res_a, is_fb ? 0.0f : rv);
    dvl = vaxpy(dl * im, R.dir, dvl);
    dvw = vaxpy(dl, R.mang, dvw);
#pragma unroll
    for (int j = 0; j < nb; ++j) dva[j] = fmaf(R.ma[j], dl, dva[j]);
  };
  // two sets of coefficients, A and B, used alternately (no copies): while one row is solved the next one loads
  GRow<ARM> A, B;
  grow_load<ARM>(rows, 0, 0, A);
  for (int k = 0; k < ncon; k += 2) {         // normals, in contact order
    const bool two = k + 1 < ncon;
    grow_load<ARM>(rows, two ? k + 1 : k, 0, B);
    solve(A, rows[k].lam[0], std::true_type{});
    if (!two) break;
    grow_load<ARM>(rows, k + 2 < ncon ? k + 2 : k + 1, 0, A);
    solve(B, rows[k + 1].lam[0], std::true_type{});
  }
  grow_load<ARM>(rows, 0, 1, A);              // (after the normals: a friction row's interval is mu * its normal impulse)
  for (int k = 0; k < ncon; ++k) {            // friction rows: directions 1 and 2 of contact k
    grow_load<ARM>(rows, k, 2, B);
    solve(A, rows[k].lam[1], std::false_type{});
    grow_load<ARM>(rows, k + 1 < ncon ? k + 1 : k, 1, A);
    solve(B, rows[k].lam[2], std::false_type{});
  }
#pragma unroll
  for (int j = 0; j < nb; ++j) io.dva[j] = dva[j];
  io.dvl = dvl; io.dvw = dvw; io.res_a = res_a; io.res_f = res_f;
}

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
    // random_pos_in_panda_space (envs/utils.py:97-107): x ~ U(-.7,.7), y = (sqrt(U(0,.49-x^2)) - U(0,.4)) * (+-1), drawn
    // again while .49 - x^2 - y^2 < 0; z = sqrt(.49 - x^2 - y^2) + 0.2.  Attempt k draws from Philox stream FLY + k.
    float x = 0.0f, y = 0.0f, zz = 0.49f;
    for (int k = 0; k < PIH_FLY_MAX_ATTEMPTS; ++k) {
      float v[4];
      rng4(P, gid, (unsigned)S.episode, PIH_STREAM_FLY + k, v);
      const float xk = fmaf(v[0], 1.4f, -0.7f);
      const float a2 = fmaf(-xk, xk, 0.49f);
      const float yk = (sqrtf(v[1] * (a2 > 0.0f ? a2 : 0.0f)) - v[2] * 0.4f) * ((v[3] < 0.5f) ? -1.0f : 1.0f);
      const float rem = fmaf(-yk, yk, a2);
      if (!(rem < 0.0f)) { x = xk; y = yk; zz = rem; break; }
    }
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
  S.grasp_ang = (float)PIH_LABEL_NONE;
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
// the scene of the scripted task's grasp image (PIH_S_LATCH_*): written straight to the env's state slots 12..15
PIH_D void latch_scene(const SimParams& P, int e, const EnvState& S) {
  float w[16];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = S.q[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[9 + k] = S.ppos[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[12 + k] = S.pquat[k];
#pragma unroll
  for (int s = 0; s < 4; ++s) P.state[(size_t)(12 + s) * P.n + e] = make_float4(w[4 * s], w[4 * s + 1], w[4 * s + 2], w[4 * s + 3]);
}

// latch_env: env index whose grasp scene may be latched by this call (-1: none - parity hooks, frozen or padding lanes)
template <int ARM, int TASK> PIH_D bool apply_action(const SimParams& P, EnvState& S, const FK<ARM>& F, const float* action, v3 off, float* qstar,
                                           float* kpc, float* mfc, int& ik_n, int latch_env = -1) {
  ik_n = 0;
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
    const bool entered = (S.fsm_t == lim);
    if (entered) {
      S.fsm_state += 1; S.fsm_t = 0;
      if (S.fsm_state >= PIH_FSM_NSTATES) S.fsm_state = 0;
    }
    S.fsm_t += 1;
    const int st = S.fsm_state;
    m3 Rp = quat_to_m3(S.pquat);
    // entering state 2 the reference renders the grasp image and labels it with the rectangle whose angle is
    // atan2 of R(q_peg) random_vector = rand_y * (second column of R) (peg_in_hole.py:58,65-72)
    if (entered && st == 2) {
      S.grasp_ang = det_atan2(S.rand_y * Rp.m[1][1], S.rand_y * Rp.m[0][1]);
      if (latch_env >= 0) latch_scene(P, latch_env, S);       // self.grasp_img = self.render() (peg_in_hole.py:65-66)
    }
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
      ik_dls<ARM>(P, F, S.q, tgt, tq, qik, ik_n);
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
  ik_dls<ARM>(P, F, S.q, tgt, tq, qstar, ik_n);
#pragma unroll
  for (int j = na; j < nb; ++j) qstar[j] = action[3];
  return false;
}

// Mailbox from the action-application warp to the dynamics warp of the warp-specialised kernels (one column per lane).
#define PIH_MB_WORDS 40
#define PIH_MB_QSTAR 0
#define PIH_MB_KP 9
#define PIH_MB_MAXF 18
#define PIH_MB_FSM_DONE 27
#define PIH_MB_FSM_STATE 28
#define PIH_MB_FSM_T 29
#define PIH_MB_MTGT 30
#define PIH_MB_GRASP_ANG 37

// ---- three-warp small-batch kernels: exchange block in shared memory (one column per lane) and the named barriers
// between the dynamics warp (warp 1) and the contacts warp (warp 2); barrier 0 (__syncthreads) is the hand-over of the
// joint targets from the action warp (warp 0)
#define PIH_WS_THREADS 96
#define PIH_BAR_ABA 1     // dynamics -> contacts: the ABA factorisation is in X.ab
#define PIH_BAR_FRONT 2   // contacts -> dynamics: contact summary / solved free-body island are in X.front, X.armc
#define PIH_BAR_COLS 3    // contacts -> dynamics: the M^-1 columns [COLS_D, NB) are in X.cols
template <int ARM> struct WsXch {
  static constexpr int NB = ArmDims<ARM>::NB;
  static constexpr int COLS_D = (NB + 1) / 2;   // M^-1 columns [0, COLS_D) are computed by the dynamics warp itself
  float ab[NB][13][32];                         // per body: U (6), U / D (6), 1 / D
  float cols[NB - COLS_D][NB][32];
  float front[12][32];                          // contacts, pair mask, row-iterations, iterations, dv linear (3), dv angular (3)
  int armc;                                     // some env of the warp has a contact that involves the arm
};
#define PIH_XF_N 0
#define PIH_XF_PAIRS 1
#define PIH_XF_ROWIT 2
#define PIH_XF_IT 3
#define PIH_XF_DVL 4
#define PIH_XF_DVW 7
// producer side: make the shared-memory stores visible, then arrive without waiting; consumer side: wait
PIH_D void bar_arrive(int id) { __threadfence_block(); asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
PIH_D void bar_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// Contacts warp of the small-batch kernels.  While the dynamics warp runs the articulated-body algorithm this warp does
// the narrow phase and - when no env of the warp has a contact that involves the arm, the usual case - builds the rows
// of the free-body island (free body vs static geometry) and solves it: that island depends neither on M^-1 nor on the
// joint targets.  Otherwise it only raises X.armc and the dynamics warp takes its general path (collide + all rows).
// Then it takes the ABA factorisation from shared memory and computes its share of the M^-1 columns.  Same operations in
// the same order as the single-warp body (pih_step_body.inc), so every value is the oracle's, bit for bit.
template <int ARM, int TASK> PIH_D void contacts_warp(const SimParams& P, const EnvState& S, const FK<ARM>& F, WsXch<ARM>& X, int lane) {
  constexpr int nb = ArmDims<ARM>::NB;
  const float dt = P.dt, idt = P.idt, im = P.fb_im;
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
  ContactSet CS;
  CS.n = 0;
  // Scripted task: while the grasp constraint is on in some env of the warp (states 4..6; the state machine is
  // time-driven, so normally in all of them) that env has a row on the arm - the constraint itself - and the dynamics
  // warp will run the narrow phase and build every row anyway: skip the narrow phase here, so that this warp's M^-1
  // columns are ready as soon as the factorisation is and the dynamics warp does not wait for a result it discards.
  bool w_grasp = false;
  if constexpr (TASK == PIH_TASK_SCRIPTED) w_grasp = __any_sync(0xffffffffu, grasp_active(TASK, fsm_state_now<TASK>(S, false)));
  if (!w_grasp) collide<ARM, TASK, 3>(P, S, F, Rfb, fsm_state_now<TASK>(S, false), CS);
  int pairmask = 0;
  bool arm_contacts = false;
  for (int k = 0; k < CS.n; ++k) { pairmask |= CS.c[k].pair; arm_contacts = arm_contacts || (CS.c[k].arm_body >= 0); }
  const bool w_armc = w_grasp || __any_sync(0xffffffffu, arm_contacts);
  v3 dvl = mk(0.0f, 0.0f, 0.0f), dvw = dvl;
  int st_rowit = 0, st_it = 0;
  if (!w_armc) {
    struct FbRow { v3 dir[3], jang[3], mang[3]; float dinv[3], diag[3], b[3], lam[3], mu; } rows[PIH_MAXC];
    const v3 pfb = mk(S.ppos[0], S.ppos[1], S.ppos[2]);
    for (int k = 0; k < CS.n; ++k) {
      const Contact& c = CS.c[k];
      FbRow& r = rows[k];
      r.mu = c.mu;
      r.dir[0] = c.n; plane_space(c.n, r.dir[1], r.dir[2]);
      v3 rr = vsub(c.p, pfb);
      for (int d = 0; d < 3; ++d) {
        r.jang[d] = vcross(rr, r.dir[d]);
        r.mang[d] = mv(Iinv, r.jang[d]);
        const float diag = im + vdot(r.jang[d], r.mang[d]);
        const float vrel = vdot(r.dir[d], fv) + vdot(r.jang[d], fw);
        r.dinv[d] = 1.0f / diag;
        r.diag[d] = diag;
        float tv = 0.0f;
        if (d == 0) tv = (c.dist > 0.0f) ? -(c.dist * idt) : -((float)PIH_ERP * c.dist * idt);
        r.b[d] = (tv - vrel) * r.dinv[d];
        r.lam[d] = 0.0f;
      }
    }
    bool fb_on = true;
    for (int it = 0; it < P.solver_iters; ++it) {
      float res_f = 0.0f;
      st_rowit += fb_on ? 3 * CS.n : 0; st_it += fb_on ? 1 : 0;
      if (fb_on) {
        for (int k = 0; k < CS.n; ++k) {
          FbRow& r = rows[k];
          float rel = vdot(r.dir[0], dvl) + vdot(r.jang[0], dvw);
          float dl = fmaf(-rel, r.dinv[0], r.b[0]);
          dl = fmaxf(dl, -r.lam[0]);
          r.lam[0] += dl;
          float rv = dl * r.diag[0];
          res_f = fmaxf(res_f, fabsf(rv));
          dvl = vaxpy(dl * im, r.dir[0], dvl);
          dvw = vaxpy(dl, r.mang[0], dvw);
        }
        for (int k = 0; k < CS.n; ++k) {
          FbRow& r = rows[k];
          const float lim = r.mu * r.lam[0];
#pragma unroll
          for (int d = 1; d < 3; ++d) {
            float rel = vdot(r.dir[d], dvl) + vdot(r.jang[d], dvw);
            float dl = fmaf(-rel, r.dinv[d], r.b[d]);
            dl = fminf(fmaxf(dl, -lim - r.lam[d]), lim - r.lam[d]);
            r.lam[d] += dl;
            float rv = dl * r.diag[d];
            res_f = fmaxf(res_f, fabsf(rv));
            dvl = vaxpy(dl * im, r.dir[d], dvl);
            dvw = vaxpy(dl, r.mang[d], dvw);
          }
        }
      }
      res_f = res_f * res_f;   // (the rows collected max |dv|; rounded squaring is monotone: the oracle's max dv^2 exactly)
      if (res_f <= P.solver_residual) fb_on = false;
      if (!__any_sync(0xffffffffu, fb_on)) break;
    }
  }
  X.front[PIH_XF_N][lane] = __int_as_float(CS.n); X.front[PIH_XF_PAIRS][lane] = __int_as_float(pairmask);
  X.front[PIH_XF_ROWIT][lane] = __int_as_float(st_rowit); X.front[PIH_XF_IT][lane] = __int_as_float(st_it);
  X.front[PIH_XF_DVL][lane] = dvl.x; X.front[PIH_XF_DVL + 1][lane] = dvl.y; X.front[PIH_XF_DVL + 2][lane] = dvl.z;
  X.front[PIH_XF_DVW][lane] = dvw.x; X.front[PIH_XF_DVW + 1][lane] = dvw.y; X.front[PIH_XF_DVW + 2][lane] = dvw.z;
  if (lane == 0) X.armc = w_armc ? 1 : 0;
  PIH_TICK(10);
  bar_arrive(PIH_BAR_FRONT);
  bar_wait(PIH_BAR_ABA);
  ABA<ARM> AB;
#pragma unroll
  for (int i = 0; i < nb; ++i) {
    AB.U[i].a = mk(X.ab[i][0][lane], X.ab[i][1][lane], X.ab[i][2][lane]); AB.U[i].l = mk(X.ab[i][3][lane], X.ab[i][4][lane], X.ab[i][5][lane]);
    AB.UD[i].a = mk(X.ab[i][6][lane], X.ab[i][7][lane], X.ab[i][8][lane]); AB.UD[i].l = mk(X.ab[i][9][lane], X.ab[i][10][lane], X.ab[i][11][lane]);
    AB.Dinv[i] = X.ab[i][12][lane];
  }
  constexpr int c0 = WsXch<ARM>::COLS_D;
  for (int j = c0; j < nb; ++j) {
    float col[nb];
    minv_column<ARM>(F, AB, j, col);
#pragma unroll
    for (int i = 0; i < nb; ++i) X.cols[j - c0][i][lane] = col[i];
  }
  PIH_TICK(11);
  bar_arrive(PIH_BAR_COLS);
}

// Everything of one env.step after apply_action: stepSimulation (ABA, M^-1, contacts, PGS, integration) and get_info.
// The body lives in pih_step_body.inc and is instantiated twice: the classic single-warp order and the warp-specialised
// order, in which the joint targets arrive from the sibling warp through the shared-memory mailbox `mb` at the one
// point where they are needed (motor rows), after this warp has already done the target-independent work.
template <int ARM, int TASK>
__device__ __forceinline__ void step_back(const SimParams& P, EnvState& S, const FK<ARM>& F, float* qstar, float* kpc, float* mfc, bool fsm_done,
                                          bool commit, v3 off, unsigned gid, float (&obs)[5], float& reward, unsigned char& done) {
#define PIH_STEP_WS 0
#include "pih_step_body.inc"
#undef PIH_STEP_WS
}
template <int ARM, int TASK, int NW>
__device__ __forceinline__ void step_back_ws(const SimParams& P, EnvState& S, const FK<ARM>& F, float* qstar, float* kpc, float* mfc, bool fsm_done,
                                             float (*mb)[32], WsXch<ARM>& X, int lane, bool commit, v3 off, unsigned gid, float (&obs)[5],
                                             float& reward, unsigned char& done) {
#define PIH_STEP_WS 1
#include "pih_step_body.inc"
#undef PIH_STEP_WS
}

// One env.step for one env.  Returns false if the env was frozen (done and not reset): outputs untouched.
template <int ARM, int TASK> __device__ __noinline__ bool step_env(const SimParams& P, EnvState& S, const float* action, v3 off, unsigned gid,
                                                              float (&obs)[5], float& reward, unsigned char& done) {
  const int latch_env = (int)(gid - (unsigned)P.env_id_base);   // local env index
  if (S.flags & PIH_FLAG_DONE) return false;
  constexpr int nb = ArmDims<ARM>::NB;
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qstar[nb], kpc[nb], mfc[nb];
  int ik_n;
  const bool fsm_done = apply_action<ARM, TASK>(P, S, F, action, off, qstar, kpc, mfc, ik_n, latch_env);
  stat_add(P, PIH_STAT_IK_ITERS, (unsigned)ik_n);
  step_back<ARM, TASK>(P, S, F, qstar, kpc, mfc, fsm_done, true, off, gid, obs, reward, done);
  return true;
}

// one env's action row: a single 128-bit load for Box(4), three 64-bit loads for Box(6) (rows are 16 B / 24 B apart)
template <int A> PIH_D void load_action(const float* __restrict__ action, int e, float (&a)[6]) {
  if (A == 4) {
    float4 v = __ldg(reinterpret_cast<const float4*>(action) + e);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; a[4] = 0.0f; a[5] = 0.0f;
  } else {
#pragma unroll
    for (int k = 0; k < A / 2; ++k) {
      float2 v = __ldg(reinterpret_cast<const float2*>(action) + (size_t)(A / 2) * e + k);
      a[2 * k] = v.x; a[2 * k + 1] = v.y;
    }
  }
}

}  // namespace pih
