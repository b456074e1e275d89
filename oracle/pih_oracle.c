/* CPU ORACLE driver - TEST INFRASTRUCTURE ONLY (see pih_oracle_impl.h header: PARITY UNPINNED vs PyBullet).
 * Batched wrapper around the per-env restatement: BaseEnvMp.step / BaseEnv.step loop semantics
 * (peg_in_hole_gym/envs/base_env_mp.py:40-63, peg_in_hole_gym/envs/base_env.py:35-94). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pih.h"
#include "../peg_in_hole_gym_b200/csrc/pih_constants.h"
#include "../peg_in_hole_gym_b200/csrc/pih_model_data.h"

/* diagnostics only (not thread-safe counters; approximate under threads) */
unsigned long long pih_dbg_pgs_iters = 0, pih_dbg_pgs_steps = 0;

/* (set by pih_oracle_capture_step; read by step_env) */
#define REAL float
#define SUFFIX _f
#define REAL_IS_DOUBLE 0
#include "pih_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE

#define REAL double
#define SUFFIX _d
#define REAL_IS_DOUBLE 1
#include "pih_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE

typedef struct {
  pih_config cfg;
  int use_double;
  int was_reset;
  model_f mf;
  model_d md;
  simcfg_f cf;
  simcfg_d cd;
  envstate_f* sf;
  envstate_d* sd;
  pipemodel_f pmf;      /* flexible pipe (cfg.peg_model == PIH_PEG_PIPE) */
  pipemodel_d pmd;
  pipestate_f* pf;
  pipestate_d* pd;
  float* off; /* [n][3] */
  uint32_t global_step;
  int* last_nc; /* contacts found by the last step of each env (diagnostics) */
} oracle;

/* BaseEnv._create_env (base_env.py:35-55): line layout offset*i when offset.x==0 or offset.y==0, otherwise a
 * ceil(sqrt(n)) square grid (sx*i, sy*j, oz), row-major in (i,j).  The layout repeats for every mp client. */
void pih_oracle_env_offset(const pih_config* cfg, int local_env, float out[3]) {
  int sub = cfg->sub_num > 0 ? cfg->sub_num : cfg->num_envs;
  int e = (cfg->env_id_base + local_env) % sub;
  if (cfg->offset[0] == 0.0f || cfg->offset[1] == 0.0f) {
    out[0] = cfg->offset[0] * (float)e; out[1] = cfg->offset[1] * (float)e; out[2] = cfg->offset[2] * (float)e;
  } else {
    int s = 0;
    while (s * s < sub) ++s;
    int i = e / s, j = e % s;
    out[0] = cfg->offset[0] * (float)i; out[1] = cfg->offset[1] * (float)j; out[2] = cfg->offset[2];
  }
}

void* pih_oracle_create(const pih_config* cfg, int use_double) {
  oracle* o = (oracle*)calloc(1, sizeof(oracle));
  o->cfg = *cfg;
  o->use_double = use_double;
  int n = cfg->num_envs;
  model_init_f(&o->mf, cfg->arm, cfg->task, cfg->fly_object);
  model_init_d(&o->md, cfg->arm, cfg->task, cfg->fly_object);
  o->cf.task = cfg->task; o->cf.arm = cfg->arm; o->cf.fly_object = cfg->fly_object; o->cf.dt = cfg->dt; o->cf.solver_residual = cfg->solver_residual;
  o->cf.solver_iters = cfg->solver_iters; o->cf.ik_iters = cfg->ik_iters; o->cf.max_steps = cfg->max_steps;
  o->cf.auto_reset = cfg->auto_reset; o->cf.action_frame_env = cfg->action_frame_env; o->cf.seed = cfg->seed;
  o->cd.task = cfg->task; o->cd.arm = cfg->arm; o->cd.fly_object = cfg->fly_object; o->cd.dt = (double)cfg->dt; o->cd.solver_residual = (double)cfg->solver_residual;
  o->cd.solver_iters = cfg->solver_iters; o->cd.ik_iters = cfg->ik_iters; o->cd.max_steps = cfg->max_steps;
  o->cd.auto_reset = cfg->auto_reset; o->cd.action_frame_env = cfg->action_frame_env; o->cd.seed = cfg->seed;
  o->sf = (envstate_f*)calloc(n, sizeof(envstate_f));
  o->sd = (envstate_d*)calloc(n, sizeof(envstate_d));
  pipemodel_init_f(&o->pmf); pipemodel_init_d(&o->pmd);
  o->pf = (pipestate_f*)calloc(n, sizeof(pipestate_f));
  o->pd = (pipestate_d*)calloc(n, sizeof(pipestate_d));
  for (int i = 0; i < n; ++i) { o->pf[i].bquat[3] = 1.0f; o->pd[i].bquat[3] = 1.0; }
  o->off = (float*)calloc(3 * (size_t)n, sizeof(float));
  o->last_nc = (int*)calloc(n, sizeof(int));
  for (int i = 0; i < n; ++i) {
    pih_oracle_env_offset(cfg, i, &o->off[3 * i]);
    o->sf[i].episode = -1; o->sd[i].episode = -1;
    o->sf[i].pquat[3] = 1.0f; o->sd[i].pquat[3] = 1.0;
  }
  return o;
}

void pih_oracle_destroy(void* h) {
  oracle* o = (oracle*)h;
  if (!o) return;
  free(o->sf); free(o->sd); free(o->pf); free(o->pd); free(o->off); free(o->last_nc); free(o);
}

int pih_oracle_reset(void* h, const uint8_t* mask, float* obs) {
  oracle* o = (oracle*)h;
  int n = o->cfg.num_envs;
  for (int i = 0; i < n; ++i) {
    if (mask && !mask[i]) continue;
    uint32_t gid = (uint32_t)(o->cfg.env_id_base + i);
    if (o->use_double) {
      o->sd[i].episode += 1;
      if (o->cfg.peg_model == PIH_PEG_PIPE) reset_env_pipe_d(&o->md, &o->cd, &o->sd[i], &o->pd[i], gid);
      else reset_env_d(&o->md, &o->cd, &o->sd[i], gid);
      if (obs) {
        double ob[5], off[3] = {o->off[3 * i], o->off[3 * i + 1], o->off[3 * i + 2]};
        observe_d(&o->md, &o->cd, &o->sd[i], off, ob);
        for (int k = 0; k < 5; ++k) obs[5 * i + k] = (float)ob[k];
      }
    } else {
      o->sf[i].episode += 1;
      if (o->cfg.peg_model == PIH_PEG_PIPE) reset_env_pipe_f(&o->mf, &o->cf, &o->sf[i], &o->pf[i], gid);
      else reset_env_f(&o->mf, &o->cf, &o->sf[i], gid);
      if (obs) observe_f(&o->mf, &o->cf, &o->sf[i], &o->off[3 * i], &obs[5 * i]);
    }
  }
  o->was_reset = 1;
  return 0;
}

static int step_one(oracle* o, int i, const float* act, float* obs, float* reward, uint8_t* done) {
  uint32_t gid = (uint32_t)(o->cfg.env_id_base + i);
  if (o->use_double) {
    double a[6], ob[5], rw = 0, off[3] = {o->off[3 * i], o->off[3 * i + 1], o->off[3 * i + 2]};
    for (int k = 0; k < 6; ++k) a[k] = act[k];
    uint8_t dn = 0;
    if (o->sd[i].flags & PIH_FLAG_DONE) return 0;
    int nc = (o->cfg.peg_model == PIH_PEG_PIPE) ? step_env_pipe_d(&o->md, &o->pmd, &o->cd, &o->sd[i], &o->pd[i], a, off, gid, ob, &rw, &dn)
                                               : step_env_d(&o->md, &o->cd, &o->sd[i], a, off, gid, ob, &rw, &dn);
    o->last_nc[i] = nc;
    if (obs) for (int k = 0; k < 5; ++k) obs[k] = (float)ob[k];
    if (reward) *reward = (float)rw;
    if (done) *done = dn;
    return nc;
  } else {
    float ob[5], rw = 0;
    uint8_t dn = 0;
    if (o->sf[i].flags & PIH_FLAG_DONE) return 0;
    int nc = (o->cfg.peg_model == PIH_PEG_PIPE) ? step_env_pipe_f(&o->mf, &o->pmf, &o->cf, &o->sf[i], &o->pf[i], act, &o->off[3 * i], gid, ob, &rw, &dn)
                                               : step_env_f(&o->mf, &o->cf, &o->sf[i], act, &o->off[3 * i], gid, ob, &rw, &dn);
    o->last_nc[i] = nc;
    if (obs) for (int k = 0; k < 5; ++k) obs[k] = ob[k];
    if (reward) *reward = rw;
    if (done) *done = dn;
    return nc;
  }
}

int pih_oracle_step(void* h, const float* action, float* obs, float* reward, uint8_t* done) {
  oracle* o = (oracle*)h;
  if (!o->was_reset) return PIH_ERR_STATE;
  int n = o->cfg.num_envs, A = (o->cfg.arm == PIH_ARM_PANDA) ? 4 : 6;
  for (int i = 0; i < n; ++i) {
    float a[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < A; ++k) a[k] = action[A * i + k];
    step_one(o, i, a, obs ? &obs[5 * i] : 0, reward ? &reward[i] : 0, done ? &done[i] : 0);
  }
  o->global_step += 1;
  return 0;
}

/* synthetic rollout: actions U[-1,1]^A from Philox (seed; env, global step, stream ACTION) - same generator as the
 * CUDA rollout kernel.  Processes envs [lo,hi) so that the caller may thread over env ranges. */
int pih_oracle_rollout_range(void* h, int steps, int lo, int hi, float* obs, float* reward, uint8_t* done, unsigned long long* counters) {
  oracle* o = (oracle*)h;
  if (!o->was_reset) return PIH_ERR_STATE;
  int A = (o->cfg.arm == PIH_ARM_PANDA) ? 4 : 6;
  unsigned long long nsteps = 0, ndone = 0;
  for (int i = lo; i < hi; ++i) {
    uint32_t gid = (uint32_t)(o->cfg.env_id_base + i);
    for (int t = 0; t < steps; ++t) {
      float u[8], a[6] = {0, 0, 0, 0, 0, 0};
      uint32_t gs = o->global_step + (uint32_t)t;
      pih_rng4(o->cfg.seed, gid, gs, PIH_STREAM_ACTION, u);
      if (A > 4) pih_rng4(o->cfg.seed, gid, gs, PIH_STREAM_ACTION + 1, u + 4);
      for (int k = 0; k < A; ++k) a[k] = fmaf(u[k], 2.0f, -1.0f);
      if (o->cfg.action_frame_env == 0) { /* world-absolute actions: centre the sample on the env origin */
        a[0] += o->off[3 * i]; a[1] += o->off[3 * i + 1]; a[2] += o->off[3 * i + 2];
      }
      int frozen = o->use_double ? (o->sd[i].flags & PIH_FLAG_DONE) : (o->sf[i].flags & PIH_FLAG_DONE);
      if (frozen) continue;
      uint8_t dn = 0;
      step_one(o, i, a, obs ? &obs[5 * i] : 0, reward ? &reward[i] : 0, &dn);
      if (done) done[i] = dn;
      nsteps += 1;
      ndone += dn;
    }
  }
  if (counters) { __atomic_fetch_add(&counters[0], nsteps, __ATOMIC_RELAXED); __atomic_fetch_add(&counters[1], ndone, __ATOMIC_RELAXED); }
  return 0;
}
void pih_oracle_advance_global_step(void* h, int steps) { ((oracle*)h)->global_step += (uint32_t)steps; }

int pih_oracle_rollout(void* h, int steps, float* obs, float* reward, uint8_t* done, unsigned long long* counters) {
  oracle* o = (oracle*)h;
  int rc = pih_oracle_rollout_range(h, steps, 0, o->cfg.num_envs, obs, reward, done, counters);
  if (rc == 0) o->global_step += (uint32_t)steps;
  return rc;
}

void pih_oracle_last_contacts(void* h, int* out) {
  oracle* o = (oracle*)h;
  for (int i = 0; i < o->cfg.num_envs; ++i) out[i] = o->last_nc[i];
}

/* state exchange in the device layout: words[slot][env][4] (float4 SoA, pih_constants.h) */
#define WIDX(w, i, n) ((size_t)((w) >> 2) * (size_t)(n) * 4 + (size_t)(i) * 4 + ((w) & 3))
int pih_oracle_state_words(const pih_config* cfg) {
  return cfg->peg_model == PIH_PEG_PIPE ? 4 * PIH_PIPE_STATE_F4 : (cfg->task == PIH_TASK_SCRIPTED ? 4 * PIH_SCRIPTED_STATE_F4 : PIH_STATE_WORDS);
}

int pih_oracle_get_state(void* h, float* words) {
  oracle* o = (oracle*)h;
  int n = o->cfg.num_envs;
  memset(words, 0, sizeof(float) * (size_t)pih_oracle_state_words(&o->cfg) * (size_t)n);
  if (pih_oracle_state_words(&o->cfg) > PIH_S_LATCH_Q)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 16; ++k) words[WIDX(PIH_S_LATCH_Q + k, i, n)] = o->use_double ? (float)o->sd[i].latch[k] : o->sf[i].latch[k];
  if (o->cfg.peg_model == PIH_PEG_PIPE)
    for (int i = 0; i < n; ++i) {
      const int D = o->use_double;
      for (int k = 0; k < 3; ++k) {
        words[WIDX(PIH_PS_POS + k, i, n)] = D ? (float)o->pd[i].bpos[k] : o->pf[i].bpos[k];
        words[WIDX(PIH_PS_VEL + k, i, n)] = D ? (float)o->pd[i].bvel[k] : o->pf[i].bvel[k];
        words[WIDX(PIH_PS_ANG + k, i, n)] = D ? (float)o->pd[i].bang[k] : o->pf[i].bang[k];
      }
      for (int k = 0; k < 4; ++k) words[WIDX(PIH_PS_QUAT + k, i, n)] = D ? (float)o->pd[i].bquat[k] : o->pf[i].bquat[k];
      int32_t gi = D ? o->pd[i].grasp_idx : o->pf[i].grasp_idx;
      memcpy(&words[WIDX(PIH_PS_GRASP, i, n)], &gi, 4);
      for (int b = 1; b < PIH_PIPE_NBODY; ++b) {
        words[WIDX(PIH_PS_QJ + b - 1, i, n)] = D ? (float)o->pd[i].qj[b] : o->pf[i].qj[b];
        words[WIDX(PIH_PS_QDJ + b - 1, i, n)] = D ? (float)o->pd[i].qdj[b] : o->pf[i].qdj[b];
      }
    }
  for (int i = 0; i < n; ++i) {
    int32_t ints[3];
    for (int k = 0; k < PIH_NB_MAX; ++k) {
      words[WIDX(PIH_S_Q + k, i, n)] = o->use_double ? (float)o->sd[i].q[k] : o->sf[i].q[k];
      words[WIDX(PIH_S_QD + k, i, n)] = o->use_double ? (float)o->sd[i].qd[k] : o->sf[i].qd[k];
    }
    for (int k = 0; k < 3; ++k) {
      words[WIDX(PIH_S_PPOS + k, i, n)] = o->use_double ? (float)o->sd[i].ppos[k] : o->sf[i].ppos[k];
      words[WIDX(PIH_S_PVEL + k, i, n)] = o->use_double ? (float)o->sd[i].pvel[k] : o->sf[i].pvel[k];
      words[WIDX(PIH_S_PANG + k, i, n)] = o->use_double ? (float)o->sd[i].pang[k] : o->sf[i].pang[k];
    }
    for (int k = 0; k < 4; ++k) words[WIDX(PIH_S_PQUAT + k, i, n)] = o->use_double ? (float)o->sd[i].pquat[k] : o->sf[i].pquat[k];
    {
      int32_t fs = o->use_double ? o->sd[i].fsm_state : o->sf[i].fsm_state, ft = o->use_double ? o->sd[i].fsm_t : o->sf[i].fsm_t;
      memcpy(&words[WIDX(PIH_S_FSM_STATE, i, n)], &fs, 4);
      memcpy(&words[WIDX(PIH_S_FSM_T, i, n)], &ft, 4);
      words[WIDX(PIH_S_RANDY, i, n)] = o->use_double ? (float)o->sd[i].rand_y : o->sf[i].rand_y;
      words[WIDX(PIH_S_GRASP_ANG, i, n)] = o->use_double ? (float)o->sd[i].grasp_ang : o->sf[i].grasp_ang;
      for (int k = 0; k < 7; ++k) words[WIDX(PIH_S_MTGT + k, i, n)] = o->use_double ? (float)o->sd[i].mtgt[k] : o->sf[i].mtgt[k];
    }
    ints[0] = o->use_double ? o->sd[i].step : o->sf[i].step;
    ints[1] = o->use_double ? o->sd[i].episode : o->sf[i].episode;
    ints[2] = o->use_double ? o->sd[i].flags : o->sf[i].flags;
    memcpy(&words[WIDX(PIH_S_STEP, i, n)], &ints[0], 4);
    memcpy(&words[WIDX(PIH_S_EPISODE, i, n)], &ints[1], 4);
    memcpy(&words[WIDX(PIH_S_FLAGS, i, n)], &ints[2], 4);
  }
  return 0;
}

int pih_oracle_set_state(void* h, const float* words) {
  oracle* o = (oracle*)h;
  int n = o->cfg.num_envs;
  if (pih_oracle_state_words(&o->cfg) > PIH_S_LATCH_Q)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 16; ++k) { o->sf[i].latch[k] = words[WIDX(PIH_S_LATCH_Q + k, i, n)]; o->sd[i].latch[k] = o->sf[i].latch[k]; }
  if (o->cfg.peg_model == PIH_PEG_PIPE)
    for (int i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) {
        o->pf[i].bpos[k] = words[WIDX(PIH_PS_POS + k, i, n)]; o->pd[i].bpos[k] = o->pf[i].bpos[k];
        o->pf[i].bvel[k] = words[WIDX(PIH_PS_VEL + k, i, n)]; o->pd[i].bvel[k] = o->pf[i].bvel[k];
        o->pf[i].bang[k] = words[WIDX(PIH_PS_ANG + k, i, n)]; o->pd[i].bang[k] = o->pf[i].bang[k];
      }
      for (int k = 0; k < 4; ++k) { o->pf[i].bquat[k] = words[WIDX(PIH_PS_QUAT + k, i, n)]; o->pd[i].bquat[k] = o->pf[i].bquat[k]; }
      int32_t gi;
      memcpy(&gi, &words[WIDX(PIH_PS_GRASP, i, n)], 4);
      o->pf[i].grasp_idx = o->pd[i].grasp_idx = gi;
      for (int b = 1; b < PIH_PIPE_NBODY; ++b) {
        o->pf[i].qj[b] = words[WIDX(PIH_PS_QJ + b - 1, i, n)]; o->pd[i].qj[b] = o->pf[i].qj[b];
        o->pf[i].qdj[b] = words[WIDX(PIH_PS_QDJ + b - 1, i, n)]; o->pd[i].qdj[b] = o->pf[i].qdj[b];
      }
    }
  for (int i = 0; i < n; ++i) {
    int32_t ints[3];
    memcpy(&ints[0], &words[WIDX(PIH_S_STEP, i, n)], 4);
    memcpy(&ints[1], &words[WIDX(PIH_S_EPISODE, i, n)], 4);
    memcpy(&ints[2], &words[WIDX(PIH_S_FLAGS, i, n)], 4);
    for (int k = 0; k < PIH_NB_MAX; ++k) {
      o->sf[i].q[k] = words[WIDX(PIH_S_Q + k, i, n)]; o->sd[i].q[k] = o->sf[i].q[k];
      o->sf[i].qd[k] = words[WIDX(PIH_S_QD + k, i, n)]; o->sd[i].qd[k] = o->sf[i].qd[k];
    }
    for (int k = 0; k < 3; ++k) {
      o->sf[i].ppos[k] = words[WIDX(PIH_S_PPOS + k, i, n)]; o->sd[i].ppos[k] = o->sf[i].ppos[k];
      o->sf[i].pvel[k] = words[WIDX(PIH_S_PVEL + k, i, n)]; o->sd[i].pvel[k] = o->sf[i].pvel[k];
      o->sf[i].pang[k] = words[WIDX(PIH_S_PANG + k, i, n)]; o->sd[i].pang[k] = o->sf[i].pang[k];
    }
    for (int k = 0; k < 4; ++k) { o->sf[i].pquat[k] = words[WIDX(PIH_S_PQUAT + k, i, n)]; o->sd[i].pquat[k] = o->sf[i].pquat[k]; }
    {
      int32_t fs, ft;
      memcpy(&fs, &words[WIDX(PIH_S_FSM_STATE, i, n)], 4);
      memcpy(&ft, &words[WIDX(PIH_S_FSM_T, i, n)], 4);
      o->sf[i].fsm_state = o->sd[i].fsm_state = fs; o->sf[i].fsm_t = o->sd[i].fsm_t = ft;
      o->sf[i].rand_y = words[WIDX(PIH_S_RANDY, i, n)]; o->sd[i].rand_y = o->sf[i].rand_y;
      o->sf[i].grasp_ang = words[WIDX(PIH_S_GRASP_ANG, i, n)]; o->sd[i].grasp_ang = o->sf[i].grasp_ang;
      for (int k = 0; k < 7; ++k) { o->sf[i].mtgt[k] = words[WIDX(PIH_S_MTGT + k, i, n)]; o->sd[i].mtgt[k] = o->sf[i].mtgt[k]; }
    }
    o->sf[i].step = o->sd[i].step = ints[0];
    o->sf[i].episode = o->sd[i].episode = ints[1];
    o->sf[i].flags = o->sd[i].flags = ints[2];
  }
  o->was_reset = 1;
  return 0;
}

/* ------------------------------------------------------------------ unit-test hooks (known-answer tests) */
void pih_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { pih_philox4x32_10(ctr, key, out); }
void pih_oracle_rng4(uint64_t seed, uint32_t env, uint32_t index, uint32_t stream, float u[4]) { pih_rng4(seed, env, index, stream, u); }

/* FK: end-effector position (env frame) and rotation for joint vector q (getLinkState) */
void pih_oracle_fk(int arm, int use_double, const double* q, double* pos, double* rot) {
  if (use_double) {
    model_d M; model_init_d(&M, arm, 0, 0);
    fkres_d F; double qq[PIH_NB_MAX] = {0};
    for (int i = 0; i < M.nb; ++i) qq[i] = q[i];
    fk_d(&M, qq, M.nb, &F);
    v3_d p; m3_d R; ee_pose_d(&M, &F, &p, &R);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    for (int i = 0; i < 9; ++i) rot[i] = R.m[i / 3][i % 3];
  } else {
    model_f M; model_init_f(&M, arm, 0, 0);
    fkres_f F; float qq[PIH_NB_MAX] = {0};
    for (int i = 0; i < M.nb; ++i) qq[i] = (float)q[i];
    fk_f(&M, qq, M.nb, &F);
    v3_f p; m3_f R; ee_pose_f(&M, &F, &p, &R);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    for (int i = 0; i < 9; ++i) rot[i] = R.m[i / 3][i % 3];
  }
}

/* IK: joint targets for an end-effector pose target (calculateInverseKinematics) */
void pih_oracle_ik(int arm, int use_double, int iters, const double* q, const double* tpos, const double* tquat, double* qout) {
  if (use_double) {
    model_d M; model_init_d(&M, arm, 0, 0);
    simcfg_d C; memset(&C, 0, sizeof(C)); C.ik_iters = iters;
    fkres_d F; double qq[PIH_NB_MAX] = {0}, qo[PIH_NB_MAX] = {0};
    for (int i = 0; i < M.nb; ++i) qq[i] = q[i];
    fk_d(&M, qq, M.nb, &F);
    ik_dls_d(&M, &C, &F, qq, v3_make_d(tpos[0], tpos[1], tpos[2]), tquat, qo);
    for (int i = 0; i < M.narm; ++i) qout[i] = qo[i];
  } else {
    model_f M; model_init_f(&M, arm, 0, 0);
    simcfg_f C; memset(&C, 0, sizeof(C)); C.ik_iters = iters;
    fkres_f F; float qq[PIH_NB_MAX] = {0}, qo[PIH_NB_MAX] = {0}, tq[4] = {(float)tquat[0], (float)tquat[1], (float)tquat[2], (float)tquat[3]};
    for (int i = 0; i < M.nb; ++i) qq[i] = (float)q[i];
    fk_f(&M, qq, M.nb, &F);
    ik_dls_f(&M, &C, &F, qq, v3_make_f((float)tpos[0], (float)tpos[1], (float)tpos[2]), tq, qo);
    for (int i = 0; i < M.narm; ++i) qout[i] = qo[i];
  }
}

/* dynamics hooks (double): joint accelerations from ABA, and M^-1 from the impulse-response pass */
void pih_oracle_aba(int arm, const double* q, const double* qd, const double* tau, double g, double* qdd, double* minv) {
  model_d M; model_init_d(&M, arm, 0, 0);
  fkres_d F; abares_d R;
  double qq[PIH_NB_MAX] = {0}, vv[PIH_NB_MAX] = {0}, tt[PIH_NB_MAX] = {0};
  for (int i = 0; i < M.nb; ++i) { qq[i] = q[i]; vv[i] = qd[i]; tt[i] = tau[i]; }
  fk_d(&M, qq, M.nb, &F);
  aba_d(&M, &F, vv, tt, g, &R);
  for (int i = 0; i < M.nb; ++i) qdd[i] = R.qdd[i];
  if (minv)
    for (int j = 0; j < M.nb; ++j) {
      double col[PIH_NB_MAX];
      minv_column_d(&M, &F, &R, j, col);
      for (int i = 0; i < M.nb; ++i) minv[i * M.nb + j] = col[i];
    }
}

/* model dump for independent checks in the tests (mass matrix via CRBA in numpy) */
int pih_oracle_model_nb(int arm) { return arm == PIH_ARM_PANDA ? PIH_PANDA_NB : PIH_UR5_NB; }

void pih_oracle_sincos(float x, float* s, float* c) { pih_sincosf_det(x, s, c); }
float pih_oracle_atan2_q1(float y, float x) { return det_atan2_q1_f(y, x); }

/* Label images of random_grasp (peg_in_hole.py:41-45,65-98) for envs [lo, hi): labels float[hi-lo][4][H][W] =
 * (pos_img, sin_img, cos_img, wid_img), params float[hi-lo][5] = (x, y, angle in degrees, width, length).  The rectangle
 * a, c, b, d of peg_in_hole.py:78-83 (first coordinate <-> column, second <-> row, because the reference writes
 * pos_img[cc, rr]) is filled where the integer pixel lies inside the closed quad (skimage.draw.polygon [3P]: interior
 * plus boundary).  Float arithmetic, same operation order as the CUDA kernel. */
int pih_oracle_grasp_labels(void* h, int W, int H, int lo, int hi, float* labels, float* params) {
  oracle* o = (oracle*)h;
  for (int i = lo; i < hi; ++i) {
    float ang = o->use_double ? (float)o->sd[i].grasp_ang : o->sf[i].grasp_ang;
    int any = ang > (float)(0.5 * PIH_LABEL_NONE);
    float sa, ca, s2, c2;
    det_sincos_f(any ? ang : 0.0f, &sa, &ca);
    det_sincos_f(any ? 2.0f * ang : 0.0f, &s2, &c2);
    const float L = (float)PIH_LABEL_LENGTH, Wd = (float)PIH_LABEL_WIDTH, S0 = (float)W, S1 = (float)H;
    float vx[4], vy[4];
    vx[0] = (1.0f + L * ca + Wd * sa) / 2.0f * S0; vy[0] = (1.0f - L * sa + Wd * ca) / 2.0f * S1;
    vx[1] = (1.0f - L * ca + Wd * sa) / 2.0f * S0; vy[1] = (1.0f + L * sa + Wd * ca) / 2.0f * S1;
    vx[2] = (1.0f - L * ca - Wd * sa) / 2.0f * S0; vy[2] = (1.0f + L * sa - Wd * ca) / 2.0f * S1;
    vx[3] = (1.0f + L * ca - Wd * sa) / 2.0f * S0; vy[3] = (1.0f - L * sa - Wd * ca) / 2.0f * S1;
    float dx = vx[0] - vx[3], dy = vy[0] - vy[3];
    float wid = sqrtf(fmaf(dx, dx, dy * dy));
    dx = vx[0] - vx[1]; dy = vy[0] - vy[1];
    float len = sqrtf(fmaf(dx, dx, dy * dy));
    if (params) {
      float* p = params + 5 * (size_t)(i - lo);
      p[0] = 0.0f; p[1] = 0.0f;
      p[2] = any ? ang / (float)3.141592653589793 * 180.0f : 0.0f;
      p[3] = any ? wid : 0.0f; p[4] = any ? len : 0.0f;
    }
    float* img = labels + (size_t)(i - lo) * 4 * H * W;
    const size_t np = (size_t)H * W;
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        int pos = 1, neg = 1;
        for (int k = 0; k < 4; ++k) {
          int m = (k + 1) & 3;
          float cr = fmaf(vx[m] - vx[k], (float)r - vy[k], -((vy[m] - vy[k]) * ((float)c - vx[k])));
          pos = pos && (cr >= 0.0f); neg = neg && (cr <= 0.0f);
        }
        int in = any && (pos || neg);
        size_t px = (size_t)r * W + c;
        img[px] = in ? (float)PIH_LABEL_POS_VALUE : 0.0f;
        img[np + px] = in ? s2 : 0.0f;
        img[2 * np + px] = in ? c2 : 1.0f;
        img[3 * np + px] = in ? wid : 0.0f;
      }
  }
  return 0;
}

/* PegInHole.render for envs [lo, hi): out float[hi-lo][H][W][4] = (depth, r, g, b).  latched != 0: the scene captured
 * when the state machine entered state 2 (the reference's grasp_img, peg_in_hole.py:65-66) instead of the present one. */
static int render_range(oracle* o, int W, int H, int lo, int hi, int latched, float* out);
int pih_oracle_render(void* h, int W, int H, int lo, int hi, float* out) { return render_range((oracle*)h, W, H, lo, hi, 0, out); }
int pih_oracle_render_grasp(void* h, int W, int H, int lo, int hi, float* out) { return render_range((oracle*)h, W, H, lo, hi, 1, out); }
static int render_range(oracle* o, int W, int H, int lo, int hi, int latched, float* out) {
  for (int i = lo; i < hi; ++i) {
    float* dst = out + (size_t)(i - lo) * H * W * 4;
    envstate_f vf = o->sf[i];
    envstate_d vd = o->sd[i];
    if (latched) {
      const float* lf = o->sf[i].latch; const double* ld = o->sd[i].latch;
      if (lf[12] != 0.0f || lf[13] != 0.0f || lf[14] != 0.0f || lf[15] != 0.0f) {
        for (int k = 0; k < 9; ++k) vf.q[k] = lf[k];
        for (int k = 0; k < 3; ++k) vf.ppos[k] = lf[9 + k];
        for (int k = 0; k < 4; ++k) vf.pquat[k] = lf[12 + k];
      }
      if (ld[12] != 0.0 || ld[13] != 0.0 || ld[14] != 0.0 || ld[15] != 0.0) {
        for (int k = 0; k < 9; ++k) vd.q[k] = ld[k];
        for (int k = 0; k < 3; ++k) vd.ppos[k] = ld[9 + k];
        for (int k = 0; k < 4; ++k) vd.pquat[k] = ld[12 + k];
      }
    }
    if (o->use_double) {
      scene_d SC; build_scene_d(&o->md, &vd, &SC);
      for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) render_pixel_d(&o->md, &o->cd, &SC, r, c, W, H, dst + ((size_t)r * W + c) * 4);
    } else {
      scene_f SC; build_scene_f(&o->mf, &vf, &SC);
      for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) render_pixel_f(&o->mf, &o->cf, &SC, r, c, W, H, dst + ((size_t)r * W + c) * 4);
    }
  }
  return 0;
}

/* flexible-pipe dynamics hook (double): unconstrained velocities after dt and the unit-impulse responses col[i][j]
 * (tests pin them against an independent numpy model: mass matrix from body Jacobians + recursive Newton-Euler bias) */
void pih_oracle_pipe_dyn(const double* bquat, const double* bvel, const double* bang, const double* qj, const double* qdj, double dt, double g,
                         double* vstar, double* col) {
  pipemodel_d PM; pipemodel_init_d(&PM);
  pipestate_d PS; memset(&PS, 0, sizeof(PS));
  for (int k = 0; k < 4; ++k) PS.bquat[k] = bquat[k];
  for (int k = 0; k < 3; ++k) { PS.bvel[k] = bvel[k]; PS.bang[k] = bang[k]; }
  for (int b = 1; b < PIH_PIPE_NBODY; ++b) { PS.qj[b] = qj[b - 1]; PS.qdj[b] = qdj[b - 1]; }
  pipefk_d K; pipe_fk_d(&PS, &K);
  pipedyn_d D; pipe_dynamics_d(&PM, &PS, &K, dt, g, &D);
  for (int i = 0; i < PIH_PIPE_NDOF; ++i) {
    vstar[i] = D.vstar[i];
    for (int j = 0; j < PIH_PIPE_NDOF; ++j) col[i * PIH_PIPE_NDOF + j] = D.col[i][j];
  }
}
/* world poses of the 24 pipe bodies for a given state: pos[24][3], rot[24][9] */
void pih_oracle_pipe_fk(const double* bpos, const double* bquat, const double* qj, double* pos, double* rot) {
  pipestate_d PS; memset(&PS, 0, sizeof(PS));
  for (int k = 0; k < 4; ++k) PS.bquat[k] = bquat[k];
  for (int k = 0; k < 3; ++k) PS.bpos[k] = bpos[k];
  for (int b = 1; b < PIH_PIPE_NBODY; ++b) PS.qj[b] = qj[b - 1];
  pipefk_d K; pipe_fk_d(&PS, &K);
  for (int b = 0; b < PIH_PIPE_NBODY; ++b) {
    pos[3 * b] = K.pw[b].x; pos[3 * b + 1] = K.pw[b].y; pos[3 * b + 2] = K.pw[b].z;
    for (int i = 0; i < 9; ++i) rot[9 * b + i] = K.Rw[b].m[i / 3][i % 3];
  }
}


/* test hook: steps env `i` once with `action` and returns the assembled problem + solution of that step (see pih_capture) */
pih_capture* pih_dbg_capture = 0;
int pih_oracle_capture_step(void* h, int i, const float* action, pih_capture* out) {
  oracle* o = (oracle*)h;
  if (o->cfg.peg_model == PIH_PEG_PIPE) return PIH_ERR_INVALID;
  float a[6] = {0, 0, 0, 0, 0, 0}, ob[5], rw;
  uint8_t dn;
  int A = (o->cfg.arm == PIH_ARM_PANDA) ? 4 : 6;
  for (int k = 0; k < A; ++k) a[k] = action[k];
  pih_dbg_capture = out;
  step_one(o, i, a, ob, &rw, &dn);
  pih_dbg_capture = 0;
  return 0;
}
int pih_oracle_capture_size(void) { return (int)sizeof(pih_capture); }
