// pih_kernels.cu - kernels + C ABI (include/pih.h) of the B200-native batched peg-in-hole simulator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC -shared
// There is NO CPU fallback in this library: every entry point drives a CUDA kernel or fails with PIH_ERR_CUDA.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>

#include "../../include/pih.h"
#include "pih_render.cuh"
#include "pih_pipe_host.h"

namespace pih {

// register budget of the step kernels: PIH_MIN_BLOCKS CTAs of 128 threads per SM (1 -> 255 regs, 4 -> 128 regs)
#ifndef PIH_MIN_BLOCKS
#define PIH_MIN_BLOCKS 1
#endif

// ------------------------------------------------------------------------------------------------ kernels
// thread-per-env; grid = ceil(n / blockDim).  State is read once and written once per launch as float4 SoA.

template <int ARM, int TASK> __global__ void __launch_bounds__(128) k_reset(SimParams P, const unsigned char* __restrict__ mask, float* __restrict__ obs) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  if (mask && !mask[e]) return;
  EnvState S;
  load_state(P, e, S);
  S.episode += 1;
  reset_env<ARM, TASK>(P, S, (unsigned)(P.env_id_base + e));
  store_state(P, e, S);
  if (obs) {
    float o[5];
    observe<ARM>(S, env_offset(P, e), o);
#pragma unroll
    for (int k = 0; k < 5; ++k) obs[5 * (size_t)e + k] = o[k];
  }
}

// A done env is not stepped and keeps its last outputs (base_env.py:61-71).  Those outputs are a function of the frozen
// state (observation of the post-step state, reward = success flag, done = 1), so they are re-emitted from it: the
// caller may hand in fresh output buffers every step (the zero-copy host path does).
template <int ARM> PIH_D void frozen_outputs(const EnvState& S, v3 off, float (&o)[5], float& r, unsigned char& d) {
  observe<ARM>(S, off, o);
  r = (S.flags & PIH_FLAG_SUCCESS) ? 1.0f : 0.0f;
  d = 1;
}

// A warp's 32 observation rows are 640 contiguous bytes: transposed through shared memory so that the warp writes five
// full 128-byte lines instead of 160 scattered words (what counts when obs is peer memory over NVLink or mapped host
// memory).  A partial warp (tail of the batch) stores its rows directly.
PIH_D void store_obs_rows(float* __restrict__ obs, int e, const float (&o)[5], float* so_block) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  if (m != 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < 5; ++k) obs[5 * (size_t)e + k] = o[k];
    return;
  }
  float* so = so_block + 160 * (threadIdx.x >> 5);
#pragma unroll
  for (int k = 0; k < 5; ++k) so[5 * lane + k] = o[k];
  __syncwarp();
  float* dst = obs + 5 * (size_t)(e - lane);
#pragma unroll
  for (int k = 0; k < 5; ++k) dst[32 * k + lane] = so[32 * k + lane];
}

template <int ARM, int TASK> __global__ void __launch_bounds__(128, PIH_MIN_BLOCKS) k_step(SimParams P, const float* __restrict__ action, float* __restrict__ obs,
                                                                float* __restrict__ reward, unsigned char* __restrict__ done) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int A = ArmDims<ARM>::ADIM;
  EnvState S;
  load_state(P, e, S);
  float a[6];
  load_action<A>(action, e, a);
  float o[5], r;
  unsigned char d;
  if (step_env<ARM, TASK>(P, S, a, env_offset(P, e), (unsigned)(P.env_id_base + e), o, r, d)) store_state(P, e, S);
  else frozen_outputs<ARM>(S, env_offset(P, e), o, r, d);
  __shared__ float so_block[4 * 160];
  __syncwarp(__activemask());   // (stepped and frozen lanes meet again here)
  if (obs) store_obs_rows(obs, e, o, so_block);
  if (reward) reward[e] = r;
  if (done) done[e] = d;
}

// ---- large batches, split step: apply_action (FK -> vel_constraint -> IK) needs far fewer registers than the dynamics,
// so it runs as its own kernel at 2-3x the occupancy and hands the joint targets over through a small SoA scratch
// buffer ([nb][n] floats, 36 B per env); the dynamics kernel is the classic step without the IK code.  Same functions,
// same arithmetic as the fused k_step.  The scripted task keeps the fused kernel (its state machine lives in apply_action).
#ifndef PIH_ACTION_MIN_BLOCKS
#define PIH_ACTION_MIN_BLOCKS 3
#endif
template <int ARM, int TASK> __global__ void __launch_bounds__(128, PIH_ACTION_MIN_BLOCKS) k_apply_action(SimParams P, const float* __restrict__ action,
                                                                                                   float* __restrict__ qt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int A = ArmDims<ARM>::ADIM, nb = ArmDims<ARM>::NB;
  float a[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  const v3 off = env_offset(P, e);
  if (action) load_action<A>(action, e, a);
  else {   // synthetic rollouts: action_space.sample() ~ U[-1,1]^A from Philox (seed; env, global step), as in k_rollout
    const unsigned gid = (unsigned)(P.env_id_base + e);
    float u[8], u0[4];
    rng4(P, gid, P.global_step, PIH_STREAM_ACTION, u0);
    u[0] = u0[0]; u[1] = u0[1]; u[2] = u0[2]; u[3] = u0[3];
    if (A > 4) { rng4(P, gid, P.global_step, PIH_STREAM_ACTION + 1, u0); u[4] = u0[0]; u[5] = u0[1]; u[6] = u0[2]; u[7] = u0[3]; }
#pragma unroll
    for (int k = 0; k < A; ++k) a[k] = fmaf(u[k], 2.0f, -1.0f);
    if (!P.action_frame_env) { a[0] += off.x; a[1] += off.y; a[2] += off.z; }
  }
  EnvState S;
  load_state(P, e, S);
  if (S.flags & PIH_FLAG_DONE) return;     // frozen env: the dynamics kernel will not read its targets
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qs[nb], kpc[nb], mfc[nb];
  int ik_n;
  apply_action<ARM, TASK>(P, S, F, a, off, qs, kpc, mfc, ik_n);
  stat_add(P, PIH_STAT_IK_ITERS, (unsigned)ik_n);
#pragma unroll
  for (int j = 0; j < nb; ++j) qt[(size_t)j * P.n + e] = qs[j];
}

template <int ARM, int TASK> __global__ void __launch_bounds__(128, PIH_MIN_BLOCKS) k_step_dyn(SimParams P, const float* __restrict__ qt, float* __restrict__ obs,
                                                                    float* __restrict__ reward, unsigned char* __restrict__ done,
                                                                    unsigned long long* __restrict__ counters) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int nb = ArmDims<ARM>::NB;
  const ArmTables& T = c_arm[ARM];
  EnvState S;
  load_state(P, e, S);
  const v3 off = env_offset(P, e);
  float o[5], r;
  unsigned char d;
  unsigned nsteps = 0, ndone = 0;
  if (S.flags & PIH_FLAG_DONE) frozen_outputs<ARM>(S, off, o, r, d);
  else {
    FK<ARM> F;
    fk<ARM, nb>(S.q, F);
    float qs[nb], kpc[nb], mfc[nb];
#pragma unroll
    for (int j = 0; j < nb; ++j) { qs[j] = __ldg(&qt[(size_t)j * P.n + e]); kpc[j] = T.kp[j]; mfc[j] = T.maxf[j]; }
    step_back<ARM, TASK>(P, S, F, qs, kpc, mfc, false, true, off, (unsigned)(P.env_id_base + e), o, r, d);
    store_state(P, e, S);
    nsteps = 1; ndone = d;
  }
  if (counters) {   // rollout accounting: one atomic pair per warp
    const unsigned m = __activemask();
    unsigned ws = __reduce_add_sync(m, nsteps), wd = __reduce_add_sync(m, ndone);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1) && ws) { atomicAdd(&counters[0], (unsigned long long)ws); atomicAdd(&counters[1], (unsigned long long)wd); }
  }
  __shared__ float so_block[4 * 160];
  __syncwarp(__activemask());
  if (obs) store_obs_rows(obs, e, o, so_block);
  if (reward) reward[e] = r;
  if (done) done[e] = d;
}

template <int ARM, int TASK> __global__ void __launch_bounds__(128, PIH_MIN_BLOCKS) k_rollout(SimParams P, int steps, float* __restrict__ obs, float* __restrict__ reward,
                                                                   unsigned char* __restrict__ done, unsigned long long* __restrict__ counters) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int A = ArmDims<ARM>::ADIM;
  EnvState S;
  load_state(P, e, S);
  const v3 off = env_offset(P, e);
  const unsigned gid = (unsigned)(P.env_id_base + e);
  float o[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, r = 0.0f;
  unsigned char d = 0;
  bool any = false;
  unsigned nsteps = 0, ndone = 0;
  for (int t = 0; t < steps; ++t) {
    float u[8], a[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    float u0[4];
    rng4(P, gid, P.global_step + (unsigned)t, PIH_STREAM_ACTION, u0);
    u[0] = u0[0]; u[1] = u0[1]; u[2] = u0[2]; u[3] = u0[3];
    if (A > 4) { rng4(P, gid, P.global_step + (unsigned)t, PIH_STREAM_ACTION + 1, u0); u[4] = u0[0]; u[5] = u0[1]; u[6] = u0[2]; u[7] = u0[3]; }
#pragma unroll
    for (int k = 0; k < A; ++k) a[k] = fmaf(u[k], 2.0f, -1.0f);
    if (!P.action_frame_env) { a[0] += off.x; a[1] += off.y; a[2] += off.z; }
    float o2[5], r2;
    unsigned char d2;
    if (step_env<ARM, TASK>(P, S, a, off, gid, o2, r2, d2)) {
      any = true;
#pragma unroll
      for (int k = 0; k < 5; ++k) o[k] = o2[k];
      r = r2; d = d2;
      nsteps += 1; ndone += d2;
    }
  }
  store_state(P, e, S);
  if (any) {
    if (obs) {
#pragma unroll
      for (int k = 0; k < 5; ++k) obs[5 * (size_t)e + k] = o[k];
    }
    if (reward) reward[e] = r;
    if (done) done[e] = d;
  }
  if (counters) {
    // warp-aggregated counters: one atomic pair per warp (reduction over the lanes that are still active)
    const unsigned m = __activemask();
    unsigned ws = __reduce_add_sync(m, nsteps), wd = __reduce_add_sync(m, ndone);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) { atomicAdd(&counters[0], (unsigned long long)ws); atomicAdd(&counters[1], (unsigned long long)wd); }
  }
}


// ---- warp-specialised kernel for small batches (4096 envs are 128 groups of 32: the step is as long as ONE warp's
// instruction stream, and an SM sub-partition issues an FP32 instruction every second cycle - DESIGN.md 5c).  A CTA owns
// 32 envs and spreads one step's independent recurrences over NW warps, one per sub-partition:
//   warp 0  apply_action: FK -> vel_constraint -> 20 branch-free IK iterations; targets into the mailbox `mb`
//   warp 1  dynamics: FK, ABA (factorisation published in X.ab), its share of the M^-1 columns, motor / limit rows, and
//           after the hand-over of the targets (barrier 0) the solver loop, integration, get_info, stores
//   warp 2  (NW == 3) contacts: FK, narrow phase, rows and solve of the free-body island, the other M^-1 columns
// Same arithmetic as k_step / k_rollout, bit for bit.  Every lane executes the whole code (frozen / padding lanes
// compute on a copy and never commit) so that the warps reach their barriers converged.  pih_rollout on these batches is
// K launches of this kernel with action == NULL (actions drawn in-kernel): a kernel that loops over the steps is
// compiled worse (profiles/r02_summary.md 2d).
// Programmatic dependent launch.  The small-batch kernels are launched with programmatic stream serialisation: once
// every CTA of step N has passed its solver loop (pdl_trigger, in pih_step_body.inc) the CTAs of step N+1 may be
// scheduled; they wait here until step N has completed and flushed its memory.  This overlaps the launch latency of
// the back-to-back ~40 us kernels of a rollout with the tail of the previous step.  With any other predecessor in the
// stream the wait returns only after that predecessor has finished (normal stream order).
PIH_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int ARM, int TASK> PIH_D void mailbox_put(float (*mb)[32], int lane, const float* qs, const float* kpc, const float* mfc, bool fsm_done,
                                                     const EnvState& S) {
  constexpr int nb = ArmDims<ARM>::NB;
#pragma unroll
  for (int j = 0; j < nb; ++j) { mb[PIH_MB_QSTAR + j][lane] = qs[j]; mb[PIH_MB_KP + j][lane] = kpc[j]; mb[PIH_MB_MAXF + j][lane] = mfc[j]; }
  mb[PIH_MB_FSM_DONE][lane] = fsm_done ? 1.0f : 0.0f;
  mb[PIH_MB_FSM_STATE][lane] = __int_as_float(S.fsm_state); mb[PIH_MB_FSM_T][lane] = __int_as_float(S.fsm_t);
#pragma unroll
  for (int k = 0; k < 7; ++k) mb[PIH_MB_MTGT + k][lane] = S.mtgt[k];
  mb[PIH_MB_GRASP_ANG][lane] = S.grasp_ang;
}

// fused "rows delivered" signal (nullable flag) and, for the learner's rank, the consumer-side wait folded into the
// next launch: before touching anything the CTAs wait until `wait_count` flags have reached `wait_value` (nullable)
struct StepSignal {
  unsigned* flag; unsigned value; unsigned* cta_counter;
  const unsigned* wait_flags; int wait_count; unsigned wait_value; unsigned long long timeout_ns; unsigned* timed_out;
  // cta_counter == NULL with a flag: deferred publication (pih_step_publish_prev) - this launch announces `value` at its
  // START (everything the stream did before it, i.e. the previous step's rows, has been delivered) and nothing at its end
};
// bounded spin of one thread on one flag word (steps count up, so >=); records the index of a flag that never came
// acquire load at system scope: what the flag's writer (a kernel on another GPU) released is visible after it - without a
// full system fence on the waiting side
PIH_D unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PIH_D void wait_flag(const unsigned* flag, unsigned value, unsigned long long timeout_ns, unsigned* timed_out, int idx) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(ld_acquire_sys(flag) - value) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) { atomicExch(timed_out, 1u + (unsigned)idx); break; }
    __nanosleep(100);
  }
}

template <int ARM, int TASK, int NW> __global__ void __launch_bounds__(NW * 32, 1) k_step_ws(SimParams P, const float* __restrict__ action, float* __restrict__ obs,
                                                                          float* __restrict__ reward, unsigned char* __restrict__ done,
                                                                          StepSignal sig, unsigned long long* __restrict__ counters) {
  __shared__ float mb[PIH_MB_WORDS][32];
  __shared__ WsXch<ARM> X;
  constexpr int A = ArmDims<ARM>::ADIM, nb = ArmDims<ARM>::NB;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int ec = e < P.n ? e : P.n - 1;
  pdl_wait();
  if (sig.flag && !sig.cta_counter && blockIdx.x == 0 && threadIdx.x == (NW - 1) * 32) {
    // The previous launch of this stream (the step that wrote the rows being announced) has completed - that is what
    // pdl_wait / stream order guarantee - so one system-scope fence orders all of its stores before the flag
    // (cumulativity), and no CTA of any step pays for a fence on its way out.  Lane 0 of the warp with the most slack.
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(sig.flag) = sig.value;
  }
  if (sig.wait_flags) {
    // learner rank: the rows of the previous step of every rank must have landed before this step's rows may follow
    // (the consumer's wait, folded into its next launch instead of a kernel of its own)
    if ((int)threadIdx.x < sig.wait_count) wait_flag(sig.wait_flags + threadIdx.x, sig.wait_value, sig.timeout_ns, sig.timed_out, threadIdx.x);
    __syncthreads();
  }
  // the action row is requested first: with zero-copy host buffers it comes over PCIe, and the state load and the
  // forward kinematics below hide that latency
  float a[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (w == 0 && action) load_action<A>(action, ec, a);
  EnvState S;
  load_state(P, ec, S);
  const bool active = (e < P.n) && !(S.flags & PIH_FLAG_DONE);
  const v3 off = env_offset(P, ec);
  if (w == 0 && !action) {
    // rollout step (pih_rollout): action_space.sample() in-kernel, Philox keyed (seed; global env id, global step)
    const unsigned gid = (unsigned)(P.env_id_base + ec);
    float u[8], u0[4];
    rng4(P, gid, P.global_step, PIH_STREAM_ACTION, u0);
    u[0] = u0[0]; u[1] = u0[1]; u[2] = u0[2]; u[3] = u0[3];
    if (A > 4) { rng4(P, gid, P.global_step, PIH_STREAM_ACTION + 1, u0); u[4] = u0[0]; u[5] = u0[1]; u[6] = u0[2]; u[7] = u0[3]; }
#pragma unroll
    for (int k = 0; k < A; ++k) a[k] = fmaf(u[k], 2.0f, -1.0f);
    if (!P.action_frame_env) { a[0] += off.x; a[1] += off.y; a[2] += off.z; }
  }
  PIH_TICK(w == 0 ? 0 : (w == 1 ? 2 : 9));
  if (w == 1) PIH_TICK_CTA(0);
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qs[nb], kpc[nb], mfc[nb];
  if (w == 0) {
    int ik_n;
    const bool fd = apply_action<ARM, TASK>(P, S, F, a, off, qs, kpc, mfc, ik_n, active ? e : -1);
    mailbox_put<ARM, TASK>(mb, lane, qs, kpc, mfc, fd, S);
    stat_add(P, PIH_STAT_IK_ITERS, active ? (unsigned)ik_n : 0u);
    PIH_TICK(1);
    __syncthreads();
    return;
  }
  if (NW == 3 && w == 2) {
    contacts_warp<ARM, TASK>(P, S, F, X, lane);
    __syncthreads();
    return;
  }
  float o[5], r;
  unsigned char d;
  step_back_ws<ARM, TASK, NW>(P, S, F, qs, kpc, mfc, false, mb, X, lane, active, off, (unsigned)(P.env_id_base + ec), o, r, d);
  PIH_TICK(7);
  PIH_TICK_CTA(1);
  if (active) store_state(P, e, S);
  else frozen_outputs<ARM>(S, off, o, r, d);   // done env (or a padding lane, which writes nothing below)
  if (counters) {   // rollouts: env-steps executed and episodes finished (one atomic pair per warp)
    const unsigned ws = __reduce_add_sync(0xffffffffu, active ? 1u : 0u), wd = __reduce_add_sync(0xffffffffu, (active && d) ? 1u : 0u);
    if (lane == 0 && ws) { atomicAdd(&counters[0], (unsigned long long)ws); atomicAdd(&counters[1], (unsigned long long)wd); }
  }
  if (obs) {
    // the CTA's 32 observation rows are 640 contiguous bytes: transpose through the (now idle) mailbox so that the
    // warp writes five full 128-byte lines instead of 160 scattered words (matters when obs is mapped host memory)
    float* so = &mb[0][0];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 5; ++k) so[5 * lane + k] = o[k];
    __syncwarp();
    const int rows = min(32, P.n - (int)blockIdx.x * 32);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int idx = 32 * k + lane;
      if (idx < 5 * rows) obs[(size_t)blockIdx.x * 160 + idx] = so[idx];
    }
  }
  if (e < P.n) {
    if (reward) reward[e] = r;
    if (done) done[e] = d;
  }
  if (sig.flag && sig.cta_counter) {
    // Fused gather epilogue: the output pointers may be peer mappings of a learner buffer on another
    // GPU.  Every CTA makes its rows visible system-wide and counts itself; the last one publishes the step number in
    // this rank's flag word (and re-arms the counter), so no separate signalling launch follows the step.
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
      if (atomicAdd(sig.cta_counter, 1u) == gridDim.x - 1) {
        *sig.cta_counter = 0u;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(sig.flag) = sig.value;
      }
    }
  }
  PIH_TICK(8);
}

template <int ARM, int TASK> __global__ void __launch_bounds__(128) k_ik(SimParams P, const float* __restrict__ action, float* __restrict__ qtarget) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int A = ArmDims<ARM>::ADIM, nb = ArmDims<ARM>::NB;
  EnvState S;
  load_state(P, e, S);
  float a[6];
  load_action<A>(action, e, a);
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  float qs[nb], kpc[nb], mfc[nb];
  int ik_n;
  apply_action<ARM, TASK>(P, S, F, a, env_offset(P, e), qs, kpc, mfc, ik_n);
#pragma unroll
  for (int k = 0; k < PIH_NB_MAX; ++k) qtarget[(size_t)PIH_NB_MAX * e + k] = (k < nb) ? qs[k < nb ? k : 0] : 0.0f;
}

template <int ARM> __global__ void __launch_bounds__(128) k_ee(SimParams P, float* __restrict__ pos, float* __restrict__ rot) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.n) return;
  constexpr int nb = ArmDims<ARM>::NB;
  EnvState S;
  load_state(P, e, S);
  FK<ARM> F;
  fk<ARM, nb>(S.q, F);
  v3 pee; m3 Ree;
  ee_pose<ARM>(F, pee, Ree);
  v3 off = env_offset(P, e);
  if (pos) { pos[3 * (size_t)e] = pee.x + off.x; pos[3 * (size_t)e + 1] = pee.y + off.y; pos[3 * (size_t)e + 2] = pee.z + off.z; }
  if (rot) {
#pragma unroll
    for (int i = 0; i < 9; ++i) rot[9 * (size_t)e + i] = Ree.m[i / 3][i % 3];
  }
}

// ---- gather-to-learner without a collective: every rank's step kernel stores its (obs, reward, done) rows straight
// into rank 0's buffer through a peer mapping (NVLink); a one-thread kernel then publishes the step number in that rank's
// flag, and rank 0 waits on the flags in its own stream.  The stores of the step kernel are complete before the
// signal kernel starts (stream order); the fences order the flag against them across devices.
__global__ void k_signal_flag(volatile unsigned* flag, unsigned value) {
  __threadfence_system();
  *flag = value;
  __threadfence_system();
}
// waits until every flag has reached `value` (steps count up, so >=); gives up after `timeout_ns` and records it
__global__ void k_wait_flags(const unsigned* flags, int count, unsigned value, unsigned long long timeout_ns, unsigned* timed_out) {
  const int i = threadIdx.x;
  if (i < count) wait_flag(flags + i, value, timeout_ns, timed_out, i);
  __threadfence_system();
}

__global__ void k_init_state(float4* state, int n, int slots) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int s = 0; s < slots; ++s) state[(size_t)s * n + e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (slots > (PIH_PS_QUAT + 3) / 4)   // flexible pipe: identity base orientation
    reinterpret_cast<float*>(&state[(size_t)((PIH_PS_QUAT + 3) >> 2) * n + e])[(PIH_PS_QUAT + 3) & 3] = 1.0f;
  // episode = -1 (no reset yet), identity quaternion
  reinterpret_cast<float*>(&state[(size_t)(PIH_S_EPISODE >> 2) * n + e])[PIH_S_EPISODE & 3] = __int_as_float(-1);
  reinterpret_cast<float*>(&state[(size_t)((PIH_S_PQUAT + 3) >> 2) * n + e])[(PIH_S_PQUAT + 3) & 3] = 1.0f;
}

// FP32 FMA throughput probe (roofline denominator): 8 independent FFMA chains per thread
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;
}

}  // namespace pih

// ------------------------------------------------------------------------------------------------ host side
using namespace pih;

#define PIH_ALIAS_SLOTS 16
struct pih_handle {
  pih_config cfg;
  SimParams P;
  int was_reset;
  int block;
  int ws;  // warp-specialised kernels for small batches
  char err[256];
  // staging for pih_step_host
  float *h_action, *h_obs, *h_reward;
  unsigned char* h_done;
  float *d_action, *d_obs, *d_reward;
  unsigned char* d_done;
  float* d_scene;  // render pre-pass output [n][PIH_SCENE_WORDS], allocated on first pih_render
  float* d_qt;     // joint targets [nb][n] between k_apply_action and k_step_dyn (large batches), allocated on first use
  int split;       // large batches: step = k_apply_action + k_step_dyn (0: fused k_step); pih_config.kernel_family overrides
  // host buffers bound with pih_host_bind (the caller guarantees their lifetime until pih_host_unbind): base, size, device alias
  const char* bind_host[PIH_ALIAS_SLOTS];
  size_t bind_bytes[PIH_ALIAS_SLOTS];
  char* bind_dev[PIH_ALIAS_SLOTS];
  int n_bind, no_zero_copy;
  float* d_stage;            // large batches with a learner buffer: local [obs n*5 | reward n | done n] block the rows are copied from
  int pipe;                  // peg_model == PIH_PEG_PIPE: the flexible-pipe kernels (pih_pipe.cu)
  float* d_pipe_scratch;     // per-env block between k_pipe_arm and k_pipe_step
  unsigned* d_cta_counter;   // fused step signal: CTAs that have delivered their rows (re-armed by the last one)
};

static int g_tables_device = -1;

static int fail(pih_handle* h, int code, const char* what, cudaError_t ce) {
  if (h) snprintf(h->err, sizeof(h->err), "%s%s%s", what, ce != cudaSuccess ? ": " : "", ce != cudaSuccess ? cudaGetErrorString(ce) : "");
  return code;
}

// Makes the handle's device current for the duration of an entry point (launches must go to the device that owns the
// state) and restores the caller's current device on return: a library call must not change the caller's device.
struct DeviceGuard {
  int prev, dev;
  bool ok;
  explicit DeviceGuard(int device) : prev(-1), dev(device), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define USE_DEVICE(h)                                                                  \
  DeviceGuard guard_((h)->cfg.device);                                                 \
  if (!guard_.ok) return fail((h), PIH_ERR_CUDA, "cudaSetDevice", cudaGetLastError())

#include "pih_tables_host.h"


// launch of the small-batch kernels with programmatic stream serialisation (see pdl_wait)
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

extern "C" {

int pih_abi_version(void) { return PIH_ABI_VERSION; }
int pih_state_words(void) { return PIH_STATE_WORDS; }
static int state_slots(const pih_config* cfg) {
  if (cfg && cfg->peg_model == PIH_PEG_PIPE) return PIH_PIPE_STATE_F4;
  if (cfg && cfg->task == PIH_TASK_SCRIPTED) return PIH_SCRIPTED_STATE_F4;
  return PIH_STATE_F4;
}
int pih_state_words_for(const pih_config* cfg) { return 4 * state_slots(cfg); }
int pih_action_dim(int arm) { return arm == PIH_ARM_PANDA ? 4 : (arm == PIH_ARM_UR5 ? 6 : PIH_ERR_INVALID); }
int pih_obs_dim(void) { return 5; }

int pih_env_offset(const pih_config* cfg, int32_t local_env, float out[3]) {
  if (!cfg || !out) return PIH_ERR_INVALID;
  int sub = cfg->sub_num > 0 ? cfg->sub_num : cfg->num_envs;
  if (sub <= 0) return PIH_ERR_INVALID;
  int e = (cfg->env_id_base + local_env) % sub;
  if (cfg->offset[0] == 0.0f || cfg->offset[1] == 0.0f) {
    out[0] = cfg->offset[0] * (float)e; out[1] = cfg->offset[1] * (float)e; out[2] = cfg->offset[2] * (float)e;
  } else {
    int s = 0;
    while (s * s < sub) ++s;
    out[0] = cfg->offset[0] * (float)(e / s); out[1] = cfg->offset[1] * (float)(e % s); out[2] = cfg->offset[2];
  }
  return PIH_OK;
}

int pih_create(const pih_config* cfg, void* state_dev, pih_handle** out) {
  if (!cfg || !out || !state_dev) return PIH_ERR_INVALID;
  *out = nullptr;
  if (cfg->num_envs <= 0 || cfg->sub_num <= 0) return PIH_ERR_INVALID;
  if (cfg->task != PIH_TASK_PEG_IN_HOLE && cfg->task != PIH_TASK_RANDOM_FLY && cfg->task != PIH_TASK_SCRIPTED) return PIH_ERR_INVALID;
  if (cfg->arm != PIH_ARM_PANDA && cfg->arm != PIH_ARM_UR5) return PIH_ERR_INVALID;
  // the ur5 variant exists for the peg-in-hole task (utils.py:40-48,70-82); random-fly and the scripted episode are
  // panda tasks (README.md:38, peg_in_hole.py:233), and their ur5 kernels are not instantiated
  if (cfg->arm == PIH_ARM_UR5 && cfg->task != PIH_TASK_PEG_IN_HOLE) return PIH_ERR_INVALID;
  if (cfg->peg_model != PIH_PEG_RIGID && cfg->peg_model != PIH_PEG_PIPE) return PIH_ERR_INVALID;
  // the pipe is the panda tasks' object (peg_in_hole.py:233-245): 'peg-in-hole' and the scripted episode
  if (cfg->peg_model == PIH_PEG_PIPE && (cfg->arm != PIH_ARM_PANDA || cfg->task == PIH_TASK_RANDOM_FLY)) return PIH_ERR_INVALID;
  if (cfg->kernel_family < 0 || cfg->kernel_family > 4 || cfg->host_path < 0 || cfg->host_path > 1) return PIH_ERR_INVALID;
  if (!(cfg->dt > 0.0f) || cfg->solver_iters < 0 || cfg->ik_iters < 0 || !(cfg->solver_residual >= 0.0f)) return PIH_ERR_INVALID;
  // reference assert: task_num == 1 or offset != [0,0,0] (envs/base_env.py:17)
  if (cfg->sub_num > 1 && cfg->offset[0] == 0.0f && cfg->offset[1] == 0.0f && cfg->offset[2] == 0.0f) return PIH_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return PIH_ERR_CUDA;
  DeviceGuard guard_(cfg->device);
  if (!guard_.ok) return PIH_ERR_CUDA;
  pih_handle* h = new (std::nothrow) pih_handle();
  if (!h) return PIH_ERR_INVALID;
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  if (g_tables_device != cfg->device) {
    cudaError_t ce = upload_tables();
    if (ce != cudaSuccess) { delete h; return PIH_ERR_CUDA; }
    g_tables_device = cfg->device;
  }
  SimParams& P = h->P;
  P.n = cfg->num_envs; P.task = cfg->task; P.arm = cfg->arm;
  P.sub_num = cfg->sub_num;
  P.line_layout = (cfg->offset[0] == 0.0f || cfg->offset[1] == 0.0f) ? 1 : 0;
  int s = 0;
  while (s * s < cfg->sub_num) ++s;
  P.grid_side = s > 0 ? s : 1;
  P.env_id_base = cfg->env_id_base;
  for (int k = 0; k < 3; ++k) P.offset[k] = cfg->offset[k];
  P.dt = cfg->dt; P.idt = 1.0f / cfg->dt; P.solver_residual = cfg->solver_residual;
  P.solver_iters = cfg->solver_iters; P.ik_iters = cfg->ik_iters; P.max_steps = cfg->max_steps;
  P.auto_reset = cfg->auto_reset; P.action_frame_env = cfg->action_frame_env;
  P.seed_lo = (unsigned)cfg->seed; P.seed_hi = (unsigned)(cfg->seed >> 32);
  P.global_step = 0;
  double hx, hy, hz, mass, mu;
  if (cfg->task == PIH_TASK_RANDOM_FLY) {
    P.fb_shape = 1;
    if (cfg->fly_object == 0) { hx = PIH_BANANA_HX; hy = PIH_BANANA_HY; hz = PIH_BANANA_HZ; }
    else { hx = PIH_AMICELLI_HX; hy = PIH_AMICELLI_HY; hz = PIH_AMICELLI_HZ; }
    mass = PIH_FLY_MASS; mu = PIH_FLY_MU;
  } else {
    P.fb_shape = 0; hx = PIH_PEG_R; hy = PIH_PEG_HL; hz = PIH_PEG_R; mass = PIH_PEG_MASS; mu = PIH_PEG_MU;
  }
  double lx = 2 * (hx + 0.001), ly = 2 * (hy + 0.001), lz = 2 * (hz + 0.001);
  double ix = mass / 12.0 * (ly * ly + lz * lz), iy = mass / 12.0 * (lx * lx + lz * lz), iz = mass / 12.0 * (lx * lx + ly * ly);
  P.fb_mass = (float)mass; P.fb_im = 1.0f / P.fb_mass; P.fb_mu = (float)mu;
  P.fb_half[0] = (float)hx; P.fb_half[1] = (float)hy; P.fb_half[2] = (float)hz;
  P.fb_idiag[0] = (float)ix; P.fb_idiag[1] = (float)iy; P.fb_idiag[2] = (float)iz;
  P.fb_iinv[0] = (float)(1.0 / ix); P.fb_iinv[1] = (float)(1.0 / iy); P.fb_iinv[2] = (float)(1.0 / iz);
  P.state = reinterpret_cast<float4*>(state_dev);
  h->block = 32;
  // warp-specialised kernels while a 64-thread CTA per 32 envs fits the register file (4 CTAs x 64 x 255 registers per
  // SM): measured faster than the classic kernels up to 16384 envs for both arms (round 1, profiles/r01_summary.md), slower
  // from 24576 envs on, where the classic kernels' one warp per 32 envs fills the schedulers first
  // (h->ws: 0 classic kernels, 2 action warp + dynamics warp, 3 = + contacts warp: 96 threads x 255 registers, two
  // resident CTAs per SM, so only while the batch fits one wave of those)
  h->ws = (P.n <= 148 * 32 * 2) ? 3 : ((P.n <= 148 * 32 * 4) ? 2 : 0);
  // measured (round 1, profiles/r01_summary.md): fused 216 / 434 / 781 us against split 196 / 378 / 692 us per step at 65536 /
  // 131072 / 262144 envs; below ~2 waves of the fused kernel the extra launch does not pay (104 vs 106 us at 32768)
  h->split = (P.n >= 49152) ? 1 : 0;
  if (cfg->kernel_family == PIH_KERNELS_WARP_SPECIALISED) { h->ws = (P.n <= 148 * 32 * 2) ? 3 : 2; h->split = 0; }
  else if (cfg->kernel_family == PIH_KERNELS_WARP_SPECIALISED_2) { h->ws = 2; h->split = 0; }
  else if (cfg->kernel_family == PIH_KERNELS_FUSED) { h->ws = 0; h->split = 0; }
  else if (cfg->kernel_family == PIH_KERNELS_SPLIT) { h->ws = 0; h->split = 1; }
  h->no_zero_copy = cfg->host_path == 1;
  h->pipe = cfg->peg_model == PIH_PEG_PIPE;
  k_init_state<<<(P.n + 127) / 128, 128>>>(P.state, P.n, state_slots(cfg));
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess && h->pipe) ce = pih_pipe_upload_tables();
  if (ce == cudaSuccess && h->pipe) ce = cudaMalloc(&h->d_pipe_scratch, pih_pipe_scratch_bytes(P.n));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_cta_counter, sizeof(unsigned));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_cta_counter, 0, sizeof(unsigned));
  if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { delete h; return PIH_ERR_CUDA; }
  *out = h;
  return PIH_OK;
}

void pih_destroy(pih_handle* h) {
  if (!h) return;
  if (h->d_cta_counter) cudaFree(h->d_cta_counter);
  if (h->d_pipe_scratch) cudaFree(h->d_pipe_scratch);
  if (h->d_stage) cudaFree(h->d_stage);
  if (h->d_qt) cudaFree(h->d_qt);
  if (h->h_action) cudaFreeHost(h->h_action);
  if (h->h_obs) cudaFreeHost(h->h_obs);     // one pinned block: obs | reward | done
  if (h->d_action) cudaFree(h->d_action);
  if (h->d_obs) cudaFree(h->d_obs);         // one device block: obs | reward | done
  if (h->d_scene) cudaFree(h->d_scene);
  delete h;
}

const char* pih_last_error(const pih_handle* h) { return h ? h->err : "null handle"; }

#ifdef PIH_FAST_BUILD   // experiment builds (tools/): only ONE <arm, task> instantiation (default panda, peg-in-hole): a 6x shorter compile
#ifndef PIH_FAST_ARM
#define PIH_FAST_ARM PIH_ARM_PANDA
#endif
#ifndef PIH_FAST_TASK
#define PIH_FAST_TASK PIH_TASK_PEG_IN_HOLE
#endif
#define LAUNCH_T(kern, ARMV, h, st, ...) kern<PIH_FAST_ARM, PIH_FAST_TASK><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__)
#define LAUNCH(kern, h, st, ...)                                                  \
  do {                                                                            \
    int grid_ = ((h)->P.n + (h)->block - 1) / (h)->block;                         \
    LAUNCH_T(kern, PIH_ARM_PANDA, h, st, __VA_ARGS__);                            \
  } while (0)
#define LAUNCH_WS(kern, h, st, ...)                                               \
  do {                                                                            \
    int gridw_ = ((h)->P.n + 31) / 32;                                           \
    if ((h)->ws == 3) launch_pdl(kern<PIH_FAST_ARM, PIH_FAST_TASK, 3>, gridw_, 96, (cudaStream_t)(st), __VA_ARGS__); \
    else launch_pdl(kern<PIH_FAST_ARM, PIH_FAST_TASK, 2>, gridw_, 64, (cudaStream_t)(st), __VA_ARGS__); \
  } while (0)
#else
#define LAUNCH_T(kern, ARMV, h, st, ...)                                                                                \
  do {                                                                                                                 \
    if ((h)->P.task == PIH_TASK_PEG_IN_HOLE) kern<ARMV, PIH_TASK_PEG_IN_HOLE><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
    else if ((h)->P.task == PIH_TASK_RANDOM_FLY) kern<ARMV, PIH_TASK_RANDOM_FLY><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
    else kern<ARMV, PIH_TASK_SCRIPTED><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__);                     \
  } while (0)
#define LAUNCH(kern, h, st, ...)                                                  \
  do {                                                                            \
    int grid_ = ((h)->P.n + (h)->block - 1) / (h)->block;                         \
    if ((h)->P.arm == PIH_ARM_PANDA) LAUNCH_T(kern, PIH_ARM_PANDA, h, st, __VA_ARGS__); \
    else kern<PIH_ARM_UR5, PIH_TASK_PEG_IN_HOLE><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); /* ur5: peg-in-hole only (pih_create) */ \
  } while (0)
#define LAUNCH_WS_N(kern, ARMV, TASKV, h, st, ...)                                                                       \
  do {                                                                                                                 \
    if ((h)->ws == 3) launch_pdl(kern<ARMV, TASKV, 3>, gridw_, 96, (cudaStream_t)(st), __VA_ARGS__);                       \
    else launch_pdl(kern<ARMV, TASKV, 2>, gridw_, 64, (cudaStream_t)(st), __VA_ARGS__);                                    \
  } while (0)
#define LAUNCH_WS_T(kern, ARMV, h, st, ...)                                                                             \
  do {                                                                                                                 \
    if ((h)->P.task == PIH_TASK_PEG_IN_HOLE) LAUNCH_WS_N(kern, ARMV, PIH_TASK_PEG_IN_HOLE, h, st, __VA_ARGS__);           \
    else if ((h)->P.task == PIH_TASK_RANDOM_FLY) LAUNCH_WS_N(kern, ARMV, PIH_TASK_RANDOM_FLY, h, st, __VA_ARGS__);        \
    else LAUNCH_WS_N(kern, ARMV, PIH_TASK_SCRIPTED, h, st, __VA_ARGS__);                                                  \
  } while (0)
#define LAUNCH_WS(kern, h, st, ...)                                               \
  do {                                                                            \
    int gridw_ = ((h)->P.n + 31) / 32;                                           \
    if ((h)->P.arm == PIH_ARM_PANDA) LAUNCH_WS_T(kern, PIH_ARM_PANDA, h, st, __VA_ARGS__); \
    else LAUNCH_WS_N(kern, PIH_ARM_UR5, PIH_TASK_PEG_IN_HOLE, h, st, __VA_ARGS__); \
  } while (0)
#endif
// peg-in-hole / random-fly only (kernels that are never instantiated for the scripted task)
#ifdef PIH_FAST_BUILD
#define LAUNCH_PF(kern, h, st, ...) LAUNCH(kern, h, st, __VA_ARGS__)
#else
#define LAUNCH_PF(kern, h, st, ...)                                               \
  do {                                                                            \
    int grid_ = ((h)->P.n + (h)->block - 1) / (h)->block;                         \
    if ((h)->P.arm == PIH_ARM_UR5) kern<PIH_ARM_UR5, PIH_TASK_PEG_IN_HOLE><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
    else if ((h)->P.task == PIH_TASK_RANDOM_FLY) kern<PIH_ARM_PANDA, PIH_TASK_RANDOM_FLY><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
    else kern<PIH_ARM_PANDA, PIH_TASK_PEG_IN_HOLE><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
  } while (0)
#endif
#define LAUNCH_ARM(kern, h, st, ...)                                              \
  do {                                                                            \
    int grid_ = ((h)->P.n + (h)->block - 1) / (h)->block;                         \
    if ((h)->P.arm == PIH_ARM_PANDA) kern<PIH_ARM_PANDA><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
    else kern<PIH_ARM_UR5><<<grid_, (h)->block, 0, (cudaStream_t)(st)>>>(__VA_ARGS__); \
  } while (0)

int pih_reset(pih_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!h) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  if (h->pipe) pih_pipe_launch_reset(h->P, mask_dev, obs_dev, (cudaStream_t)stream);
  else LAUNCH(k_reset, h, stream, h->P, mask_dev, obs_dev);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_reset launch", ce);
  h->was_reset = 1;
  return PIH_OK;
}

static int step_impl(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, StepSignal sig, void* stream);

int pih_step(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream) {
  StepSignal sig = {};
  return step_impl(h, action_dev, obs_dev, reward_dev, done_dev, sig, stream);
}

int pih_step_signal(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                    uint32_t flag_value, void* stream) {
  if (!flag_dev) return PIH_ERR_INVALID;
  StepSignal sig = {};
  sig.flag = flag_dev; sig.value = flag_value;
  return step_impl(h, action_dev, obs_dev, reward_dev, done_dev, sig, stream);
}

int pih_step_signal_wait(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                         uint32_t flag_value, const uint32_t* wait_flags_dev, int32_t wait_count, uint32_t wait_value, uint32_t timeout_ms,
                         uint32_t* timed_out_dev, void* stream) {
  if (!flag_dev || !wait_flags_dev || !timed_out_dev || wait_count <= 0 || wait_count > 64) return PIH_ERR_INVALID;
  StepSignal sig = {};
  sig.flag = flag_dev; sig.value = flag_value;
  sig.wait_flags = wait_flags_dev; sig.wait_count = wait_count; sig.wait_value = wait_value;
  sig.timeout_ns = (unsigned long long)timeout_ms * 1000000ull; sig.timed_out = timed_out_dev;
  return step_impl(h, action_dev, obs_dev, reward_dev, done_dev, sig, stream);
}

int pih_step_publish_prev(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                          uint32_t prev_value, const uint32_t* wait_flags_dev, int32_t wait_count, uint32_t wait_value, uint32_t timeout_ms,
                          uint32_t* timed_out_dev, void* stream) {
  if (!flag_dev || (wait_flags_dev && (!timed_out_dev || wait_count <= 0 || wait_count > 64))) return PIH_ERR_INVALID;
  StepSignal sig = {};
  sig.flag = flag_dev; sig.value = prev_value; sig.wait_count = -1;   // (-1: deferred; step_impl keeps cta_counter NULL)
  if (wait_flags_dev) {
    sig.wait_flags = wait_flags_dev; sig.wait_count = -wait_count - 1; sig.wait_value = wait_value;
    sig.timeout_ns = (unsigned long long)timeout_ms * 1000000ull; sig.timed_out = timed_out_dev;
  }
  return step_impl(h, action_dev, obs_dev, reward_dev, done_dev, sig, stream);
}

static int step_impl(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, StepSignal sig, void* stream) {
  if (!h || !action_dev) return PIH_ERR_INVALID;
  if (!h->was_reset) return fail(h, PIH_ERR_STATE, "step() before reset()", cudaSuccess);
  if ((uintptr_t)action_dev & (h->P.arm == PIH_ARM_PANDA ? 15 : 7)) return fail(h, PIH_ERR_INVALID, "action rows must be 16-byte (Box(4)) / 8-byte (Box(6)) aligned", cudaSuccess);
  USE_DEVICE(h);
  const bool deferred = sig.wait_count < 0;        // pih_step_publish_prev: wait_count arrives as -(count) - 1
  if (deferred) sig.wait_count = -sig.wait_count - 1;
  sig.cta_counter = deferred ? nullptr : h->d_cta_counter;
  uint32_t* const pub_flag = deferred ? sig.flag : nullptr;      // announced at the START of this call
  uint32_t* const flag_dev = deferred ? nullptr : sig.flag;      // announced at its end
  const uint32_t flag_value = sig.value;
  // signalled step of a large batch: the kernels write a local block, one copy kernel delivers it and publishes the flag
  float* const out_obs = obs_dev; float* const out_rew = reward_dev; uint8_t* const out_done = done_dev;
  const bool staged_rows = (flag_dev || pub_flag) && !h->ws && obs_dev && reward_dev && done_dev;
  // deferred publication outside the small-batch kernels: a one-thread launch in front of the step
  if (pub_flag && (h->pipe || !h->ws)) k_signal_flag<<<1, 1, 0, (cudaStream_t)stream>>>(pub_flag, flag_value);
  if (staged_rows) {
    const size_t n = (size_t)h->P.n;
    if (!h->d_stage && cudaMalloc(&h->d_stage, n * 24 + ((n + 15) & ~(size_t)15)) != cudaSuccess)
      return fail(h, PIH_ERR_CUDA, "row stage allocation", cudaGetLastError());
    obs_dev = h->d_stage; reward_dev = h->d_stage + 5 * n; done_dev = reinterpret_cast<uint8_t*>(h->d_stage + 6 * n);
  }
  if (h->pipe) {
    if (sig.wait_flags) k_wait_flags<<<1, 64, 0, (cudaStream_t)stream>>>(sig.wait_flags, sig.wait_count, sig.wait_value, sig.timeout_ns, sig.timed_out);
    pih_pipe_launch_step(h->P, action_dev, h->d_pipe_scratch, obs_dev, reward_dev, done_dev, nullptr, (cudaStream_t)stream);
    if (flag_dev && !staged_rows) k_signal_flag<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, flag_value);
  } else if (h->ws) { LAUNCH_WS(k_step_ws, h, stream, h->P, action_dev, obs_dev, reward_dev, done_dev, sig, (unsigned long long*)nullptr); }
  else {
    // large batches: the wait and the signal are launches of their own (noise next to a multi-hundred-microsecond step)
    if (sig.wait_flags) k_wait_flags<<<1, 64, 0, (cudaStream_t)stream>>>(sig.wait_flags, sig.wait_count, sig.wait_value, sig.timeout_ns, sig.timed_out);
    if (h->split && h->P.task != PIH_TASK_SCRIPTED) {
      if (!h->d_qt && cudaMalloc(&h->d_qt, (size_t)h->P.n * PIH_NB_MAX * sizeof(float)) != cudaSuccess)
        return fail(h, PIH_ERR_CUDA, "target scratch allocation", cudaGetLastError());
      LAUNCH_PF(k_apply_action, h, stream, h->P, action_dev, h->d_qt);
      LAUNCH_PF(k_step_dyn, h, stream, h->P, h->d_qt, obs_dev, reward_dev, done_dev, (unsigned long long*)nullptr);
    } else {
      LAUNCH(k_step, h, stream, h->P, action_dev, obs_dev, reward_dev, done_dev);
    }
    if (flag_dev && !staged_rows) k_signal_flag<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, flag_value);
  }
  if (staged_rows) {
    // Large batches: the (obs, reward, done) rows reach the learner buffer as a device-to-device copy after the step (the
    // copy engine moves the 13 MB of 524288 envs over NVLink in ~40 us; per-thread peer stores inside the step kernel
    // cost 180 us, a copy kernel with word stores 200 us), then a one-thread kernel publishes the flag.  One copy when the
    // destination slices are laid out like the local block (PeerLearnerBuffer's layout), else three.
    const size_t n = (size_t)h->P.n;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t cc;
    if (out_rew == out_obs + 5 * n && out_done == reinterpret_cast<uint8_t*>(out_rew + n)) cc = cudaMemcpyAsync(out_obs, obs_dev, n * 25, cudaMemcpyDeviceToDevice, st);
    else {
      cc = cudaMemcpyAsync(out_obs, obs_dev, n * 20, cudaMemcpyDeviceToDevice, st);
      if (cc == cudaSuccess) cc = cudaMemcpyAsync(out_rew, reward_dev, n * 4, cudaMemcpyDeviceToDevice, st);
      if (cc == cudaSuccess) cc = cudaMemcpyAsync(out_done, done_dev, n, cudaMemcpyDeviceToDevice, st);
    }
    if (cc != cudaSuccess) return fail(h, PIH_ERR_CUDA, "row copy to the learner buffer", cc);
    if (flag_dev) k_signal_flag<<<1, 1, 0, st>>>(flag_dev, flag_value);
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_step launch", ce);
  h->P.global_step += 1;
  return PIH_OK;
}

// staging for the host-buffer call: ONE pinned block and ONE device block holding [obs n*5 f32 | reward n f32 | done n u8],
// so that the results come back in a single D2H copy
static size_t out_bytes(size_t n) { return n * 5 * sizeof(float) + n * sizeof(float) + n; }

static int ensure_staging(pih_handle* h) {
  if (h->h_action) return PIH_OK;
  size_t n = (size_t)h->P.n;
  cudaError_t ce = cudaMallocHost(&h->h_action, n * 6 * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMallocHost(&h->h_obs, out_bytes(n));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_action, n * 6 * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_obs, out_bytes(n));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_obs, 0, out_bytes(n));
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "staging allocation", ce);
  h->h_reward = h->h_obs + n * 5;
  h->h_done = reinterpret_cast<unsigned char*>(h->h_reward + n);
  h->d_reward = h->d_obs + n * 5;
  h->d_done = reinterpret_cast<unsigned char*>(h->d_reward + n);
  return PIH_OK;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Device alias of a page-locked, device-mapped host buffer (NULL when the buffer is pageable or not mapped).
// Buffers the caller has bound with pih_host_bind (lifetime guaranteed by the caller) resolve from the handle's table;
// any other pointer is re-queried on EVERY call - a remembered answer could outlive the allocation it described (a
// freed pinned buffer whose address is reused by a pageable one), and the query costs about a microsecond.
static void* mapped_alias(pih_handle* h, const void* p) {
  if (!p) return nullptr;
  const char* c = static_cast<const char*>(p);
  for (int i = 0; i < h->n_bind; ++i)
    if (c >= h->bind_host[i] && c < h->bind_host[i] + h->bind_bytes[i]) return h->bind_dev[i] + (c - h->bind_host[i]);
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost) return nullptr;
  return a.devicePointer;   // NULL when the registration is not device-mapped
}

int pih_host_bind(pih_handle* h, const void* host_ptr, uint64_t bytes) {
  if (!h || !host_ptr || bytes == 0) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return fail(h, PIH_ERR_INVALID, "pih_host_bind: not a CUDA-visible host buffer", cudaSuccess); }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return fail(h, PIH_ERR_INVALID, "pih_host_bind: buffer is not page-locked and device-mapped", cudaSuccess);
  // the whole range must belong to the same mapping: its last byte has to resolve to the matching alias
  cudaPointerAttributes e;
  const char* last = static_cast<const char*>(host_ptr) + bytes - 1;
  if (cudaPointerGetAttributes(&e, last) != cudaSuccess || e.type != cudaMemoryTypeHost ||
      static_cast<char*>(e.devicePointer) - static_cast<char*>(a.devicePointer) != (ptrdiff_t)(bytes - 1)) {
    cudaGetLastError();
    return fail(h, PIH_ERR_INVALID, "pih_host_bind: range exceeds the page-locked allocation", cudaSuccess);
  }
  for (int i = 0; i < h->n_bind; ++i)
    if (h->bind_host[i] == host_ptr) { h->bind_bytes[i] = (size_t)bytes; h->bind_dev[i] = static_cast<char*>(a.devicePointer); return PIH_OK; }
  if (h->n_bind >= PIH_ALIAS_SLOTS) return fail(h, PIH_ERR_INVALID, "pih_host_bind: table full", cudaSuccess);
  h->bind_host[h->n_bind] = static_cast<const char*>(host_ptr); h->bind_bytes[h->n_bind] = (size_t)bytes;
  h->bind_dev[h->n_bind] = static_cast<char*>(a.devicePointer);
  h->n_bind += 1;
  return PIH_OK;
}

int pih_host_unbind(pih_handle* h, const void* host_ptr) {
  if (!h) return PIH_ERR_INVALID;
  for (int i = 0; i < h->n_bind; ++i)
    if (h->bind_host[i] == host_ptr || !host_ptr) {
      h->n_bind -= 1;
      h->bind_host[i] = h->bind_host[h->n_bind]; h->bind_bytes[i] = h->bind_bytes[h->n_bind]; h->bind_dev[i] = h->bind_dev[h->n_bind];
      if (host_ptr) return PIH_OK;
      --i;
    }
  return host_ptr ? PIH_ERR_INVALID : PIH_OK;
}

int pih_step_host(pih_handle* h, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host, void* stream) {
  if (!h || !action_host) return PIH_ERR_INVALID;
  if (!h->was_reset) return fail(h, PIH_ERR_STATE, "step() before reset()", cudaSuccess);
  USE_DEVICE(h);
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)h->P.n, A = (size_t)pih_action_dim(h->P.arm);
  // Zero-copy: when every caller buffer is page-locked and device-mapped, the step kernel reads the actions from and
  // writes (obs, reward, done) to host memory itself (PCIe reads/writes issued by the SMs: 16 B in and 25 B out per
  // env), so the call is ONE launch plus one stream synchronisation - no staging copies and no copy-engine latency.
  if (!h->no_zero_copy && obs_host && reward_host && done_host) {
    float* da = (float*)mapped_alias(h, action_host);
    float* dob = (float*)mapped_alias(h, obs_host);
    float* dr = (float*)mapped_alias(h, reward_host);
    uint8_t* dd = (uint8_t*)mapped_alias(h, done_host);
    if (da && dob && dr && dd && ((uintptr_t)da & 15) == 0 && ((uintptr_t)dob & 15) == 0) {
      int rc = pih_step(h, da, dob, dr, dd, stream);
      if (rc) return rc;
      cudaError_t ce = cudaStreamSynchronize(st);
      if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "step (zero-copy) sync", ce);
      return PIH_OK;
    }
  }
  int rc = ensure_staging(h);
  if (rc) return rc;
  // page-locked caller buffers are used for the DMA directly; pageable ones go through the pinned staging block
  const float* src = action_host;
  if (!is_pinned(action_host)) { memcpy(h->h_action, action_host, n * A * sizeof(float)); src = h->h_action; }
  cudaError_t ce = cudaMemcpyAsync(h->d_action, src, n * A * sizeof(float), cudaMemcpyHostToDevice, st);
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "H2D", ce);
  rc = pih_step(h, h->d_action, h->d_obs, h->d_reward, h->d_done, stream);
  if (rc) return rc;
  // one D2H straight into the caller's memory when it is one pinned block laid out [obs | reward | done]
  const bool direct = obs_host && reward_host && done_host && reward_host == obs_host + n * 5 &&
                      done_host == reinterpret_cast<uint8_t*>(reward_host + n) && is_pinned(obs_host);
  ce = cudaMemcpyAsync(direct ? obs_host : h->h_obs, h->d_obs, out_bytes(n), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "D2H", ce);
  if (!direct) {
    if (obs_host) memcpy(obs_host, h->h_obs, n * 5 * sizeof(float));
    if (reward_host) memcpy(reward_host, h->h_reward, n * sizeof(float));
    if (done_host) memcpy(done_host, h->h_done, n);
  }
  return PIH_OK;
}

int pih_rollout(pih_handle* h, int32_t steps, float* obs_dev, float* reward_dev, uint8_t* done_dev, unsigned long long* counters_dev, void* stream) {
  if (!h || steps < 0) return PIH_ERR_INVALID;
  if (!h->was_reset) return fail(h, PIH_ERR_STATE, "rollout() before reset()", cudaSuccess);
  USE_DEVICE(h);
  if (h->pipe) {
    // the pipe step is two launches (thread-per-env arm half, warp-per-env chain): `steps` of them, Philox actions in-kernel
    for (int t = 0; t < steps; ++t) {
      pih_pipe_launch_step(h->P, nullptr, h->d_pipe_scratch, obs_dev, reward_dev, done_dev, counters_dev, (cudaStream_t)stream);
      h->P.global_step += 1;
    }
    cudaError_t cep = cudaGetLastError();
    if (cep != cudaSuccess) return fail(h, PIH_ERR_CUDA, "pipe rollout launch", cep);
    return PIH_OK;
  }
  if (h->ws) {
    // small batches: `steps` back-to-back launches of the step kernel (actions from Philox in-kernel), chained by
    // programmatic dependent launch.  A kernel that loops over the steps was measured slower per step (52 against 44 us
    // at 4096 envs): what the enclosing loop keeps alive takes the spare registers the scheduler needs in the solver loop.
    StepSignal sig = {};
    for (int t = 0; t < steps; ++t) {
      LAUNCH_WS(k_step_ws, h, stream, h->P, (const float*)nullptr, obs_dev, reward_dev, done_dev, sig, counters_dev);
      h->P.global_step += 1;
    }
    cudaError_t cew = cudaGetLastError();
    if (cew != cudaSuccess) return fail(h, PIH_ERR_CUDA, "rollout launch", cew);
    return PIH_OK;
  }
  else if (h->split && h->P.task != PIH_TASK_SCRIPTED) {
    // large batches: the two-kernel step (see k_apply_action), `steps` times; at these sizes a launch is noise
    if (!h->d_qt && cudaMalloc(&h->d_qt, (size_t)h->P.n * PIH_NB_MAX * sizeof(float)) != cudaSuccess)
      return fail(h, PIH_ERR_CUDA, "target scratch allocation", cudaGetLastError());
    for (int t = 0; t < steps; ++t) {
      LAUNCH_PF(k_apply_action, h, stream, h->P, (const float*)nullptr, h->d_qt);
      LAUNCH_PF(k_step_dyn, h, stream, h->P, h->d_qt, obs_dev, reward_dev, done_dev, counters_dev);
      h->P.global_step += 1;
    }
    cudaError_t ce2 = cudaGetLastError();
    if (ce2 != cudaSuccess) return fail(h, PIH_ERR_CUDA, "split rollout launch", ce2);
    return PIH_OK;
  }
  else { LAUNCH(k_rollout, h, stream, h->P, (int)steps, obs_dev, reward_dev, done_dev, counters_dev); }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_rollout launch", ce);
  h->P.global_step += (unsigned)steps;
  return PIH_OK;
}

int pih_set_stats(pih_handle* h, unsigned long long* stats_dev) {
  if (!h) return PIH_ERR_INVALID;
  h->P.stats = stats_dev;
  return PIH_OK;
}

int pih_ik_targets(pih_handle* h, const float* action_dev, float* qtarget_dev, void* stream) {
  if (!h || !action_dev || !qtarget_dev) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  LAUNCH(k_ik, h, stream, h->P, action_dev, qtarget_dev);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_ik launch", ce);
  return PIH_OK;
}

int pih_ee_pose(pih_handle* h, float* pos_dev, float* rot_dev, void* stream) {
  if (!h) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  LAUNCH_ARM(k_ee, h, stream, h->P, pos_dev, rot_dev);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_ee launch", ce);
  return PIH_OK;
}

static int render_impl(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, int latched, void* stream);
int pih_render(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, void* stream) { return render_impl(h, width, height, rgbd_dev, 0, stream); }
int pih_render_grasp(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, void* stream) {
  if (h && h->P.task != PIH_TASK_SCRIPTED) return fail(h, PIH_ERR_INVALID, "pih_render_grasp: scripted task only", cudaSuccess);
  return render_impl(h, width, height, rgbd_dev, 1, stream);
}

static int render_impl(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, int latched, void* stream) {
  if (!h || !rgbd_dev || width <= 0 || height <= 0) return PIH_ERR_INVALID;
  if (!h->was_reset) return fail(h, PIH_ERR_STATE, "render() before reset()", cudaSuccess);
  USE_DEVICE(h);
  if (!h->d_scene) {
    cudaError_t ce = cudaMalloc(&h->d_scene, (size_t)h->P.n * PIH_SCENE_WORDS * sizeof(float));
    if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "scene allocation", ce);
  }
  LAUNCH_ARM(k_scene, h, stream, h->P, h->d_scene, (int)width, (int)height, latched);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_scene launch", ce);
  dim3 grid((unsigned)((width * height + 256 * PIH_RENDER_PPT - 1) / (256 * PIH_RENDER_PPT)), (unsigned)h->P.n);
  cudaStream_t st = (cudaStream_t)stream;
  float4* out = reinterpret_cast<float4*>(rgbd_dev);
#define RLAUNCH(ARMV)                                                                                              \
  do {                                                                                                             \
    if (h->P.task == PIH_TASK_RANDOM_FLY) k_render<ARMV, PIH_TASK_RANDOM_FLY><<<grid, 256, 0, st>>>(h->P, h->d_scene, width, height, out); \
    else k_render<ARMV, PIH_TASK_PEG_IN_HOLE><<<grid, 256, 0, st>>>(h->P, h->d_scene, width, height, out);         \
  } while (0)
  if (h->P.arm == PIH_ARM_PANDA) RLAUNCH(PIH_ARM_PANDA); else RLAUNCH(PIH_ARM_UR5);
#undef RLAUNCH
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_render launch", ce);
  return PIH_OK;
}

// learner buffer shared between the rank processes of one node: allocated by its owner with cudaMalloc (a whole
// allocation, so that its IPC handle maps exactly this buffer) and opened by every other rank ON THAT RANK'S OWN device,
// which maps the owner's HBM into the opener's address space as peer memory (NVLink)
int pih_peer_alloc(int32_t device, uint64_t bytes, void** ptr_out, uint8_t handle_out[64]) {
  if (!ptr_out || !handle_out || bytes == 0) return PIH_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  DeviceGuard guard_(device);
  if (!guard_.ok || cudaMalloc(ptr_out, (size_t)bytes) != cudaSuccess ||
      cudaMemset(*ptr_out, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaIpcGetMemHandle(&hd, *ptr_out) != cudaSuccess) { cudaGetLastError(); return PIH_ERR_CUDA; }
  memcpy(handle_out, &hd, 64);
  return PIH_OK;
}
int pih_peer_open(int32_t device, const uint8_t handle[64], void** ptr_out) {
  if (!ptr_out || !handle) return PIH_ERR_INVALID;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  DeviceGuard guard_(device);
  if (!guard_.ok || cudaIpcOpenMemHandle(ptr_out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return PIH_ERR_CUDA;
  }
  return PIH_OK;
}
int pih_peer_release(void* ptr, int32_t is_owner) {
  if (!ptr) return PIH_OK;
  cudaError_t ce = is_owner ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
  if (ce != cudaSuccess) { cudaGetLastError(); return PIH_ERR_CUDA; }
  return PIH_OK;
}

int pih_enable_peer(pih_handle* h, int32_t peer_device) {
  if (!h) return PIH_ERR_INVALID;
  if (peer_device == h->cfg.device) return PIH_OK;
  USE_DEVICE(h);
  int can = 0;
  cudaError_t ce = cudaDeviceCanAccessPeer(&can, h->cfg.device, peer_device);
  if (ce != cudaSuccess || !can) return fail(h, PIH_ERR_CUDA, "no peer access to that device", ce);
  ce = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (ce == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ce = cudaSuccess; }
  return ce == cudaSuccess ? PIH_OK : fail(h, PIH_ERR_CUDA, "cudaDeviceEnablePeerAccess", ce);
}

int pih_signal_flag(pih_handle* h, uint32_t* flag_dev, uint32_t value, void* stream) {
  if (!h || !flag_dev) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  k_signal_flag<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? PIH_OK : fail(h, PIH_ERR_CUDA, "k_signal_flag launch", ce);
}

int pih_wait_flags(pih_handle* h, const uint32_t* flags_dev, int32_t count, uint32_t value, uint32_t timeout_ms, uint32_t* timed_out_dev,
                   void* stream) {
  if (!h || !flags_dev || !timed_out_dev || count <= 0 || count > 1024) return PIH_ERR_INVALID;
  USE_DEVICE(h);
  k_wait_flags<<<1, ((count + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags_dev, (int)count, value, (unsigned long long)timeout_ms * 1000000ull,
                                                                        timed_out_dev);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? PIH_OK : fail(h, PIH_ERR_CUDA, "k_wait_flags launch", ce);
}

int pih_grasp_labels(pih_handle* h, int32_t width, int32_t height, float* labels_dev, float* params_dev, void* stream) {
  if (!h || !labels_dev || width <= 0 || height <= 0) return PIH_ERR_INVALID;
  if (!h->was_reset) return fail(h, PIH_ERR_STATE, "grasp_labels() before reset()", cudaSuccess);
  USE_DEVICE(h);
  dim3 grid((unsigned)((width * height + 256 * PIH_RENDER_PPT - 1) / (256 * PIH_RENDER_PPT)), (unsigned)h->P.n);
  k_labels<<<grid, 256, 0, (cudaStream_t)stream>>>(h->P, (int)width, (int)height, labels_dev, params_dev);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(h, PIH_ERR_CUDA, "k_labels launch", ce);
  return PIH_OK;
}

#ifdef PIH_PHASE_TICKS
// experiment builds only: SM clock values recorded by CTA 0 at the phase boundaries of the last step kernel
int pih_debug_ticks(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_ticks, 16 * sizeof(long long)) == cudaSuccess ? PIH_OK : PIH_ERR_CUDA;
}
int pih_debug_cta_ns(unsigned long long* out2048) {
  return cudaMemcpyFromSymbol(out2048, g_cta_ns, 2048 * sizeof(unsigned long long)) == cudaSuccess ? PIH_OK : PIH_ERR_CUDA;
}
#endif

int pih_measure_fp32_peak(int device, double* tflops_out) {
  if (!tflops_out) return PIH_ERR_INVALID;
  DeviceGuard guard_(device);
  if (!guard_.ok) return PIH_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PIH_ERR_CUDA;
  float* d_out = nullptr;
  if (cudaMalloc(&d_out, 4) != cudaSuccess) return PIH_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_fma_peak<<<blocks, threads>>>(d_out, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d_out); return PIH_ERR_CUDA; }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * (double)blocks * threads * (double)iters * 16.0 * 8.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops_out = best;
  return PIH_OK;
}

}  // extern "C"
