/* pih.h - C ABI of the B200-native batched simulator that replaces peg-in-hole-gym's vectorised env.step.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Each entry
 * point names the reference interface it replaces (paths relative to the reference repository).  The Python
 * host (peg_in_hole_gym_b200/envs/*.py) binds these with ctypes and mirrors the reference's gym surface;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * All device pointers are CUDA device pointers on the handle's device; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  Calls are asynchronous on that stream unless stated; a handle is not
 * thread-safe; one handle per GPU / rank.  Every function returns 0 on success, <0 on error
 * (see pih_last_error).  There is NO CPU fallback: without a CUDA device pih_create fails.
 */
#ifndef PIH_H
#define PIH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIH_ABI_VERSION 2

/* error codes */
#define PIH_OK 0
#define PIH_ERR_INVALID (-1)   /* bad argument / config (reference: assert in BaseEnv.__init__, base_env.py:16-17) */
#define PIH_ERR_CUDA (-2)      /* CUDA runtime error or no device */
#define PIH_ERR_STATE (-3)     /* step before reset (reference: AttributeError, base_env_mp.py:42) */

/* Construction arguments: the kwargs of gym.make('peg-in-hole-mp-v0', client, task, mp_num, sub_num, offset,
 * args, is_test) (README.md:38, base_env_mp.py:15, base_env.py:15) plus the NEW knobs of SURVEY.md Appendix D. */
typedef struct pih_config {
  int32_t task;             /* 0 'peg-in-hole' (base_env.py:9-11), 1 'random-fly' (README.md:38) */
  int32_t arm;              /* 0 panda (utils.py:30-38), 1 ur5 (utils.py:40-48; task 0 only) */
  int32_t fly_object;       /* args[0]: 0 'Banana', 1 'Amicelli' (README.md:38) */
  int32_t num_envs;         /* envs resident on this device = local mp_num * sub_num */
  int32_t sub_num;          /* envs per reference client; period of the offset layout (base_env.py:35-55) */
  int32_t env_id_base;      /* global id of local env 0 (rank sharding; RNG is keyed by global id) */
  float offset[3];          /* base_env.py:15 offset */
  float dt;                 /* 1/240 (peg_in_hole.py:254) or args[1] (README.md:38) */
  float solver_residual;    /* PyBullet solverResidualThreshold [3P], default 1e-7: PGS stops early once the largest squared
                               velocity change of any row in an iteration is <= this; 0 = always run solver_iters */
  int32_t solver_iters;     /* Bullet numSolverIterations, default 50 */
  int32_t ik_iters;         /* calculateInverseKinematics maxNumIterations, default 20 */
  int32_t max_steps;        /* NEW time limit per episode, 0 = none (reference has none) */
  int32_t auto_reset;       /* NEW 0 = reference behaviour (done envs frozen until reset, base_env.py:61-71) */
  int32_t action_frame_env; /* NEW 0 = reference (actions are absolute world coordinates), 1 = relative to env origin */
  int32_t device;           /* CUDA device ordinal */
  uint64_t seed;            /* Philox key (reference: unseeded global MT19937, peg_in_hole.py:239) */
  int32_t kernel_family;    /* NEW which step kernels run: 0 = chosen from num_envs (default), 1 = warp-specialised small-batch kernels
                               (three warps per 32 envs: action / dynamics / contacts), 2 = fused thread-per-env kernel, 3 = split
                               apply_action + dynamics kernels, 4 = warp-specialised with two warps per 32 envs (action / dynamics).
                               All families compute the same bits (tests/test_gpu_parity.py); the field exists so that every family
                               can be tested at any size */
  int32_t host_path;        /* NEW pih_step_host: 0 = zero-copy when the caller's buffers are page-locked, else staged; 1 = always staged */
  int32_t peg_model;        /* what the reference's `pipe.urdf` object is simulated as (peg_in_hole.py:240-245): 0 = one rigid end
                               segment (the round-1 stand-in), 1 = the 25-link chain with 23 continuous joints (panda, tasks 0 and 2).
                               The state buffer of model 1 has pih_state_words_for(cfg) words per env instead of pih_state_words() */
} pih_config;

#define PIH_KERNELS_AUTO 0
#define PIH_KERNELS_WARP_SPECIALISED 1
#define PIH_KERNELS_FUSED 2
#define PIH_KERNELS_SPLIT 3
#define PIH_KERNELS_WARP_SPECIALISED_2 4

typedef struct pih_handle pih_handle;

/* ABI / layout queries */
int pih_abi_version(void);
int pih_state_words(void);                 /* 32-bit words of state per env (= 4 * float4 slots), rigid peg model */
int pih_state_words_for(const pih_config* cfg);   /* the same for any config: 48; 64 for the scripted task (+ latched grasp scene); 128 with peg_model = 1 */
int pih_action_dim(int arm);               /* 4 panda (peg_in_hole.py:12), 6 ur5 (utils.py:70-72) */
int pih_obs_dim(void);                     /* 5 (peg_in_hole.py:13) */

/* BaseEnvMp.__init__ / BaseEnv.__init__ (base_env_mp.py:15-37, base_env.py:15-33).  `state_dev` is a caller-owned
 * (torch-owned) device buffer of pih_state_words()*num_envs 32-bit words, laid out [slot][env][4] (float4 SoA). */
int pih_create(const pih_config* cfg, void* state_dev, pih_handle** out);
/* BaseEnv.close (base_env.py:97-98) */
void pih_destroy(pih_handle* h);
const char* pih_last_error(const pih_handle* h);

/* Env placement, BaseEnv._create_env (base_env.py:35-55): writes the world offset of local env i to out[3]. Host only. */
int pih_env_offset(const pih_config* cfg, int32_t local_env, float out[3]);

/* BaseEnvMp.reset / BaseEnv.reset (base_env_mp.py:54-63, base_env.py:84-94).  mask_dev: nullable uint8[num_envs]
 * (NULL = every env, the reference behaviour).  Writes the first observation when obs_dev != NULL. */
int pih_reset(pih_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* BaseEnvMp.step -> worker -> BaseEnv.step (base_env_mp.py:40-51,75-87, base_env.py:60-75): apply_action
 * (panda_execute / ur_execute, utils.py:60-82), one stepSimulation, get_info.  action_dev float[num_envs][A];
 * obs_dev float[num_envs][5]; reward_dev float[num_envs]; done_dev uint8[num_envs].  Envs that are done are not
 * stepped and report their last outputs again (base_env.py:61-71) unless auto_reset.  action_dev must be 16-byte
 * aligned (Box(4) rows are read as one 128-bit word) / 8-byte aligned (Box(6)). */
int pih_step(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream);

/* Same call with HOST buffers: the reference-facing plugin path (host lists in, host lists out); the stream is
 * synchronised before returning.  When all four buffers are page-locked and device-mapped (cudaHostAlloc /
 * cudaHostRegister, e.g. torch pin_memory) the step kernel reads the actions from and writes the results to host
 * memory itself: ONE launch, no staging copies (zero-copy over PCIe).  Otherwise pageable buffers are staged through
 * the handle's pinned block: H2D, step, ONE D2H of [obs n*5 f32 | reward n f32 | done n u8].  Done envs re-emit their
 * last outputs (they are a function of the frozen state), so the output buffers may be fresh every call. */
int pih_step_host(pih_handle* h, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host, void* stream);

/* Lifetime contract for the zero-copy path of pih_step_host.  A page-locked, device-mapped host range that the caller
 * BINDS stays resolved inside the handle (no per-call driver query) until it is unbound; the caller guarantees that the
 * memory stays allocated and registered for that long.  Buffers that are not bound are looked up in the driver on every
 * call, so a stale answer can never be used.  host_ptr == NULL in pih_host_unbind drops every binding. */
int pih_host_bind(pih_handle* h, const void* host_ptr, uint64_t bytes);
int pih_host_unbind(pih_handle* h, const void* host_ptr);

/* Execution statistics (NEW; measurement only).  stats_dev: nullable device uint64[8]; while set, every
 * step / rollout kernel adds to it: [0] env-steps, [1] IK iterations executed, [2] constraint rows solved summed over
 * solver iterations, [3] solver iterations, [4] contact points, [5] joint-limit rows, [6] envs force-reset by the
 * non-finite guard, [7] max solver iterations of any env-step.  The caller zeroes and reads the buffer. */
int pih_set_stats(pih_handle* h, unsigned long long* stats_dev);

/* Benchmark / synthetic rollouts: `steps` consecutive env.steps in ONE launch with actions
 * action_space.sample() ~ U[-1,1]^A (README.md:48) drawn in-kernel from Philox (seed; env, global step).
 * Outputs of the LAST step are written (nullable).  counters_dev: nullable uint64[2] += {env-steps, episodes done}. */
int pih_rollout(pih_handle* h, int32_t steps, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                unsigned long long* counters_dev, void* stream);

/* apply_action only (panda_execute/ur_execute up to the IK result, utils.py:60-82): writes the joint position
 * targets float[num_envs][9] without stepping.  Parity hook for SURVEY.md rows a9-a13. */
int pih_ik_targets(pih_handle* h, const float* action_dev, float* qtarget_dev, void* stream);

/* getLinkState(ee)[0:2] for every env (utils.py:62-63): pos float[num_envs][3] (world), rot float[num_envs][9]. */
int pih_ee_pose(pih_handle* h, float* pos_dev, float* rot_dev, void* stream);

/* PegInHole.render (peg_in_hole.py:276-304) for every env: camera at the end effector looking along -z, FOV 60 deg,
 * near 0.001, far 1000.  rgbd_dev float[num_envs][height][width][4] = (OpenGL depth buffer value, r, g, b in 0..255),
 * the layout of the reference's np.concatenate((depth, rgb), 2).  The reference renders 300 x 300. */
int pih_render(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, void* stream);

/* Scripted task only: the same camera on the scene LATCHED when the grasp state machine entered state 2 - the reference's
 * `self.grasp_img = self.render()` (peg_in_hole.py:65-66), the observation the shipped get_info returns (peg_in_hole.py:33-37)
 * together with the label images below.  Envs that have not reached state 2 yet show the present scene. */
int pih_render_grasp(pih_handle* h, int32_t width, int32_t height, float* rgbd_dev, void* stream);

/* The label images of the shipped task's random_grasp (peg_in_hole.py:41-45,65-98), for every env: a rotated
 * rectangle (half extents 0.1 x 0.2 in normalised image units; angle captured by the scripted state machine when it
 * entered state 2) filled into labels_dev float[num_envs][4][height][width] = (pos_img: 50 inside, sin(2 angle),
 * cos(2 angle): 1 outside, wid_img: rectangle width in pixels); params_dev (nullable) float[num_envs][5] =
 * (x = 0, y = 0, angle in degrees, width, length in pixels), the second element of the reference's return value.
 * Before state 2 has been entered the images are the reference's initial ones (zeros, cos = 1). */
int pih_grasp_labels(pih_handle* h, int32_t width, int32_t height, float* labels_dev, float* params_dev, void* stream);

/* Gather to a rank-0 learner buffer WITHOUT a collective (preferred form).  The reference gathers the
 * workers' results through multiprocessing queues (base_env_mp.py:44-50).  Here every rank hands pih_step output
 * pointers that are PEER mappings of its slice of rank 0's buffer, so the step kernel itself stores the rows over
 * NVLink; pih_signal_flag then publishes `value` (the step number) in the rank's flag word inside that buffer
 * (stream-ordered after the step, fenced system-wide), and rank 0 enqueues pih_wait_flags, which returns control to
 * its stream once all `count` flags have reached `value`.  A wait that exceeds timeout_ms writes 1 + the index of the
 * late flag to *timed_out_dev and gives up (no unbounded device-side spin). */
/* pih_step whose kernel also publishes flag_value in *flag_dev once all its rows have been stored (fused epilogue;
 * for batches that run the classic kernels a one-thread signalling kernel follows instead). */
int pih_step_signal(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                    uint32_t flag_value, void* stream);
/* The learner's rank (the owner of the buffer): pih_step_signal whose kernel FIRST waits until `wait_count` flags have
 * reached `wait_value` - the consumer's wait for the previous step's rows of every rank, folded into its next launch
 * instead of a pih_wait_flags kernel of its own (bounded by timeout_ms like pih_wait_flags). */
int pih_step_signal_wait(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                         uint32_t flag_value, const uint32_t* wait_flags_dev, int32_t wait_count, uint32_t wait_value, uint32_t timeout_ms,
                         uint32_t* timed_out_dev, void* stream);

/* Deferred publication of the same gather (NEW).  pih_step_signal makes every CTA of the step fence at system scope before
 * it leaves (its rows must have landed in the learner's memory before the flag may follow): ~6 us per step at 4096 envs.
 * Here the step stores its rows and leaves; the NEXT launch of the stream announces them: at its start (the previous
 * launch has completed by stream order) one thread fences once and stores prev_value into flag_dev.  So a loop of
 * pih_step_publish_prev(..., prev_value = t - 1, ...) for t = 1, 2, ... delivers the flags one launch later and costs
 * the steps nothing; the rows of the last step are announced by pih_signal_flag.  wait_flags_dev (nullable) as in
 * pih_step_signal_wait: the learner's rank waits for wait_count flags >= wait_value before it touches anything. */
int pih_step_publish_prev(pih_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, uint32_t* flag_dev,
                          uint32_t prev_value, const uint32_t* wait_flags_dev, int32_t wait_count, uint32_t wait_value, uint32_t timeout_ms,
                          uint32_t* timed_out_dev, void* stream);
int pih_peer_alloc(int32_t device, uint64_t bytes, void** ptr_out, uint8_t handle_out[64]);  /* owner: zeroed buffer + its IPC handle */
int pih_peer_open(int32_t device, const uint8_t handle[64], void** ptr_out);   /* other rank: map it on ITS device (peer memory) */
int pih_peer_release(void* ptr, int32_t is_owner);
int pih_enable_peer(pih_handle* h, int32_t peer_device);   /* same process, two devices: let this handle's kernels address peer_device */
int pih_signal_flag(pih_handle* h, uint32_t* flag_dev, uint32_t value, void* stream);
int pih_wait_flags(pih_handle* h, const uint32_t* flags_dev, int32_t count, uint32_t value, uint32_t timeout_ms,
                   uint32_t* timed_out_dev, void* stream);

/* FP32 FMA-throughput micro-benchmark (roofline denominator): returns measured TFLOP/s. */
int pih_measure_fp32_peak(int device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif
