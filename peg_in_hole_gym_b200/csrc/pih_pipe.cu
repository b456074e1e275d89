// pih_pipe.cu - translation unit of the flexible-pipe kernels (pih_pipe.cuh); launched through pih_pipe_host.h by the
// C ABI in pih_kernels.cu.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
#include "pih_pipe.cuh"
#include "pih_pipe_host.h"

using namespace pih;
#include "pih_tables_host.h"

cudaError_t pih_pipe_upload_tables() {
  // the step kernel's CTA holds PIH_PIPE_WARPS warp-envs: its dynamic shared memory can exceed the 48 KB default limit
  cudaFuncSetAttribute(k_pipe_step<PIH_ARM_PANDA, PIH_TASK_PEG_IN_HOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIH_PIPE_WARPS * kPipeSmemBytes);
  cudaFuncSetAttribute(k_pipe_step<PIH_ARM_PANDA, PIH_TASK_SCRIPTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIH_PIPE_WARPS * kPipeSmemBytes);
  return upload_tables();
}

size_t pih_pipe_scratch_bytes(int num_envs) { return (size_t)num_envs * (PIH_PA_WORDS + kRowBufWords) * sizeof(float); }

// the pipe exists for the panda tasks that use the reference's pipe.urdf: 'peg-in-hole' and the scripted episode
#define PIPE_DISPATCH(P, CALL)                                                        \
  do {                                                                                \
    if ((P).task == PIH_TASK_SCRIPTED) { constexpr int TASK = PIH_TASK_SCRIPTED; CALL; } \
    else { constexpr int TASK = PIH_TASK_PEG_IN_HOLE; CALL; }                          \
  } while (0)

void pih_pipe_launch_reset(const SimParams& P, const unsigned char* mask, float* obs, cudaStream_t st) {
  PIPE_DISPATCH(P, (k_pipe_reset<PIH_ARM_PANDA, TASK><<<P.n, 32, kPipeSmemBytes, st>>>(P, mask, obs)));
}

void pih_pipe_launch_step(const SimParams& P, const float* action, float* scratch, float* obs, float* reward, unsigned char* done,
                          unsigned long long* counters, cudaStream_t st) {
  PIPE_DISPATCH(P, (k_pipe_arm<PIH_ARM_PANDA, TASK><<<(P.n + 31) / 32, 32, 0, st>>>(P, action, scratch)));
  float* rowbuf = scratch + (size_t)P.n * PIH_PA_WORDS;     // [n][3 * PIH_PIPE_MAXC][96]: constraint rows of the solver
  PIPE_DISPATCH(P, (k_pipe_step<PIH_ARM_PANDA, TASK><<<(P.n + PIH_PIPE_WARPS - 1) / PIH_PIPE_WARPS, 32 * PIH_PIPE_WARPS, PIH_PIPE_WARPS * kPipeSmemBytes, st>>>(P, scratch, rowbuf, obs, reward, done, counters)));
}
