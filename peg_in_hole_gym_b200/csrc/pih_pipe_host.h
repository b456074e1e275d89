// pih_pipe_host.h - internal interface between the C ABI (pih_kernels.cu) and the flexible-pipe translation unit
// (pih_pipe.cu).  Not part of the boundary: include/pih.h is.
#pragma once
#include <cuda_runtime.h>

namespace pih { struct SimParams; }

// uploads this translation unit's copy of the model tables to the current device
cudaError_t pih_pipe_upload_tables();
// bytes of the per-env scratch: the block handed from k_pipe_arm to k_pipe_step + the solver's constraint rows
size_t pih_pipe_scratch_bytes(int num_envs);
// PegInHole.reset with the pipe (mask nullable)
void pih_pipe_launch_reset(const pih::SimParams& P, const unsigned char* mask, float* obs, cudaStream_t st);
// one env.step: k_pipe_arm (thread per env) + k_pipe_step (warp per env).  action == nullptr: in-kernel Philox actions.
void pih_pipe_launch_step(const pih::SimParams& P, const float* action, float* scratch, float* obs, float* reward, unsigned char* done,
                          unsigned long long* counters, cudaStream_t st);
