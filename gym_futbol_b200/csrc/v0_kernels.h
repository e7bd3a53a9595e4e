// Host-side launchers of the v0 kernels (defined in v0_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/futbol_b200.h"
#include "v0_step.cuh"

namespace futbol {
// How a rollout is launched: kernel 0 = standard (20 warps per SM), 1 = dense (28 warps per SM); slices > 1 = the time-sliced
// work-queue launch of the standard kernel.
struct V0RolloutChoice { int kernel; int slices; };
size_t v0_state_bytes(int n_envs);
cudaError_t v0_launch_reset(const V0Params &P, void *state, const uint8_t *mask, void *obs, int obs_f64, int init,
                            cudaStream_t st);
cudaError_t v0_launch_step(const V0Params &P, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs,
                           void *reward, uint8_t *done, void *final_obs, int out_f64, cudaStream_t st);
cudaError_t v0_launch_rollout(const V0Params &P, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions,
                              float *obs, float *reward, uint8_t *done, FutbolStats *stats, int slices, int variant, cudaStream_t st);
V0RolloutChoice v0_plan_rollout(const V0Params &P, int K, int slices, int variant);
cudaError_t v0_launch_get_state(int n, const void *state, void *aos, cudaStream_t st);
cudaError_t v0_launch_set_state(int n, void *state, const void *aos, cudaStream_t st);
}  // namespace futbol
