// Host-side launchers of the v1 kernels (defined in v1_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/futbol_b200.h"
#include "v1_step.cuh"

namespace futbol {
namespace v1 {
size_t state_bytes(int n_envs, int n_players);
cudaError_t launch_reset(const V1Params &P, void *state, const uint8_t *mask, void *obs, int obs_f64, int init, cudaStream_t st);
cudaError_t launch_step(const V1Params &P, void *state, const uint8_t *actions, const uint8_t *opp_actions, void *obs, void *reward,
                        uint8_t *done, void *final_obs, int out_f64, cudaStream_t st);
cudaError_t launch_rollout(const V1Params &P, void *state, int K, const uint8_t *actions, const uint8_t *opp_actions, float *obs,
                           float *reward, uint8_t *done, FutbolStats *stats, int slices, cudaStream_t st);
int plan_rollout_slices(const V1Params &P, int K, int slices);   // time slices launch_rollout will use (1 = the plain kernel)
size_t env_state_bytes(int n_players);    // one AoS record: FutbolV1EnvState header + the env's arbiter cache
cudaError_t launch_get_state(int n, int n_players, const void *state, void *aos, cudaStream_t st);
cudaError_t launch_set_state(int n, int n_players, void *state, const void *aos, cudaStream_t st);
}  // namespace v1
}  // namespace futbol
