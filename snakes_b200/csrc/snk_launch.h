// snk_launch.h -- host-visible launch interface between snk_api.cu and snk_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/snk.h"

struct Params;
struct PeerArgs;
enum { KIND_LANE = 0, KIND_TILE = 1, KIND_DENSE = 2, KIND_ROWS = 3 };

struct LaunchPlan {
  int kind;         // KIND_LANE (lane per env, chain bodies), KIND_TILE (warp per env), KIND_ROWS / KIND_DENSE (CTA per env)
  bool ws;          // lane path as ONE warp-specialised kernel (k_step_lane_ws): paint warps + logic warps per CTA
  bool pdl;         // fused lane kernel launched with programmatic stream serialization (prologue overlaps the previous launch's tail)
  bool split;       // lane path as two kernels: k_lane_logic (thread per env) + k_lane_paint (observation writer)
  bool paint2;      // split form: the observation writer is k_lane_paint2 (two warps per image buffer, programmatic launches)
  int grid, block;
  size_t smem;
  int occupancy;    // resident CTAs per SM
  int max_grid;     // SMs * occupancy: persistent grid size
};

cudaError_t snk_plan(int rules, LaunchPlan& plan, int n_sm, int S, int K);
bool snk_lane_supported(int S, int K);
cudaError_t snk_launch_step(const Params& p, int rules, const LaunchPlan& plan, cudaStream_t stream);
cudaError_t snk_launch_upscale84(const uint8_t* native, uint8_t* out, long long N, int V, int C, int n_sm, cudaStream_t stream);
cudaError_t snk_launch_dump(const Params& p, uint8_t* blob, const snk_state_layout& lay, long long first, long long count, cudaStream_t stream);
cudaError_t snk_launch_sum_inbox(const PeerArgs& a, const double* stats, double* glob, cudaStream_t stream);
cudaError_t snk_launch_extract_views(const uint8_t* src, uint8_t* dst, long long n_pixels, int C, int n_out, cudaStream_t stream);
cudaError_t snk_launch_load(const Params& p, const uint8_t* blob, const snk_state_layout& lay, cudaStream_t stream);
cudaError_t snk_launch_scripted_actions(const Params& p, int8_t* actions, uint64_t step, uint64_t seed, int eps_permille,
                                        const uint8_t* occ_obs, cudaStream_t stream);
cudaError_t snk_launch_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                           const uint8_t* last_dones, double gamma, double lam, int T, long long N, float* advs, float* returns,
                           cudaStream_t stream);
cudaError_t snk_launch_gen_actions(int8_t* actions, long long N, int S, long long env_id_base, uint64_t step,
                                   uint64_t seed, int n_actions, cudaStream_t stream);
