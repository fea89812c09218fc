// npk_holdem_launch.h -- parameter blocks and host-callable launchers of the kernels in npk_holdem.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/npk_holdem.h"
#include "npk_device.cuh"

namespace npk {

struct HoldemInit {
    int n_players, max_raises;
    double initial_stacks, small_blind, big_blind;
    uint8_t autoplay[NPK_MAX_SEATS];
};

struct HoldemAgents {
    uint8_t kind[NPK_MAX_SEATS];
    double min_call_equity[NPK_MAX_SEATS];
    double min_bet_equity[NPK_MAX_SEATS];
};

cudaError_t launch_holdem_init(const DeviceTables& tab, void* tables, long long n, const HoldemInit& cfg, uint64_t seed,
                               long long table_offset, cudaStream_t s);
cudaError_t launch_holdem_reset_done(const DeviceTables& tab, void* tables, long long n, uint64_t seed, long long table_offset,
                                     cudaStream_t s);
cudaError_t launch_holdem_step(const DeviceTables& tab, void* tables, long long n, const int8_t* actions, double* rewards,
                               uint64_t seed, long long table_offset, int restart_finished, cudaStream_t s);
cudaError_t launch_holdem_attach(void* tables, long long n, double* stage_data, cudaStream_t s);
cudaError_t launch_holdem_observe(const void* tables, long long n, const double* equity, double* obs, cudaStream_t s);
cudaError_t launch_holdem_queries(const void* tables, long long n, uint8_t* hole, uint8_t* board, uint8_t* n_players,
                                  uint8_t* active, cudaStream_t s);
cudaError_t launch_holdem_decide(const DeviceTables& tab, const void* tables, long long n, const uint64_t* wins,
                                 const uint64_t* ties, long long runs, const double* equity, const HoldemAgents& ag,
                                 uint64_t seed, unsigned long long decision_counter, long long table_offset, int8_t* actions,
                                 cudaStream_t s);

}  // namespace npk
