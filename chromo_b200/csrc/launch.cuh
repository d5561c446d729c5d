// launch.cuh -- launch plumbing shared by the translation units.
// The MC kernels are instantiated for (RNG mode) x (number of binders 1..4);
// to keep nvcc wall-time down each (RNG, nb-pair) group is its own translation
// unit (mc_inst.cu compiled four times, see __graft_entry__.build()).
#pragma once
#ifdef CHROMO_HOST_EMU
// test-only lockstep emulator (tests/host_emu); never part of the product build
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define CB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define CB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CB_NOINLINE __noinline__
#define CB_GRID_CONSTANT __grid_constant__
__device__ __forceinline__ void cb_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif
#include "params.cuh"

struct McSimArgs {
    DevCtx d;
    long long num_mc_steps;
    double mu_adjust;
    unsigned long long seed;
    int cap;
    size_t smem;
    cudaStream_t stream;
};
struct McStepArgs {
    DevCtx d;
    int replica, move;
    double amp_move;
    int amp_bead;
    double mu_adjust;
    unsigned long long seed;
    int force_accept;
    DebugOut *dbg;
    int cap;
    size_t smem;
    cudaStream_t stream;
};
// each returns a cudaError_t as int; defined in mc_inst.cu
int cb_mc_sim_replay_12(const McSimArgs &a);
int cb_mc_sim_replay_34(const McSimArgs &a);
int cb_mc_sim_philox_12(const McSimArgs &a);
int cb_mc_sim_philox_34(const McSimArgs &a);
int cb_mc_step_replay_12(const McStepArgs &a);
int cb_mc_step_replay_34(const McStepArgs &a);
int cb_mc_step_philox_12(const McStepArgs &a);
int cb_mc_step_philox_34(const McStepArgs &a);
