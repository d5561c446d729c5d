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
#include <cstdint>
#define CB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define CB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CB_NOINLINE __noinline__
#define CB_GRID_CONSTANT __grid_constant__
__device__ __forceinline__ void cb_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cb_backoff() { __nanosleep(20); }
// 32-bit shared-state-space addresses for the fire-and-forget atomic adds of the delta-density cells: a
// generic pointer costs a 64-bit address computation and a generic-to-shared conversion per atomic
typedef uint32_t cb_saddr;
__device__ __forceinline__ cb_saddr cb_shared_addr(const void *p) { return (cb_saddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cb_red_add_u32(cb_saddr a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// named barrier `id` (1..15) for `count` threads of the block
__device__ __forceinline__ void cb_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif
#include <cstddef>
// shared-memory bytes of one warp's delta-density table: cap x (ncol 64-bit cells + key + list entry)
static inline __host__ __device__ size_t cb_table_bytes(int cap, int ncol) {
    return (size_t)cap * ((size_t)ncol * 8 + 8);
}
// shared memory of the MC kernel per replica / per warp (>= sizeof(ReplicaSh), sizeof(WarpSh);
// checked in mc_kernel.cuh) and the largest number of warps per replica that is instantiated
#ifndef CB_KSEL
// tangent rotations of up to this many beads take the prepared (lane-parallel) path; <= 32.  SimpleControl settles
// the window at 14-18 beads on the reference workloads: with 16 the ~6 % of attempts above it went down the
// sequential path (bead set and axis draws by one lane, index scratch in HBM) and cost 6 % of the whole step.
#define CB_KSEL 24
#endif
#define CB_REPLICA_SH_BYTES 6528
#define CB_WARP_SH_BYTES ((368 + 52 * CB_KSEL + 127) / 128 * 128)
#define CB_MAX_WARPS 2
// shared memory of one replica: ReplicaSh | WarpSh x warps | table x warps
static inline __host__ __device__ size_t cb_replica_smem(int cap, int ncol, int warps) {
    return CB_REPLICA_SH_BYTES + (size_t)warps * (CB_WARP_SH_BYTES + cb_table_bytes(cap, ncol));
}
// replicas that share one thread block (one block per SM; their warps go through the move types
// of a sweep together, see mc_sim_kernel)
#ifndef CB_MAX_RPB
#define CB_MAX_RPB 7
#endif
#include "params.cuh"

struct McSimArgs {
    DevCtx d;
    long long num_mc_steps;
    double mu_adjust;
    unsigned long long seed;
    int cap;   // slots of each warp's table
    int warps; // warps per replica (Philox kernels: 1 or 2; the replay kernels always run 1)
    int rpb;   // replicas per block (1..CB_MAX_RPB)
    cudaStream_t stream;
    int rep0 = 0, nrep = -1; // replica sub-range [rep0, rep0 + nrep) (nrep < 0: all of the context's)
    const int *order = nullptr; // move ids in the controller list's order (host; nullptr = all_moves' order)
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
// SSTWLC (twist) builds of mc_inst.cu (-DCB_TWIST=1), one or two binders
int cb_mc_sim_replay_tw_12(const McSimArgs &a);
int cb_mc_sim_philox_tw_12(const McSimArgs &a);
int cb_mc_step_replay_tw_12(const McStepArgs &a);
int cb_mc_step_philox_tw_12(const McStepArgs &a);
