// mc_inst.cu -- instantiates the fused MC kernels for one RNG mode and one pair
// of binder counts.  Compile with
//   -DCB_INST_REPLAY=0|1  -DCB_INST_HI=0|1     (HI: nb in {3,4}, else {1,2})
#include "launch.cuh"
#ifndef CB_TWIST
#define CB_TWIST 0
#endif
#include "geometry.cuh"
#include "params.cuh"
#include "rng.cuh"
#if CB_TWIST
// the twist build of the kernels lives in its own namespace: same template arguments as the default
// build, different symbols (kernels, and every inline function whose body depends on CB_TWIST)
#define TW_ _tw
namespace cb_twist_build {
#include "mc_kernel.cuh"
}
using namespace cb_twist_build;
#else
#define TW_
#include "mc_kernel.cuh"
#endif

#ifndef CB_INST_REPLAY
#error "define CB_INST_REPLAY"
#endif
#if CB_INST_REPLAY
typedef ReplayRng InstRng;
#define NAME3(a, t, b) cb_mc_##a##_replay##t##_##b
#else
typedef PhiloxRng InstRng;
#define NAME3(a, t, b) cb_mc_##a##_philox##t##_##b
#endif
#define NAME3X(a, t, b) NAME3(a, t, b)
#define NAME2(a, b) NAME3X(a, TW_, b)
#if CB_INST_HI
#define NAME(a) NAME2(a, 34)
constexpr int NB_A = 3, NB_B = 4;
#else
#define NAME(a) NAME2(a, 12)
constexpr int NB_A = 1, NB_B = 2;
#endif

template <int NB, int NW>
static int sim_launch(const McSimArgs &a) {
    auto k = mc_sim_kernel<InstRng, NB, NW>;
    const size_t smem = (size_t)a.rpb * cb_replica_smem(a.cap, a.d.ncol, NW);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int nrep = a.nrep < 0 ? a.d.R - a.rep0 : a.nrep;
    if (nrep <= 0) return 0;
    CB_LAUNCH(k, (nrep + a.rpb - 1) / a.rpb, 32 * NW * a.rpb, smem, a.stream, a.d, a.num_mc_steps, a.mu_adjust,
              a.seed, a.cap, a.rpb, a.rep0, a.rep0 + nrep);
    return (int)cudaGetLastError();
}
template <int NB>
static int sim_one(const McSimArgs &a) {
#if !CB_INST_REPLAY
    if (a.warps == 2) return sim_launch<NB, 2>(a);
#endif
    return sim_launch<NB, 1>(a);
}
template <int NB>
static int step_one(const McStepArgs &a) {
    auto k = mc_step_kernel<InstRng, NB>;
    const size_t smem = cb_table_bytes(a.cap, a.d.ncol);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    CB_LAUNCH(k, 1, 32, smem, a.stream, a.d, a.replica, a.move, a.amp_move, a.amp_bead, a.mu_adjust,
              a.seed, a.force_accept, a.dbg, a.cap);
    return (int)cudaGetLastError();
}

int NAME(sim)(const McSimArgs &a) { return a.d.nb == NB_A ? sim_one<NB_A>(a) : sim_one<NB_B>(a); }
int NAME(step)(const McStepArgs &a) { return a.d.nb == NB_A ? step_one<NB_A>(a) : step_one<NB_B>(a); }

#if defined(CB_PHASE_TIMERS) && !CB_INST_REPLAY && !CB_INST_HI && !CB_TWIST
// development only: read and reset the phase timers of the Philox nb<=2 kernels
extern "C" int cb_phase_read(unsigned long long *out) {
    cudaError_t e = cudaMemcpyFromSymbol(out, cb_phase_acc, sizeof(unsigned long long) * CHROMO_NUM_MOVES * CB_NPHASE);
    if (e != cudaSuccess) return (int)e;
    static unsigned long long zero[CHROMO_NUM_MOVES * CB_NPHASE];
    return (int)cudaMemcpyToSymbol(cb_phase_acc, zero, sizeof(zero));
}
#endif
