// chromo_b200.cu -- host side of the C ABI declared in include/chromo_b200.h.
// One context = R replicas resident in the HBM of one B200; every compute entry
// point launches hand-written sm_100a kernels on the context's stream.  There is
// no CPU fallback: if CUDA is unavailable every call fails with CHROMO_ERR_CUDA.
#include "launch.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <cstdlib>
#include <vector>

#include "../../include/chromo_b200.h"
#include "field_kernels.cuh"
#include "rng.cuh"

static_assert(sizeof(chromo_move_state) == 88, "ABI: chromo_move_state");
static_assert(sizeof(chromo_shape) == 112, "ABI: chromo_shape");
static_assert(sizeof(chromo_step_report) == 48, "ABI: chromo_step_report");

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
// error sink of the other translation units (rediscretize.cu)
int cb_set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(CHROMO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

struct chromo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    chromo_shape shape{};
    DevCtx d{};
    std::vector<void *> allocs;
    int64_t bytes = 0;
    // owned device buffers that can be replaced
    double *d_bond = nullptr, *d_twist = nullptr, *d_chi = nullptr, *d_mu = nullptr, *d_bindF = nullptr, *d_access = nullptr;
    double *d_partial = nullptr, *d_out = nullptr, *d_detailed = nullptr;
    int move_order[CHROMO_NUM_MOVES] = {0, 1, 2, 3, 4}; // chromo_ctx_set_move_order
    int *d_dcount = nullptr;
    int *d_bad = nullptr; // set by the narrowing kernel when a state / mark is out of range
    // replica exchange (chromo_exchange_*): the chi ladder(s), the replica on every rung, counters
    double *d_ex_ladder = nullptr;
    int *d_ex_rung = nullptr;
    unsigned long long *d_ex_counters = nullptr;
    int64_t ex_total = 0, ex_first = 0, ex_ladder_len = 0;
    long long *d_stage = nullptr;
    int64_t stage_elems = 0;
    // host-array path (chromo_mc_sim_host): one stream + int64 staging buffer per replica chunk
    std::vector<cudaStream_t> chunk_streams;
    std::vector<cudaEvent_t> chunk_uploaded; // chunk k's upload is complete (chunk k+1's upload waits for it)
    std::vector<long long *> chunk_stage;
    int64_t chunk_stage_elems = 0;
    DebugOut *d_dbg = nullptr;
    long long *d_dbg_inds = nullptr, *d_dbg_touched = nullptr;
    double *d_dbg_rows = nullptr, *d_dbg_dtrial = nullptr;
    int64_t dbg_inds_cap = 0, dbg_touched_cap = 0;
    int nblk_bins = 1, nblk_bonds = 1;
    int cap = 0;           // slots of each warp's delta-density table in the MC kernel
    int warps = 1;            // warps per replica of the production (Philox) MC kernel (1 or 2)
    int rpb = 1;              // replicas per thread block
    int rpb_fixed = 0;        // > 0: set by chromo_ctx_set_replicas_per_block
    double roundK = 0.0;
    double min_access_vol = 0.0; // smallest positive per-voxel volume (0: uniform voxels)
    bool have_binders = false, have_bonds = false, have_state = false;
    bool have_mods = false; // every replica's chemical_mods have been uploaded
    bool force_l2_density = false; // test knob: full recompute through the L2-atomic kernel even on a small grid
    int64_t last_attempts = 0;
    int sm_count = 148;
    size_t smem_optin = 227 * 1024; // largest block
    size_t smem_sm = 228 * 1024;    // per SM
};

// exponent base of the MC kernel's fixed-point delta-density cells (mc_kernel.cuh, fx_format):
// floor(log2(V_min / max_state))
static void refresh_fx_base(chromo_ctx *c) {
    DevCtx &d = c->d;
    double vmin = (d.access_vol && c->min_access_vol > 0.0) ? c->min_access_vol : d.vol_bin;
    int smax = 1;
    for (int a = 0; a < d.nb; a++) smax = std::max(smax, d.sites[a]);
    d.fx_base = (vmin > 0.0) ? (int)ilogb(vmin / (double)smax) : 0;
}

// kernel instantiation for (rng mode, number of binders, twist); the twist (SSTWLC) kernels are their own
// translation units, built for one or two binders
static int launch_sim_once(const McSimArgs &a, int rng_mode) {
    const DevCtx &d = a.d;
    if (d.twist) {
        if (d.nb > 2) return -1000;
        return rng_mode == CHROMO_RNG_REPLAY ? cb_mc_sim_replay_tw_12(a) : cb_mc_sim_philox_tw_12(a);
    }
    if (rng_mode == CHROMO_RNG_REPLAY) return d.nb <= 2 ? cb_mc_sim_replay_12(a) : cb_mc_sim_replay_34(a);
    return d.nb <= 2 ? cb_mc_sim_philox_12(a) : cb_mc_sim_philox_34(a);
}
// One launch runs every MC step with the move types in all_moves' order (the kernel's loop).  A controller list
// in another order (mc_sim.pyx:92-103 walks the list as given) becomes one launch per (MC step, move type) with
// only that type enabled -- the replicas' RNG streams, controller state and densities carry over on the device,
// the counters of chromo_last_attempts / _algo_bytes accumulate.  (A run-time order inside the kernel's loop cost
// the common case 1.4-2.3 % on B200; the order is rare.)
static int launch_sim(const McSimArgs &a, int rng_mode) {
    bool canonical = true;
    for (int i = 0; i < CHROMO_NUM_MOVES; i++) canonical = canonical && (!a.order || a.order[i] == i);
    McSimArgs b = a;
    b.d.type_mask = (1 << CHROMO_NUM_MOVES) - 1;
    b.d.accumulate = 0;
    if (canonical || a.num_mc_steps == 0) return launch_sim_once(b, rng_mode);
    b.num_mc_steps = 1;
    for (long long k = 0; k < a.num_mc_steps; k++)
        for (int i = 0; i < CHROMO_NUM_MOVES; i++) {
            b.d.type_mask = 1 << a.order[i];
            const int e = launch_sim_once(b, rng_mode);
            if (e) return e;
            b.d.accumulate = 1;
        }
    return 0;
}
static int launch_step(const McStepArgs &a, int rng_mode) {
    const DevCtx &d = a.d;
    if (d.twist) {
        if (d.nb > 2) return -1000;
        return rng_mode == CHROMO_RNG_REPLAY ? cb_mc_step_replay_tw_12(a) : cb_mc_step_philox_tw_12(a);
    }
    if (rng_mode == CHROMO_RNG_REPLAY) return d.nb <= 2 ? cb_mc_step_replay_12(a) : cb_mc_step_replay_34(a);
    return d.nb <= 2 ? cb_mc_step_philox_12(a) : cb_mc_step_philox_34(a);
}
static const char *launch_error(int e) {
    return e == -1000 ? "twist (SSTWLC) kernels are built for one or two binders" : cudaGetErrorString((cudaError_t)e);
}

extern "C" const char *chromo_last_error(void) { return g_err.c_str(); }
extern "C" int chromo_version(void) { return 100; }

template <class T>
static int dev_alloc(chromo_ctx *c, T **p, size_t n) {
    void *q = nullptr;
    size_t b = n * sizeof(T);
    if (b == 0) b = sizeof(T);
    CK(cudaMalloc(&q, b));
    CK(cudaMemsetAsync(q, 0, b, c->stream));
    c->allocs.push_back(q);
    c->bytes += (int64_t)b;
    *p = (T *)q;
    return 0;
}

static int choose_table(chromo_ctx *c) {
    // One thread block per SM holds `rpb` replicas (mc_kernel.cuh, mc_sim_kernel); all replicas
    // are resident at once when R <= SMs x CB_MAX_RPB.  Pick the largest per-warp table (a multiple
    // of 32 slots, never below 128) that lets the block's replicas share the SM's shared memory;
    // at most 768 slots when replicas share a block (measured: beyond that the shared-memory
    // carve-out only takes L1 away from the bead rows), 2048 otherwise.
    const int ncol = c->d.ncol;
    int rpb = (c->d.R + c->sm_count - 1) / c->sm_count;
    rpb = std::max(1, std::min(rpb, CB_MAX_RPB));
    if (c->rpb_fixed > 0) rpb = c->rpb_fixed;
    const size_t per_replica = c->smem_optin / (size_t)rpb;
    int cap = rpb > 1 ? 768 : 2048;
    while (cap > 128 && cb_replica_smem(cap, ncol, c->warps) > per_replica) cap -= 32;
    // The shared-memory carve-out comes in steps (..., 164, 196, 228 KB of the SM's 256 KB; the rest is L1, which
    // holds the bead rows of the attempts in flight): a block that needs a byte more than 196 KB (its 1 KB of
    // static and the 1 KB the driver reserves included) leaves 28 KB of L1 instead of 60.  Measured at the
    // stationary working point: 704 slots inside the 196 KB step beat 736 / 768 slots outside it by 3 % (v17
    // layout); with the prepared tangent sets overlaid on Prop::M the step holds 832 slots (+0.8 % over 768).
    const size_t step = (size_t)196 * 1024 - 2048;
    int inside = rpb > 1 ? 1024 : cap;
    while (inside > 512 && ((size_t)rpb * cb_replica_smem(inside, ncol, c->warps) > step ||
                            cb_replica_smem(inside, ncol, c->warps) > per_replica))
        inside -= 32;
    if ((size_t)rpb * cb_replica_smem(inside, ncol, c->warps) <= step) cap = inside;
    c->cap = cap;
    c->rpb = rpb;
    return 0;
}

extern "C" int chromo_ctx_create(chromo_ctx **out, int device, const chromo_shape *s) {
    if (!out || !s) return fail(CHROMO_ERR_ARG, "null argument");
    if (s->n_replicas < 1 || s->num_beads < 2 || s->num_binders < 1 || s->num_binders > CHROMO_MAX_BINDERS)
        return fail(CHROMO_ERR_ARG, "bad shape: R=%lld N=%lld nb=%lld (nb must be 1..%d)",
                    (long long)s->n_replicas, (long long)s->num_beads, (long long)s->num_binders,
                    CHROMO_MAX_BINDERS);
    if (s->num_beads > 0x3fffffff) return fail(CHROMO_ERR_ARG, "N too large");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(CHROMO_ERR_CUDA, "no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    chromo_ctx *c = new chromo_ctx();
    c->device = device;
    c->shape = *s;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_sm = prop.sharedMemPerMultiprocessor;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    DevCtx &d = c->d;
    d.R = (int)s->n_replicas;
    d.N = (int)s->num_beads;
    d.nb = (int)s->num_binders;
    d.ncol = d.nb + 1;
    d.nx = (int)s->nx;
    d.ny = (int)s->ny;
    d.nz = (int)s->nz;
    long long nbins = (long long)s->nx * s->ny * s->nz;
    if (nbins > 0x7fffffffLL) return fail(CHROMO_ERR_ARG, "grid too large");
    d.n_bins = (int)nbins;
    d.field_active = nbins > 0;
    d.confine_type = s->confine_type;
    d.confine_length = s->confine_length;
    d.vf_limit = (double)s->vf_limit; // C float -> double, as the reference's comparisons do
    d.bead_vol = s->bead_vol;
    d.max_binders = s->max_binders;
    if (d.field_active) {
        int n3[3] = {d.nx, d.ny, d.nz};
        for (int j = 0; j < 3; j++) { // init_grid fields.pyx:536-575
            d.width[j] = s->width[j];
            d.dxyz[j] = s->width[j] / n3[j];
            d.half_width[j] = 0.5 * s->width[j];
            d.half_step[j] = 0.5 * d.dxyz[j];
            d.inv_width[j] = 1.0 / d.width[j];
            d.inv_dxyz[j] = 1.0 / d.dxyz[j];
        }
        d.vol_bin = s->width[0] * s->width[1] * s->width[2] / (double)d.n_bins;
        d.inv_vol_bin = 1.0 / d.vol_bin;
        // smallest hundredth K with double(K/100) > vf_limit (see round2_exceeds)
        double K = 0.0;
        while (K / 100.0 <= d.vf_limit && K < 1e7) K += 1.0;
        c->roundK = K;
    }
    size_t RN = (size_t)d.R * d.N;
    int rc;
    if ((rc = dev_alloc(c, &d.r, RN * 3))) return rc;
    if ((rc = dev_alloc(c, &d.t3, RN * 3))) return rc;
    if ((rc = dev_alloc(c, &d.t2, RN * 3))) return rc;
    if ((rc = dev_alloc(c, &d.states, RN * d.nb))) return rc;
    if ((rc = dev_alloc(c, &d.mods, RN * d.nb))) return rc;
    if ((rc = dev_alloc(c, &d.density, (size_t)d.R * (d.n_bins > 0 ? d.n_bins : 1) * d.ncol))) return rc;
    if ((rc = dev_alloc(c, &d.moves, (size_t)d.R * CHROMO_NUM_MOVES))) return rc;
    if ((rc = dev_alloc(c, &d.glibc, (size_t)d.R * CB_GLIBC_WORDS))) return rc;
    if ((rc = dev_alloc(c, &d.mt, (size_t)d.R * CB_MT_WORDS))) return rc;
    if ((rc = dev_alloc(c, &d.philox_ctr, (size_t)d.R))) return rc;
    if ((rc = dev_alloc(c, &c->d_bad, (size_t)1))) return rc;
    d.rep_offset = 0u;
    d.batch = 32;
    d.type_mask = (1 << CHROMO_NUM_MOVES) - 1;
    d.accumulate = 0;
    for (int i = 0; i < CHROMO_NUM_MOVES; i++) c->move_order[i] = i;
    if ((rc = dev_alloc(c, &d.tan_inds, RN))) return rc;
    if ((rc = dev_alloc(c, &d.sel_bits, (size_t)d.R * ((d.N + 31) / 32)))) return rc;
    if ((rc = dev_alloc(c, &d.st_new, RN))) return rc;
    if ((rc = dev_alloc(c, &d.attempts, (size_t)d.R))) return rc;
    if ((rc = dev_alloc(c, &d.algo_bytes, (size_t)d.R))) return rc;
    if ((rc = dev_alloc(c, &c->d_chi, (size_t)d.R))) return rc;
    if ((rc = dev_alloc(c, &c->d_mu, (size_t)d.R * d.nb))) return rc;
    d.chi = c->d_chi;
    d.mu = c->d_mu;
    // reduction scratch
    c->nblk_bins = d.n_bins > 0 ? std::min(64, (d.n_bins + FK_THREADS - 1) / FK_THREADS) : 1;
    c->nblk_bonds = std::min(64, (d.N - 1 + FK_THREADS - 1) / FK_THREADS);
    int nblk = std::max(c->nblk_bins, c->nblk_bonds);
    if ((rc = dev_alloc(c, &c->d_partial, (size_t)d.R * nblk * (d.ncol + 1)))) return rc;
    if ((rc = dev_alloc(c, &c->d_out, (size_t)d.R * (d.ncol + 1)))) return rc;
    if ((rc = dev_alloc(c, &c->d_dcount, (size_t)d.R * d.nb))) return rc;
    // staging for int64 <-> int8 conversion
    c->stage_elems = (int64_t)std::min<size_t>(RN * d.nb, (size_t)1 << 24);
    if ((rc = dev_alloc(c, &c->d_stage, (size_t)c->stage_elems))) return rc;
    // single-step instrumentation
    c->dbg_inds_cap = d.N;
    c->dbg_touched_cap = d.n_bins > 0 ? std::min<int64_t>(d.n_bins, 16LL * d.N) : 1;
    if ((rc = dev_alloc(c, &c->d_dbg, 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_dbg_inds, (size_t)c->dbg_inds_cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_dbg_rows, (size_t)c->dbg_inds_cap * (9 + d.nb)))) return rc;
    if ((rc = dev_alloc(c, &c->d_dbg_touched, (size_t)c->dbg_touched_cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_dbg_dtrial, (size_t)c->dbg_touched_cap * d.ncol))) return rc;
    // defaults: chi = 1, seeds as a fresh process (glibc seed 1, numpy seed 0)
    {
        std::vector<double> chi(d.R, 1.0);
        CK(cudaMemcpyAsync(c->d_chi, chi.data(), sizeof(double) * d.R, cudaMemcpyHostToDevice, c->stream));
        std::vector<uint32_t> g((size_t)d.R * CB_GLIBC_WORDS, 0), m((size_t)d.R * CB_MT_WORDS, 0);
        glibc_srand_host(g.data(), 1);
        mt_seed_host(m.data(), 0);
        for (int r = 1; r < d.R; r++) {
            memcpy(&g[(size_t)r * CB_GLIBC_WORDS], g.data(), sizeof(uint32_t) * CB_GLIBC_WORDS);
            memcpy(&m[(size_t)r * CB_MT_WORDS], m.data(), sizeof(uint32_t) * CB_MT_WORDS);
        }
        CK(cudaMemcpyAsync(d.glibc, g.data(), g.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d.mt, m.data(), m.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    // default: a second warp per replica while at most 4 replicas share an SM (measured on B200 at
    // 100,000 beads: 148 replicas 31 -> 43 M attempts/s, 296: 58 -> 78, 592: 111 -> 116; at 7 per SM one
    // warp is faster, 159 vs 122); chromo_ctx_set_warps_per_replica overrides
#ifndef CHROMO_HOST_EMU // (the CPU emulation of the test-suite keeps one warp: half the OS threads)
    c->warps = (long long)d.R <= 4LL * c->sm_count ? 2 : 1;
#endif
    choose_table(c);
    refresh_fx_base(c);
    *out = c;
    return CHROMO_OK;
}

extern "C" int chromo_ctx_destroy(chromo_ctx *c) {
    if (!c) return CHROMO_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (cudaStream_t st : c->chunk_streams) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    for (cudaEvent_t e : c->chunk_uploaded) cudaEventDestroy(e);
    for (long long *p : c->chunk_stage) cudaFree(p);
    for (void *p : c->allocs) cudaFree(p);
    cudaStreamDestroy(c->stream);
    delete c;
    return CHROMO_OK;
}
extern "C" int chromo_ctx_sync(chromo_ctx *c) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_table_capacity(chromo_ctx *c, int64_t cap, int64_t *cap_out) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (cap == 0) {
        choose_table(c);
    } else {
        if (cap < 128 || cap > 4096 || (cap % 32)) return fail(CHROMO_ERR_ARG, "capacity must be a multiple of 32 in [128, 4096]");
        if ((size_t)c->rpb * cb_replica_smem((int)cap, c->d.ncol, c->warps) > c->smem_optin)
            return fail(CHROMO_ERR_ARG, "capacity does not fit in shared memory");
        c->cap = (int)cap;
    }
    if (cap_out) *cap_out = c->cap;
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_warps_per_replica(chromo_ctx *c, int64_t warps, int64_t *warps_out) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (warps != 0) {
        if (warps < 1 || warps > CB_MAX_WARPS) return fail(CHROMO_ERR_ARG, "warps per replica must be in [1, %d]", CB_MAX_WARPS);
        c->warps = (int)warps;
        choose_table(c);
    }
    if (warps_out) *warps_out = c->warps;
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_replicas_per_block(chromo_ctx *c, int64_t rpb, int64_t *rpb_out) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (rpb != 0) {
        if (rpb < -1 || rpb > CB_MAX_RPB) return fail(CHROMO_ERR_ARG, "replicas per block must be in [1, %d] (-1 = automatic)", CB_MAX_RPB);
        c->rpb_fixed = rpb < 0 ? 0 : (int)rpb;
        choose_table(c);
    }
    if (rpb_out) *rpb_out = c->rpb;
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_replica_offset(chromo_ctx *c, int64_t offset) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (offset < 0 || offset > 0xffffffffLL - c->d.R) return fail(CHROMO_ERR_ARG, "replica offset out of range");
    c->d.rep_offset = (unsigned)offset;
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_move_order(chromo_ctx *c, const int32_t *order) {
    if (!c || !order) return fail(CHROMO_ERR_ARG, "null argument");
    unsigned seen = 0;
    for (int i = 0; i < CHROMO_NUM_MOVES; i++) {
        if (order[i] < 0 || order[i] >= CHROMO_NUM_MOVES || (seen >> order[i] & 1u))
            return fail(CHROMO_ERR_ARG, "move order must be a permutation of 0..%d", CHROMO_NUM_MOVES - 1);
        seen |= 1u << order[i];
    }
    for (int i = 0; i < CHROMO_NUM_MOVES; i++) c->move_order[i] = order[i];
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_batch_size(chromo_ctx *c, int64_t batch) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (batch < 1 || batch > 32) return fail(CHROMO_ERR_ARG, "batch size must be in [1, 32]");
    c->d.batch = (int)batch;
    return CHROMO_OK;
}
// page-lock caller-owned host arrays so that chromo_mc_sim_host / upload / download run at link speed
extern "C" int chromo_host_register(void *p, uint64_t bytes) {
    if (!p || bytes == 0) return fail(CHROMO_ERR_ARG, "null or empty host range");
#ifndef CHROMO_HOST_EMU
    cudaPointerAttributes at; // memory from cudaHostAlloc (torch pinned tensors) or an earlier registration
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered)
        return fail(CHROMO_ERR_STATE, "host range is already page-locked");
    (void)cudaGetLastError();
#endif
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return fail(CHROMO_ERR_STATE, "host range is already page-locked");
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(CHROMO_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
    }
    return CHROMO_OK;
}
extern "C" int chromo_host_unregister(void *p) {
    if (!p) return fail(CHROMO_ERR_ARG, "null host pointer");
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(CHROMO_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    }
    return CHROMO_OK;
}
extern "C" int chromo_ctx_set_fast_field(chromo_ctx *c, int64_t n_points) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (n_points < 0 || n_points > 1000000) return fail(CHROMO_ERR_ARG, "n_points out of range");
    if (n_points > 0 && !c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    n_points += n_points % 2; // fields.pyx:603
    c->d.fast_n = (int)n_points;
    for (int j = 0; j < 3; j++) c->d.fast_sbw[j] = n_points ? c->d.dxyz[j] / (double)n_points : 0.0;
    return CHROMO_OK;
}
extern "C" int chromo_get_rng_counters(chromo_ctx *c, int64_t first, int64_t n, uint64_t *counters) {
    if (!c || !counters) return fail(CHROMO_ERR_ARG, "null argument");
    if (first < 0 || n < 0 || first + n > c->d.R) return fail(CHROMO_ERR_ARG, "replica range out of bounds");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(counters, c->d.philox_ctr + first, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" int chromo_set_rng_counters(chromo_ctx *c, int64_t first, int64_t n, const uint64_t *counters) {
    if (!c || !counters) return fail(CHROMO_ERR_ARG, "null argument");
    if (first < 0 || n < 0 || first + n > c->d.R) return fail(CHROMO_ERR_ARG, "replica range out of bounds");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->d.philox_ctr + first, counters, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" void *chromo_ctx_stream(chromo_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t chromo_ctx_bytes(chromo_ctx *c) { return c ? c->bytes : 0; }

// ------------------------------------------------------------- parameters
template <class T>
static int replace_buf(chromo_ctx *c, T **slot, const T *host, size_t n) {
    if (!*slot) {
        int rc = dev_alloc(c, slot, n);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(*slot, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int chromo_set_binders(chromo_ctx *c, const int64_t *sites, const double *pref,
                                  const double *e_intra, const double *xpref, const double *bind_F,
                                  int64_t S) {
    if (!c || !sites || !pref || !e_intra || !xpref || !bind_F) return fail(CHROMO_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    for (int a = 0; a < d.nb; a++) {
        if (sites[a] < 0 || sites[a] > S || S > 126) return fail(CHROMO_ERR_ARG, "bad sites_per_bead");
        d.sites[a] = (int)sites[a];
        d.pref[a] = pref[a];
        d.e_intra[a] = e_intra[a];
        for (int b = 0; b < d.nb; b++) d.xpref[a * d.nb + b] = xpref[a * d.nb + b];
    }
    d.any_cross = 0;
    for (int a = 0; a < d.nb * d.nb; a++) d.any_cross |= (d.xpref[a] != 0.0);
    d.S1 = (int)S + 1;
    if (c->d_bindF) return fail(CHROMO_ERR_STATE, "binders already set for this context");
    int rc = replace_buf(c, &c->d_bindF, bind_F, (size_t)d.nb * d.S1 * d.S1);
    if (rc) return rc;
    d.bindF = c->d_bindF;
    c->have_binders = true;
    refresh_fx_base(c);
    return CHROMO_OK;
}

extern "C" int chromo_set_replica_params(chromo_ctx *c, const double *chi, const double *mu) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    if (chi) CK(cudaMemcpyAsync(c->d_chi, chi, sizeof(double) * c->d.R, cudaMemcpyHostToDevice, c->stream));
    if (mu) CK(cudaMemcpyAsync(c->d_mu, mu, sizeof(double) * c->d.R * c->d.nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

extern "C" int chromo_set_bond_params(chromo_ctx *c, int64_t n_sets, const double *eps_bend,
                                      const double *eps_par, const double *eps_perp,
                                      const double *gamma, const double *eta) {
    if (!c || !eps_bend || !eps_par || !eps_perp || !gamma || !eta) return fail(CHROMO_ERR_ARG, "null argument");
    if (n_sets != 1 && n_sets != c->d.R) return fail(CHROMO_ERR_ARG, "n_sets must be 1 or R");
    if (c->d_bond) return fail(CHROMO_ERR_STATE, "bond parameters already set for this context");
    CK(cudaSetDevice(c->device));
    size_t nbonds = (size_t)c->d.N - 1;
    std::vector<double> packed((size_t)n_sets * nbonds * 5);
    for (size_t s = 0; s < (size_t)n_sets; s++)
        for (size_t b = 0; b < nbonds; b++) {
            double *o = &packed[(s * nbonds + b) * 5];
            o[0] = eps_bend[s * nbonds + b];
            o[1] = eps_par[s * nbonds + b];
            o[2] = eps_perp[s * nbonds + b];
            o[3] = gamma[s * nbonds + b];
            o[4] = eta[s * nbonds + b];
        }
    int rc = replace_buf(c, &c->d_bond, packed.data(), packed.size());
    if (rc) return rc;
    c->d.bond = c->d_bond;
    c->d.bond_stride = (n_sets == 1) ? 0 : (long long)nbonds * 5;
    c->have_bonds = true;
    return CHROMO_OK;
}

extern "C" int chromo_set_twist_params(chromo_ctx *c, int64_t n_sets, const double *eps_twist,
                                       const double *natural_twist) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    if (!eps_twist && !natural_twist) { // back to a chain without twist
        c->d.twist = nullptr;
        c->d.twist_stride = 0;
        return CHROMO_OK;
    }
    if (!eps_twist || !natural_twist) return fail(CHROMO_ERR_ARG, "null argument");
    if (n_sets != 1 && n_sets != c->d.R) return fail(CHROMO_ERR_ARG, "n_sets must be 1 or n_replicas");
    if (c->d.nb > 2) return fail(CHROMO_ERR_ARG, "twist (SSTWLC) kernels are built for one or two binders");
    size_t nbonds = (size_t)c->d.N - 1;
    std::vector<double> packed((size_t)n_sets * nbonds * 2);
    for (size_t i = 0; i < (size_t)n_sets * nbonds; i++) {
        if (!(eps_twist[i] > 0.0)) return fail(CHROMO_ERR_ARG, "Twist modulus must be positive."); // polymers.pyx:2095
        packed[2 * i] = eps_twist[i];
        packed[2 * i + 1] = natural_twist[i];
    }
    int rc = replace_buf(c, &c->d_twist, packed.data(), packed.size());
    if (rc) return rc;
    c->d.twist = c->d_twist;
    c->d.twist_stride = (n_sets == 1) ? 0 : (long long)nbonds * 2;
    return CHROMO_OK;
}

extern "C" int chromo_set_detailed_nucleosomes(chromo_ctx *c, const double *consts20) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    if (!consts20) {
        c->d.detailed = nullptr;
        return CHROMO_OK;
    }
    if (!c->d.twist) return fail(CHROMO_ERR_STATE, "DetailedChromatin is an SSTWLC: call chromo_set_twist_params first");
    int rc = replace_buf(c, &c->d_detailed, consts20, (size_t)20);
    if (rc) return rc;
    c->d.detailed = c->d_detailed;
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

extern "C" int chromo_set_access_volumes(chromo_ctx *c, const double *access_vol) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    if (!access_vol) {
        c->d.access_vol = nullptr;
        refresh_fx_base(c);
        return CHROMO_OK;
    }
    if (!c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    int rc = replace_buf(c, &c->d_access, access_vol, (size_t)c->d.n_bins);
    if (rc) return rc;
    c->d.access_vol = c->d_access;
    c->min_access_vol = 0.0;
    for (int i = 0; i < c->d.n_bins; i++)
        if (access_vol[i] > 0.0 && (c->min_access_vol == 0.0 || access_vol[i] < c->min_access_vol))
            c->min_access_vol = access_vol[i];
    refresh_fx_base(c);
    return CHROMO_OK;
}

// ------------------------------------------------------------------ state
static int check_range(chromo_ctx *c, int64_t first, int64_t n) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (first < 0 || n < 0 || first + n > c->d.R)
        return fail(CHROMO_ERR_ARG, "replica range [%lld, %lld) outside [0, %d)", (long long)first,
                    (long long)(first + n), c->d.R);
    return 0;
}

// largest value a state / mark may take: the binding free-energy table is [nb][S1][S1]
static int value_limit(const chromo_ctx *c) {
    int hi = 0;
    for (int a = 0; a < c->d.nb; a++) hi = std::max(hi, c->d.sites[a]);
    return c->have_binders ? hi : 127;
}
static int check_bad_values(chromo_ctx *c) {
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (bad) {
        CK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
        return fail(CHROMO_ERR_ARG, "states / chemical_mods must lie in [0, %d] (the largest sites_per_bead)", value_limit(c));
    }
    return 0;
}
static int upload_i64_as_i8(chromo_ctx *c, signed char *dst, const int64_t *src, size_t n) {
    for (size_t o = 0; o < n; o += (size_t)c->stage_elems) {
        size_t m = std::min(n - o, (size_t)c->stage_elems);
        CK(cudaMemcpyAsync(c->d_stage, src + o, m * 8, cudaMemcpyHostToDevice, c->stream));
        CB_LAUNCH(narrow_i64_kernel, (unsigned)((m + 255) / 256), 256, 0, c->stream, (const long long *)c->d_stage, dst + o, (long long)m,
                  value_limit(c), c->d_bad);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
    }
    return check_bad_values(c);
}
static int download_i8_as_i64(chromo_ctx *c, int64_t *dst, const signed char *src, size_t n) {
    for (size_t o = 0; o < n; o += (size_t)c->stage_elems) {
        size_t m = std::min(n - o, (size_t)c->stage_elems);
        CB_LAUNCH(widen_i8_kernel, (unsigned)((m + 255) / 256), 256, 0, c->stream, src + o, c->d_stage, (long long)m);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(dst + o, c->d_stage, m * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

extern "C" int chromo_upload_state(chromo_ctx *c, int64_t first, int64_t n, const double *r,
                                   const double *t3, const double *t2, const int64_t *states,
                                   const int64_t *mods) {
    int rc = check_range(c, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    size_t off = (size_t)first * d.N, cnt = (size_t)n * d.N;
    if (r) CK(cudaMemcpyAsync(d.r + off * 3, r, cnt * 24, cudaMemcpyHostToDevice, c->stream));
    if (t3) CK(cudaMemcpyAsync(d.t3 + off * 3, t3, cnt * 24, cudaMemcpyHostToDevice, c->stream));
    if (t2) CK(cudaMemcpyAsync(d.t2 + off * 3, t2, cnt * 24, cudaMemcpyHostToDevice, c->stream));
    if (states && (rc = upload_i64_as_i8(c, d.states + off * d.nb, states, cnt * d.nb))) return rc;
    if (mods && (rc = upload_i64_as_i8(c, d.mods + off * d.nb, mods, cnt * d.nb))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->have_state = true;
    if (mods && first == 0 && n == d.R) c->have_mods = true;
    return CHROMO_OK;
}

extern "C" int chromo_download_state(chromo_ctx *c, int64_t first, int64_t n, double *r, double *t3,
                                     double *t2, int64_t *states) {
    int rc = check_range(c, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    size_t off = (size_t)first * d.N, cnt = (size_t)n * d.N;
    if (r) CK(cudaMemcpyAsync(r, d.r + off * 3, cnt * 24, cudaMemcpyDeviceToHost, c->stream));
    if (t3) CK(cudaMemcpyAsync(t3, d.t3 + off * 3, cnt * 24, cudaMemcpyDeviceToHost, c->stream));
    if (t2) CK(cudaMemcpyAsync(t2, d.t2 + off * 3, cnt * 24, cudaMemcpyDeviceToHost, c->stream));
    if (states && (rc = download_i8_as_i64(c, states, d.states + off * d.nb, cnt * d.nb))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

extern "C" int chromo_download_density(chromo_ctx *c, int64_t first, int64_t n, double *density) {
    int rc = check_range(c, first, n);
    if (rc) return rc;
    if (!density) return fail(CHROMO_ERR_ARG, "null argument");
    if (!c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    CK(cudaSetDevice(c->device));
    size_t per = (size_t)c->d.n_bins * c->d.ncol;
    CK(cudaMemcpyAsync(density, c->d.density + first * per, n * per * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" int chromo_upload_density(chromo_ctx *c, int64_t first, int64_t n, const double *density) {
    int rc = check_range(c, first, n);
    if (rc) return rc;
    if (!density) return fail(CHROMO_ERR_ARG, "null argument");
    if (!c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    CK(cudaSetDevice(c->device));
    size_t per = (size_t)c->d.n_bins * c->d.ncol;
    CK(cudaMemcpyAsync(c->d.density + first * per, density, n * per * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

// -------------------------------------------------------- full recompute
#define DISPATCH_NB(nb, ...)                              \
    switch (nb) {                                         \
    case 1: { constexpr int NB = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NB = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NB = 3; __VA_ARGS__; } break; \
    default: { constexpr int NB = 4; __VA_ARGS__; } break; \
    }

static int launch_recompute(chromo_ctx *c, int clamp) {
    DevCtx &d = c->d;
    if (!d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    if (!c->have_state) return fail(CHROMO_ERR_STATE, "upload the polymer state first");
    size_t n = (size_t)d.R * d.n_bins * d.ncol;
    const size_t col_bytes = (size_t)d.n_bins * 12; // three 32-bit words per voxel and column
    if (col_bytes <= c->smem_optin && !c->force_l2_density) {
        // the grid fits one block's shared memory: privatised, bit-reproducible accumulation (field_kernels.cuh).
        // A term w / V (x state) is an integer of 3p bits spread over three 32-bit words that collect p bits each;
        // the 32 - p bits above are head room for the adds of one chunk of beads (a voxel receives at most one
        // term per bead and column, two / four / eight on a grid with one / two / three dimensions of a single
        // voxel), after which the carries are folded.  p = 21 (63-bit terms) for chains of up to 2,048 beads,
        // 18 (54-bit terms, quantum ~5e-21 nm^-3 at C2: 1/200 of the 1e-18 below which the reference itself
        // zeroes a density) without a fold up to 16,384, folds every 16,384 beads beyond.
        int kdeg = 1;
        for (int n1 : {d.nx, d.ny, d.nz}) kdeg *= (n1 == 1 ? 2 : 1);
        int hbits = 11;
        while (hbits < 14 && (1LL << hbits) < (long long)d.N * kdeg) hbits++;
        const int pbits = 32 - hbits, fold_beads = std::max(1, (1 << hbits) / kdeg);
        double vmin = (d.access_vol && c->min_access_vol > 0.0) ? c->min_access_vol : d.vol_bin;
        int smax = 1;
        for (int a = 0; a < d.nb; a++) smax = std::max(smax, d.sites[a]);
        int sbits = 0;
        while ((1 << sbits) < smax) sbits++;
        // (w / V) * state <= 2^(sbits - ilogb(vmin)) must stay below 2^(3p - 1); the top word is never folded, so
        // the terms of ALL beads must fit its 32 bits: chains beyond 2^(33 - p) beads give up a bit per doubling
        int tot_bits = 0;
        while ((1LL << tot_bits) < (long long)d.N * kdeg) tot_bits++;
        const int extra = std::max(0, pbits - 1 + tot_bits - 32);
        const int fx_e = 3 * pbits - 1 - extra - sbits + (int)ilogb(vmin);
        const int cpb = (int)std::min<size_t>((size_t)d.ncol, c->smem_optin / col_bytes); // columns per block
        const size_t priv_bytes = col_bytes * (size_t)cpb;
        dim3 grid(d.R, (d.ncol + cpb - 1) / cpb);
        int lrc = 0;
        DISPATCH_NB(d.nb, {
            auto k = density_private_kernel<NB>;
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)priv_bytes);
            if (e != cudaSuccess) lrc = (int)e;
            else CB_LAUNCH(k, grid, FK_PRIV_THREADS, priv_bytes, c->stream, d, clamp, fx_e, pbits, fold_beads, cpb);
        });
        if (lrc) return fail(CHROMO_ERR_CUDA, "density_private_kernel: %s", cudaGetErrorString((cudaError_t)lrc));
        CK(cudaGetLastError());
        return 0;
    }
    CK(cudaMemsetAsync(d.density, 0, n * 8, c->stream));
    dim3 grid((d.N + FK_THREADS - 1) / FK_THREADS, d.R);
    DISPATCH_NB(d.nb, { auto k = density_scatter_kernel<NB>; CB_LAUNCH(k, grid, FK_THREADS, 0, c->stream, d); });
    CK(cudaGetLastError());
    if (clamp) {
        CB_LAUNCH(density_clamp_kernel, (unsigned)((n + 255) / 256), 256, 0, c->stream, d.density, (long long)n);
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int chromo_field_recompute(chromo_ctx *c, int clamp) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    return launch_recompute(c, clamp);
}

static int field_reduce(chromo_ctx *c, int chi_observable) {
    DevCtx &d = c->d;
    CK(cudaMemsetAsync(c->d_dcount, 0, sizeof(int) * d.R * d.nb, c->stream));
    dim3 grid(c->nblk_bins, d.R);
    DISPATCH_NB(d.nb, {
        auto k = field_energy_kernel<NB>;
        CB_LAUNCH(k, grid, FK_THREADS, 0, c->stream, d, c->roundK, c->d_partial, c->d_dcount, chi_observable);
    });
    CK(cudaGetLastError());
    int tot = d.R * d.ncol;
    CB_LAUNCH(partial_finish_kernel, (tot + 127) / 128, 128, 0, c->stream, (const double *)c->d_partial, c->d_out, d.R, c->nblk_bins, d.ncol);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int chromo_field_energy(chromo_ctx *c, double *E, double *sum_sq, int64_t *doubly,
                                   double *nonspecific) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (!c->have_binders) return fail(CHROMO_ERR_STATE, "call chromo_set_binders first");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    int rc = launch_recompute(c, 0); // compute_E recomputes the densities (fields.pyx:1962-1964)
    if (rc) return rc;
    if ((rc = field_reduce(c, 0))) return rc;
    std::vector<double> out((size_t)d.R * d.ncol);
    std::vector<int> dc((size_t)d.R * d.nb);
    CK(cudaMemcpyAsync(out.data(), c->d_out, out.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(dc.data(), c->d_dcount, dc.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < d.R; r++) {
        double e = 0.0; // get_E_binders_and_beads fields.pyx:2236-2253
        for (int a = 0; a < d.nb; a++) {
            e += d.pref[a] * out[(size_t)r * d.ncol + a];
            e += d.e_intra[a] * (double)dc[(size_t)r * d.nb + a];
            if (sum_sq) sum_sq[(size_t)r * d.nb + a] = out[(size_t)r * d.ncol + a];
            if (doubly) doubly[(size_t)r * d.nb + a] = dc[(size_t)r * d.nb + a];
        }
        e += out[(size_t)r * d.ncol + d.nb];
        if (nonspecific) nonspecific[r] = out[(size_t)r * d.ncol + d.nb];
        if (E) E[r] = e;
    }
    return CHROMO_OK;
}

extern "C" int chromo_chi_observable(chromo_ctx *c, double *Phi) {
    if (!c || !Phi) return fail(CHROMO_ERR_ARG, "null argument");
    if (!c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    int rc = field_reduce(c, 1);
    if (rc) return rc;
    std::vector<double> out((size_t)d.R * d.ncol);
    CK(cudaMemcpyAsync(out.data(), c->d_out, out.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < d.R; r++) Phi[r] = out[(size_t)r * d.ncol + d.nb];
    return CHROMO_OK;
}

// ---------------------------------------------------------- replica exchange
extern "C" int chromo_exchange_init(chromo_ctx *c, const double *chi_by_replica, int64_t n_total, int64_t first,
                                    int64_t ladder_len) {
    if (!c || !chi_by_replica) return fail(CHROMO_ERR_ARG, "null argument");
    DevCtx &d = c->d;
    if (n_total < 1 || first < 0 || first + d.R > n_total) return fail(CHROMO_ERR_ARG, "this context's replicas [first, first + R) must lie in [0, n_total)");
    if (ladder_len <= 0) ladder_len = n_total;
    if (n_total % ladder_len) return fail(CHROMO_ERR_ARG, "n_total must be a multiple of ladder_len");
    if (n_total > 0x7fffffffLL) return fail(CHROMO_ERR_ARG, "too many replicas");
    CK(cudaSetDevice(c->device));
    // rung k of ladder l = the k-th smallest chi among replicas [l * ladder_len, (l + 1) * ladder_len)
    std::vector<double> ladder((size_t)n_total);
    std::vector<int> rung((size_t)n_total);
    for (int64_t l = 0; l < n_total / ladder_len; l++) {
        std::vector<int> idx((size_t)ladder_len);
        for (int64_t i = 0; i < ladder_len; i++) idx[(size_t)i] = (int)(l * ladder_len + i);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return chi_by_replica[x] < chi_by_replica[y]; });
        for (int64_t i = 0; i < ladder_len; i++) {
            rung[(size_t)(l * ladder_len + i)] = idx[(size_t)i];
            ladder[(size_t)(l * ladder_len + i)] = chi_by_replica[idx[(size_t)i]];
        }
    }
    int rc;
    if (c->ex_total != n_total) {
        if ((rc = dev_alloc(c, &c->d_ex_ladder, (size_t)n_total))) return rc;
        if ((rc = dev_alloc(c, &c->d_ex_rung, (size_t)n_total))) return rc;
        if (!c->d_ex_counters && (rc = dev_alloc(c, &c->d_ex_counters, (size_t)2))) return rc;
    }
    c->ex_total = n_total;
    c->ex_first = first;
    c->ex_ladder_len = ladder_len;
    CK(cudaMemcpyAsync(c->d_ex_ladder, ladder.data(), (size_t)n_total * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_ex_rung, rung.data(), (size_t)n_total * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->d_ex_counters, 0, 16, c->stream));
    if (!c->d_chi && (rc = dev_alloc(c, &c->d_chi, (size_t)d.R))) return rc;
    CK(cudaMemcpyAsync(c->d_chi, chi_by_replica + first, (size_t)d.R * 8, cudaMemcpyHostToDevice, c->stream));
    d.chi = c->d_chi;
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

extern "C" int chromo_exchange_observable(chromo_ctx *c, double *phi_dev) {
    if (!c || !phi_dev) return fail(CHROMO_ERR_ARG, "null argument");
    if (!c->d.field_active) return fail(CHROMO_ERR_STATE, "context has no field");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    int rc = field_reduce(c, 1);
    if (rc) return rc;
    CB_LAUNCH(pick_column_kernel, (d.R + 127) / 128, 128, 0, c->stream, (const double *)c->d_out, phi_dev, d.R, d.ncol, d.nb);
    CK(cudaGetLastError());
    return CHROMO_OK; // asynchronous: ordered on the context's stream
}

extern "C" int chromo_exchange_step(chromo_ctx *c, const double *phi_all_dev, int64_t round, uint64_t seed) {
    if (!c || !phi_all_dev) return fail(CHROMO_ERR_ARG, "null argument");
    if (c->ex_total <= 0) return fail(CHROMO_ERR_STATE, "call chromo_exchange_init first");
    if (round < 0) return fail(CHROMO_ERR_ARG, "negative round");
    CK(cudaSetDevice(c->device));
    const long long pairs = (c->ex_total + 1) / 2;
    CB_LAUNCH(exchange_kernel, (unsigned)((pairs + 127) / 128), 128, 0, c->stream, phi_all_dev,
              (const double *)c->d_ex_ladder, c->d_ex_rung, c->d_chi, (long long)c->ex_total, (long long)c->ex_ladder_len,
              (long long)c->ex_first, c->d.R, (long long)round, (unsigned long long)seed, c->d_ex_counters);
    CK(cudaGetLastError());
    return CHROMO_OK; // asynchronous: the next mc_sim on this context's stream sees the new chi
}

extern "C" int chromo_exchange_state(chromo_ctx *c, int32_t *rung_replica, double *chi_local, uint64_t *pairs_tried,
                                     uint64_t *swaps_accepted) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (c->ex_total <= 0) return fail(CHROMO_ERR_STATE, "call chromo_exchange_init first");
    CK(cudaSetDevice(c->device));
    unsigned long long cnt[2] = {0, 0};
    if (rung_replica) CK(cudaMemcpyAsync(rung_replica, c->d_ex_rung, (size_t)c->ex_total * 4, cudaMemcpyDeviceToHost, c->stream));
    if (chi_local) CK(cudaMemcpyAsync(chi_local, c->d_chi, (size_t)c->d.R * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt, c->d_ex_counters, 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (pairs_tried) *pairs_tried = cnt[0];
    if (swaps_accepted) *swaps_accepted = cnt[1];
    return CHROMO_OK;
}

extern "C" int chromo_elastic_energy(chromo_ctx *c, double *E) {
    if (!c || !E) return fail(CHROMO_ERR_ARG, "null argument");
    if (!c->have_bonds || !c->have_state) return fail(CHROMO_ERR_STATE, "set bond parameters and state first");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    dim3 grid(c->nblk_bonds, d.R);
    CB_LAUNCH(elastic_energy_kernel, grid, FK_THREADS, 0, c->stream, d, c->d_partial);
    CK(cudaGetLastError());
    CB_LAUNCH(partial_finish_kernel, (d.R + 127) / 128, 128, 0, c->stream, (const double *)c->d_partial, c->d_out, d.R, c->nblk_bonds, 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(E, c->d_out, sizeof(double) * d.R, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

// -------------------------------------------------------------------- RNG
extern "C" int chromo_srand(chromo_ctx *c, const uint32_t *seeds) {
    if (!c || !seeds) return fail(CHROMO_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    std::vector<uint32_t> g((size_t)c->d.R * CB_GLIBC_WORDS, 0);
    for (int r = 0; r < c->d.R; r++) glibc_srand_host(&g[(size_t)r * CB_GLIBC_WORDS], seeds[r]);
    CK(cudaMemcpyAsync(c->d.glibc, g.data(), g.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" int chromo_numpy_seed(chromo_ctx *c, const uint32_t *seeds) {
    if (!c || !seeds) return fail(CHROMO_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    std::vector<uint32_t> m((size_t)c->d.R * CB_MT_WORDS, 0);
    for (int r = 0; r < c->d.R; r++) mt_seed_host(&m[(size_t)r * CB_MT_WORDS], seeds[r]);
    CK(cudaMemcpyAsync(c->d.mt, m.data(), m.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

// ----------------------------------------------------------- the hot path
static int check_ready(chromo_ctx *c) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (!c->have_binders) return fail(CHROMO_ERR_STATE, "call chromo_set_binders first");
    if (!c->have_bonds) return fail(CHROMO_ERR_STATE, "call chromo_set_bond_params first");
    if (!c->have_state) return fail(CHROMO_ERR_STATE, "call chromo_upload_state first");
    return 0;
}

static int validate_moves(chromo_ctx *c, const chromo_move_state *mv) {
    const int N = c->d.N;
    for (int r = 0; r < c->d.R; r++)
        for (int m = 0; m < CHROMO_NUM_MOVES; m++) {
            const chromo_move_state &s = mv[(size_t)r * CHROMO_NUM_MOVES + m];
            if (!s.move_on) continue;
            // bead_selection.pyx:85-88,138-141: a window larger than the chain raises
            if (s.amp_bead > N || s.bead_amp_hi > N)
                return fail(CHROMO_ERR_ARG, "Bead selection window size must be less than polymer length"
                                            " (replica %d move %d: amp_bead %d, N %d)", r, m, s.amp_bead, N);
            if (m == CHROMO_TANGENT_ROTATION && s.amp_bead < 1)
                return fail(CHROMO_ERR_ARG, "tangent_rotation needs amp_bead >= 1");
            if (m == CHROMO_END_PIVOT && s.amp_bead < 1)
                return fail(CHROMO_ERR_ARG, "end_pivot needs amp_bead >= 1");
            if (s.num_per_cycle < 0) return fail(CHROMO_ERR_ARG, "negative num_per_cycle");
        }
    return 0;
}

extern "C" int chromo_set_moves(chromo_ctx *c, const chromo_move_state *mv) {
    if (!c || !mv) return fail(CHROMO_ERR_ARG, "null argument");
    int rc = validate_moves(c, mv);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->d.moves, mv, sizeof(chromo_move_state) * c->d.R * CHROMO_NUM_MOVES,
                       cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
extern "C" int chromo_get_moves(chromo_ctx *c, chromo_move_state *mv) {
    if (!c || !mv) return fail(CHROMO_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(mv, c->d.moves, sizeof(chromo_move_state) * c->d.R * CHROMO_NUM_MOVES,
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}

extern "C" int chromo_mc_sim(chromo_ctx *c, int64_t num_mc_steps, chromo_move_state *moves,
                             double mu_adjust_factor, uint64_t seed, int rng_mode,
                             const uint32_t *numpy_seeds) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (num_mc_steps < 0) return fail(CHROMO_ERR_ARG, "negative num_mc_steps");
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    if (moves && (rc = chromo_set_moves(c, moves))) return rc;
    if (rng_mode == CHROMO_RNG_REPLAY && numpy_seeds && (rc = chromo_numpy_seed(c, numpy_seeds))) return rc;
    McSimArgs a{d, (long long)num_mc_steps, mu_adjust_factor, (unsigned long long)seed, c->cap, c->warps, c->rpb, c->stream};
    a.order = c->move_order;
    if (rng_mode != CHROMO_RNG_REPLAY && rng_mode != CHROMO_RNG_PHILOX) return fail(CHROMO_ERR_ARG, "unknown rng_mode %d", rng_mode);
    const int e = launch_sim(a, rng_mode);
    if (e) return fail(CHROMO_ERR_CUDA, "mc_sim launch failed: %s", launch_error(e));
    CK(cudaGetLastError());
    if (moves) {
        CK(cudaMemcpyAsync(moves, d.moves, sizeof(chromo_move_state) * d.R * CHROMO_NUM_MOVES,
                           cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return CHROMO_OK;
}

// mc_sim on host arrays, pipelined over replica chunks (see include/chromo_b200.h)
extern "C" int chromo_mc_sim_host(chromo_ctx *c, int64_t num_mc_steps, chromo_move_state *moves,
                                  double mu_adjust_factor, uint64_t seed, int rng_mode,
                                  const uint32_t *numpy_seeds, double *r, double *t3, double *t2,
                                  int64_t *states, const int64_t *mods, int64_t n_chunks) {
    if (!c) return fail(CHROMO_ERR_ARG, "null context");
    if (!r || !t3 || !t2 || !states) return fail(CHROMO_ERR_ARG, "null host array");
    if (!mods && !c->have_mods) return fail(CHROMO_ERR_ARG, "chemical_mods were never uploaded to this context");
    if (num_mc_steps < 0) return fail(CHROMO_ERR_ARG, "negative num_mc_steps");
    if (rng_mode != CHROMO_RNG_REPLAY && rng_mode != CHROMO_RNG_PHILOX) return fail(CHROMO_ERR_ARG, "unknown rng_mode %d", rng_mode);
    if (n_chunks < 0 || n_chunks > 64) return fail(CHROMO_ERR_ARG, "n_chunks must be in [0, 64]");
    c->have_state = true; // the state arrives with this call
    int rc = check_ready(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    DevCtx &d = c->d;
    if (moves && (rc = chromo_set_moves(c, moves))) return rc;
    if (rng_mode == CHROMO_RNG_REPLAY && numpy_seeds && (rc = chromo_numpy_seed(c, numpy_seeds))) return rc;
    CK(cudaStreamSynchronize(c->stream)); // everything queued on the context's own stream comes first
    // chunks are whole thread blocks; automatic = as many (<= 8) as keep every chunk's blocks resident at once
    const int rpb = c->rpb, nblk = (d.R + rpb - 1) / rpb;
    int chunks = (int)n_chunks;
    if (chunks == 0) {
        chunks = 8; // the call ends with the LAST chunk's kernel + download: the finer, the less is left exposed
        while (chunks > 1 && chunks * ((nblk + chunks - 1) / chunks) > c->sm_count) chunks--;
    }
    chunks = std::max(1, std::min(chunks, nblk));
    const int blk_per_chunk = (nblk + chunks - 1) / chunks;
    const int64_t rep_per_chunk = (int64_t)blk_per_chunk * rpb;
    const int64_t stage_need = rep_per_chunk * d.N * d.nb;
    if (stage_need > c->chunk_stage_elems) { // (re)size the staging buffers
        for (long long *p : c->chunk_stage) cudaFree(p);
        c->chunk_stage.clear();
        c->chunk_stage_elems = stage_need;
    }
    while ((int)c->chunk_streams.size() < chunks) {
        cudaStream_t st;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        c->chunk_streams.push_back(st);
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_uploaded.push_back(e);
    }
    while ((int)c->chunk_stage.size() < chunks) {
        void *q = nullptr;
        CK(cudaMalloc(&q, (size_t)c->chunk_stage_elems * 8));
        c->chunk_stage.push_back((long long *)q);
    }
#ifndef CHROMO_HOST_EMU
    // development aid (CHROMO_TIMELINE=1): per chunk, ms from the start of the call to the end of its
    // upload, its kernel and its download, printed to stderr
    const bool timeline = getenv("CHROMO_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> evs;
    if (timeline) {
        evs.resize(1 + 3 * (size_t)chunks);
        for (auto &e : evs) cudaEventCreate(&e);
        cudaEventRecord(evs[0], c->chunk_streams[0]);
    }
#define CB_MARK(i) do { if (timeline) cudaEventRecord(evs[1 + 3 * k + (i)], st); } while (0)
#else
#define CB_MARK(i) do { } while (0)
#endif
    for (int k = 0; k < chunks; k++) {
        const int64_t first = (int64_t)k * rep_per_chunk, n = std::min<int64_t>(rep_per_chunk, d.R - first);
        if (n <= 0) break;
        cudaStream_t st = c->chunk_streams[k];
        long long *stage = c->chunk_stage[k];
        const size_t off = (size_t)first * d.N, cnt = (size_t)n * d.N, cs = cnt * d.nb;
        // uploads go over the link one chunk at a time, in order: left to itself the copy engine
        // time-slices the streams and every chunk's data arrives at the end (measured on B200)
        if (k > 0) CK(cudaStreamWaitEvent(st, c->chunk_uploaded[k - 1], 0));
        CK(cudaMemcpyAsync(d.r + off * 3, r + off * 3, cnt * 24, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.t3 + off * 3, t3 + off * 3, cnt * 24, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.t2 + off * 3, t2 + off * 3, cnt * 24, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(stage, states + off * d.nb, cs * 8, cudaMemcpyHostToDevice, st));
        CB_LAUNCH(narrow_i64_kernel, (unsigned)((cs + 255) / 256), 256, 0, st, (const long long *)stage, d.states + off * d.nb, (long long)cs,
                  value_limit(c), c->d_bad);
        if (mods) { // NULL: mc_sim never modifies the marks, the copy already on the device is current
            CK(cudaMemcpyAsync(stage, mods + off * d.nb, cs * 8, cudaMemcpyHostToDevice, st));
            CB_LAUNCH(narrow_i64_kernel, (unsigned)((cs + 255) / 256), 256, 0, st, (const long long *)stage, d.mods + off * d.nb, (long long)cs,
                      value_limit(c), c->d_bad);
        }
        CK(cudaEventRecord(c->chunk_uploaded[k], st));
        CB_MARK(0);
        McSimArgs a{d, (long long)num_mc_steps, mu_adjust_factor, (unsigned long long)seed, c->cap, c->warps, c->rpb, st,
                    (int)first, (int)n};
        a.order = c->move_order;
        const int e = launch_sim(a, rng_mode);
        if (e) return fail(CHROMO_ERR_CUDA, "mc_sim launch failed: %s", launch_error(e));
        CB_MARK(1);
        CK(cudaMemcpyAsync(r + off * 3, d.r + off * 3, cnt * 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(t3 + off * 3, d.t3 + off * 3, cnt * 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(t2 + off * 3, d.t2 + off * 3, cnt * 24, cudaMemcpyDeviceToHost, st));
        CB_LAUNCH(widen_i8_kernel, (unsigned)((cs + 255) / 256), 256, 0, st, d.states + off * d.nb, stage, (long long)cs);
        CK(cudaMemcpyAsync(states + off * d.nb, stage, cs * 8, cudaMemcpyDeviceToHost, st));
        CB_MARK(2);
        CK(cudaGetLastError());
    }
    for (int k = 0; k < chunks; k++) CK(cudaStreamSynchronize(c->chunk_streams[k]));
#ifndef CHROMO_HOST_EMU
    if (timeline) {
        for (int k = 0; k < chunks; k++) {
            float t[3] = {0, 0, 0};
            for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], evs[0], evs[1 + 3 * k + i]);
            fprintf(stderr, "[chromo timeline] chunk %d: uploaded %.2f ms, kernel done %.2f ms, downloaded %.2f ms\n", k, t[0], t[1], t[2]);
        }
        for (auto &e : evs) cudaEventDestroy(e);
    }
#endif
#undef CB_MARK
    if ((rc = check_bad_values(c))) return rc;
    if (mods) c->have_mods = true;
    if (moves) {
        CK(cudaMemcpyAsync(moves, d.moves, sizeof(chromo_move_state) * d.R * CHROMO_NUM_MOVES,
                           cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return CHROMO_OK;
}

extern "C" int64_t chromo_last_attempts(chromo_ctx *c) {
    if (!c) return -1;
    cudaSetDevice(c->device);
    std::vector<unsigned long long> a(c->d.R);
    if (cudaMemcpyAsync(a.data(), c->d.attempts, a.size() * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    int64_t t = 0;
    for (auto v : a) t += (int64_t)v;
    return t;
}

extern "C" int64_t chromo_last_algo_bytes(chromo_ctx *c) {
    if (!c) return -1;
    cudaSetDevice(c->device);
    std::vector<unsigned long long> a(c->d.R);
    if (cudaMemcpyAsync(a.data(), c->d.algo_bytes, a.size() * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    int64_t t = 0;
    for (auto v : a) t += (int64_t)v;
    return t;
}

extern "C" int chromo_mc_step(chromo_ctx *c, int64_t replica, int move, double amp_move,
                              int64_t amp_bead, double mu_adjust_factor, int rng_mode, uint64_t seed,
                              int force_accept, chromo_step_report *report, int64_t *inds,
                              int64_t inds_cap, double *trial_rows, int64_t rows_cap, int64_t *touched,
                              double *dtrial, int64_t touched_cap) {
    int rc = check_ready(c);
    if (rc) return rc;
    DevCtx &d = c->d;
    if (replica < 0 || replica >= d.R) return fail(CHROMO_ERR_ARG, "replica out of range");
    if (move < 0 || move >= CHROMO_NUM_MOVES) return fail(CHROMO_ERR_ARG, "unknown move %d", move);
    if (amp_bead > d.N) // bead_selection.pyx:85-88,138-141
        return fail(CHROMO_ERR_ARG, "Bead selection window size must be less than polymer length");
    if ((move == CHROMO_TANGENT_ROTATION || move == CHROMO_END_PIVOT) && amp_bead < 1)
        return fail(CHROMO_ERR_ARG, "amp_bead must be >= 1 for this move");
    CK(cudaSetDevice(c->device));
    DebugOut h{};
    h.inds_cap = c->dbg_inds_cap;
    h.rows_cap = c->dbg_inds_cap;
    h.touched_cap = c->dbg_touched_cap;
    h.inds = c->d_dbg_inds;
    h.rows = c->d_dbg_rows;
    h.touched = c->d_dbg_touched;
    h.dtrial = c->d_dbg_dtrial;
    CK(cudaMemcpyAsync(c->d_dbg, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    McStepArgs a{d, (int)replica, move, amp_move, (int)amp_bead, mu_adjust_factor, (unsigned long long)seed,
                 force_accept, c->d_dbg, c->cap, c->stream};
    if (rng_mode != CHROMO_RNG_REPLAY && rng_mode != CHROMO_RNG_PHILOX) return fail(CHROMO_ERR_ARG, "unknown rng_mode %d", rng_mode);
    const int e = launch_step(a, rng_mode);
    if (e) return fail(CHROMO_ERR_CUDA, "mc_step launch failed: %s", launch_error(e));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h, c->d_dbg, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (report) {
        report->n_inds = h.n_inds;
        report->n_touched = h.n_touched;
        report->dE_poly = h.dE_poly;
        report->dE_field = h.dE_field;
        report->u = h.u;
        report->accepted = h.accepted;
        report->passes = h.passes;
    }
    int64_t ni = std::min<int64_t>(h.n_inds, c->dbg_inds_cap);
    int64_t nt = std::min<int64_t>(h.n_touched, c->dbg_touched_cap);
    if (inds && ni > 0)
        CK(cudaMemcpyAsync(inds, c->d_dbg_inds, 8 * std::min(ni, inds_cap), cudaMemcpyDeviceToHost, c->stream));
    if (trial_rows && ni > 0)
        CK(cudaMemcpyAsync(trial_rows, c->d_dbg_rows, 8 * (9 + d.nb) * std::min(ni, rows_cap),
                           cudaMemcpyDeviceToHost, c->stream));
    if (touched && nt > 0)
        CK(cudaMemcpyAsync(touched, c->d_dbg_touched, 8 * std::min(nt, touched_cap), cudaMemcpyDeviceToHost, c->stream));
    if (dtrial && nt > 0)
        CK(cudaMemcpyAsync(dtrial, c->d_dbg_dtrial, 8 * d.ncol * std::min(nt, touched_cap),
                           cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CHROMO_OK;
}
