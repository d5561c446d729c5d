// mc_kernel.cuh -- the fused Monte-Carlo kernel: proposal -> elastic dE ->
// field dE -> Metropolis -> commit, for one replica per warp.
//
// Why a warp per replica: moves inside a replica are strictly serial (each
// reads the density the previous one wrote), a typical move touches 1-30 beads
// (16 voxel contributions each), so the parallelism inside a move fits 32 lanes
// and the parallelism across the GPU comes from replicas.  The replica's voxel
// field stays in HBM/L2 (C2: 148 KB per replica, 1,024 replicas = 152 MB, most
// of it never touched because the polymer sits in the inscribed sphere); what
// lives in shared memory is the per-move delta-density table:
//
//   an open-addressing hash keyed by voxel super-index holding the (nb+1)
//   delta-rho columns of every voxel the move touches (the reference's
//   `density_trial` rows + `bins_found` set, fields.pyx:1427-1522), filled by
//   shared-memory atomics, then reduced over the touched voxels with warp
//   shuffles into the Flory-Huggins + reader-protein energy change
//   (fields.pyx:1675-1875).
//
// Moves whose touched-voxel set does not fit the table are evaluated in P
// hash-partition passes (voxels with bin % P == p per pass): the energy is a
// sum over voxels, so the passes are independent.
//
// Work mapping inside a move (v1, after the first ncu profile: the kernel was
// instruction-fetch bound at 2.8k warp-instructions per attempt and 5 active
// lanes): the 16 voxel contributions of a bead are spread over up to 16 lanes
// (G lanes per bead, G = 16/8/4/2/1 by segment length), the four bond energies
// of an elastic dE are evaluated by four lanes at once, and every heavy routine
// has ONE call site or is CB_NOINLINE, which keeps the SASS small enough for
// the instruction cache.
#pragma once
#include "geometry.cuh"
#include "launch.cuh"
#include "params.cuh"
#include "rng.cuh"

#define FULL_MASK 0xffffffffu
#define HASH_EMPTY (-1)

// Development-only phase timers (tools/build_variant.sh WORK x.so -DCB_PHASE_TIMERS): cycles per phase
// of an attempt, summed over warps, per move type.  Never defined in the product build.
#ifdef CB_PHASE_TIMERS
#define CB_NPHASE 16
__device__ unsigned long long cb_phase_acc[CHROMO_NUM_MOVES][CB_NPHASE];
#define CB_T0() long long cb_t_ = clock64()
#define CB_LAP(ph)                                           \
    do {                                                     \
        const long long cb_n_ = clock64();                   \
        tacc[ph] += (unsigned long long)(cb_n_ - cb_t_);     \
        cb_t_ = cb_n_;                                       \
    } while (0)
#else
#define CB_T0() do { } while (0)
#define CB_LAP(ph) do { } while (0)
#endif

#ifndef CB_MIN_BLOCKS
#define CB_MIN_BLOCKS 1 // resident blocks per SM the register allocation aims for (A/B knob)
#endif
#ifndef CB_UNIT_UNROLL
#define CB_UNIT_UNROLL 1 // voxel-contribution loop of scatter_pass (A/B knob)
#endif
constexpr int cb_unit_unroll = CB_UNIT_UNROLL;
// CB_KSEL (launch.cuh): tangent rotation: bead sets up to this size are drawn in the batched prepare
#define CB_TAN_SMALL CB_KSEL // tangent rotation: new tangents of up to this many beads are staged in shared memory
#define CB_NEWST 128    // binding: new states of up to this many beads are staged in shared memory

// one prepared attempt (see McWarp::prepare)
struct Prop {
    double ax[3];  // rotation axis (end-pivot, single-bead tangent rotation) or translation vector (slide)
    double sn, cs; // sin / cos of the rotation angle
    double u;      // Metropolis uniform (batched mode)
    int ind0, indf, n;
    int aux;       // binder (binding) | left-hand side (end-pivot) | bead (single-bead tangent rotation)
    int newst0;    // binding, n == 1: the new state
    uint32_t used; // draws of the attempt's stream consumed by prepare (batched mode)
    // segment moves: the state-DEPENDENT scalar half, from the rows as they were when it was computed
    // (segment_rows_prepare); redone if an accepted attempt of the batch changed those rows since
    double M[12];   // 3x4 affine map (slide: translation column only)
    double dE_poly; // elastic energy change of the two bonds at the segment's ends
};

// per-warp shared state (each warp of a replica's block works on its own attempt)
struct WarpSh {
    double M[12];                    // affine map of the current move
    double tan_new[CB_TAN_SMALL * 6]; // tangent rotation: new t3 | t2 of the selected beads
    int tinds[CB_TAN_SMALL];         // tangent rotation: selected beads (sequential-stream path)
    int count;                       // occupied hash slots
    int overflow;
    int last_U;                      // touched voxels of the last field dE (all passes)
    int passes;
    uint32_t draws[32];              // per-bead axis draws of tangent rotation (16 beads per chunk)
    signed char newst[CB_NEWST];     // new binding states (small path)
};

// per-replica shared state
struct ReplicaSh {
    Prop prop[32];                   // prepared attempts of the current batch
    chromo_move_state mv[CHROMO_NUM_MOVES];
    unsigned long long attempt_base; // Philox: index of the batch's first attempt in the replica's stream
    unsigned long long algo_bytes;   // SURVEY 8(d) algorithmic bytes, accumulated under the commit token
    int token;                       // attempt of the batch whose turn it is to evaluate the field and commit
    unsigned accepted;               // bit j: attempt j of the batch was accepted
    // (tangent rotation: the bead set drawn by the batched prepare lives in the attempt's own Prop::M, which only
    // segment moves use -- see tsel())
    uint32_t grs[CB_GLIBC_WORDS];    // ReplayRng state
    uint32_t rng_save[CB_GLIBC_WORDS];
    uint32_t rng_after[CB_GLIBC_WORDS];
};
static_assert(sizeof(ReplicaSh) <= CB_REPLICA_SH_BYTES && sizeof(WarpSh) <= CB_WARP_SH_BYTES, "update launch.cuh");
static_assert(CB_KSEL <= CB_TAN_SMALL, "prepared bead sets are staged in shared memory");
static_assert(CB_KSEL * sizeof(int) <= sizeof(Prop::M), "the prepared bead set of a tangent rotation overlays Prop::M");
// prepared bead set of tangent-rotation attempt `slot` of the batch
__device__ __forceinline__ int *tsel(ReplicaSh &B, int slot) { return reinterpret_cast<int *>(B.prop[slot].M); }
__device__ __forceinline__ const int *tsel(const ReplicaSh &B, int slot) { return reinterpret_cast<const int *>(B.prop[slot].M); }

struct HashTable {
    int *keys;      // [cap]
    int *list;      // [cap]   occupied slots, in claim order
    uint32_t *vals; // [cap][ncol][2]  delta-rho in fixed point (two words per cell, see Fx)
    cb_saddr vals_s; // the same as a shared-state-space address
    int cap, limit; // cap: any size >= 128 (not necessarily a power of two)
    bool prefetch;  // prefetch the density row of a voxel that joins the table (a compile-time fact of the kernel)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// ------------------------------------------------- fixed-point delta-rho cells
// Shared memory has no native 64-bit or floating-point atomic add on sm_100a
// (both compile to ATOMS.CAST.SPIN loops, three dependent shared-memory round
// trips per try -- the top stall of the v3-v6 profiles).  The delta-rho of a move
// is therefore accumulated as an integer in units of 2^-E, split into two
// INDEPENDENT 32-bit words -- word 0 collects the low 20 bits of every term
// (unsigned; 2n <= 4096 terms per cell cannot overflow it), word 1 the terms'
// arithmetic shift right by 20 (two's complement; the final sum fits) -- so that
// a term is two fire-and-forget native atomic adds with no carry to wait for
// (v8-v14 took the carry from the low word's returned value: two dependent round
// trips).  E is chosen per move from the segment length so that the sum cannot
// overflow (fx_format): at n = 32 beads the quantum is 2^-57 nm^-3, a factor 2^20
// below the 1e-9 relative tolerance on a 1e-5 nm^-3 term, and the result does not
// depend on the order in which lanes arrive (bit-reproducible runs).  Moves of more
// than 2,048 beads (no reference workload has them) keep the exact 64-bit form
// with a carry (lo_bits = 0).
struct Fx {
    double scale, inv_scale;
    int lo_bits;
};
// 2^e as a double
__device__ __forceinline__ double pow2_double(int e) { return __hiloint2double((1023 + e) << 20, 0); }
// every voxel receives at most 2n terms of magnitude <= max_state / V_min = 2^-fx_base per move
// (C.fx_base = floor(log2(V_min / max_state)), host): scaled by 2^E the sum stays below 2^(29 + lo_bits)
// (2^61 in the 64-bit form)
__device__ __forceinline__ Fx fx_format(const DevCtx &C, int n) {
    Fx f;
    f.lo_bits = n <= 2048 ? 20 : 0;
    // E <= 58 in the two-word form: a term the reference drops (|w/V| <= 1e-18, quirk 3) then rounds to 0 by
    // itself (1e-18 * 2^58 = 0.29), so the hot path needs no explicit threshold
    int e = (f.lo_bits ? 28 + f.lo_bits : 60) + C.fx_base - (n > 1 ? 32 - __clz(n - 1) : 0);
    if (f.lo_bits) e = min(e, 58);
    f.scale = pow2_double(e);
    f.inv_scale = pow2_double(-e);
    return f;
}
static __device__ CB_NOINLINE void fx_add_carry(uint32_t *cell, long long v) {
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)((unsigned long long)v >> 32);
    const uint32_t old = atomicAdd(&cell[0], lo);
    const uint32_t carry = (uint32_t)(old + lo) < lo ? 1u : 0u;
    atomicAdd(&cell[1], hi + carry);
}
__device__ __forceinline__ void fx_add(uint32_t *cell, cb_saddr cell_s, long long v, int lo_bits) {
    if (lo_bits) {
#ifdef CB_DBG_GENERIC_ATOMICS
        atomicAdd(&cell[0], (uint32_t)v & 0xFFFFFu);
        atomicAdd(&cell[1], (uint32_t)(v >> 20));
#else
        cb_red_add_u32(cell_s, (uint32_t)v & 0xFFFFFu);
        cb_red_add_u32(cell_s + 4, (uint32_t)(v >> 20));
#endif
    } else {
        fx_add_carry(cell, v);
    }
}
__device__ __forceinline__ double fx_read(const uint32_t *cell, const Fx &f) {
    const long long v = f.lo_bits ? ((long long)(int)cell[1] << 20) + (long long)cell[0]
                                  : (long long)(((unsigned long long)cell[1] << 32) | (unsigned long long)cell[0]);
    return (double)v * f.inv_scale;
}

// ---------------------------------------------------------------- hash table
__device__ __forceinline__ void table_reset_all(HashTable &H, WarpSh &S, int ncol, int lane) {
    for (int i = lane; i < H.cap; i += 32) H.keys[i] = HASH_EMPTY;
    for (int i = lane; i < H.cap * ncol * 2; i += 32) H.vals[i] = 0u;
    if (lane == 0) {
        S.count = 0;
        S.overflow = 0;
    }
    __syncwarp();
}
// clear only what the last move used
__device__ __forceinline__ void table_clear(HashTable &H, WarpSh &S, int ncol, int lane) {
    __syncwarp();
    int cnt = min(S.count, H.cap);
    for (int j = lane; j < cnt; j += 32) {
        int slot = H.list[j];
        H.keys[slot] = HASH_EMPTY;
        for (int c = 0; c < ncol; c++) *(unsigned long long *)&H.vals[(slot * ncol + c) * 2] = 0ull;
    }
    __syncwarp();
    if (lane == 0) {
        S.count = 0;
        S.overflow = 0;
    }
    __syncwarp();
}
// claim (or find) the slot of `bin`: one ATOMS.CAS on the key in the common
// case.  `checked` (moves that could overflow the table: 16 n > limit): -1 once
// the overflow flag is up; the flag rises when `limit` slots are taken, at most
// 32 more claims can be in flight, so the table never fills and every claimed
// slot is listed (table_clear relies on that).  `fresh`: this call claimed the slot.
__device__ __forceinline__ int table_claim(HashTable &H, WarpSh &S, int bin, bool checked, bool &fresh) {
    uint32_t slot = __umulhi((uint32_t)bin * 2654435761u, (uint32_t)H.cap);
    fresh = false;
    if (checked && *(volatile int *)&S.overflow) return -1;
    while (true) {
        const int prev = atomicCAS(&H.keys[slot], HASH_EMPTY, bin);
        if (prev == bin) return (int)slot;
        if (prev == HASH_EMPTY) {
            const int pos = atomicAdd(&S.count, 1);
            H.list[pos] = (int)slot;
            if (checked && pos + 1 >= H.limit) S.overflow = 1;
            fresh = true;
            return (int)slot;
        }
        slot = slot + 1 == (uint32_t)H.cap ? 0u : slot + 1;
    }
}

// w / V_access in fixed point: exact division by the constant voxel volume, or by the per-voxel one (out of
// line: only fields with assume_fully_accessible = 0 have it); |x| <= 1e-18 terms are dropped (quirk 3)
static __device__ CB_NOINLINE double div_access(const double *access_vol, int bin, double w) {
    return w / access_vol[bin];
}
template <bool GEN>
__device__ __forceinline__ long long fx_term(const DevCtx &C, double w, int bin, double scale) {
    const double d = (GEN && C.access_vol) ? div_access(C.access_vol, bin, w) : div_const(w, C.vol_bin, C.inv_vol_bin);
    if (!GEN) return __double2ll_rn(d * scale); // (the 1e-18 threshold is implied by the format, see fx_format)
    if (C.fast_n) return __double2ll_rn(d * scale); // fast_field adds every term (fields.pyx:1350-1366)
    return fabs(d) > 1E-18 ? __double2ll_rn(d * scale) : 0ll;
}
// one voxel contribution: `v` to the bead column (unless `no_bead`), v * mult[m] to binder column m
template <int NB, bool GEN>
__device__ __forceinline__ void unit_add(const DevCtx &C, HashTable &H, WarpSh &S, const double *dens_rows,
                                         bool checked, int P, int p, int bin, long long v, bool no_bead,
                                         const int mult[NB], int lo_bits_rt) {
    constexpr int NCOL = NB + 1;
    const int lo_bits = GEN ? lo_bits_rt : 20;
    if (GEN && P > 1 && (bin & (P - 1)) != p) return;
    bool fresh; // the voxel joins the touched set even when its term is zero (fields.pyx:1499-1520)
    const int slot = table_claim(H, S, bin, checked, fresh);
    if (slot < 0) return;
    // the density row is needed by table_energy: worth a prefetch when the replica has the SM to itself (the
    // two-warp kernels, C4: +3 %); with seven replicas per SM it costs 3 % (the rows of all seven compete for
    // 60 KB of L1)
    if (H.prefetch && fresh) cb_prefetch(dens_rows + (long long)bin * NCOL);
    if (GEN && v == 0) return;
    uint32_t *cell = H.vals + (size_t)slot * NCOL * 2;
    const cb_saddr cell_s = H.vals_s + (cb_saddr)slot * (NCOL * 8);
    if (!no_bead) fx_add(cell, cell_s, v, lo_bits);
#pragma unroll
    for (int m = 0; m < NB; m++)
        if (mult[m] != 0) fx_add(cell + 2 * (1 + m), cell_s + 8 * (1 + m), v * (long long)mult[m], lo_bits);
}

// ---------------------------------------------------------- scatter of a move
// One pass of get_change_in_density (fields.pyx:1430-1520) for the segment
// [ind0, ind0+n): every bead contributes -w/V*{1,state} at its current position
// and +w/V*{1,state'} at its trial position to the 8 voxels around each.
//   kind 0: trial = M r (crank-shaft, end-pivot)   kind 1: trial = r + t (slide)
//   kind 2: trial position = current position, state' = newst (binding)
// v15: two passes per chunk of up to 32 beads.
//   pass 1  the 8 voxels of the bead's CURRENT cell, G = 1/2/4/8 lanes per bead (by the beads left to
//           do): trial minus current for a bead whose trial position falls in the same cell ("merged":
//           most slides and small rotations, every binding move), minus current otherwise;
//   pass 2  the 8 voxels of the TRIAL cell of the (few) beads that left their cell, 8 lanes per bead --
//           one corner each -- whatever the segment length: the owner lane hands the trial cell over by
//           shuffles.
// (v14 walked 16 contributions per lane as soon as ONE bead of the chunk had left its cell, with most
// lanes idle in the second half, and selected cell / weight per contribution at run time.)
// Returns the confinement counters of get_confinement_dE (fields.pyx:160-193):
// x = # trial positions outside, y = # current positions outside (this lane's).
// GEN = false: the hot instance (uniform voxel volumes, a single pass, at most 2,048 beads: every reference
// workload); GEN = true: the general one behind scatter_pass_cold (per-voxel accessible volumes, partition
// passes, the 64-bit cell format).
template <int NB, bool GEN, int KIND_CT>
__device__ __forceinline__ int2 scatter_pass(const DevCtx &C, HashTable &H, WarpSh &S, int rep, int lane,
                                          int kind_rt, int ind0, int n, int binder,
                                          const signed char *newst, int P, int p, const Fx &fx) {
    constexpr int NCOL = NB + 1;
    // KIND_CT = 2: binding moves; 0: crank-shaft / end-pivot / slide (kind_rt tells which); -1: kind_rt decides
    const int kind = KIND_CT == 2 ? 2 : (KIND_CT == 0 ? (kind_rt == 1 ? 1 : 0) : kind_rt);
    const double *Rr = C.r + (long long)rep * C.N * 3;
    const signed char *ST = C.states + (long long)rep * C.N * NB;
    const double *dens_rows = C.density + (long long)rep * C.n_bins * NCOL;
    const bool checked = P > 1 || 16 * n > H.limit;
    const int nxy = C.nx * C.ny;
    const double scale = fx.scale;
    const int lo_bits = fx.lo_bits;
    int out_t = 0, out_c = 0;
    int per_iter;
#pragma unroll 1
    for (int base = 0; base < n; base += per_iter) {
        const int rem = n - base;
        const int G = rem <= 4 ? 8 : rem <= 8 ? 4 : rem <= 16 ? 2 : 1;
        per_iter = 32 / G;
        const int sub = lane & (G - 1);
        const int i = base + lane / G;
        const bool active = i < n;
        int mult[NB]; // binder column m receives (bead-column term) * mult[m]
        int clo[3], chi[3], tlo[3], thi[3];
        double cw[3], tw[3];
        bool merged = true;
#pragma unroll
        for (int m = 0; m < NB; m++) mult[m] = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) tlo[j] = thi[j] = 0, tw[j] = 0.0;
        if (active) {
            const int bead = ind0 + i;
            double x[3];
            load3(Rr + 3 * bead, x);
#pragma unroll
            for (int m = 0; m < NB; m++) mult[m] = ST[bead * NB + m];
            if (GEN && C.fast_n) bin_axes_fast(C, x, clo, chi, cw);
            else bin_axes(C, x, clo, chi, cw);
            if (kind == 2) {
                // state change only: the bead column cancels exactly (quirk 4), the binder's column gets
                // w/V * (s' - s)
#pragma unroll
                for (int m = 0; m < NB; m++) mult[m] = (m == binder) ? (int)newst[i] - mult[m] : 0;
            } else {
                double y[3];
                if (kind == 0) apply_affine(S.M, x, y);
                else
                    for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                if (p == 0 && sub == 0) {
                    if (C.confine_type == CHROMO_CONFINE_SPHERICAL) {
                        out_t += sqrt(dot3(y, y)) > C.confine_length;
                        out_c += sqrt(dot3(x, x)) > C.confine_length;
                    } else if (C.confine_type == CHROMO_CONFINE_CUBICAL) {
                        // fields.pyx:178-193: the current configuration is never counted
                        for (int j = 0; j < 3; j++) out_t += (fabs(y[j]) > C.confine_length / 2);
                    }
                }
                if (GEN && C.fast_n) bin_axes_fast(C, y, tlo, thi, tw);
                else bin_axes(C, y, tlo, thi, tw);
                merged = clo[0] == tlo[0] && clo[1] == tlo[1] && clo[2] == tlo[2];
            }
            // ---- pass 1: the current cell's corners l (bit0 x, bit1 y, bit2 z), this lane's share ----
            const int cpl = 8 / G;
            // a lane that walks all 8 corners starts at a lane-dependent corner, so that neighbouring
            // beads (which mostly share a cell) do not hit one slot together
            const int rot = G == 1 ? (lane & 7) : 0;
            const int yl = C.nx * clo[1], yh = C.nx * chi[1], zl = nxy * clo[2], zh = nxy * chi[2];
#pragma unroll 1
            for (int u = sub * cpl; u < (sub + 1) * cpl; u++) {
                const int l = (u + rot) & 7;
                const bool bx = l & 1, by = l & 2, bz = l & 4;
                const int bin = (bx ? chi[0] : clo[0]) + (by ? yh : yl) + (bz ? zh : zl);
                // products in the reference's order (x*y)*z
                const double w_c = ((bx ? 1.0 - cw[0] : cw[0]) * (by ? 1.0 - cw[1] : cw[1])) * (bz ? 1.0 - cw[2] : cw[2]);
                const long long f_c = fx_term<GEN>(C, w_c, bin, scale);
                long long v;
                if (kind == 2) {
                    v = f_c;
                } else {
                    const double w_t = ((bx ? 1.0 - tw[0] : tw[0]) * (by ? 1.0 - tw[1] : tw[1])) * (bz ? 1.0 - tw[2] : tw[2]);
                    const long long f_t = fx_term<GEN>(C, w_t, bin, scale);
                    v = merged ? f_t - f_c : -f_c;
                }
                unit_add<NB, GEN>(C, H, S, dens_rows, checked, P, p, bin, v, kind == 2, mult, lo_bits);
            }
        }
        // ---- pass 2: the trial cell of the beads that left theirs, one corner per lane ----
        if (kind != 2) {
            unsigned left = __ballot_sync(FULL_MASK, active && !merged && sub == 0);
#pragma unroll 1
            while (left) {
                // the (lane / 8)-th set bit of `left` (the beads of this round): peel the lowest set bit q times
                // (__fns is a software loop of ~30 instructions)
                unsigned peel = left;
                const int q = lane >> 3;
                if (q > 0) peel &= peel - 1;
                if (q > 1) peel &= peel - 1;
                if (q > 2) peel &= peel - 1;
                const int owner = peel ? (int)__ffs((int)peel) - 1 : -1;
                const bool has = owner >= 0 && owner < 32;
                const int src = has ? owner : 0;
                int qlo[3], qhi[3], qm[NB];
                double qw[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    qlo[j] = __shfl_sync(FULL_MASK, tlo[j], src);
                    qhi[j] = __shfl_sync(FULL_MASK, thi[j], src);
                    qw[j] = __shfl_sync(FULL_MASK, tw[j], src);
                }
#pragma unroll
                for (int m = 0; m < NB; m++) qm[m] = __shfl_sync(FULL_MASK, mult[m], src);
                if (has) {
                    const int l = lane & 7;
                    const bool bx = l & 1, by = l & 2, bz = l & 4;
                    const int bin = (bx ? qhi[0] : qlo[0]) + C.nx * (by ? qhi[1] : qlo[1]) + nxy * (bz ? qhi[2] : qlo[2]);
                    const double w_t = ((bx ? 1.0 - qw[0] : qw[0]) * (by ? 1.0 - qw[1] : qw[1])) * (bz ? 1.0 - qw[2] : qw[2]);
                    unit_add<NB, GEN>(C, H, S, dens_rows, checked, P, p, bin, fx_term<GEN>(C, w_t, bin, scale), false, qm, lo_bits);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) left &= left - 1; // the four beads just done
            }
        }
    }
    return make_int2(out_t, out_c);
}

// Partial sums of get_dE_binders_and_beads / nonspecific_interact_dE over the
// voxels currently in the table (fields.pyx:1723-1750, 1825-1840).
template <int NB>
struct FieldSums {
    double sq[NB];
    double cross[NB * NB];
    double chi;
};
template <int NB>
__device__ __forceinline__ void table_energy(const DevCtx &C, const HashTable &H, const WarpSh &S,
                                             int rep, double chi, int lane, FieldSums<NB> &F,
                                             bool want_cross, const Fx &fx) {
    constexpr int NCOL = NB + 1;
    int cnt = S.count;
    const double *dens = C.density + (long long)rep * C.n_bins * NCOL;
    const double chi_Vv = chi * (C.vol_bin / C.bead_vol); // chi * (V / v_bead) of a uniform voxel: one division per move
    for (int j = lane; j < cnt; j += 32) {
        int slot = H.list[j];
        int bin = H.keys[slot];
        const double *row = dens + (long long)bin * NCOL;
        double rho[NCOL], rn[NCOL], dr0 = 0.0;
#pragma unroll
        for (int c = 0; c < NCOL; c++) {
            rho[c] = row[c];
            const double dr = fx_read(H.vals + ((size_t)slot * NCOL + c) * 2, fx);
            if (c == 0) dr0 = dr;
            rn[c] = rho[c] + dr;
        }
#pragma unroll
        for (int a = 0; a < NB; a++) {
            double t = rn[a + 1] * rn[a + 1] - rho[a + 1] * rho[a + 1];
            if (fabs(t) < 1E-18) t = 0.0;
            F.sq[a] += t;
        }
        if (want_cross) {
#pragma unroll
            for (int a = 0; a < NB; a++)
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    double t = (rn[a + 1] * rn[b + 1]) - (rho[a + 1] * rho[b + 1]);
                    if (fabs(t) < 1E-18) t = 0.0;
                    F.cross[a * NB + b] += t;
                }
        }
        // chi * (V_access / v_bead) * phi^2 (fields.pyx:1829-1840), evaluated as (chi * (V / v)) * (phi * phi)
        const double k = C.access_vol ? chi * (C.access_vol[bin] / C.bead_vol) : chi_Vv;
        double vf0 = rho[0] * C.bead_vol;
        double vf1 = vf0 + (dr0 * C.bead_vol);
        double e = 0.0;
        if (vf1 > C.vf_limit) e += CB_E_HUGE_FIELD * vf1;
        else e += k * (vf1 * vf1);
        if (vf0 > C.vf_limit) e -= CB_E_HUGE_FIELD * vf0;
        else e -= k * (vf0 * vf0);
        F.chi += e;
    }
}
// update_affected_densities fields.pyx:1968-1975 for the voxels in the table
__device__ __forceinline__ void table_commit(const DevCtx &C, const HashTable &H, const WarpSh &S,
                                             int rep, int lane, const Fx &fx) {
    int cnt = S.count;
    double *dens = C.density + (long long)rep * C.n_bins * C.ncol;
    for (int j = lane; j < cnt; j += 32) {
        int slot = H.list[j];
        double *row = dens + (long long)H.keys[slot] * C.ncol;
        for (int c = 0; c < C.ncol; c++) row[c] += fx_read(H.vals + ((size_t)slot * C.ncol + c) * 2, fx);
    }
}
// instrumentation of the single-step kernel: the voxels in the table and their delta-rho rows, appended after
// the `base` entries of the earlier partition passes (S.last_U already counts this pass)
__device__ __forceinline__ void table_debug_dump(const DevCtx &C, const HashTable &H, const WarpSh &S,
                                              int lane, DebugOut *dbg, const Fx &fx) {
    __syncwarp();
    const int cnt = __shfl_sync(FULL_MASK, S.count, 0);
    const long long base = __shfl_sync(FULL_MASK, S.last_U - S.count, 0); // lane 0 wrote last_U itself
    for (int j = lane; j < cnt; j += 32) {
        long long o = base + j;
        if (o < dbg->touched_cap) {
            int slot = H.list[j];
            dbg->touched[o] = H.keys[slot];
            for (int c = 0; c < C.ncol; c++)
                dbg->dtrial[o * C.ncol + c] = fx_read(H.vals + ((size_t)slot * C.ncol + c) * 2, fx);
        }
    }
    if (lane == 0) dbg->n_touched = base + cnt;
    __syncwarp();
}

// Field dE of a continuous segment: compute_dE fields.pyx:1149-1233 =
// confinement + get_change_in_density + get_dE_binders_and_beads, in two stages.
//
//   field_scatter (stage 1) depends only on the positions / states of the
//       segment's own beads: it fills this warp's delta-rho table and counts the
//       confinement violations.  It does NOT read the density.
//   field_finish  (stage 2) gathers the touched voxels' density rows, forms the
//       energy change and returns dE on all lanes.  It reads the density and so
//       has to run after every earlier attempt of the replica has committed.
//
// The split is what lets the warps of a replica overlap: stage 1 of attempt
// j+1 runs while attempt j is still in stage 2 (see McWarp::attempt).
// S.passes = number of partition passes; with one pass the table still holds
// the delta-rho rows afterwards (used by the commit).  Moves whose touched set
// overflows the table are re-scattered in hash-partition passes inside stage 2.
// ddbl[a] = change in the number of doubly-bound beads (count_doubly_bound).
// the general scatter: fields with per-voxel accessible volumes, the partition passes of a move whose touched
// set overflowed the table, moves of more than 2,048 beads.  (Inlined into its three callers, two of which
// are themselves out of line: as a separate out-of-line function it lost table entries on B200 -- warp-level
// primitives behind a call from a loop with divergent lanes -- although the CPU emulation was clean.)
template <int NB>
__device__ __forceinline__ int2 scatter_pass_cold(const DevCtx &C, HashTable H, WarpSh *Sp, int rep, int lane, int kind,
                                              int ind0, int n, int binder, const signed char *newst, int P, int p) {
    return scatter_pass<NB, true, -1>(C, H, *Sp, rep, lane, kind, ind0, n, binder, newst, P, p, fx_format(C, n));
}
template <int NB>
__device__ __forceinline__ int2 field_scatter(const DevCtx &C, HashTable &H, WarpSh &S, int rep, int lane,
                                              int kind, int ind0, int n, int binder, const signed char *newst) {
    int2 conf;
    if (C.access_vol || n > 2048 || C.fast_n) conf = scatter_pass_cold<NB>(C, H, &S, rep, lane, kind, ind0, n, binder, newst, 1, 0);
    else if (kind == 2) conf = scatter_pass<NB, false, 2>(C, H, S, rep, lane, kind, ind0, n, binder, newst, 1, 0, fx_format(C, n));
    else conf = scatter_pass<NB, false, 0>(C, H, S, rep, lane, kind, ind0, n, binder, newst, 1, 0, fx_format(C, n));
    __syncwarp();
    return conf;
}
// the touched set does not fit the table (rare): re-scatter in P = 2, 4, ... hash-partition passes
// (voxels with bin % P == p per pass; the energy is a sum over voxels).  Returns P; the table ends empty.
template <int NB, bool DEBUG>
__device__ CB_NOINLINE int field_energy_multipass(const DevCtx &C, HashTable H, WarpSh *Sp, int rep, int lane, int kind,
                                                  int ind0, int n, int binder, const signed char *newst, double chi,
                                                  FieldSums<NB> *Fp, DebugOut *dbg) {
    constexpr int NCOL = NB + 1;
    WarpSh &S = *Sp;
    FieldSums<NB> &F = *Fp;
    const Fx fx = fx_format(C, n);
    const bool want_cross = C.any_cross != 0;
    int P = 1;
    bool failed = true;
    while (failed) {
        P *= 2;
#pragma unroll
        for (int a = 0; a < NB; a++) F.sq[a] = 0.0;
#pragma unroll
        for (int a = 0; a < NB * NB; a++) F.cross[a] = 0.0;
        F.chi = 0.0;
        if (DEBUG && lane == 0) dbg->n_touched = 0;
        failed = false;
        for (int p = 0; p < P; p++) {
            table_clear(H, S, NCOL, lane);
            (void)scatter_pass_cold<NB>(C, H, &S, rep, lane, kind, ind0, n, binder, newst, P, p);
            __syncwarp();
            if (S.overflow) {
                failed = true;
                break;
            }
            table_energy<NB>(C, H, S, rep, chi, lane, F, want_cross, fx);
            if (lane == 0) S.last_U = (p == 0 ? 0 : S.last_U) + S.count;
            if (DEBUG) table_debug_dump(C, H, S, lane, dbg, fx);
        }
    }
    table_clear(H, S, NCOL, lane);
    return P;
}
template <int NB, bool DEBUG>
__device__ __forceinline__ double field_finish(const DevCtx &C, HashTable &H, WarpSh &S, int rep, int lane,
                                               int kind, int ind0, int n, int binder, const signed char *newst,
                                               int2 conf, const int *ddbl, DebugOut *dbg) {
    constexpr int NCOL = NB + 1;
    const double chi = C.chi[rep];
    const Fx fx = fx_format(C, n);
    FieldSums<NB> F;
    const bool want_cross = C.any_cross != 0;
#pragma unroll
    for (int a = 0; a < NB; a++) F.sq[a] = 0.0;
#pragma unroll
    for (int a = 0; a < NB * NB; a++) F.cross[a] = 0.0;
    F.chi = 0.0;
    if (DEBUG && lane == 0) dbg->n_touched = 0;
    int P = 1;
    if (!S.overflow) {
        table_energy<NB>(C, H, S, rep, chi, lane, F, want_cross, fx);
        if (lane == 0) S.last_U = S.count;
        if (DEBUG) table_debug_dump(C, H, S, lane, dbg, fx);
    } else {
        P = field_energy_multipass<NB, DEBUG>(C, H, &S, rep, lane, kind, ind0, n, binder, newst, chi, &F, dbg);
    }
    if (lane == 0) S.passes = P;
    // ---- reduce and assemble in the reference's order ----
    double dE = 0.0;
    if (kind != 2) { // compute_dE fields.pyx:1209-1211
        int nt = warp_sum_int(conf.x), nc = warp_sum_int(conf.y);
        dE += (double)nt * CB_E_HUGE_FIELD;
        dE -= (double)nc * CB_E_HUGE_FIELD;
    }
    double bb = 0.0; // get_dE_binders_and_beads fields.pyx:1760-1790
#pragma unroll
    for (int a = 0; a < NB; a++) {
        double tot = warp_sum(F.sq[a]);
        bb += C.pref[a] * tot;
        bb += C.e_intra[a] * (double)(ddbl ? ddbl[a] : 0);
    }
#pragma unroll
    for (int a = 0; a < NB; a++)
#pragma unroll
        for (int b = 0; b < NB; b++) {
            double tot = want_cross ? warp_sum(F.cross[a * NB + b]) : 0.0;
            bb += C.xpref[a * NB + b] * tot;
        }
    bb += warp_sum(F.chi);
    dE += bb;
    __syncwarp();
    return dE;
}

// accepted move whose delta-rho needed several partition passes (rare): redo
// each pass and apply it (update_affected_densities fields.pyx:1968-1975)
template <int NB>
__device__ CB_NOINLINE void field_commit_multipass(const DevCtx &C, HashTable H, WarpSh *Sp, int rep, int lane,
                                                   int kind, int ind0, int n, int binder,
                                                   const signed char *newst) {
    WarpSh &S = *Sp;
    const int passes = S.passes;
    const Fx fx = fx_format(C, n);
    for (int p = 0; p < passes; p++) {
        table_clear(H, S, NB + 1, lane);
        (void)scatter_pass_cold<NB>(C, H, Sp, rep, lane, kind, ind0, n, binder, newst, passes, p);
        __syncwarp();
        table_commit(C, H, S, rep, lane, fx);
    }
}

// NullField.compute_dE fields.pyx:300-318: only the confinement acts
__device__ __forceinline__ double confinement_dE_segment(const DevCtx &C, const WarpSh &S, int rep, int lane,
                                                      int kind, int ind0, int n) {
    const double *Rr = C.r + (long long)rep * C.N * 3;
    int out_t = 0, out_c = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        if (i < n) {
            double x[3], y[3];
            load3(Rr + 3 * (ind0 + i), x);
            if (kind == 0) apply_affine(S.M, x, y);
            else
                for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
            if (C.confine_type == CHROMO_CONFINE_SPHERICAL) {
                out_t += (sqrt(dot3(y, y)) > C.confine_length);
                out_c += (sqrt(dot3(x, x)) > C.confine_length);
            } else {
                for (int j = 0; j < 3; j++) out_t += (fabs(y[j]) > C.confine_length / 2);
            }
        }
    }
    int nt = warp_sum_int(out_t), nc = warp_sum_int(out_c);
    double dE = (double)nt * CB_E_HUGE_FIELD;
    dE -= (double)nc * CB_E_HUGE_FIELD;
    return dE;
}

// ------------------------------------------------------------- elastic pairs
// Energy of bond `bond` (beads bond, bond+1) with the (r, t3) of one of its
// beads optionally replaced by trial values: moved = 0 none, 1 first bead,
// 2 second bead.  E_pair with dr / dr_par / dr_perp / bend built as in
// bead_pair_dE_poly_forward / _reverse (polymers.pyx:1148-1175, 1253-1346).
// SSTWLC builds (-DCB_TWIST=1, their own translation units) add the twist term of E_pair_with_twist
// (polymers.pyx:2050-2102) to every bond energy: TwistRows carries the replica's t2 rows and its
// {eps_twist, natural twist} rows, `t2n` the trial t2 of the moved bead.  In the default build the
// struct is empty and nothing below changes.
#ifndef CB_TWIST
#define CB_TWIST 0
#endif
struct TwistRows {
#if CB_TWIST
    const double *T2, *tw;
    const double *det; // DetailedChromatin: nucleosome constants, nullptr for a plain SSTWLC
#endif
};
__device__ __forceinline__ TwistRows twist_rows(const DevCtx &C, int rep) {
    TwistRows t;
#if CB_TWIST
    t.T2 = C.t2 + (long long)rep * C.N * 3;
    t.tw = C.twist + (long long)rep * C.twist_stride;
    t.det = C.detailed;
#endif
    return t;
}
__device__ __forceinline__ double pair_energy_p(const double *Rr, const double *T3, const double *bondp, int bond,
                                                int moved, const double rn[3], const double tn[3],
                                                const TwistRows &TW, const double *t2n);
__device__ __forceinline__ double pair_energy(const DevCtx &C, int rep, int bond, int moved,
                                              const double rn[3], const double tn[3], const double *t2n) {
    return pair_energy_p(C.r + (long long)rep * C.N * 3, C.t3 + (long long)rep * C.N * 3,
                         C.bond + (long long)rep * C.bond_stride, bond, moved, rn, tn, twist_rows(C, rep), t2n);
}

// the same from explicit row pointers (`bondp` = the replica's bond-parameter rows)
__device__ __forceinline__ double pair_energy_p(const double *Rr, const double *T3, const double *bondp, int bond,
                                                int moved, const double rn[3], const double tn[3],
                                                const TwistRows &TW, const double *t2n) {
    double r0[3], r1[3], t0[3], t1[3];
    load3(Rr + 3 * bond, r0);
    load3(Rr + 3 * bond + 3, r1);
    load3(T3 + 3 * bond, t0);
    load3(T3 + 3 * bond + 3, t1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        r0[j] = moved == 1 ? rn[j] : r0[j];
        t0[j] = moved == 1 ? tn[j] : t0[j];
        r1[j] = moved == 2 ? rn[j] : r1[j];
        t1[j] = moved == 2 ? tn[j] : t1[j];
    }
    const double *p = bondp + (long long)bond * 5;
    Bond B;
    B.eps_bend = p[0];
    B.eps_par = p[1];
    B.eps_perp = p[2];
    B.gamma = p[3];
    B.eta = p[4];
#if CB_TWIST
    double u0[3], u1[3];
    load3(TW.T2 + 3 * bond, u0);
    load3(TW.T2 + 3 * bond + 3, u1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        u0[j] = moved == 1 ? t2n[j] : u0[j];
        u1[j] = moved == 2 ? t2n[j] : u1[j];
    }
    if (TW.det) { // DetailedChromatin.continuous_dE_poly polymers.pyx:2503-2607: exit of bead 0 -> entry of bead 1
        nucleosome_frame(TW.det, r0, t0, u0, 1);
        nucleosome_frame(TW.det, r1, t1, u1, 0);
    }
    return bond_energy(B, r0, r1, t0, t1) + twist_energy(TW.tw + 2 * bond, twist_omega(u0, t0, u1, t1));
#else
    return bond_energy(B, r0, r1, t0, t1);
#endif
}

// The state-dependent scalar half of a segment move (crank-shaft, end-pivot, slide), ONE THREAD per
// attempt: the 3x4 affine map from the current positions (arbitrary_axis_rotation linalg.pyx:62-139,
// get_crank_shaft_axis / _fulcrum move_funcs.pyx:157-280, get_end_pivot_fulcrum 349-398) and the elastic
// energy change of the two bonds at the segment's ends (continuous_dE_poly polymers.pyx:1084-1146:
// (E_left' - E_left) + (E_right' - E_right)), written to P->M and P->dE_poly.
// Out of line and on explicit pointers: the one copy serves the batched prepare (32 attempts on 32
// lanes, off the per-attempt path) and the recompute of an attempt whose rows were changed by an
// accepted attempt of the same batch.  Returns 1 without writing when the crank-shaft axis is
// degenerate and `axis_in` is null: the caller draws one (move_funcs.pyx:229-230) and calls again.
static __device__ CB_NOINLINE int segment_rows_prepare(const double *Rr, const double *T3, const double *bondp, int N,
                                                       int mtype, Prop *P, const double *axis_in,
                                                       const TwistRows TW) {
    const int ind0 = P->ind0, indf = P->indf;
    double M[12];
    if (mtype == CHROMO_SLIDE) {
#pragma unroll
        for (int j = 0; j < 12; j++) M[j] = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) M[4 * j + 3] = P->ax[j]; // generate_translation_mat linalg.pyx:172-199
    } else {
        int ful;
        double axis[3], pt[3];
        if (mtype == CHROMO_CRANK_SHAFT) {
            int a, b; // get_crank_shaft_axis move_funcs.pyx:157-234
            if (ind0 == indf - 1 && ind0 == 0) { a = indf; b = ind0; }
            else if (ind0 == indf - 1 && ind0 == N - 1) { a = ind0; b = ind0 - 1; }
            else if (ind0 == 0 && indf == N) { a = indf - 1; b = ind0; }
            else if (ind0 == 0) { a = indf; b = ind0; }
            else if (indf == N) { a = indf - 1; b = ind0 - 1; }
            else { a = indf; b = ind0 - 1; }
            // get_crank_shaft_fulcrum move_funcs.pyx:237-280
            if (ind0 == 0 && indf != N) ful = indf;
            else if (ind0 != 0 && indf == N) ful = ind0 - 1;
            else if (ind0 == 0 && indf == N) ful = ind0;
            else ful = ind0 - 1;
#pragma unroll
            for (int j = 0; j < 3; j++) axis[j] = Rr[3 * a + j] - Rr[3 * b + j];
            const double mag = sqrt((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2]);
            if (mag < 1E-5) {
                if (!axis_in) return 1;
#pragma unroll
                for (int j = 0; j < 3; j++) axis[j] = axis_in[j];
            } else {
                const double sc = 1.0 / mag;
#pragma unroll
                for (int j = 0; j < 3; j++) axis[j] = axis[j] * sc;
            }
        } else { // end pivot: get_end_pivot_fulcrum move_funcs.pyx:349-398
            if (ind0 == 0 && indf != N) ful = indf;
            else if (ind0 != 0 && indf == N) ful = ind0 - 1;
            else if (ind0 == 0 && indf == N && P->aux == 1) ful = indf - 1;
            else ful = ind0;
#pragma unroll
            for (int j = 0; j < 3; j++) axis[j] = P->ax[j];
        }
        load3(Rr + 3 * ful, pt);
        double Rm[9], tv[3];
        rotation_3x3(axis, P->sn, P->cs, Rm);
        rotation_translation(axis, pt, P->sn, P->cs, tv);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            M[4 * j] = Rm[3 * j];
            M[4 * j + 1] = Rm[3 * j + 1];
            M[4 * j + 2] = Rm[3 * j + 2];
            M[4 * j + 3] = tv[j];
        }
    }
    // the two end bonds: which = 0 left (bond ind0-1, its second bead moves), 1 right (bond indf-1, its
    // first bead moves)
    double d[2] = {0.0, 0.0};
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        const bool present = which == 0 ? (ind0 != 0) : (indf != N);
        if (!present) continue;
        const int bond = which == 0 ? ind0 - 1 : indf - 1;
        const int mbead = which == 0 ? ind0 : indf - 1;
        double r[3], t[3], rn[3], tn[3];
        load3(Rr + 3 * mbead, r);
        load3(T3 + 3 * mbead, t);
        if (mtype == CHROMO_SLIDE) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                rn[j] = r[j] + M[4 * j + 3];
                tn[j] = t[j];
            }
        } else {
            apply_affine(M, r, rn);
            apply_rot(M, t, tn);
        }
        const double *t2n = nullptr;
#if CB_TWIST
        double u[3], un[3];
        load3(TW.T2 + 3 * mbead, u);
        if (mtype == CHROMO_SLIDE) {
#pragma unroll
            for (int j = 0; j < 3; j++) un[j] = u[j];
        } else {
            apply_rot(M, u, un);
        }
        t2n = un;
#endif
        const double e_trial = pair_energy_p(Rr, T3, bondp, bond, which == 0 ? 2 : 1, rn, tn, TW, t2n);
        const double e_cur = pair_energy_p(Rr, T3, bondp, bond, 0, rn, tn, TW, t2n);
        d[which] = e_trial - e_cur;
    }
#pragma unroll
    for (int j = 0; j < 12; j++) P->M[j] = M[j];
    P->dE_poly = d[0] + d[1];
    return 0;
}

// ------------------------------------------------------- bead selection (lane 0)
// (double)rand() / RAND_MAX, exactly (div_const)
__device__ __forceinline__ double u01(uint32_t x) {
    return div_const((double)x, CB_RAND_MAX, 1.0 / CB_RAND_MAX);
}

// capped_exponential bead_selection.pyx:19-67
template <class Rng>
__device__ __forceinline__ int capped_exponential(Rng &g, int window, int cap) {
    long long r;
    do {
        r = (long long)(-log10(u01(g.next31()) + 0.00001) * (double)window * 0.45 + 1.0001);
    } while (r > cap);
    return (int)r;
}
// check_bead_bounds bead_selection.pyx:157-192
__device__ __forceinline__ void check_bead_bounds(int b0, int b1, int N, int &ind0, int &indf) {
    b0 = max(min(b0, N), 0);
    if (b1 > N) {
        ind0 = b0;
        indf = N;
    } else if (b1 < 0) {
        ind0 = 0;
        indf = b0 + 1;
    } else if (b0 == b1) {
        ind0 = b0;
        indf = b0 + 1;
    } else {
        ind0 = min(b0, b1);
        indf = max(b0, b1);
    }
}

// =================================================================== moves
// An attempt is split into
//   prepare : everything that does NOT depend on the replica's state -- RNG
//             draws, exponential window (log10), the trigonometry, the bead set
//             of a tangent rotation -- written to a `Prop` record in shared
//             memory.  With the counter-based Philox generator attempt t has its
//             own stream, so a batch of up to 32 attempts of one move type is
//             prepared by 32 lanes AT ONCE (same code, different data).  With
//             the replayed reference streams (sequential by nature) the batch
//             size is 1 and lane 0 prepares, in the reference's draw order.
//   stage 1 : reads the replica's bead rows only: affine map from the current
//             positions, elastic / binding dE, scatter of the delta-density
//             into this warp's table (or the new tangents of a tangent
//             rotation).
//   stage 2 : reads the density: gather + energy reduction, Metropolis, commit.
//
// NW warps work on one replica.  Warp w takes attempts w, w+NW, ... of the
// batch; stage 1 runs AHEAD of the attempt's turn, stage 2 runs when the
// replica's commit token reaches the attempt, so stage 1 of attempt j+1
// overlaps stage 2 of attempt j while the attempts still take effect strictly
// in order.  If an attempt that committed in between was accepted and touched
// beads this attempt reads (rare: segments of ~30 beads out of 10^4), stage 1 is
// simply redone once the turn has come -- so the result is the sequential
// one, bit for bit (tests/test_gpu_scale.py compares NW = 1 against NW = 2).
// Every heavy routine has exactly ONE call site (inlined once: no ABI spills,
// shared-memory address spaces stay visible to the compiler).
template <class Rng, bool DEBUG, int NB, int NW>
struct McWarp {
    static constexpr int NCOL = NB + 1;
    static constexpr bool BATCH = Rng::kBatched;
    static_assert(NW == 1 || Rng::kBatched, "several warps per replica need the counter-based generator");
    const DevCtx &C;
    ReplicaSh &B;
    WarpSh &S;
    HashTable &H;
    Rng &rng;
    int rep, lane, wid, local; // replica, lane, warp within the replica, replica within the block
    double mu_adjust;
    int force_accept; // DEBUG only: -1 Metropolis, 0/1 forced
    DebugOut *dbg;
    unsigned long long abase; // Philox: index of the batch's first attempt in the replica's stream
#ifdef CB_PHASE_TIMERS
    unsigned long long tacc[CB_NPHASE] = {};
    __device__ __forceinline__ void flush_timers(int mtype) {
        if (lane == 0)
            for (int i = 0; i < CB_NPHASE; i++) atomicAdd(&cb_phase_acc[mtype][i], tacc[i]);
        for (int i = 0; i < CB_NPHASE; i++) tacc[i] = 0;
    }
#endif

    __device__ __forceinline__ double *R_() const { return C.r + (long long)rep * C.N * 3; }
    __device__ __forceinline__ double *T3_() const { return C.t3 + (long long)rep * C.N * 3; }
    __device__ __forceinline__ double *T2_() const { return C.t2 + (long long)rep * C.N * 3; }
    __device__ __forceinline__ signed char *ST_() const { return C.states + (long long)rep * C.N * NB; }
    __device__ __forceinline__ const signed char *MOD_() const { return C.mods + (long long)rep * C.N * NB; }
    __device__ __forceinline__ bool has_field() const {
        return C.field_active || C.confine_type != CHROMO_CONFINE_NONE;
    }
    __device__ __forceinline__ void block_sync() const {
        if (NW > 1) cb_bar_sync(1 + local, 32 * NW); // the replica's own named barrier
        else __syncwarp();
    }

    // ---- the replica's commit token -------------------------------------------
    __device__ __forceinline__ void wait_turn(int slot) const {
        if (NW > 1) {
            if (lane == 0)
                while (*(volatile int *)&B.token != slot) cb_backoff();
            __syncwarp();
            __threadfence_block();
        }
    }
    __device__ __forceinline__ void pass_turn(int slot, int acc) const {
        if (NW == 1) { // same warp: the next attempt reads the bit after the __syncwarp that ends this one
            if (acc && lane == 0) B.accepted |= 1u << slot;
        } else {
            __threadfence_block(); // this lane's commit stores, before the token moves
            __syncwarp();
            if (lane == 0) {
                if (acc) *(volatile unsigned *)&B.accepted = *(volatile unsigned *)&B.accepted | (1u << slot);
                __threadfence_block();
                *(volatile int *)&B.token = slot + 1;
            }
        }
    }
    // did an attempt that committed after this warp's previous turn change rows that stage 1 of
    // attempt `slot` has read?  Writes of a segment move: beads [ind0, indf); reads: one more bead
    // on either side.  Tangent rotation: the selected beads; reads: their neighbours as well.
    __device__ __forceinline__ bool stale(int mtype, int slot) const {
        if (NW == 1) return false;
        const unsigned accepted = *(volatile unsigned *)&B.accepted;
        const Prop &P = B.prop[slot];
        bool hit = false;
        for (int j = max(0, slot - NW + 1); j < slot; j++) {
            if (!((accepted >> j) & 1u)) continue;
            const Prop &Q = B.prop[j];
            if (mtype != CHROMO_TANGENT_ROTATION) {
                hit |= (Q.ind0 < P.indf + 1) && (P.ind0 - 1 < Q.indf);
            } else if (Q.n > CB_KSEL) {
                hit = true; // its bead set was drawn at execute time: assume the worst
            } else {
                // lane a walks this attempt's beads against all of the other's
                const int a = lane;
                if (a < P.n) {
                    const int mine = P.n == 1 ? P.aux : tsel(B, slot)[a];
                    for (int q = 0; q < Q.n; q++) {
                        const int other = Q.n == 1 ? Q.aux : tsel(B, j)[q];
                        hit |= (other - mine <= 1) && (mine - other <= 1);
                    }
                }
            }
        }
        return __any_sync(FULL_MASK, hit);
    }

    // ---- prepared affine map / elastic dE of a segment move (segment_rows_prepare) ----
    __device__ __forceinline__ const double *bond_rows() const { return C.bond + (long long)rep * C.bond_stride; }
    // accepted attempts of the batch before `slot` (bits only ever get set while a batch runs)
    __device__ __forceinline__ unsigned accepted_before(int slot) const {
        return *(volatile unsigned *)&B.accepted & ((1u << slot) - 1u);
    }
    // did one of the accepted attempts in `bits` write rows that attempt `slot` reads?  A segment move
    // writes beads [ind0, indf) and reads one more bead on either side (batches hold ONE move type).
    __device__ __forceinline__ bool rows_changed(unsigned bits, int slot) const {
        const Prop &P = B.prop[slot];
        bool hit = false;
        if ((bits >> lane) & 1u) {
            const Prop &Q = B.prop[lane];
            hit = (Q.ind0 < P.indf + 1) && (P.ind0 - 1 < Q.indf);
        }
        return __any_sync(FULL_MASK, hit);
    }
    // one thread redoes the state-dependent half of attempt `slot` from the rows as they are now
    __device__ __forceinline__ void rows_recompute(int mtype, int slot) {
        if (lane == 0) {
            Prop *Pp = &B.prop[slot];
            if (segment_rows_prepare(R_(), T3_(), bond_rows(), C.N, mtype, Pp, nullptr, twist_rows(C, rep))) {
                if (BATCH) rng.seek_attempt(abase + (unsigned long long)slot, Pp->used);
                double axis[3];
                const uint32_t d1 = rng.next31(), d2 = rng.next31();
                degenerate_axis(d1, d2, axis);
                (void)segment_rows_prepare(R_(), T3_(), bond_rows(), C.N, mtype, Pp, axis, twist_rows(C, rep));
            }
        }
        __syncwarp();
    }

    // ======================================================== prepare
    // lanes [0, cnt) of warp 0 each prepare one attempt of move type `mtype`
    __device__ __forceinline__ void prepare(int mtype, int cnt) {
        if (lane >= cnt) return;
        const int N = C.N;
        Prop &P = B.prop[lane];
        const chromo_move_state &mv = B.mv[mtype];
        if (BATCH) {
            rng.seek_attempt(abase + (unsigned long long)lane);
            P.u = u01(rng.next31()); // Metropolis uniform: draw 0 of the attempt's own stream
        }
        const bool crank = mtype == CHROMO_CRANK_SHAFT, pivot = mtype == CHROMO_END_PIVOT;
        const bool slide = mtype == CHROMO_SLIDE, bind = mtype == CHROMO_CHANGE_BINDING_STATE;
        const bool tangent = mtype == CHROMO_TANGENT_ROTATION;
        double amp = 0.0;
        uint32_t d1 = 0, d2 = 0;
        int b0 = 0, lhs = 0, binder = 0, ind0 = 0, indf = 0, k = 0, bead = 0;
        bool sphere = pivot || slide;
        // ---- integer draws, in the reference's order ----
        if (crank) { // move_funcs.pyx:80-84
            amp = mv.amp_move * (u01(rng.next31()) - 0.5);
            b0 = (int)(u01(rng.next31()) * (double)N);
        } else if (pivot) { // move_funcs.pyx:325-326
            amp = mv.amp_move * (u01(rng.next31()) - 0.5);
            lhs = (int)(rng.next31() % 2u);
        } else if (slide) { // move_funcs.pyx:441-445
            amp = mv.amp_move * u01(rng.next31());
            d1 = rng.next31();
            d2 = rng.next31();
            b0 = (int)(rng.next31() % (uint32_t)N);
        } else if (bind) { // move_funcs.pyx:763-766
            binder = (int)(rng.next31() % (uint32_t)NB);
            b0 = (int)(rng.next31() % (uint32_t)N);
        } else { // tangent_rotation move_funcs.pyx:503-504
            amp = mv.amp_move * (u01(rng.next31()) - 0.5);
            k = (int)(rng.next31() % (uint32_t)mv.amp_bead) + 1;
            if (k == 1) { // one bead: index, then its axis (get_inds 552-582, rotate_select_beads 517-549)
                bead = (int)(rng.next31() % (uint32_t)N);
                d1 = rng.next31();
                d2 = rng.next31();
                sphere = true;
            } else if (BATCH && k <= CB_KSEL) {
                // get_inds move_funcs.pyx:552-582: k distinct draws, duplicates redrawn; the per-bead
                // axis draws follow in the attempt's stream and are made at execute time
                int *sel = tsel(B, lane);
                for (int i = 0; i < k; i++) {
                    int c;
                    bool dup;
                    do {
                        c = (int)(rng.next31() % (uint32_t)N);
                        dup = false;
                        for (int q = 0; q < i; q++) dup |= (sel[q] == c);
                    } while (dup);
                    sel[i] = c;
                }
            }
        }
        if (!tangent) {
            // exponential window: from_left / from_right / from_point (bead_selection.pyx:69-154)
            int side = 0, ws = mv.amp_bead, ce = 0;
            bool do_ce = true;
            if (!pivot) {
                if (mv.amp_bead < 1) do_ce = false;
                else {
                    side = (int)(rng.next31() % 2u);
                    ws = side == 0 ? max(min(mv.amp_bead, b0), 1) : max(min(mv.amp_bead, N - b0), 1);
                }
            }
            if (do_ce) ce = capped_exponential(rng, ws, ws);
            if (pivot) {
                if (lhs == 1) {
                    ind0 = 0;
                    indf = ce + 1;
                } else {
                    ind0 = N - ce;
                    indf = N;
                }
                d1 = rng.next31(); // la.uniform_sample_unit_sphere() move_funcs.pyx:336
                d2 = rng.next31();
            } else {
                int b1 = !do_ce ? b0 : (side == 0 ? max(b0, 1) - ce : ce + b0);
                if (crank) b1 = max(b1, 1);
                check_bead_bounds(b0, b1, N, ind0, indf);
            }
        }
        const int n = tangent ? k : indf - ind0;
        P.ind0 = ind0;
        P.indf = indf;
        P.n = n;
        P.aux = bind ? binder : (pivot ? lhs : bead);
        P.newst0 = 0;
        if (bind && n >= 1) { // conduct_change_binding_states move_funcs.pyx:778-820
            if (!BATCH) { // lane 0, sequential stream: draw all n states now, as the reference does
                signed char *dst = (n <= CB_NEWST) ? S.newst : (C.st_new + (long long)rep * N);
                for (int i = 0; i < n; i++) dst[i] = (signed char)rng.randint(C.sites[binder] + 1);
            } else if (n == 1) {
                P.newst0 = rng.randint(C.sites[binder] + 1);
            } // n > 1 in a batch: drawn at execute time from this attempt's stream
        }
        // ---- the trigonometry (one out-of-line sincos / acos each, shared by all call sites) ----
        // uniform_sample_unit_sphere linalg.pyx:23-59
        double sn = 0.0, cs = 1.0;
        if (!(slide || bind)) {
            const double2 sc = sincos_ni(amp);
            sn = sc.x;
            cs = sc.y;
        }
        double axis[3] = {0.0, 0.0, 0.0};
        if (sphere) {
            unit_sphere_point<!BATCH>(u01(d1), u01(d2), axis);
        }
        P.sn = sn;
        P.cs = cs;
#pragma unroll
        for (int j = 0; j < 3; j++) P.ax[j] = slide ? axis[j] * amp : axis[j]; // generate_translation_mat linalg.pyx:172-199
        if (BATCH) P.used = rng.position();
        // ---- the state-dependent half of a segment move, from the rows as they are now ----
        if ((crank || pivot || slide) && n > 0) {
            if (segment_rows_prepare(R_(), T3_(), bond_rows(), N, mtype, &P, nullptr, twist_rows(C, rep))) {
                // degenerate crank-shaft axis: two more draws (move_funcs.pyx:229-230); in the batched
                // mode they are draws `used`, `used + 1` of the attempt's own stream
                double dax[3];
                const uint32_t e1 = rng.next31(), e2 = rng.next31();
                degenerate_axis(e1, e2, dax);
                (void)segment_rows_prepare(R_(), T3_(), bond_rows(), N, mtype, &P, dax, twist_rows(C, rep));
            }
        }
    }

    // ======================================================== stage 1
    // ---- binding part of dE (the elastic part of the other segment moves is prepared per batch) ----
    __device__ __forceinline__ double binding_dE_poly(int ind0, int n, int binder, const signed char *newst,
                                                      int ddbl[NB]) {
        // binding_dE / bead_binding_dE polymers.pyx:1383-1538
        const signed char *ST = ST_();
        const signed char *MOD = MOD_();
        double de = 0.0;
#pragma unroll
        for (int m = 0; m < NB; m++) ddbl[m] = 0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                int bead = ind0 + i;
                double d = 0.0;
                int sc[NB], sn[NB];
#pragma unroll
                for (int m = 0; m < NB; m++) {
                    sc[m] = ST[bead * NB + m];
                    sn[m] = (m == binder) ? (int)newst[i] : sc[m];
                    ddbl[m] += (sn[m] == 2) - (sc[m] == 2); // count_doubly_bound fields.pyx:1877-1937
                }
                if (C.max_binders != -1) {
                    long long tot = 0;
#pragma unroll
                    for (int m = 0; m < NB; m++) tot += sn[m];
                    if (tot > C.max_binders) d += CB_E_HUGE_POLY * (double)(tot - C.max_binders);
                    tot = 0;
#pragma unroll
                    for (int m = 0; m < NB; m++) tot += sc[m];
                    if (tot > C.max_binders) d -= CB_E_HUGE_POLY * (double)(tot - C.max_binders);
                }
#pragma unroll
                for (int m = 0; m < NB; m++) {
                    int Nm = MOD[bead * NB + m];
                    const double *Ft = C.bindF + (m * C.S1 + Nm) * C.S1;
                    double mu = C.mu[(long long)rep * NB + m];
                    d += Ft[sn[m]];
                    d -= Ft[sc[m]];
                    if (mu > 0) {
                        d -= (double)sn[m] * (mu * (-mu_adjust + 2.0));
                        d += (double)sc[m] * (mu * (-mu_adjust + 2.0));
                    } else {
                        d -= (double)sn[m] * mu * mu_adjust;
                        d += (double)sc[m] * mu * mu_adjust;
                    }
                }
                de += d;
            }
        }
        double dE_poly = warp_sum(de);
        if (n == 1) dE_poly = __shfl_sync(FULL_MASK, de, 0); // exact for the default amp_bead = 1
#pragma unroll
        for (int m = 0; m < NB; m++) ddbl[m] = warp_sum_int(ddbl[m]);
        return dE_poly;
    }

    // MCAdapter.accept moves.pyx:190-226 for the segment moves
    __device__ __forceinline__ void segment_commit(int kind, int ind0, int n, int binder,
                                                   const signed char *newst) {
        if (C.field_active) {
            if (S.passes == 1) table_commit(C, H, S, rep, lane, fx_format(C, n));
            else field_commit_multipass<NB>(C, H, &S, rep, lane, kind, ind0, n, binder, newst);
        }
        if (kind == 2) {
            __syncwarp();
            signed char *ST = ST_();
            for (int base = 0; base < n; base += 32) {
                int i = base + lane;
                if (i < n) ST[(ind0 + i) * NB + binder] = newst[i];
            }
            return;
        }
        double *Rr = R_(), *T3 = T3_(), *T2 = T2_();
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                const int o = 3 * (ind0 + i);
                double x[3], y[3];
                load3(Rr + o, x);
                if (kind == 0) {
                    // all three rows first: the arrays could alias as far as the compiler knows, so a load placed
                    // after a store waits for it -- three trips to HBM in a row instead of one
                    double a[3], b[3];
                    load3(T3 + o, a);
                    load3(T2 + o, b);
                    apply_affine(S.M, x, y);
                    store3(Rr + o, y);
                    apply_rot(S.M, a, y);
                    store3(T3 + o, y);
                    apply_rot(S.M, b, y);
                    store3(T2 + o, y);
                } else {
#pragma unroll
                    for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                    store3(Rr + o, y);
                }
            }
        }
    }

    // instrumentation of the single-step kernel: moved beads + their trial rows
    __device__ __forceinline__ void debug_report(int kind, int ind0, int n, int binder,
                                                 const signed char *newst, double dE_poly, double dE_field) {
        const double *Rr = R_(), *T3 = T3_(), *T2 = T2_();
        const signed char *ST = ST_();
        const int W = 9 + NB;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                if (i < dbg->inds_cap) dbg->inds[i] = ind0 + i;
                if (i < dbg->rows_cap) {
                    const int o = 3 * (ind0 + i);
                    double x[3], t[3], y[3], tn[3];
                    double *row = dbg->rows + (long long)i * W;
                    load3(Rr + o, x);
                    load3(T3 + o, t);
                    if (kind == 2) {
                        store3(row, x);
                        store3(row + 3, t);
                    } else {
                        if (kind == 0) {
                            apply_affine(S.M, x, y);
                            apply_rot(S.M, t, tn);
                        } else {
                            for (int j = 0; j < 3; j++) {
                                y[j] = x[j] + S.M[4 * j + 3];
                                tn[j] = t[j];
                            }
                        }
                        store3(row, y);
                        store3(row + 3, tn);
                    }
                    load3(T2 + o, t);
                    if (kind == 0) apply_rot(S.M, t, tn);
                    else
                        for (int j = 0; j < 3; j++) tn[j] = t[j];
                    store3(row + 6, tn);
                    for (int m = 0; m < NB; m++)
                        row[9 + m] = (kind == 2 && m == binder) ? (double)newst[i]
                                                                : (double)ST[(ind0 + i) * NB + m];
                }
            }
        }
        if (lane == 0) {
            dbg->n_inds = n;
            dbg->dE_poly = dE_poly;
            dbg->dE_field = dE_field;
            dbg->passes = C.field_active ? S.passes : 0;
        }
        __syncwarp();
    }

    // ---- tangent_rotation move_funcs.pyx:470-582 --------------------------
    // per selected bead: own random axis, rotate t3/t2, both adjacent bonds
    // against the CURRENT neighbours (polymers.pyx:1075-1080; quirk 8); never
    // touches the field (mc_sim.pyx:145).
    // One bead is evaluated by 2 lanes -- its left bond and its right bond, each with the bead's trial
    // tangents and as it is -- up to 16 beads per call (v14: 4 lanes per bead, 8 beads per call: the typical
    // window of 14 needed two calls with a nearly empty second one).  The axis comes either from the prepared
    // record (`axfix`, the common single-bead case) or from per-bead draws.  New tangents of chunk bead j go
    // to out[6j..6j+5] (shared) when `out` is given, or straight to global memory when `store`.
    // Returns acc_in + the chunk's dE on all lanes (bead-by-bead sum).
    __device__ __forceinline__ double tangent_chunk(const int *beads, int cnt, const uint32_t *draws,
                                                    const double *axfix, double sn, double cs, double *out,
                                                    bool store, int dbg_base, double acc_in) {
        const double *Rr = R_();
        double *T3 = T3_(), *T2 = T2_();
        const int N = C.N;
        const int j = lane >> 1, right = lane & 1;
        double dside = 0.0;
        if (j < cnt) {
            const int bead = beads[j];
            double Rm[9], t3c[3], t2c[3], t3n[3], t2n[3], rc[3], axis[3];
            if (axfix) {
#pragma unroll
                for (int q = 0; q < 3; q++) axis[q] = axfix[q];
            } else {
                unit_sphere_point<!BATCH>(u01(draws[2 * j]), u01(draws[2 * j + 1]), axis);
            }
            rotation_3x3(axis, sn, cs, Rm);
            load3(T3 + 3 * bead, t3c);
            load3(T2 + 3 * bead, t2c);
            load3(Rr + 3 * bead, rc);
            apply_rot3(Rm, t3c, t3n);
            apply_rot3(Rm, t2c, t2n);
            const bool present = right ? (bead + 1 != N) : (bead != 0);
            if (!store && present) {
                // E(bond with the bead's trial tangents) - E(bond as it is); the bead is the second bead of
                // its left bond and the first of its right bond
                const int bond = right ? bead : bead - 1;
                const double e_trial = pair_energy(C, rep, bond, right ? 1 : 2, rc, t3n, t2n);
                const double e_cur = pair_energy(C, rep, bond, 0, rc, t3n, t2n);
                dside = e_trial - e_cur;
            }
            if (!right) {
                if (out) {
                    store3(out + 6 * j, t3n);
                    store3(out + 6 * j + 3, t2n);
                }
                if (store) {
                    store3(T3 + 3 * bead, t3n);
                    store3(T2 + 3 * bead, t2n);
                }
                if (DEBUG && !store && dbg_base + j < dbg->rows_cap) {
                    double *row = dbg->rows + (long long)(dbg_base + j) * (9 + NB);
                    store3(row, rc);
                    store3(row + 3, t3n);
                    store3(row + 6, t2n);
                    for (int m = 0; m < NB; m++) row[9 + m] = (double)ST_()[bead * NB + m];
                    if (dbg_base + j < dbg->inds_cap) dbg->inds[dbg_base + j] = bead;
                }
            }
        }
        // per bead: (0 + (E_left' - E_left)) + (E_right' - E_right); then bead by bead
        const double d = dside + __shfl_down_sync(FULL_MASK, dside, 1); // valid on the left-bond lanes
        double tot = acc_in;
        for (int b = 0; b < cnt; b++) tot += __shfl_sync(FULL_MASK, d, 2 * b);
        return tot;
    }

    // k > 1 beads, 16 per chunk: energies or stores.  The two axis draws of bead b are draws
    // `first + 2b`, `first + 2b + 1` of the attempt's stream: with the counter-based generator
    // (`par`) every bead's lane fetches its own, otherwise lane 0 draws them in order.
    __device__ __forceinline__ double tangent_eval_multi(const int *inds, int k, double sn, double cs,
                                                         bool small, bool store, bool par,
                                                         unsigned long long att, uint32_t first) {
        double dE_poly = 0.0;
        for (int base = 0; base < k; base += 16) {
            const int cnt = min(16, k - base);
            if (BATCH && par) {
                if (lane < cnt) {
                    Rng g = rng;
                    g.seek_attempt(att, first + 2u * (uint32_t)(base + lane));
                    S.draws[2 * lane] = g.next31();
                    S.draws[2 * lane + 1] = g.next31();
                }
            } else if (lane == 0) {
                for (int i = 0; i < 2 * cnt; i++) S.draws[i] = rng.next31();
            }
            __syncwarp();
            dE_poly = tangent_chunk(inds + base, cnt, S.draws, nullptr, sn, cs,
                                    (small && !store) ? S.tan_new + 6 * base : nullptr, store, base, dE_poly);
            __syncwarp();
        }
        return dE_poly;
    }

    // SimpleControl.update_move_amplitude mc_controller.py:148-213 (one thread)
    __device__ __forceinline__ void update_amplitudes(int mtype) {
        if (lane != 0 || wid != 0) return;
        chromo_move_state &mv = B.mv[mtype];
        if (mv.controller != 1) return;
        const double setpoint = 0.5, factor = 0.95;
        double a = mv.acceptance_rate;
        if (a < setpoint) {
            double prop = mv.amp_move * factor;
            if (prop > mv.move_amp_lo) mv.amp_move = prop;
            else {
                double nbd = (double)(mv.amp_bead - 1);
                mv.amp_move = mv.move_amp_hi;
                mv.amp_bead = (int)(mv.bead_amp_lo > nbd ? mv.bead_amp_lo : nbd);
            }
        } else if (a > setpoint) {
            double prop = mv.amp_move / factor;
            if (prop < mv.move_amp_hi) mv.amp_move = prop;
            else {
                double nbd = (double)(mv.amp_bead + 1);
                mv.amp_move = mv.move_amp_lo;
                mv.amp_bead = (int)(mv.bead_amp_hi < nbd ? mv.bead_amp_hi : nbd);
            }
        }
    }

    // ---- attempt `slot` of the batch (mc_step, mc_sim.pyx:106-182) ------------
    __device__ __forceinline__ void attempt(int mtype, int slot) {
        const int N = C.N;
        const Prop &P = B.prop[slot];
        const bool tangent = mtype == CHROMO_TANGENT_ROTATION;
        const int ind0 = P.ind0, n = P.n;
        const int binder = mtype == CHROMO_CHANGE_BINDING_STATE ? P.aux : 0;
        const int kind = mtype == CHROMO_SLIDE ? 1 : (mtype == CHROMO_CHANGE_BINDING_STATE ? 2 : 0);
        const unsigned long long att = abase + (unsigned long long)slot;
        if (n <= 0) { // mc_sim.pyx:151-152: counted, nothing else happens
            wait_turn(slot);
            if (lane == 0) B.mv[mtype].num_attempt += 1; // MCAdapter.propose moves.pyx:151
            pass_turn(slot, 0);
            return;
        }
        const signed char *newst = nullptr;
        const int *tinds = S.tinds;
        const bool small = n <= CB_TAN_SMALL;
        const bool presel = BATCH && n <= CB_KSEL; // tangent rotation: bead set drawn in prepare
        double dE_poly = 0.0, dE_field = 0.0;
        int2 conf = make_int2(0, 0);
        int ddbl[NB];
        // stage 1 may run ahead of the attempt's turn unless it needs sequential draws or the
        // replica-wide HBM scratch (large tangent / binding moves)
        bool my_turn = !(NW > 1 && (tangent ? presel : (kind != 2 || n <= CB_NEWST)));
        const bool segmove = !tangent && kind != 2;
        unsigned seen = 0u; // accepted attempts of the batch whose writes the prepared map / elastic dE reflect
        CB_T0();
        if (my_turn) wait_turn(slot);
        CB_LAP(2);
        while (true) {
            // ================= stage 1: bead rows only =================
#pragma unroll
            for (int m = 0; m < NB; m++) ddbl[m] = 0;
            if (!tangent) {
                if (segmove) {
                    // the map and the elastic dE were prepared with the batch; redo them if an attempt
                    // accepted since then wrote this segment's rows or its neighbours'
                    const unsigned now = accepted_before(slot);
                    if (rows_changed(now & ~seen, slot)) rows_recompute(mtype, slot);
                    seen = now;
                    if (lane < 12) S.M[lane] = P.M[lane];
                    dE_poly = P.dE_poly;
                } else if (lane == 0) {
                    if (BATCH) { // new states of this attempt (sequential mode drew them in prepare)
                        signed char *dst = (n <= CB_NEWST) ? S.newst : (C.st_new + (long long)rep * N);
                        if (n == 1) dst[0] = (signed char)P.newst0;
                        else {
                            rng.seek_attempt(att, P.used);
                            for (int i = 0; i < n; i++) dst[i] = (signed char)rng.randint(C.sites[binder] + 1);
                        }
                    }
                }
                __syncwarp();
                CB_LAP(12);
                if (kind == 2) newst = (n <= CB_NEWST) ? S.newst : (C.st_new + (long long)rep * N);
                if (kind == 2) dE_poly = binding_dE_poly(ind0, n, binder, newst, ddbl);
                CB_LAP(13);
                if (C.field_active) {
                    conf = field_scatter<NB>(C, H, S, rep, lane, kind, ind0, n, binder, newst);
                    CB_LAP(14);
                } else if (kind != 2 && C.confine_type != CHROMO_CONFINE_NONE)
                    dE_field = confinement_dE_segment(C, S, rep, lane, kind, ind0, n);
            } else if (n == 1) {
                if (lane == 0) S.tinds[0] = P.aux;
                __syncwarp();
                dE_poly = tangent_chunk(S.tinds, 1, nullptr, P.ax, P.sn, P.cs, S.tan_new, false, 0, 0.0);
            } else if (presel) {
                tinds = tsel(B, slot);
                dE_poly = tangent_eval_multi(tinds, n, P.sn, P.cs, true, false, true, att, P.used);
            } else {
                if (BATCH && lane == 0) rng.seek_attempt(att, P.used);
                if (small) { // get_inds move_funcs.pyx:552-582: k distinct draws, redraw duplicates
                    int my = -1;
                    for (int i = 0; i < n; i++) {
                        int c = 0;
                        bool dup;
                        do {
                            if (lane == 0) c = (int)(rng.next31() % (uint32_t)N);
                            c = __shfl_sync(FULL_MASK, c, 0);
                            dup = __any_sync(FULL_MASK, lane < i && my == c);
                        } while (dup);
                        if (lane == i) my = c;
                    }
                    if (lane < n) S.tinds[lane] = my;
                } else {
                    tinds = tangent_select_large(n);
                }
                __syncwarp();
                dE_poly = tangent_eval_multi(tinds, n, P.sn, P.cs, small, false, false, att, 0u);
            }
            CB_LAP(1);
            if (my_turn) break;
            wait_turn(slot);
            CB_LAP(2);
            my_turn = true;
            if (segmove ? !rows_changed(accepted_before(slot) & ~seen, slot) : !stale(mtype, slot)) break;
            if (!tangent && C.field_active) table_clear(H, S, NCOL, lane); // rows changed under stage 1: redo it
            CB_LAP(3);
        }
        CB_LAP(3);
        // ================= stage 2: density, Metropolis, commit =================
        if (!tangent) {
            if (C.field_active)
                dE_field = field_finish<NB, DEBUG>(C, H, S, rep, lane, kind, ind0, n, binder, newst, conf, ddbl, dbg);
            if (DEBUG) debug_report(kind, ind0, n, binder, newst, dE_poly, dE_field);
        } else if (DEBUG) {
            if (lane == 0) {
                dbg->n_inds = n;
                dbg->dE_poly = dE_poly;
                dbg->dE_field = 0.0;
                dbg->n_touched = 0;
                dbg->passes = 0;
            }
            __syncwarp();
        }
        CB_LAP(4);
        // Metropolis (mc_sim.pyx:163-171)
        double dE = 0.0;
        dE += dE_poly;
        if (!tangent && has_field()) dE += dE_field;
        int acc = 0;
        if (lane == 0) {
            double u = __longlong_as_double(0x7ff8000000000000LL);
            if (DEBUG && force_accept >= 0) {
                acc = force_accept;
            } else {
                u = BATCH ? P.u : u01(rng.next31());
                // mc_sim.pyx:163-171: accept iff u < exp(-dE).  dE < 0: exp(-dE) > 1 >= u whatever u is -- the same
                // decision without the exponential (about half of the attempts)
                acc = (dE < 0.0) ? 1 : ((u < exp(-dE)) ? 1 : 0);
            }
            if (DEBUG) {
                dbg->u = u;
                dbg->accepted = acc;
            }
            // counters + AcceptanceTracker.update_acceptance_rate mc_stat.py:190-207
            chromo_move_state &mv = B.mv[mtype];
            mv.num_attempt += 1; // MCAdapter.propose moves.pyx:151
            if (acc) mv.num_success += 1;
            mv.acceptance_rate = (mv.alpha * (acc ? 1.0 : 0.0)) + (1.0 - mv.alpha) * mv.acceptance_rate;
            // algorithmic bytes of this attempt
            unsigned long long ab;
            if (tangent) {
                ab = 48ull * n + 144ull * n + (acc ? 48ull * n : 0ull);
            } else {
                unsigned long long U = C.field_active ? (unsigned long long)S.last_U : 0ull;
                if (kind == 2)
                    ab = 24ull * n + (unsigned long long)NB * n * (acc ? 2ull : 1ull) +
                         8ull * NCOL * U * (acc ? 3ull : 1ull);
                else
                    ab = 72ull * (n + 2) + (acc ? 72ull * n : 0ull) + (unsigned long long)NB * n +
                         8ull * NCOL * U * (acc ? 3ull : 1ull) + 80ull;
            }
            B.algo_bytes += ab;
        }
        acc = __shfl_sync(FULL_MASK, acc, 0);
        CB_LAP(5);
        // accept (moves.pyx:156-239)
        if (acc) {
            if (!tangent) segment_commit(kind, ind0, n, binder, newst);
            else if (small) {
                if (lane < n) {
                    store3(T3_() + 3 * tinds[lane], S.tan_new + 6 * lane);
                    store3(T2_() + 3 * tinds[lane], S.tan_new + 6 * lane + 3);
                }
            } else {
                tangent_commit_large(tinds, n, P.sn, P.cs);
            }
        }
        CB_LAP(6);
        pass_turn(slot, acc);
        if (!tangent && C.field_active) table_clear(H, S, NCOL, lane);
        __syncwarp();
        CB_LAP(7);
#ifdef CB_PHASE_TIMERS
        tacc[11] += 1;
        tacc[15] += (unsigned long long)n;
#endif
    }

    // a batch of `cnt` attempts of one move type: prepare (lanes of warp 0), then the warps take
    // the attempts round-robin
    __device__ __forceinline__ void run(int mtype, int cnt) {
        CB_T0();
        if (wid == 0) {
            prepare(mtype, cnt);
            if (lane == 0) {
                *(volatile int *)&B.token = 0;
                *(volatile unsigned *)&B.accepted = 0u;
            }
        }
        CB_LAP(8);
        block_sync();
        CB_LAP(9);
#pragma unroll 1
        for (int j = wid; j < cnt; j += NW) {
            CB_T0();
            CB_LAP(0);
            attempt(mtype, j);
        }
        abase += (unsigned long long)cnt;
        {
            CB_T0();
            block_sync();
            CB_LAP(9);
        }
    }

    // tangent rotation of more than CB_TAN_SMALL beads (rare): indices in HBM scratch,
    // membership in a bitmap, per-bead draws regenerated on commit
    __device__ __forceinline__ const int *tangent_select_large(int k) {
        const int N = C.N;
        int *inds = C.tan_inds + (long long)rep * N;
        uint32_t *bits = C.sel_bits + (long long)rep * ((N + 31) / 32);
        for (int i = lane; i < (N + 31) / 32; i += 32) bits[i] = 0u;
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < k; i++) {
                int c;
                do {
                    c = (int)(rng.next31() % (uint32_t)N);
                } while ((bits[c >> 5] >> (c & 31)) & 1u);
                bits[c >> 5] |= 1u << (c & 31);
                inds[i] = c;
            }
            rng.save(B.rng_save);
        }
        __syncwarp();
        return inds;
    }
    __device__ __forceinline__ void tangent_commit_large(const int *inds, int k, double sn, double cs) {
        // a bead's new tangents depend on its own t3/t2 only, so regenerating the
        // per-bead draws from the saved RNG state and storing chunk by chunk is safe
        if (lane == 0) {
            rng.save(B.rng_after);
            rng.restore(B.rng_save);
        }
        (void)tangent_eval_multi(inds, k, sn, cs, false, true, false, 0ull, 0u);
        if (lane == 0) rng.restore(B.rng_after);
    }
};

// ------------------------------------------------------------------ kernels
__device__ __forceinline__ HashTable carve_table(unsigned char *dyn, int cap, int ncol) {
    HashTable H;
    H.vals = (uint32_t *)dyn;
    H.vals_s = cb_shared_addr(dyn);
    H.keys = (int *)(dyn + (size_t)cap * ncol * 8);
    H.list = H.keys + cap;
    H.cap = cap;
    H.limit = cap - cap / 4 - 32;
    return H;
}

template <class Rng>
__device__ __forceinline__ unsigned long long rng_load(Rng &rng, const DevCtx &C, ReplicaSh &B, int rep, int tid,
                                                       unsigned long long seed);
template <>
__device__ __forceinline__ unsigned long long rng_load<ReplayRng>(ReplayRng &rng, const DevCtx &C, ReplicaSh &B,
                                                                  int rep, int tid, unsigned long long) {
    for (int i = tid; i < CB_GLIBC_WORDS; i += 32) B.grs[i] = C.glibc[(long long)rep * CB_GLIBC_WORDS + i];
    rng.st = B.grs;
    rng.mt = C.mt + (long long)rep * CB_MT_WORDS;
    return 0ull;
}
template <>
__device__ __forceinline__ unsigned long long rng_load<PhiloxRng>(PhiloxRng &rng, const DevCtx &C, ReplicaSh &,
                                                                  int rep, int, unsigned long long seed) {
    rng.k0 = (uint32_t)seed;
    rng.k1 = (uint32_t)(seed >> 32);
    rng.rep = C.rep_offset + (uint32_t)rep; // global replica index: shards of an ensemble draw different streams
    rng.seek_attempt(C.philox_ctr[rep]);
    return C.philox_ctr[rep]; // attempts this replica has ever made
}
template <class Rng>
__device__ __forceinline__ void rng_store(const DevCtx &C, ReplicaSh &B, int rep, int tid, unsigned long long abase);
template <>
__device__ __forceinline__ void rng_store<ReplayRng>(const DevCtx &C, ReplicaSh &B, int rep, int tid,
                                                     unsigned long long) {
    for (int i = tid; i < CB_GLIBC_WORDS; i += 32) C.glibc[(long long)rep * CB_GLIBC_WORDS + i] = B.grs[i];
}
template <>
__device__ __forceinline__ void rng_store<PhiloxRng>(const DevCtx &C, ReplicaSh &, int rep, int tid,
                                                     unsigned long long abase) {
    if (tid == 0) C.philox_ctr[rep] = abase;
}

// mc_sim mc_sim.pyx:26-103 for every replica.  One thread block per SM holds `rpb`
// replicas (NW warps each); grid = ceil(R / rpb).  The replicas of a block are independent
// simulations, but they go through the move types of a sweep TOGETHER (a block-wide
// barrier after each): the kernel is bound by instruction fetch (140 KB of SASS against a
// 32 KB L1.5 instruction cache), and warps that run the same move type share its code.
template <class Rng, int NB, int NW>
__global__ void __launch_bounds__(32 * NW * CB_MAX_RPB, CB_MIN_BLOCKS)
    mc_sim_kernel(const CB_GRID_CONSTANT DevCtx C, long long num_mc_steps, double mu_adjust,
                  unsigned long long seed, int cap, int rpb, int rep0, int rep_end) {
    CB_DYN_SMEM(dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, local = warp / NW, wid = warp % NW;
    // replicas [rep0, rep_end) of the context: the host-array path launches one grid per replica chunk
    const int rep = rep0 + blockIdx.x * rpb + local, rtid = wid * 32 + lane; // thread index within the replica
    const bool active = rep < rep_end; // the last block may hold fewer replicas; its spare warps only keep the barriers
    unsigned char *base = dyn + (size_t)local * cb_replica_smem(cap, C.ncol, NW);
    ReplicaSh &B = *(ReplicaSh *)base;
    WarpSh &S = *(WarpSh *)(base + CB_REPLICA_SH_BYTES + (size_t)wid * CB_WARP_SH_BYTES);
    HashTable H = carve_table(base + CB_REPLICA_SH_BYTES + (size_t)NW * CB_WARP_SH_BYTES +
                                  (size_t)wid * cb_table_bytes(cap, C.ncol), cap, C.ncol);
    H.prefetch = NW > 1;
    Rng rng;
    unsigned long long abase0 = 0;
    if (active) {
        table_reset_all(H, S, C.ncol, lane);
        if (rtid < CHROMO_NUM_MOVES) B.mv[rtid] = C.moves[(long long)rep * CHROMO_NUM_MOVES + rtid];
        if (rtid == 0) B.algo_bytes = 0;
        if (lane == 0) {
            S.last_U = 0;
            S.passes = 1;
        }
        abase0 = rng_load<Rng>(rng, C, B, rep, rtid, seed);
    }
    McWarp<Rng, false, NB, NW> W{C, B, S, H, rng, rep, lane, wid, local, mu_adjust, -1, nullptr, abase0};
    __syncthreads();
    long long a0 = 0;
    if (active)
        for (int m = 0; m < CHROMO_NUM_MOVES; m++) a0 += B.mv[m].num_attempt;
    const int BS = Rng::kBatched ? C.batch : 1;
    for (long long k = 0; k < num_mc_steps; k++)
        for (int m = 0; m < CHROMO_NUM_MOVES; m++) {
#ifdef CB_NO_TYPE_MASK // (A/B knob: what the per-type mask costs)
            if (active) {
#else
            if (active && (C.type_mask >> m & 1)) {
#endif
                if (B.mv[m].move_on == 1) {
                    const int npc = B.mv[m].num_per_cycle;
#pragma unroll 1
                    for (int j0 = 0; j0 < npc; j0 += BS) W.run(m, min(BS, npc - j0));
                }
                W.update_amplitudes(m); // also for moves that are off (mc_sim.pyx:103)
            }
#ifdef CB_PHASE_TIMERS
            {
                long long t0_ = clock64();
                __syncthreads();
                W.tacc[10] += (unsigned long long)(clock64() - t0_);
                if (active) W.flush_timers(m);
            }
#endif
            // The block's replicas enter the next move type together.  (Measured on B200: skipping
            // the barrier before the short move types -- 1 end-pivot, 10 binding attempts -- costs
            // more in lost instruction-cache sharing than the wait for the slowest replica does.)
#if defined(CB_PAIR_BARRIER)
            // experiment: only the warps that share an SM sub-partition (warp w and w + 4) wait for each other
            if (NW == 1 && (int)blockDim.x == 32 * 7) {
                if ((warp ^ 4) < 7) cb_bar_sync(1 + (warp & 3), 64);
            } else __syncthreads();
#elif defined(CB_SKIP_BARRIER_AFTER)
            if (m != CB_SKIP_BARRIER_AFTER) __syncthreads();
#elif !defined(CB_NO_TYPE_BARRIER)
            __syncthreads();
#endif
        }
    if (!active) return;
    long long a1 = 0;
    for (int m = 0; m < CHROMO_NUM_MOVES; m++) a1 += B.mv[m].num_attempt;
    if (rtid == 0) {
        C.attempts[rep] = (C.accumulate ? C.attempts[rep] : 0ull) + (unsigned long long)(a1 - a0);
        C.algo_bytes[rep] = (C.accumulate ? C.algo_bytes[rep] : 0ull) + B.algo_bytes;
    }
    if (rtid < CHROMO_NUM_MOVES) C.moves[(long long)rep * CHROMO_NUM_MOVES + rtid] = B.mv[rtid];
    rng_store<Rng>(C, B, rep, rtid, W.abase);
}

// one instrumented mc_step of one replica (chromo_mc_step)
template <class Rng, int NB>
__global__ void __launch_bounds__(32) mc_step_kernel(const CB_GRID_CONSTANT DevCtx C, int rep, int mtype,
                                                     double amp_move, int amp_bead, double mu_adjust,
                                                     unsigned long long seed, int force_accept,
                                                     DebugOut *dbg, int cap) {
    CB_DYN_SMEM(dyn);
    __shared__ ReplicaSh B;
    __shared__ WarpSh S;
    const int lane = threadIdx.x;
    HashTable H = carve_table(dyn, cap, C.ncol);
    H.prefetch = false;
    table_reset_all(H, S, C.ncol, lane);
    if (lane < CHROMO_NUM_MOVES) B.mv[lane] = C.moves[(long long)rep * CHROMO_NUM_MOVES + lane];
    __syncwarp();
    if (lane == 0) {
        B.mv[mtype].amp_move = amp_move;
        B.mv[mtype].amp_bead = amp_bead;
        dbg->n_inds = 0;
        dbg->n_touched = 0;
        dbg->dE_poly = dbg->dE_field = 0.0;
        dbg->accepted = 0;
        dbg->passes = 0;
        B.algo_bytes = 0;
        S.last_U = 0;
        S.passes = 1;
    }
    Rng rng;
    const unsigned long long abase0 = rng_load<Rng>(rng, C, B, rep, lane, seed);
    __syncwarp();
    McWarp<Rng, true, NB, 1> W{C, B, S, H, rng, rep, lane, 0, 0, mu_adjust, force_accept, dbg, abase0};
    W.run(mtype, 1);
    __syncwarp();
    rng_store<Rng>(C, B, rep, lane, W.abase);
}
