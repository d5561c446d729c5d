// field_kernels.cuh -- whole-replica kernels (A8): full density binning,
// total field energy, total elastic energy, the chi-conjugate observable.
// These are the HBM-streaming kernels of the path: every bead / voxel is read
// once, coalesced; per-replica results come from fixed-order block partials so
// they are bit-reproducible run to run.
#pragma once
#include "geometry.cuh"
#include "rng.cuh"
#include "launch.cuh"
#include "params.cuh"

#define FK_THREADS 256

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sh[i];
    return t; // valid on thread 0
}

// update_all_densities fields.pyx:1977-2039: scatter every bead into its 8
// voxels (rho += w / V_access * {1, state}).  One thread per bead; a warp's
// 32 beads are one contiguous 768-byte run of r.  Accumulation is a native
// fp64 reduction at L2 (RED.E.ADD.F64); the grid was zeroed by the caller.
template <int NB>
__global__ void __launch_bounds__(FK_THREADS) density_scatter_kernel(DevCtx C) {
    constexpr int NCOL = NB + 1;
    const int rep = blockIdx.y;
    const int bead = blockIdx.x * blockDim.x + threadIdx.x;
    if (bead >= C.N) return;
    const double *p = C.r + ((long long)rep * C.N + bead) * 3;
    const signed char *st = C.states + ((long long)rep * C.N + bead) * NB;
    double *dens = C.density + (long long)rep * C.n_bins * NCOL;
    int idx[8];
    double w[8];
    bin_point(C, p[0], p[1], p[2], idx, w);
    double s[NB];
#pragma unroll
    for (int m = 0; m < NB; m++) s[m] = (double)st[m];
#pragma unroll
    for (int l = 0; l < 8; l++) {
        double V = C.access_vol ? C.access_vol[idx[l]] : C.vol_bin;
        double d = w[l] / V;
        double *row = dens + (long long)idx[l] * NCOL;
        atomicAdd(row, d);
#pragma unroll
        for (int m = 0; m < NB; m++)
            if (s[m] != 0.0) atomicAdd(row + 1 + m, d * s[m]);
    }
}

// The same for grids that fit one block's shared memory (16 bytes per voxel: C1-C3; the C2 grid is 148 KB):
// one block per (replica, density column) accumulates its column PRIVATELY in shared memory and writes every
// voxel exactly once -- no memset of the grid, no atomics at L2, no separate clamp pass.  Each term w / V (x
// state) is added as an 80-bit fixed-point number in units of 2^-E (E = 60 + floor(log2 V_min), quantum ~5e-23
// nm^-3: below the smallest non-zero term a double weight can produce, 1e-16 / V), split over four 32-bit words
// that collect 16 bits each, so a term is four fire-and-forget native shared atomics with no carries (folded
// once per 65,536 beads).  Integer sums do not depend on the order in which beads arrive: the densities are
// bit-reproducible run to run (the L2-atomic version is order-dependent in the last ulp), and each voxel holds
// the correctly rounded value of its exact fixed-point sum.
#define FK_PRIV_THREADS 1024
template <int NB>
__global__ void __launch_bounds__(FK_PRIV_THREADS, 1)
    density_private_kernel(DevCtx C, int clamp, int fx_e, int pbits, int fold_beads, int cols_per_block) {
    constexpr int NCOL = NB + 1;
    CB_DYN_SMEM(dyn);
    // cells[column][word][bin], word-major: the 32 lanes' adds to one word of their voxels fall on banks bin % 32
    // (with a voxel's words side by side one instruction only ever touched a quarter of the banks)
    uint32_t *cells = (uint32_t *)dyn;
    const int nbw = C.n_bins;
    const cb_saddr cells_s = cb_shared_addr(dyn);
    const int rep = blockIdx.x, tid = threadIdx.x;
    const int col0 = blockIdx.y * cols_per_block, ncb = min(cols_per_block, NCOL - col0); // this block's columns
    const uint32_t pmask = (1u << pbits) - 1u;
    const double scale = __hiloint2double((1023 + fx_e) << 20, 0), inv_scale = __hiloint2double((1023 - fx_e) << 20, 0);
    for (int i = tid; i < nbw * 3 * ncb; i += blockDim.x) cells[i] = 0u;
    __syncthreads();
    const double *R = C.r + (long long)rep * C.N * 3;
    const signed char *ST = C.states + (long long)rep * C.N * NB;
    for (int chunk = 0; chunk < C.N; chunk += fold_beads) {
        const int end = min(C.N, chunk + fold_beads);
        // blocked assignment: a thread walks k CONSECUTIVE beads, so the lanes of a warp are k beads (~6 voxels
        // at C2) apart and rarely meet in a voxel -- with one bead per lane neighbouring lanes hit the same
        // cells and the shared-memory atomics serialise 8-fold (ncu: 7.6 wavefronts per instruction)
        const int k = (end - chunk + (int)blockDim.x - 1) / (int)blockDim.x;
        for (int j = 0; j < k; j++) {
            const int bead = chunk + tid * k + j;
            if (bead >= end) break;
            unsigned long long mult[NCOL]; // column c of this block receives (w / V) * mult[c]
            bool any = false;
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                const int col = col0 + c;
                mult[c] = (c >= ncb) ? 0ull : (col == 0 ? 1ull : (unsigned long long)ST[bead * NB + col - 1]);
                any = any || mult[c] != 0ull;
            }
            if (!any) continue;
            int idx[8];
            double w[8];
            bin_point(C, R[3 * bead], R[3 * bead + 1], R[3 * bead + 2], idx, w);
#pragma unroll
            for (int l = 0; l < 8; l++) {
                const double d = C.access_vol ? w[l] / C.access_vol[idx[l]] : div_const(w[l], C.vol_bin, C.inv_vol_bin);
                const unsigned long long t1 = (unsigned long long)__double2ll_rn(d * scale);
#pragma unroll
                for (int c = 0; c < NCOL; c++) {
                    if (mult[c] == 0ull) continue;
                    const unsigned long long t = t1 * mult[c];
                    const cb_saddr a = cells_s + 4u * (cb_saddr)(c * 3 * nbw + idx[l]);
                    cb_red_add_u32(a, (uint32_t)t & pmask);
                    cb_red_add_u32(a + 4 * nbw, (uint32_t)(t >> pbits) & pmask);
                    cb_red_add_u32(a + 8 * nbw, (uint32_t)(t >> (2 * pbits)));
                }
            }
        }
        __syncthreads();
        if (end < C.N) { // more beads to come: fold the carries so that the payload lanes start empty again
            for (int i = tid; i < nbw * ncb; i += blockDim.x) {
                uint32_t *c = cells + (i / nbw) * 3 * nbw + (i % nbw);
                uint32_t w0 = c[0], w1 = c[nbw], w2 = c[2 * nbw];
                w1 += w0 >> pbits, w0 &= pmask;
                w2 += w1 >> pbits, w1 &= pmask;
                c[0] = w0, c[nbw] = w1, c[2 * nbw] = w2;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < nbw * ncb; i += blockDim.x) {
        const int c = i / nbw, bin = i % nbw;
        const uint32_t *q = cells + c * 3 * nbw + bin;
        // exact value = (w2 2^(2p) + w1 2^p + w0) 2^-E, written as B 2^32 + A with A < 2^32 and B < 2^53: two
        // exactly representable halves, one rounding
        const unsigned long long low = ((unsigned long long)q[nbw] << pbits) + (unsigned long long)q[0];
        const unsigned long long A = low & 0xFFFFFFFFull;
        const unsigned long long B = (low >> 32) + ((unsigned long long)q[2 * nbw] << (2 * pbits - 32));
        double v = ((double)B * 4294967296.0 + (double)A) * inv_scale;
        if (clamp && fabs(v) < 1E-18) v = 0.0; // update_all_densities_for_all_polymers fields.pyx:2101-2105
        C.density[((long long)rep * nbw + bin) * NCOL + col0 + c] = v;
    }
}

// the |rho| < 1e-18 -> 0 pass of update_all_densities_for_all_polymers
// (fields.pyx:2101-2105)
__global__ void density_clamp_kernel(double *d, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && fabs(d[i]) < 1E-18) d[i] = 0.0;
}

// Python's round(x, 2) > limit, exactly: round-half-even of the EXACT decimal
// value of x (float.__round__), compared through the smallest hundredth K
// whose double K/100 exceeds the limit (K computed on the host).
__device__ __forceinline__ bool round2_exceeds(double x, double K) {
    double T = K - 0.5;
    double p = 100.0 * x;
    if (p > T) return true;
    if (p < T) return false;
    if (!(p == T)) return false;     // NaN
    double e = fma(100.0, x, -p);    // exact rounding error of p
    if (e > 0.0) return true;
    if (e < 0.0) return false;
    return fmod(K, 2.0) == 0.0;      // exact tie: to even
}

// get_E_binders_and_beads + nonspecific_interact_E fields.pyx:2208-2315.
// partial[rep][blk][0..NB-1] = sum rho_a^2, [NB] = nonspecific, over the
// block's voxels;  dcount[rep][a] += # beads with state == 2.
template <int NB>
__global__ void __launch_bounds__(FK_THREADS) field_energy_kernel(DevCtx C, double roundK,
                                                                  double *partial, int *dcount,
                                                                  int chi_observable) {
    constexpr int NCOL = NB + 1;
    __shared__ double sh[FK_THREADS / 32];
    const int rep = blockIdx.y;
    const double *dens = C.density + (long long)rep * C.n_bins * NCOL;
    const double chi = C.chi[rep];
    double sq[NB], ns = 0.0;
#pragma unroll
    for (int a = 0; a < NB; a++) sq[a] = 0.0;
    for (int bin = blockIdx.x * blockDim.x + threadIdx.x; bin < C.n_bins; bin += gridDim.x * blockDim.x) {
        const double *row = dens + (long long)bin * NCOL;
#pragma unroll
        for (int a = 0; a < NB; a++) sq[a] += row[a + 1] * row[a + 1];
        double V = C.access_vol ? C.access_vol[bin] : C.vol_bin;
        double vf = row[0] * C.bead_vol;
        if (chi_observable) ns += (V / C.bead_vol) * (vf * vf);
        else if (round2_exceeds(vf, roundK)) ns += CB_E_HUGE_FIELD * vf;
        else ns += chi * (V / C.bead_vol) * vf * (1.0 - vf);
    }
    double *out = partial + ((long long)rep * gridDim.x + blockIdx.x) * NCOL;
#pragma unroll
    for (int a = 0; a < NB; a++) {
        double t = block_sum(sq[a], sh);
        if (threadIdx.x == 0) out[a] = t;
    }
    double t = block_sum(ns, sh);
    if (threadIdx.x == 0) out[NB] = t;
    if (!chi_observable) {
        const signed char *st = C.states + (long long)rep * C.N * NB;
        int cnt[NB];
#pragma unroll
        for (int a = 0; a < NB; a++) cnt[a] = 0;
        for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < C.N; b += gridDim.x * blockDim.x)
#pragma unroll
            for (int a = 0; a < NB; a++) cnt[a] += (st[(long long)b * NB + a] == 2);
#pragma unroll
        for (int a = 0; a < NB; a++) {
            int v = cnt[a];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&dcount[rep * NB + a], v);
        }
    }
}

// SSWLC.compute_E polymers.pyx:1348-1381: one thread per bond, fixed-order
// block partials.  Note the reference builds `bend` here as
// t3_1 + (-t3_0 - eta*dr_perp), not as in the dE routines.
__global__ void __launch_bounds__(FK_THREADS) elastic_energy_kernel(DevCtx C, double *partial) {
    __shared__ double sh[FK_THREADS / 32];
    const int rep = blockIdx.y;
    const double *Rr = C.r + (long long)rep * C.N * 3;
    const double *T3 = C.t3 + (long long)rep * C.N * 3;
    double e = 0.0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < C.N - 1; b += gridDim.x * blockDim.x) {
        double r0[3], r1[3], t0[3], t1[3], dr[3], perp[3], bend[3];
        load3(Rr + 3 * (long long)b, r0);
        load3(Rr + 3 * (long long)(b + 1), r1);
        load3(T3 + 3 * (long long)b, t0);
        load3(T3 + 3 * (long long)(b + 1), t1);
        Bond B = load_bond(C, rep, b);
        for (int i = 0; i < 3; i++) dr[i] = r1[i] - r0[i];
        double par = dot3(t0, dr);
        for (int i = 0; i < 3; i++) {
            perp[i] = dr[i] - t0[i] * par;
            bend[i] = t1[i] + (-t0[i] - B.eta * perp[i]);
        }
        double tp = par - B.gamma;
        double eb = (0.5 * B.eps_bend * dot3(bend, bend) + 0.5 * B.eps_par * (tp * tp)) +
                    0.5 * B.eps_perp * dot3(perp, perp);
        if (C.twist) { // SSTWLC.compute_E polymers.pyx:2250-2285
            double u0[3], u1[3];
            const double *T2 = C.t2 + (long long)rep * C.N * 3;
            load3(T2 + 3 * (long long)b, u0);
            load3(T2 + 3 * (long long)(b + 1), u1);
            eb += twist_energy(C.twist + (long long)rep * C.twist_stride + 2 * (long long)b, twist_omega(u0, t0, u1, t1));
        }
        e += eb;
    }
    double t = block_sum(e, sh);
    if (threadIdx.x == 0) partial[(long long)rep * gridDim.x + blockIdx.x] = t;
}

// out[rep][c] = sum_blk partial[rep][blk][c], in block order
__global__ void partial_finish_kernel(const double *partial, double *out, int R, int nblk, int ncomp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * ncomp) return;
    int rep = i / ncomp, c = i % ncomp;
    double t = 0.0;
    for (int b = 0; b < nblk; b++) t += partial[((long long)rep * nblk + b) * ncomp + c];
    out[i] = t;
}

// int64 host layout <-> int8 device layout for states / chemical_mods
__global__ void widen_i8_kernel(const signed char *src, long long *dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// Values index the binding free-energy table (binding_dE_poly): anything outside [0, hi] (hi = the largest
// sites_per_bead) is clamped and reported through *bad, and the call that uploaded it fails with
// CHROMO_ERR_ARG instead of reading out of bounds on the device.
__global__ void narrow_i64_kernel(const long long *src, signed char *dst, long long n, int hi, int *bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        long long v = src[i];
        if (v < 0 || v > hi) {
            *bad = 1;
            v = v < 0 ? 0 : hi;
        }
        dst[i] = (signed char)v;
    }
}

// ------------------------------------------------------------ replica exchange
// column `col` of the per-replica reduction results -> a caller-owned device buffer
__global__ void pick_column_kernel(const double *out, double *dst, int R, int ncol, int col) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) dst[r] = out[(size_t)r * ncol + col];
}

// One even / odd round of neighbour swaps on a chi ladder, decided identically on every rank from the
// all-gathered observable Phi_g = sum_bins (V/v) phi^2 (global replica id g).  The Hamiltonian sampled by the
// moves is H0 + chi Phi (nonspecific_interact_dE fields.pyx:1829-1840), so exchanging the chi LABELS of the
// replicas on rungs k, k+1 changes the total energy by (chi_k - chi_{k+1}) (Phi_b - Phi_a); it is accepted
// with probability min(1, exp(-dE)).  ladder[k] is the k-th smallest chi, rung_replica[k] the replica that
// holds it; rungs are grouped in independent ladders of `ladder_len` (no pair across a boundary).  One
// thread per pair: pairs are disjoint, so the permutation is updated in place.  The uniform of pair p in
// round t is word 0 of Philox4x32-10(counter = (p, t_lo, t_hi, 'EXCH'), key = seed): the same on every rank.
__global__ void exchange_kernel(const double *phi_all, const double *ladder, int *rung_replica, double *chi_local,
                                long long n_total, long long ladder_len, long long first, int R, long long round,
                                unsigned long long seed, unsigned long long *counters) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long k = (round & 1) + 2 * p;
    if (k + 1 >= n_total) return;
    if ((k + 1) % ladder_len == 0) return; // rungs k and k+1 belong to different ladders
    const int a = rung_replica[k], b = rung_replica[k + 1];
    const double dE = (ladder[k] - ladder[k + 1]) * (phi_all[b] - phi_all[a]);
    uint32_t o[4];
    philox4x32_10((uint32_t)p, (uint32_t)round, (uint32_t)((unsigned long long)round >> 32), 0x45584348u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double u = (double)(o[0] >> 1) / CB_RAND_MAX;
    atomicAdd(&counters[0], 1ull);
    if (u < exp(-dE)) {
        rung_replica[k] = b;
        rung_replica[k + 1] = a;
        if (b >= first && b < first + R) chi_local[b - first] = ladder[k];
        if (a >= first && a < first + R) chi_local[a - first] = ladder[k + 1];
        atomicAdd(&counters[1], 1ull);
    }
}
