// rediscretize.cu -- the coarse-grain / refine steps either side of the MC path,
// batched over replicas.  The reference does these one polymer at a time in Python loops
// (chromo/util/rediscretize.py); here one launch handles every replica of an ensemble.
//
//   cg_reduce_kernel      get_cg_chromatin            rediscretize.py:401-471
//                         (get_cg_bead_intervals 24-54, get_avg_in_intervals 57-84,
//                          get_orientations_in_intervals 87-122, get_majority_state_in_interval 125-160)
//   refine_path_kernel    get_refined_path            rediscretize.py:756-807
//                         (get_refined_intervals 537-583, brownian_bridge 586-685,
//                          gaussian_walk util/poly_paths.py:269-295,
//                          get_refined_orientations 810-842)
//   confine_kernel        enforce_spherical_confinement rediscretize.py:708-753
//
// All three are streaming kernels: inputs are staged through shared memory with coalesced loads,
// each output element is written once (measured 27-53 % of the HBM copy peak for the coarse-graining
// and the confinement; the refinement is bound by its fp64 arithmetic).
// Arithmetic follows numpy's evaluation
// order (sequential row sums for axis-0 means, literal cross products, IEEE divisions; the library
// is compiled with -fmad=false) so that interval means, orientations and majority states are
// bit-identical to the reference's.
#include "launch.cuh"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>

#include "../../include/chromo_b200.h"
#include "rng.cuh"

int cb_set_error(int code, const char *fmt, ...); // chromo_b200.cu

#define RCK(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            rc = cb_set_error(CHROMO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                              __FILE__, __LINE__);                                                \
            goto done;                                                                            \
        }                                                                                         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// coarse-graining
// ------------------------------------------------------------------------------------------------
#define CG_THREADS 256
#define CG_MAX_VALUE 16 // states / marks are small non-negative integers (sites_per_bead + 1 values)

static inline __host__ __device__ int cg_row_stride(int k, int width) {
    // row of one interval in shared memory, padded to an odd number of 64-bit words so that the
    // per-interval threads of a warp hit different banks
    int s = k * width;
    return (s & 1) ? s : s + 1;
}

struct CgArgs {
    long long R, N, nb, k, M;
    double r_div; // cg_factor ** (1/3)
    const double *r, *t3;
    const long long *states, *mods;
    double *r_cg, *t3_cg, *t2_cg;
    long long *states_cg, *mods_cg;
    int *err;
    int T; // intervals per block
};

// stage rows [b0, b1) x width of one replica into padded per-interval rows
template <class V>
__device__ __forceinline__ void cg_stage(V *sh, const V *src, long long b0, long long b1, int width, int k,
                                         int stride) {
    const int n = (int)((b1 - b0) * width); // a block's rows fit shared memory: well below 2^31
    const int kw = k * width, nt = (int)blockDim.x;
    const int pad = stride - kw;          // 0 or 1 padding word per interval row
    const int dg = nt / kw, dr = nt % kw; // (interval, offset) advance of a thread from one load to its next
    src += b0 * width;
    int i = threadIdx.x, g = i / kw, o = i % kw; // element i = row g, offset o; kept incrementally (no division per load)
    for (; i + 7 * nt < n; i += 8 * nt) { // eight loads in flight per thread
        V v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = src[i + u * nt];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            sh[i + u * nt + pad * g] = v[u];
            g += dg;
            o += dr;
            if (o >= kw) {
                o -= kw;
                g++;
            }
        }
    }
    for (; i < n; i += nt) {
        sh[i + pad * g] = src[i];
        g += dg;
        o += dr;
        if (o >= kw) {
            o -= kw;
            g++;
        }
    }
}

// mean of the rows of one interval, numpy order: ((x0 + x1) + x2) + ... then one division by the count
// (np.average(axis=0) -> add.reduce over the outer axis, rediscretize.py:81-83)
__device__ __forceinline__ void cg_mean3(const double *row, int cnt, double m[3]) {
    double sx = row[0], sy = row[1], sz = row[2];
    for (int j = 1; j < cnt; j++) {
        sx = sx + row[3 * j];
        sy = sy + row[3 * j + 1];
        sz = sz + row[3 * j + 2];
    }
    const double c = (double)cnt;
    m[0] = sx / c;
    m[1] = sy / c;
    m[2] = sz / c;
}

// t3 / |t3| and t2 = t3 x e_x (or t3 x e_y when t3 == e_x exactly), normalised
// (rediscretize.py:107-121 and 829-841; np.cross evaluates a1*b2 - a2*b1 etc. literally)
__device__ __forceinline__ void cg_orient(double a[3], double t2[3]) {
    const double mag = sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
    a[0] = a[0] / mag;
    a[1] = a[1] / mag;
    a[2] = a[2] / mag;
    double b0 = 1.0, b1 = 0.0, b2 = 0.0;
    if (a[0] == 1.0 && a[1] == 0.0 && a[2] == 0.0) {
        b0 = 0.0;
        b1 = 1.0;
    }
    t2[0] = a[1] * b2 - a[2] * b1;
    t2[1] = a[2] * b0 - a[0] * b2;
    t2[2] = a[0] * b1 - a[1] * b0;
    const double m2 = sqrt((t2[0] * t2[0] + t2[1] * t2[1]) + t2[2] * t2[2]);
    t2[0] = t2[0] / m2;
    t2[1] = t2[1] / m2;
    t2[2] = t2[2] / m2;
}

__global__ void __launch_bounds__(CG_THREADS) cg_reduce_kernel(const CB_GRID_CONSTANT CgArgs a) {
    CB_DYN_SMEM(smem_raw);
    const long long rep = blockIdx.y;
    const long long i0 = (long long)blockIdx.x * a.T;         // first interval of this block
    const long long i1 = min(a.M, i0 + (long long)a.T);       // one past the last
    const long long b0 = i0 * a.k, b1 = min(a.N, i1 * a.k);   // bead range
    const int k = (int)a.k, nb = (int)a.nb;
    const int s3 = cg_row_stride(k, 3), sn = cg_row_stride(k, nb > 0 ? nb : 1);
    // shared memory: r rows | t3 rows | states rows | marks rows, all staged before ONE barrier so that
    // every load of the block is in flight together
    double *sh_r = (double *)smem_raw, *sh_t = sh_r + (size_t)a.T * s3;
    long long *sh_s = (long long *)(sh_t + (size_t)a.T * s3), *sh_m = sh_s + (size_t)a.T * sn;
    cg_stage(sh_r, a.r + rep * a.N * 3, b0, b1, 3, k, s3);
    cg_stage(sh_t, a.t3 + rep * a.N * 3, b0, b1, 3, k, s3);
    if (a.states) cg_stage(sh_s, a.states + rep * a.N * nb, b0, b1, nb, k, sn);
    if (a.mods) cg_stage(sh_m, a.mods + rep * a.N * nb, b0, b1, nb, k, sn);
    __syncthreads();
    const long long me = i0 + threadIdx.x;
    if (me >= i1) return;
    const int cnt = (int)(min(a.N, (me + 1) * a.k) - me * a.k);
    double m[3], t2[3];
    // positions: interval mean, pulled inwards by cg_factor^(1/3) (rediscretize.py:452)
    cg_mean3(sh_r + (size_t)threadIdx.x * s3, cnt, m);
    double *o = a.r_cg + (rep * a.M + me) * 3;
    o[0] = m[0] / a.r_div;
    o[1] = m[1] / a.r_div;
    o[2] = m[2] / a.r_div;
    // orientations
    cg_mean3(sh_t + (size_t)threadIdx.x * s3, cnt, m);
    cg_orient(m, t2);
    double *o3 = a.t3_cg + (rep * a.M + me) * 3, *o2 = a.t2_cg + (rep * a.M + me) * 3;
    for (int c = 0; c < 3; c++) {
        o3[c] = m[c];
        o2[c] = t2[c];
    }
    // binding states and marks: most frequent value, smallest on ties (argmax(bincount))
    for (int which = 0; which < 2; which++) {
        const long long *row = (which == 0 ? sh_s : sh_m) + (size_t)threadIdx.x * sn;
        long long *dst = which == 0 ? a.states_cg : a.mods_cg;
        if (!(which == 0 ? a.states : a.mods)) continue;
        for (int c = 0; c < nb; c++) {
            // occurrence counts of the values 0..15 packed as 16 bytes (two 64-bit words): one shift
            // and add per bead; intervals longer than 255 beads are counted in chunks
            unsigned long long lo = 0, hi = 0;
            int count[CG_MAX_VALUE];
            bool bad = false, wide = cnt > 255;
            if (wide) {
#pragma unroll
                for (int v = 0; v < CG_MAX_VALUE; v++) count[v] = 0;
            }
            int vmax = 0;
            for (int j0 = 0; j0 < cnt; j0 += 255) {
                const int j1 = min(cnt, j0 + 255);
                for (int j = j0; j < j1; j++) {
                    const long long v = row[j * nb + c];
                    if (v < 0 || v >= CG_MAX_VALUE) {
                        bad = true;
                        continue;
                    }
                    vmax = max(vmax, (int)v);
                    const unsigned long long one = 1ull << (8 * ((int)v & 7));
                    lo += v < 8 ? one : 0ull;
                    hi += v < 8 ? 0ull : one;
                }
                if (wide) {
#pragma unroll
                    for (int w = 0; w < CG_MAX_VALUE; w++)
                        count[w] += (int)(((w < 8 ? lo : hi) >> (8 * (w & 7))) & 0xffull);
                    lo = hi = 0;
                }
            }
            int best = 0, best_count = -1;
            for (int w = 0; w <= vmax; w++) { // argmax(bincount): the smallest value among the most frequent
                const int cw = wide ? count[w] : (int)(((w < 8 ? lo : hi) >> (8 * (w & 7))) & 0xffull);
                if (cw > best_count) {
                    best = w;
                    best_count = cw;
                }
            }
            if (bad) *a.err = 1;
            dst[(rep * a.M + me) * nb + c] = cnt > 1 ? (long long)best : row[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// refinement
// ------------------------------------------------------------------------------------------------
#define RF_WALK_CHUNK 32

struct RefineLayout {
    long long M, Nref;
    long long seg, half1, half2, left; // get_refined_intervals rediscretize.py:565-583
    long long nseg;                    // M (+1 when left > 0)
    long long points, draws;
};

static inline __host__ __device__ int refine_layout(long long M, long long Nref, RefineLayout *L) {
    if (M < 2 || Nref < 1) return -1;
    L->M = M;
    L->Nref = Nref;
    L->seg = Nref / (M - 1);
    L->half1 = L->seg / 2;
    L->half2 = L->seg - L->half1;
    L->left = Nref % (M - 1);
    if (L->half2 - 1 < 1) return -2; // brownian_bridge(0, ...) divides by zero in the reference
    L->nseg = M + (L->left > 0 ? 1 : 0);
    L->points = L->half1 + (M - 2) * L->seg + (L->half2 - 1) + (L->left > 0 ? L->left + 1 : 0);
    L->draws = L->half1 + (M - 2) * (L->seg - 1) + (L->half2 - 2) + L->left;
    return 0;
}

// segment s of a path: number of steps, first output row, first Gaussian triple
__device__ __forceinline__ void refine_segment(const RefineLayout &L, long long s, long long *n, long long *row,
                                               long long *draw) {
    if (s == 0) {
        *n = L.half1;
        *row = 0;
        *draw = 0;
    } else if (s < L.M - 1) {
        *n = L.seg;
        *row = L.half1 + (s - 1) * L.seg;
        *draw = L.half1 + (s - 1) * (L.seg - 1);
    } else if (s == L.M - 1) {
        *n = L.half2 - 1;
        *row = L.half1 + (L.M - 2) * L.seg;
        *draw = L.half1 + (L.M - 2) * (L.seg - 1);
    } else {
        *n = L.left;
        *row = L.half1 + (L.M - 2) * L.seg + (L.half2 - 1);
        *draw = L.half1 + (L.M - 2) * (L.seg - 1) + (L.half2 - 2);
    }
}

struct RefineArgs {
    RefineLayout L;
    long long R;
    double spacing;   // bead_spacing / avg_step_target
    double out_scale; // r_refine *= scaling (rediscretize.py:1056); ignored when orient != 0
    int orient;       // get_refined_orientations: rows normalised, t2 completed
    const double *cg; // [R][M][3]
    const double *xi; // [R][draws][3] standard normal deviates in the reference's draw order, or NULL
    unsigned long long seed;
    double *out, *out_t2; // [R][points][3]
    int wpb;              // warps per block
    int per_warp;         // doubles of shared memory per warp
    int G;                // lanes per inner bridge
    long long inner_warps; // warps that carry the inner bridges of one replica
};

// production deviates: triple q of replica rep = Box-Muller of Philox4x32-10(counter (q_lo, q_hi, j, rep), key seed)
__device__ __forceinline__ void philox_normal3(unsigned long long seed, uint32_t rep, unsigned long long q,
                                               double z[3]) {
    uint32_t o[4], p[4];
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 0u, rep, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5EEDu, o);
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 1u, rep, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5EEDu, p);
    const double u1 = ((double)o[0] + 1.0) * (1.0 / 4294967296.0), u2 = (double)o[1] * (1.0 / 4294967296.0);
    const double u3 = ((double)o[2] + 1.0) * (1.0 / 4294967296.0), u4 = (double)p[0] * (1.0 / 4294967296.0);
    double s, c;
    const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
    sincospi(2.0 * u2, &s, &c);
    z[0] = r1 * c;
    z[1] = r1 * s;
    sincospi(2.0 * u4, &s, &c);
    z[2] = r2 * c;
}

__device__ __forceinline__ void refine_fetch_xi(const RefineArgs &a, long long rep, long long q, double z[3]) {
    if (a.xi) {
        const double *p = a.xi + (rep * a.L.draws + q) * 3;
        z[0] = p[0];
        z[1] = p[1];
        z[2] = p[2];
    } else {
        philox_normal3(a.seed, (uint32_t)rep, (unsigned long long)q, z);
    }
}

// write one finished row (position mode: scaled; orientation mode: normalised + t2)
__device__ __forceinline__ void refine_store(const RefineArgs &a, long long rep, long long row, double x, double y,
                                             double z) {
    double *o = a.out + (rep * a.L.points + row) * 3;
    if (a.orient) {
        double v[3] = {x, y, z}, t2[3];
        cg_orient(v, t2);
        double *o2 = a.out_t2 + (rep * a.L.points + row) * 3;
        for (int c = 0; c < 3; c++) {
            o[c] = v[c];
            o2[c] = t2[c];
        }
    } else {
        o[0] = x * a.out_scale;
        o[1] = y * a.out_scale;
        o[2] = z * a.out_scale;
    }
}

// sum over the G lanes of a group (G a power of two; every lane of the warp takes part)
__device__ __forceinline__ double group_sum(double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// np.linspace(p0, p1, n + 1)[j] (numpy/_core/function_base.py: step = delta / div; y = j * step + start,
// or (j / div) * delta + start when any component of step is zero; the last sample is p1 itself)
__device__ __forceinline__ double linspace_at(double p0, double p1, double delta, double step, bool any_zero, int j,
                                              int n) {
    if (j == n) return p1;
    return any_zero ? ((double)j / (double)n) * delta + p0 : (double)j * step + p0;
}

// one group of G lanes (G = 4..32, a power of two; `lane` = lane within the group): a Brownian bridge of n
// steps from p0 to p1 (rows 0..n-1 are emitted; row n = p1 belongs to the next segment),
// rediscretize.py:634-685.  All groups of a warp run this with the same n, so the warp-wide barriers and
// shuffles are reached uniformly; a group without a bridge (`active` false) only keeps step.
__device__ void refine_bridge(const RefineArgs &a, long long rep, long long n_, long long row0, long long draw0,
                              const double *p0, const double *p1, double *sh, int G, int lane, bool active) {
    const int n = (int)n_;
    if (n == 1) { // trivial case: [p0, p1][:1]
        if (lane == 0 && active) refine_store(a, rep, row0, p0[0], p0[1], p0[2]);
        return;
    }
    if (!active) { // same barriers and shuffles as below, no memory traffic
        __syncwarp();
        __syncwarp();
        __syncwarp();
        (void)group_sum(0.0, G);
        return;
    }
    double *B = sh;                  // [(n + 1)][3]
    double *coef = sh + 3 * (n + 1); // [n]
    const double dt = 1.0 / (double)n, dt_sqrt = sqrt(dt);
    // increments and decay factors, all lanes
    for (int j = lane; j < n - 1; j += G) {
        double z[3];
        refine_fetch_xi(a, rep, draw0 + j, z);
        const double t = (double)j * dt;
        coef[j] = 1.0 - dt / (1.0 - t);
        B[3 * (j + 1)] = z[0] * dt_sqrt;
        B[3 * (j + 1) + 1] = z[1] * dt_sqrt;
        B[3 * (j + 1) + 2] = z[2] * dt_sqrt;
    }
    __syncwarp();
    // B[j+1] = B[j] * coef[j] + xi[j]: serial in j, one lane per coordinate
    if (lane < 3) {
        double b = 0.0;
        B[lane] = 0.0;
        for (int j = 0; j < n - 1; j++) {
            b = b * coef[j] + B[3 * (j + 1) + lane];
            B[3 * (j + 1) + lane] = b;
        }
        B[3 * n + lane] = 0.0;
    }
    __syncwarp();
    double d[3], step[3];
    for (int c = 0; c < 3; c++) d[c] = p1[c] - p0[c];
    const double dpl = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    const double direct_step = dpl / (double)n;
    bool any_zero = false;
    for (int c = 0; c < 3; c++) {
        step[c] = d[c] / (double)n;
        any_zero = any_zero || step[c] == 0.0;
    }
    // B *= direct_path_length; path length of direct + B
    for (int i = lane; i < 3 * (n + 1); i += G) B[i] = B[i] * dpl;
    __syncwarp();
    double len = 0.0;
    for (int j = lane; j < n; j += G) {
        double q[3];
        for (int c = 0; c < 3; c++) {
            const double x0 = linspace_at(p0[c], p1[c], d[c], step[c], any_zero, j, n) + B[3 * j + c];
            const double x1 = linspace_at(p0[c], p1[c], d[c], step[c], any_zero, j + 1, n) + B[3 * (j + 1) + c];
            q[c] = x1 - x0;
        }
        len += sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    }
    len = group_sum(len, G);
    double avg = len / (double)n;
    if (a.spacing < direct_step) avg = direct_step; // the reference prints a notice and adjusts the spacing
    const double actual_to_direct = avg / direct_step, target_to_direct = a.spacing / direct_step;
    const double den = actual_to_direct / target_to_direct;
    for (int j = lane; j < n; j += G) {
        double x[3];
        for (int c = 0; c < 3; c++)
            x[c] = linspace_at(p0[c], p1[c], d[c], step[c], any_zero, j, n) + B[3 * j + c] / den;
        refine_store(a, rep, row0 + j, x[0], x[1], x[2]);
    }
}

// one warp: a free end, n unit-direction steps of length `spacing` away from `start`
// (gaussian_walk poly_paths.py:288-295).  flip: rows n, n-1, ..., 1 (segment 0); else rows 0..n.
__device__ void refine_walk(const RefineArgs &a, long long rep, long long n, long long row0, long long draw0,
                            const double *start, bool flip, double *sh) {
    const int lane = threadIdx.x & 31;
    double carry = 0.0; // running cumsum of this lane's coordinate (lanes 0..2)
    if (!flip && lane == 0) refine_store(a, rep, row0, 0.0 + start[0], 0.0 + start[1], 0.0 + start[2]);
    for (long long base = 0; base < n; base += RF_WALK_CHUNK) {
        const int m = (int)min((long long)RF_WALK_CHUNK, n - base);
        for (int j = lane; j < m; j += 32) {
            double z[3];
            refine_fetch_xi(a, rep, draw0 + base + j, z);
            const double mag = sqrt((z[0] * z[0] + z[1] * z[1]) + z[2] * z[2]);
            for (int c = 0; c < 3; c++) sh[3 * j + c] = (z[c] / mag) * a.spacing;
        }
        __syncwarp();
        if (lane < 3) {
            for (int j = 0; j < m; j++) {
                carry = (base + j == 0) ? sh[3 * j + lane] : carry + sh[3 * j + lane];
                sh[3 * j + lane] = carry;
            }
        }
        __syncwarp();
        for (int j = lane; j < m; j += 32) {
            const long long step = base + j; // point index step + 1 of the walk
            const long long row = flip ? row0 + (n - 1 - step) : row0 + 1 + step;
            refine_store(a, rep, row, sh[3 * j] + start[0], sh[3 * j + 1] + start[1], sh[3 * j + 2] + start[2]);
        }
        __syncwarp();
    }
}

// lanes per inner bridge: the smallest power of two >= its number of steps, within [4, 32]
static inline __host__ __device__ int refine_group(long long seg) {
    int g = 4;
    while (g < 32 && g < seg) g <<= 1;
    return g;
}

__global__ void refine_path_kernel(const CB_GRID_CONSTANT RefineArgs a) {
    CB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sh = (double *)smem_raw + (size_t)warp * a.per_warp;
    const long long rep = blockIdx.y;
    const long long wi = (long long)blockIdx.x * a.wpb + warp; // warp index within the replica
    const double *cg = a.cg + rep * a.L.M * 3;
    long long n, row, draw;
    if (wi < a.inner_warps) {
        // inner bridges 1 .. M-2 (all of L.seg steps): 32 / G of them per warp
        const int G = a.G, gpw = 32 / G, grp = lane / G;
        const long long s = 1 + wi * gpw + grp;
        const bool active = s < a.L.M - 1;
        refine_segment(a.L, active ? s : 1, &n, &row, &draw);
        refine_bridge(a, rep, n, row, draw, cg + (s - 1) * 3, cg + s * 3, sh + (size_t)grp * 4 * (a.L.seg + 1), G,
                      lane & (G - 1), active);
        return;
    }
    // the two free ends and the last (shorter) bridge: one warp each
    const long long k = wi - a.inner_warps;
    const long long s = k == 0 ? 0 : k == 1 ? a.L.M - 1 : a.L.M;
    if (k > 2 || s >= a.L.nseg) return;
    refine_segment(a.L, s, &n, &row, &draw);
    if (s == 0) {
        refine_walk(a, rep, n, row, draw, cg, true, sh);
    } else if (s == a.L.M) {
        refine_walk(a, rep, n, row, draw, cg + (a.L.M - 1) * 3, false, sh);
    } else {
        refine_bridge(a, rep, n, row, draw, cg + (s - 1) * 3, cg + s * 3, sh, 32, lane, true);
    }
}

// ------------------------------------------------------------------------------------------------
// spherical confinement of a refined path
// ------------------------------------------------------------------------------------------------
struct ConfineArgs {
    long long R, N;
    double rad;
    const double *in;
    double *out;
};

#define CF_TILE 1024
__global__ void __launch_bounds__(256) confine_kernel(const CB_GRID_CONSTANT ConfineArgs a) {
    // each bead is rescaled once per violator within three beads of it, in ascending violator order,
    // with the violation measured on the ORIGINAL path (rediscretize.py:728-752)
    __shared__ double tile[(CF_TILE + 6) * 3];
    __shared__ double inv_f[CF_TILE + 6]; // 1 / (dist / rad) of a violator, 0 otherwise
    const long long rep = blockIdx.y;
    const long long j0 = (long long)blockIdx.x * CF_TILE;
    const double *src = a.in + rep * a.N * 3;
    const long long lo = max(0LL, j0 - 3), hi = min(a.N, j0 + CF_TILE + 3);
    {
        // all of a thread's loads are issued before the first store to shared memory
        constexpr int PER = ((CF_TILE + 6) * 3 + 255) / 256;
        double v[PER];
        const long long e0 = lo * 3 + threadIdx.x, e1 = hi * 3;
#pragma unroll
        for (int u = 0; u < PER; u++) v[u] = e0 + u * 256 < e1 ? src[e0 + u * 256] : 0.0;
#pragma unroll
        for (int u = 0; u < PER; u++)
            if (e0 + u * 256 < e1) tile[e0 + u * 256 - (j0 - 3) * 3] = v[u];
    }
    for (int t = threadIdx.x; t < CF_TILE + 6; t += 256) inv_f[t] = 0.0; // slots outside [0, N): no violator
    __syncthreads();
    for (int t = (int)(lo - (j0 - 3)) + threadIdx.x; t < (int)(hi - (j0 - 3)); t += 256) {
        const double *p = tile + t * 3;
        const double dist = sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]);
        if (dist > a.rad) inv_f[t] = 1.0 / (dist / a.rad);
    }
    __syncthreads();
    const double kern[7] = {0.98, 0.97, 0.96, 0.95, 0.96, 0.97, 0.98};
    // (beads outside [0, N) have inv_f = 0: the halo slots were zeroed first)
    const int nbeads = (int)(min(a.N, j0 + CF_TILE) - j0);
    // one thread per bead: its window factors once, applied to the three coordinates in place
    for (int b = threadIdx.x; b < nbeads; b += 256) {
        double *p = tile + (b + 3) * 3;
        double x = p[0], y = p[1], z = p[2];
        bool any = false;
#pragma unroll
        for (int o = -3; o <= 3; o++) {
            const double f = inv_f[b + 3 + o]; // violator b + o
            if (f != 0.0) {
                const double w = kern[3 - o] * f; // bead sits at offset 3 - o of that violator's window
                x = x * w;
                y = y * w;
                z = z * w;
                any = true;
            }
        }
        if (any) {
            p[0] = x;
            p[1] = y;
            p[2] = z;
        }
    }
    __syncthreads();
    double *dst = a.out + (rep * a.N + j0) * 3;
    for (int e = threadIdx.x; e < nbeads * 3; e += 256) dst[e] = tile[e + 9]; // coalesced
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool on = false;
};

#ifdef CHROMO_HOST_EMU
static inline void timer_start(Timer &, double *, cudaStream_t) {}
static inline void timer_stop(Timer &, double *ms, cudaStream_t) {
    if (ms) *ms = 0.0;
}
#else
static inline void timer_start(Timer &t, double *ms, cudaStream_t s) {
    if (!ms) return;
    t.on = cudaEventCreate(&t.e0) == cudaSuccess && cudaEventCreate(&t.e1) == cudaSuccess;
    if (t.on) cudaEventRecord(t.e0, s);
}
static inline void timer_stop(Timer &t, double *ms, cudaStream_t s) {
    if (!ms) return;
    *ms = 0.0;
    if (t.on) {
        float f = 0.f;
        cudaEventRecord(t.e1, s);
        cudaEventSynchronize(t.e1);
        cudaEventElapsedTime(&f, t.e0, t.e1);
        *ms = f;
    }
    if (t.e0) cudaEventDestroy(t.e0);
    if (t.e1) cudaEventDestroy(t.e1);
}
#endif

static int select_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        return cb_set_error(CHROMO_ERR_CUDA, "no CUDA device (chromo_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return cb_set_error(CHROMO_ERR_ARG, "device %d out of range (%d present)", device, n);
    if (cudaSetDevice(device) != cudaSuccess) return cb_set_error(CHROMO_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    return 0;
}

template <class T>
static cudaError_t dev_alloc(T **p, size_t n) {
    return cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
}

extern "C" int64_t chromo_cg_num_beads(int64_t num_beads, int64_t cg_factor) {
    if (num_beads < 1 || cg_factor < 1 || num_beads < cg_factor) return -1;
    return num_beads / cg_factor + (num_beads % cg_factor ? 1 : 0);
}

extern "C" int chromo_cg_chromatin(int device, int64_t R, int64_t N, int64_t nb, int64_t cg_factor, double r_divisor,
                                   const double *r, const double *t3, const int64_t *states, const int64_t *mods,
                                   double *r_cg, double *t3_cg, double *t2_cg, int64_t *states_cg, int64_t *mods_cg,
                                   double *kernel_ms) {
    const int64_t M = chromo_cg_num_beads(N, cg_factor);
    if (R < 1 || M < 1 || nb < 0 || nb > 8)
        return cb_set_error(CHROMO_ERR_ARG, "chromo_cg_chromatin: need R >= 1, 1 <= cg_factor <= num_beads, nb <= 8");
    if (!r || !t3 || !r_cg || !t3_cg || !t2_cg || (nb && states && !states_cg) || (nb && mods && !mods_cg))
        return cb_set_error(CHROMO_ERR_ARG, "chromo_cg_chromatin: null array");
    if (!(r_divisor > 0.0)) return cb_set_error(CHROMO_ERR_ARG, "chromo_cg_chromatin: r_divisor must be positive");
    int rc = select_device(device);
    if (rc) return rc;
    const size_t budget = 56 * 1024 / 8; // 64-bit words of staging per block: four blocks per SM
    const int s3 = cg_row_stride((int)cg_factor, 3), sn = cg_row_stride((int)cg_factor, (int)std::max<int64_t>(nb, 1));
    const size_t per_interval = 2 * (size_t)s3 + 2 * (size_t)sn;
    if (per_interval > budget) return cb_set_error(CHROMO_ERR_ARG, "chromo_cg_chromatin: cg_factor too large");
    const int T = (int)std::max<size_t>(1, std::min<size_t>(CG_THREADS, budget / per_interval));
    const size_t smem = (size_t)T * per_interval * 8;
    const size_t nr = (size_t)R * N * 3, ns = (size_t)R * N * nb, ncg = (size_t)R * M * 3, nscg = (size_t)R * M * nb;
    double *d_r = nullptr, *d_t3 = nullptr, *d_rcg = nullptr, *d_t3cg = nullptr, *d_t2cg = nullptr;
    long long *d_st = nullptr, *d_md = nullptr, *d_stcg = nullptr, *d_mdcg = nullptr;
    int *d_err = nullptr, h_err = 0;
    cudaStream_t s = nullptr;
    Timer tm;
    CgArgs a{};
    RCK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    RCK(dev_alloc(&d_r, nr));
    RCK(dev_alloc(&d_t3, nr));
    RCK(dev_alloc(&d_rcg, ncg));
    RCK(dev_alloc(&d_t3cg, ncg));
    RCK(dev_alloc(&d_t2cg, ncg));
    RCK(dev_alloc(&d_err, 1));
    RCK(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    RCK(cudaMemcpyAsync(d_r, r, nr * 8, cudaMemcpyHostToDevice, s));
    RCK(cudaMemcpyAsync(d_t3, t3, nr * 8, cudaMemcpyHostToDevice, s));
    if (nb && states) {
        RCK(dev_alloc(&d_st, ns));
        RCK(dev_alloc(&d_stcg, nscg));
        RCK(cudaMemcpyAsync(d_st, states, ns * 8, cudaMemcpyHostToDevice, s));
    }
    if (nb && mods) {
        RCK(dev_alloc(&d_md, ns));
        RCK(dev_alloc(&d_mdcg, nscg));
        RCK(cudaMemcpyAsync(d_md, mods, ns * 8, cudaMemcpyHostToDevice, s));
    }
    a.R = R; a.N = N; a.nb = nb; a.k = cg_factor; a.M = M; a.r_div = r_divisor;
    a.r = d_r; a.t3 = d_t3; a.states = d_st; a.mods = d_md;
    a.r_cg = d_rcg; a.t3_cg = d_t3cg; a.t2_cg = d_t2cg; a.states_cg = d_stcg; a.mods_cg = d_mdcg;
    a.err = d_err; a.T = T;
    RCK(cudaFuncSetAttribute(cg_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_start(tm, kernel_ms, s);
    CB_LAUNCH(cg_reduce_kernel, dim3((unsigned)((M + T - 1) / T), (unsigned)R), dim3(CG_THREADS), smem, s, a);
    RCK(cudaGetLastError());
    timer_stop(tm, kernel_ms, s);
    RCK(cudaMemcpyAsync(r_cg, d_rcg, ncg * 8, cudaMemcpyDeviceToHost, s));
    RCK(cudaMemcpyAsync(t3_cg, d_t3cg, ncg * 8, cudaMemcpyDeviceToHost, s));
    RCK(cudaMemcpyAsync(t2_cg, d_t2cg, ncg * 8, cudaMemcpyDeviceToHost, s));
    if (d_stcg) RCK(cudaMemcpyAsync(states_cg, d_stcg, nscg * 8, cudaMemcpyDeviceToHost, s));
    if (d_mdcg) RCK(cudaMemcpyAsync(mods_cg, d_mdcg, nscg * 8, cudaMemcpyDeviceToHost, s));
    RCK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    RCK(cudaStreamSynchronize(s));
    if (h_err)
        rc = cb_set_error(CHROMO_ERR_ARG, "chromo_cg_chromatin: states / marks must lie in [0, %d)", CG_MAX_VALUE);
done:
    for (void *p : {(void *)d_r, (void *)d_t3, (void *)d_rcg, (void *)d_t3cg, (void *)d_t2cg, (void *)d_st, (void *)d_md,
                    (void *)d_stcg, (void *)d_mdcg, (void *)d_err})
        if (p) cudaFree(p);
    if (s) cudaStreamDestroy(s);
    return rc;
}

extern "C" int64_t chromo_refined_num_points(int64_t num_beads_cg, int64_t num_beads_refined) {
    RefineLayout L;
    return refine_layout(num_beads_cg, num_beads_refined, &L) ? -1 : L.points;
}

extern "C" int64_t chromo_refined_num_draws(int64_t num_beads_cg, int64_t num_beads_refined) {
    RefineLayout L;
    return refine_layout(num_beads_cg, num_beads_refined, &L) ? -1 : L.draws;
}

extern "C" int chromo_refine_path(int device, int64_t R, int64_t num_beads_cg, int64_t num_beads_refined,
                                  double bead_spacing, const double *cg_r, const double *xi, uint64_t seed,
                                  double out_scale, int orientations, double *out, double *out_t2,
                                  double *kernel_ms) {
    RefineLayout L;
    const int lrc = refine_layout(num_beads_cg, num_beads_refined, &L);
    if (lrc == -1) return cb_set_error(CHROMO_ERR_ARG, "chromo_refine_path: need >= 2 coarse beads and >= 1 refined bead");
    if (lrc == -2)
        return cb_set_error(CHROMO_ERR_ARG, "chromo_refine_path: fewer than 3 refined beads per coarse bond "
                                            "(the reference's brownian_bridge divides by zero)");
    if (R < 1 || !cg_r || !out || (orientations && !out_t2)) return cb_set_error(CHROMO_ERR_ARG, "chromo_refine_path: null array");
    if (L.seg > 4096) return cb_set_error(CHROMO_ERR_ARG, "chromo_refine_path: more than 4096 refined beads per coarse bond");
    int rc = select_device(device);
    if (rc) return rc;
    const int G = refine_group(L.seg), gpw = 32 / G;
    const long long inner_warps = (L.M - 2 + gpw - 1) / gpw, warps = inner_warps + 3;
    const int per_warp = (int)std::max<long long>((long long)gpw * 4 * (L.seg + 1), 3 * RF_WALK_CHUNK);
    const int wpb = (int)std::max<long long>(1, std::min<long long>(8, (160 * 1024 / 8) / per_warp));
    const size_t smem = (size_t)wpb * per_warp * 8;
    const size_t ncg = (size_t)R * L.M * 3, nxi = (size_t)R * L.draws * 3, nout = (size_t)R * L.points * 3;
    double *d_cg = nullptr, *d_xi = nullptr, *d_out = nullptr, *d_t2 = nullptr;
    cudaStream_t s = nullptr;
    Timer tm;
    RefineArgs a{};
    RCK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    RCK(dev_alloc(&d_cg, ncg));
    RCK(dev_alloc(&d_out, nout));
    RCK(cudaMemcpyAsync(d_cg, cg_r, ncg * 8, cudaMemcpyHostToDevice, s));
    if (xi) {
        RCK(dev_alloc(&d_xi, nxi));
        RCK(cudaMemcpyAsync(d_xi, xi, nxi * 8, cudaMemcpyHostToDevice, s));
    }
    if (orientations) RCK(dev_alloc(&d_t2, nout));
    a.L = L; a.R = R; a.spacing = bead_spacing; a.out_scale = out_scale; a.orient = orientations;
    a.cg = d_cg; a.xi = d_xi; a.seed = seed; a.out = d_out; a.out_t2 = d_t2; a.wpb = wpb; a.per_warp = per_warp;
    a.G = G; a.inner_warps = inner_warps;
    RCK(cudaFuncSetAttribute(refine_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    timer_start(tm, kernel_ms, s);
    CB_LAUNCH(refine_path_kernel, dim3((unsigned)((warps + wpb - 1) / wpb), (unsigned)R), dim3(32 * wpb), smem, s, a);
    RCK(cudaGetLastError());
    timer_stop(tm, kernel_ms, s);
    RCK(cudaMemcpyAsync(out, d_out, nout * 8, cudaMemcpyDeviceToHost, s));
    if (orientations) RCK(cudaMemcpyAsync(out_t2, d_t2, nout * 8, cudaMemcpyDeviceToHost, s));
    RCK(cudaStreamSynchronize(s));
done:
    for (void *p : {(void *)d_cg, (void *)d_xi, (void *)d_out, (void *)d_t2})
        if (p) cudaFree(p);
    if (s) cudaStreamDestroy(s);
    return rc;
}

extern "C" int chromo_enforce_spherical_confinement(int device, int64_t R, int64_t N, double *r, double rad,
                                                    double *kernel_ms) {
    if (R < 1 || N < 1 || !r) return cb_set_error(CHROMO_ERR_ARG, "chromo_enforce_spherical_confinement: bad arguments");
    if (N < 7) return cb_set_error(CHROMO_ERR_ARG, "chromo_enforce_spherical_confinement: needs at least 7 beads");
    int rc = select_device(device);
    if (rc) return rc;
    const size_t n = (size_t)R * N * 3;
    double *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s = nullptr;
    Timer tm;
    ConfineArgs a{};
    RCK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    RCK(dev_alloc(&d_in, n));
    RCK(dev_alloc(&d_out, n));
    RCK(cudaMemcpyAsync(d_in, r, n * 8, cudaMemcpyHostToDevice, s));
    a.R = R; a.N = N; a.rad = rad; a.in = d_in; a.out = d_out;
    timer_start(tm, kernel_ms, s);
    CB_LAUNCH(confine_kernel, dim3((unsigned)((N + CF_TILE - 1) / CF_TILE), (unsigned)R), dim3(256), 0, s, a);
    RCK(cudaGetLastError());
    timer_stop(tm, kernel_ms, s);
    RCK(cudaMemcpyAsync(r, d_out, n * 8, cudaMemcpyDeviceToHost, s));
    RCK(cudaStreamSynchronize(s));
done:
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (s) cudaStreamDestroy(s);
    return rc;
}
