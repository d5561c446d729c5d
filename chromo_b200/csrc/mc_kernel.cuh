// mc_kernel.cuh -- the fused Monte-Carlo kernel: proposal -> elastic dE ->
// field dE -> Metropolis -> commit, for one replica per warp.
//
// Why a warp per replica: moves inside a replica are strictly serial (each
// reads the density the previous one wrote), a typical move touches 3-30 beads
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
#pragma once
#include "geometry.cuh"
#include "launch.cuh"
#include "params.cuh"
#include "rng.cuh"

#define FULL_MASK 0xffffffffu
#define HASH_EMPTY (-1)

// per-warp shared state
struct WarpSh {
    chromo_move_state mv[CHROMO_NUM_MOVES];
    uint32_t grs[CB_GLIBC_WORDS]; // ReplayRng state
    uint32_t rng_save[CB_GLIBC_WORDS];
    double M[12];                 // affine map of the current move
    int ind0, indf, n, binder;
    int count;                    // occupied hash slots
    int overflow;
    int last_U;                   // touched voxels of the last field dE (all passes)
    unsigned long long algo_bytes; // SURVEY 8(d) algorithmic bytes, accumulated by lane 0
    uint32_t draws[64];           // per-bead axis draws of tangent rotation
    signed char newst[256];       // new binding states (small path)
};

struct HashTable {
    int *keys;    // [cap]
    int *list;    // [cap]   occupied slots, in claim order
    double *vals; // [cap][ncol]
    int cap, shift, limit;
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

// ---------------------------------------------------------------- hash table
__device__ __forceinline__ void table_reset_all(HashTable &H, WarpSh &S, int ncol, int lane) {
    for (int i = lane; i < H.cap; i += 32) H.keys[i] = HASH_EMPTY;
    for (int i = lane; i < H.cap * ncol; i += 32) H.vals[i] = 0.0;
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
        for (int c = 0; c < ncol; c++) H.vals[slot * ncol + c] = 0.0;
    }
    __syncwarp();
    if (lane == 0) {
        S.count = 0;
        S.overflow = 0;
    }
    __syncwarp();
}
// claim (or find) the slot of `bin`; -1 on overflow.  Claims stop once `limit`
// slots are taken (at most 32 more can be in flight), so the list never
// overruns and every claimed slot is listed (table_clear relies on that).
__device__ __forceinline__ int table_claim(HashTable &H, WarpSh &S, int bin) {
    uint32_t slot = ((uint32_t)bin * 2654435761u) >> H.shift;
    for (int probe = 0; probe < H.cap; probe++) {
        int cur = *(volatile int *)&H.keys[slot];
        if (cur == bin) return (int)slot;
        if (cur == HASH_EMPTY) {
            if (*(volatile int *)&S.count >= H.limit) {
                S.overflow = 1;
                return -1;
            }
            int prev = atomicCAS(&H.keys[slot], HASH_EMPTY, bin);
            if (prev == bin) return (int)slot;
            if (prev == HASH_EMPTY) {
                int pos = atomicAdd(&S.count, 1);
                H.list[pos] = (int)slot;
                return (int)slot;
            }
        }
        slot = (slot + 1) & (uint32_t)(H.cap - 1);
    }
    S.overflow = 1;
    return -1;
}

// add one bead's 8-voxel stencil: sign * w[l] / V_access * {1, state_1..state_nb}
// (fields.pyx:1476-1520; contributions with |x| <= 1e-18 are dropped, quirk 3)
template <int NB>
__device__ __forceinline__ void table_scatter(const DevCtx &C, HashTable &H, WarpSh &S,
                                              const int idx[8], const double w[8], double sign,
                                              const signed char st[NB], int P, int p) {
#pragma unroll
    for (int l = 0; l < 8; l++) {
        int bin = idx[l];
        if (P > 1 && (bin & (P - 1)) != p) continue;
        int slot = table_claim(H, S, bin);
        if (slot < 0) continue;
        double V = C.access_vol ? C.access_vol[bin] : C.vol_bin;
        double base = w[l] / V;
        double t0 = sign * base;
        if (fabs(t0) > 1E-18) atomicAdd(&H.vals[slot * (NB + 1)], t0);
#pragma unroll
        for (int m = 0; m < NB; m++) {
            double t = sign * (base * (double)st[m]);
            if (fabs(t) > 1E-18) atomicAdd(&H.vals[slot * (NB + 1) + 1 + m], t);
        }
    }
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
                                             bool want_cross) {
    constexpr int NCOL = NB + 1;
    int cnt = S.count;
    const double *dens = C.density + (long long)rep * C.n_bins * NCOL;
    for (int j = lane; j < cnt; j += 32) {
        int slot = H.list[j];
        int bin = H.keys[slot];
        const double *row = dens + (long long)bin * NCOL;
        double rho[NCOL], rn[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; c++) {
            rho[c] = row[c];
            rn[c] = rho[c] + H.vals[slot * NCOL + c];
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
        double V = C.access_vol ? C.access_vol[bin] : C.vol_bin;
        double vf0 = rho[0] * C.bead_vol;
        double vf1 = vf0 + (H.vals[slot * NCOL] * C.bead_vol);
        double e = 0.0;
        if (vf1 > C.vf_limit) e += CB_E_HUGE_FIELD * vf1;
        else e += chi * (V / C.bead_vol) * (vf1 * vf1);
        if (vf0 > C.vf_limit) e -= CB_E_HUGE_FIELD * vf0;
        else e -= chi * (V / C.bead_vol) * (vf0 * vf0);
        F.chi += e;
    }
}
// update_affected_densities fields.pyx:1968-1975 for the voxels in the table
__device__ __forceinline__ void table_commit(const DevCtx &C, const HashTable &H, const WarpSh &S,
                                             int rep, int lane) {
    int cnt = S.count;
    double *dens = C.density + (long long)rep * C.n_bins * C.ncol;
    for (int j = lane; j < cnt; j += 32) {
        int slot = H.list[j];
        double *row = dens + (long long)H.keys[slot] * C.ncol;
        for (int c = 0; c < C.ncol; c++) row[c] += H.vals[slot * C.ncol + c];
    }
}
__device__ __forceinline__ void table_debug_dump(const DevCtx &C, const HashTable &H,
                                                 const WarpSh &S, int lane, DebugOut *dbg) {
    int cnt = S.count;
    long long base = dbg->n_touched;
    for (int j = lane; j < cnt; j += 32) {
        long long o = base + j;
        if (o < dbg->touched_cap) {
            int slot = H.list[j];
            dbg->touched[o] = H.keys[slot];
            for (int c = 0; c < C.ncol; c++) dbg->dtrial[o * C.ncol + c] = H.vals[slot * C.ncol + c];
        }
    }
    __syncwarp();
    if (lane == 0) dbg->n_touched = base + cnt;
    __syncwarp();
}

// ------------------------------------------------------- bead selection (lane 0)
// capped_exponential bead_selection.pyx:19-67
template <class Rng>
__device__ int capped_exponential(Rng &g, int window, int cap) {
    long long r;
    do {
        r = (long long)(-log10(g.uniform() + 0.00001) * (double)window * 0.45 + 1.0001);
    } while (r > cap);
    return (int)r;
}
// from_point bead_selection.pyx:115-154 (from_left 69-90, from_right 93-112)
template <class Rng>
__device__ int from_point(Rng &g, int window, int N, int ind0) {
    if (window < 1) return ind0;
    int side = (int)(g.next31() % 2u);
    if (side == 0) {
        int ws = max(min(window, ind0), 1);
        int ub = max(ind0, 1);
        return ub - capped_exponential(g, ws, ws);
    }
    int ws = max(min(window, N - ind0), 1);
    return capped_exponential(g, ws, ws) + ind0;
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
template <class Rng, bool DEBUG, int NB>
struct McWarp {
    static constexpr int NCOL = NB + 1;
    const DevCtx &C;
    WarpSh &S;
    HashTable &H;
    Rng &rng;
    int rep, lane;
    double mu_adjust;
    int force_accept; // DEBUG only: -1 Metropolis, 0/1 forced
    DebugOut *dbg;

    __device__ double *R_() const { return C.r + (long long)rep * C.N * 3; }
    __device__ double *T3_() const { return C.t3 + (long long)rep * C.N * 3; }
    __device__ double *T2_() const { return C.t2 + (long long)rep * C.N * 3; }
    __device__ signed char *ST_() const { return C.states + (long long)rep * C.N * NB; }
    __device__ const signed char *MOD_() const { return C.mods + (long long)rep * C.N * NB; }

    // Metropolis test mc_sim.pyx:163-171 (lane 0 draws; result broadcast)
    __device__ bool metropolis(double dE) {
        int acc = 0;
        double u = __longlong_as_double(0x7ff8000000000000LL);
        if (lane == 0) {
            if (DEBUG && force_accept >= 0) {
                acc = force_accept;
            } else {
                double e = exp(-dE);
                u = rng.uniform();
                acc = (u < e) ? 1 : 0;
            }
            if (DEBUG) {
                dbg->u = u;
                dbg->accepted = acc;
            }
        }
        return __shfl_sync(FULL_MASK, acc, 0) != 0;
    }
    // AcceptanceTracker.update_acceptance_rate mc_stat.py:190-207 + counters
    __device__ void track(int mtype, bool acc) {
        if (lane == 0) {
            chromo_move_state &mv = S.mv[mtype];
            if (acc) mv.num_success += 1;
            mv.acceptance_rate = (mv.alpha * (acc ? 1.0 : 0.0)) + (1.0 - mv.alpha) * mv.acceptance_rate;
        }
    }

    // ---- field dE of a continuous segment under the affine map S.M -------
    // kind: 0 = rotation (crank / pivot), 1 = translation (slide),
    //       2 = state change (binding; positions unchanged)
    // Returns dE_field (valid on all lanes).  Leaves the table holding the
    // delta-rho rows when passes == 1.
    __device__ double field_dE_segment(int kind, int ind0, int n, int binder, const signed char *newst,
                                       int &passes_out) {
        const double *Rr = R_();
        const signed char *ST = ST_();
        double chi = C.chi[rep];
        int P = 1;
        FieldSums<NB> F;
        int out_t = 0, out_c = 0, dbl_t[NB], dbl_c[NB];
        bool want_cross = false;
        for (int a = 0; a < NB * NB; a++) want_cross |= (C.xpref[a] != 0.0);
        while (true) {
            for (int a = 0; a < NB; a++) {
                F.sq[a] = 0.0;
                dbl_t[a] = dbl_c[a] = 0;
            }
            for (int a = 0; a < NB * NB; a++) F.cross[a] = 0.0;
            F.chi = 0.0;
            out_t = out_c = 0;
            if (DEBUG && lane == 0) dbg->n_touched = 0;
            bool failed = false;
            for (int p = 0; p < P && !failed; p++) {
                for (int base = 0; base < n; base += 32) {
                    int i = base + lane;
                    if (i < n) {
                        int bead = ind0 + i;
                        double x[3], y[3];
                        load3(Rr + 3 * (long long)bead, x);
                        signed char sc[NB], sn[NB];
                        for (int m = 0; m < NB; m++) sn[m] = sc[m] = ST[(long long)bead * NB + m];
                        int idx[8];
                        double w[8];
                        bin_point(C, x[0], x[1], x[2], idx, w);
                        table_scatter<NB>(C, H, S, idx, w, -1.0, sc, P, p);
                        if (kind == 2) {
                            sn[binder] = newst[i];
                            if (p == 0)
                                for (int m = 0; m < NB; m++) {
                                    dbl_c[m] += (sc[m] == 2);
                                    dbl_t[m] += (sn[m] == 2);
                                }
                        } else {
                            if (kind == 0) apply_affine(S.M, x, y);
                            else
                                for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                            if (p == 0 && C.confine_type == CHROMO_CONFINE_SPHERICAL) {
                                // get_confinement_dE fields.pyx:160-175
                                out_t += (sqrt(dot3(y, y)) > C.confine_length);
                                out_c += (sqrt(dot3(x, x)) > C.confine_length);
                            } else if (p == 0 && C.confine_type == CHROMO_CONFINE_CUBICAL) {
                                // fields.pyx:178-193: the current configuration is never counted
                                for (int j = 0; j < 3; j++) out_t += (fabs(y[j]) > C.confine_length / 2);
                            }
                            bin_point(C, y[0], y[1], y[2], idx, w);
                        }
                        table_scatter<NB>(C, H, S, idx, w, 1.0, sn, P, p);
                    }
                }
                __syncwarp();
                if (S.overflow) {
                    failed = true;
                    break;
                }
                table_energy<NB>(C, H, S, rep, chi, lane, F, want_cross);
                if (lane == 0) S.last_U = (p == 0 ? 0 : S.last_U) + S.count;
                if (DEBUG) table_debug_dump(C, H, S, lane, dbg);
                if (P > 1) table_clear(H, S, NCOL, lane);
            }
            if (!failed) break;
            table_clear(H, S, NCOL, lane);
            P *= 2;
        }
        passes_out = P;
        // ---- reduce and assemble in the reference's order ----
        double dE = 0.0;
        if (kind != 2) { // compute_dE fields.pyx:1209-1211
            int nt = warp_sum_int(out_t), nc = warp_sum_int(out_c);
            dE += (double)nt * CB_E_HUGE_FIELD;
            dE -= (double)nc * CB_E_HUGE_FIELD;
        }
        double bb = 0.0; // get_dE_binders_and_beads fields.pyx:1760-1790
        for (int a = 0; a < NB; a++) {
            double tot = warp_sum(F.sq[a]);
            bb += C.pref[a] * tot;
            int dd = (kind == 2) ? warp_sum_int(dbl_t[a] - dbl_c[a]) : 0;
            bb += C.e_intra[a] * (double)dd;
        }
        for (int a = 0; a < NB; a++)
            for (int b = 0; b < NB; b++) {
                double tot = want_cross ? warp_sum(F.cross[a * NB + b]) : 0.0;
                bb += C.xpref[a * NB + b] * tot;
            }
        bb += warp_sum(F.chi);
        dE += bb;
        return dE;
    }

    // NullField.compute_dE fields.pyx:300-318: only the confinement acts
    __device__ double confinement_dE_segment(int kind, int ind0, int n) {
        const double *Rr = R_();
        int out_t = 0, out_c = 0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                double x[3], y[3];
                load3(Rr + 3 * (long long)(ind0 + i), x);
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

    // apply the accepted move's density change (update_affected_densities)
    __device__ void field_commit_segment(int kind, int ind0, int n, int binder, const signed char *newst,
                                         int passes) {
        if (passes == 1) {
            table_commit(C, H, S, rep, lane);
            return;
        }
        const double *Rr = R_();
        const signed char *ST = ST_();
        for (int p = 0; p < passes; p++) {
            table_clear(H, S, NCOL, lane);
            for (int base = 0; base < n; base += 32) {
                int i = base + lane;
                if (i < n) {
                    int bead = ind0 + i;
                    double x[3], y[3];
                    load3(Rr + 3 * (long long)bead, x);
                    signed char sc[NB], sn[NB];
                    for (int m = 0; m < NB; m++) sn[m] = sc[m] = ST[(long long)bead * NB + m];
                    int idx[8];
                    double w[8];
                    bin_point(C, x[0], x[1], x[2], idx, w);
                    table_scatter<NB>(C, H, S, idx, w, -1.0, sc, passes, p);
                    if (kind == 2) sn[binder] = newst[i];
                    else {
                        if (kind == 0) apply_affine(S.M, x, y);
                        else
                            for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                        bin_point(C, y[0], y[1], y[2], idx, w);
                    }
                    table_scatter<NB>(C, H, S, idx, w, 1.0, sn, passes, p);
                }
            }
            __syncwarp();
            table_commit(C, H, S, rep, lane);
        }
    }

    // ---- crank-shaft / end-pivot / slide ---------------------------------
    __device__ void segment_move(int mtype) {
        chromo_move_state &mv = S.mv[mtype];
        const int N = C.N;
        double *Rr = R_(), *T3 = T3_(), *T2 = T2_();
        if (lane == 0) {
            mv.num_attempt += 1; // MCAdapter.propose moves.pyx:151
            int ind0 = 0, indf = 0;
            double ang = 0.0;
            if (mtype == CHROMO_CRANK_SHAFT) { // move_funcs.pyx:80-99
                ang = mv.amp_move * (rng.uniform() - 0.5);
                int b0 = (int)(rng.uniform() * (double)N);
                int b1 = max(from_point(rng, mv.amp_bead, N, b0), 1);
                check_bead_bounds(b0, b1, N, ind0, indf);
                if (indf > ind0) {
                    int a, b, ful; // get_crank_shaft_axis move_funcs.pyx:157-234
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
                    double ra[3], rb[3], pt[3], dir[3];
                    load3(Rr + 3 * (long long)a, ra);
                    load3(Rr + 3 * (long long)b, rb);
                    load3(Rr + 3 * (long long)ful, pt);
                    for (int j = 0; j < 3; j++) dir[j] = ra[j] - rb[j];
                    double mag = sqrt((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
                    if (mag < 1E-5) {
                        uint32_t d1 = rng.next31(), d2 = rng.next31();
                        sphere_from_draws(d1, d2, dir);
                    } else {
                        double sc = 1.0 / mag;
                        for (int j = 0; j < 3; j++) dir[j] = dir[j] * sc;
                    }
                    rotation_matrix(dir, pt, ang, S.M);
                }
            } else if (mtype == CHROMO_END_PIVOT) { // move_funcs.pyx:325-344
                ang = mv.amp_move * (rng.uniform() - 0.5);
                int lhs = (int)(rng.next31() % 2u);
                if (lhs == 1) {
                    ind0 = 0;
                    indf = capped_exponential(rng, mv.amp_bead, mv.amp_bead) + 1;
                } else {
                    ind0 = N - capped_exponential(rng, mv.amp_bead, mv.amp_bead);
                    indf = N;
                }
                uint32_t d1 = rng.next31(), d2 = rng.next31();
                double axis[3], pt[3];
                sphere_from_draws(d1, d2, axis);
                int ful; // get_end_pivot_fulcrum move_funcs.pyx:349-398
                if (ind0 == 0 && indf != N) ful = indf;
                else if (ind0 != 0 && indf == N) ful = ind0 - 1;
                else if (ind0 == 0 && indf == N && lhs == 1) ful = indf - 1;
                else ful = ind0;
                load3(Rr + 3 * (long long)ful, pt);
                rotation_matrix(axis, pt, ang, S.M);
            } else { // slide move_funcs.pyx:441-463
                double amp = mv.amp_move * rng.uniform();
                uint32_t d1 = rng.next31(), d2 = rng.next31();
                double dir[3];
                sphere_from_draws(d1, d2, dir);
                for (int j = 0; j < 3; j++) S.M[4 * j + 3] = dir[j] * amp;
                int b0 = (int)(rng.next31() % (uint32_t)N);
                int b1 = from_point(rng, mv.amp_bead, N, b0);
                check_bead_bounds(b0, b1, N, ind0, indf);
            }
            S.ind0 = ind0;
            S.indf = indf;
            S.n = indf - ind0;
        }
        __syncwarp();
        const int ind0 = S.ind0, indf = S.indf, n = S.n;
        if (n <= 0) return; // mc_sim.pyx:151-152
        const int kind = (mtype == CHROMO_SLIDE) ? 1 : 0;

        // ---- elastic dE: continuous_dE_poly polymers.pyx:1084-1146 -------
        // lane 0: bond left of ind0 ("forward"), lane 1: bond right of indf-1 ("reverse")
        double de = 0.0;
        if (lane == 0 && ind0 != 0) {
            double r0[3], r1[3], t0[3], t1[3], r1n[3], t1n[3];
            load3(Rr + 3 * (long long)(ind0 - 1), r0);
            load3(Rr + 3 * (long long)ind0, r1);
            load3(T3 + 3 * (long long)(ind0 - 1), t0);
            load3(T3 + 3 * (long long)ind0, t1);
            if (kind == 0) {
                apply_affine(S.M, r1, r1n);
                apply_rot(S.M, t1, t1n);
            } else {
                for (int j = 0; j < 3; j++) {
                    r1n[j] = r1[j] + S.M[4 * j + 3];
                    t1n[j] = t1[j];
                }
            }
            Bond B = load_bond(C, rep, ind0 - 1);
            de = bond_energy(B, r0, r1n, t0, t1n) - bond_energy(B, r0, r1, t0, t1);
        } else if (lane == 1 && indf != N) {
            double r0[3], r1[3], t0[3], t1[3], r0n[3], t0n[3];
            load3(Rr + 3 * (long long)(indf - 1), r0);
            load3(Rr + 3 * (long long)indf, r1);
            load3(T3 + 3 * (long long)(indf - 1), t0);
            load3(T3 + 3 * (long long)indf, t1);
            if (kind == 0) {
                apply_affine(S.M, r0, r0n);
                apply_rot(S.M, t0, t0n);
            } else {
                for (int j = 0; j < 3; j++) {
                    r0n[j] = r0[j] + S.M[4 * j + 3];
                    t0n[j] = t0[j];
                }
            }
            Bond B = load_bond(C, rep, indf - 1);
            de = bond_energy(B, r0n, r1, t0n, t1) - bond_energy(B, r0, r1, t0, t1);
        }
        double dE_poly = __shfl_sync(FULL_MASK, de, 0) + __shfl_sync(FULL_MASK, de, 1);

        // ---- field dE ----------------------------------------------------
        double dE_field = 0.0;
        int passes = 1;
        if (C.field_active) dE_field = field_dE_segment(kind, ind0, n, 0, nullptr, passes);
        else if (C.confine_type != CHROMO_CONFINE_NONE) dE_field = confinement_dE_segment(kind, ind0, n);
        if (DEBUG) debug_report_segment(kind, ind0, n, dE_poly, dE_field, passes);

        double dE = 0.0;
        dE += dE_poly;
        if (C.field_active || C.confine_type != CHROMO_CONFINE_NONE) dE += dE_field;
        bool acc = metropolis(dE);
        if (acc) { // MCAdapter.accept moves.pyx:190-226
            if (C.field_active) field_commit_segment(kind, ind0, n, 0, nullptr, passes);
            for (int base = 0; base < n; base += 32) {
                int i = base + lane;
                if (i < n) {
                    long long o = 3 * (long long)(ind0 + i);
                    double x[3], y[3];
                    load3(Rr + o, x);
                    if (kind == 0) {
                        apply_affine(S.M, x, y);
                        store3(Rr + o, y);
                        load3(T3 + o, x);
                        apply_rot(S.M, x, y);
                        store3(T3 + o, y);
                        load3(T2 + o, x);
                        apply_rot(S.M, x, y);
                        store3(T2 + o, y);
                    } else {
                        for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                        store3(Rr + o, y);
                    }
                }
            }
        }
        if (C.field_active) table_clear(H, S, NCOL, lane);
        if (lane == 0) { // B = 72(n+2) + 72 n a + nb n + 8(nb+1) U (1+2a) + 80   
            unsigned long long U = C.field_active ? (unsigned long long)S.last_U : 0ull;
            S.algo_bytes += 72ull * (n + 2) + (acc ? 72ull * n : 0ull) + (unsigned long long)NB * n +
                            8ull * NCOL * U * (acc ? 3ull : 1ull) + 80ull;
        }
        track(mtype, acc);
    }

    __device__ void debug_report_segment(int kind, int ind0, int n, double dE_poly, double dE_field,
                                         int passes) {
        const double *Rr = R_(), *T3 = T3_(), *T2 = T2_();
        const signed char *ST = ST_();
        int W = 9 + NB;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                if (i < dbg->inds_cap) dbg->inds[i] = ind0 + i;
                if (i < dbg->rows_cap) {
                    long long o = 3 * (long long)(ind0 + i);
                    double x[3], y[3];
                    double *row = dbg->rows + (long long)i * W;
                    load3(Rr + o, x);
                    if (kind == 0) apply_affine(S.M, x, y);
                    else
                        for (int j = 0; j < 3; j++) y[j] = x[j] + S.M[4 * j + 3];
                    store3(row, y);
                    load3(T3 + o, x);
                    if (kind == 0) apply_rot(S.M, x, y);
                    else
                        for (int j = 0; j < 3; j++) y[j] = x[j];
                    store3(row + 3, y);
                    load3(T2 + o, x);
                    if (kind == 0) apply_rot(S.M, x, y);
                    else
                        for (int j = 0; j < 3; j++) y[j] = x[j];
                    store3(row + 6, y);
                    for (int m = 0; m < NB; m++) row[9 + m] = (double)ST[(long long)(ind0 + i) * NB + m];
                }
            }
        }
        if (lane == 0) {
            dbg->n_inds = n;
            dbg->dE_poly = dE_poly;
            dbg->dE_field = dE_field;
            dbg->passes = passes;
        }
        __syncwarp();
    }

    // ---- change_binding_state move_funcs.pyx:717-820 ----------------------
    __device__ void binding_move() {
        chromo_move_state &mv = S.mv[CHROMO_CHANGE_BINDING_STATE];
        const int N = C.N;
        signed char *ST = ST_();
        const signed char *MOD = MOD_();
        if (lane == 0) {
            mv.num_attempt += 1;
            int binder = (int)(rng.next31() % (uint32_t)NB);
            int b0 = (int)(rng.next31() % (uint32_t)N);
            int b1 = from_point(rng, mv.amp_bead, N, b0);
            int ind0, indf;
            check_bead_bounds(b0, b1, N, ind0, indf);
            S.ind0 = ind0;
            S.indf = indf;
            S.n = indf - ind0;
            S.binder = binder;
            int n = indf - ind0;
            signed char *dst = (n <= 256) ? S.newst : (C.st_new + (long long)rep * N);
            for (int i = 0; i < n; i++) dst[i] = (signed char)rng.randint(C.sites[binder] + 1);
        }
        __syncwarp();
        const int ind0 = S.ind0, n = S.n, binder = S.binder;
        if (n <= 0) return;
        const signed char *newst = (n <= 256) ? S.newst : (C.st_new + (long long)rep * N);

        // ---- binding_dE / bead_binding_dE polymers.pyx:1383-1538 ----------
        double de = 0.0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            if (i < n) {
                int bead = ind0 + i;
                double d = 0.0;
                int sc[NB], sn[NB];
                for (int m = 0; m < NB; m++) sn[m] = sc[m] = ST[(long long)bead * NB + m];
                sn[binder] = newst[i];
                if (C.max_binders != -1) {
                    long long tot = 0;
                    for (int m = 0; m < NB; m++) tot += sn[m];
                    if (tot > C.max_binders) d += CB_E_HUGE_POLY * (double)(tot - C.max_binders);
                    tot = 0;
                    for (int m = 0; m < NB; m++) tot += sc[m];
                    if (tot > C.max_binders) d -= CB_E_HUGE_POLY * (double)(tot - C.max_binders);
                }
                for (int m = 0; m < NB; m++) {
                    int Nm = MOD[(long long)bead * NB + m];
                    const double *Ft = C.bindF + ((long long)m * C.S1 + Nm) * C.S1;
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

        double dE_field = 0.0;
        int passes = 1;
        if (C.field_active) dE_field = field_dE_segment(2, ind0, n, binder, newst, passes);
        if (DEBUG) {
            int W = 9 + NB;
            for (int base = 0; base < n; base += 32) {
                int i = base + lane;
                if (i < n) {
                    if (i < dbg->inds_cap) dbg->inds[i] = ind0 + i;
                    if (i < dbg->rows_cap) {
                        long long o = 3 * (long long)(ind0 + i);
                        double *row = dbg->rows + (long long)i * W;
                        for (int j = 0; j < 3; j++) {
                            row[j] = R_()[o + j];
                            row[3 + j] = T3_()[o + j];
                            row[6 + j] = T2_()[o + j];
                        }
                        for (int m = 0; m < NB; m++)
                            row[9 + m] = (m == binder) ? (double)newst[i]
                                                       : (double)ST[(long long)(ind0 + i) * NB + m];
                    }
                }
            }
            if (lane == 0) {
                dbg->n_inds = n;
                dbg->dE_poly = dE_poly;
                dbg->dE_field = dE_field;
                dbg->passes = passes;
            }
            __syncwarp();
        }
        double dE = 0.0;
        dE += dE_poly;
        if (C.field_active) dE += dE_field;
        bool acc = metropolis(dE);
        if (acc) {
            if (C.field_active) field_commit_segment(2, ind0, n, binder, newst, passes);
            __syncwarp();
            for (int base = 0; base < n; base += 32) {
                int i = base + lane;
                if (i < n) ST[(long long)(ind0 + i) * NB + binder] = newst[i];
            }
        }
        if (C.field_active) table_clear(H, S, NCOL, lane);
        if (lane == 0) { // binding: 24 n (positions) + nb n (1+a) + 8(nb+1) U (1+2a)
            unsigned long long U = C.field_active ? (unsigned long long)S.last_U : 0ull;
            S.algo_bytes += 24ull * n + (unsigned long long)NB * n * (acc ? 2ull : 1ull) +
                            8ull * NCOL * U * (acc ? 3ull : 1ull);
        }
        track(CHROMO_CHANGE_BINDING_STATE, acc);
    }

    // ---- tangent_rotation move_funcs.pyx:470-582 --------------------------
    // per selected bead: own random axis, rotate t3/t2, both adjacent bonds
    // against the CURRENT neighbours (polymers.pyx:1075-1080; quirk 8); never
    // touches the field (mc_sim.pyx:145).
    __device__ double tangent_bead(int bead, uint32_t d1, uint32_t d2, double ang, double t3n[3],
                                   double t2n[3]) {
        const double *Rr = R_(), *T3 = T3_(), *T2 = T2_();
        const int N = C.N;
        double axis[3], M[12], t3c[3], t2c[3], rc[3];
        const double origin[3] = {0.0, 0.0, 0.0};
        sphere_from_draws(d1, d2, axis);
        rotation_matrix(axis, origin, ang, M);
        long long o = 3 * (long long)bead;
        load3(T3 + o, t3c);
        load3(T2 + o, t2c);
        load3(Rr + o, rc);
        apply_rot(M, t3c, t3n);
        apply_rot(M, t2c, t2n);
        double d = 0.0;
        if (bead != 0) {
            double r0[3], t0[3];
            load3(Rr + o - 3, r0);
            load3(T3 + o - 3, t0);
            Bond B = load_bond(C, rep, bead - 1);
            d += bond_energy(B, r0, rc, t0, t3n) - bond_energy(B, r0, rc, t0, t3c);
        }
        if (bead + 1 != N) {
            double r1[3], t1[3];
            load3(Rr + o + 3, r1);
            load3(T3 + o + 3, t1);
            Bond B = load_bond(C, rep, bead);
            d += bond_energy(B, rc, r1, t3n, t1) - bond_energy(B, rc, r1, t3c, t1);
        }
        return d;
    }

    __device__ void tangent_move() {
        chromo_move_state &mv = S.mv[CHROMO_TANGENT_ROTATION];
        const int N = C.N;
        double *T3 = T3_(), *T2 = T2_();
        double ang = 0.0;
        int k = 0;
        if (lane == 0) {
            mv.num_attempt += 1;
            ang = mv.amp_move * (rng.uniform() - 0.5);
            k = (int)(rng.next31() % (uint32_t)mv.amp_bead) + 1;
        }
        ang = __shfl_sync(FULL_MASK, ang, 0);
        k = __shfl_sync(FULL_MASK, k, 0);
        double dE_poly = 0.0;
        bool acc;
        if (k <= 32) {
            // get_inds move_funcs.pyx:552-582: k distinct draws, redraw on duplicates
            int my = -1;
            for (int i = 0; i < k; i++) {
                int c = 0;
                bool dup;
                do {
                    if (lane == 0) c = (int)(rng.next31() % (uint32_t)N);
                    c = __shfl_sync(FULL_MASK, c, 0);
                    dup = __any_sync(FULL_MASK, lane < i && my == c);
                } while (dup);
                if (lane == i) my = c;
            }
            if (lane == 0)
                for (int i = 0; i < 2 * k; i++) S.draws[i] = rng.next31();
            __syncwarp();
            double t3n[3], t2n[3], d = 0.0;
            if (lane < k) d = tangent_bead(my, S.draws[2 * lane], S.draws[2 * lane + 1], ang, t3n, t2n);
            // ordered sum over beads (the reference accumulates bead by bead)
            for (int i = 0; i < k; i++) dE_poly += __shfl_sync(FULL_MASK, d, i);
            if (DEBUG) {
                int W = 9 + NB;
                if (lane < k) {
                    if (lane < dbg->inds_cap) dbg->inds[lane] = my;
                    if (lane < dbg->rows_cap) {
                        double *row = dbg->rows + (long long)lane * W;
                        for (int j = 0; j < 3; j++) {
                            row[j] = R_()[3 * (long long)my + j];
                            row[3 + j] = t3n[j];
                            row[6 + j] = t2n[j];
                        }
                        for (int m = 0; m < NB; m++) row[9 + m] = (double)ST_()[(long long)my * NB + m];
                    }
                }
                if (lane == 0) {
                    dbg->n_inds = k;
                    dbg->dE_poly = dE_poly;
                    dbg->dE_field = 0.0;
                    dbg->n_touched = 0;
                    dbg->passes = 0;
                }
                __syncwarp();
            }
            double dE = 0.0;
            dE += dE_poly;
            acc = metropolis(dE);
            if (acc && lane < k) {
                store3(T3 + 3 * (long long)my, t3n);
                store3(T2 + 3 * (long long)my, t2n);
            }
        } else {
            // large path: indices in HBM scratch, membership in a bitmap,
            // per-bead draws regenerated on commit from a saved RNG state
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
                rng.save(S.rng_save);
            }
            __syncwarp();
            double d = 0.0;
            for (int base = 0; base < k; base += 32) {
                int cnt = min(32, k - base);
                if (lane == 0)
                    for (int i = 0; i < 2 * cnt; i++) S.draws[i] = rng.next31();
                __syncwarp();
                if (lane < cnt) {
                    double t3n[3], t2n[3];
                    int bead = inds[base + lane];
                    d += tangent_bead(bead, S.draws[2 * lane], S.draws[2 * lane + 1], ang, t3n, t2n);
                    if (DEBUG && base + lane < dbg->rows_cap) {
                        double *row = dbg->rows + (long long)(base + lane) * (9 + NB);
                        for (int j = 0; j < 3; j++) {
                            row[j] = R_()[3 * (long long)bead + j];
                            row[3 + j] = t3n[j];
                            row[6 + j] = t2n[j];
                        }
                        for (int m = 0; m < NB; m++) row[9 + m] = (double)ST_()[(long long)bead * NB + m];
                    }
                }
                __syncwarp();
            }
            dE_poly = warp_sum(d);
            if (DEBUG) {
                if (lane == 0) {
                    dbg->n_inds = k;
                    dbg->dE_poly = dE_poly;
                    dbg->dE_field = 0.0;
                    dbg->n_touched = 0;
                    dbg->passes = 0;
                }
                for (int i = lane; i < k && i < dbg->inds_cap; i += 32) dbg->inds[i] = inds[i];
                __syncwarp();
            }
            double dE = 0.0;
            dE += dE_poly;
            acc = metropolis(dE);
            if (acc) {
                // every selected bead is distinct and its energy used only the
                // current neighbours, so all new tangents are computed from the
                // pre-move state before any is stored
                uint32_t after[CB_GLIBC_WORDS];
                if (lane == 0) {
                    rng.save(after);
                    rng.restore(S.rng_save);
                }
                // pass 1: recompute and park the new tangents in registers per chunk,
                // storing only after the whole chunk's loads are done; neighbours in
                // other chunks may already be rotated, but tangents of a bead depend
                // on its OWN t3/t2 only, so the stored values are unaffected.
                for (int base = 0; base < k; base += 32) {
                    int cnt = min(32, k - base);
                    if (lane == 0)
                        for (int i = 0; i < 2 * cnt; i++) S.draws[i] = rng.next31();
                    __syncwarp();
                    if (lane < cnt) {
                        int bead = inds[base + lane];
                        double axis[3], M[12], v[3], o3[3], o2[3];
                        const double origin[3] = {0.0, 0.0, 0.0};
                        sphere_from_draws(S.draws[2 * lane], S.draws[2 * lane + 1], axis);
                        rotation_matrix(axis, origin, ang, M);
                        load3(T3 + 3 * (long long)bead, v);
                        apply_rot(M, v, o3);
                        load3(T2 + 3 * (long long)bead, v);
                        apply_rot(M, v, o2);
                        store3(T3 + 3 * (long long)bead, o3);
                        store3(T2 + 3 * (long long)bead, o2);
                    }
                    __syncwarp();
                }
                if (lane == 0) rng.restore(after);
            }
        }
        if (lane == 0) // tangent rotation: B = 48 n + 72*2n + 48 n a   
            S.algo_bytes += 48ull * k + 144ull * k + (acc ? 48ull * k : 0ull);
        track(CHROMO_TANGENT_ROTATION, acc);
    }

    // SimpleControl.update_move_amplitude mc_controller.py:148-213 (lane 0)
    __device__ void update_amplitudes(int mtype) {
        if (lane != 0) return;
        chromo_move_state &mv = S.mv[mtype];
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

    __device__ void step(int mtype) {
        if (mtype == CHROMO_TANGENT_ROTATION) tangent_move();
        else if (mtype == CHROMO_CHANGE_BINDING_STATE) binding_move();
        else segment_move(mtype);
        __syncwarp();
    }
};

// ------------------------------------------------------------------ kernels
__device__ __forceinline__ HashTable carve_table(unsigned char *dyn, int cap, int ncol) {
    HashTable H;
    H.vals = (double *)dyn;
    H.keys = (int *)(dyn + (size_t)cap * ncol * sizeof(double));
    H.list = H.keys + cap;
    H.cap = cap;
    int lg = 0;
    while ((1 << lg) < cap) lg++;
    H.shift = 32 - lg;
    H.limit = cap - cap / 4 - 32;
    return H;
}
static_assert(sizeof(WarpSh) <= 2048, "update kWarpShBytes in chromo_b200.cu");

template <class Rng>
__device__ __forceinline__ void rng_load(Rng &rng, const DevCtx &C, WarpSh &S, int rep, int lane,
                                         unsigned long long seed);
template <>
__device__ __forceinline__ void rng_load<ReplayRng>(ReplayRng &rng, const DevCtx &C, WarpSh &S, int rep,
                                                    int lane, unsigned long long) {
    for (int i = lane; i < CB_GLIBC_WORDS; i += 32) S.grs[i] = C.glibc[(long long)rep * CB_GLIBC_WORDS + i];
    rng.st = S.grs;
    rng.mt = C.mt + (long long)rep * CB_MT_WORDS;
    __syncwarp();
}
template <>
__device__ __forceinline__ void rng_load<PhiloxRng>(PhiloxRng &rng, const DevCtx &C, WarpSh &, int rep, int,
                                                    unsigned long long seed) {
    rng.k0 = (uint32_t)seed;
    rng.k1 = (uint32_t)(seed >> 32);
    rng.rep = (uint32_t)rep;
    rng.ctr = C.philox_ctr[rep];
    rng.have = 0;
}
template <class Rng>
__device__ __forceinline__ void rng_store(Rng &rng, const DevCtx &C, WarpSh &S, int rep, int lane);
template <>
__device__ __forceinline__ void rng_store<ReplayRng>(ReplayRng &, const DevCtx &C, WarpSh &S, int rep,
                                                     int lane) {
    __syncwarp();
    for (int i = lane; i < CB_GLIBC_WORDS; i += 32) C.glibc[(long long)rep * CB_GLIBC_WORDS + i] = S.grs[i];
}
template <>
__device__ __forceinline__ void rng_store<PhiloxRng>(PhiloxRng &rng, const DevCtx &C, WarpSh &, int rep,
                                                     int lane) {
    if (lane == 0) C.philox_ctr[rep] = rng.ctr;
}

// mc_sim mc_sim.pyx:26-103 for every replica: grid = R blocks of one warp.
template <class Rng, int NB>
__global__ void __launch_bounds__(32, 8) mc_sim_kernel(DevCtx C, long long num_mc_steps, double mu_adjust,
                                                    unsigned long long seed, int cap) {
    CB_DYN_SMEM(dyn);
    __shared__ WarpSh S;
    const int rep = blockIdx.x, lane = threadIdx.x;
    if (rep >= C.R) return;
    HashTable H = carve_table(dyn, cap, C.ncol);
    table_reset_all(H, S, C.ncol, lane);
    if (lane < CHROMO_NUM_MOVES) S.mv[lane] = C.moves[(long long)rep * CHROMO_NUM_MOVES + lane];
    if (lane == 0) {
        S.algo_bytes = 0;
        S.last_U = 0;
    }
    Rng rng;
    rng_load<Rng>(rng, C, S, rep, lane, seed);
    __syncwarp();
    McWarp<Rng, false, NB> W{C, S, H, rng, rep, lane, mu_adjust, -1, nullptr};
    long long a0 = 0;
    for (int m = 0; m < CHROMO_NUM_MOVES; m++) a0 += S.mv[m].num_attempt;
    for (long long k = 0; k < num_mc_steps; k++)
        for (int m = 0; m < CHROMO_NUM_MOVES; m++) {
            if (S.mv[m].move_on == 1)
                for (int j = 0; j < S.mv[m].num_per_cycle; j++) W.step(m);
            W.update_amplitudes(m); // also for moves that are off (mc_sim.pyx:103)
            __syncwarp();
        }
    __syncwarp();
    long long a1 = 0;
    for (int m = 0; m < CHROMO_NUM_MOVES; m++) a1 += S.mv[m].num_attempt;
    if (lane == 0) {
        C.attempts[rep] = (unsigned long long)(a1 - a0);
        C.algo_bytes[rep] = S.algo_bytes;
    }
    if (lane < CHROMO_NUM_MOVES) C.moves[(long long)rep * CHROMO_NUM_MOVES + lane] = S.mv[lane];
    rng_store<Rng>(rng, C, S, rep, lane);
}

// one instrumented mc_step of one replica (chromo_mc_step)
template <class Rng, int NB>
__global__ void __launch_bounds__(32) mc_step_kernel(DevCtx C, int rep, int mtype, double amp_move,
                                                     int amp_bead, double mu_adjust,
                                                     unsigned long long seed, int force_accept,
                                                     DebugOut *dbg, int cap) {
    CB_DYN_SMEM(dyn);
    __shared__ WarpSh S;
    const int lane = threadIdx.x;
    HashTable H = carve_table(dyn, cap, C.ncol);
    table_reset_all(H, S, C.ncol, lane);
    if (lane < CHROMO_NUM_MOVES) S.mv[lane] = C.moves[(long long)rep * CHROMO_NUM_MOVES + lane];
    __syncwarp();
    if (lane == 0) {
        S.mv[mtype].amp_move = amp_move;
        S.mv[mtype].amp_bead = amp_bead;
        dbg->n_inds = 0;
        dbg->n_touched = 0;
        dbg->dE_poly = dbg->dE_field = 0.0;
        dbg->accepted = 0;
        dbg->passes = 0;
        S.algo_bytes = 0;
        S.last_U = 0;
    }
    Rng rng;
    rng_load<Rng>(rng, C, S, rep, lane, seed);
    __syncwarp();
    McWarp<Rng, true, NB> W{C, S, H, rng, rep, lane, mu_adjust, force_accept, dbg};
    W.step(mtype);
    __syncwarp();
    rng_store<Rng>(rng, C, S, rep, lane);
}
