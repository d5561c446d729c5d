// params.cuh -- device-side view of a context (R replicas resident in HBM).
//
// HBM layout (one context = one GPU):
//   r, t3, t2   [R][N][3]  fp64   AoS xyz, the reference's own layout
//                                 (polymers.pxd:21) so a move segment
//                                 [ind0, indf) is ONE contiguous 24n-byte run
//                                 per array -> fully coalesced warp loads
//   states,mods [R][N][nb] int8   (int64 in the reference; values 0..sites)
//   density     [R][n_bins][nb+1] fp64, row = one voxel (fields.pxd:54), so a
//                                 touched voxel is one 16/24/32-byte gather
//   bond        [sets][N-1][5]    eps_bend, eps_par, eps_perp, gamma, eta
//   twist       [sets][N-1][2]    eps_twist, natural twist (SSTWLC only)
//   moves       [R][5]            controller / tracker state
#pragma once
#include <cstdint>
#include "../../include/chromo_b200.h"

#define CB_MAXNB CHROMO_MAX_BINDERS
#define CB_E_HUGE_FIELD 1E99 /* fields.pyx:32 */
#define CB_E_HUGE_POLY 1E25  /* polymers.pyx:35 */
#define CB_RAND_MAX 2147483647.0

struct DevCtx {
    int R, N, nb, ncol;
    int nx, ny, nz, n_bins;
    int field_active, confine_type;
    double width[3], dxyz[3], half_width[3], half_step[3];
    double inv_width[3], inv_dxyz[3]; // correctly rounded reciprocals, for div_const
    double vol_bin, inv_vol_bin, bead_vol, confine_length, vf_limit;
    long long max_binders;
    double *r, *t3, *t2;
    signed char *states, *mods;
    double *density;
    const double *access_vol; // nullptr -> vol_bin
    const double *bond;
    long long bond_stride; // 0 (shared) or (N-1)*5
    const double *twist;   // SSTWLC: [sets][N-1][2] eps_twist, natural twist per bond; nullptr = no twist term
    long long twist_stride; // 0 (shared) or (N-1)*2
    const double *detailed; // DetailedChromatin: 20 nucleosome constants (geometry.cuh nucleosome_frame), nullptr = off
    const double *chi;     // [R]
    const double *mu;      // [R][nb]
    double pref[CB_MAXNB], e_intra[CB_MAXNB], xpref[CB_MAXNB * CB_MAXNB];
    int sites[CB_MAXNB];
    int any_cross; // some cross-talk prefactor is non-zero
    int fx_base;   // floor(log2(V_min / max_state)): exponent base of the fixed-point delta-density cells (fx_format)
    const double *bindF; // [nb][S1][S1]
    int S1;
    chromo_move_state *moves;        // [R][5]
    uint32_t *glibc;                 // [R][GLIBC_WORDS]
    uint32_t *mt;                    // [R][MT_WORDS]
    unsigned long long *philox_ctr;  // [R]
    int fast_n;                      // fast_field: sub-bins per voxel edge (even), 0 = exact binning
    double fast_sbw[3];              // fast_field: sub-bin widths dx / n_points (fields.pyx:604-606)
    unsigned rep_offset;             // global index of replica 0 (key of the production streams)
    int batch;                       // attempts prepared at once in the production kernels (1..32)
    int type_mask;                   // bit m: mc_sim_kernel runs move type m (31 = all; see launch_sim, chromo_b200.cu)
    int accumulate;                  // add this launch's attempt / byte counters to the stored ones
    int *tan_inds;                   // [R][N]   tangent-rotation large path
    uint32_t *sel_bits;              // [R][ceil(N/32)]
    signed char *st_new;             // [R][N]   binding large path
    unsigned long long *attempts;    // [R]
    unsigned long long *algo_bytes;  // [R]   algorithmic bytes of the last mc_sim (SURVEY 8d formula)
};

#define CB_GLIBC_WORDS 36 // r[31], f, b (+pad)
#define CB_MT_WORDS 628   // mt[624], pos (+pad)

// instrumentation written by the single-step parity kernel (chromo_mc_step)
struct DebugOut {
    long long n_inds, n_touched;
    double dE_poly, dE_field, u;
    int accepted, passes;
    long long inds_cap, rows_cap, touched_cap;
    long long *inds;    // [inds_cap]
    double *rows;       // [rows_cap][9+nb]
    long long *touched; // [touched_cap]
    double *dtrial;     // [touched_cap][ncol]
};
