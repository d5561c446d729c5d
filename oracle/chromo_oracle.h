/*
 * chromo_oracle.h -- CPU restatement (plain C) of the reference's Monte-Carlo
 * energy-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker, never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (chromo_b200/) never includes, links or calls anything here.
 *
 * Parity status: PINNED.  Every function below is checked against the
 * reference itself (the unmodified Cython build in oracle/_ref, driven by
 * tests/golden/make_golden.py in the authoring container) and against the
 * committed golden vectors in tests/golden/ (tests/test_oracle_*.py).
 *
 * All citations are file:line in /root/reference/chromo/.
 */
#ifndef CHROMO_ORACLE_H
#define CHROMO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { OC_CRANK = 0, OC_PIVOT = 1, OC_SLIDE = 2, OC_TANGENT = 3, OC_BINDING = 4,
       OC_NMOVES = 5 };
enum { OC_CONFINE_NONE = 0, OC_CONFINE_SPHERICAL = 1, OC_CONFINE_CUBICAL = 2 };

/* glibc rand() (TYPE_3 additive feedback) -- SURVEY Appendix B */
typedef struct {
    int32_t r[34];
    int f, b;
    void *alt;   /* NULL: the reference's streams.  Else an oc_philox: the product's
                    production streams (see below), every draw comes from there */
} oc_glibc_rand;

/* The product's PRODUCTION random streams, restated (chromo_b200/csrc/rng.cuh,
 * PhiloxRng): draw i of attempt t of replica r is word (i & 3) of
 * Philox4x32-10(counter = (t_lo, t_hi, i >> 2, r), key = seed).  Draw 0 of an
 * attempt is its Metropolis uniform, the proposal's draws follow in the
 * reference's order; binding states come from the same
 * stream (masked rejection on 32-bit words) instead of numpy's MT19937, and
 * sphere points use the closed form cos(theta) = 2u-1 (the reference calls
 * acos).  This is NOT the reference's RNG: it exists so that the kernel that is
 * benchmarked can be replayed attempt for attempt on the CPU. */
typedef struct {
    uint32_t k0, k1, rep, pos;
    uint64_t attempt;       /* stream in use */
    uint64_t next_attempt;  /* attempts this replica has made (persists across mc_sim calls) */
    uint32_t blk[4];
} oc_philox;

/* numpy legacy RandomState (MT19937) -- move_funcs.pyx:819 np.random.randint */
typedef struct {
    uint32_t mt[624];
    int pos;
} oc_mt19937;

/* one controller + MCAdapter + AcceptanceTracker (moves.pyx:58-135,
 * mc_controller.py:89-213, mc_stat.py:72,190-207) */
typedef struct {
    int64_t move_on;
    int64_t num_per_cycle;
    double amp_move;
    int64_t amp_bead;
    int64_t num_attempt;
    int64_t num_success;
    double acceptance_rate;
    double alpha;             /* 2/(moves_in_average+1) */
    double move_amp_lo, move_amp_hi;
    double bead_amp_lo, bead_amp_hi;   /* kept as double: end_pivot's lower
                                          bound is min(50, N/4), a float */
    int64_t controller;       /* 0 = NoControl, 1 = SimpleControl */
} oc_move;

/* DetailedChromatin (polymers.pyx:2455-2607): every nucleosome has fixed entry / exit positions and an exit
 * frame in its own frame (beads.py:448-574, util/nucleo_geom.py); constants of one value of bp_wrap */
typedef struct {
    double t3_local[3], t2_local[3];      /* beads.py:493-502 */
    double r_enter_unit[3], r_enter_norm; /* beads.py:503-515 */
    double r_exit_unit[3], r_exit_norm;
    double a3[3];                         /* T3_exit = a3[0] T3 + a3[1] T2 + a3[2] T1   (nucleo_geom.get_T3) */
    double a1[3];                         /* T1_exit = a1[0] T3 + a1[1] T2 + a1[2] T1   (nucleo_geom.get_T1) */
} oc_detailed;

/* polymer + field of ONE replica; all arrays caller-owned (numpy) */
typedef struct {
    /* --- polymer (polymers.pxd:12-32) --- */
    int64_t N, nb;
    double *r, *t3, *t2;                     /* [N,3] C-contiguous */
    double *r_trial, *t3_trial, *t2_trial;   /* [N,3] */
    int64_t *states, *states_trial, *mods;   /* [N,nb] */
    double *eps_bend, *eps_par, *eps_perp, *gamma, *eta;   /* [N-1] */
    int64_t max_binders;
    double mu_adjust_factor;
    double bead_vol;                         /* beads[0].vol, beads.py:415 */
    /* --- binders (binders.pyx:51-131) --- */
    int64_t *sites_per_bead;                 /* [nb] */
    double *bind_energy_mod, *bind_energy_no_mod, *chemical_potential; /* [nb] */
    double *field_pref;                      /* [nb]  fields.pyx:696-700 */
    double *e_intra;                         /* [nb]  fields.pyx:701-704 */
    double *xpref;                           /* [nb,nb] fields.pyx:705-712 */
    /* --- field (fields.pxd:40-66) --- */
    int64_t field_active;                    /* 0 = NullField */
    int64_t nx, ny, nz, n_bins;
    double width[3], dxyz[3], half_width[3], half_step[3];
    double vol_bin;
    double *access_vol;                      /* [n_bins] */
    double *density, *density_trial;         /* [n_bins, nb+1] */
    int64_t *affected;                       /* [n_bins] 0/1 */
    int64_t confine_type;
    double confine_length;
    double chi;
    float vf_limit;                          /* C float: fields.pxd:61 */
    /* --- scratch for get_change_in_density --- */
    int64_t *touched;                        /* [n_bins] first-touch order */
    int64_t n_touched;
    int64_t *touch_stamp;                    /* [n_bins] */
    int64_t stamp;
    /* --- RNG --- */
    oc_glibc_rand crng;
    oc_mt19937 mt;
    /* --- last move bookkeeping --- */
    double last_dE_poly, last_dE_field;
    int64_t last_accept;
    double last_u;
    /* --- SSTWLC twist (polymers.pyx:1889-2319); NULL for SSWLC / Chromatin --- */
    double *eps_twist;                       /* [N-1] lt / (delta * lp), polymers.pyx:2000 */
    double *twist0;                          /* [N-1] bead_length * NATURAL_TWIST_BARE / LENGTH_BP, 2088-2090 */
    /* --- production streams (crng.alt points here when in use) --- */
    oc_philox philox;
    /* --- fast_field (fields.pyx:577-671, 1235-1368): sub-bins per voxel edge, 0 = exact binning --- */
    int64_t fast_n_points;
    /* --- DetailedChromatin: NULL for every other polymer class --- */
    const oc_detailed *detailed;
} oc_sim;

/* RNG */
void oc_srand(oc_glibc_rand *s, uint32_t seed);
int32_t oc_rand(oc_glibc_rand *s);
void oc_mt_seed(oc_mt19937 *s, uint32_t seed);
uint32_t oc_mt_next(oc_mt19937 *s);
int64_t oc_mt_randint(oc_mt19937 *s, int64_t low, int64_t high);
void oc_philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]);
void oc_philox_init(oc_sim *s, uint64_t seed, uint32_t replica, uint64_t next_attempt);

/* A1: per-bead binning */
void oc_bin_point(const oc_sim *s, const double xyz[3], int64_t idx[8], double w[8]);

/* A8 */
void oc_update_all_densities(oc_sim *s, int for_all_polymers);
double oc_field_E(oc_sim *s);
double oc_poly_E(const oc_sim *s);

/* A1-A6 */
double oc_field_dE(oc_sim *s, const int64_t *inds, int64_t n, int state_change);
/* A7 */
void oc_update_affected_densities(oc_sim *s);
void oc_nucleosome_frames(const oc_detailed *d, const double r[3], const double t3[3], const double t2[3],
                          double r_enter[3], double r_exit[3], double t3_exit[3], double t2_exit[3]);
/* A9, A10 */
double oc_poly_dE(oc_sim *s, int move, const int64_t *inds, int64_t n);
double oc_binding_free_energy(int64_t Nn, int64_t Nm, int64_t s, double e_mod, double e_nomod);

/* A11 */
void oc_rotation_matrix(const double axis[3], const double point[3], double ang, double m[16]);
void oc_transform_rows(oc_sim *s, const double m[16], const int64_t *inds, int64_t n);
int64_t oc_from_point(oc_glibc_rand *g, int64_t window, int64_t N, int64_t ind0);
int64_t oc_from_left(oc_glibc_rand *g, int64_t window, int64_t N);
int64_t oc_from_right(oc_glibc_rand *g, int64_t window, int64_t N);
int64_t oc_propose(oc_sim *s, int move, double amp_move, int64_t amp_bead, int64_t *inds_out);

/* A12 */
void oc_accept(oc_sim *s, oc_move *mv, int move, const int64_t *inds, int64_t n);
void oc_reject(oc_sim *s, oc_move *mv, int move, const int64_t *inds, int64_t n);
void oc_update_amplitudes(oc_move *mv);
int oc_mc_step(oc_sim *s, oc_move *mv, int move, int64_t *inds_scratch);
void oc_mc_sim_ordered(oc_sim *s, oc_move mv[OC_NMOVES], int64_t num_mc_steps, uint32_t mt_seed,
                       int64_t *inds, const int32_t *order); /* order: the controller list's move ids, or NULL */
void oc_mc_sim(oc_sim *s, oc_move mv[OC_NMOVES], int64_t num_mc_steps, uint32_t mt_seed,
               int64_t *inds_scratch);

#ifdef __cplusplus
}
#endif
#endif
