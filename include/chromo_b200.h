/*
 * chromo_b200.h -- C ABI of the B200-native Monte-Carlo energy-evaluation path.
 *
 * This is the drop-in boundary for chromo's hot path.  The
 * reference has no FFI for this path: the seam is the Cython entry point
 *
 *   cpdef void mc_sim(list polymers, readerproteins, long num_mc_steps,
 *                     list mc_move_controllers, FieldBase field,
 *                     double mu_adjust_factor, long random_seed)
 *                                   (chromo/mc/mc_sim.pyx:26-30, mc_sim.pxd:11-15)
 *
 * plus the fine-grained cdef/cpdef methods it calls.  Every entry point below
 * names the reference interface it replaces (file:line under
 * /root/reference/chromo/).  INTEGRATION.md shows the ctypes stub a chromo
 * maintainer would add to bind them.
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature;
 *   - a context (`chromo_ctx`) holds R independent replicas of one problem
 *     shape (N beads, nb binders, one nx*ny*nz grid) resident in HBM on ONE
 *     device; replica i of a multi-GPU job lives on rank i % world_size;
 *   - all array arguments are caller-owned HOST buffers unless the name ends in
 *     `_dev`; layouts are the reference's (C-contiguous, fp64 / int64);
 *   - every function returns 0 on success or a negative code; the message is
 *     available from chromo_last_error() (thread-local);
 *   - one CUDA stream per context; calls on one context are not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with CHROMO_ERR_CUDA.
 */
#ifndef CHROMO_B200_H
#define CHROMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHROMO_OK 0
#define CHROMO_ERR_ARG -1
#define CHROMO_ERR_CUDA -2
#define CHROMO_ERR_STATE -3

#define CHROMO_MAX_BINDERS 4

/* move order = controller order in mc_sim (mc_controller.py:255-258) */
enum chromo_move_id {
    CHROMO_CRANK_SHAFT = 0,          /* move_funcs.pyx:39  */
    CHROMO_END_PIVOT = 1,            /* move_funcs.pyx:285 */
    CHROMO_SLIDE = 2,                /* move_funcs.pyx:403 */
    CHROMO_TANGENT_ROTATION = 3,     /* move_funcs.pyx:470 */
    CHROMO_CHANGE_BINDING_STATE = 4, /* move_funcs.pyx:717 */
    CHROMO_NUM_MOVES = 5
};

enum chromo_confine { CHROMO_CONFINE_NONE = 0, CHROMO_CONFINE_SPHERICAL = 1,
                      CHROMO_CONFINE_CUBICAL = 2 }; /* fields.pyx:160-202 */

/* RNG behind the proposals and the Metropolis test.
 *   PHILOX: production; Philox4x32-10 keyed by (seed, replica).
 *   REPLAY: restates the reference's generators -- glibc rand() (never seeded
 *           by the reference, mc_sim.pyx:52-56) and numpy's legacy MT19937
 *           (np.random.seed(random_seed), mc_sim.pyx:81) -- so that a run can
 *           be compared draw for draw with the reference. */
enum chromo_rng_mode { CHROMO_RNG_PHILOX = 0, CHROMO_RNG_REPLAY = 1 };

/* One MCAdapter + its controller + its AcceptanceTracker, per replica and move
 * (moves.pyx:58-135; mc_controller.py:89-213; mc_stat.py:72,190-207). In/out. */
typedef struct chromo_move_state {
    double amp_move;        /* MCAdapter.amp_move */
    double move_amp_lo;     /* Controller.move_amp_bounds */
    double move_amp_hi;
    double bead_amp_lo;     /* Controller.bead_amp_bounds (may be fractional:  */
    double bead_amp_hi;     /*  end_pivot lower bound is min(50, N/4))         */
    double acceptance_rate; /* AcceptanceTracker.acceptance_rate (EWMA)        */
    double alpha;           /* 2 / (moves_in_average + 1)                      */
    int64_t num_attempt;    /* MCAdapter.num_attempt */
    int64_t num_success;    /* MCAdapter.num_success */
    int32_t amp_bead;       /* MCAdapter.amp_bead */
    int32_t num_per_cycle;  /* MCAdapter.num_per_cycle */
    int32_t move_on;        /* MCAdapter.move_on */
    int32_t controller;     /* 0 = NoControl, 1 = SimpleControl */
} chromo_move_state;

/* Problem shape shared by all replicas of a context.
 * Replaces the constructor state of UniformDensityField (fields.pyx:453-575)
 * and PolymerBase/SSWLC (polymers.pyx:186-300, 959-1012). */
typedef struct chromo_shape {
    int64_t n_replicas;
    int64_t num_beads;      /* N */
    int64_t num_binders;    /* nb >= 1 (null_reader counts, polymers.pyx:322) */
    int64_t nx, ny, nz;     /* 0,0,0 = NullField (fields.pyx:280) */
    double width[3];        /* x_width, y_width, z_width */
    int32_t confine_type;   /* enum chromo_confine */
    double confine_length;
    float vf_limit;         /* stored as C float by the reference (fields.pxd:61) */
    double bead_vol;        /* beads[0].vol = 4/3*pi*rad^3 (beads.py:415) */
    int64_t max_binders;    /* PolymerBase.max_binders, -1 = no limit */
} chromo_shape;

typedef struct chromo_ctx chromo_ctx;

const char *chromo_last_error(void);
int chromo_version(void);

/* ---- lifetime ---------------------------------------------------------- */
int chromo_ctx_create(chromo_ctx **out, int device, const chromo_shape *shape);
int chromo_ctx_destroy(chromo_ctx *ctx);
int chromo_ctx_sync(chromo_ctx *ctx);
/* the context's cudaStream_t, for CUDA-event timing by the caller */
void *chromo_ctx_stream(chromo_ctx *ctx);
/* bytes of HBM held by the context */
int64_t chromo_ctx_bytes(chromo_ctx *ctx);
/* Tuning knob: slots (a multiple of 32, 128..4096) of the per-warp shared-memory
 * delta-density hash used by the MC kernel; 0 = choose from the replica count so
 * that all replicas are resident at once.  Moves that touch more voxels than fit
 * are evaluated in several hash-partition passes (same result).  Returns the
 * capacity in effect through *cap_out (may be NULL). */
int chromo_ctx_set_table_capacity(chromo_ctx *ctx, int64_t cap, int64_t *cap_out);
/* Tuning knob: warps that work on one replica in the production (Philox) MC
 * kernel, 1 or 2 (0 = leave unchanged; default 2 while at most 4 replicas share an SM,
 * i.e. n_replicas <= 4 x SM count, else 1).  With 2, the bead-row stage of
 * attempt j+1 overlaps the density stage of attempt j; attempts still take
 * effect strictly in order and the results are bit-identical to 1 warp (on
 * B200 the second warp costs more in instruction-cache misses than the overlap
 * gains at 7 replicas per SM, and gains 4-40 % when 1-4 replicas share an SM).  The
 * replayed-RNG kernels (sequential streams) always use 1.  Re-chooses the table
 * capacity.  Returns the value in effect through *warps_out (may be NULL). */
int chromo_ctx_set_warps_per_replica(chromo_ctx *ctx, int64_t warps, int64_t *warps_out);
/* Tuning knob: replicas that share one thread block of the MC kernel, 1..7
 * (-1 = automatic: ceil(replicas / SMs), at most 7; 0 = leave unchanged).  The
 * replicas of a block stay independent simulations but enter each move type of
 * a sweep together, so that their warps share instruction-cache lines.  Results
 * do not depend on it.  Re-chooses the table capacity.  Returns the value in
 * effect through *rpb_out (may be NULL). */
int chromo_ctx_set_replicas_per_block(chromo_ctx *ctx, int64_t rpb, int64_t *rpb_out);
/* Index of this context's replica 0 in the whole ensemble.  The production
 * random streams are keyed by (seed, GLOBAL replica index, attempt): a shard of
 * a multi-GPU ensemble must set its offset (chromo_b200.parallel does), or
 * replica j of every shard would draw the same numbers.  Default 0. */
int chromo_ctx_set_replica_offset(chromo_ctx *ctx, int64_t offset);
/* Test / debugging knob: attempts of one move type whose state-independent half
 * is prepared at once by the lanes of a warp in the production kernels, 1..32
 * (default 32).  Results must not depend on it (tests/test_philox_parity.py). */
int chromo_ctx_set_batch_size(chromo_ctx *ctx, int64_t batch);
/* DetailedChromatin (polymers.pyx:2455-2607): nucleosomes with fixed entry / exit points and an exit frame
 * (DetailedNucleosome beads.py:448-574).  consts20 = t3_local[3] | t2_local[3] | r_enter_unit[3] | r_enter_norm |
 * r_exit_unit[3] | r_exit_norm | a3[3] | a1[3] for one bp_wrap (chromo_b200.util.nucleo_geom.nucleosome_constants):
 * every bond energy of the elastic dE then runs from the exit of one bead to the entry of the next.  Needs the
 * twist parameters (it is an SSTWLC); NULL switches it off.  DetailedChromatin2 (polymers.pyx:2627-2735, bonds between
 * the bead centres): the same constants with both norms 0.  compute_E is the SSTWLC one, as in the reference. */
int chromo_set_detailed_nucleosomes(chromo_ctx *ctx, const double *consts20);
/* Order in which chromo_mc_sim goes through the move types within one MC step: the reference walks its
 * controller LIST (`for controller in mc_move_controllers`, mc_sim.pyx:92-103), whatever order the caller built
 * it in.  `order` is a permutation of the CHROMO_* move ids; default 0,1,2,3,4 (mc_controller.all_moves).
 * The default order is one kernel launch per chromo_mc_sim call; any other order is one launch per (MC step,
 * move type), state carried on the device -- same results, ~10 us of launch overhead per phase. */
int chromo_ctx_set_move_order(chromo_ctx *ctx, const int32_t order[CHROMO_NUM_MOVES]);
/* Page-lock (cudaHostRegister) a caller-owned host array -- the polymers' r / t3 / t2 / states buffers -- so that
 * chromo_mc_sim_host, chromo_upload_state and chromo_download_state move it at link speed instead of staging it
 * through the driver's bounce buffers (pageable numpy memory: ~3x slower).  The reference has no counterpart (its
 * arrays never leave the host).  CHROMO_ERR_STATE if the range is already page-locked (e.g. torch pinned memory):
 * nothing to do, and nothing to unregister.  Unregister before the array is freed. */
int chromo_host_register(void *host_ptr, uint64_t bytes);
int chromo_host_unregister(void *host_ptr);
/* fast_field = 1 of UniformDensityField (init_fast_field fields.pyx:577-671, get_change_in_density_quickly
 * 1235-1368): the dE path bins positions quantised to n_points sub-bins per voxel edge (rounded up to an even
 * number, as the reference does) and adds every term (no 1e-18 filter).  0 = exact binning (default).  The full
 * recompute and compute_E are not affected (they are not in the reference either). */
int chromo_ctx_set_fast_field(chromo_ctx *ctx, int64_t n_points);
/* Attempts every replica has made so far = position of its production random
 * stream (counters[n] for replicas [first, first+n)).  Saved and restored with a
 * snapshot so that a resumed run continues the stream instead of replaying it. */
int chromo_get_rng_counters(chromo_ctx *ctx, int64_t first, int64_t n, uint64_t *counters);
int chromo_set_rng_counters(chromo_ctx *ctx, int64_t first, int64_t n, const uint64_t *counters);

/* ---- parameters -------------------------------------------------------- */
/* Reader-protein tables.
 *  sites_per_bead[nb]                    binders.pyx:38-49
 *  field_pref[nb], e_intra[nb], xpref[nb*nb]
 *                                        init_field_energy_prefactors fields.pyx:687-712
 *  bind_F[nb][S+1][S+1]  with S = max sites_per_bead:
 *      bind_F[b][Nm][s] = -ln sum_i C(Nm,i) C(Nn-Nm,s-i) exp(-(i e_mod + (s-i) e_nomod))
 *                                        bead_binding_dE polymers.pyx:1493-1517 */
int chromo_set_binders(chromo_ctx *ctx, const int64_t *sites_per_bead, const double *field_pref,
                       const double *e_intra, const double *xpref, const double *bind_F,
                       int64_t S);
/* Per-replica sweep parameters: chi[R] (UniformDensityField.chi, fields.pyx:526)
 * and chemical_potential[R][nb] (ReaderProtein.chemical_potential). */
int chromo_set_replica_params(chromo_ctx *ctx, const double *chi, const double *mu);
/* SSWLC bond parameters [n_sets][N-1] each (SSWLC._find_parameters
 * polymers.pyx:1545-1601); n_sets is 1 (shared by all replicas) or R. */
int chromo_set_bond_params(chromo_ctx *ctx, int64_t n_sets, const double *eps_bend,
                           const double *eps_par, const double *eps_perp, const double *gamma,
                           const double *eta);
/* SSTWLC._find_parameters / E_pair_with_twist (polymers.pyx:1957-2001, 2050-2102): per bond the twist
 * modulus eps_twist = lt / (delta * lp) and the natural twist bead_length * NATURAL_TWIST_BARE /
 * LENGTH_BP; [n_sets][N-1] each, n_sets = 1 (shared) or n_replicas.  Once set, every bond energy of
 * chromo_mc_sim / chromo_mc_step / chromo_elastic_energy carries the twist term
 * 0.5 * eps_twist * wrap(omega - natural_twist)^2 with omega = compute_twist_angle_omega
 * (polymers.pyx:3427-3461).  Both pointers NULL: back to a chain without twist.  One or two binders. */
int chromo_set_twist_params(chromo_ctx *ctx, int64_t n_sets, const double *eps_twist,
                            const double *natural_twist);

/* access_vols[n_bins] (fields.pyx:523-525, 714-770); NULL = vol_bin everywhere */
int chromo_set_access_volumes(chromo_ctx *ctx, const double *access_vol);

/* ---- state ------------------------------------------------------------- */
/* r, t3, t2: [n][N][3] fp64; states, chemical_mods: [n][N][nb] int64
 * (polymers.pxd:21-24) for replicas first .. first+n-1.  NULL = leave as is. */
int chromo_upload_state(chromo_ctx *ctx, int64_t first, int64_t n, const double *r,
                        const double *t3, const double *t2, const int64_t *states,
                        const int64_t *chemical_mods);
int chromo_download_state(chromo_ctx *ctx, int64_t first, int64_t n, double *r, double *t3,
                          double *t2, int64_t *states);
/* density[n][n_bins][nb+1] (UniformDensityField.density, fields.pxd:54) */
int chromo_download_density(chromo_ctx *ctx, int64_t first, int64_t n, double *density);
int chromo_upload_density(chromo_ctx *ctx, int64_t first, int64_t n, const double *density);

/* ---- full recompute (A8) ---------------------------------------------- */
/* update_all_densities (fields.pyx:1977-2039) for every replica; clamp != 0
 * also applies the |rho| < 1e-18 -> 0 pass of
 * update_all_densities_for_all_polymers (fields.pyx:2041-2106). */
int chromo_field_recompute(chromo_ctx *ctx, int clamp);
/* UniformDensityField.compute_E (fields.pyx:1939-1966, 2208-2315): recomputes
 * the densities, then returns per replica
 *   sum_sq[R][nb]   = sum_bins rho_a^2
 *   doubly[R][nb]   = # beads with state == 2
 *   nonspecific[R]  = sum_bins (round(phi,2) > vf_limit ? 1e99 phi : chi (V/v) phi (1-phi))
 *   E[R]            = sum_a (pref_a sum_sq_a + e_intra_a doubly_a) + nonspecific  */
int chromo_field_energy(chromo_ctx *ctx, double *E, double *sum_sq, int64_t *doubly,
                        double *nonspecific);
/* SSWLC.compute_E (polymers.pyx:1348-1381): E[R] */
int chromo_elastic_energy(chromo_ctx *ctx, double *E);
/* chi-conjugate observable used by the replica-exchange step (new functionality,
 * SURVEY.md 8e): Phi[R] = sum_bins (V/v) phi^2, i.e. dE_chi-term / chi
 * (nonspecific_interact_dE fields.pyx:1829-1840).  Uses the current densities. */
int chromo_chi_observable(chromo_ctx *ctx, double *Phi);

/* ---- replica exchange on the device (SURVEY.md 8b `exchange_step`, 8e) ---
 * Parallel tempering over chi for an ensemble sharded over GPUs: replica g of
 * n_total lives in the context that covers [first, first + R).  The reference has
 * no such step (its runs are independent processes); this is BASELINE config 5.
 * One round, all on the context's stream and without any host round trip:
 *   1. chromo_exchange_observable(ctx, phi_local_dev)   Phi of the local replicas
 *      into a caller-owned DEVICE buffer [R];
 *   2. the caller all-gathers the shards into phi_all_dev [n_total] (NCCL over
 *      NVLink: 8 B per replica; chromo_b200.parallel does it with
 *      torch.distributed on the same stream);
 *   3. chromo_exchange_step(ctx, phi_all_dev, round, seed)   every rank decides the
 *      same even / odd neighbour swaps on the ladder (a counter-based uniform per
 *      pair and round), updates its copy of the rung -> replica permutation and
 *      the chi of its own replicas.  Labels move, configurations never do.
 * chromo_exchange_init sets the ladder from chi_by_replica[n_total] (the chi of
 * every GLOBAL replica); replicas [l * ladder_len, (l+1) * ladder_len) form ladder l
 * (ladder_len <= 0: one ladder of n_total rungs). */
int chromo_exchange_init(chromo_ctx *ctx, const double *chi_by_replica, int64_t n_total, int64_t first,
                         int64_t ladder_len);
int chromo_exchange_observable(chromo_ctx *ctx, double *phi_local_dev);
int chromo_exchange_step(chromo_ctx *ctx, const double *phi_all_dev, int64_t round, uint64_t seed);
/* Synchronising read-back for checks: rung_replica[n_total] (global replica on each rung),
 * chi_local[R], pairs tried and swaps accepted since chromo_exchange_init (any may be NULL). */
int chromo_exchange_state(chromo_ctx *ctx, int32_t *rung_replica, double *chi_local, uint64_t *pairs_tried,
                          uint64_t *swaps_accepted);

/* ---- RNG --------------------------------------------------------------- */
/* REPLAY: seed each replica's glibc rand() state, as srand(seed) would. */
int chromo_srand(chromo_ctx *ctx, const uint32_t *seeds /* [R] */);
/* REPLAY: np.random.seed(seed) for each replica (mc_sim.pyx:81 does this once
 * per mc_sim call; chromo_mc_sim does it too when reseed_numpy != 0). */
int chromo_numpy_seed(chromo_ctx *ctx, const uint32_t *seeds /* [R] */);

/* ---- the hot path (A1-A7, A9-A12) -------------------------------------- */
/* mc_sim (mc_sim.pyx:26-103) for all replicas at once: num_mc_steps sweeps of
 * the five move types, num_per_cycle attempts each, Metropolis accept/reject,
 * incremental density update, amplitude controller once per type per sweep.
 *   moves[R][5]      in/out (NULL = keep the context's current controllers)
 *   numpy_seeds[R]   REPLAY only: per-replica np.random.seed value (NULL = keep)
 *   seed             PHILOX key
 * Asynchronous on the context's stream when `moves` is NULL; otherwise the
 * call returns after the controllers have been copied back. */
int chromo_mc_sim(chromo_ctx *ctx, int64_t num_mc_steps, chromo_move_state *moves,
                  double mu_adjust_factor, uint64_t seed, int rng_mode,
                  const uint32_t *numpy_seeds);
int chromo_get_moves(chromo_ctx *ctx, chromo_move_state *moves /* [R][5] */);
int chromo_set_moves(chromo_ctx *ctx, const chromo_move_state *moves /* [R][5] */);
/* The same call on caller-owned HOST arrays, in and out -- what the reference's
 * mc_sim does to the polymers' numpy arrays (mc_sim.pyx:26-103 mutates poly.r, t3, t2,
 * states in place; moves.pyx:156-299).  r / t3 / t2: [R][N][3] f64, states / mods:
 * [R][N][nb] int64, in the reference's layouts; states and the three coordinate
 * arrays are overwritten with the result, mods is only read.  The replicas are
 * processed in `n_chunks` chunks (0 = automatic: as many as keep every chunk's thread
 * blocks resident at once, at most 8), each on its own stream: while chunk k runs,
 * chunk k+1 is still arriving over PCIe and chunk k-1 is already on its way back.
 * Pinned (page-locked) host arrays make the copies asynchronous; pageable ones work,
 * serialised by the driver.  The voxel densities are the field's state, not the
 * polymers' (fields.pxd:54): they stay on the device between calls, exactly as
 * with chromo_upload_state + chromo_mc_sim + chromo_download_state, whose result
 * this call reproduces bit for bit (each replica has its own RNG stream).
 * `mods` may be NULL once every replica's marks have been uploaded (by an earlier
 * call of this function or by chromo_upload_state over all replicas): mc_sim never
 * modifies chemical_mods, so the copy on the device stays current. */
int chromo_mc_sim_host(chromo_ctx *ctx, int64_t num_mc_steps, chromo_move_state *moves,
                       double mu_adjust_factor, uint64_t seed, int rng_mode,
                       const uint32_t *numpy_seeds, double *r, double *t3, double *t2,
                       int64_t *states, const int64_t *mods, int64_t n_chunks);

/* attempts executed by the last chromo_mc_sim call, summed over replicas */
int64_t chromo_last_attempts(chromo_ctx *ctx);
/* algorithmic bytes those attempts needed (per attempt
 * 72(n+2) + 72 n a + nb n + 8(nb+1) U (1+2a) + 80 for segment moves, n beads,
 * U touched voxels, a = accepted), summed over replicas: the numerator of the
 * roofline figure bench.py reports */
int64_t chromo_last_algo_bytes(chromo_ctx *ctx);

/* One mc_step (mc_sim.pyx:106-182) of ONE replica through the same device code
 * as chromo_mc_sim, with everything the reference exposes after a step
 * written back for parity checks:
 *   MCAdapter.propose (moves.pyx:137-154)       -> inds, trial rows
 *   PolymerBase.compute_dE (polymers.pxd:33-38)  -> dE_poly
 *   FieldBase.compute_dE (fields.pxd:16-19)      -> dE_field, touched bins
 *                                                   (affected_bins_last_move),
 *                                                   density_trial rows
 *   accept / reject (moves.pxd:24-33), update_affected_densities (fields.pxd:20)
 * force_accept: -1 = Metropolis with one RNG draw (mc_sim.pyx:171),
 *                0 / 1 = forced reject / accept without consuming a draw. */
typedef struct chromo_step_report {
    int64_t n_inds;
    int64_t n_touched;
    double dE_poly;
    double dE_field;
    double u;             /* Metropolis uniform (NaN when forced) */
    int32_t accepted;
    int32_t passes;       /* hash-partition passes the field dE needed */
} chromo_step_report;

int chromo_mc_step(chromo_ctx *ctx, int64_t replica, int move, double amp_move,
                   int64_t amp_bead, double mu_adjust_factor, int rng_mode, uint64_t seed,
                   int force_accept, chromo_step_report *report,
                   int64_t *inds, int64_t inds_cap,            /* [n_inds] */
                   double *trial_rows, int64_t rows_cap,       /* [n_inds][9+nb]: r,t3,t2,states */
                   int64_t *touched, double *dtrial, int64_t touched_cap /* [n_touched],[..][nb+1] */);

/* ---- coarse-grain / refine, the steps either side of the MC path (chromo/util/rediscretize.py) ----
 * Batched over R replicas; context-free (host arrays in, host arrays out, one launch per call).
 * `kernel_ms` (may be NULL) receives the device time of the kernel alone (CUDA events). */

/* number of coarse-grained beads: len(get_cg_bead_intervals(num_beads, cg_factor)),
 * rediscretize.py:24-54 (floor(N / cg) intervals plus one for the left-over beads); -1 on bad input */
int64_t chromo_cg_num_beads(int64_t num_beads, int64_t cg_factor);

/* get_cg_chromatin rediscretize.py:401-471 for every replica at once:
 *   r_cg      = interval means of r / r_divisor            (get_avg_in_intervals 57-84, line 452;
 *                                                           r_divisor = cg_factor ** (1/3))
 *   t3_cg,t2_cg = normalised interval means of t3, t2 = t3 x e_x (e_y if t3 == e_x), normalised
 *                                                          (get_orientations_in_intervals 87-122)
 *   states_cg, mods_cg = most frequent value per interval and column, smallest on ties
 *                                                          (get_majority_state_in_interval 125-160)
 * r, t3: [R][N][3]; states, mods: [R][N][nb] (either may be NULL); outputs [R][M][.] with
 * M = chromo_cg_num_beads(N, cg_factor). Bit-identical to the reference's numpy arithmetic. */
int chromo_cg_chromatin(int device, int64_t R, int64_t N, int64_t nb, int64_t cg_factor,
                        double r_divisor, const double *r, const double *t3, const int64_t *states,
                        const int64_t *mods, double *r_cg, double *t3_cg, double *t2_cg,
                        int64_t *states_cg, int64_t *mods_cg, double *kernel_ms);

/* rows of get_refined_path(cg_r, num_beads_refined) (rediscretize.py:756-807): num_beads_refined when
 * it is not a multiple of (num_beads_cg - 1), one fewer when it is (get_refined_intervals 565-583 gives
 * the last bridge half a segment minus one); and the number of standard-normal TRIPLES one path draws
 * from numpy's global generator. -1 when the reference itself would fail (fewer than 3 refined beads
 * per coarse bond: brownian_bridge(0, ...) divides by zero). */
int64_t chromo_refined_num_points(int64_t num_beads_cg, int64_t num_beads_refined);
int64_t chromo_refined_num_draws(int64_t num_beads_cg, int64_t num_beads_refined);

/* get_refined_path rediscretize.py:756-807 for every replica at once: free Gaussian-direction ends
 * (gaussian_walk util/poly_paths.py:269-295) and one Brownian bridge per coarse bond (brownian_bridge
 * 586-685) with the average step normalised to bead_spacing.
 *   cg_r [R][M][3]; out [R][P][3], P = chromo_refined_num_points(M, num_beads_refined)
 *   xi   [R][D][3] standard-normal deviates in the reference's draw order (replay of np.random), or
 *        NULL: Philox4x32-10 + Box-Muller keyed by (seed, replica, draw) on the device
 *   orientations = 0: rows are multiplied by out_scale (refine_chromatin 1056)
 *   orientations = 1: get_refined_orientations 810-842: rows normalised, out_t2 completed as in
 *                     chromo_cg_chromatin (pass bead_spacing = pi, the reference's default) */
int chromo_refine_path(int device, int64_t R, int64_t num_beads_cg, int64_t num_beads_refined,
                       double bead_spacing, const double *cg_r, const double *xi, uint64_t seed,
                       double out_scale, int orientations, double *out, double *out_t2,
                       double *kernel_ms);

/* enforce_spherical_confinement rediscretize.py:708-753, in place on r [R][N][3]: every bead outside
 * the sphere pulls itself and its three neighbours on either side inwards by
 * {0.98,0.97,0.96,0.95,0.96,0.97,0.98} / (dist / rad), violations measured on the incoming path */
int chromo_enforce_spherical_confinement(int device, int64_t R, int64_t N, double *r, double rad,
                                         double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* CHROMO_B200_H */
