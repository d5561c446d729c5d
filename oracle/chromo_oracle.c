/*
 * chromo_oracle.c -- sequential CPU restatement of chromo's MC hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see chromo_oracle.h).  Follows the reference's
 * arithmetic order statement by statement so that it can be diffed against
 * the reference's own Cython build (oracle/_ref) bit for bit on binning and
 * to the last few ulp on energies.  Python containers of the reference
 * (dict `access_vols`, set `bins_found`) become plain arrays; the only
 * observable difference is the summation ORDER over touched bins (the
 * reference iterates a Python set; we iterate in first-touch order).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: the reference is built for generic x86-64
 * (no FMA), and a contracted x/dx - ind could flip the last ulp of a weight.
 *
 * Citations: file:line in /root/reference/chromo/.
 */
#include "chromo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OC_RAND_MAX 2147483647
static const double E_HUGE_FIELD = 1E99; /* fields.pyx:32 */
static const double E_HUGE_POLY = 1E25;  /* polymers.pyx:35 */

/* ------------------------------------------------------------------ RNGs */

/* glibc srandom_r/random_r, TYPE_3 (degree 31, separation 3).  The reference
 * draws everything except get_new_state from libc rand()
 * (move_funcs.pyx:14, bead_selection.pyx:10, linalg.pyx:9, mc_sim.pyx:11). */
void oc_srand(oc_glibc_rand *s, uint32_t seed)
{
    int32_t *r = s->r;
    int i;
    s->alt = NULL; /* srand() selects the reference's streams */
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (i = 1; i < 31; i++) {
        long hi = r[i - 1] / 127773;
        long lo = r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    s->f = 3;
    s->b = 0;
    for (i = 0; i < 310; i++) (void)oc_rand(s);
}

/* ---- the product's production streams (chromo_b200/csrc/rng.cuh, PhiloxRng) ---- */
void oc_philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    int i;
    for (i = 0; i < 10; i++) { /* Philox4x32-10 (Salmon et al., SC'11) */
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_seek(oc_philox *p, uint64_t attempt)
{
    p->attempt = attempt;
    p->pos = 0;
}

static uint32_t philox_next32(oc_philox *p)
{
    uint32_t w = p->pos & 3u;
    if (w == 0)
        oc_philox_block((uint32_t)p->attempt, (uint32_t)(p->attempt >> 32), p->pos >> 2, p->rep,
                        p->k0, p->k1, p->blk);
    p->pos++;
    return p->blk[w];
}

/* PhiloxRng::randint: np.random.randint's masked rejection on the stream's 32-bit words */
static int64_t philox_randint(oc_philox *p, int64_t high)
{
    uint32_t rng = (uint32_t)(high - 1), mask = rng, v;
    if (rng == 0) return 0;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    do {
        v = philox_next32(p) & mask;
    } while (v > rng);
    return (int64_t)v;
}

void oc_philox_init(oc_sim *s, uint64_t seed, uint32_t replica, uint64_t next_attempt)
{
    s->philox.k0 = (uint32_t)seed;
    s->philox.k1 = (uint32_t)(seed >> 32);
    s->philox.rep = replica;
    s->philox.next_attempt = next_attempt;
    philox_seek(&s->philox, next_attempt);
    s->crng.alt = &s->philox;
}

int32_t oc_rand(oc_glibc_rand *s)
{
    uint32_t *r = (uint32_t *)s->r;
    if (s->alt) return (int32_t)(philox_next32((oc_philox *)s->alt) >> 1); /* PhiloxRng::next31 */
    {
    uint32_t val = r[s->f] += r[s->b];
    int32_t result = (int32_t)(val >> 1);
    if (++s->f >= 31) s->f = 0;
    if (++s->b >= 31) s->b = 0;
    return result;
    }
}

static double oc_uniform(oc_glibc_rand *g)
{
    /* `<double>rand() / RAND_MAX` everywhere in the reference */
    return (double)oc_rand(g) / OC_RAND_MAX;
}

/* numpy legacy seeding for an integer seed: init_genrand (Knuth) */
void oc_mt_seed(oc_mt19937 *s, uint32_t seed)
{
    int i;
    s->mt[0] = seed;
    for (i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->pos = 624;
}

uint32_t oc_mt_next(oc_mt19937 *s)
{
    uint32_t y;
    if (s->pos == 624) {
        int k;
        uint32_t *mt = s->mt;
        for (k = 0; k < 624 - 397; k++) {
            y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
            mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; k < 623; k++) {
            y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
            mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        s->pos = 0;
    }
    y = s->mt[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* np.random.randint(low, high) on the legacy global RandomState, default
 * dtype int64: masked rejection on 32-bit draws when the range fits 32 bits
 * (numpy/random/_bounded_integers + legacy use_masked=True).  rng==0 consumes
 * no draw.  (move_funcs.pyx:819 calls it with low=0, high=sites_per_bead+1.) */
int64_t oc_mt_randint(oc_mt19937 *s, int64_t low, int64_t high)
{
    uint64_t rng = (uint64_t)(high - 1 - low);
    uint64_t mask = rng;
    uint32_t val;
    if (rng == 0) return low;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    if (rng > 0xFFFFFFFFull) { /* not reachable on the hot path */
        uint64_t v;
        do {
            uint64_t hi = oc_mt_next(s), lo = oc_mt_next(s);
            v = ((hi << 32) | lo) & mask;
        } while (v > rng);
        return low + (int64_t)v;
    }
    do {
        val = oc_mt_next(s) & (uint32_t)mask;
    } while (val > (uint32_t)rng);
    return low + (int64_t)val;
}

/* --------------------------------------------------------------- helpers */

/* Python-semantics float modulo: the generated helper of `cdivision=False`
 * (setup.py:89): fmod, then add b when the signs differ.  SURVEY quirk 1. */
static double py_mod(double a, double b)
{
    double r = fmod(a, b);
    r += ((r != 0) & ((r < 0) ^ (b < 0))) * b;
    return r;
}

static double dot3(const double *a, const double *b)
{
    /* linalg.pyx:374-395 */
    double d = 0;
    int i;
    for (i = 0; i < 3; i++) d += a[i] * b[i];
    return d;
}

static int64_t super_index(const oc_sim *s, int64_t ix, int64_t iy, int64_t iz)
{
    /* inds_xyz_to_super[i,j,k] = (i%nx) + nx*(j%ny) + nx*ny*(k%nz)
     * fields.pyx:673-685, 2379-2402 */
    return (ix % s->nx) + s->nx * (iy % s->ny) + s->nx * s->ny * (iz % s->nz);
}

/* ----------------------------------------------------------- A1: binning */

/* fields.pyx:1437-1462 (per-axis wrap/index/weight), 1524-1598 (8 weights),
 * 1600-1673 (8 super-indices).  Quirk 2: the weight uses the UNPATCHED ind. */
/* fast_field = 1: init_fast_field fields.pyx:577-671 precomputes, for every sub-bin i of an axis (n_points per
 * voxel edge), the lower voxel index and its weight from the sub-bin's LOWER EDGE
 *   x_i = (i - n_points/2) * sub_bin_width,  ind = floor(x_i / d),  lower = (ind == -1 ? n-1 : ind),
 *   lower_weight = 1 - (x_i / d - ind);
 * get_change_in_density_quickly 1235-1368 looks a position up by
 *   i = floor((x + W/2) / sub_bin_width) % (n_points * n)   (numpy floor, Python modulo),
 * and adds the terms without the 1e-18 filters of the exact path.  Same quantities, computed on the fly. */
static void bin_point_fast(const oc_sim *s, const double xyz[3], int64_t idx[8], double w[8])
{
    int64_t n3[3] = {s->nx, s->ny, s->nz};
    int64_t ind3[3];
    double w3[3];
    int j, l;
    for (j = 0; j < 3; j++) {
        double sbw = s->dxyz[j] / (double)s->fast_n_points;
        double nsub = (double)(s->fast_n_points * n3[j]);
        double sub = py_mod(floor((xyz[j] + s->half_width[j]) / sbw), nsub);
        double xi = ((double)((int64_t)sub - s->fast_n_points / 2) * sbw) / s->dxyz[j];
        int64_t ind = (int64_t)floor(xi);
        ind3[j] = (ind == -1) ? n3[j] - 1 : ind;
        w3[j] = 1 - (xi - (double)ind);
    }
    for (l = 0; l < 8; l++) {
        int bx = l & 1, by = (l >> 1) & 1, bz = (l >> 2) & 1;
        double wx = bx ? (1 - w3[0]) : w3[0];
        double wy = by ? (1 - w3[1]) : w3[1];
        double wz = bz ? (1 - w3[2]) : w3[2];
        w[l] = wx * wy * wz;
        idx[l] = super_index(s, ind3[0] + bx, ind3[1] + by, ind3[2] + bz);
    }
}

void oc_bin_point(const oc_sim *s, const double xyz[3], int64_t idx[8], double w[8])
{
    int64_t n_m1[3] = {s->nx - 1, s->ny - 1, s->nz - 1};
    int64_t ind3[3];
    double w3[3];
    int j, l;
    for (j = 0; j < 3; j++) {
        double x = py_mod(xyz[j] + s->half_width[j], s->width[j]) - s->half_step[j];
        int64_t ind = (int64_t)floor(x / s->dxyz[j]);
        ind3[j] = (ind == -1) ? n_m1[j] : ind;
        w3[j] = 1 - (x / s->dxyz[j] - ind);
    }
    for (l = 0; l < 8; l++) {
        int bx = l & 1, by = (l >> 1) & 1, bz = (l >> 2) & 1;
        double wx = bx ? (1 - w3[0]) : w3[0];
        double wy = by ? (1 - w3[1]) : w3[1];
        double wz = bz ? (1 - w3[2]) : w3[2];
        w[l] = wx * wy * wz;
        idx[l] = super_index(s, ind3[0] + bx, ind3[1] + by, ind3[2] + bz);
    }
}

/* ---------------------------------------------------- A8: full recompute */

/* update_all_densities (fields.pyx:1977-2039) when for_all_polymers == 0;
 * update_all_densities_for_all_polymers (2041-2106) otherwise: that variant
 * zeroes only columns 0..nb-1 (quirk 7) and clamps |rho| < 1e-18 to 0. */
void oc_update_all_densities(oc_sim *s, int for_all_polymers)
{
    int64_t ncol = s->nb + 1, i, j, l, m;
    int64_t zero_cols = for_all_polymers ? s->nb : ncol;
    for (i = 0; i < s->n_bins; i++)
        for (j = 0; j < zero_cols; j++) {
            s->density[i * ncol + j] = 0;
            s->density_trial[i * ncol + j] = 0;
        }
    for (i = 0; i < s->N; i++) {
        int64_t idx[8];
        double w[8];
        oc_bin_point(s, &s->r[3 * i], idx, w);
        for (l = 0; l < 8; l++) {
            double density = w[l] / s->access_vol[idx[l]];
            s->density[idx[l] * ncol] += density;
            for (m = 1; m < ncol; m++)
                s->density[idx[l] * ncol + m] += density * (double)s->states[i * s->nb + m - 1];
        }
    }
    if (for_all_polymers)
        for (i = 0; i < s->n_bins * ncol; i++)
            if (fabs(s->density[i]) < 1E-18) s->density[i] = 0;
}

/* Python's round(x, 2) on a float (fields.pyx:2281): correctly-rounded
 * decimal rounding, ties to even on the exact binary value. */
static double py_round2(double x)
{
    char buf[400];
    if (!isfinite(x)) return x;
    snprintf(buf, sizeof buf, "%.2f", x); /* glibc: exact, round-half-even */
    return strtod(buf, NULL);
}

/* compute_E -> get_E_binders_and_beads + nonspecific_interact_E
 * (fields.pyx:1939-1966, 2208-2315).  Quirk 6: phi*(1-phi), round(phi,2), no
 * cross-talk.  Densities are recomputed first, as compute_E does. */
double oc_field_E(oc_sim *s)
{
    int64_t ncol = s->nb + 1, i, j;
    double E = 0, nonspecific = 0;
    oc_update_all_densities(s, 0);
    for (i = 0; i < s->nb; i++) {
        double tot = 0;
        int64_t doubly = 0;
        for (j = 0; j < s->n_bins; j++) {
            double d = s->density[j * ncol + i + 1];
            tot += d * d;
        }
        for (j = 0; j < s->N; j++)
            if (s->states[j * s->nb + i] == 2) doubly++;
        E += s->field_pref[i] * tot;
        E += s->e_intra[i] * (double)doubly;
    }
    for (i = 0; i < s->n_bins; i++) {
        double vf = s->density[i * ncol] * s->bead_vol;
        if (py_round2(vf) > s->vf_limit)
            nonspecific += E_HUGE_FIELD * vf;
        else
            nonspecific += s->chi * (s->access_vol[i] / s->bead_vol) * vf * (1 - vf);
    }
    E += nonspecific;
    return E;
}

/* ------------------------------------------------ A9: elastic energies */

static double E_pair(const oc_sim *s, const double *bend, double dr_par,
                     const double *dr_perp, int64_t b)
{
    /* polymers.pyx:1148-1175 */
    double t = dr_par - s->gamma[b];
    return 0.5 * s->eps_bend[b] * dot3(bend, bend) + 0.5 * s->eps_par[b] * (t * t) +
           0.5 * s->eps_perp[b] * dot3(dr_perp, dr_perp);
}

/* compute_twist_angle_omega polymers.pyx:3427-3461 */
static double twist_omega(const double *t2_0, const double *t3_0, const double *t2_1, const double *t3_1)
{
    double t1_0[3], t1_1[3];
    t1_0[0] = t2_0[1] * t3_0[2] - t2_0[2] * t3_0[1];
    t1_0[1] = t2_0[2] * t3_0[0] - t2_0[0] * t3_0[2];
    t1_0[2] = t2_0[0] * t3_0[1] - t2_0[1] * t3_0[0];
    t1_1[0] = t2_1[1] * t3_1[2] - t2_1[2] * t3_1[1];
    t1_1[1] = t2_1[2] * t3_1[0] - t2_1[0] * t3_1[2];
    t1_1[2] = t2_1[0] * t3_1[1] - t2_1[1] * t3_1[0];
    return atan2(dot3(t2_0, t1_1) - dot3(t1_0, t2_1), dot3(t1_0, t1_1) + dot3(t2_0, t2_1));
}

/* the twist term of E_pair_with_twist polymers.pyx:2050-2102 (0 for chains without twist) */
static double E_twist(const oc_sim *s, double omega, int64_t b)
{
    const double two_pi = 2 * M_PI;
    double d = omega - s->twist0[b];
    d -= two_pi * floor((d + M_PI) / two_pi);
    return 0.5 * s->eps_twist[b] * (d * d);
}

/* SSWLC.compute_E polymers.pyx:1348-1381; SSTWLC.compute_E 2250-2285 */
double oc_poly_E(const oc_sim *s)
{
    double E = 0;
    int64_t i;
    int j;
    for (i = 1; i < s->N; i++) {
        const double *r0 = &s->r[3 * (i - 1)], *r1 = &s->r[3 * i];
        const double *t0 = &s->t3[3 * (i - 1)], *t1 = &s->t3[3 * i];
        double dr[3], dr_perp[3], bend[3], dr_par;
        for (j = 0; j < 3; j++) dr[j] = r1[j] - r0[j];
        dr_par = dot3(t0, dr);
        for (j = 0; j < 3; j++) dr_perp[j] = dr[j] - t0[j] * dr_par;
        for (j = 0; j < 3; j++) bend[j] = t1[j] + (-t0[j] - s->eta[i - 1] * dr_perp[j]);
        if (s->eps_twist)
            E += E_pair(s, bend, dr_par, dr_perp, i - 1) +
                 E_twist(s, twist_omega(&s->t2[3 * (i - 1)], t0, &s->t2[3 * i], t1), i - 1);
        else
            E += E_pair(s, bend, dr_par, dr_perp, i - 1);
    }
    return E;
}

/* bead_pair_dE_poly_forward polymers.pyx:1177-1277 */
static double pair_dE_forward(const oc_sim *s, const double *r_0, const double *r_1,
                              const double *test_r_1, const double *t3_0,
                              const double *t3_1, const double *test_t3_1, const double *t2_0,
                              const double *t2_1, const double *test_t2_1, int64_t b)
{
    double dr[3], dr_test[3], dr_perp[3], dr_perp_test[3], bend[3], bend_test[3];
    double dr_par, dr_par_test;
    int i;
    for (i = 0; i < 3; i++) {
        dr_test[i] = test_r_1[i] - r_0[i];
        dr[i] = r_1[i] - r_0[i];
    }
    dr_par_test = dot3(t3_0, dr_test);
    dr_par = dot3(t3_0, dr);
    for (i = 0; i < 3; i++) {
        dr_perp_test[i] = dr_test[i] - t3_0[i] * dr_par_test;
        dr_perp[i] = dr[i] - t3_0[i] * dr_par;
        bend_test[i] = test_t3_1[i] - t3_0[i] - dr_perp_test[i] * s->eta[b];
        bend[i] = t3_1[i] - t3_0[i] - dr_perp[i] * s->eta[b];
    }
    if (s->eps_twist) /* bead_pair_dE_poly_forward_with_twist polymers.pyx:2104-2175 */
        return (E_pair(s, bend_test, dr_par_test, dr_perp_test, b) +
                E_twist(s, twist_omega(t2_0, t3_0, test_t2_1, test_t3_1), b)) -
               (E_pair(s, bend, dr_par, dr_perp, b) + E_twist(s, twist_omega(t2_0, t3_0, t2_1, t3_1), b));
    return E_pair(s, bend_test, dr_par_test, dr_perp_test, b) -
           E_pair(s, bend, dr_par, dr_perp, b);
}

/* bead_pair_dE_poly_reverse polymers.pyx:1279-1346 */
static double pair_dE_reverse(const oc_sim *s, const double *r_0, const double *test_r_0,
                              const double *r_1, const double *t3_0,
                              const double *test_t3_0, const double *t3_1, const double *t2_0,
                              const double *test_t2_0, const double *t2_1, int64_t b)
{
    double dr[3], dr_test[3], dr_perp[3], dr_perp_test[3], bend[3], bend_test[3];
    double dr_par, dr_par_test;
    int i;
    for (i = 0; i < 3; i++) {
        dr_test[i] = r_1[i] - test_r_0[i];
        dr[i] = r_1[i] - r_0[i];
    }
    dr_par_test = dot3(test_t3_0, dr_test);
    dr_par = dot3(t3_0, dr);
    for (i = 0; i < 3; i++) {
        dr_perp_test[i] = dr_test[i] - test_t3_0[i] * dr_par_test;
        dr_perp[i] = dr[i] - t3_0[i] * dr_par;
        bend_test[i] = t3_1[i] - test_t3_0[i] - dr_perp_test[i] * s->eta[b];
        bend[i] = t3_1[i] - t3_0[i] - dr_perp[i] * s->eta[b];
    }
    if (s->eps_twist) /* bead_pair_dE_poly_reverse_with_twist polymers.pyx:2177-2248 */
        return (E_pair(s, bend_test, dr_par_test, dr_perp_test, b) +
                E_twist(s, twist_omega(test_t2_0, test_t3_0, t2_1, t3_1), b)) -
               (E_pair(s, bend, dr_par, dr_perp, b) + E_twist(s, twist_omega(t2_0, t3_0, t2_1, t3_1), b));
    return E_pair(s, bend_test, dr_par_test, dr_perp_test, b) -
           E_pair(s, bend, dr_par, dr_perp, b);
}

/* ---- DetailedChromatin: entry / exit frames of a nucleosome (DetailedNucleosome.update_configuration
 * beads.py:536-574) ---- */
static void mat3_mul(const double a[9], const double b[9], double o[9])
{
    int i, j, k;
    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) {
            double t = 0;
            for (k = 0; k < 3; k++) t += a[3 * i + k] * b[3 * k + j];
            o[3 * i + j] = t;
        }
}
static void mat3_vec(const double a[9], const double v[3], double o[3])
{
    int i;
    for (i = 0; i < 3; i++) o[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
/* rotation_matrix_from_vectors linalg.pyx:478-510 */
static void rot_from_vectors(const double v1[3], const double v2[3], double R[9])
{
    double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    double a[3], b[3], v[3], K[9], K2[9];
    int i;
    for (i = 0; i < 3; i++) a[i] = v1[i] / n1, b[i] = v2[i] / n2;
    cross3(a, b, v);
    for (i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (v[0] != 0 || v[1] != 0 || v[2] != 0) {
        double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
        double sn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        double f = (1 - c) / (sn * sn);
        K[0] = 0, K[1] = -v[2], K[2] = v[1];
        K[3] = v[2], K[4] = 0, K[5] = -v[0];
        K[6] = -v[1], K[7] = v[0], K[8] = 0;
        mat3_mul(K, K, K2);
        for (i = 0; i < 9; i++) R[i] = R[i] + K[i] + K2[i] * f;
    }
}
static int allclose3(const double x[3], const double y[3])
{ /* np.allclose defaults: |x - y| <= 1e-8 + 1e-5 |y| */
    int i;
    for (i = 0; i < 3; i++)
        if (!(fabs(x[i] - y[i]) <= 1e-8 + 1e-5 * fabs(y[i]))) return 0;
    return 1;
}
/* get_rotation_matrix linalg.pyx:537-575 */
static void local_to_global(const oc_detailed *d, const double t3[3], const double t2[3], double R[9])
{
    double R1[9], R2[9], t2r[3], chk[3], neg[3];
    int i;
    rot_from_vectors(d->t3_local, t3, R1);
    mat3_vec(R1, d->t2_local, t2r);
    rot_from_vectors(t2r, t2, R2);
    mat3_mul(R2, R1, R);
    mat3_vec(R, d->t3_local, chk);
    for (i = 0; i < 3; i++) neg[i] = -t3[i];
    if (allclose3(chk, neg)) { /* get_arbitrary_axis_rotation_matrix(t2, pi) linalg.pyx:512-535 */
        double n = sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
        double ux = t2[0] / n, uy = t2[1] / n, uz = t2[2] / n, ct = cos(M_PI), st = sin(M_PI);
        double R3[9], Rn[9];
        R3[0] = ct + ux * ux * (1 - ct), R3[1] = ux * uy * (1 - ct) - uz * st, R3[2] = uz * ux * (1 - ct) + uy * st;
        R3[3] = ux * uy * (1 - ct) + uz * st, R3[4] = ct + uy * uy * (1 - ct), R3[5] = uy * uz * (1 - ct) - ux * st;
        R3[6] = uz * ux * (1 - ct) - uy * st, R3[7] = uy * uz * (1 - ct) + ux * st, R3[8] = ct + uz * uz * (1 - ct);
        mat3_mul(R3, R, Rn);
        memcpy(R, Rn, sizeof Rn);
    }
}
void oc_nucleosome_frames(const oc_detailed *d, const double r[3], const double t3[3], const double t2[3],
                          double r_enter[3], double r_exit[3], double t3_exit[3], double t2_exit[3])
{
    double R[9], t1[3], v[3], t1_exit[3];
    int i;
    cross3(t2, t3, t1); /* beads.py:566 */
    local_to_global(d, t3, t2, R);
    mat3_vec(R, d->r_enter_unit, v);
    for (i = 0; i < 3; i++) r_enter[i] = v[i] * d->r_enter_norm + r[i];
    mat3_vec(R, d->r_exit_unit, v);
    for (i = 0; i < 3; i++) r_exit[i] = v[i] * d->r_exit_norm + r[i];
    for (i = 0; i < 3; i++) { /* nucleo_geom.get_T3 / get_T1 / get_T2 */
        t3_exit[i] = d->a3[0] * t3[i] + d->a3[1] * t2[i] + d->a3[2] * t1[i];
        t1_exit[i] = d->a1[0] * t3[i] + d->a1[1] * t2[i] + d->a1[2] * t1[i];
    }
    cross3(t3_exit, t1_exit, t2_exit);
}

/* DetailedChromatin.continuous_dE_poly polymers.pyx:2503-2607: the linker DNA runs from the EXIT point / frame
 * of one nucleosome to the ENTRY point / frame (= the nucleosome's own t3, t2) of the next */
static double continuous_dE_poly_detailed(const oc_sim *s, int64_t ind0, int64_t indf)
{
    const oc_detailed *d = s->detailed;
    double dE = 0;
    double ri0[3], ro0[3], t3o0[3], t2o0[3], ri1[3], ro1[3], t3o1[3], t2o1[3], rit[3], rot[3], t3ot[3], t2ot[3];
    if (ind0 != 0) {
        int64_t a = ind0 - 1, b = ind0;
        oc_nucleosome_frames(d, &s->r[3 * a], &s->t3[3 * a], &s->t2[3 * a], ri0, ro0, t3o0, t2o0);
        oc_nucleosome_frames(d, &s->r_trial[3 * b], &s->t3_trial[3 * b], &s->t2_trial[3 * b], rit, rot, t3ot, t2ot);
        oc_nucleosome_frames(d, &s->r[3 * b], &s->t3[3 * b], &s->t2[3 * b], ri1, ro1, t3o1, t2o1);
        dE += pair_dE_forward(s, ro0, ri1, rit, t3o0, &s->t3[3 * b], &s->t3_trial[3 * b], t2o0, &s->t2[3 * b],
                              &s->t2_trial[3 * b], a);
    }
    if (indf != s->N) {
        int64_t a = indf - 1, b = indf;
        oc_nucleosome_frames(d, &s->r_trial[3 * a], &s->t3_trial[3 * a], &s->t2_trial[3 * a], rit, rot, t3ot, t2ot);
        oc_nucleosome_frames(d, &s->r[3 * a], &s->t3[3 * a], &s->t2[3 * a], ri0, ro0, t3o0, t2o0);
        oc_nucleosome_frames(d, &s->r[3 * b], &s->t3[3 * b], &s->t2[3 * b], ri1, ro1, t3o1, t2o1);
        dE += pair_dE_reverse(s, ro0, rot, ri1, t3o0, t3ot, &s->t3[3 * b], t2o0, t2ot, &s->t2[3 * b], a);
    }
    return dE;
}

/* continuous_dE_poly polymers.pyx:1084-1146 */
static double continuous_dE_poly(const oc_sim *s, int64_t ind0, int64_t indf)
{
    double dE = 0;
    if (s->detailed) return continuous_dE_poly_detailed(s, ind0, indf);
    if (ind0 != 0)
        dE += pair_dE_forward(s, &s->r[3 * (ind0 - 1)], &s->r[3 * ind0], &s->r_trial[3 * ind0],
                              &s->t3[3 * (ind0 - 1)], &s->t3[3 * ind0],
                              &s->t3_trial[3 * ind0], &s->t2[3 * (ind0 - 1)], &s->t2[3 * ind0],
                              &s->t2_trial[3 * ind0], ind0 - 1);
    if (indf != s->N)
        dE += pair_dE_reverse(s, &s->r[3 * (indf - 1)], &s->r_trial[3 * (indf - 1)],
                              &s->r[3 * indf], &s->t3[3 * (indf - 1)],
                              &s->t3_trial[3 * (indf - 1)], &s->t3[3 * indf], &s->t2[3 * (indf - 1)],
                              &s->t2_trial[3 * (indf - 1)], &s->t2[3 * indf], indf - 1);
    return dE;
}

/* ---------------------------------------------------- A10: binding dE */

static double comb_small(int64_t n, int64_t k)
{
    /* scipy.special.comb(N, k) (exact=False): 0 outside 0<=k<=N */
    double c = 1;
    int64_t i;
    if (k < 0 || n < 0 || k > n) return 0;
    for (i = 1; i <= k; i++) c = c * (double)(n - k + i) / (double)i;
    return floor(c + 0.5);
}

/* single-site Helmholtz free energy, polymers.pyx:1493-1517 */
double oc_binding_free_energy(int64_t Nn, int64_t Nm, int64_t st, double e_mod, double e_nomod)
{
    double sum = 0;
    int64_t i;
    for (i = 0; i <= st; i++)
        sum += comb_small(Nm, i) * comb_small(Nn - Nm, st - i) *
               exp(-((double)i * e_mod + (double)(st - i) * e_nomod));
    return -log(sum);
}

/* bead_binding_dE polymers.pyx:1408-1538 */
static double bead_binding_dE(const oc_sim *s, int64_t ind)
{
    const int64_t *st_t = &s->states_trial[ind * s->nb];
    const int64_t *st_c = &s->states[ind * s->nb];
    const int64_t *mod = &s->mods[ind * s->nb];
    double dE = 0;
    int64_t b;
    if (s->max_binders != -1) {
        int64_t tot = 0;
        for (b = 0; b < s->nb; b++) tot += st_t[b];
        if (tot > s->max_binders) dE += E_HUGE_POLY * (double)(tot - s->max_binders);
        tot = 0;
        for (b = 0; b < s->nb; b++) tot += st_c[b];
        if (tot > s->max_binders) dE -= E_HUGE_POLY * (double)(tot - s->max_binders);
    }
    for (b = 0; b < s->nb; b++) {
        int64_t Nn = s->sites_per_bead[b], Nm = mod[b];
        double mu = s->chemical_potential[b];
        dE += oc_binding_free_energy(Nn, Nm, st_t[b], s->bind_energy_mod[b],
                                     s->bind_energy_no_mod[b]);
        dE -= oc_binding_free_energy(Nn, Nm, st_c[b], s->bind_energy_mod[b],
                                     s->bind_energy_no_mod[b]);
        if (mu > 0) {
            dE -= (double)st_t[b] * (mu * (-s->mu_adjust_factor + 2));
            dE += (double)st_c[b] * (mu * (-s->mu_adjust_factor + 2));
        } else {
            dE -= (double)st_t[b] * mu * s->mu_adjust_factor;
            dE += (double)st_c[b] * mu * s->mu_adjust_factor;
        }
    }
    return dE;
}

/* SSWLC.compute_dE polymers.pyx:1024-1082 */
double oc_poly_dE(oc_sim *s, int move, const int64_t *inds, int64_t n)
{
    double dE = 0;
    int64_t i;
    if (move == OC_BINDING) {
        int64_t ind0 = inds[0];
        for (i = 0; i < n; i++) dE += bead_binding_dE(s, ind0 + i); /* binding_dE 1383-1406 */
    } else if (move == OC_SLIDE || move == OC_PIVOT || move == OC_CRANK) {
        dE += continuous_dE_poly(s, inds[0], inds[n - 1] + 1);
    } else if (move == OC_TANGENT) {
        for (i = 0; i < n; i++) dE += continuous_dE_poly(s, inds[i], inds[i] + 1);
    }
    return dE;
}

/* ---------------------------------------------- A6: confinement energy */

/* FieldBase.get_confinement_dE fields.pyx:121-202 (quirk 5: the cubical
 * branch never counts the current configuration) */
static double confinement_E(const oc_sim *s, const int64_t *inds, int64_t n, int trial)
{
    int64_t out = 0, i;
    int j;
    if (s->confine_type == OC_CONFINE_NONE) return 0.;
    if (s->confine_type == OC_CONFINE_SPHERICAL) {
        const double *r = trial ? s->r_trial : s->r;
        for (i = 0; i < n; i++) {
            double dist = sqrt(dot3(&r[3 * inds[i]], &r[3 * inds[i]]));
            if (dist > s->confine_length) out++;
        }
        return (double)out * E_HUGE_FIELD;
    }
    if (trial == 1)
        for (i = 0; i < n; i++)
            for (j = 0; j < 3; j++)
                if (fabs(s->r_trial[3 * inds[i] + j]) > s->confine_length / 2) out++;
    return (double)out * E_HUGE_FIELD;
}

/* ------------------------------------- A1-A5: field dE of a proposed move */

/* get_change_in_density fields.pyx:1370-1522.  Fills density_trial rows
 * (assigned on a bin's first touch, += afterwards; quirk 3) and the touched
 * list (the reference's `bins_found` set, here in first-touch order). */
static void change_in_density(oc_sim *s, const int64_t *inds, int64_t n, int state_change)
{
    int64_t ncol = s->nb + 1, i, m;
    int k, l;
    s->n_touched = 0;
    s->stamp++;
    for (i = 0; i < n; i++) {
        int64_t idx[2][8];
        double w[2][8];
        int64_t bead = inds[i];
        const int fast = s->fast_n_points > 0;
        if (fast) bin_point_fast(s, &s->r[3 * bead], idx[0], w[0]);
        else oc_bin_point(s, &s->r[3 * bead], idx[0], w[0]);
        if (state_change == 0) {
            if (fast) bin_point_fast(s, &s->r_trial[3 * bead], idx[1], w[1]);
            else oc_bin_point(s, &s->r_trial[3 * bead], idx[1], w[1]);
        } else { /* quirk 4: trial coords = current coords */
            memcpy(idx[1], idx[0], sizeof idx[0]);
            memcpy(w[1], w[0], sizeof w[0]);
        }
        for (k = 0; k < 2; k++) {
            double prefactor = (k == 0) ? -1. : 1.;
            for (l = 0; l < 8; l++) {
                int64_t bin = idx[k][l];
                double base = w[k][l] / s->access_vol[bin];
                int first = (s->touch_stamp[bin] != s->stamp);
                if (first) {
                    s->touch_stamp[bin] = s->stamp;
                    s->touched[s->n_touched++] = bin;
                }
                for (m = 0; m < ncol; m++) {
                    double dens = base, temp;
                    if (m > 0) {
                        const int64_t *st =
                            (k == 0 || state_change == 0) ? s->states : s->states_trial;
                        dens = base * (double)st[bead * s->nb + m - 1];
                    }
                    temp = prefactor * dens;
                    if (fast) { /* fields.pyx:1350-1366: no threshold in the fast path */
                        if (first) s->density_trial[bin * ncol + m] = temp;
                        else s->density_trial[bin * ncol + m] += temp;
                    } else if (first)
                        s->density_trial[bin * ncol + m] = (fabs(temp) > 1E-18) ? temp : 0;
                    else if (fabs(temp) > 1E-18)
                        s->density_trial[bin * ncol + m] += temp;
                }
            }
        }
    }
}

/* nonspecific_interact_dE + get_volume_fractions_with_trial fields.pyx:1792-1875 */
static double nonspecific_dE(const oc_sim *s)
{
    int64_t ncol = s->nb + 1, i;
    double dE = 0;
    for (i = 0; i < s->n_touched; i++) {
        int64_t bin = s->touched[i];
        double access = s->access_vol[bin];
        double vf0 = s->density[bin * ncol] * s->bead_vol;
        double vf1 = vf0 + (s->density_trial[bin * ncol] * s->bead_vol);
        if (vf1 > s->vf_limit)
            dE += E_HUGE_FIELD * vf1;
        else
            dE += s->chi * (access / s->bead_vol) * (vf1 * vf1);
        if (vf0 > s->vf_limit)
            dE -= E_HUGE_FIELD * vf0;
        else
            dE -= s->chi * (access / s->bead_vol) * (vf0 * vf0);
    }
    return dE;
}

/* get_dE_binders_and_beads fields.pyx:1675-1790 (+ count_doubly_bound 1877-1937) */
static double dE_binders_and_beads(const oc_sim *s, const int64_t *inds, int64_t n,
                                   int state_change)
{
    int64_t ncol = s->nb + 1, nb = s->nb, a, b, k, i;
    double dE = 0;
    for (a = 0; a < nb; a++) {
        double tot = 0;
        int64_t d_cur = 0, d_trial = 0;
        for (k = 0; k < s->n_touched; k++) {
            int64_t bin = s->touched[k];
            double rho = s->density[bin * ncol + a + 1];
            double rn = rho + s->density_trial[bin * ncol + a + 1];
            double t = rn * rn - rho * rho;
            if (fabs(t) < 1E-18) t = 0;
            tot += t;
        }
        dE += s->field_pref[a] * tot;
        for (i = 0; i < n; i++) {
            if (s->states[inds[i] * nb + a] == 2) d_cur++;
            if (state_change) {
                if (s->states_trial[inds[i] * nb + a] == 2) d_trial++;
            } else if (s->states[inds[i] * nb + a] == 2) {
                d_trial++;
            }
        }
        dE += s->e_intra[a] * (double)(d_trial - d_cur);
    }
    for (a = 0; a < nb; a++)
        for (b = 0; b < nb; b++) {
            double tot = 0;
            for (k = 0; k < s->n_touched; k++) {
                int64_t bin = s->touched[k];
                double ra = s->density[bin * ncol + a + 1], rb = s->density[bin * ncol + b + 1];
                double t = ((ra + s->density_trial[bin * ncol + a + 1]) *
                            (rb + s->density_trial[bin * ncol + b + 1])) -
                           (ra * rb);
                if (fabs(t) < 1E-18) t = 0;
                tot += t;
            }
            dE += s->xpref[a * nb + b] * tot;
        }
    dE += nonspecific_dE(s);
    return dE;
}

/* UniformDensityField.compute_dE fields.pyx:1149-1233 */
double oc_field_dE(oc_sim *s, const int64_t *inds, int64_t n, int state_change)
{
    double dE = 0;
    int64_t i;
    if (state_change == 0) {
        dE += confinement_E(s, inds, n, 1);
        dE -= confinement_E(s, inds, n, 0);
    }
    change_in_density(s, inds, n, state_change);
    for (i = 0; i < s->n_bins; i++) s->affected[i] = 0;
    for (i = 0; i < s->n_touched; i++) s->affected[s->touched[i]] = 1;
    dE += dE_binders_and_beads(s, inds, n, state_change);
    return dE;
}

/* A7: update_affected_densities fields.pyx:1968-1975 */
void oc_update_affected_densities(oc_sim *s)
{
    int64_t ncol = s->nb + 1, i, j;
    for (i = 0; i < s->n_bins; i++)
        if (s->affected[i] == 1)
            for (j = 0; j < ncol; j++) {
                s->density[i * ncol + j] += s->density_trial[i * ncol + j];
                s->density_trial[i * ncol + j] = 0;
            }
}

/* ------------------------------------------------------- A11: proposals */

/* capped_exponential bead_selection.pyx:19-67 */
static int64_t capped_exponential(oc_glibc_rand *g, int64_t window, int64_t cap)
{
    int64_t r = (int64_t)(-log10(oc_uniform(g) + 0.00001) * (double)window * 0.45 + 1.0001);
    while (r > cap)
        r = (int64_t)(-log10(oc_uniform(g) + 0.00001) * (double)window * 0.45 + 1.0001);
    return r;
}

int64_t oc_from_left(oc_glibc_rand *g, int64_t window, int64_t N)
{
    (void)N; /* bead_selection.pyx:69-90 (window > N raises in the reference) */
    return capped_exponential(g, window, window);
}

int64_t oc_from_right(oc_glibc_rand *g, int64_t window, int64_t N)
{
    return N - oc_from_left(g, window, N); /* bead_selection.pyx:93-112 */
}

static int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

/* bead_selection.pyx:115-154 */
int64_t oc_from_point(oc_glibc_rand *g, int64_t window, int64_t N, int64_t ind0)
{
    int64_t side, window_side, upper;
    if (window < 1) return ind0;
    side = oc_rand(g) % 2;
    if (side == 0) {
        window_side = imax(imin(window, ind0), 1);
        upper = imax(ind0, 1);
        return oc_from_right(g, window_side, upper);
    }
    window_side = imax(imin(window, N - ind0), 1);
    upper = imax(N - ind0, 1);
    return oc_from_left(g, window_side, upper) + ind0;
}

/* bead_selection.pyx:157-192 */
static void check_bead_bounds(int64_t b0, int64_t b1, int64_t N, int64_t *ind0, int64_t *indf)
{
    b0 = imin(b0, N);
    b0 = imax(b0, 0);
    if (b1 > N) {
        *ind0 = b0;
        *indf = N;
    } else if (b1 < 0) {
        *ind0 = 0;
        *indf = b0 + 1;
    } else if (b0 == b1) {
        *ind0 = b0;
        *indf = b0 + 1;
    } else {
        *ind0 = imin(b0, b1);
        *indf = imax(b0, b1);
    }
}

/* uniform_sample_unit_sphere(_inplace) linalg.pyx:23-59 */
static void sample_sphere_exact(oc_glibc_rand *g, double v[3])
{
    double phi = oc_uniform(g) * (2.0 * M_PI);
    double theta = acos(oc_uniform(g) * 2 - 1);
    v[0] = cos(phi) * sin(theta);
    v[1] = sin(phi) * sin(theta);
    v[2] = cos(theta);
}

/* the same point; with the production streams in the product's closed form
 * (geometry.cuh unit_sphere_point<false>: cos(theta) = x = 2 u2 - 1,
 * sin(theta) = sqrt((1 - x)(1 + x)), sincospi(2 u1)) */
static void sample_sphere(oc_glibc_rand *g, double v[3])
{
    double phi, theta;
    if (g->alt) {
        double u1 = oc_uniform(g), u2 = oc_uniform(g);
        double x = u2 * 2.0 - 1.0;
        double st = sqrt((1.0 - x) * (1.0 + x));
        double a = M_PI * (u1 * 2.0);
        v[0] = cos(a) * st;
        v[1] = sin(a) * st;
        v[2] = x;
        return;
    }
    phi = oc_uniform(g) * (2.0 * M_PI);
    theta = acos(oc_uniform(g) * 2 - 1);
    v[0] = cos(phi) * sin(theta);
    v[1] = sin(phi) * sin(theta);
    v[2] = cos(theta);
}

/* arbitrary_axis_rotation linalg.pyx:62-139; m is the 4x4 row-major matrix */
void oc_rotation_matrix(const double axis[3], const double point[3], double ang, double m[16])
{
    double c = cos(ang), sn = sin(ang);
    double rot[3];
    int i;
    for (i = 0; i < 16; i++) m[i] = 0;
    for (i = 0; i < 4; i++) m[5 * i] = 1;
    m[0] = axis[0] * axis[0] + (axis[1] * axis[1] + axis[2] * axis[2]) * c;
    m[1] = axis[0] * axis[1] * (1 - c) - axis[2] * sn;
    m[2] = axis[0] * axis[2] * (1 - c) + axis[1] * sn;
    m[4] = axis[0] * axis[1] * (1 - c) + axis[2] * sn;
    m[5] = axis[1] * axis[1] + (axis[0] * axis[0] + axis[2] * axis[2]) * c;
    m[6] = axis[1] * axis[2] * (1 - c) - axis[0] * sn;
    m[8] = axis[0] * axis[2] * (1 - c) - axis[1] * sn;
    m[9] = axis[1] * axis[2] * (1 - c) + axis[0] * sn;
    m[10] = axis[2] * axis[2] + (axis[0] * axis[0] + axis[1] * axis[1]) * c;
    rot[0] = (point[1] * axis[2] - point[2] * axis[1]) * sn;
    rot[1] = (point[2] * axis[0] - point[0] * axis[2]) * sn;
    rot[2] = (point[0] * axis[1] - point[1] * axis[0]) * sn;
    rot[0] += (point[0] * (1 - axis[0] * axis[0]) -
               axis[0] * (point[1] * axis[1] + point[2] * axis[2])) * (1 - c);
    rot[1] += (point[1] * (1 - axis[1] * axis[1]) -
               axis[1] * (point[0] * axis[0] + point[2] * axis[2])) * (1 - c);
    rot[2] += (point[2] * (1 - axis[2] * axis[2]) -
               axis[2] * (point[0] * axis[0] + point[1] * axis[1])) * (1 - c);
    for (i = 0; i < 3; i++) m[4 * i + 3] = rot[i];
}

/* transform_r_t3_t2 move_funcs.pyx:121-154 */
void oc_transform_rows(oc_sim *s, const double m[16], const int64_t *inds, int64_t n)
{
    int64_t i;
    int j, k;
    for (i = 0; i < n; i++) {
        int64_t b = inds[i];
        for (j = 0; j < 3; j++) {
            double er = 0, e3 = 0, e2 = 0;
            for (k = 0; k < 3; k++) {
                er += m[4 * j + k] * s->r[3 * b + k];
                e3 += m[4 * j + k] * s->t3[3 * b + k];
                e2 += m[4 * j + k] * s->t2[3 * b + k];
            }
            s->r_trial[3 * b + j] = er + m[4 * j + 3];
            s->t3_trial[3 * b + j] = e3;
            s->t2_trial[3 * b + j] = e2;
        }
    }
}

/* get_crank_shaft_axis move_funcs.pyx:157-234 */
static void crank_axis(oc_sim *s, int64_t ind0, int64_t indf, double dir[3])
{
    int64_t N = s->N, a, b;
    double mag;
    int i;
    if (ind0 == indf - 1 && ind0 == 0) { a = indf; b = ind0; }
    else if (ind0 == indf - 1 && ind0 == N - 1) { a = ind0; b = ind0 - 1; }
    else if (ind0 == 0 && indf == N) { a = indf - 1; b = ind0; }
    else if (ind0 == 0) { a = indf; b = ind0; }
    else if (indf == N) { a = indf - 1; b = ind0 - 1; }
    else { a = indf; b = ind0 - 1; }
    for (i = 0; i < 3; i++) dir[i] = s->r[3 * a + i] - s->r[3 * b + i];
    mag = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    if (mag < 1E-5) {
        sample_sphere_exact(&s->crng, dir); /* the product keeps libm's acos here (geometry.cuh degenerate_axis) */
    } else {
        double scaling = 1.0 / mag;
        for (i = 0; i < 3; i++) dir[i] = dir[i] * scaling;
    }
}

/* get_crank_shaft_fulcrum move_funcs.pyx:237-280 */
static int64_t crank_fulcrum(const oc_sim *s, int64_t ind0, int64_t indf)
{
    if (ind0 == 0 && indf != s->N) return indf;
    if (ind0 != 0 && indf == s->N) return ind0 - 1;
    if (ind0 == 0 && indf == s->N) return ind0;
    return ind0 - 1;
}

/* Returns n_inds and fills inds_out; writes *_trial rows like the reference.
 * crank_shaft 39-101, end_pivot 285-346, slide 403-465, tangent_rotation
 * 470-582, change_binding_state 717-820 (move_funcs.pyx). */
int64_t oc_propose(oc_sim *s, int move, double amp_move, int64_t amp_bead, int64_t *inds)
{
    oc_glibc_rand *g = &s->crng;
    int64_t N = s->N, ind0, indf, n, i;
    double m[16];
    int j, k;
    switch (move) {
    case OC_CRANK: {
        double ang = amp_move * (oc_uniform(g) - 0.5);
        int64_t b0 = (int)(oc_uniform(g) * (double)N);
        int64_t b1 = oc_from_point(g, amp_bead, N, b0);
        double dir[3];
        int64_t ful;
        b1 = imax(b1, 1);
        check_bead_bounds(b0, b1, N, &ind0, &indf);
        n = indf - ind0;
        for (i = 0; i < n; i++) inds[i] = ind0 + i;
        if (n <= 0) return 0;
        crank_axis(s, ind0, indf, dir);
        ful = crank_fulcrum(s, ind0, indf);
        oc_rotation_matrix(dir, &s->r[3 * ful], ang, m);
        oc_transform_rows(s, m, inds, n);
        return n;
    }
    case OC_PIVOT: {
        double ang = amp_move * (oc_uniform(g) - 0.5);
        int64_t lhs = oc_rand(g) % 2, ful;
        double axis[3];
        if (lhs == 1) {
            ind0 = 0;
            indf = oc_from_left(g, amp_bead, N) + 1;
        } else {
            ind0 = oc_from_right(g, amp_bead, N);
            indf = N;
        }
        n = indf - ind0;
        for (i = 0; i < n; i++) inds[i] = ind0 + i;
        sample_sphere(g, axis);
        /* get_end_pivot_fulcrum move_funcs.pyx:349-398 */
        if (ind0 == 0 && indf != N) ful = indf;
        else if (ind0 != 0 && indf == N) ful = ind0 - 1;
        else if (ind0 == 0 && indf == N && lhs == 1) ful = indf - 1;
        else ful = ind0;
        oc_rotation_matrix(axis, &s->r[3 * ful], ang, m);
        oc_transform_rows(s, m, inds, n);
        return n;
    }
    case OC_SLIDE: {
        double amp = amp_move * oc_uniform(g);
        double dir[3];
        int64_t b0, b1;
        sample_sphere(g, dir);
        for (j = 0; j < 3; j++) dir[j] *= amp;
        b0 = oc_rand(g) % N;
        b1 = oc_from_point(g, amp_bead, N, b0);
        check_bead_bounds(b0, b1, N, &ind0, &indf);
        n = indf - ind0;
        for (i = 0; i < n; i++) {
            inds[i] = ind0 + i;
            for (j = 0; j < 3; j++) {
                s->r_trial[3 * inds[i] + j] = s->r[3 * inds[i] + j] + dir[j];
                s->t3_trial[3 * inds[i] + j] = s->t3[3 * inds[i] + j];
                s->t2_trial[3 * inds[i] + j] = s->t2[3 * inds[i] + j];
            }
        }
        return n;
    }
    case OC_TANGENT: {
        double ang = amp_move * (oc_uniform(g) - 0.5);
        static const double origin[3] = {0., 0., 0.};
        n = oc_rand(g) % amp_bead + 1;
        /* get_inds move_funcs.pyx:552-582: distinct draws, redraw duplicates */
        for (i = 0; i < n; i++) {
            int redraw = 1;
            while (redraw) {
                int64_t c = oc_rand(g) % N, q;
                redraw = 0;
                for (q = 0; q < i; q++)
                    if (inds[q] == c) { redraw = 1; break; }
                if (!redraw) inds[i] = c;
            }
        }
        /* rotate_select_beads move_funcs.pyx:517-549 */
        for (i = 0; i < n; i++) {
            double dir[3];
            int64_t b = inds[i];
            sample_sphere(g, dir);
            oc_rotation_matrix(dir, origin, ang, m);
            for (j = 0; j < 3; j++) {
                double e3 = 0, e2 = 0;
                for (k = 0; k < 3; k++) {
                    e3 += m[4 * j + k] * s->t3[3 * b + k];
                    e2 += m[4 * j + k] * s->t2[3 * b + k];
                }
                s->t3_trial[3 * b + j] = e3;
                s->t2_trial[3 * b + j] = e2;
            }
        }
        return n;
    }
    case OC_BINDING: {
        int64_t binder = oc_rand(g) % s->nb;
        int64_t tails = s->sites_per_bead[binder];
        int64_t b0 = oc_rand(g) % N;
        int64_t b1 = oc_from_point(g, amp_bead, N, b0);
        check_bead_bounds(b0, b1, N, &ind0, &indf);
        n = indf - ind0;
        for (i = 0; i < n; i++) {
            inds[i] = ind0 + i;
            s->states_trial[inds[i] * s->nb + binder] =
                g->alt ? philox_randint((oc_philox *)g->alt, tails + 1) : oc_mt_randint(&s->mt, 0, tails + 1);
        }
        return n;
    }
    }
    return 0;
}

/* ------------------------------------------------ A12: accept / reject */

static void tracker_update(oc_move *mv, double accept)
{
    /* mc_stat.py:190-207 */
    mv->acceptance_rate = (mv->alpha * accept) + (1 - mv->alpha) * mv->acceptance_rate;
}

/* MCAdapter.accept moves.pyx:156-239 (quirk 9) */
void oc_accept(oc_sim *s, oc_move *mv, int move, const int64_t *inds, int64_t n)
{
    int64_t i, j;
    for (i = 0; i < n; i++) {
        int64_t b = inds[i];
        if (move == OC_BINDING) {
            for (j = 0; j < s->nb; j++) s->states[b * s->nb + j] = s->states_trial[b * s->nb + j];
        } else if (move == OC_SLIDE) {
            for (j = 0; j < 3; j++) {
                s->r[3 * b + j] = s->r_trial[3 * b + j];
                s->t3_trial[3 * b + j] = s->t3[3 * b + j];
                s->t2_trial[3 * b + j] = s->t2[3 * b + j];
            }
        } else if (move == OC_TANGENT) {
            for (j = 0; j < 3; j++) {
                s->t3[3 * b + j] = s->t3_trial[3 * b + j];
                s->t2[3 * b + j] = s->t2_trial[3 * b + j];
                s->r_trial[3 * b + j] = s->r[3 * b + j];
            }
        } else {
            for (j = 0; j < 3; j++) {
                s->r[3 * b + j] = s->r_trial[3 * b + j];
                s->t3[3 * b + j] = s->t3_trial[3 * b + j];
                s->t2[3 * b + j] = s->t2_trial[3 * b + j];
            }
        }
    }
    mv->num_success += 1;
    tracker_update(mv, 1.0);
}

/* MCAdapter.reject moves.pyx:241-299 */
void oc_reject(oc_sim *s, oc_move *mv, int move, const int64_t *inds, int64_t n)
{
    int64_t i, j;
    for (i = 0; i < n; i++) {
        int64_t b = inds[i];
        if (move == OC_BINDING) {
            for (j = 0; j < s->nb; j++) s->states_trial[b * s->nb + j] = s->states[b * s->nb + j];
        } else {
            for (j = 0; j < 3; j++) {
                s->r_trial[3 * b + j] = s->r[3 * b + j];
                s->t3_trial[3 * b + j] = s->t3[3 * b + j];
                s->t2_trial[3 * b + j] = s->t2[3 * b + j];
            }
        }
    }
    tracker_update(mv, 0.0);
}

/* SimpleControl.update_move_amplitude mc_controller.py:148-213 */
void oc_update_amplitudes(oc_move *mv)
{
    const double setpoint = 0.5, factor = 0.95;
    double acc = mv->acceptance_rate;
    if (mv->controller != 1) return;
    if (acc < setpoint) {
        double prop = mv->amp_move * factor;
        if (prop > mv->move_amp_lo) {
            mv->amp_move = prop;
        } else {
            double nb = (double)(mv->amp_bead - 1);
            mv->amp_move = mv->move_amp_hi;
            mv->amp_bead = (int64_t)(mv->bead_amp_lo > nb ? mv->bead_amp_lo : nb);
        }
    } else if (acc > setpoint) {
        double prop = mv->amp_move / factor;
        if (prop < mv->move_amp_hi) {
            mv->amp_move = prop;
        } else {
            double nb = (double)(mv->amp_bead + 1);
            mv->amp_move = mv->move_amp_lo;
            mv->amp_bead = (int64_t)(mv->bead_amp_hi < nb ? mv->bead_amp_hi : nb);
        }
    }
}

/* mc_step mc_sim.pyx:106-182.  Returns 1 on accept, 0 on reject, -1 if the
 * proposal was empty (n_inds == 0 -> early return, mc_sim.pyx:151-152). */
int oc_mc_step(oc_sim *s, oc_move *mv, int move, int64_t *inds)
{
    int check_field = (s->field_active && move != OC_TANGENT);
    double dE = 0, exp_dE, u = 0, dEp, dEf = 0;
    int64_t n;
    const int prod = s->crng.alt != NULL;
    if (prod) { /* production streams: attempt t has its own stream, whose draw 0 is the Metropolis uniform */
        philox_seek(&s->philox, s->philox.next_attempt++);
        u = oc_uniform(&s->crng);
    }
    mv->num_attempt += 1; /* MCAdapter.propose moves.pyx:151 */
    n = oc_propose(s, move, mv->amp_move, mv->amp_bead, inds);
    if (n == 0) return -1;
    dEp = oc_poly_dE(s, move, inds, n);
    dE += dEp;
    if (check_field) {
        dEf = oc_field_dE(s, inds, n, move == OC_BINDING);
        dE += dEf;
    }
    s->last_dE_poly = dEp;
    s->last_dE_field = dEf;
    exp_dE = exp(-dE);
    if (!prod) u = oc_uniform(&s->crng);
    s->last_u = u;
    if (u < exp_dE) {
        oc_accept(s, mv, move, inds, n);
        if (check_field) oc_update_affected_densities(s);
        s->last_accept = 1;
        return 1;
    }
    oc_reject(s, mv, move, inds, n);
    s->last_accept = 0;
    return 0;
}

/* mc_sim mc_sim.pyx:26-103 (one polymer; np.random.seed(random_seed) at :81) */
void oc_mc_sim_ordered(oc_sim *s, oc_move mv[OC_NMOVES], int64_t num_mc_steps, uint32_t mt_seed,
                       int64_t *inds, const int32_t *order)
{
    int64_t k, j;
    int ci, c;
    oc_mt_seed(&s->mt, mt_seed);
    for (k = 0; k < num_mc_steps; k++)
        for (ci = 0; ci < OC_NMOVES; ci++) { /* `for controller in mc_move_controllers` mc_sim.pyx:92 */
            c = order ? order[ci] : ci;
            if (mv[c].move_on == 1)
                for (j = 0; j < mv[c].num_per_cycle; j++) (void)oc_mc_step(s, &mv[c], c, inds);
            oc_update_amplitudes(&mv[c]); /* quirk 15: also for moves that are off */
        }
}

void oc_mc_sim(oc_sim *s, oc_move mv[OC_NMOVES], int64_t num_mc_steps, uint32_t mt_seed,
               int64_t *inds)
{
    oc_mc_sim_ordered(s, mv, num_mc_steps, mt_seed, inds, 0);
}
