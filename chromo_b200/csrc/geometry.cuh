// geometry.cuh -- per-bead device functions: trilinear binning, affine
// transforms, SSWLC pair energy.  Compiled with -fmad=false: the reference is
// built for generic x86-64 (no FMA contraction), and a contracted
// `x/dx - ind` could flip the last ulp of a weight or a floor().
#pragma once
#include "launch.cuh"
#include "params.cuh"

// ---------------------------------------------------------------- binning
// Reference: UniformDensityField.get_change_in_density fields.pyx:1437-1462
// (wrap, lower index, lower weight), _generate_weight_vector_with_trial
// 1524-1598 (8 weights), _generate_index_vector_with_trial 1600-1673 and the
// wrap table 673-685 (8 super-indices  ix + nx*iy + nx*ny*iz).
//
// Bit-exact contract: `%` is Python-modulo on doubles (cdivision=False,
// setup.py:89): fmod, then +W when the remainder is negative.  The weight uses
// the UNPATCHED floor index even when it is -1 (fields.pyx:1451-1462).
// a / b for a divisor whose correctly rounded reciprocal y = RN(1/b) is known:
// q = RN(a y), r = a - q b (exact, one fma), RN(q + r y) is the correctly
// rounded quotient (Markstein's division step) -- bit-identical to IEEE
// division (checked on 8.6e8 operands incl. near-integer quotients,
// profiles/README) at 3 instructions instead of ~13.
__device__ __forceinline__ double div_const(double a, double b, double y) {
    double q = a * y;
    double r = fma(-q, b, a);
    return fma(r, y, q);
}

// exact fmod for w > 0: the remainder is exactly representable, so one
// division, one trunc and one fma reproduce C's fmod bit for bit (a / w can only
// round UP to the next integer, which shows as a remainder of the wrong sign).
__device__ __forceinline__ double fmod_exact(double a, double w, double inv_w) {
    double q = trunc(div_const(a, w, inv_w));
    double r = fma(-q, w, a);
    if (a >= 0.0) {
        if (r < 0.0) r += w;
    } else if (r > 0.0) {
        r -= w;
    }
    return r;
}

__device__ __forceinline__ void bin_axis(double x, double half_width, double width, double inv_width,
                                         double half_step, double d, double inv_d, int n, int &ind_lo,
                                         int &ind_hi, double &w_lo) {
    double a = x + half_width;
    double m = a; // a % width == a for 0 <= a < width: every bead of a confined chain, no division needed
    if (!(a >= 0.0 && a < width)) {
        m = fmod_exact(a, width, inv_width);
        if (m != 0.0 && m < 0.0) m = m + width; // Python modulo: result takes the divisor's sign
    }
    double xs = m - half_step;
    double q = div_const(xs, d, inv_d); // == xs / d
    double fl = floor(q);
    int ind = (int)fl;
    w_lo = 1.0 - (q - fl);
    ind_lo = (ind == -1) ? n - 1 : ind;
    ind_hi = (ind_lo + 1 >= n) ? ind_lo + 1 - n : ind_lo + 1;
}

// fast_field = 1 (init_fast_field fields.pyx:577-671, get_change_in_density_quickly 1235-1368): the position is
// quantised to sub-bin i = floor((x + W/2) / sbw) % (n_points n) (numpy floor, Python modulo) and binned from the
// sub-bin's lower edge x_i = (i - n_points/2) sbw: ind = floor(x_i / d), weight 1 - (x_i / d - ind).  The
// reference reads these from per-axis tables built once; the same IEEE operations here give the same doubles.
__device__ __forceinline__ void bin_axis_fast(double x, double half_width, double sbw, double d, int n, int npts,
                                              int &ind_lo, int &ind_hi, double &w_lo) {
    const double nsub = (double)npts * (double)n;
    double f = floor((x + half_width) / sbw);
    double sub = fmod(f, nsub);
    if (sub != 0.0 && sub < 0.0) sub += nsub;
    const double xi = ((double)((long long)sub - npts / 2) * sbw) / d;
    const double fl = floor(xi);
    const int ind = (int)fl;
    w_lo = 1.0 - (xi - fl);
    ind_lo = (ind == -1) ? n - 1 : ind;
    ind_hi = (ind_lo + 1 >= n) ? ind_lo + 1 - n : ind_lo + 1;
}
static __device__ CB_NOINLINE void bin_axes_fast(const DevCtx &C, const double p[3], int lo[3], int hi[3], double wl[3]) {
    bin_axis_fast(p[0], C.half_width[0], C.fast_sbw[0], C.dxyz[0], C.nx, C.fast_n, lo[0], hi[0], wl[0]);
    bin_axis_fast(p[1], C.half_width[1], C.fast_sbw[1], C.dxyz[1], C.ny, C.fast_n, lo[1], hi[1], wl[1]);
    bin_axis_fast(p[2], C.half_width[2], C.fast_sbw[2], C.dxyz[2], C.nz, C.fast_n, lo[2], hi[2], wl[2]);
}

// lower / upper voxel index and lower-voxel weight along each axis
__device__ __forceinline__ void bin_axes(const DevCtx &C, const double p[3], int lo[3], int hi[3],
                                         double wl[3]) {
    bin_axis(p[0], C.half_width[0], C.width[0], C.inv_width[0], C.half_step[0], C.dxyz[0], C.inv_dxyz[0], C.nx, lo[0], hi[0], wl[0]);
    bin_axis(p[1], C.half_width[1], C.width[1], C.inv_width[1], C.half_step[1], C.dxyz[1], C.inv_dxyz[1], C.ny, lo[1], hi[1], wl[1]);
    bin_axis(p[2], C.half_width[2], C.width[2], C.inv_width[2], C.half_step[2], C.dxyz[2], C.inv_dxyz[2], C.nz, lo[2], hi[2], wl[2]);
}

__device__ __forceinline__ void bin_point(const DevCtx &C, double x, double y, double z,
                                          int idx[8], double w[8]) {
    int x0, x1, y0, y1, z0, z1;
    double wx, wy, wz;
    bin_axis(x, C.half_width[0], C.width[0], C.inv_width[0], C.half_step[0], C.dxyz[0], C.inv_dxyz[0], C.nx, x0, x1, wx);
    bin_axis(y, C.half_width[1], C.width[1], C.inv_width[1], C.half_step[1], C.dxyz[1], C.inv_dxyz[1], C.ny, y0, y1, wy);
    bin_axis(z, C.half_width[2], C.width[2], C.inv_width[2], C.half_step[2], C.dxyz[2], C.inv_dxyz[2], C.nz, z0, z1, wz);
    double ux = 1.0 - wx, uy = 1.0 - wy, uz = 1.0 - wz;
    // l = bit0:x, bit1:y, bit2:z ; products in the reference's order (x*y)*z
    w[0] = wx * wy * wz;
    w[1] = ux * wy * wz;
    w[2] = wx * uy * wz;
    w[3] = ux * uy * wz;
    w[4] = wx * wy * uz;
    w[5] = ux * wy * uz;
    w[6] = wx * uy * uz;
    w[7] = ux * uy * uz;
    int nxy = C.nx * C.ny;
    int r00 = C.nx * y0 + nxy * z0, r10 = C.nx * y1 + nxy * z0;
    int r01 = C.nx * y0 + nxy * z1, r11 = C.nx * y1 + nxy * z1;
    idx[0] = x0 + r00;
    idx[1] = x1 + r00;
    idx[2] = x0 + r10;
    idx[3] = x1 + r10;
    idx[4] = x0 + r01;
    idx[5] = x1 + r01;
    idx[6] = x0 + r11;
    idx[7] = x1 + r11;
}

// ------------------------------------------------------------- transforms
// libdevice's fp64 sincos / acos are ~100 instructions each when inlined; one
// out-of-line copy keeps the kernel inside the instruction cache
static __device__ CB_NOINLINE double2 sincos_ni(double x) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(s, c);
}
static __device__ CB_NOINLINE double acos_ni(double x) { return acos(x); }
static __device__ CB_NOINLINE double2 sincospi_ni(double x) {
    double s, c;
    sincospi(x, &s, &c);
    return make_double2(s, c);
}

// Point on the unit sphere from two uniforms u1, u2 in [0, 1]
// (uniform_sample_unit_sphere linalg.pyx:23-59: phi = 2 pi u1, theta = acos(2 u2 - 1),
// (cos phi sin theta, sin phi sin theta, cos theta)).
//   EXACT = true : the reference's own sequence of libm calls (replay / parity mode: the same
//                  draws give the same axis to the last bit under the same libm);
//   EXACT = false: cos theta = x and sin theta = sqrt((1 - x)(1 + x)) for x = 2 u2 - 1, and
//                  sincospi(2 u1): the same point to ~1 ulp at a third of the instructions and
//                  without the out-of-line acos / sincos (production mode, whose counter-based
//                  draws are not the reference's anyway).
template <bool EXACT>
__device__ __forceinline__ void unit_sphere_point(double u1, double u2, double v[3]) {
    if (EXACT) {
        const double2 ph = sincos_ni(u1 * (2.0 * 3.14159265358979323846));
        const double2 th = sincos_ni(acos_ni(u2 * 2.0 - 1.0));
        v[0] = ph.y * th.x;
        v[1] = ph.x * th.x;
        v[2] = th.y;
    } else {
        const double x = u2 * 2.0 - 1.0;
        const double st = sqrt((1.0 - x) * (1.0 + x));
        const double2 ph = sincospi_ni(u1 * 2.0);
        v[0] = ph.y * st;
        v[1] = ph.x * st;
        v[2] = x;
    }
}

// arbitrary_axis_rotation linalg.pyx:62-139 given sin / cos of the angle.  M is
// 3x4 row-major (M[4*j+3] is the translation column).
__device__ __forceinline__ void rotation_matrix_sc(const double ax[3], const double pt[3],
                                                   double sn, double c, double M[12]) {
    double omc = 1.0 - c;
    M[0] = ax[0] * ax[0] + (ax[1] * ax[1] + ax[2] * ax[2]) * c;
    M[1] = ax[0] * ax[1] * omc - ax[2] * sn;
    M[2] = ax[0] * ax[2] * omc + ax[1] * sn;
    M[4] = ax[0] * ax[1] * omc + ax[2] * sn;
    M[5] = ax[1] * ax[1] + (ax[0] * ax[0] + ax[2] * ax[2]) * c;
    M[6] = ax[1] * ax[2] * omc - ax[0] * sn;
    M[8] = ax[0] * ax[2] * omc - ax[1] * sn;
    M[9] = ax[1] * ax[2] * omc + ax[0] * sn;
    M[10] = ax[2] * ax[2] + (ax[0] * ax[0] + ax[1] * ax[1]) * c;
    double r0 = (pt[1] * ax[2] - pt[2] * ax[1]) * sn;
    double r1 = (pt[2] * ax[0] - pt[0] * ax[2]) * sn;
    double r2 = (pt[0] * ax[1] - pt[1] * ax[0]) * sn;
    r0 += (pt[0] * (1.0 - ax[0] * ax[0]) - ax[0] * (pt[1] * ax[1] + pt[2] * ax[2])) * omc;
    r1 += (pt[1] * (1.0 - ax[1] * ax[1]) - ax[1] * (pt[0] * ax[0] + pt[2] * ax[2])) * omc;
    r2 += (pt[2] * (1.0 - ax[2] * ax[2]) - ax[2] * (pt[0] * ax[0] + pt[1] * ax[1])) * omc;
    M[3] = r0;
    M[7] = r1;
    M[11] = r2;
}

// the same matrix in two parts: 3x3 rotation (state-independent) ...
__device__ __forceinline__ void rotation_3x3(const double ax[3], double sn, double c, double R[9]) {
    double omc = 1.0 - c;
    R[0] = ax[0] * ax[0] + (ax[1] * ax[1] + ax[2] * ax[2]) * c;
    R[1] = ax[0] * ax[1] * omc - ax[2] * sn;
    R[2] = ax[0] * ax[2] * omc + ax[1] * sn;
    R[3] = ax[0] * ax[1] * omc + ax[2] * sn;
    R[4] = ax[1] * ax[1] + (ax[0] * ax[0] + ax[2] * ax[2]) * c;
    R[5] = ax[1] * ax[2] * omc - ax[0] * sn;
    R[6] = ax[0] * ax[2] * omc - ax[1] * sn;
    R[7] = ax[1] * ax[2] * omc + ax[0] * sn;
    R[8] = ax[2] * ax[2] + (ax[0] * ax[0] + ax[1] * ax[1]) * c;
}
// ... and the translation column, which needs the fulcrum (linalg.pyx:115-139)
__device__ __forceinline__ void rotation_translation(const double ax[3], const double pt[3], double sn,
                                                     double c, double tv[3]) {
    double omc = 1.0 - c;
    double r0 = (pt[1] * ax[2] - pt[2] * ax[1]) * sn;
    double r1 = (pt[2] * ax[0] - pt[0] * ax[2]) * sn;
    double r2 = (pt[0] * ax[1] - pt[1] * ax[0]) * sn;
    r0 += (pt[0] * (1.0 - ax[0] * ax[0]) - ax[0] * (pt[1] * ax[1] + pt[2] * ax[2])) * omc;
    r1 += (pt[1] * (1.0 - ax[1] * ax[1]) - ax[1] * (pt[0] * ax[0] + pt[2] * ax[2])) * omc;
    r2 += (pt[2] * (1.0 - ax[2] * ax[2]) - ax[2] * (pt[0] * ax[0] + pt[1] * ax[1])) * omc;
    tv[0] = r0;
    tv[1] = r1;
    tv[2] = r2;
}
__device__ __forceinline__ void apply_rot3(const double *R, const double v[3], double o[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++) o[j] = (R[3 * j] * v[0] + R[3 * j + 1] * v[1]) + R[3 * j + 2] * v[2];
}
// axis of a degenerate crank-shaft (end beads coincide): rare, kept out of line
static __device__ CB_NOINLINE void degenerate_axis(uint32_t d1, uint32_t d2, double *axis) {
    double phi = (double)d1 / CB_RAND_MAX * (2.0 * 3.14159265358979323846);
    double theta = acos_ni(((double)d2 / CB_RAND_MAX) * 2.0 - 1.0);
    const double2 p = sincos_ni(phi), t = sincos_ni(theta);
    axis[0] = p.y * t.x;
    axis[1] = p.x * t.x;
    axis[2] = t.y;
}

// transform_r_t3_t2 move_funcs.pyx:121-154: ((m0 x + m1 y) + m2 z) (+ t)
__device__ __forceinline__ void apply_rot(const double *M, const double v[3], double o[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++) o[j] = (M[4 * j] * v[0] + M[4 * j + 1] * v[1]) + M[4 * j + 2] * v[2];
}
__device__ __forceinline__ void apply_affine(const double *M, const double v[3], double o[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++)
        o[j] = ((M[4 * j] * v[0] + M[4 * j + 1] * v[1]) + M[4 * j + 2] * v[2]) + M[4 * j + 3];
}

// uniform_sample_unit_sphere linalg.pyx:23-59 from two rand() outputs
__device__ __forceinline__ void sphere_from_draws(uint32_t d1, uint32_t d2, double v[3]) {
    double phi = (double)d1 / CB_RAND_MAX * (2.0 * 3.14159265358979323846);
    double theta = acos_ni(((double)d2 / CB_RAND_MAX) * 2.0 - 1.0);
    const double2 p = sincos_ni(phi), t = sincos_ni(theta);
    v[0] = p.y * t.x;
    v[1] = p.x * t.x;
    v[2] = t.y;
}

// ---------------------------------------------------------- SSWLC energy
struct Bond {
    double eps_bend, eps_par, eps_perp, gamma, eta;
};
__device__ __forceinline__ Bond load_bond(const DevCtx &C, int rep, int b) {
    const double *p = C.bond + (long long)rep * C.bond_stride + (long long)b * 5;
    Bond o;
    o.eps_bend = p[0];
    o.eps_par = p[1];
    o.eps_perp = p[2];
    o.gamma = p[3];
    o.eta = p[4];
    return o;
}

__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; // ((0+p0)+p1)+p2, linalg.pyx:374-395
}

// E of one bond given the tangent of its first bead, polymers.pyx:1148-1175
// with dr / dr_par / dr_perp / bend built as in bead_pair_dE_poly_forward
// (polymers.pyx:1253-1271).
__device__ __forceinline__ double bond_energy(const Bond &B, const double r0[3],
                                              const double r1[3], const double t0[3],
                                              const double t1[3]) {
    double dr[3], perp[3], bend[3];
#pragma unroll
    for (int i = 0; i < 3; i++) dr[i] = r1[i] - r0[i];
    double par = dot3(t0, dr);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        perp[i] = dr[i] - t0[i] * par;
        bend[i] = (t1[i] - t0[i]) - perp[i] * B.eta;
    }
    double tp = par - B.gamma;
    return (0.5 * B.eps_bend * dot3(bend, bend) + 0.5 * B.eps_par * (tp * tp)) +
           0.5 * B.eps_perp * dot3(perp, perp);
}

// compute_twist_angle_omega polymers.pyx:3427-3461: t1 = t2 x t3 for both beads,
// omega = atan2(t2_0.t1_1 - t1_0.t2_1, t1_0.t1_1 + t2_0.t2_1)
__device__ __forceinline__ double twist_omega(const double t2_0[3], const double t3_0[3], const double t2_1[3],
                                              const double t3_1[3]) {
    double a[3], b[3];
    a[0] = t2_0[1] * t3_0[2] - t2_0[2] * t3_0[1];
    a[1] = t2_0[2] * t3_0[0] - t2_0[0] * t3_0[2];
    a[2] = t2_0[0] * t3_0[1] - t2_0[1] * t3_0[0];
    b[0] = t2_1[1] * t3_1[2] - t2_1[2] * t3_1[1];
    b[1] = t2_1[2] * t3_1[0] - t2_1[0] * t3_1[2];
    b[2] = t2_1[0] * t3_1[1] - t2_1[1] * t3_1[0];
    return atan2(dot3(t2_0, b) - dot3(a, t2_1), dot3(a, b) + dot3(t2_0, t2_1));
}
// the twist term of E_pair_with_twist polymers.pyx:2050-2102; tw = {eps_twist, natural twist} of the bond
__device__ __forceinline__ double twist_energy(const double *tw, double omega) {
    const double pi = 3.141592653589793, two_pi = 2.0 * pi;
    double d = omega - tw[1];
    d -= two_pi * floor((d + pi) / two_pi);
    return 0.5 * tw[0] * (d * d);
}

// ---- DetailedChromatin (polymers.pyx:2455-2607): the linker DNA leaves a nucleosome at its EXIT point with the
// exit frame and arrives at the next one's ENTRY point with that nucleosome's own (t3, t2)
// (DetailedNucleosome.update_configuration beads.py:536-574).  det[20] = t3_local[3] | t2_local[3] |
// r_enter_unit[3] | r_enter_norm | r_exit_unit[3] | r_exit_norm | a3[3] | a1[3] (constants of one bp_wrap, built on
// the host by chromo_b200.util.nucleo_geom).  The rotation local -> global is built exactly as the reference does:
// align t3_local with t3, then the rotated t2_local with t2 (rotation_matrix_from_vectors linalg.pyx:478-510,
// get_rotation_matrix 537-575 incl. its antiparallel fix-up).  Rare model, kept out of line.
__device__ __forceinline__ void cb_cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void cb_mat3_vec(const double a[9], const double v[3], double o[3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
__device__ __forceinline__ void cb_mat3_mul(const double a[9], const double b[9], double o[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) t += a[3 * i + k] * b[3 * k + j];
            o[3 * i + j] = t;
        }
}
__device__ __forceinline__ void cb_rot_from_vectors(const double v1[3], const double v2[3], double Rm[9]) {
    const double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    const double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    double a[3], b[3], v[3], K[9], K2[9];
#pragma unroll
    for (int i = 0; i < 3; i++) a[i] = v1[i] / n1, b[i] = v2[i] / n2;
    cb_cross3(a, b, v);
#pragma unroll
    for (int i = 0; i < 9; i++) Rm[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (v[0] != 0.0 || v[1] != 0.0 || v[2] != 0.0) {
        const double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
        const double sn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const double f = (1.0 - c) / (sn * sn);
        K[0] = 0.0, K[1] = -v[2], K[2] = v[1];
        K[3] = v[2], K[4] = 0.0, K[5] = -v[0];
        K[6] = -v[1], K[7] = v[0], K[8] = 0.0;
        cb_mat3_mul(K, K, K2);
#pragma unroll
        for (int i = 0; i < 9; i++) Rm[i] = Rm[i] + K[i] + K2[i] * f;
    }
}
// in place: (r, t3, t2) of a nucleosome -> exit point and exit frame (`exit` != 0) or entry point (frame unchanged)
static __device__ CB_NOINLINE void nucleosome_frame(const double *det, double *r, double *t3, double *t2, int exit) {
    const double *t3l = det, *t2l = det + 3, *ren = det + 6, *rex = det + 10, *a3 = det + 14, *a1 = det + 17;
    double R1[9], R2[9], Rm[9], t2r[3], chk[3];
    cb_rot_from_vectors(t3l, t3, R1);
    cb_mat3_vec(R1, t2l, t2r);
    cb_rot_from_vectors(t2r, t2, R2);
    cb_mat3_mul(R2, R1, Rm);
    cb_mat3_vec(Rm, t3l, chk);
    bool anti = true; // np.allclose(R t3_local, -t3): |x - y| <= 1e-8 + 1e-5 |y|
#pragma unroll
    for (int i = 0; i < 3; i++) anti = anti && (fabs(chk[i] + t3[i]) <= 1e-8 + 1e-5 * fabs(t3[i]));
    if (anti) { // get_arbitrary_axis_rotation_matrix(t2, pi) linalg.pyx:512-535
        const double n = sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
        const double ux = t2[0] / n, uy = t2[1] / n, uz = t2[2] / n;
        const double ct = -1.0, st = 1.2246467991473532e-16; // cos(pi), sin(pi) as libm returns them
        double R3[9], Rn[9];
        R3[0] = ct + ux * ux * (1 - ct), R3[1] = ux * uy * (1 - ct) - uz * st, R3[2] = uz * ux * (1 - ct) + uy * st;
        R3[3] = ux * uy * (1 - ct) + uz * st, R3[4] = ct + uy * uy * (1 - ct), R3[5] = uy * uz * (1 - ct) - ux * st;
        R3[6] = uz * ux * (1 - ct) - uy * st, R3[7] = uy * uz * (1 - ct) + ux * st, R3[8] = ct + uz * uz * (1 - ct);
        cb_mat3_mul(R3, Rm, Rn);
#pragma unroll
        for (int i = 0; i < 9; i++) Rm[i] = Rn[i];
    }
    double v[3];
    if (exit) {
        double t1[3], t3e[3], t1e[3], t2e[3];
        cb_cross3(t2, t3, t1);
        cb_mat3_vec(Rm, rex, v);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            r[i] = v[i] * rex[3] + r[i];
            t3e[i] = a3[0] * t3[i] + a3[1] * t2[i] + a3[2] * t1[i];
            t1e[i] = a1[0] * t3[i] + a1[1] * t2[i] + a1[2] * t1[i];
        }
        cb_cross3(t3e, t1e, t2e);
#pragma unroll
        for (int i = 0; i < 3; i++) t3[i] = t3e[i], t2[i] = t2e[i];
    } else {
        cb_mat3_vec(Rm, ren, v);
#pragma unroll
        for (int i = 0; i < 3; i++) r[i] = v[i] * ren[3] + r[i];
    }
}

__device__ __forceinline__ void load3(const double *p, double v[3]) {
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
}
__device__ __forceinline__ void store3(double *p, const double v[3]) {
    p[0] = v[0];
    p[1] = v[1];
    p[2] = v[2];
}
