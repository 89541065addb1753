/*
 * xgrid_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the loop nests the reference generates for the workload
 * kernels of BASELINE.json.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs
 * may load this file; nothing under xgrid_b200/ does.
 *
 * Shape of every function = shape of the reference's generated C
 * (xgrid/lang/generator.py:285-364):
 *   - one full-grid traversal PER stencil statement, in program order, each an
 *     `omp parallel for` (generator.py:353-356);
 *   - the body runs only where the stored grid's int32 boundary mask equals the
 *     statement's mask value (generator.py:297-298); other points are not
 *     written;
 *   - "implicit" statements -- store to level 0 while loading level 0 of the
 *     same grid -- go through a malloc'ed full-grid scratch in two sweeps with
 *     a barrier in between (generator.py:312-352);
 *   - expressions are fully parenthesised in Python-AST order
 *     (generator.py:377-385); `x ** 2.0` is written as a product because gcc
 *     folds pow(x, 2.0) at -O2/-O3;
 *   - loads default to time level 1 ("-1"), stores to level 0
 *     (xgrid/lang/parser.py:523).
 * Deliberate deviation: linear indices use correct C-order
 * strides and 64-bit arithmetic; for 1-D and square 2-D grids this is identical
 * to the reference's formula (generator.py:171-179), for non-square / 3-D it is
 * what the reference intended.
 *
 * Level pointers point at the first real element of a buffer that the Python
 * wrapper pads on both sides, so neighbour reads off the ends are harmless, as
 * they are (by luck) in the reference.
 *
 * Build: gcc -O3 -fopenmp -shared -fpic xgrid_oracle.c -o _build/libxgrid_oracle.so -lm
 * (no -march: the reference never passes one, so no FMA contraction.)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

#define SQ(x) ((x) * (x))

int xo_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* README.md:26-28  result[0] = a[0] * b[0] */
#define DEF_EWMUL(NAME, T)                                                              \
    void NAME(T *out, const T *a, const T *b, const int32_t *mask, i64 n) {             \
        _Pragma("omp parallel for")                                                     \
        for (i64 i = 0; i < n; ++i)                                                     \
            if (mask[i] == 0) out[i] = (a[i] * b[i]);                                   \
    }
DEF_EWMUL(xo_ewmul_f64, double)
DEF_EWMUL(xo_ewmul_f32, float)

/* test.py:214-218  u[0] = u[0] - c * dt / dx * (u[0] - u[-1]); boundary(1): u[0] = 1.0 */
void xo_conv1d_f64(double *u0, const double *u1, const int32_t *mask, i64 n, double c, double dt,
                   double dx) {
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 0) u0[i] = (u1[i] - (((c * dt) / dx) * (u1[i] - u1[i - 1])));
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 1) u0[i] = 1.0;
}

/* test.py:240-244  u[0] = u[0] - u[0] * dt / dx * (u[0] - u[-1]) */
void xo_conv1d_nonlinear_f64(double *u0, const double *u1, const int32_t *mask, i64 n, double dt,
                             double dx) {
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 0) u0[i] = (u1[i] - (((u1[i] * dt) / dx) * (u1[i] - u1[i - 1])));
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 1) u0[i] = 1.0;
}

/* test.py:269-273  u[0] = u[0] + nu * dt / dx ** 2.0 * (u[1] - 2.0 * u[0] + u[-1]) */
void xo_diff1d_f64(double *u0, const double *u1, const int32_t *mask, i64 n, double nu, double dt,
                   double dx) {
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 0)
            u0[i] = (u1[i] + (((nu * dt) / SQ(dx)) * ((u1[i + 1] - (2.0 * u1[i])) + u1[i - 1])));
#pragma omp parallel for
    for (i64 i = 0; i < n; ++i)
        if (mask[i] == 1) u0[i] = 1.0;
}

/* test.py:300-307  cdx = c*dt/dx; cdy = c*dt/dy;
 * u[0,0] = u[0,0] + cdx * (u[0,0] - u[-1,0]) - cdy * (u[0,0] - u[0,-1]); boundary(1): 1.0 */
#define DEF_CONV2D(NAME, T)                                                             \
    void NAME(T *u0, const T *u1, const int32_t *mask, i64 n0, i64 n1, T c, T dt, T dx, T dy) { \
        const T cdx = ((c * dt) / dx);                                                  \
        const T cdy = ((c * dt) / dy);                                                  \
        const i64 n = n0 * n1;                                                          \
        _Pragma("omp parallel for")                                                     \
        for (i64 q = 0; q < n; ++q)                                                     \
            if (mask[q] == 0)                                                           \
                u0[q] = ((u1[q] + (cdx * (u1[q] - u1[q - n1]))) - (cdy * (u1[q] - u1[q - 1]))); \
        _Pragma("omp parallel for")                                                     \
        for (i64 q = 0; q < n; ++q)                                                     \
            if (mask[q] == 1) u0[q] = 1.0;                                              \
    }
DEF_CONV2D(xo_conv2d_f64, double)

/* fp32 build of the same kernel: literals stay double, so the
 * boundary store is (float)1.0 and the arithmetic is float (no literal inside). */
DEF_CONV2D(xo_conv2d_f32, float)

/* 5-point diffusion used for BASELINE config 3:
 * u[0,0] = u[0,0] + a * (u[0,1] + u[0,-1] + u[1,0] + u[-1,0] - 4.0 * u[0,0]); boundary(1): 1.0 */
void xo_diff2d_f64(double *u0, const double *u1, const int32_t *mask, i64 n0, i64 n1, double a) {
    const i64 n = n0 * n1;
#pragma omp parallel for
    for (i64 q = 0; q < n; ++q)
        if (mask[q] == 0)
            u0[q] = (u1[q] + (a * ((((u1[q + 1] + u1[q - 1]) + u1[q + n1]) + u1[q - n1]) - (4.0 * u1[q]))));
#pragma omp parallel for
    for (i64 q = 0; q < n; ++q)
        if (mask[q] == 1) u0[q] = 1.0;
}

/* 3-D 7-point heat / Jacobi step for BASELINE config 5:
 * u[0,0,0] = u[0,0,0] + a * (u[1,0,0] + u[-1,0,0] + u[0,1,0] + u[0,-1,0] + u[0,0,1] + u[0,0,-1]
 *                            - 6.0 * u[0,0,0]); boundary(1): u[0,0,0] = 0.0 */
void xo_heat3d_f64(double *u0, const double *u1, const int32_t *mask, i64 n0, i64 n1, i64 n2, double a) {
    const i64 n = n0 * n1 * n2, s0 = n1 * n2, s1 = n2;
#pragma omp parallel for
    for (i64 q = 0; q < n; ++q)
        if (mask[q] == 0)
            u0[q] = (u1[q] + (a * (((((((u1[q + s0] + u1[q - s0]) + u1[q + s1]) + u1[q - s1]) + u1[q + 1]) +
                                     u1[q - 1]) - (6.0 * u1[q])))));
#pragma omp parallel for
    for (i64 q = 0; q < n; ++q)
        if (mask[q] == 1) u0[q] = 0.0;
}

/* ------------------------------------------------------------------------------
 * examples/cavity.py:74-142 -- one call of cavity_kernel = 311 full-grid sweeps.
 * Pointers: X0 = level 0 (written), X1 = level 1 (default loads).  Masks per grid.
 * ------------------------------------------------------------------------------ */
struct xo_cavity_cfg { double rho, nu, dt, dx, dy; };

#define FOR_MASK(M, K)                                                                  \
    _Pragma("omp parallel for")                                                         \
    for (i64 q = 0; q < n; ++q)                                                         \
        if ((M)[q] == (K))

static void cavity_p_bcs(double *p0, const int32_t *mp, i64 n, i64 s) {
    FOR_MASK(mp, 1) p0[q] = p0[q - 1];   /* dp/dx = 0 at x = 2 */
    FOR_MASK(mp, 2) p0[q] = p0[q + s];   /* dp/dy = 0 at y = 0 */
    FOR_MASK(mp, 3) p0[q] = p0[q + 1];   /* dp/dx = 0 at x = 0 */
    FOR_MASK(mp, 4) p0[q] = 0.0;         /* p = 0 at y = 2 */
}

void xo_cavity_f64(double *b0, double *p0, const double *p1, double *u0, const double *u1, double *v0,
                   const double *v1, const int32_t *mb, const int32_t *mp, const int32_t *mu,
                   const int32_t *mv, i64 n0, i64 n1, struct xo_cavity_cfg cfg, int nit) {
    const i64 n = n0 * n1, s = n1;
    const double rho = cfg.rho, nu = cfg.nu, dt = cfg.dt, dx = cfg.dx, dy = cfg.dy;

    /* b[0,0] = rho * (1/dt * (du/dx + dv/dy) - (du/dx)^2 - 2 (du/dy dv/dx) - (dv/dy)^2) */
    FOR_MASK(mb, 0)
        b0[q] = (rho * (((((1.0 / dt) * (((u1[q + 1] - u1[q - 1]) / (2.0 * dx)) +
                                          ((v1[q + s] - v1[q - s]) / (2.0 * dy)))) -
                          SQ(((u1[q + 1] - u1[q - 1]) / (2.0 * dx)))) -
                         (2.0 * ((((u1[q + s] - u1[q - s]) / (2.0 * dy)) * (v1[q + 1] - v1[q - 1])) /
                                 (2.0 * dx)))) -
                        SQ(((v1[q + s] - v1[q - s]) / (2.0 * dy)))));

    /* first pressure sweep: explicit, loads p at level 1 and b at level 0 */
    FOR_MASK(mp, 0)
        p0[q] = ((((((p1[q + 1] + p1[q - 1]) * SQ(dy)) + ((p1[q + s] + p1[q - s]) * SQ(dx)))) /
                  (2.0 * (SQ(dx) + SQ(dy)))) -
                 (((SQ(dx) * SQ(dy)) / (2.0 * (SQ(dx) + SQ(dy)))) * b0[q]));
    cavity_p_bcs(p0, mp, n, s);

    /* nit Jacobi sweeps: implicit (loads p at level 0) -> scratch + barrier + copy-back,
     * with a fresh malloc/free per statement execution like the reference. */
    for (int it = 0; it < nit; ++it) {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel
        {
#pragma omp for
            for (i64 q = 0; q < n; ++q)
                if (mp[q] == 0)
                    tmp[q] = ((((((p0[q + 1] + p0[q - 1]) * SQ(dy)) + ((p0[q + s] + p0[q - s]) * SQ(dx)))) /
                               (2.0 * (SQ(dx) + SQ(dy)))) -
                              (((SQ(dx) * SQ(dy)) / (2.0 * (SQ(dx) + SQ(dy)))) * b0[q]));
#pragma omp barrier
#pragma omp for
            for (i64 q = 0; q < n; ++q)
                if (mp[q] == 0) p0[q] = tmp[q];
        }
        free(tmp);
        cavity_p_bcs(p0, mp, n, s);
    }

    FOR_MASK(mu, 0)
        u0[q] = ((((u1[q] - (((u1[q] * dt) / dx) * (u1[q] - u1[q - 1]))) -
                   (((v1[q] * dt) / dy) * (u1[q] - u1[q - s]))) -
                  ((dt / ((2.0 * rho) * dx)) * (p0[q + 1] - p0[q - 1]))) +
                 (nu * (((dt / SQ(dx)) * ((u1[q + 1] - (2.0 * u1[q])) + u1[q - 1])) +
                        ((dt / SQ(dy)) * ((u1[q + s] - (2.0 * u1[q])) + u1[q - s])))));

    FOR_MASK(mv, 0)
        v0[q] = ((((v1[q] - (((u1[q] * dt) / dx) * (v1[q] - v1[q - 1]))) -
                   (((v1[q] * dt) / dy) * (v1[q] - v1[q - s]))) -
                  ((dt / ((2.0 * rho) * dy)) * (p0[q + s] - p0[q - s]))) +
                 (nu * (((dt / SQ(dx)) * ((v1[q + 1] - (2.0 * v1[q])) + v1[q - 1])) +
                        ((dt / SQ(dy)) * ((v1[q + s] - (2.0 * v1[q])) + v1[q - s])))));

    FOR_MASK(mu, 1) u0[q] = 0.0;
    FOR_MASK(mv, 1) v0[q] = 0.0;
    FOR_MASK(mu, 2) u0[q] = 1.0;
}

/* test.py:171-172  a[0, 0] = 4 on an int grid */
void xo_fill_i32(int32_t *a0, const int32_t *mask, i64 n, int32_t value) {
#pragma omp parallel for
    for (i64 q = 0; q < n; ++q)
        if (mask[q] == 0) a0[q] = value;
}
