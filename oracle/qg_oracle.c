/* C restatement of the reference's hot path — TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Nothing here is shipped or called by the product (libqgb200).  It exists so that
 *   (a) the GPU box can time "the reference's algorithm on host cores" in the same run as the
 *       CUDA path (bench.py cpu_baseline and `--impl reference`), the reference itself being
 *       Julia, which is not installed in this image, and
 *   (b) tests have a third, independent implementation of the inversion (x-FFT + cyclic Thomas
 *       with Sherman-Morrison) next to the NumPy oracle's SuperLU and 2-D FFT back-ends.
 * It is validated against oracle/qg_oracle.py in tests/test_oracle_c.py.
 *
 * Structure follows the reference operator by operator (one full-field temporary per
 * operator, ghost refresh after each, as in src/schemes/*.jl); loops over y are OpenMP
 * parallel so that "all host threads" can be used.  The sparse Cholesky solves of
 * src/model.jl:186,191 are replaced by an exact spectral solve of the same matrices
 * (src/schemes/laplacian.jl:54-75) because CHOLMOD is not available here and its fill-in
 * makes grids beyond ~1024^2 impractical; this favours the CPU baseline.
 *
 * Arrays: reference layout, column-major (M+2, P+2, 2, 3), level 0 newest.
 * Parity: pinned only through the reference's unit-level known-answer tests as restated for the
 * NumPy oracle, which this file is checked against; for multi-step trajectories it is, like that
 * oracle, PARITY UNPINNED (the reference holds no golden trajectory and cannot run here).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int M, P;
    double dx, dt, visc, r, U, beta1, beta2, alpha;
    double Pinv[4], Pfwd[4];
    double H1, H2, S1;
} qgo_params;

#define IDX(i, j) ((size_t)(i) + (size_t)(W) * (size_t)(j))

/* src/schemes/boundary_conditions.jl:2-13 */
static void bc(double* b, int M, int P) {
    const int W = M + 2, Hh = P + 2;
    for (int i = 1; i <= M; ++i) { b[IDX(i, 0)] = b[IDX(i, Hh - 2)]; b[IDX(i, Hh - 1)] = b[IDX(i, 1)]; }
    for (int j = 1; j <= P; ++j) { b[IDX(0, j)] = b[IDX(W - 2, j)]; b[IDX(W - 1, j)] = b[IDX(1, j)]; }
    b[IDX(0, 0)] = b[IDX(W - 2, Hh - 2)];
    b[IDX(0, Hh - 1)] = b[IDX(W - 2, 1)];
    b[IDX(W - 1, Hh - 1)] = b[IDX(1, 1)];
    b[IDX(W - 1, 0)] = b[IDX(1, Hh - 2)];
}

static double* newfield(int M, int P) { return (double*)calloc((size_t)(M + 2) * (P + 2), sizeof(double)); }

/* src/schemes/laplacian.jl:15-27 */
static double* laplace_5p(const double* u, int M, int P, double dx) {
    const int W = M + 2;
    const double i1 = 1.0 / dx, idx2 = i1 * i1;
    double* lap = newfield(M, P);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= P; ++j)
        for (int i = 1; i <= M; ++i)
            lap[IDX(i, j)] = (u[IDX(i - 1, j)] + u[IDX(i + 1, j)] - 4 * u[IDX(i, j)] + u[IDX(i, j - 1)] + u[IDX(i, j + 1)]) * idx2;
    bc(lap, M, P);
    return lap;
}

/* src/model.jl:68-80 */
static double* cd(const double* u, int M, int P, double dx) {
    const int W = M + 2;
    const double h = 0.5 * (1.0 / dx);
    double* out = newfield(M, P);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= P; ++j)
        for (int i = 1; i <= M; ++i) out[IDX(i, j)] = h * (u[IDX(i + 1, j)] - u[IDX(i - 1, j)]);
    bc(out, M, P);
    return out;
}

/* src/schemes/arakawa.jl:7-62 */
static double* jacobian(double dx, const double* z, const double* p, int M, int P) {
    const int W = M + 2;
    double* jpp = newfield(M, P);
    double* jpt = newfield(M, P);
    double* jtp = newfield(M, P);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= P; ++j)
        for (int i = 1; i <= M; ++i) {
            jpp[IDX(i, j)] = (z[IDX(i + 1, j)] - z[IDX(i - 1, j)]) * (p[IDX(i, j + 1)] - p[IDX(i, j - 1)]) -
                             (z[IDX(i, j + 1)] - z[IDX(i, j - 1)]) * (p[IDX(i + 1, j)] - p[IDX(i - 1, j)]);
            jpt[IDX(i, j)] = z[IDX(i + 1, j)] * (p[IDX(i + 1, j + 1)] - p[IDX(i + 1, j - 1)]) -
                             z[IDX(i - 1, j)] * (p[IDX(i - 1, j + 1)] - p[IDX(i - 1, j - 1)]) -
                             z[IDX(i, j + 1)] * (p[IDX(i + 1, j + 1)] - p[IDX(i - 1, j + 1)]) +
                             z[IDX(i, j - 1)] * (p[IDX(i + 1, j - 1)] - p[IDX(i - 1, j - 1)]);
            jtp[IDX(i, j)] = z[IDX(i + 1, j + 1)] * (p[IDX(i, j + 1)] - p[IDX(i + 1, j)]) -
                             z[IDX(i - 1, j - 1)] * (p[IDX(i - 1, j)] - p[IDX(i, j - 1)]) -
                             z[IDX(i - 1, j + 1)] * (p[IDX(i, j + 1)] - p[IDX(i - 1, j)]) +
                             z[IDX(i + 1, j - 1)] * (p[IDX(i + 1, j)] - p[IDX(i, j - 1)]);
        }
    const double den = 3 * 4 * (dx * dx);
    const size_t n = (size_t)(M + 2) * (P + 2);
#pragma omp parallel for schedule(static)
    for (size_t e = 0; e < n; ++e) jpp[e] = (jpp[e] + jpt[e] + jtp[e]) / den;
    bc(jpp, M, P);
    free(jpt);
    free(jtp);
    return jpp;
}

/* src/model.jl:139-153 */
static double* zeta_rhs(const qgo_params* m, int layer, const double* zeta, const double* psi) {
    const int M = m->M, P = m->P;
    const size_t n = (size_t)(M + 2) * (P + 2);
    double* l1 = laplace_5p(psi, M, P, m->dx);
    double* v = laplace_5p(l1, M, P, m->dx);
    double* J = jacobian(m->dx, zeta, psi, M, P);
    double* bt = cd(psi, M, P, m->dx);
    double* last = layer == 0 ? cd(zeta, M, P, m->dx) : l1;
    const double beta = layer == 0 ? m->beta1 : m->beta2;
    const double cl = layer == 0 ? m->U : m->r;
#pragma omp parallel for schedule(static)
    for (size_t e = 0; e < n; ++e) v[e] = m->visc * v[e] - J[e] - beta * bt[e] - cl * last[e];
    free(J);
    free(bt);
    if (layer == 0) free(last);
    free(l1);
    return v;
}

static void store_new_state(double* arr, const double* ns, int z, size_t fs) {
    /* src/model.jl:102-106; arr index (z + 2*t)*fs */
    memcpy(arr + (z + 4) * fs, arr + (z + 2) * fs, fs * sizeof(double));
    memcpy(arr + (z + 2) * fs, arr + (z + 0) * fs, fs * sizeof(double));
    memcpy(arr + (z + 0) * fs, ns, fs * sizeof(double));
}

/* src/model.jl:155-170 */
void qgo_evolve_zeta(const qgo_params* m, double* zeta, const double* psi, int timestep, double* f_store) {
    const size_t fs = (size_t)(m->M + 2) * (m->P + 2);
    for (int layer = 0; layer < 2; ++layer) {
        double* f1 = zeta_rhs(m, layer, zeta + layer * fs, psi + layer * fs);
        store_new_state(f_store, f1, layer, fs);
        double* nz = (double*)malloc(fs * sizeof(double));
        const double* z0 = zeta + layer * fs;
        if (timestep == 1 || timestep == 2) {
#pragma omp parallel for schedule(static)
            for (size_t e = 0; e < fs; ++e) nz[e] = z0[e] + m->dt * f1[e];
        } else {
            const double* f2 = f_store + (layer + 2) * fs;
            const double* f3 = f_store + (layer + 4) * fs;
#pragma omp parallel for schedule(static)
            for (size_t e = 0; e < fs; ++e)
                nz[e] = z0[e] + m->dt * ((23.0 / 12.0) * f1[e] - (16.0 / 12.0) * f2[e] + (5.0 / 12.0) * f3[e]);
        }
        store_new_state(zeta, nz, layer, fs);
        free(nz);
        free(f1);
    }
}

/* ---- spectral stand-in for the two CHOLMOD solves ------------------------------------- */
typedef struct { double re, im; } cplx;

static void fft_inplace(cplx* a, int n, int sign, const cplx* tw) {
    /* iterative radix-2, n power of two; tw[k] = exp(-2 pi i k / n) */
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                cplx w = tw[k * step];
                if (sign > 0) w.im = -w.im;
                const cplx u = a[i + k], x = a[i + k + half];
                const cplx v = {x.re * w.re - x.im * w.im, x.re * w.im + x.im * w.re};
                a[i + k].re = u.re + v.re; a[i + k].im = u.im + v.im;
                a[i + k + half].re = u.re - v.re; a[i + k + half].im = u.im - v.im;
            }
    }
}

static void dft_naive(const cplx* in, cplx* out, int n, int sign, const cplx* tw) {
    for (int k = 0; k < n; ++k) {
        double sr = 0, si = 0;
        int idx = 0;
        for (int t = 0; t < n; ++t) {
            cplx w = tw[idx];
            if (sign > 0) w.im = -w.im;
            sr += in[t].re * w.re - in[t].im * w.im;
            si += in[t].re * w.im + in[t].im * w.re;
            idx += k; if (idx >= n) idx -= n;
        }
        out[k].re = sr; out[k].im = si;
    }
}

static void xform(cplx* a, cplx* tmp, int n, int sign, const cplx* tw) {
    if ((n & (n - 1)) == 0) fft_inplace(a, n, sign, tw);
    else { dft_naive(a, tmp, n, sign, tw); memcpy(a, tmp, n * sizeof(cplx)); }
}

/* cyclic tridiagonal [1 d 1] x = g, d < -2 (Thomas + Sherman-Morrison) */
static void cyclic_solve(double d, double* g, int n, double* cp, double* zz) {
    const double gamma = -d;
    /* modified diagonal: b0 = d - gamma, b_{n-1} = d - 1/gamma, others d */
    double bb0 = d - gamma, bbn = d - 1.0 / gamma;
    /* solve A' x = g and A' z = u (u0 = gamma, u_{n-1} = 1) */
    for (int j = 0; j < n; ++j) zz[j] = 0.0;
    zz[0] = gamma; zz[n - 1] = 1.0;
    double den = bb0;
    cp[0] = 1.0 / den; g[0] /= den; zz[0] /= den;
    for (int j = 1; j < n; ++j) {
        const double bj = (j == n - 1) ? bbn : d;
        den = bj - cp[j - 1];
        cp[j] = 1.0 / den;
        g[j] = (g[j] - g[j - 1]) / den;
        zz[j] = (zz[j] - zz[j - 1]) / den;
    }
    for (int j = n - 2; j >= 0; --j) { g[j] -= cp[j] * g[j + 1]; zz[j] -= cp[j] * zz[j + 1]; }
    const double fact = (g[0] + g[n - 1] / gamma) / (1.0 + zz[0] + zz[n - 1] / gamma);
    for (int j = 0; j < n; ++j) g[j] -= fact * zz[j];
}

/* src/model.jl:172-199 with the solves of src/schemes/laplacian.jl:60-75 done spectrally */
void qgo_evolve_psi(const qgo_params* m, const double* zeta, double* psi) {
    const int M = m->M, P = m->P, W = M + 2;
    const size_t fs = (size_t)(M + 2) * (P + 2);
    const double PI = 3.14159265358979323846;
    cplx* tw = (cplx*)malloc(M * sizeof(cplx));
    for (int k = 0; k < M; ++k) { tw[k].re = cos(2 * PI * k / M); tw[k].im = -sin(2 * PI * k / M); }
    /* Z[j][k] = FFT_x(q~1 + i q~2) */
    cplx* Z = (cplx*)malloc((size_t)M * P * sizeof(cplx));
    const double* q1 = zeta;
    const double* q2 = zeta + fs;
    double tot = 0.0;
    for (int j = 1; j <= P; ++j)
        for (int i = 1; i <= M; ++i)
            if (!(i == 1 && j == 1)) tot += m->Pinv[0] * q1[IDX(i, j)] + m->Pinv[1] * q2[IDX(i, j)];
#pragma omp parallel
    {
        cplx* tmp = (cplx*)malloc(M * sizeof(cplx));
#pragma omp for schedule(static)
        for (int j = 0; j < P; ++j) {
            cplx* row = Z + (size_t)j * M;
            for (int i = 0; i < M; ++i) {
                const double a = q1[IDX(i + 1, j + 1)], b = q2[IDX(i + 1, j + 1)];
                row[i].re = m->Pinv[0] * a + m->Pinv[1] * b;
                row[i].im = m->Pinv[2] * a + m->Pinv[3] * b;
            }
            if (j == 0) row[0].re = -tot;   /* b[1] = 0 and the pinned row: rhs(0,0) := -sum(others) */
            xform(row, tmp, M, -1, tw);
        }
        free(tmp);
    }
    /* per wavenumber: untangle, solve both fields along y, re-tangle */
    const double dx2 = m->dx * m->dx;
#pragma omp parallel
    {
        double* col = (double*)malloc(4 * (size_t)P * sizeof(double));
        double* cp = (double*)malloc(P * sizeof(double));
        double* zz = (double*)malloc(P * sizeof(double));
#pragma omp for schedule(dynamic, 8)
        for (int k = 0; k <= M / 2; ++k) {
            const int km = (M - k) % M;
            double *a1r = col, *a1i = col + P, *a2r = col + 2 * P, *a2i = col + 3 * P;
            for (int j = 0; j < P; ++j) {
                const cplx A = Z[(size_t)j * M + k], B = Z[(size_t)j * M + km];
                a1r[j] = 0.5 * (A.re + B.re) * dx2; a1i[j] = 0.5 * (A.im - B.im) * dx2;
                a2r[j] = 0.5 * (A.im + B.im) * dx2; a2i[j] = 0.5 * (B.re - A.re) * dx2;
            }
            const double lam = 2 * cos(2 * PI * k / M) - 2;
            const double dP = lam - 2.0, dH = lam - 2.0 + m->alpha * dx2;
            if (k == 0) {
                /* singular Poisson line: D[j] = x[j+1]-x[j], D[j]-D[j-1] = g[j], sum D = 0, x[0] = 0 */
                double run = 0, s = 0;
                for (int j = 0; j < P; ++j) { run += a1r[j]; cp[j] = run; s += run; }
                const double dm1 = -s / P;
                double x = 0;
                for (int j = 0; j < P; ++j) { const double D = dm1 + cp[j]; a1r[j] = x; x += D; }
                for (int j = 0; j < P; ++j) a1i[j] = 0.0;
            } else {
                cyclic_solve(dP, a1r, P, cp, zz);
                cyclic_solve(dP, a1i, P, cp, zz);
            }
            cyclic_solve(dH, a2r, P, cp, zz);
            cyclic_solve(dH, a2i, P, cp, zz);
            for (int j = 0; j < P; ++j) {
                cplx* zk = &Z[(size_t)j * M + k];
                zk->re = a1r[j] - a2i[j]; zk->im = a1i[j] + a2r[j];
                if (km != k) {
                    cplx* zm = &Z[(size_t)j * M + km];
                    zm->re = a1r[j] + a2i[j]; zm->im = a2r[j] - a1i[j];
                }
            }
        }
        free(col); free(cp); free(zz);
    }
    /* inverse x transform, gauge, back-projection; history shift */
    double* n1 = newfield(M, P);
    double* n2 = newfield(M, P);
#pragma omp parallel
    {
        cplx* tmp = (cplx*)malloc(M * sizeof(cplx));
#pragma omp for schedule(static)
        for (int j = 0; j < P; ++j) {
            cplx* row = Z + (size_t)j * M;
            xform(row, tmp, M, +1, tw);
        }
        free(tmp);
    }
    const double gauge = Z[0].re / M;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < P; ++j)
        for (int i = 0; i < M; ++i) {
            const double t1 = Z[(size_t)j * M + i].re / M - gauge, t2 = Z[(size_t)j * M + i].im / M;
            n1[IDX(i + 1, j + 1)] = m->Pfwd[0] * t1 + m->Pfwd[1] * t2;
            n2[IDX(i + 1, j + 1)] = m->Pfwd[2] * t1 + m->Pfwd[3] * t2;
        }
    bc(n1, M, P);
    bc(n2, M, P);
    store_new_state(psi, n1, 0, fs);
    store_new_state(psi, n2, 1, fs);
    free(n1); free(n2); free(Z); free(tw);
}

/* loop body of src/run_model_no_output.jl:10-13 */
void qgo_step(const qgo_params* m, double* zeta, double* psi, double* f_store, int first_timestep, int nsteps,
              int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int t = first_timestep; t < first_timestep + nsteps; ++t) {
        qgo_evolve_zeta(m, zeta, psi, t, f_store);
        qgo_evolve_psi(m, zeta, psi);
    }
}

int qgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
