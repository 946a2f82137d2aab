// C ABI of libqgb200 (include/qgb200.h): handle life cycle, spectral plan, host <-> device
// state transfer in the reference's array layout, the step loop, diagnostics.
#include <cudaTypedefs.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "qg_internal.cuh"

namespace qg {

static thread_local std::string g_create_error = "";

static int fail(Handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define QG_CUDA(h, expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            char _b[512];                                                                  \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                  \
            return fail(h, _e == cudaErrorMemoryAllocation ? QG_ERR_NOMEM : QG_ERR_CUDA, _b); \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------------------------------
// host <-> device layout conversion
// ------------------------------------------------------------------------------------------
// Host arrays are (M+2, P+2, 2, 3) column-major per member, level 0 newest, ONE ghost ring.  A host
// field is therefore a dense (P+2) x (M+2) block, and it coincides with the sub-rectangle of the padded
// device field that starts one ghost cell out: host [ih, jh] <-> device row YPAD-1+jh, column XPAD-1+ih.
// Transfers are plain pitched DMA copies between the two (cudaMemcpy3DAsync, the two layers of a
// (level, member) per call) - no staging buffer, no pack / unpack pass over the data.  What remains for
// a kernel is the periodic ghost ring (two cells wide on the device) after an upload.

// Periodic images of the interior into the ghost cells of `nz` consecutive fields
// (update_doubly_periodic_bc!, src/schemes/boundary_conditions.jl:2-13, widened to two cells).
// x_only = 1 (y-slab mode): only the x ghosts of the interior rows; the ghost rows belong to the ring
// neighbours and are filled by the halo exchange.
__global__ void k_fill_ghosts(double* __restrict__ dev, Geom g, int x_only) {
    double* __restrict__ f = dev + (int64_t)blockIdx.y * g.fstride;
    const int W = g.M + 2 * GHOST;
    const int nrow = x_only ? 0 : 2 * GHOST * W;      // cells of the four ghost rows
    const int ncol = 2 * GHOST * g.P;                 // x ghosts of the interior rows
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nrow + ncol; e += gridDim.x * blockDim.x) {
        int ix, iy;
        if (e < nrow) {
            const int rr = e / W;                      // 0..3 -> rows -2, -1, P, P+1
            ix = e - rr * W - GHOST;
            iy = rr < GHOST ? rr - GHOST : g.P + rr - GHOST;
        } else {
            const int c = e - nrow;
            iy = c / (2 * GHOST);
            const int cc = c - iy * 2 * GHOST;         // 0..3 -> columns -2, -1, M, M+1
            ix = cc < GHOST ? cc - GHOST : g.M + cc - GHOST;
        }
        const int i = ((ix % g.M) + g.M) % g.M, j = ((iy % g.P) + g.P) % g.P;
        f[g.at(ix, iy)] = f[g.at(i, j)];
    }
}

__global__ void k_pack(const double* __restrict__ dev, double* __restrict__ host, Geom g, int nm,
                       int slot0, int s1, int s2, int y_from_ghost, int hfields) {
    const int ih = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. M+1 (host index incl. ghost)
    const int jh = blockIdx.y;                              // 0 .. P+1
    if (ih >= g.M + 2) return;
    int z = blockIdx.z;
    const int layer = z & 1;
    z >>= 1;
    const int member = z % nm, level = z / nm;
    const int slot = level == 0 ? slot0 : (level == 1 ? s1 : s2);
    // x wraps locally; y wraps locally too unless this is a y-slab, whose ghost rows hold the
    // neighbours' rows (filled by the halo exchange)
    const int i = ((ih - 1) % g.M + g.M) % g.M;
    const int j = y_from_ghost ? jh - 1 : ((jh - 1) % g.P + g.P) % g.P;
    const int64_t hs = (int64_t)(g.M + 2) * (g.P + 2);
    double v = dev[((int64_t)(slot * nm + member) * 2 + layer) * g.fstride + g.at(i, j)];
    // hfields = fields per member in the host array: 6 (three levels) or 2 (a snapshot of level 1)
    host[(int64_t)member * hfields * hs + (int64_t)(level * 2 + layer) * hs + (int64_t)jh * (g.M + 2) + ih] = v;
}

// ------------------------------------------------------------------------------------------
// initial condition on the device (reference src/model.jl:37-62)
// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (node lo, node hi, field, 0), key = seed: the
// number at a node depends only on (seed, member, layer, i, j), never on the launch shape.
__device__ __forceinline__ double philox_uniform(uint64_t node, uint32_t field, uint64_t seed) {
    uint32_t c0 = (uint32_t)node, c1 = (uint32_t)(node >> 32), c2 = field, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint64_t bits = (((uint64_t)c0 << 32) | c1) >> 11;   // 53 random bits
    return (double)bits * (1.0 / 9007199254740992.0);          // [0, 1)
}

// psi_l = amplitude * rand at every interior node (src/model.jl:41-42); the ghost cells get the
// periodic images (update_doubly_periodic_bc!, :44-45).  One thread per padded device cell.  The draw
// at a node is a pure function of its GLOBAL index (row0 + j) * M + i, so a y-slab rank (row0 = first
// global row it owns, Pglob = global row count) produces exactly its part of the single-GPU field,
// neighbours' rows in its ghost rows included, with no exchange.
__device__ __forceinline__ double ic_psi_at(int i, int jg, int M, uint32_t fz, double amplitude, uint64_t seed) {
    return amplitude * philox_uniform((uint64_t)jg * M + i, fz, seed);
}

__global__ void k_ic_psi(double* __restrict__ psi, Geom g, int nm, double amplitude, uint64_t seed, int row0,
                         int Pglob) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x - GHOST;
    const int iy = blockIdx.y - GHOST;
    if (ix >= g.M + GHOST) return;
    const int fz = blockIdx.z;   // member * 2 + layer
    const int i = ((ix % g.M) + g.M) % g.M, jg = (((row0 + iy) % Pglob) + Pglob) % Pglob;
    psi[(int64_t)fz * g.fstride + g.at(ix, iy)] = ic_psi_at(i, jg, g.M, (uint32_t)fz, amplitude, seed);
}

// q1 = lap(psi1) + S1 (psi2 - psi1), q2 = lap(psi2) + S2 (psi1 - psi2)  (src/model.jl:47-48,
// laplace_5p of src/schemes/laplacian.jl:15-27), ghost images included.  The stencil's psi values are
// re-drawn from the counter-based generator instead of read back, so ghost rows of a y-slab (whose
// own neighbours lie two ranks' rows away) need no halo exchange either.
__global__ void k_ic_q(double* __restrict__ q, Geom g, int nm, double idx2, double S1, double S2,
                       double amplitude, uint64_t seed, int row0, int Pglob) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x - GHOST;
    const int iy = blockIdx.y - GHOST;
    if (ix >= g.M + GHOST) return;
    const int fz = blockIdx.z, layer = fz & 1;
    const int M = g.M;
    const int i = ((ix % M) + M) % M, jg = (((row0 + iy) % Pglob) + Pglob) % Pglob;
    const int iw = i == 0 ? M - 1 : i - 1, ie = i == M - 1 ? 0 : i + 1;
    const int js = jg == 0 ? Pglob - 1 : jg - 1, jn = jg == Pglob - 1 ? 0 : jg + 1;
    const uint32_t fo = (uint32_t)fz, fx = (uint32_t)(fz ^ 1);
    const double c = ic_psi_at(i, jg, M, fo, amplitude, seed);
    const double lap = (ic_psi_at(iw, jg, M, fo, amplitude, seed) + ic_psi_at(ie, jg, M, fo, amplitude, seed) - 4.0 * c +
                        ic_psi_at(i, js, M, fo, amplitude, seed) + ic_psi_at(i, jn, M, fo, amplitude, seed)) * idx2;
    q[(int64_t)fz * g.fstride + g.at(ix, iy)] = lap + (layer == 0 ? S1 : S2) * (ic_psi_at(i, jg, M, fx, amplitude, seed) - c);
}

// ------------------------------------------------------------------------------------------
// diagnostics (DESIGN.md "Diagnostics")
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k5_diag_partial(const double* __restrict__ q, const double* __restrict__ psi, Geom g, double hdx,
                double H1, double H2, double S1, double* __restrict__ part) {
    __shared__ double sh[32];
    const int member = blockIdx.y;
    const double* __restrict__ q1 = q + (int64_t)member * 2 * g.fstride;
    const double* __restrict__ q2 = q1 + g.fstride;
    const double* __restrict__ p1 = psi + (int64_t)member * 2 * g.fstride;
    const double* __restrict__ p2 = p1 + g.fstride;
    double e = 0.0, zz = 0.0;
    for (int j = blockIdx.x; j < g.P; j += gridDim.x) {
        for (int i = threadIdx.x; i < g.M; i += blockDim.x) {
            const int64_t o = g.at(i, j);
            const double ux1 = hdx * (p1[o + 1] - p1[o - 1]), uy1 = hdx * (p1[o + g.pitch] - p1[o - g.pitch]);
            const double ux2 = hdx * (p2[o + 1] - p2[o - 1]), uy2 = hdx * (p2[o + g.pitch] - p2[o - g.pitch]);
            const double dp = p1[o] - p2[o];
            e += H1 * (ux1 * ux1 + uy1 * uy1) + H2 * (ux2 * ux2 + uy2 * uy2) + H1 * S1 * dp * dp;
            zz += H1 * q1[o] * q1[o] + H2 * q2[o] * q2[o];
        }
    }
    const double et = block_sum(e, sh);
    const double zt = block_sum(zz, sh);
    if (threadIdx.x == 0) {
        part[((int64_t)member * gridDim.x + blockIdx.x) * 2 + 0] = et;
        part[((int64_t)member * gridDim.x + blockIdx.x) * 2 + 1] = zt;
    }
}

__global__ void __launch_bounds__(256)
k5_diag_final(const double* __restrict__ part, int nb, double scale, double* __restrict__ out) {
    __shared__ double sh[32];
    const int member = blockIdx.x;
    double e = 0.0, zz = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        e += part[((int64_t)member * nb + b) * 2 + 0];
        zz += part[((int64_t)member * nb + b) * 2 + 1];
    }
    const double et = block_sum(e, sh);
    const double zt = block_sum(zz, sh);
    if (threadIdx.x == 0) {
        out[member * 2 + 0] = scale * et;
        out[member * 2 + 1] = scale * zt;
    }
}

// Running extrema (the reference's update_max / update_min, src/run_model.jl:41-53, which scan a whole
// matrix on the host): max and min over the interior of q_1, q_2, psi_1, psi_2 of the newest level.
// out[member][8] = {max q1, min q1, max q2, min q2, max psi1, min psi1, max psi2, min psi2}; stage 1 leaves
// one row per block with the minima NEGATED so that both stages (and the y-slab all-reduce) are a max.
__device__ __forceinline__ double block_max(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = lane < nw ? sh[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256)
k5_extrema_partial(const double* __restrict__ q, const double* __restrict__ psi, Geom g, double* __restrict__ part) {
    __shared__ double sh[32];
    const int member = blockIdx.y;
    const double* __restrict__ fld[4] = {q + (int64_t)member * 2 * g.fstride, q + (int64_t)(member * 2 + 1) * g.fstride,
                                         psi + (int64_t)member * 2 * g.fstride, psi + (int64_t)(member * 2 + 1) * g.fstride};
    double m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int j = blockIdx.x; j < g.P; j += gridDim.x)
        for (int i = threadIdx.x; i < g.M; i += blockDim.x) {
            const int64_t o = g.at(i, j);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double v = fld[k][o];
                m[2 * k] = fmax(m[2 * k], v);
                m[2 * k + 1] = fmax(m[2 * k + 1], -v);
            }
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double r = block_max(m[k], sh);
        if (threadIdx.x == 0) part[((int64_t)member * gridDim.x + blockIdx.x) * 8 + k] = r;
    }
}

__global__ void __launch_bounds__(256) k5_extrema_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    __shared__ double sh[32];
    const int member = blockIdx.x;
    for (int k = 0; k < 8; ++k) {
        double m = -INFINITY;
        for (int b = threadIdx.x; b < nb; b += blockDim.x) m = fmax(m, part[((int64_t)member * nb + b) * 8 + k]);
        m = block_max(m, sh);
        if (threadIdx.x == 0) out[member * 8 + k] = m;   // minima still negated
    }
}

// ------------------------------------------------------------------------------------------
// spectral plan
// ------------------------------------------------------------------------------------------
static int ilog2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return (1 << l) == n ? l : -1;
}

static void exact_twiddle(int n, int N, long double* c, long double* s) {
    // exp(-2 pi i n / N) evaluated in 80-bit extended precision (argument error ~3e-19, far
    // below half a double ulp); multiples of a quarter turn are forced to their exact values.
    const long double PI = 3.14159265358979323846264338327950288L;
    if (n == 0) { *c = 1.0L; *s = 0.0L; return; }
    if (4LL * n == N) { *c = 0.0L; *s = -1.0L; return; }
    if (2LL * n == N) { *c = -1.0L; *s = 0.0L; return; }
    if (4LL * n == 3LL * N) { *c = 0.0L; *s = 1.0L; return; }
    const long double th = 2.0L * PI * (long double)n / (long double)N;
    *c = cosl(th);
    *s = -sinl(th);
}

template <typename T>
static cudaError_t upload_vec(T** dptr, const std::vector<T>& v) {
    cudaError_t e = cudaMalloc((void**)dptr, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// ---- pure host arithmetic of the plan (also reachable without a GPU through qg_plan_probe) ----------
// Real column `col` of the packed spectral layout (qg_internal.cuh, Plan): which modal field (0 = Poisson /
// barotropic, 1 = Helmholtz / baroclinic), which x wavenumber, real or imaginary part.
static void column_mode(int col, int M, int* field, int* k, bool* is_re) {
    const int s = col >> 1, part = col & 1;
    if (s == 0) { *field = part; *k = 0; }
    else if ((M % 2 == 0) && s == M / 2) { *field = part; *k = M / 2; }
    else if (2 * s < M) { *field = 0; *k = s; }
    else { *field = 1; *k = M - s; }
    *is_re = (s == 0 || ((M % 2 == 0) && s == M / 2)) ? true : (part == 0);
}

// Root inside the unit circle of r^2 + d r + 1 = 0, d = -(2 + e), e = 4 sin^2(pi k / M) [- alpha dx^2]: the
// decay per row of the two first-order y-recurrences (k3_ysolve.cu).  e <= 0 (the Poisson k = 0 column): singular.
static long double column_root(int field, int k, int M, long double dx2, double alpha, bool* singular) {
    const long double PI = 3.14159265358979323846264338327950288L;
    const long double sn = sinl(PI * k / M);
    long double e = 4.0L * sn * sn;
    if (field == 1) e -= (long double)alpha * dx2;
    *singular = !(e > 0.0L);
    return *singular ? 0.0L : 2.0L / ((2.0L + e) + sqrtl(e * (e + 4.0L)));
}

// y-slab mode: the (32-column tile, 32-row segment) pairs in which the carry terms added by k3_rank_correct are
// still above 2^-60 of their value at the slab edge: rows closer than n_cut = 41.6 / -ln r to the bottom or the
// top edge, n_cut taken as the maximum over the tile's columns.
static void edge_worklist(const std::vector<double>& rtab, const std::vector<double>& kap,
                          const std::vector<double>& logr, int ncol, int P, std::vector<int2>* work) {
    const int nseg = P / 32;
    for (int t = 0; t < (ncol + 31) / 32; ++t) {
        int ncut = 0;
        for (int col = 32 * t; col < 32 * t + 32 && col < ncol; ++col) {
            if (!(rtab[col] > 0.0) || kap[col] == 0.0) continue;
            const double n = -41.6 / logr[col];
            const int c = n >= (double)P ? P : (int)n + 1;
            if (c > ncut) ncut = c;
        }
        for (int sg = 0; sg < nseg; ++sg) {
            const int i0 = 32 * sg;
            if (i0 < ncut || P - (i0 + 31) <= ncut) work->push_back(make_int2(t, sg));
        }
    }
}

cudaError_t build_plan(Handle* h) {
    Plan& pl = h->plan;
    memset(&pl, 0, sizeof(pl));
    const int M = h->g.M, P = h->g.P;
    pl.M = M;
    pl.P = P;
    pl.ncol = 2 * M;
    pl.log2M = ilog2_exact(M);
    pl.pow2 = (pl.log2M >= 3 && M <= 16384) ? 1 : 0;
    if (pl.pow2) {
        pl.tpr = M / 8;
        pl.rpb = pl.tpr >= 128 ? 1 : 128 / pl.tpr;
    }
    pl.C = (P + 31) / 32;
    pl.lenLast = P - 32 * (pl.C - 1);
    pl.wpc = pl.C < 16 ? pl.C : 16;
    {
        int need = (pl.C + pl.wpc - 1) / pl.wpc;
        int cs = 1;
        while (cs < need && cs < 8) cs *= 2;
        pl.CS = cs;
        pl.m = (pl.C + cs * pl.wpc - 1) / (cs * pl.wpc);
    }
    pl.k0scale = h->prm.dx * h->prm.dx / M;
    {   // TMA-staged y-solve: at most 16 chunks (512 rows, 64 KB) per CTA, cluster of <= 8
        int cs = 1;
        while (cs < 8 && (pl.C + cs - 1) / cs > 16) cs *= 2;
        pl.ts_CS = cs;
        pl.ts_nchunk = (pl.C + cs - 1) / cs;
        pl.ts_ok = (pl.ts_nchunk <= 16 && (pl.ncol % 16) == 0) ? 1 : 0;
    }

    std::vector<double2> tw(M);
    for (int n = 0; n < M; ++n) {
        long double c, s;
        exact_twiddle(n, M, &c, &s);
        tw[n] = make_double2((double)c, (double)s);
    }
    const long double dx2 = (long double)h->prm.dx * (long double)h->prm.dx;
    std::vector<double> rtab(pl.ncol), kap(pl.ncol), rho32(pl.ncol), h32(pl.ncol), rhoL(pl.ncol),
        hL(pl.ncol), inv1(pl.ncol), pinw(pl.ncol), gw(pl.ncol), logr(pl.ncol), g1mr2(pl.ncol);
    std::vector<long double> rlong(pl.ncol);
    for (int col = 0; col < pl.ncol; ++col) {
        int field, k;
        bool is_re, singular;
        column_mode(col, M, &field, &k, &is_re);
        const long double r = column_root(field, k, M, dx2, h->prm.alpha, &singular);
        const long double r32 = powl(r, 32), rL = powl(r, pl.lenLast), rP = powl(r, h->Pglob > 0 ? h->Pglob : P);
        const long double geo32 = singular ? 0.0L : r * (1.0L - r32 * r32) / (1.0L - r * r);
        const long double geoL = singular ? 0.0L : r * (1.0L - rL * rL) / (1.0L - r * r);
        rtab[col] = (double)r;
        rlong[col] = r;
        kap[col] = singular ? 0.0 : (double)(-r * dx2 / M);
        rho32[col] = (double)r32;
        rhoL[col] = (double)rL;
        h32[col] = (double)geo32;
        hL[col] = (double)geoL;
        inv1[col] = singular ? 1.0 : (double)(1.0L / (1.0L - rP));
        logr[col] = singular ? 0.0 : (double)logl(r);
        g1mr2[col] = singular ? 0.0 : (double)(1.0L / (1.0L - r * r));
        pinw[col] = (field == 0 && is_re) ? 1.0 : 0.0;
        // weight of this column in psi~1(0,0) = sum_k U1[k] over all M wavenumbers (Hermitian pairs count twice)
        gw[col] = (field == 0 && is_re) ? ((k == 0 || 2 * k == M) ? 1.0 : 2.0) : 0.0;
    }
    cudaError_t e;
    if ((e = upload_vec(&pl.tw, tw)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.rtab, rtab)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.kap, kap)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.rho32, rho32)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.h32, h32)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.rhoL, rhoL)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.hL, hL)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.inv1mrP, inv1)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.pinw, pinw)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.gw, gw)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.logr, logr)) != cudaSuccess) return e;
    if ((e = upload_vec(&pl.g1mr2, g1mr2)) != cudaSuccess) return e;
    if (h->dist_n > 1 && P % 32 == 0) {   // y-slab mode: the work list of k3_rank_correct
        std::vector<int2> work;
        edge_worklist(rtab, kap, logr, pl.ncol, P, &work);
        pl.ncorr = (int)work.size();
        if (pl.ncorr > 0 && (e = upload_vec(&pl.corr_work, work)) != cudaSuccess) return e;
    }
    pl.ngp = (pl.ncol + 15) / 16;
    // persistent y-solve (k3_ysolve_pipe): needs whole 32-row chunks; one TMA box of <= 256 rows
    // (or two equal ones) per CTA tile, and all per-column constants in one table so that they
    // arrive with the tile:  rows 0-31 cA[i] = r^(i+1) * sum_{m<32-i} r^(2m), rows 32-63
    // cB[i] = r^(32-i), then r, kap, r^32, h32, 1/(1-r^P), pin weight, gauge weight, spare.
    {
        int cs = 1;
        while (cs < 16 && (pl.C + cs - 1) / cs > 16) cs *= 2;
        pl.tp_CS = cs;
        pl.tp_nchunk = (pl.C + cs - 1) / cs;
    }
    pl.tp_ok = (pl.tp_nchunk <= 16 && (pl.ncol % 16) == 0 && P % 32 == 0) ? 1 : 0;
    pl.tp_boxrows = pl.tp_nchunk * 32 <= 256 ? pl.tp_nchunk * 32 : pl.tp_nchunk * 16;
    if (pl.tp_ok) {
        const int NR = 72;
        std::vector<double> ct((size_t)NR * pl.ncol, 0.0);
        for (int col = 0; col < pl.ncol; ++col) {
            const long double r = (long double)rtab[col] == 0.0L ? 0.0L : rlong[col];
            long double rp[34];
            rp[0] = 1.0L;
            for (int i = 1; i < 34; ++i) rp[i] = rp[i - 1] * r;
            long double gs[34];   // gs[n] = sum_{m<n} r^(2m)
            gs[0] = 0.0L;
            for (int n = 1; n < 34; ++n) gs[n] = gs[n - 1] + rp[n - 1] * rp[n - 1];
            for (int i = 0; i < 32; ++i) {
                ct[(size_t)i * pl.ncol + col] = (double)(rp[i + 1] * gs[32 - i]);
                ct[(size_t)(32 + i) * pl.ncol + col] = (double)rp[32 - i];
            }
            ct[(size_t)64 * pl.ncol + col] = rtab[col];
            ct[(size_t)65 * pl.ncol + col] = kap[col];
            ct[(size_t)66 * pl.ncol + col] = rho32[col];
            ct[(size_t)67 * pl.ncol + col] = h32[col];
            ct[(size_t)68 * pl.ncol + col] = inv1[col];
            ct[(size_t)69 * pl.ncol + col] = pinw[col];
            ct[(size_t)70 * pl.ncol + col] = gw[col];
        }
        if ((e = upload_vec(&pl.coltab, ct)) != cudaSuccess) return e;
    }
    h->plan_ok = true;
    return cudaSuccess;
}

void free_plan(Handle* h) {
    Plan& pl = h->plan;
    cudaFree(pl.tw); cudaFree(pl.rtab); cudaFree(pl.kap); cudaFree(pl.rho32); cudaFree(pl.h32);
    cudaFree(pl.rhoL); cudaFree(pl.hL); cudaFree(pl.inv1mrP); cudaFree(pl.pinw); cudaFree(pl.gw);
    cudaFree(pl.coltab); cudaFree(pl.logr); cudaFree(pl.g1mr2); cudaFree(pl.corr_work);
    memset(&pl, 0, sizeof(pl));
    h->plan_ok = false;
}

static int make_tensor_map(Handle* h, CUtensorMap* tm, double* base, int bx, int by) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(h, QG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)h->g.pitch, (cuuint64_t)h->g.rows, (cuuint64_t)h->nfields};
    const cuuint64_t strides[2] = {(cuuint64_t)h->g.pitch * sizeof(double),
                                   (cuuint64_t)h->g.fstride * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(h, QG_ERR_CUDA, b);
    }
    return QG_OK;
}

static int make_tensor_map_S(Handle* h) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(h, QG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const cuuint64_t dims[2] = {(cuuint64_t)h->plan.ncol, (cuuint64_t)h->plan.P * h->nm};
    const cuuint64_t strides[1] = {(cuuint64_t)h->plan.ncol * sizeof(double)};
    const cuuint32_t box[2] = {16, 32};   // TS_WC columns x one chunk (k3_ysolve.cu)
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&h->tm_S, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->S, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (spectral) failed with CUresult %d", (int)r);
        return fail(h, QG_ERR_CUDA, b);
    }
    if (h->plan.tp_ok) {   // persistent y-solve: big tile boxes and the column table
        const cuuint32_t box2[2] = {16, (cuuint32_t)h->plan.tp_boxrows};
        r = encode(&h->tm_S2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->S, dims, strides, box2, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            const cuuint64_t dimsT[2] = {(cuuint64_t)h->plan.ncol, 72};
            const cuuint32_t boxT[2] = {16, 72};
            r = encode(&h->tm_T, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->plan.coltab, dimsT, strides, boxT, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) {
            char b[128];
            snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (persistent y-solve) failed with CUresult %d", (int)r);
            return fail(h, QG_ERR_CUDA, b);
        }
    }
    return QG_OK;
}

cudaError_t launch_diag(Handle* h) {
    const int nb = h->diag_blocks;
    const double inv = 1.0 / h->prm.dx;
    {
        KernelTimer t(h, QG_K_DIAG);
        k5_diag_partial<<<dim3(nb, h->nm), 256, 0, h->stream>>>(
            h->field(h->q, h->qcur, 0, 0), h->field(h->psi, h->pcur, 0, 0), h->g, 0.5 * inv, h->prm.H1,
            h->prm.H2, h->prm.S1, h->diag_part);
    }
    {
        KernelTimer t(h, QG_K_DIAG);
        k5_diag_final<<<h->nm, 256, 0, h->stream>>>(h->diag_part, nb, 0.5 * h->prm.dx * h->prm.dx,
                                                   h->diag_part + (int64_t)h->nm * nb * 2);
    }
    return cudaGetLastError();
}

static int do_evolve_psi(Handle* h) {
    const int nxt = (h->pcur + 1) % 3;
    QG_CUDA(h, launch_fft_forward(h, h->field(h->q, h->qcur, 0, 0), 0));
    QG_CUDA(h, launch_ysolve(h, 1, 0));
    QG_CUDA(h, launch_fft_inverse(h, h->field(h->psi, nxt, 0, 0), 1));
    h->pcur = nxt;
    // y-slab: psi's ghost rows come from the ring neighbours - NCCL send/recv, or (peer mode) K4 has
    // stored them itself and the barrier orders those stores (and K1's q rows) before the next step
    if (h->dist_n > 1) QG_CUDA(h, h->peer_ok ? dist_barrier(h) : dist_halo_exchange(h, h->psi, nxt));
    return QG_OK;
}

static int do_evolve_zeta(Handle* h, int timestep) {
    // peer mode: K1 stores its edge rows into the neighbours' ghost rows; they are read by the next
    // step's K1 only, after the barriers of evolve_psi (two evolve_zeta calls in a row - not a
    // reference pattern - get a barrier of their own)
    if (h->peer_ok && h->q_halo_pending) QG_CUDA(h, dist_barrier(h));
    QG_CUDA(h, launch_zeta(h, timestep));
    if (h->dist_n > 1 && !h->peer_ok) QG_CUDA(h, dist_halo_exchange(h, h->q, h->qcur));
    h->q_halo_pending = h->peer_ok;
    return QG_OK;
}

}  // namespace qg

using namespace qg;

extern "C" {

int qg_abi_version(void) { return QG_ABI_VERSION; }

const char* qg_last_error(const qg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

const char* qg_kernel_name(int id) {
    static const char* names[QG_NKERNELS] = {"k1_zeta_step", "k2_fft_forward", "k3_pre", "k3_ysolve",
                                             "k3_gauge", "k4_fft_inverse", "k5_diag", "k_pack"};
    return (id >= 0 && id < QG_NKERNELS) ? names[id] : "?";
}

int qg_create(const qg_params* p, int device, int nmembers, void* stream, qg_handle** out) {
    if (!p || !out) return fail(nullptr, QG_ERR_INVALID, "qg_create: null argument");
    *out = nullptr;
    if (p->M < 3 || p->P < 3) return fail(nullptr, QG_ERR_INVALID, "qg_create: M and P must be >= 3");
    if (nmembers < 1 || nmembers > 16384) return fail(nullptr, QG_ERR_INVALID, "qg_create: bad nmembers");
    if (!(p->dx > 0.0)) return fail(nullptr, QG_ERR_INVALID, "qg_create: dx must be positive");
    if (!(p->alpha < 0.0))
        return fail(nullptr, QG_ERR_INVALID, "qg_create: alpha (S_eig) must be negative (modified Helmholtz)");
    if (p->P > 16384) return fail(nullptr, QG_ERR_INVALID, "qg_create: P > 16384 not supported");
    const bool pow2 = (p->M & (p->M - 1)) == 0 && p->M >= 8;
    if (pow2 && p->M > 16384) return fail(nullptr, QG_ERR_INVALID, "qg_create: M > 16384 not supported");
    if (!pow2 && p->M > 4096)
        return fail(nullptr, QG_ERR_INVALID, "qg_create: non-power-of-two M > 4096 not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(nullptr, QG_ERR_NODEVICE, "qg_create: no CUDA device (there is no CPU fallback)");
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return fail(nullptr, QG_ERR_NODEVICE, "qg_create: device is not sm_100 class (there is no CPU fallback)");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, QG_ERR_CUDA, "cudaSetDevice failed");

    qg_handle* h = new (std::nothrow) qg_handle();
    if (!h) return fail(nullptr, QG_ERR_NOMEM, "qg_create: out of host memory");
    h->prm = *p;
    h->device = device;
    h->nm = nmembers;
    h->g.M = p->M;
    h->g.P = p->P;
    h->g.pitch = ((XPAD + p->M + GHOST + 15) / 16) * 16;
    h->g.rows = p->P + 2 * YPAD;
    h->g.fstride = (int64_t)h->g.pitch * h->g.rows;
    h->nfields = 3 * nmembers * 2;
    int rc = QG_OK;
    auto bail = [&](int code) { std::string m = h->err; qg_destroy(h); g_create_error = m; return code; };
#define QG_TRY(expr)                                                        \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) {                                            \
            h->err = std::string(#expr) + ": " + cudaGetErrorString(_e);    \
            return bail(_e == cudaErrorMemoryAllocation ? QG_ERR_NOMEM : QG_ERR_CUDA); \
        }                                                                   \
    } while (0)
    if (stream) {
        h->stream = (cudaStream_t)stream;
    } else {
        QG_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    const size_t fbytes = (size_t)h->nfields * h->g.fstride * sizeof(double);
    QG_TRY(cudaMalloc((void**)&h->q, fbytes));
    QG_TRY(cudaMalloc((void**)&h->psi, fbytes));
    QG_TRY(cudaMalloc((void**)&h->f, fbytes));
    QG_TRY(cudaMemsetAsync(h->q, 0, fbytes, h->stream));
    QG_TRY(cudaMemsetAsync(h->psi, 0, fbytes, h->stream));
    QG_TRY(cudaMemsetAsync(h->f, 0, fbytes, h->stream));
    QG_TRY(cudaMalloc((void**)&h->S, (size_t)nmembers * p->P * 2 * p->M * sizeof(double)));
    QG_TRY(cudaMalloc((void**)&h->k0sol, (size_t)nmembers * p->P * sizeof(double)));
    QG_TRY(cudaMalloc((void**)&h->col0, (size_t)nmembers * p->P * sizeof(double)));
    QG_TRY(cudaMalloc((void**)&h->gpart, (size_t)nmembers * ((2 * p->M + 15) / 16) * sizeof(double)));
    h->Pglob = p->P;
    QG_TRY(cudaMalloc((void**)&h->scal, (size_t)nmembers * 4 * sizeof(double)));
    QG_TRY(cudaMemsetAsync(h->scal, 0, (size_t)nmembers * 4 * sizeof(double), h->stream));
    h->diag_blocks = p->P < 592 ? p->P : 592;
    QG_TRY(cudaMalloc((void**)&h->diag_part, ((size_t)nmembers * h->diag_blocks * 2 + 2 * nmembers) * sizeof(double)));
    QG_TRY(build_plan(h));
    if (getenv("QG_NO_GRAPH")) h->use_graph = false;
    {
        const char* ty = getenv("QG_K1_TY");
        const int v = ty ? atoi(ty) : (p->P >= 512 ? 24 : 16);   // measured on B200: 24 rows is best at 4096^2
        h->k1_ty = (v == 8 || v == 12 || v == 24) ? v : 16;
    }
    rc = make_tensor_map(h, &h->tm_q, h->q, K1_TX + 2 * GHOST, h->k1_ty + 2);
    if (rc == QG_OK) rc = make_tensor_map(h, &h->tm_psi, h->psi, K1_TX + 2 * GHOST, h->k1_ty + 2 * GHOST);
    if (rc == QG_OK && (h->plan.ts_ok || h->plan.tp_ok)) rc = make_tensor_map_S(h);
    if (rc != QG_OK) return bail(rc);
    QG_TRY(cudaStreamSynchronize(h->stream));
#undef QG_TRY
    *out = h;
    return QG_OK;
}

int qg_destroy(qg_handle* h) {
    if (!h) return QG_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    dist_destroy(h);
    free_plan(h);
    cudaFree(h->col0);
    cudaFree(h->gpart);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); cudaEventDestroy(h->snap_ev); }
    cudaFree(h->snap_stage);
    cudaFree(h->q); cudaFree(h->psi); cudaFree(h->f); cudaFree(h->S); cudaFree(h->k0sol);
    cudaFree(h->scal); cudaFree(h->solve_tmp); cudaFree(h->diag_part); cudaFree(h->ext_part);
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            if (h->gexec[a][b]) cudaGraphExecDestroy(h->gexec[a][b]);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return QG_OK;
}

// One pitched DMA copy: the two layers of (slot, member) of a device array <-> the two layers of
// (level, member) of a host array in the reference layout.  hfields = fields per member on the host
// (6: three levels; 2: level 1 only).
static cudaError_t copy_fields(qg_handle* h, double* dev, int slot, double* host, int level, int member, int hfields,
                               bool to_device, cudaStream_t st) {
    const Geom& g = h->g;
    const size_t hs = (size_t)(g.M + 2) * (g.P + 2);
    cudaMemcpy3DParms p{};
    double* d = h->field(dev, slot, member, 0) + (int64_t)(YPAD - 1) * g.pitch + (XPAD - 1);
    double* hp = host + ((size_t)member * hfields + (size_t)level * 2) * hs;
    const cudaPitchedPtr dptr = make_cudaPitchedPtr(d, (size_t)g.pitch * sizeof(double), (size_t)g.pitch, (size_t)g.rows);
    const cudaPitchedPtr hptr = make_cudaPitchedPtr(hp, (size_t)(g.M + 2) * sizeof(double), (size_t)(g.M + 2), (size_t)(g.P + 2));
    p.srcPtr = to_device ? hptr : dptr;
    p.dstPtr = to_device ? dptr : hptr;
    p.extent = make_cudaExtent((size_t)(g.M + 2) * sizeof(double), (size_t)(g.P + 2), 2);
    p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    return cudaMemcpy3DAsync(&p, st);
}

static cudaError_t fill_ghosts(qg_handle* h, double* first_field, int nz) {
    const int work = 2 * GHOST * (h->g.M + 2 * GHOST + h->g.P);
    const int nb = (work + 255) / 256 < 64 ? (work + 255) / 256 : 64;
    KernelTimer t(h, QG_K_PACK);
    k_fill_ghosts<<<dim3(nb, nz), 256, 0, h->stream>>>(first_field, h->g, h->dist_n > 1 ? 1 : 0);
    return cudaGetLastError();
}

// all three levels of one array; level 0 goes to slot `cur`
static int upload_one(qg_handle* h, const double* host, double* dev, int cur) {
    const int slot_of[3] = {cur, (cur + 2) % 3, (cur + 1) % 3};
    for (int level = 0; level < 3; ++level)
        for (int m = 0; m < h->nm; ++m)
            QG_CUDA(h, copy_fields(h, dev, slot_of[level], const_cast<double*>(host), level, m, 6, true, h->stream));
    QG_CUDA(h, fill_ghosts(h, dev, h->nfields));
    if (h->dist_n > 1)   // y-slab: the ghost rows are the neighbours' rows, not local periodic images
        for (int s = 0; s < 3; ++s) QG_CUDA(h, dist_halo_exchange(h, dev, s));
    return QG_OK;
}

static int ensure_copy_stream(qg_handle* h) {
    if (!h->copy_stream) {
        QG_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        QG_CUDA(h, cudaEventCreateWithFlags(&h->snap_ev, cudaEventDisableTiming));
    }
    return QG_OK;
}

// needs_ghosts: the array's ghost ring is not maintained by the step kernels (f_store).
// (Splitting the copies over two streams / DMA queues was measured and does not help: one queue
// already saturates the PCIe link, 47 GB/s on the bench boxes.)
static int download_one(qg_handle* h, double* dev, double* host, int cur, bool needs_ghosts) {
    const int slot_of[3] = {cur, (cur + 2) % 3, (cur + 1) % 3};
    if (needs_ghosts) QG_CUDA(h, fill_ghosts(h, dev, h->nfields));
    if (h->dist_n > 1)
        for (int s = 0; s < 3; ++s) QG_CUDA(h, dist_halo_exchange(h, dev, s));
    for (int level = 0; level < 3; ++level)
        for (int m = 0; m < h->nm; ++m)
            QG_CUDA(h, copy_fields(h, dev, slot_of[level], host, level, m, 6, false, h->stream));
    return QG_OK;
}

int qg_upload_state(qg_handle* h, const double* zeta, const double* psi, const double* f_store) {
    if (!h) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if (zeta) {
        // zeta and f_store share one rotation counter (both are pushed by evolve_zeta!).  Uploaded
        // together, or first, they start at slot 0; zeta alone on a handle that has stepped goes into
        // the slots f_store's history currently occupies, so the next AB3 step still pairs them up.
        if (f_store || !h->have_state) h->qcur = 0;
        if ((rc = upload_one(h, zeta, h->q, h->qcur))) return rc;
        if (!f_store && !h->have_state)
            QG_CUDA(h, cudaMemsetAsync(h->f, 0, (size_t)h->nfields * h->g.fstride * sizeof(double), h->stream));
    }
    if (f_store) {
        if ((rc = upload_one(h, f_store, h->f, h->qcur))) return rc;
    }
    if (psi) {
        h->pcur = 0;
        if ((rc = upload_one(h, psi, h->psi, 0))) return rc;
    }
    h->q_halo_pending = false;
    QG_CUDA(h, cudaStreamSynchronize(h->stream));   // host buffers are borrowed only for the call
    h->have_state = true;
    return QG_OK;
}

// level 1 of one array only (2 fields per member), the other levels zeroed
static int upload_level1(qg_handle* h, const double* host, double* dev) {
    QG_CUDA(h, cudaMemsetAsync(dev, 0, (size_t)h->nfields * h->g.fstride * sizeof(double), h->stream));
    for (int m = 0; m < h->nm; ++m)
        QG_CUDA(h, copy_fields(h, dev, 0, const_cast<double*>(host), 0, m, 6, true, h->stream));
    QG_CUDA(h, fill_ghosts(h, h->field(dev, 0, 0, 0), h->nm * 2));
    if (h->dist_n > 1) QG_CUDA(h, dist_halo_exchange(h, dev, 0));
    return QG_OK;
}

int qg_upload_initial_state(qg_handle* h, const double* zeta, const double* psi) {
    if (!h || !zeta || !psi) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc;
    h->qcur = 0;
    h->pcur = 0;
    h->q_halo_pending = false;
    if ((rc = upload_level1(h, zeta, h->q))) return rc;
    if ((rc = upload_level1(h, psi, h->psi))) return rc;
    QG_CUDA(h, cudaMemsetAsync(h->f, 0, (size_t)h->nfields * h->g.fstride * sizeof(double), h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_state = true;
    return QG_OK;
}

int qg_init_state(qg_handle* h, uint64_t seed, double amplitude, double S1, double S2) {
    if (!h) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->nfields * h->g.fstride * sizeof(double);
    QG_CUDA(h, cudaMemsetAsync(h->q, 0, bytes, h->stream));
    QG_CUDA(h, cudaMemsetAsync(h->psi, 0, bytes, h->stream));
    QG_CUDA(h, cudaMemsetAsync(h->f, 0, bytes, h->stream));
    h->qcur = 0;
    h->pcur = 0;
    h->q_halo_pending = false;
    dim3 block(128), grid((h->g.M + 2 * GHOST + 127) / 128, h->g.P + 2 * GHOST, h->nm * 2);
    const double inv = 1.0 / h->prm.dx;
    // y-slab mode: this rank draws its rows of the global field (and its neighbours' rows into the ghosts)
    const int row0 = h->dist_n > 1 ? h->dist_rank * h->g.P : 0;
    {
        KernelTimer t(h, QG_K_PACK);
        k_ic_psi<<<grid, block, 0, h->stream>>>(h->field(h->psi, 0, 0, 0), h->g, h->nm, amplitude, seed, row0, h->Pglob);
    }
    QG_CUDA(h, cudaGetLastError());
    {
        KernelTimer t(h, QG_K_PACK);
        k_ic_q<<<grid, block, 0, h->stream>>>(h->field(h->q, 0, 0, 0), h->g, h->nm, inv * inv, S1, S2, amplitude, seed,
                                              row0, h->Pglob);
    }
    QG_CUDA(h, cudaGetLastError());
    // y-slab peer mode: no rank may start stepping (and storing edge rows into its neighbours' arrays) before
    // every rank's memsets above have run
    if (h->peer_ok) QG_CUDA(h, dist_barrier(h));
    h->have_state = true;
    return QG_OK;
}

int qg_download_state(qg_handle* h, double* zeta, double* psi, double* f_store) {
    if (!h) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_download_state: no state uploaded");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc;
    // q and psi carry their periodic ghost ring on the device (written by K1 / K4 with every step);
    // f_store's is produced here.  The copies run back to back on the handle's stream.
    if (zeta && (rc = download_one(h, h->q, zeta, h->qcur, false))) return rc;
    if (psi && (rc = download_one(h, h->psi, psi, h->pcur, false))) return rc;
    if (f_store && (rc = download_one(h, h->f, f_store, h->qcur, true))) return rc;
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return dist_poll_error(h);
}

// Snapshot of the newest level (src/run_model.jl:70-73, 86-90 write exactly zeta[:,:,:,1] and
// psi[:,:,:,1]): packed on the handle's stream into a staging buffer of its own, copied to the
// host on a separate stream, so the steps queued after qg_snapshot_begin overlap the PCIe copy.
int qg_snapshot_end(qg_handle* h) {
    if (!h) return QG_ERR_INVALID;
    if (!h->snap_pending) return QG_OK;
    QG_CUDA(h, cudaSetDevice(h->device));
    QG_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    h->snap_pending = false;
    return QG_OK;
}

int qg_snapshot_begin(qg_handle* h, double* zeta1, double* psi1) {
    if (!h || (!zeta1 && !psi1)) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_snapshot_begin: no state uploaded");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = qg_snapshot_end(h);   // the staging buffer is single: finish the previous snapshot first
    if (rc) return rc;
    const size_t n = (size_t)h->nm * 2 * (h->g.M + 2) * (h->g.P + 2);
    if (!h->snap_stage) QG_CUDA(h, cudaMalloc((void**)&h->snap_stage, 2 * n * sizeof(double)));
    if ((rc = ensure_copy_stream(h))) return rc;
    dim3 block(128), grid((h->g.M + 2 + 127) / 128, h->g.P + 2, h->nm * 2);
    for (int which = 0; which < 2; ++which) {
        double* host = which == 0 ? zeta1 : psi1;
        if (!host) continue;
        double* dev = which == 0 ? h->q : h->psi;
        const int cur = which == 0 ? h->qcur : h->pcur;
        if (h->dist_n > 1) QG_CUDA(h, dist_halo_exchange(h, dev, cur));
        {
            KernelTimer t(h, QG_K_PACK);
            k_pack<<<grid, block, 0, h->stream>>>(dev, h->snap_stage + which * n, h->g, h->nm, cur, cur, cur,
                                                  h->dist_n > 1 ? 1 : 0, 2);
        }
        QG_CUDA(h, cudaGetLastError());
    }
    QG_CUDA(h, cudaEventRecord(h->snap_ev, h->stream));
    QG_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->snap_ev, 0));
    if (zeta1)
        QG_CUDA(h, cudaMemcpyAsync(zeta1, h->snap_stage, n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
    if (psi1)
        QG_CUDA(h, cudaMemcpyAsync(psi1, h->snap_stage + n, n * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
    h->snap_pending = true;
    return QG_OK;
}

int qg_evolve_zeta(qg_handle* h, int timestep) {
    if (!h) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_evolve_zeta: no state uploaded");
    if (timestep < 1) return fail(h, QG_ERR_INVALID, "qg_evolve_zeta: timestep is 1-based");
    QG_CUDA(h, cudaSetDevice(h->device));
    return do_evolve_zeta(h, timestep);
}

int qg_evolve_psi(qg_handle* h) {
    if (!h) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_evolve_psi: no state uploaded");
    QG_CUDA(h, cudaSetDevice(h->device));
    return do_evolve_psi(h);
}

// Replay (capturing on first use) the 3-step AB3 cycle that starts from the current slot phase.
static int step_cycle_graph(qg_handle* h, int t) {
    const int a = h->qcur, b = h->pcur;
    if (!h->gexec[a][b]) {
        const int64_t l0 = h->launches;
        int64_t k0[QG_NKERNELS];
        for (int i = 0; i < QG_NKERNELS; ++i) k0[i] = h->kcount[i];
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            h->use_graph = false;
            return 1;
        }
        int rc = QG_OK;
        for (int i = 0; i < 3 && rc == QG_OK; ++i) {
            rc = do_evolve_zeta(h, t + i);
            if (rc == QG_OK) rc = do_evolve_psi(h);
        }
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (rc != QG_OK || e != cudaSuccess || !graph ||
            cudaGraphInstantiate(&h->gexec[a][b], graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            h->gexec[a][b] = nullptr;
            h->use_graph = false;
            h->qcur = a;
            h->pcur = b;
            h->launches = l0;
            for (int i = 0; i < QG_NKERNELS; ++i) h->kcount[i] = k0[i];
            return 1;   // caller falls back to plain launches
        }
        cudaGraphDestroy(graph);
        h->glaunches[a][b] = h->launches - l0;
        for (int i = 0; i < QG_NKERNELS; ++i) h->gkcount[a][b][i] = h->kcount[i] - k0[i];
        // capture only recorded the work: undo the bookkeeping, the launch below redoes it
        h->launches = l0;
        for (int i = 0; i < QG_NKERNELS; ++i) h->kcount[i] = k0[i];
    }
    QG_CUDA(h, cudaGraphLaunch(h->gexec[a][b], h->stream));
    h->launches += h->glaunches[a][b];
    for (int i = 0; i < QG_NKERNELS; ++i) h->kcount[i] += h->gkcount[a][b][i];
    // three steps rotate every slot back: qcur and pcur are unchanged
    h->qcur = a;
    h->pcur = b;
    return QG_OK;
}

int qg_step(qg_handle* h, int first_timestep, int nsteps) {
    if (!h) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_step: no state uploaded");
    if (first_timestep < 1 || nsteps < 0) return fail(h, QG_ERR_INVALID, "qg_step: bad timestep range");
    QG_CUDA(h, cudaSetDevice(h->device));
    const int end = first_timestep + nsteps;
    int t = first_timestep;
    while (t < end) {
        if (h->use_graph && h->warm && !h->profiling && h->dist_n == 1 && t >= 3 && end - t >= 3) {
            const int rc = step_cycle_graph(h, t);
            if (rc == QG_OK) { t += 3; continue; }
            if (rc < 0) return rc;
        }
        int rc = do_evolve_zeta(h, t);
        if (rc) return rc;
        rc = do_evolve_psi(h);
        if (rc) return rc;
        h->warm = true;
        ++t;
    }
    return QG_OK;
}

int qg_sync(qg_handle* h) {
    if (!h) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    QG_CUDA(h, cudaGetLastError());
    return dist_poll_error(h);
}

int qg_diagnostics(qg_handle* h, double* energy, double* enstrophy) {
    if (!h || !energy || !enstrophy) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_diagnostics: no state uploaded");
    QG_CUDA(h, cudaSetDevice(h->device));
    QG_CUDA(h, launch_diag(h));
    if (h->dist_n > 1)   // local sums -> domain sums
        QG_CUDA(h, dist_allreduce_sum(h, h->diag_part + (int64_t)h->nm * h->diag_blocks * 2, 2 * (size_t)h->nm));
    std::vector<double> out(2 * h->nm);
    QG_CUDA(h, cudaMemcpyAsync(out.data(), h->diag_part + (int64_t)h->nm * h->diag_blocks * 2,
                               out.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (int rc = dist_poll_error(h)) return rc;
    for (int m = 0; m < h->nm; ++m) {
        energy[m] = out[2 * m];
        enstrophy[m] = out[2 * m + 1];
    }
    return QG_OK;
}

int qg_extrema(qg_handle* h, double* out) {
    if (!h || !out) return QG_ERR_INVALID;
    if (!h->have_state) return fail(h, QG_ERR_STATE, "qg_extrema: no state uploaded");
    QG_CUDA(h, cudaSetDevice(h->device));
    const int nb = h->diag_blocks;
    if (!h->ext_part) QG_CUDA(h, cudaMalloc((void**)&h->ext_part, ((size_t)h->nm * nb + h->nm) * 8 * sizeof(double)));
    double* res = h->ext_part + (size_t)h->nm * nb * 8;
    {
        KernelTimer t(h, QG_K_DIAG);
        k5_extrema_partial<<<dim3(nb, h->nm), 256, 0, h->stream>>>(h->field(h->q, h->qcur, 0, 0),
                                                                   h->field(h->psi, h->pcur, 0, 0), h->g, h->ext_part);
    }
    {
        KernelTimer t(h, QG_K_DIAG);
        k5_extrema_final<<<h->nm, 256, 0, h->stream>>>(h->ext_part, nb, res);
    }
    QG_CUDA(h, cudaGetLastError());
    if (h->dist_n > 1) QG_CUDA(h, dist_allreduce_max(h, res, 8 * (size_t)h->nm));   // slab extrema -> domain extrema
    std::vector<double> tmp(8 * (size_t)h->nm);
    QG_CUDA(h, cudaMemcpyAsync(tmp.data(), res, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (int rc = dist_poll_error(h)) return rc;
    for (size_t k = 0; k < tmp.size(); ++k) out[k] = (k & 1) ? -tmp[k] : tmp[k];
    return QG_OK;
}

int qg_solve(qg_handle* h, int pinned, const double* f, double* u) {
    if (!h || !f || !u) return QG_ERR_INVALID;
    if (h->dist_n > 1) return fail(h, QG_ERR_STATE, "qg_solve: not available on a y-slab handle");
    if (!h->plan_ok) return fail(h, QG_ERR_STATE, "qg_solve: the handle has no plan");
    QG_CUDA(h, cudaSetDevice(h->device));
    const Geom& g = h->g;
    // Four padded fields of its own (two in, two out), kept on the handle: the state arrays stay untouched.
    if (!h->solve_tmp) QG_CUDA(h, cudaMalloc((void**)&h->solve_tmp, 4 * g.fstride * sizeof(double)));
    double* tmp = h->solve_tmp;
    // both layers carry f: field 1 solves Poisson, field 2 Helmholtz
    cudaError_t e = cudaSuccess;
    for (int l = 0; l < 2 && e == cudaSuccess; ++l)
        e = cudaMemcpy2DAsync(tmp + l * g.fstride + (int64_t)(YPAD - 1) * g.pitch + (XPAD - 1), (size_t)g.pitch * sizeof(double), f,
                              (size_t)(g.M + 2) * sizeof(double), (size_t)(g.M + 2) * sizeof(double), (size_t)(g.P + 2),
                              cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = fill_ghosts(h, tmp, 2);
    // identity projections: psi~ = solve(f) per field
    qg_params saved = h->prm;
    const double I4[4] = {1.0, 0.0, 0.0, 1.0};
    memcpy(h->prm.Pinv, I4, sizeof(I4));
    memcpy(h->prm.Pfwd, I4, sizeof(I4));
    const int nm_saved = h->nm;
    h->nm = 1;
    if (e == cudaSuccess) e = launch_fft_forward(h, tmp, 0);
    if (e == cudaSuccess) e = launch_ysolve(h, pinned ? 1 : 0, 0);
    if (e == cudaSuccess) e = launch_fft_inverse(h, tmp + 2 * g.fstride, pinned ? 1 : 0);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(u, (size_t)(g.M + 2) * sizeof(double),
                              tmp + (2 + (pinned ? 0 : 1)) * g.fstride + (int64_t)(YPAD - 1) * g.pitch + (XPAD - 1),
                              (size_t)g.pitch * sizeof(double), (size_t)(g.M + 2) * sizeof(double), (size_t)(g.P + 2),
                              cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    h->nm = nm_saved;
    h->prm = saved;
    QG_CUDA(h, e);
    return QG_OK;
}

int qg_plan_probe(const qg_params* p, int rows_local, double* r, double* kappa, int32_t* worklist,
                  int worklist_capacity, int* worklist_len) {
    if (!p || p->M < 3 || !(p->dx > 0.0)) return QG_ERR_INVALID;
    const int M = p->M, ncol = 2 * M;
    const long double dx2 = (long double)p->dx * (long double)p->dx;
    std::vector<double> rt(ncol), kp(ncol), lr(ncol);
    for (int col = 0; col < ncol; ++col) {
        int field, k;
        bool is_re, singular;
        column_mode(col, M, &field, &k, &is_re);
        const long double rr = column_root(field, k, M, dx2, p->alpha, &singular);
        rt[col] = (double)rr;
        kp[col] = singular ? 0.0 : (double)(-rr * dx2 / M);
        lr[col] = singular ? 0.0 : (double)logl(rr);
        if (r) r[col] = rt[col];
        if (kappa) kappa[col] = kp[col];
    }
    if (worklist_len) {
        if (rows_local < 32 || rows_local % 32 != 0) return QG_ERR_INVALID;
        std::vector<int2> work;
        edge_worklist(rt, kp, lr, ncol, rows_local, &work);
        *worklist_len = (int)work.size();
        if (worklist)
            for (int i = 0; i < (int)work.size() && i < worklist_capacity; ++i) {
                worklist[2 * i] = work[i].x;
                worklist[2 * i + 1] = work[i].y;
            }
    }
    return QG_OK;
}

int qg_nccl_unique_id(void* out128) {
    if (!out128) return QG_ERR_INVALID;
    std::string err;
    int rc = dist_unique_id(out128, &err);
    if (rc) g_create_error = err;
    return rc;
}

int qg_dist_init(qg_handle* h, int rank, int nranks, const void* unique_id128) {
    if (!h || !unique_id128) return QG_ERR_INVALID;
    if (h->dist_n > 1) return fail(h, QG_ERR_STATE, "qg_dist_init: already initialised");
    if (nranks < 2 || nranks > 8 || rank < 0 || rank >= nranks)
        return fail(h, QG_ERR_INVALID, "qg_dist_init: need 2 <= nranks <= 8 and 0 <= rank < nranks");
    if (h->nm != 1) return fail(h, QG_ERR_INVALID, "qg_dist_init: y-slab mode takes one member per handle");
    if (h->g.P % 32 != 0 || !h->plan.tp_ok)
        return fail(h, QG_ERR_INVALID, "qg_dist_init: local row count must be a multiple of 32 and at most 8192");
    if ((int64_t)h->g.P * nranks > 16384) return fail(h, QG_ERR_INVALID, "qg_dist_init: global P > 16384");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = dist_init(h, rank, nranks, unique_id128);
    if (rc) return rc;
    // the cyclic closure now spans the global row count: rebuild the coefficient tables
    free_plan(h);
    QG_CUDA(h, build_plan(h));
    if ((h->plan.ts_ok || h->plan.tp_ok) && (rc = make_tensor_map_S(h))) return rc;   // the column table moved with the plan
    h->have_state = false;
    return QG_OK;
}

int qg_dist_ipc_export(qg_handle* h, void* out256) {
    if (!h || !out256) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    return dist_ipc_export(h, out256);
}

int qg_dist_ipc_blobs_share_device(const void* all_ranks, int nranks) {
    if (!all_ranks || nranks < 1 || nranks > 8) return QG_ERR_INVALID;
    return dist_blobs_share_device(all_ranks, nranks);
}

int qg_dist_ipc_import(qg_handle* h, const void* all_ranks) {
    if (!h || !all_ranks) return QG_ERR_INVALID;
    if (getenv("QG_DIST_NCCL") || getenv("QG_K3_V1")) return QG_OK;   // diagnostics: stay on the NCCL path
    QG_CUDA(h, cudaSetDevice(h->device));
    return dist_ipc_import(h, all_ranks);
}

int qg_set_profiling(qg_handle* h, int enabled) {
    if (!h) return QG_ERR_INVALID;
    h->profiling = enabled ? 1 : 0;
    for (int i = 0; i < QG_NKERNELS; ++i) { h->kms[i] = 0.0; h->kcount[i] = 0; }
    h->evkernel.clear();
    return QG_OK;
}

int qg_kernel_times(qg_handle* h, double* ms, int64_t* launches) {
    if (!h) return QG_ERR_INVALID;
    QG_CUDA(h, cudaSetDevice(h->device));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    for (size_t n = 0; n < h->evkernel.size(); ++n) {   // fold the recorded pairs into the sums
        float t = 0.f;
        QG_CUDA(h, cudaEventElapsedTime(&t, h->evpool[2 * n], h->evpool[2 * n + 1]));
        h->kms[h->evkernel[n]] += t;
    }
    h->evkernel.clear();
    for (int i = 0; i < QG_NKERNELS; ++i) {
        if (ms) ms[i] = h->kms[i];
        if (launches) launches[i] = h->kcount[i];
    }
    return QG_OK;
}

int64_t qg_launch_count(const qg_handle* h) { return h ? h->launches : 0; }

int qg_device_layout(qg_handle* h, int which, void** base, int64_t* pitch, int64_t* xpad, int64_t* ypad,
                     int64_t* field_stride) {
    if (!h) return QG_ERR_INVALID;
    void* b = nullptr;
    switch (which) {
        case 0: b = h->q; break;
        case 1: b = h->psi; break;
        case 2: b = h->f; break;
        case 3: b = h->S; break;
        default: return fail(h, QG_ERR_INVALID, "qg_device_layout: which must be 0..3");
    }
    if (base) *base = b;
    if (which == 3) {
        if (pitch) *pitch = h->plan.ncol;
        if (xpad) *xpad = 0;
        if (ypad) *ypad = 0;
        if (field_stride) *field_stride = (int64_t)h->plan.P * h->plan.ncol;
    } else {
        if (pitch) *pitch = h->g.pitch;
        if (xpad) *xpad = XPAD;
        if (ypad) *ypad = YPAD;
        if (field_stride) *field_stride = h->g.fstride;
    }
    return QG_OK;
}

}  // extern "C"
