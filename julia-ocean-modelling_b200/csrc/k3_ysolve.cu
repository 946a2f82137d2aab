// K3 — the y-direction solve of the inversion, for both modal fields at once.
//
// After the x-FFT (k2_fft.cu) every real column c of the spectral array obeys a cyclic
// Toeplitz tridiagonal system along y (reference matrix: src/schemes/laplacian.jl:40-58,
// periodic 1-D Laplacian pinned by src/test.jl:229-238):
//
//      u[j-1] + d_c u[j] + u[j+1] = dx^2 * qhat[j],   d_c = 2cos(2 pi k/M) - 4 + alpha dx^2,
//
// j mod P, alpha = 0 for the Poisson (barotropic) field and S_eig for the Helmholtz
// (baroclinic) one.  With r the root of r^2 + d r + 1 = 0 inside the unit circle the
// operator factorises exactly,  S^-1 + d + S = -(1/r) (1 - r S^-1)(1 - r S),  so the
// solve is two cyclic first-order recurrences (forward then backward) and a scale by -r.
//
// Each recurrence is solved by blocked cyclic reduction: a warp owns 32 adjacent columns
// (one 256-byte segment per row) and a 32-row chunk, runs the chunk serially in registers,
// and the chunk-to-chunk carries are a cyclic bidiagonal system reduced CTA-locally in
// shared memory and then across the thread-block cluster through distributed shared
// memory (cluster.sync, no global round trip).  Every element is read once and written
// once: 16 B per real column entry, 32 B per grid cell for the two fields.
//
// Special cases: the Poisson k=0 column is singular (d = -2, r = 1); k3_pre solves it by
// double prefix sums and also produces the global sum needed for the reference's pinned
// node (src/schemes/laplacian.jl:66-75, src/model.jl:185): the pin is equivalent to
// replacing rhs(0,0) so that the right-hand side sums to zero, i.e. subtracting the
// total from the real part of every Poisson coefficient of row 0.  k3_gauge evaluates
// psi~1(0,0), which k4 subtracts (the pinned unknown is exactly zero in the reference).
#include <cooperative_groups.h>

#include "qg_internal.cuh"

namespace cg = cooperative_groups;

namespace qg {

constexpr int CH = 32;   // rows per chunk

constexpr int PRE_T = 1024;
constexpr int PRE_E = 16;   // rows per thread, P <= 16384

// One block per member: statistics and singular solve of the Poisson k=0 column (column 0).
__global__ void __launch_bounds__(PRE_T) k3_pre(const YArgs a) {
    __shared__ double sh[32];
    const int member = blockIdx.x;
    const int P = a.pl.P;
    const int e = (P + PRE_T - 1) / PRE_T;
    const double* __restrict__ col = a.S + member * a.sstride;
    const int j0 = threadIdx.x * e;
    double b[PRE_E];
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        b[i] = (i < e && j < P) ? col[(int64_t)j * a.pl.ncol] : 0.0;
        loc += b[i];
    }
    const double total = block_sum(loc, sh);
    const double pin = a.pinned ? total : 0.0;
    // g = (b - delta_j0 * pin) * dx^2 / M ; if not pinned the column mean is removed instead
    // (solvability of the singular periodic problem).
    const double mean = a.pinned ? 0.0 : total / P;
    double c1[PRE_E];
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        double g = 0.0;
        if (i < e && j < P) g = (b[i] - (j == 0 ? pin : 0.0) - mean) * a.pl.k0scale;
        run += g;
        c1[i] = run;   // local inclusive prefix
    }
    const double off1 = block_exclusive_scan(run, sh);
    double s1 = 0.0;
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        c1[i] += off1;
        if (i < e && j < P) s1 += c1[i];
    }
    const double sum_c1 = block_sum(s1, sh);
    const double dm1 = -sum_c1 / P;
    // D[j] = dm1 + c1[j];  x[j] = sum_{m<j} D[m], x[0] = 0
    double run2 = 0.0;
    double xs[PRE_E];
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        xs[i] = run2;   // exclusive within the thread
        if (i < e && j < P) run2 += dm1 + c1[i];
    }
    const double off2 = block_exclusive_scan(run2, sh);
    double* __restrict__ out = a.k0sol + (int64_t)member * P;
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        if (i < e && j < P) out[j] = xs[i] + off2;
    }
    if (threadIdx.x == 0) a.scal[member * 4 + 0] = pin;
}

// One block per member: psi~1 at node (0,0) = sum over all x wavenumbers of row 0 of the
// solved Poisson field (normalisation is already folded into the column scale).
__global__ void __launch_bounds__(256) k3_gauge(const YArgs a) {
    __shared__ double sh[32];
    const int member = blockIdx.x;
    const int M = a.pl.M;
    const double* __restrict__ row0 = a.S + member * a.sstride;
    const int kmax = (M + 1) / 2;   // paired wavenumbers 1 .. kmax-1
    double loc = 0.0;
    for (int k = 1 + threadIdx.x; k < kmax; k += blockDim.x) loc += 2.0 * row0[2 * k];
    double t = block_sum(loc, sh);
    if (threadIdx.x == 0) {
        t += row0[0];
        if ((M & 1) == 0) t += row0[M];
        a.scal[member * 4 + 1] = t;
    }
}

// ---- the cluster kernel ------------------------------------------------------------------
// grid = (slabs * CS, members), cluster = (CS,1,1), block = 32 * wpc threads.
// dynamic smem: 4 arrays [wpc*m][32] (F, G, A, B) + 3 arrays [32] (FF, RR, GG).
template <bool KEEP>
__global__ void __launch_bounds__(512) k3_ysolve(const YArgs a) {
    extern __shared__ __align__(16) double ysm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = a.pl.CS, wpc = a.pl.wpc, m = a.pl.m, C = a.pl.C, P = a.pl.P;
    const int cr = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slab = blockIdx.x / CS;
    const int member = blockIdx.y;
    const int col = slab * 32 + lane;
    const bool cvalid = col < a.pl.ncol;
    const int ccol = cvalid ? col : 0;
    const int nloc = wpc * m;
    double* sF = ysm;
    double* sG = sF + nloc * 32;
    double* sA = sG + nloc * 32;
    double* sB = sA + nloc * 32;
    double* sFF = sB + nloc * 32;
    double* sRR = sFF + 32;
    double* sGG = sRR + 32;

    const double r = cvalid ? __ldg(a.pl.rtab + ccol) : 0.0;
    const double kap = cvalid ? __ldg(a.pl.kap + ccol) : 0.0;
    const double pinv = (a.pinned && cvalid) ? __ldg(a.pl.pinw + ccol) * a.scal[member * 4 + 0] : 0.0;
    double* __restrict__ base = a.S + member * a.sstride + ccol;
    const int ncol = a.pl.ncol;

    double arr[CH];
    const int cbase = (cr * wpc + warp) * m;   // first global chunk of this warp

    // ---- pass 1: zero-carry forward recurrence per chunk, chunk sums F and G -------------
    for (int ci = 0; ci < m; ++ci) {
        const int c = cbase + ci;
        const int j0 = c * CH;
        const int len = c < C ? min(CH, P - j0) : 0;
#pragma unroll
        for (int i = 0; i < CH; ++i)
            arr[i] = (i < len && cvalid) ? base[(int64_t)(j0 + i) * ncol] : 0.0;
        if (j0 == 0) arr[0] -= pinv;
        double y = 0.0;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (i < len) {
                y = fma(r, y, arr[i]);
                arr[i] = y;
            }
        }
        double z = 0.0;
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            if (i < len) z = fma(r, z, arr[i]);
        }
        sF[(warp * m + ci) * 32 + lane] = y;
        sG[(warp * m + ci) * 32 + lane] = z;
    }
    __syncthreads();

    // per-chunk decay and geometric factors for this CTA's chunks
    const double rho32 = cvalid ? __ldg(a.pl.rho32 + ccol) : 0.0;
    const double rhoL = cvalid ? __ldg(a.pl.rhoL + ccol) : 0.0;
    const double h32 = cvalid ? __ldg(a.pl.h32 + ccol) : 0.0;
    const double hL = cvalid ? __ldg(a.pl.hL + ccol) : 0.0;
    const double inv1 = cvalid ? __ldg(a.pl.inv1mrP + ccol) : 0.0;
    const int cta_c0 = cr * nloc;
    auto rho_of = [&](int c) { return c < C - 1 ? rho32 : (c == C - 1 ? rhoL : 1.0); };
    auto h_of = [&](int c) { return c < C - 1 ? h32 : (c == C - 1 ? hL : 0.0); };

    // ---- CTA-level forward aggregate ----------------------------------------------------
    if (warp == 0) {
        double t = 0.0, R = 1.0;
        for (int lc = 0; lc < nloc; ++lc) {
            const double rho = rho_of(cta_c0 + lc);
            t = fma(rho, t, sF[lc * 32 + lane]);
            R *= rho;
        }
        sFF[lane] = t;
        sRR[lane] = R;
    }
    cluster.sync();

    // ---- forward carries, then the A-dependent backward sums ---------------------------------
    if (warp == 0) {
        double t = 0.0, mine = 0.0;
        // closure over the CS CTAs: y at the last row of the column
        for (int i = 0; i < CS; ++i) {
            const double* rFF = cluster.map_shared_rank(sFF, i);
            const double* rRR = cluster.map_shared_rank(sRR, i);
            t = fma(rRR[lane], t, rFF[lane]);
        }
        double acar = t * inv1;   // carry into chunk 0 (cyclic)
        for (int i = 0; i < cr; ++i) {
            const double* rFF = cluster.map_shared_rank(sFF, i);
            const double* rRR = cluster.map_shared_rank(sRR, i);
            acar = fma(rRR[lane], acar, rFF[lane]);
        }
        (void)mine;
        double g = 0.0;
        for (int lc = 0; lc < nloc; ++lc) {
            const int c = cta_c0 + lc;
            sA[lc * 32 + lane] = acar;
            sG[lc * 32 + lane] = fma(acar, h_of(c), sG[lc * 32 + lane]);
            acar = fma(rho_of(c), acar, sF[lc * 32 + lane]);
        }
        for (int lc = nloc - 1; lc >= 0; --lc) g = fma(rho_of(cta_c0 + lc), g, sG[lc * 32 + lane]);
        sGG[lane] = g;
    }
    cluster.sync();

    // ---- backward carries ------------------------------------------------------------------
    if (warp == 0) {
        double t = 0.0;
        for (int i = CS - 1; i >= 0; --i) {
            const double* rGG = cluster.map_shared_rank(sGG, i);
            const double* rRR = cluster.map_shared_rank(sRR, i);
            t = fma(rRR[lane], t, rGG[lane]);
        }
        double bcar = t * inv1;   // carry into the last chunk (cyclic)
        for (int i = CS - 1; i > cr; --i) {
            const double* rGG = cluster.map_shared_rank(sGG, i);
            const double* rRR = cluster.map_shared_rank(sRR, i);
            bcar = fma(rRR[lane], bcar, rGG[lane]);
        }
        for (int lc = nloc - 1; lc >= 0; --lc) {
            sB[lc * 32 + lane] = bcar;
            bcar = fma(rho_of(cta_c0 + lc), bcar, sG[lc * 32 + lane]);
        }
    }
    __syncthreads();

    // ---- pass 2: apply carries, backward recurrence, scale, store -----------------------------
    const double* __restrict__ k0 = a.k0sol + (int64_t)member * P;
    for (int ci = 0; ci < m; ++ci) {
        const int c = cbase + ci;
        const int j0 = c * CH;
        const int len = c < C ? min(CH, P - j0) : 0;
        const double A = sA[(warp * m + ci) * 32 + lane];
        const double B = sB[(warp * m + ci) * 32 + lane];
        if (KEEP) {
            double pw = r;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (i < len) {
                    arr[i] = fma(pw, A, arr[i]);
                    pw *= r;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                arr[i] = (i < len && cvalid) ? base[(int64_t)(j0 + i) * ncol] : 0.0;
            if (j0 == 0) arr[0] -= pinv;
            double y = A;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (i < len) {
                    y = fma(r, y, arr[i]);
                    arr[i] = y;
                }
            }
        }
        double z = B;
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            if (i < len) {
                z = fma(r, z, arr[i]);
                arr[i] = kap * z;
            }
        }
        if (cvalid) {
            if (col == 0) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    if (i < len) arr[i] = k0[j0 + i];
            }
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (i < len) base[(int64_t)(j0 + i) * ncol] = arr[i];
        }
    }
    cluster.sync();   // keep distributed shared memory alive until every CTA has read it
}

cudaError_t launch_ysolve(Handle* h, int pinned, int /*unused*/) {
    YArgs a{};
    a.pl = h->plan;
    a.S = h->S;
    a.sstride = (int64_t)h->plan.P * h->plan.ncol;
    a.k0sol = h->k0sol;
    a.scal = h->scal;
    a.pinned = pinned;
    {
        KernelTimer t(h, QG_K_YPRE);
        k3_pre<<<h->nm, PRE_T, 0, h->stream>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    {
        const Plan& pl = h->plan;
        const int nslab = (pl.ncol + 31) / 32;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(nslab * pl.CS, h->nm, 1);
        cfg.blockDim = dim3(32 * pl.wpc, 1, 1);
        cfg.dynamicSmemBytes = (size_t)(4 * pl.wpc * pl.m + 3) * 32 * sizeof(double);
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pl.CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        KernelTimer t(h, QG_K_YSOLVE);
        if (pl.m == 1) {
            if (cfg.dynamicSmemBytes > 48 * 1024)
                cudaFuncSetAttribute(k3_ysolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)cfg.dynamicSmemBytes);
            e = cudaLaunchKernelEx(&cfg, k3_ysolve<true>, a);
        } else {
            if (cfg.dynamicSmemBytes > 48 * 1024)
                cudaFuncSetAttribute(k3_ysolve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)cfg.dynamicSmemBytes);
            e = cudaLaunchKernelEx(&cfg, k3_ysolve<false>, a);
        }
    }
    if (e != cudaSuccess) return e;
    {
        KernelTimer t(h, QG_K_GAUGE);
        k3_gauge<<<h->nm, 256, 0, h->stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace qg
