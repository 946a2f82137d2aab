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
#include <cstdio>
#include <cstdlib>

#include "qg_internal.cuh"

namespace cg = cooperative_groups;

namespace qg {

constexpr int CH = 32;   // rows per chunk

constexpr int PRE_T = 1024;

// One block per member: statistics and singular solve of the Poisson k=0 column (column 0).
// PRE_E = rows per thread (P <= 1024 * PRE_E).
template <int PRE_E>
__global__ void __launch_bounds__(PRE_T) k3_pre(const YArgs a) {
    __shared__ double sh[32];
    const int member = blockIdx.x;
    const int P = a.preP;
    const int e = (P + PRE_T - 1) / PRE_T;
    const double* __restrict__ col = a.col0 + (int64_t)member * P;
    const int j0 = threadIdx.x * e;
    double b[PRE_E];
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        b[i] = (i < e && j < P) ? col[j] : 0.0;
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
    double* __restrict__ out = a.k0sol + (int64_t)member * P;   // P == a.preP here
#pragma unroll
    for (int i = 0; i < PRE_E; ++i) {
        const int j = j0 + i;
        if (i < e && j < P) out[j] = xs[i] + off2;
    }
    if (threadIdx.x == 0) a.scal[member * 4 + 0] = pin;
}

static void launch_pre(const YArgs& a, int nblocks, cudaStream_t st) {
    const int e = (a.preP + PRE_T - 1) / PRE_T;
    if (e <= 1) k3_pre<1><<<nblocks, PRE_T, 0, st>>>(a);
    else if (e <= 2) k3_pre<2><<<nblocks, PRE_T, 0, st>>>(a);
    else if (e <= 4) k3_pre<4><<<nblocks, PRE_T, 0, st>>>(a);
    else if (e <= 8) k3_pre<8><<<nblocks, PRE_T, 0, st>>>(a);
    else k3_pre<16><<<nblocks, PRE_T, 0, st>>>(a);
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
        for (int r = 0; r < a.peer_n; ++r) a.scal_peer[r][member * 4 + 1] = t;   // y-slab peer mode: every rank's copy
        if (a.peer_n > 0) __threadfence_system();   // posted NVLink writes acknowledged before the kernel ends
    }
}

// ---- the cluster kernel ------------------------------------------------------------------
// grid = (slabs * CS, members), cluster = (CS,1,1), block = 32 * wpc threads.
// dynamic smem: 4 arrays [wpc*m][32] (F, G, A, B) + 4 arrays [32] (FF, RR, X, Y).
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

    // ---- carries: one exchange round across the cluster ---------------------------------------
    // Within the CTA the carry into local chunk lc is affine in the carry A_s into the CTA's
    // first chunk:  A_lc = pF_lc + Rpre_lc * A_s.  The backward sums need the true forward
    // values, G'_lc = G_lc + h_lc * A_lc, so the CTA's backward aggregate is affine in A_s too:
    // GG' = X + Y * A_s.  Publishing (FF, RR, X, Y) lets every CTA close both cyclic
    // recurrences after a single cluster.sync().
    double* sX = sGG;          // [32]
    double* sY = sGG + 32;     // [32]  (the launcher reserves 4 trailing rows)
    if (warp == 0) {
        double t = 0.0, R = 1.0, X = 0.0, Y = 0.0;
#pragma unroll 4
        for (int lc = 0; lc < nloc; ++lc) {
            const int c = cta_c0 + lc;
            const double rho = rho_of(c), hh = h_of(c);
            const double F = sF[lc * 32 + lane], G = sG[lc * 32 + lane];
            sA[lc * 32 + lane] = t;    // pF
            sB[lc * 32 + lane] = R;    // Rpre
            const double g0 = fma(hh, t, G), g1 = hh * R;
            X = fma(R, g0, X);
            Y = fma(R, g1, Y);
            t = fma(rho, t, F);
            R *= rho;
        }
        sFF[lane] = t;
        sRR[lane] = R;
        sX[lane] = X;
        sY[lane] = Y;
    }
    cluster.sync();

    if (warp == 0) {
        double FFi[8], RRi[8], Xi[8], Yi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < CS) {
                FFi[i] = cluster.map_shared_rank(sFF, i)[lane];
                RRi[i] = cluster.map_shared_rank(sRR, i)[lane];
                Xi[i] = cluster.map_shared_rank(sX, i)[lane];
                Yi[i] = cluster.map_shared_rank(sY, i)[lane];
            } else {
                FFi[i] = 0.0; RRi[i] = 1.0; Xi[i] = 0.0; Yi[i] = 0.0;
            }
        }
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t = fma(RRi[i], t, FFi[i]);
        double As[8];
        As[0] = t * inv1;   // carry into chunk 0 = y at the last row (cyclic closure)
#pragma unroll
        for (int i = 0; i < 7; ++i) As[i + 1] = fma(RRi[i], As[i], FFi[i]);
        double GGp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) GGp[i] = fma(Yi[i], As[i], Xi[i]);
        t = 0.0;
#pragma unroll
        for (int i = 7; i >= 0; --i) t = fma(RRi[i], t, GGp[i]);
        double Be[8];
        Be[7] = t * inv1;   // carry into the last chunk = z at row 0 (cyclic closure)
#pragma unroll
        for (int i = 7; i > 0; --i) Be[i - 1] = fma(RRi[i], Be[i], GGp[i]);
        double a_s = 0.0, b_e = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i == cr) { a_s = As[i]; b_e = Be[i]; }
        }
#pragma unroll 4
        for (int lc = 0; lc < nloc; ++lc) {
            const double A = fma(sB[lc * 32 + lane], a_s, sA[lc * 32 + lane]);
            sA[lc * 32 + lane] = A;
            sG[lc * 32 + lane] = fma(A, h_of(cta_c0 + lc), sG[lc * 32 + lane]);
        }
        double bcar = b_e;
#pragma unroll 4
        for (int lc = nloc - 1; lc >= 0; --lc) {
            sB[lc * 32 + lane] = bcar;
            bcar = fma(rho_of(cta_c0 + lc), bcar, sG[lc * 32 + lane]);
        }
    }
    __syncthreads();

    // ---- pass 2: apply carries, backward recurrence, scale, store -----------------------------
    const double* __restrict__ k0 = a.k0sol + (int64_t)member * a.preP;
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

// ---- TMA-staged variant ---------------------------------------------------------------------
// Same algorithm, different residency: the CTA's part of the slab (rows_cta rows x 16 columns,
// one 128-byte line per row) is brought into shared memory by TMA tensor copies
// (cp.async.bulk.tensor.2d, one 32-row box per chunk, all completing on one mbarrier) and the
// sweeps run out of shared memory with a handful of registers, so several CTAs of different
// clusters share an SM and their load / compute / store phases overlap.  A half-warp owns a
// chunk (16 lanes = 16 adjacent columns); pass 2 stores straight to global memory.
// grid = (slabs * CS, members), cluster = (CS,1,1), block = 16 * nchunk threads.
constexpr int TS_WC = 16;
constexpr int TS_LD = 17;   // padded leading dimension of the per-chunk carry arrays

#ifdef QG_K3_TRACE
__device__ long long* g_k3_trace = nullptr;   // [CTA][8] clock64 stamps of thread 0
#define K3_STAMP(i) do { if (tid == 0 && g_k3_trace) g_k3_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = clock64(); } while (0)
#else
#define K3_STAMP(i) do { } while (0)
#endif

template <int MODE>   // 0: cyclic over the local rows; 1 / 2: y-slab mode, see YArgs::mode
__global__ void __launch_bounds__(256, 3)
k3_ysolve_tma(const __grid_constant__ CUtensorMap tmS, const YArgs a, int nchunk) {
    extern __shared__ __align__(128) unsigned char ts_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int cr = (int)cluster.block_rank();
    const int C = a.pl.C, P = a.pl.P, ncol = a.pl.ncol;
    const int tid = threadIdx.x;
    const int chunk = tid >> 4, l = tid & 15;
    const int slab = blockIdx.x / CS;
    const int member = blockIdx.y;
    const int col0 = slab * TS_WC;
    const int col = col0 + l;
    const bool cvalid = col < ncol;
    const int ccol = cvalid ? col : 0;

    double* tile = reinterpret_cast<double*>(ts_raw);                  // [nchunk*32][16]
    double* pw = tile + (size_t)nchunk * 32 * TS_WC;                   // [32][16]  r^(i+1)
    double* sF = pw + 32 * TS_WC;                                      // [nchunk][TS_LD]
    double* sG = sF + nchunk * TS_LD;
    double* sFF = sG + nchunk * TS_LD;                                 // [16] each
    double* sRR = sFF + TS_WC;
    double* sX = sRR + TS_WC;
    double* sY = sX + TS_WC;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sY + TS_WC);

    K3_STAMP(0);
    const int c0 = cr * nchunk;                 // first global chunk of this CTA
    const int nact = max(0, min(nchunk, C - c0));   // chunks that hold rows
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0 && nact > 0) {
        mbar_expect_tx(bar, (uint32_t)nact * 32 * TS_WC * sizeof(double));
        for (int k = 0; k < nact; ++k)
            tma_load_2d(tile + (size_t)k * 32 * TS_WC, &tmS, col0, member * P + (c0 + k) * 32, bar);
    }
    const double r = cvalid ? __ldg(a.pl.rtab + ccol) : 0.0;
    const double kap = cvalid ? __ldg(a.pl.kap + ccol) : 0.0;
    const double pinv = (a.pinned && cvalid) ? __ldg(a.pl.pinw + ccol) * a.scal[member * 4 + 0] : 0.0;
    if (chunk == 0) {   // first half-warp: powers r^(i+1) of this slab's columns
        double p = r;
        for (int i = 0; i < 32; ++i) {
            pw[i * TS_WC + l] = p;
            p *= r;
        }
    }
    const int c = c0 + chunk;
    const int j0 = c * 32;
    const bool mine = chunk < nchunk;   // the block is padded to whole warps
    const int len = (mine && c < C) ? min(32, P - j0) : 0;
    double* t = tile + (size_t)(mine ? chunk : 0) * 32 * TS_WC + l;
    K3_STAMP(1);
    if (nact > 0) mbar_wait(bar, 0);
    K3_STAMP(2);

    // ---- pass 1: zero-carry forward recurrence in place, forward end value and backward sum ----
    {
        double y = 0.0, G = 0.0, p = 1.0;
#pragma unroll 4
        for (int i = 0; i < len; ++i) {
            double b = t[i * TS_WC];
            if (a.row0 + j0 + i == 0) b -= pinv;
            y = fma(r, y, b);
            t[i * TS_WC] = y;
            G = fma(p, y, G);
            p *= r;
        }
        if (mine) {
            sF[chunk * TS_LD + l] = y;
            sG[chunk * TS_LD + l] = G;
        }
    }
    K3_STAMP(3);
    __syncthreads();

    // ---- carries (see k3_ysolve): one exchange round across the cluster ---------------------
    // Warp-level parallel cyclic reduction: lanes 0-15 / 16-31 of a warp hold the (up to 16) chunks
    // of two columns; the affine maps x -> rho x + F compose by Kogge-Stone shuffles.
    const int warp = tid >> 5, lane = tid & 31, nwarp = (blockDim.x + 31) >> 5;
    const int slot = lane & 15, csel = lane >> 4;
    const bool act = slot < nchunk;
    const int cc = c0 + slot;
    // forward scan of one column pair: returns the exclusive prefix (pF, Rpre) and the totals
    auto fwd_scan = [&](double rho, double F, double& pF, double& Rpre, double& FFv, double& RRv) {
        double R = rho, t = F;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const double Rp = __shfl_up_sync(0xffffffffu, R, d, 16);
            const double tp = __shfl_up_sync(0xffffffffu, t, d, 16);
            if (slot >= d) {
                t = fma(R, tp, t);
                R *= Rp;
            }
        }
        FFv = __shfl_sync(0xffffffffu, t, 15, 16);
        RRv = __shfl_sync(0xffffffffu, R, 15, 16);
        Rpre = __shfl_up_sync(0xffffffffu, R, 1, 16);
        pF = __shfl_up_sync(0xffffffffu, t, 1, 16);
        if (slot == 0) { Rpre = 1.0; pF = 0.0; }
    };
    for (int cb = warp; cb < TS_WC / 2; cb += nwarp) {
        const int cl = 2 * cb + csel;                 // column within the slab
        const int gc = (col0 + cl < ncol) ? col0 + cl : 0;
        const double rho = !act || cc > C - 1 ? 1.0 : __ldg((cc == C - 1 ? a.pl.rhoL : a.pl.rho32) + gc);
        const double hh = !act || cc > C - 1 ? 0.0 : __ldg((cc == C - 1 ? a.pl.hL : a.pl.h32) + gc);
        const double F = act ? sF[slot * TS_LD + cl] : 0.0;
        const double G = act ? sG[slot * TS_LD + cl] : 0.0;
        double pF, Rpre, FFv, RRv;
        fwd_scan(rho, F, pF, Rpre, FFv, RRv);
        double Xv = Rpre * fma(hh, pF, G), Yv = Rpre * (hh * Rpre);
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) {
            Xv += __shfl_xor_sync(0xffffffffu, Xv, d, 16);
            Yv += __shfl_xor_sync(0xffffffffu, Yv, d, 16);
        }
        if (slot == 0) { sFF[cl] = FFv; sRR[cl] = RRv; sX[cl] = Xv; sY[cl] = Yv; }
    }
    K3_STAMP(4);
    cluster.sync();
    K3_STAMP(5);
    for (int cb = warp; cb < TS_WC / 2; cb += nwarp) {
        const int cl = 2 * cb + csel;
        const int gc = (col0 + cl < ncol) ? col0 + cl : 0;
        const double inv1 = __ldg(a.pl.inv1mrP + gc);
        double FFi[8], RRi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            FFi[i] = i < CS ? cluster.map_shared_rank(sFF, i)[cl] : 0.0;
            RRi[i] = i < CS ? cluster.map_shared_rank(sRR, i)[cl] : 1.0;
        }
        if (MODE == 1) {
            // y-slab mode, first kernel: fold the cluster's CTAs into one rank-level aggregate
            // (same affine composition one level up) and stop; the ranks exchange these.
            double t = 0.0, R = 1.0, X = 0.0, Y = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double Xi = i < CS ? cluster.map_shared_rank(sX, i)[cl] : 0.0;
                const double Yi = i < CS ? cluster.map_shared_rank(sY, i)[cl] : 0.0;
                X = fma(R, fma(Yi, t, Xi), X);
                Y = fma(R, Yi * R, Y);
                t = fma(RRi[i], t, FFi[i]);
                R *= RRi[i];
            }
            if (cr == 0 && slot == 0 && col0 + cl < ncol) {
                a.aggr[0 * ncol + col0 + cl] = t;
                a.aggr[1 * ncol + col0 + cl] = R;
                a.aggr[2 * ncol + col0 + cl] = X;
                a.aggr[3 * ncol + col0 + cl] = Y;
            }
            continue;
        }
        double tt = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tt = fma(RRi[i], tt, FFi[i]);
        // carry into chunk 0: cyclic closure (y at the last row), or handed in by the rank below
        double as = MODE == 2 ? __ldg(a.Ain + gc) : tt * inv1;
        double a_s = 0.0;
        double GGp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double Xi = i < CS ? cluster.map_shared_rank(sX, i)[cl] : 0.0;
            const double Yi = i < CS ? cluster.map_shared_rank(sY, i)[cl] : 0.0;
            GGp[i] = fma(Yi, as, Xi);
            if (i == cr) a_s = as;
            as = fma(RRi[i], as, FFi[i]);
        }
        tt = 0.0;
#pragma unroll
        for (int i = 7; i >= 0; --i) tt = fma(RRi[i], tt, GGp[i]);
        // carry into the last chunk: cyclic closure (z at row 0), or handed in by the rank above
        double b_e = MODE == 2 ? __ldg(a.Bin + gc) : tt * inv1;
#pragma unroll
        for (int i = 7; i > 0; --i)
            if (i > cr) b_e = fma(RRi[i], b_e, GGp[i]);

        const double rho = !act || cc > C - 1 ? 1.0 : __ldg((cc == C - 1 ? a.pl.rhoL : a.pl.rho32) + gc);
        const double hh = !act || cc > C - 1 ? 0.0 : __ldg((cc == C - 1 ? a.pl.hL : a.pl.h32) + gc);
        const double F = act ? sF[slot * TS_LD + cl] : 0.0;
        const double G = act ? sG[slot * TS_LD + cl] : 0.0;
        double pF, Rpre, FFv, RRv;
        fwd_scan(rho, F, pF, Rpre, FFv, RRv);
        const double A = fma(Rpre, a_s, pF);
        const double Gp = fma(A, hh, G);
        // inclusive backward (suffix) scan of x -> rho x + G'
        double R = rho, t = Gp;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const double Rp = __shfl_down_sync(0xffffffffu, R, d, 16);
            const double tp = __shfl_down_sync(0xffffffffu, t, d, 16);
            if (slot + d < 16) {
                t = fma(R, tp, t);
                R *= Rp;
            }
        }
        double Rex = __shfl_down_sync(0xffffffffu, R, 1, 16), tex = __shfl_down_sync(0xffffffffu, t, 1, 16);
        if (slot == 15) { Rex = 1.0; tex = 0.0; }
        if (act) {
            sF[slot * TS_LD + cl] = A;                       // A overwrites F, B overwrites G
            sG[slot * TS_LD + cl] = fma(Rex, b_e, tex);
        }
    }
    K3_STAMP(6);
    cluster.barrier_arrive();   // remote reads are done; matched by barrier_wait() before exit
    if (MODE == 1) {
        cluster.barrier_wait();
        return;
    }
    __syncthreads();

    // ---- pass 2: add the carry, backward recurrence, scale, store --------------------------------
    {
        const double A = mine ? sF[chunk * TS_LD + l] : 0.0, B = mine ? sG[chunk * TS_LD + l] : 0.0;
        const double* __restrict__ k0 = a.k0sol + (int64_t)member * a.preP + a.row0;
        double* __restrict__ out = a.S + member * a.sstride + (int64_t)j0 * ncol + ccol;
        double z = B, v0 = 0.0;
#pragma unroll 4
        for (int i = len - 1; i >= 0; --i) {
            const double y = fma(pw[i * TS_WC + l], A, t[i * TS_WC]);
            z = fma(r, z, y);
            double v = kap * z;
            if (col == 0) v = k0[j0 + i];
            if (cvalid) out[(int64_t)i * ncol] = v;
            v0 = v;   // after the loop: the value at the chunk's first row
        }
        // psi~1 at node (0,0) = sum over the x wavenumbers of row 0 of the solved Poisson field:
        // the half-warp that owns global row 0 leaves this slab's share for K4 to add up
        if (a.gpart != nullptr && a.row0 + j0 == 0 && len > 0) {   // uniform over the half-warp
            double g = cvalid ? __ldg(a.pl.gw + ccol) * v0 : 0.0;
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) g += __shfl_xor_sync(0xffffu << (tid & 16), g, d, 16);
            if (l == 0) a.gpart[(int64_t)member * a.ngp + slab] = g;
        }
    }
    K3_STAMP(7);
    cluster.barrier_wait();   // distributed shared memory must outlive every remote read
}

// ---- persistent, software-pipelined variant (single-GPU mode) ----------------------------------
// Same slab / cluster geometry and carry algebra as k3_ysolve_tma, rebuilt around what the phase
// trace of that kernel showed (QG_K3_TRACE): a dependent FP64 operation costs ~40 cycles here,
// one thread issuing 16 TMA boxes costs ~4 k cycles, cluster.sync() waits for the CTA's global
// stores to drain, and a tile holds its shared memory for the whole CTA lifetime although it is
// in flight for a quarter of it.
//   * A cluster stays on its SMs and walks over the column slabs.  As soon as a CTA's tile has
//     landed it is moved into registers (32 rows per thread) and the TMA copies of the NEXT slab
//     (two 256-row boxes + the slab's column table) are issued into the same buffer: the load of
//     slab n+1 overlaps the sweeps, the exchange and the stores of slab n.
//   * The 32-row recurrences run as four interleaved 8-row segments with a three-step carry fix
//     (12 dependent steps instead of 32); forward then backward, both with zero carries, leave
//     z_loc in registers, F = y_loc(last row), G = z_loc(first row).
//   * Once the carries (A from below, B from above) are known the solution is element-wise,
//         u[i] = kap * ( z_loc[i] + A * cA[i] + B * cB[i] ),
//         cA[i] = r^(i+1) * sum_{m<32-i} r^(2m),   cB[i] = r^(32-i),
//     with cA, cB and every other per-column constant in one table (Plan::coltab) that arrives
//     by TMA with the tile; the stores go straight from registers (a half-warp = one 128-byte line).
//   * The cluster exchange is push-based: every CTA writes its aggregate (FF, RR, X, Y)[16] into
//     the shared memory of all peers with st.async, completing on the peer's mbarrier.  No cluster
//     barrier in the loop, hence no wait for outstanding global stores.
//   * CTA-level closure by shuffles (lane i mod 8 = CTA i), like the chunk level.
constexpr int TP_THREADS = 256;
constexpr int CT_ROWS = 72;   // rows of Plan::coltab, see build_plan()
constexpr int CT_CA = 0, CT_CB = 32, CT_R = 64, CT_KAP = 65, CT_RHO = 66, CT_H = 67, CT_INV1 = 68, CT_PINW = 69,
              CT_GW = 70;

__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_2(uint32_t raddr, double x, double y, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "l"(__double_as_longlong(x)), "l"(__double_as_longlong(y)), "r"(rbar)
                 : "memory");
}

// The singular k = 0 Poisson column inside the persistent kernel (same algorithm as k3_pre, one
// 256-thread CTA, up to 16 rows per thread): writes the solved column straight into column 0 of
// the spectral array, so no separate launch sits between K2 and K3.  Block-uniform call.
__device__ __forceinline__ void k0_column_solve(const YArgs& a, int member, double* sh) {
    constexpr int E = 16;
    const int P = a.pl.P;
    const int e = (P + TP_THREADS - 1) / TP_THREADS;
    const double* __restrict__ col = a.col0 + (int64_t)member * P;
    const int j0 = threadIdx.x * e;
    double b[E];
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int j = j0 + i;
        b[i] = (i < e && j < P) ? __ldcg(col + j) : 0.0;
        loc += b[i];
    }
    const double total = block_sum(loc, sh);
    const double pin = a.pinned ? total : 0.0;
    const double mean = a.pinned ? 0.0 : total / P;
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int j = j0 + i;
        double g = 0.0;
        if (i < e && j < P) g = (b[i] - (j == 0 ? pin : 0.0) - mean) * a.pl.k0scale;
        run += g;
        b[i] = run;   // local inclusive prefix
    }
    const double off1 = block_exclusive_scan(run, sh);
    double s1 = 0.0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int j = j0 + i;
        b[i] += off1;
        if (i < e && j < P) s1 += b[i];
    }
    const double dm1 = -block_sum(s1, sh) / P;
    double run2 = 0.0;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int j = j0 + i;
        const double x = run2;   // exclusive within the thread
        if (i < e && j < P) run2 += dm1 + b[i];
        b[i] = x;
    }
    const double off2 = block_exclusive_scan(run2, sh);
    double* __restrict__ out = a.S + member * a.sstride;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int j = j0 + i;
        if (i < e && j < P) out[(int64_t)j * a.pl.ncol] = b[i] + off2;
    }
}

// sum of the k = 0 Poisson column in a fixed order (the pin's right-hand-side correction)
__device__ __forceinline__ double k0_column_total(const YArgs& a, int member, double* sh) {
    const double* __restrict__ col = a.col0 + (int64_t)member * a.pl.P;
    double loc = 0.0;
    for (int j = threadIdx.x; j < a.pl.P; j += TP_THREADS) loc += __ldcg(col + j);
    return block_sum(loc, sh);
}


#ifdef QG_K3_TRACE
#define K3_ACC(i) do { if (tid == 0) { const long long now_ = clock64(); acc[i] += now_ - last; last = now_; } } while (0)
#else
#define K3_ACC(i) do { } while (0)
#endif

// CSW = lanes of the CTA-level closure = largest cluster size served: 8 (portable) or 16 (non-portable
// clusters, 512 rows x 16 CTAs = 8192 rows in one pass)
// MODE 0: cyclic over the local rows; y-slab modes (YArgs::mode): 1 = rank-level aggregates only, 2 = apply with
// the carries handed in, 3 = aggregates AND the solution with ZERO incoming carries in one pass (the missing
// carry terms decay geometrically away from the slab edges: k3_rank_correct adds them where they matter)
template <int MODE, int CSW>
__global__ void __launch_bounds__(TP_THREADS, 2)
k3_ysolve_pipe(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmT, const YArgs a,
               int nchunk, int boxrows, int nslab, int nwork) {
    extern __shared__ __align__(128) unsigned char ts_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int cr = (int)cluster.block_rank();
    const int ncluster = gridDim.x / CS;
    const int cid = blockIdx.x / CS;
    const int C = a.pl.C, P = a.pl.P, ncol = a.pl.ncol;
    const int tid = threadIdx.x;
    const int chunk = tid >> 4, l = tid & 15;
    const int warp = tid >> 5, lane = tid & 31;
    const int slot = lane & 15, csel = lane >> 4;
    const int cl = 2 * warp + csel;                 // exchange phases: this lane's column within the slab

    double* tile = reinterpret_cast<double*>(ts_raw);                  // [nchunk*32][16]
    double* ctab = tile + (size_t)nchunk * 32 * TS_WC;                 // [2 parities][CT_ROWS][16]
    double* sF = ctab + 2 * CT_ROWS * TS_WC;                           // [nchunk][TS_LD]
    double* sG = sF + nchunk * TS_LD;
    double* sEx = sG + nchunk * TS_LD;                                 // [2 parities][CSW CTAs][16 cols][FF,RR,X,Y]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sEx + 2 * CSW * TS_WC * 4);   // [0] tile, [1..2] exchange parity

    const int c0 = cr * nchunk;                     // first global chunk of this CTA
    const int nact = max(0, min(nchunk, C - c0));   // chunks that hold rows (all of them full: P % 32 == 0)
    const bool live = chunk < nact;                 // this thread owns a chunk with rows
    const int j0 = (c0 + chunk) * 32;
    const double* t = tile + (size_t)(live ? chunk : 0) * 32 * TS_WC + l;
    const bool act = slot < nact;                   // scan lanes that stand for a chunk with rows
    const int nbox = (nchunk * 32) / boxrows;
    const uint32_t tx_bytes = (uint32_t)((nact > 0 ? nchunk * 32 : 0) + CT_ROWS) * TS_WC * sizeof(double);

    auto issue = [&](int w, int par) {   // one thread: tile boxes + column table of work item w
        const int member = w / nslab, slab = w - member * nslab;
        mbar_expect_tx(&bar[0], tx_bytes);
        if (nact > 0)
            for (int k = 0; k < nbox; ++k)
                tma_load_2d(tile + (size_t)k * boxrows * TS_WC, &tmS, slab * TS_WC, member * P + c0 * 32 + k * boxrows,
                            &bar[0]);
        tma_load_2d(ctab + (size_t)par * CT_ROWS * TS_WC, &tmT, slab * TS_WC, 0, &bar[0]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        if (cid < nwork) issue(cid, 0);
    }
    cluster.sync();   // every CTA's mbarriers exist before anyone pushes to them

    // the k = 0 Poisson columns (one per member) are extra work items at the END of the list, so
    // they fall to clusters that would otherwise finish one slab early; the slab pass itself
    // leaves column 0 alone
    __shared__ double red_sh[32];
    const bool k0ext = MODE != 0 || a.k0_external;         // k3_pre has solved the k = 0 column (y-slab mode, P > 4096)
    const int nmember = k0ext ? 0 : nwork / nslab;
    int tot_member = -1;
    double pinscale = 0.0;

#ifdef QG_K3_TRACE
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last = clock64();
#endif
    uint32_t it = 0;
    bool pushed = false;   // this lane stored rank-level aggregates into other ranks' memory (y-slab peer mode)
    for (int w = cid; w < nwork + nmember; w += ncluster, ++it) {
        if (w >= nwork) {   // block-uniform
            if (cr == 0) k0_column_solve(a, w - nwork, red_sh);
            continue;
        }
        const int member = w / nslab, slab = w - member * nslab;
        const int col0 = slab * TS_WC;
        const int col = col0 + l;
        const int par = it & 1;
        const double* ct = ctab + (size_t)par * CT_ROWS * TS_WC;
        double* ex = sEx + (size_t)par * CSW * TS_WC * 4;
        uint64_t* xbar = &bar[1 + par];
        if (a.pinned && c0 == 0 && a.row0 == 0 && member != tot_member) {   // block-uniform: this CTA owns row 0
            pinscale = k0ext ? a.scal[member * 4 + 0] : k0_column_total(a, member, red_sh);
            tot_member = member;
        }
        if (tid == 0) mbar_expect_tx(xbar, (uint32_t)CS * TS_WC * 4 * sizeof(double));
        K3_ACC(0);
        mbar_wait(&bar[0], par);
        K3_ACC(1);
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = live ? t[i * TS_WC] : 0.0;
        K3_ACC(2);

        const double r = ct[CT_R * TS_WC + l];
        {
            // the reference's pinned node: the right-hand side of row 0 loses the column total
            if (live && a.row0 + j0 == 0) v[0] -= ct[CT_PINW * TS_WC + l] * pinscale;
            double p8[8];   // r^(k+1) = cB[31-k]
#pragma unroll
            for (int k = 0; k < 8; ++k) p8[k] = ct[(CT_CB + 31 - k) * TS_WC + l];
            const double r8 = p8[7];
            // forward, zero carry: four interleaved 8-row chains, then the segment carries
#pragma unroll
            for (int k = 1; k < 8; ++k)
#pragma unroll
                for (int sg = 0; sg < 4; ++sg) v[8 * sg + k] = fma(r, v[8 * sg + k - 1], v[8 * sg + k]);
            const double f1 = v[7], f2 = fma(r8, f1, v[15]), f3 = fma(r8, f2, v[23]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[8 + k] = fma(p8[k], f1, v[8 + k]);
                v[16 + k] = fma(p8[k], f2, v[16 + k]);
                v[24 + k] = fma(p8[k], f3, v[24 + k]);
            }
            const double F = v[31];
            // backward, zero carry
#pragma unroll
            for (int k = 6; k >= 0; --k)
#pragma unroll
                for (int sg = 0; sg < 4; ++sg) v[8 * sg + k] = fma(r, v[8 * sg + k + 1], v[8 * sg + k]);
            const double e3 = v[24], e2 = fma(r8, e3, v[16]), e1 = fma(r8, e2, v[8]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = fma(p8[7 - k], e1, v[k]);
                v[8 + k] = fma(p8[7 - k], e2, v[8 + k]);
                v[16 + k] = fma(p8[7 - k], e3, v[16 + k]);
            }
            if (chunk < nchunk) {
                sF[chunk * TS_LD + l] = F;
                sG[chunk * TS_LD + l] = v[0];
            }
        }
        __syncthreads();
        // The tile is in registers AND every load of it has delivered its value: the stores of F and G above
        // depend on all 32 of them, and a store cannot pass the barrier.  Only now may the next slab's TMA copies
        // overwrite the buffer.  (Round 1 issued them right behind a barrier that directly followed the 32 loads:
        // a barrier orders the loads' ISSUE, not their completion, and the TMA engine does not go through the
        // load/store pipe that keeps generic accesses in order - about one run in ten of 43 steps at 4096^2 then
        // differed in the last bits, scripts/determinism.py; the per-slab first-generation kernel never did.)
        if (tid == 0 && w + ncluster < nwork) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(w + ncluster, par ^ 1);
        }
        K3_ACC(3);

        // ---- carries: warp-level cyclic reduction per column pair, one push-based exchange ----------
        {
            const double rho = act ? ct[CT_RHO * TS_WC + cl] : 1.0;
            const double hh = act ? ct[CT_H * TS_WC + cl] : 0.0;
            const double inv1 = ct[CT_INV1 * TS_WC + cl];
            double asIn = 0.0, beIn = 0.0;
            if (MODE == 2) {
                const int gc = (col0 + cl < ncol) ? col0 + cl : 0;
                asIn = __ldg(a.Ain + gc);
                beIn = __ldg(a.Bin + gc);
            }
            const double F = act ? sF[slot * TS_LD + cl] : 0.0;
            const double G = act ? sG[slot * TS_LD + cl] : 0.0;
            double R = rho, tt = F;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                const double Rp = __shfl_up_sync(0xffffffffu, R, d, 16);
                const double tp = __shfl_up_sync(0xffffffffu, tt, d, 16);
                if (slot >= d) {
                    tt = fma(R, tp, tt);
                    R *= Rp;
                }
            }
            const double FFv = __shfl_sync(0xffffffffu, tt, 15, 16);
            const double RRv = __shfl_sync(0xffffffffu, R, 15, 16);
            double Rpre = __shfl_up_sync(0xffffffffu, R, 1, 16);
            double pF = __shfl_up_sync(0xffffffffu, tt, 1, 16);
            if (slot == 0) { Rpre = 1.0; pF = 0.0; }
            double Xv = Rpre * fma(hh, pF, G), Yv = Rpre * (hh * Rpre);
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) {
                Xv += __shfl_xor_sync(0xffffffffu, Xv, d, 16);
                Yv += __shfl_xor_sync(0xffffffffu, Yv, d, 16);
            }
            if (slot < CS) {   // lane `slot` pushes this column's aggregate to CTA `slot`
                const uint32_t dst = map_to_cta(smem_u32(ex + ((size_t)cr * TS_WC + cl) * 4), (uint32_t)slot);
                const uint32_t rb = map_to_cta(smem_u32(xbar), (uint32_t)slot);
                st_async_2(dst, FFv, RRv, rb);
                st_async_2(dst + 16, Xv, Yv, rb);
            }
            K3_ACC(4);
            mbar_wait(xbar, (it >> 1) & 1);
            K3_ACC(5);
            // CTA-level closure, again as a parallel cyclic reduction: lane i (mod 8) of the half-warp
            // holds CTA i's aggregate and the affine maps compose by shuffles
            const int i8 = slot & (CSW - 1);
            double FFl = 0.0, RRl = 1.0, Xl = 0.0, Yl = 0.0;
            if (i8 < CS) {
                const double2 q0 = *reinterpret_cast<const double2*>(ex + ((size_t)i8 * TS_WC + cl) * 4);
                const double2 q1 = *reinterpret_cast<const double2*>(ex + ((size_t)i8 * TS_WC + cl) * 4 + 2);
                FFl = q0.x; RRl = q0.y; Xl = q1.x; Yl = q1.y;
            }
            double R1 = RRl, T1 = FFl;   // inclusive forward composition over CTAs 0..i
#pragma unroll
            for (int d = 1; d < CSW; d <<= 1) {
                const double Rp = __shfl_up_sync(0xffffffffu, R1, d, CSW);
                const double Tp = __shfl_up_sync(0xffffffffu, T1, d, CSW);
                if (i8 >= d) {
                    T1 = fma(R1, Tp, T1);
                    R1 *= Rp;
                }
            }
            double Rx = __shfl_up_sync(0xffffffffu, R1, 1, CSW), Tx = __shfl_up_sync(0xffffffffu, T1, 1, CSW);
            if (i8 == 0) { Rx = 1.0; Tx = 0.0; }
            if (MODE == 1 || MODE == 3) {
                // y-slab mode: fold the cluster's CTAs into one rank-level aggregate (same affine
                // composition one level up); the ranks exchange these.  Mode 1 stops here.
                double Xr = Rx * fma(Yl, Tx, Xl), Yr = Rx * (Yl * Rx);
#pragma unroll
                for (int d = CSW / 2; d > 0; d >>= 1) {
                    Xr += __shfl_xor_sync(0xffffffffu, Xr, d, CSW);
                    Yr += __shfl_xor_sync(0xffffffffu, Yr, d, CSW);
                }
                const double Tt = __shfl_sync(0xffffffffu, T1, CSW - 1, CSW), Rt = __shfl_sync(0xffffffffu, R1, CSW - 1, CSW);
                if (cr == 0 && slot == 0 && col0 + cl < ncol) {
                    a.aggr[0 * ncol + col0 + cl] = Tt;
                    a.aggr[1 * ncol + col0 + cl] = Rt;
                    a.aggr[2 * ncol + col0 + cl] = Xr;
                    a.aggr[3 * ncol + col0 + cl] = Yr;
                }
                // peer mode: lane r of the half-warp pushes the aggregate into rank r's aggr_all[this rank]
                if (cr == 0 && slot < a.peer_n && col0 + cl < ncol) {
                    double* dst = a.aggr_peer[slot] + (size_t)a.peer_rank * 4 * ncol + col0 + cl;
                    dst[0] = Tt;
                    dst[ncol] = Rt;
                    dst[2 * ncol] = Xr;
                    dst[3 * ncol] = Yr;
                    pushed = true;
                }
                if (MODE == 1) continue;   // warp-uniform; sF / sG are rewritten only after the next iteration's barrier
            }
            // carry into CTA 0: cyclic closure (y at the last row), or handed in by the rank below
            const double as0 = MODE >= 2 ? asIn : __shfl_sync(0xffffffffu, T1, CSW - 1, CSW) * inv1;
            const double as_i = fma(Rx, as0, Tx);          // forward carry into CTA i
            const double GGp = fma(Yl, as_i, Xl);          // CTA i's backward aggregate with its true carry
            double R2 = RRl, T2 = GGp;   // inclusive backward composition over CTAs 7..i
#pragma unroll
            for (int d = 1; d < CSW; d <<= 1) {
                const double Rp = __shfl_down_sync(0xffffffffu, R2, d, CSW);
                const double Tp = __shfl_down_sync(0xffffffffu, T2, d, CSW);
                if (i8 + d < CSW) {
                    T2 = fma(R2, Tp, T2);
                    R2 *= Rp;
                }
            }
            // carry into the last CTA: cyclic closure (z at row 0), or handed in by the rank above
            const double blast = MODE >= 2 ? beIn : __shfl_sync(0xffffffffu, T2, 0, CSW) * inv1;
            Rx = __shfl_down_sync(0xffffffffu, R2, 1, CSW);
            Tx = __shfl_down_sync(0xffffffffu, T2, 1, CSW);
            if (i8 == CSW - 1) { Rx = 1.0; Tx = 0.0; }
            const double be_i = fma(Rx, blast, Tx);        // backward carry into CTA i
            const double a_s = __shfl_sync(0xffffffffu, as_i, cr, CSW);
            const double b_e = __shfl_sync(0xffffffffu, be_i, cr, CSW);
            const double A = fma(Rpre, a_s, pF);
            const double Gp = fma(A, hh, G);
            // inclusive backward (suffix) scan of x -> rho x + G'
            double ts = Gp;
            R = rho;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                const double Rp = __shfl_down_sync(0xffffffffu, R, d, 16);
                const double tp = __shfl_down_sync(0xffffffffu, ts, d, 16);
                if (slot + d < 16) {
                    ts = fma(R, tp, ts);
                    R *= Rp;
                }
            }
            double Rex = __shfl_down_sync(0xffffffffu, R, 1, 16), tex = __shfl_down_sync(0xffffffffu, ts, 1, 16);
            if (slot == 15) { Rex = 1.0; tex = 0.0; }
            if (slot < nchunk) {
                sF[slot * TS_LD + cl] = A;                       // A overwrites F, B overwrites G
                sG[slot * TS_LD + cl] = fma(Rex, b_e, tex);
            }
        }
        __syncthreads();
        K3_ACC(6);

        // ---- apply: element-wise from registers, a half-warp stores one 128-byte row segment ---------
        if (live) {
            const double A = sF[chunk * TS_LD + l], B = sG[chunk * TS_LD + l];
            const double kap = ct[CT_KAP * TS_WC + l];
            const bool cvalid = col < ncol;
            // column 0 (k = 0 Poisson): written by k0_column_solve (mode 0) / taken from k3_pre's solution (mode 2)
            const bool wr = col < ncol && (k0ext || col != 0);
            const double* __restrict__ k0 = a.k0sol + (int64_t)member * a.preP + a.row0 + j0;
            double* out = a.S + member * a.sstride + (int64_t)j0 * ncol + (cvalid ? col : 0);
            double u0 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                double u = kap * fma(A, ct[(CT_CA + i) * TS_WC + l], fma(B, ct[(CT_CB + i) * TS_WC + l], v[i]));
                if (k0ext && col == 0) u = k0[i];
                if (wr) *out = u;
                out += ncol;
                if (i == 0) u0 = u;   // kap = 0 for the singular column: its gauge share is its solved value 0
            }
            // psi~1 at node (0,0) = sum over the x wavenumbers of row 0 of the solved Poisson field:
            // the half-warp that owns global row 0 leaves this slab's share for K4 to add up
            if (a.gpart != nullptr && a.row0 + j0 == 0) {   // uniform over the half-warp
                double g = ct[CT_GW * TS_WC + l] * u0;
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) g += __shfl_xor_sync(0xffffu << (tid & 16), g, d, 16);
                if (l == 0) a.gpart[(int64_t)member * a.ngp + slab] = g;
            }
        }
        K3_ACC(7);
        // sF / sG and the other table parity are rewritten only after the next iteration's barriers
    }
#ifdef QG_K3_TRACE
    if (tid == 0 && g_k3_trace)
        for (int i = 0; i < 8; ++i) g_k3_trace[(size_t)blockIdx.x * 8 + i] = acc[i];
#endif
    // posted NVLink writes: acknowledged before the lane exits, so that the flag barrier enqueued behind this
    // kernel cannot overtake them (one system-scope fence covers all the lane's earlier stores; see k2_fft.cu)
    if (pushed) __threadfence_system();
    cluster.sync();   // no CTA leaves while a peer may still push to it or read its aggregates
}

// y-slab mode: cyclic closure over the ranks.  aggr_all[g][4][ncol] holds every rank's
// (FF, RR, X, Y); thread per column computes the forward carry entering this rank from below
// (Ain) and the backward carry entering it from above (Bin).  Same algebra as the CTA level.
__global__ void __launch_bounds__(256)
k3_rank_closure(const double* __restrict__ aggr_all, int nranks, int rank, int ncol,
                const double* __restrict__ inv1mrP, double* __restrict__ Ain, double* __restrict__ Bin) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double FFi[8], RRi[8], Xi[8], Yi[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const bool in = g < nranks;
        const double* p = aggr_all + (size_t)(in ? g : 0) * 4 * ncol + col;
        FFi[g] = in ? p[0] : 0.0;
        RRi[g] = in ? p[ncol] : 1.0;
        Xi[g] = in ? p[2 * ncol] : 0.0;
        Yi[g] = in ? p[3 * ncol] : 0.0;
    }
    const double inv1 = inv1mrP[col];
    double tt = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) tt = fma(RRi[g], tt, FFi[g]);
    double as = tt * inv1, a_s = 0.0;
    double GGp[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        GGp[g] = fma(Yi[g], as, Xi[g]);
        if (g == rank) a_s = as;
        as = fma(RRi[g], as, FFi[g]);
    }
    tt = 0.0;
#pragma unroll
    for (int g = 7; g >= 0; --g) tt = fma(RRi[g], tt, GGp[g]);
    double b_e = tt * inv1;
#pragma unroll
    for (int g = 7; g > 0; --g)
        if (g > rank) b_e = fma(RRi[g], b_e, GGp[g]);
    Ain[col] = a_s;
    Bin[col] = b_e;
}

// y-slab mode, single-pass variant.  k3_ysolve_pipe<3> has written u_loc, the solution with zero carries
// entering the rank.  With the true carries Ain (forward, from the rank below) and Bin (backward, from the
// rank above) the two first-order recurrences give, for local row i of P rows,
//     u[i] = u_loc[i] + kap * ( Ain * r^(i+1) * (1 - r^(2(P-i))) / (1 - r^2)  +  Bin * r^(P-i) ),
// and since |r| < 1 both terms vanish (below 2^-60 of their size at the edge) more than n_cut = 41.6 / -ln r
// rows away from the bottom / top edge: 22 rows at the shortest waves, every row only for the few longest
// ones.  A block owns 32 adjacent columns (one warp row = 256 B), computes their carries from all ranks'
// aggregates (the cyclic closure of k3_rank_closure) and its warps walk over the 32-row segments that still
// matter, with r^n re-evaluated from ln r at every segment start.  On 16384 x 1024 rows per rank this
// touches 13 % of the slab instead of re-reading and re-solving all of it.  (Two details that mattered: the 32
// row updates of a segment are 32 independent loads, then the arithmetic, then 32 stores - as a read-modify-
// write loop the compiler has to order every load behind the previous store and the kernel took 550 us; and
// only the (tile, segment) pairs that need work are launched - a full grid of mostly idle blocks took 160 us.)
__global__ void __launch_bounds__(256)
k3_rank_correct(double* __restrict__ S, int ncol, int P, const double* __restrict__ aggr_all, int nranks, int rank,
                const double* __restrict__ inv1mrP, const double* __restrict__ rtab, const double* __restrict__ kaptab,
                const double* __restrict__ logr, const double* __restrict__ g1mr2, const int2* __restrict__ work,
                int nwork) {
    // one warp per work item = (32-column tile, 32-row segment) that an edge term still reaches; the list
    // depends on the plan only and is built with it (build_plan)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * 8 + warp;
    if (item >= nwork) return;
    const int2 wk = work[item];
    const int col = wk.x * 32 + lane;
    const bool cv = col < ncol;
    const int cc = cv ? col : 0;
    const double r = cv ? rtab[cc] : 0.0, kap = cv ? kaptab[cc] : 0.0;
    const double lr = r > 0.0 ? logr[cc] : 0.0;
    const int i0 = wk.y << 5;

    // the carries entering this rank: cyclic closure over all ranks' aggregates (k3_rank_closure)
    double FFi[8], RRi[8], Xi[8], Yi[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const bool in = g < nranks;
        const double* p = aggr_all + (size_t)(in ? g : 0) * 4 * ncol + cc;
        FFi[g] = in ? p[0] : 0.0;
        RRi[g] = in ? p[ncol] : 1.0;
        Xi[g] = in ? p[2 * (size_t)ncol] : 0.0;
        Yi[g] = in ? p[3 * (size_t)ncol] : 0.0;
    }
    const double inv1 = inv1mrP[cc];
    double tt = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) tt = fma(RRi[g], tt, FFi[g]);
    double as = tt * inv1, Ain = 0.0;
    double GGp[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        GGp[g] = fma(Yi[g], as, Xi[g]);
        if (g == rank) Ain = as;
        as = fma(RRi[g], as, FFi[g]);
    }
    tt = 0.0;
#pragma unroll
    for (int g = 7; g >= 0; --g) tt = fma(RRi[g], tt, GGp[g]);
    double Bin = tt * inv1;
#pragma unroll
    for (int g = 7; g > 0; --g)
        if (g > rank) Bin = fma(RRi[g], Bin, GGp[g]);

    const double gg = r > 0.0 ? g1mr2[cc] : 0.0, rinv = r > 0.0 ? 1.0 / r : 0.0;
    const double cA = kap * Ain * gg, cB = kap * Bin;
    double pA = r > 0.0 ? exp((double)(i0 + 1) * lr) : 0.0, pB = r > 0.0 ? exp((double)(P - i0) * lr) : 0.0;
    double* __restrict__ p = S + (int64_t)i0 * ncol + cc;
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = cv ? p[(int64_t)i * ncol] : 0.0;   // 32 independent loads in flight
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        v[i] += fma(cA * pA, 1.0 - pB * pB, cB * pB);
        pA *= r;
        pB *= rinv;
    }
    if (cv) {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[(int64_t)i * ncol] = v[i];
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static cudaError_t launch_tma_kernel(Handle* h, const YArgs& a) {
    const Plan& pl = h->plan;
    static const int use_v1 = env_int("QG_K3_V1", 0);
    static const int ncl_env = env_int("QG_K3_NCL", 0);
    const int nslab = (pl.ncol + TS_WC - 1) / TS_WC;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nslab * pl.ts_CS, h->nm, 1);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.ts_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (use_v1 || !pl.tp_ok) {
        cfg.blockDim = dim3(((16 * pl.ts_nchunk + 31) / 32) * 32, 1, 1);
        cfg.dynamicSmemBytes = ((size_t)pl.ts_nchunk * 32 * TS_WC + 32 * TS_WC + 2 * pl.ts_nchunk * TS_LD +
                                4 * TS_WC) * sizeof(double) + 16;
        static size_t configured_dev[QG_MAX_DEVICES][3] = {};
        size_t* configured = configured_dev[dev_slot(h)];
        auto kern = a.mode == 1 ? k3_ysolve_tma<1> : (a.mode == 2 ? k3_ysolve_tma<2> : k3_ysolve_tma<0>);
        if (cfg.dynamicSmemBytes > configured[a.mode]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)cfg.dynamicSmemBytes);
            if (e != cudaSuccess) return e;
            configured[a.mode] = cfg.dynamicSmemBytes;
        }
        KernelTimer t(h, QG_K_YSOLVE);
        return cudaLaunchKernelEx(&cfg, kern, h->tm_S, a, pl.ts_nchunk);
    }
    // persistent clusters, one wave; clusters of up to 8 CTAs are portable, 16 (P > 4096) opt in
    const bool wide = pl.tp_CS > 8;
    const int csw = wide ? 16 : 8;
    static size_t configured_p_dev[QG_MAX_DEVICES][2] = {};
    size_t* configured_p = configured_p_dev[dev_slot(h)];
    attr[0].val.clusterDim.x = pl.tp_CS;
    cfg.blockDim = dim3(TP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = ((size_t)pl.tp_nchunk * 32 * TS_WC + 2 * CT_ROWS * TS_WC + 2 * pl.tp_nchunk * TS_LD +
                            2 * csw * TS_WC * 4) * sizeof(double) + 32;
    auto pkern = wide ? (a.mode == 1 ? k3_ysolve_pipe<1, 16> : (a.mode == 2 ? k3_ysolve_pipe<2, 16> :
                                                               (a.mode == 3 ? k3_ysolve_pipe<3, 16> : k3_ysolve_pipe<0, 16>)))
                      : (a.mode == 1 ? k3_ysolve_pipe<1, 8> : (a.mode == 2 ? k3_ysolve_pipe<2, 8> :
                                                               (a.mode == 3 ? k3_ysolve_pipe<3, 8> : k3_ysolve_pipe<0, 8>)));
    if (cfg.dynamicSmemBytes > configured_p[wide]) {
        auto set = [&](auto k) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
            if (e == cudaSuccess && wide) e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            return e;
        };
        cudaError_t e = wide ? set(k3_ysolve_pipe<0, 16>) : set(k3_ysolve_pipe<0, 8>);
        if (e == cudaSuccess) e = wide ? set(k3_ysolve_pipe<1, 16>) : set(k3_ysolve_pipe<1, 8>);
        if (e == cudaSuccess) e = wide ? set(k3_ysolve_pipe<2, 16>) : set(k3_ysolve_pipe<2, 8>);
        if (e == cudaSuccess) e = wide ? set(k3_ysolve_pipe<3, 16>) : set(k3_ysolve_pipe<3, 8>);
        if (e != cudaSuccess) return e;
        configured_p[wide] = cfg.dynamicSmemBytes;
    }
    if (h->plan.tp_ncl == 0) {   // clusters the device holds at once
        int ncl = 0;
        cfg.gridDim = dim3(nslab * pl.tp_CS * h->nm, 1, 1);
        if (cudaOccupancyMaxActiveClusters(&ncl, pkern, &cfg) == cudaSuccess && ncl > 0) h->plan.tp_ncl = ncl;
        else { h->plan.tp_ncl = 2 * 148 / pl.tp_CS; (void)cudaGetLastError(); }
        if (ncl_env > 0) h->plan.tp_ncl = ncl_env;
        if (getenv("QG_VERBOSE")) fprintf(stderr, "qgb200: y-solve persistent clusters: %d x %d CTAs\n", h->plan.tp_ncl, pl.tp_CS);
    }
    const int nwork = nslab * h->nm;
    // work items = slabs + one k=0 column per member; small grids get clusters of their own for the latter
    const int nitems = nwork + (a.mode == 0 ? h->nm : 0);
    const int ncl = h->plan.tp_ncl < nitems ? h->plan.tp_ncl : nitems;
    cfg.gridDim = dim3(ncl * pl.tp_CS, 1, 1);
#ifdef QG_K3_TRACE
    static long long* dbuf = nullptr;
    static int calls = 0;
    const size_t nct = (size_t)cfg.gridDim.x;
    if (!dbuf) {
        cudaMalloc((void**)&dbuf, nct * 8 * sizeof(long long));
        cudaMemcpyToSymbol(g_k3_trace, &dbuf, sizeof(dbuf));
    }
    cudaError_t te;
    {
        KernelTimer t(h, QG_K_YSOLVE);
        te = cudaLaunchKernelEx(&cfg, pkern, h->tm_S2, h->tm_T, a, pl.tp_nchunk, pl.tp_boxrows, nslab, nwork);
    }
    if (++calls == 20) {
        cudaStreamSynchronize(h->stream);
        std::vector<long long> hb(nct * 8);
        cudaMemcpy(hb.data(), dbuf, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum[8] = {0};
        for (size_t c = 0; c < nct; ++c)
            for (int i = 0; i < 8; ++i) sum[i] += (double)hb[c * 8 + i];
        const char* nm[8] = {"loop top", "tma wait", "tile->regs+sync+issue", "sweeps+sync", "publish scan+push", "exchange wait",
                             "closure+sync", "apply+stores"};
        const double iters = (double)nwork / ncl;
        double tot = 0;
        for (int i = 0; i < 8; ++i) tot += sum[i] / nct / iters;
        for (int i = 0; i < 8; ++i) fprintf(stderr, "K3TRACE %-22s %9.0f cycles/iter\n", nm[i], sum[i] / nct / iters);
        fprintf(stderr, "K3TRACE %-22s %9.0f cycles/iter (%zu CTAs, %.1f iters)\n", "total", tot, nct, iters);
    }
    return te;
#else
    KernelTimer t(h, QG_K_YSOLVE);
    return cudaLaunchKernelEx(&cfg, pkern, h->tm_S2, h->tm_T, a, pl.tp_nchunk, pl.tp_boxrows, nslab, nwork);
#endif
}

// y-slab mode: gather the k=0 column, solve it redundantly on every rank, sweep + exchange the
// rank-level carry aggregates, close the recurrences over the ranks, apply.
static cudaError_t launch_ysolve_dist(Handle* h, int pinned) {
    h->gauge_parts = false;   // rank 0 evaluates the gauge (k3_gauge) and hands it to the others
    YArgs a{};
    a.pl = h->plan;
    a.S = h->S;
    a.sstride = (int64_t)h->plan.P * h->plan.ncol;
    a.scal = h->scal;
    a.pinned = pinned;
    a.row0 = h->dist_rank * h->plan.P;
    a.preP = h->Pglob;
    a.col0 = h->col0_full;
    a.k0sol = h->k0sol_full;
    a.aggr = h->aggr;
    a.Ain = h->carry_in;
    a.Bin = h->carry_in + h->plan.ncol;
    static const bool force_v1 = env_int("QG_K3_V1", 0) != 0;
    const bool peer = h->peer_ok && h->plan.tp_ok && !force_v1;   // exchanges by in-kernel peer stores + flag barriers
    if (peer) {
        a.peer_n = h->dist_n;
        a.peer_rank = h->dist_rank;
        for (int r = 0; r < h->dist_n; ++r) {
            a.aggr_peer[r] = h->peer_mail[r] + (h->aggr_all - h->mailbox);
            a.scal_peer[r] = h->peer_mail[r] + (h->scal - h->mailbox);
        }
    }
    // K2 has written the k=0 column: gathered by NCCL, or already in place on every rank (peer stores)
    cudaError_t e = peer ? dist_barrier(h) : dist_allgather(h, h->col0, h->col0_full, (size_t)h->plan.P);
    if (e != cudaSuccess) return e;
    {
        KernelTimer t(h, QG_K_YPRE);
        launch_pre(a, 1, h->stream);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // Single pass (default): solve with zero incoming carries while the rank-level aggregates travel, then add
    // the carry terms where they have not decayed (k3_rank_correct).  QG_K3_TWOPASS=1 (or the first-generation
    // kernel) keeps the round-1 flow: aggregates, closure, second full pass.
    static const bool twopass = env_int("QG_K3_TWOPASS", 0) != 0;
    const bool single = !twopass && !force_v1 && h->plan.tp_ok && h->plan.corr_work != nullptr;
    a.mode = single ? 3 : 1;
    if ((e = launch_tma_kernel(h, a)) != cudaSuccess) return e;
    e = peer ? dist_barrier(h) : dist_allgather(h, h->aggr, h->aggr_all, (size_t)4 * h->plan.ncol);
    if (e != cudaSuccess) return e;
    if (single) {
        KernelTimer t(h, QG_K_GAUGE);
        k3_rank_correct<<<(h->plan.ncorr + 7) / 8, 256, 0, h->stream>>>(
            h->S, h->plan.ncol, h->plan.P, h->aggr_all, h->dist_n, h->dist_rank, h->plan.inv1mrP, h->plan.rtab,
            h->plan.kap, h->plan.logr, h->plan.g1mr2, h->plan.corr_work, h->plan.ncorr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    } else {
        {
            KernelTimer t(h, QG_K_GAUGE);
            k3_rank_closure<<<(h->plan.ncol + 255) / 256, 256, 0, h->stream>>>(
                h->aggr_all, h->dist_n, h->dist_rank, h->plan.ncol, h->plan.inv1mrP, h->carry_in,
                h->carry_in + h->plan.ncol);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        a.mode = 2;
        if ((e = launch_tma_kernel(h, a)) != cudaSuccess) return e;
    }
    if (h->dist_rank == 0) {   // global row 0 lives on rank 0
        KernelTimer t(h, QG_K_GAUGE);
        k3_gauge<<<1, 256, 0, h->stream>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return peer ? dist_barrier(h) : dist_broadcast(h, h->scal, 4, 0);
}

cudaError_t launch_ysolve(Handle* h, int pinned, int /*unused*/) {
    YArgs a{};
    a.pl = h->plan;
    a.S = h->S;
    a.sstride = (int64_t)h->plan.P * h->plan.ncol;
    if (h->dist_n > 1) return launch_ysolve_dist(h, pinned);
    a.k0sol = h->k0sol;
    a.scal = h->scal;
    a.pinned = pinned;
    a.col0 = h->col0;
    a.preP = h->plan.P;
    static const bool force_v1 = env_int("QG_K3_V1", 0) != 0;
    const bool pipe = h->plan.tp_ok && !force_v1;          // persistent kernel
    const bool tma = pipe || h->plan.ts_ok;                // either TMA-staged kernel
    h->gauge_parts = tma;
    if (tma) {   // the gauge is assembled from per-slab partial sums (no extra kernel)
        a.gpart = h->gpart;
        a.ngp = h->plan.ngp;
    }
    // the persistent kernel solves the k = 0 column itself up to 4096 rows (16 per thread of one CTA)
    a.k0_external = (pipe && h->plan.P > 16 * TP_THREADS) ? 1 : 0;
    if (!pipe || a.k0_external) {
        KernelTimer t(h, QG_K_YPRE);
        launch_pre(a, h->nm, h->stream);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (tma) {
        e = launch_tma_kernel(h, a);
    } else {
        const Plan& pl = h->plan;
        const int nslab = (pl.ncol + 31) / 32;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(nslab * pl.CS, h->nm, 1);
        cfg.blockDim = dim3(32 * pl.wpc, 1, 1);
        cfg.dynamicSmemBytes = (size_t)(4 * pl.wpc * pl.m + 4) * 32 * sizeof(double);
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
    if (!tma) {
        KernelTimer t(h, QG_K_GAUGE);
        k3_gauge<<<h->nm, 256, 0, h->stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace qg
