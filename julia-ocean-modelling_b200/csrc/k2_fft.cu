// K2 / K4 — batched FFT along the periodic x direction, fused with the modal projections.
//
// Together with k3_ysolve.cu these kernels replace the two sparse Cholesky solves of
// evolve_psi! (reference src/model.jl:179-198; factors built at
// src/schemes/laplacian.jl:54-75).
//
//   K2 (forward):  q~ = P_inv (q1, q2)            (src/model.jl:179-182)
//                  z[n] = q~1[n] + i q~2[n]       two real rows as one complex row
//                  Z = FFT_M(z), untangled in place into the packed spectral row
//                  slot k = Q1[k], slot M-k = Q2[k]  (see Plan in qg_internal.cuh)
//   K4 (inverse):  re-tangle, inverse FFT, gauge shift psi~1 -= psi~1(0,0)
//                  (the reference's pinned unknown, src/schemes/laplacian.jl:71-73),
//                  psi = P (psi~1, psi~2) (src/model.jl:195-198), periodic ghost images
//                  (src/schemes/boundary_conditions.jl:2-22).
//
// The transform is a hand-written Stockham autosort FFT: 8 points per thread, radix-8
// passes (plus one radix-4/2 pass when log2 M is not a multiple of 3), data exchanged
// through XOR-swizzled shared memory between passes only: the first pass is fed straight
// from global memory and the inverse transform's last pass stores straight to global
// memory (its outputs are coalesced by construction).  The kernels are persistent: a CTA
// loops over rows, and each thread keeps one base twiddle per pass in registers for the
// whole launch (the twiddle depends on the thread's position, not on the row) and derives
// the powers w^2..w^7 by multiplication, so the row loop issues no twiddle loads at all.
// The base twiddles come from a table evaluated in extended precision on the host.
// Non-power-of-two M (the reference benchmarks M = 8:8:128) takes a direct O(M^2) DFT
// path with the same spectral layout.
#include <cstdio>
#include <cstdlib>

#include "qg_internal.cuh"

namespace qg {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ double2 muli(double2 a) {
    return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
template <int SIGN>
__device__ __forceinline__ double2 twid(const double2* __restrict__ tw, int idx) {
    double2 w = __ldg(tw + idx);
    if (SIGN > 0) w.y = -w.y;   // table holds exp(-2 pi i n / M)
    return w;
}

// Q1[0] of one row: the compact k=0 Poisson column.  In y-slab peer mode the value goes straight
// into every rank's gathered column (NVLink peer stores) instead of a local copy + all-gather.
// `pushed` is set when peer stores were issued: the calling kernel ends with peer_store_fence(pushed).
__device__ __forceinline__ void store_col0(const FftArgs& a, int member, int row, double v, bool& pushed) {
    if (a.col0_n > 0) {
        for (int r = 0; r < a.col0_n; ++r) a.col0_peer[r][a.col0_off + row] = v;
        pushed = true;
    } else {
        a.col0[(int64_t)member * a.pl.P + row] = v;
    }
}

// Peer stores are posted writes over NVLink.  The thread that issued them waits for their acknowledgement
// (one system-scope fence, which covers all its earlier stores) before it exits, so that the flag barrier
// enqueued behind the kernel can never overtake them: the barrier kernel's own fence orders only its own
// thread's writes.  Without this an 8-rank run of 2048 x 4096 departed from the oracle by 7e-6 (stale halo rows
// at the ring's seam), profiles/r02/slab_r02s_n8.log -> slab_r02u_n8.log.
__device__ __forceinline__ void peer_store_fence(bool pushed) {
    if (pushed) __threadfence_system();
}

// bank-conflict-free placement of complex slot i (16-byte elements)
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }

// 4-point DFT, X[u] = sum_t c[t] exp(SIGN * 2 pi i t u / 4)
template <int SIGN>
__device__ __forceinline__ void bfly4(double2& c0, double2& c1, double2& c2, double2& c3) {
    const double2 s02 = cadd(c0, c2), d02 = csub(c0, c2);
    const double2 s13 = cadd(c1, c3), d13 = muli<SIGN>(csub(c1, c3));
    c0 = cadd(s02, s13);
    c2 = csub(s02, s13);
    c1 = cadd(d02, d13);
    c3 = csub(d02, d13);
}

// 8-point DFT in place, natural order output.
template <int SIGN>
__device__ __forceinline__ void bfly8(double2 (&a)[8]) {
    const double h = 0.70710678118654752440;
    double2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    double2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    double2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
    double2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    {
        const double2 t5 = muli<SIGN>(b5);   // W8^1 = h (1 + SIGN i)
        b5 = make_double2(h * (b5.x + t5.x), h * (b5.y + t5.y));
        b6 = muli<SIGN>(b6);                 // W8^2 = SIGN i
        const double2 t7 = muli<SIGN>(b7);   // W8^3 = h (-1 + SIGN i)
        b7 = make_double2(h * (t7.x - b7.x), h * (t7.y - b7.y));
    }
    bfly4<SIGN>(b0, b1, b2, b3);   // X[0], X[2], X[4], X[6]
    bfly4<SIGN>(b4, b5, b6, b7);   // X[1], X[3], X[5], X[7]
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = b4; a[3] = b5; a[5] = b6; a[7] = b7;
}

// Per-thread FFT engine for rows of N = 2^LOG2N points, 8 points per thread.
// W8S: keep the per-pass base twiddles in shared memory instead of registers (the 1024-thread
// long-row kernels have 64 registers per thread and spilled with them resident)
template <int LOG2N, int SIGN, bool W8S = false>
struct RowFft {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPR = N / 8;
    static constexpr int NB8 = LOG2N / 3;
    static constexpr int REM = LOG2N % 3;
    static constexpr int NW8 = NB8 > 1 ? NB8 - 1 : 1;
    double2 w8[W8S ? 1 : NW8];   // base twiddle of radix-8 pass p = 1 .. NB8-1 (registers)
    double2* w8s;                // ... or [NW8][blockDim.x] in shared memory, this thread's column
    double2 wr0;       // base twiddle of the remainder pass for butterfly 0 (index lt); butterfly b sits
                       // TPR = N/8 further on, i.e. its twiddle is wr0 turned by b eighths of a half turn
    // wr0 * exp(SIGN * i * pi * b / 4)
    __device__ __forceinline__ double2 wrem(int b) const {
        const double h = 0.70710678118654752440;
        switch (b & 3) {
            case 0: return wr0;
            case 1: return cmul(wr0, make_double2(h, SIGN * h));
            case 2: return muli<SIGN>(wr0);
            default: return cmul(wr0, make_double2(-h, SIGN * h));
        }
    }

    // `stride`: the table holds exp(-2 pi i n / (stride * N))
    __device__ __forceinline__ void init(const double2* __restrict__ tw, int lt, int stride = 1,
                                         double2* w8_smem = nullptr) {
        int Ns = 8;
        w8s = w8_smem + threadIdx.x;
#pragma unroll
        for (int p = 1; p < NB8; ++p) {
            const int k = lt & (Ns - 1);
            const double2 w = twid<SIGN>(tw, stride * (k * (N / (Ns * 8))));
            if (W8S) w8s[(p - 1) * blockDim.x] = w; else w8[p - 1] = w;   // only this thread reads it back
            Ns *= 8;
        }
        // remainder pass: Ns = N/4 (radix 4) or N/2 (radix 2), so j = lt + b*TPR < Ns and the table
        // index of butterfly b is simply j
        if (REM != 0) wr0 = twid<SIGN>(tw, stride * lt);
    }

    // On entry v[t] = x[lt + t*TPR].  TO_SMEM: on exit the transform sits in `s` (swizzled,
    // natural order) after a __syncthreads().  Otherwise the last pass stays in registers:
    //   REM == 0: v[u]       = X[lt + u*TPR]
    //   REM == 2: v[4b + u]  = X[lt + b*TPR + u*N/4]
    //   REM == 1: v[2b + u]  = X[lt + b*TPR + u*N/2]
    template <bool TO_SMEM>
    __device__ __forceinline__ void run(double2 (&v)[8], double2* s, int lt) {
        int Ns = 1;
#pragma unroll
        for (int p = 0; p < NB8; ++p) {
            if (p > 0) {
                __syncthreads();
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = s[swz(lt + t * TPR)];
                const double2 w1 = W8S ? w8s[(p - 1) * blockDim.x] : w8[p - 1];
                const double2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                v[5] = cmul(v[5], cmul(w4, w1));
                v[6] = cmul(v[6], cmul(w3, w3));
                v[7] = cmul(v[7], cmul(w4, w3));
            }
            bfly8<SIGN>(v);
            const bool last = (p == NB8 - 1) && REM == 0;
            if (!last || TO_SMEM) {
                if (p > 0) __syncthreads();   // every thread has loaded its inputs of this pass
                const int k = lt & (Ns - 1);
                const int j0 = (lt - k) * 8 + k;
#pragma unroll
                for (int u = 0; u < 8; ++u) s[swz(j0 + u * Ns)] = v[u];
            }
            Ns *= 8;
        }
        if (REM == 2) {
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int t = 0; t < 4; ++t) v[b * 4 + t] = s[swz(lt + b * TPR + t * 2 * TPR)];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double2 w1 = wrem(b), w2 = cmul(w1, w1);
                v[b * 4 + 1] = cmul(v[b * 4 + 1], w1);
                v[b * 4 + 2] = cmul(v[b * 4 + 2], w2);
                v[b * 4 + 3] = cmul(v[b * 4 + 3], cmul(w2, w1));
                bfly4<SIGN>(v[b * 4 + 0], v[b * 4 + 1], v[b * 4 + 2], v[b * 4 + 3]);
            }
            if (TO_SMEM) {
                __syncthreads();
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int u = 0; u < 4; ++u) s[swz(lt + b * TPR + u * Ns)] = v[b * 4 + u];   // Ns == N/4
            }
        } else if (REM == 1) {
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                v[b * 2 + 0] = s[swz(lt + b * TPR)];
                v[b * 2 + 1] = s[swz(lt + b * TPR + 4 * TPR)];
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double2 w = cmul(v[b * 2 + 1], wrem(b));
                const double2 x0 = cadd(v[b * 2], w), x1 = csub(v[b * 2], w);
                v[b * 2] = x0;
                v[b * 2 + 1] = x1;
            }
            if (TO_SMEM) {
                __syncthreads();
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    s[swz(lt + b * TPR)] = v[b * 2];
                    s[swz(lt + b * TPR + Ns)] = v[b * 2 + 1];   // Ns == N/2
                }
            }
        }
        if (TO_SMEM) __syncthreads();
    }

    // index of register slot e (0..7) of the last pass in the output row (see run<false>)
    static __device__ __forceinline__ int out_index(int lt, int e) {
        if (REM == 0) return lt + e * TPR;
        if (REM == 2) return lt + (e >> 2) * TPR + (e & 3) * (N / 4);
        return lt + (e >> 1) * TPR + (e & 1) * (N / 2);
    }
};

// psi~1(0,0) of one member: the fixed-order sum of K3's per-slab shares (or the value k3_gauge
// left in scal[1]).  Called by every thread of the block; two block barriers.
__device__ __forceinline__ double load_gauge(const FftArgs& a, int member, double* sh) {
    if (!a.use_gauge) return 0.0;
    if (a.gpart == nullptr) return a.scal[member * 4 + 1];
    double loc = 0.0;
    for (int i = threadIdx.x; i < a.ngp; i += blockDim.x) loc += a.gpart[(int64_t)member * a.ngp + i];
    return block_sum(loc, sh);
}

// L2 prefetch of a row that a later loop iteration will read: `nthr` threads (index `t`) cover
// `bytes` bytes, one request per 128-byte line.  Hides the HBM latency of the persistent loop's
// next row behind the current row's passes (there are no spare registers to double-buffer in).
__device__ __forceinline__ void prefetch_row_l2(const void* base, int bytes, int t, int nthr) {
    const char* p = static_cast<const char*>(base);
    for (int off = t * 128; off < bytes; off += nthr * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

template <int LOG2N>
struct FftLaunch {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPR = N / 8;
    static constexpr int RPB = TPR >= 128 ? 1 : 128 / TPR;
    static constexpr int THREADS = TPR * RPB;
    static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : (THREADS >= 256 ? 4 : 6));
    static constexpr size_t SMEM = (size_t)RPB * N * sizeof(double2);
};

// ---------------------------------------------------------------------------------------
// Forward kernel (persistent): row groups of RPB rows, grid-stride over (member, group).
// ---------------------------------------------------------------------------------------
template <int LOG2N>
__global__ void __launch_bounds__(FftLaunch<LOG2N>::THREADS, FftLaunch<LOG2N>::MINB)
k2_fft_forward(const FftArgs a, int ngroups_per_member, int ngroups_total) {
    using L = FftLaunch<LOG2N>;
    using F = RowFft<LOG2N, -1>;
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR;
    const int lr = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    double2* s = fft_smem + (size_t)lr * N;
    F fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int row = (grp - member * ngroups_per_member) * L::RPB + lr;
        const bool live = row < a.pl.P;
        const double* __restrict__ q1 = a.q1 + member * a.mstride + a.g.at(0, live ? row : 0);
        const double* __restrict__ q2 = a.q2 + member * a.mstride + a.g.at(0, live ? row : 0);
        double x1[8], x2[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            x1[t] = __ldg(q1 + lt + t * TPR);
            x2[t] = __ldg(q2 + lt + t * TPR);
        }
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t)
            v[t] = make_double2(A0 * x1[t] + A1 * x2[t], A2 * x1[t] + A3 * x2[t]);   // src/model.jl:180
        fft.template run<true>(v, s, lt);
        if (live) {
            double2* __restrict__ out =
                reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
            constexpr int half = N >> 1;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = lt + b * TPR;   // 0 .. N/2 - 1
                if (k == 0) {
                    const double2 z0 = s[swz(0)];
                    out[0] = z0;
                    out[half] = s[swz(half)];
                    store_col0(a, member, row, z0.x, pushed);   // Q1[0]: the Poisson k=0 column
                } else {
                    const double2 X = s[swz(k)], Y = s[swz(N - k)];
                    out[k] = make_double2(0.5 * (X.x + Y.x), 0.5 * (X.y - Y.y));        // Q1[k]
                    out[N - k] = make_double2(0.5 * (X.y + Y.y), 0.5 * (Y.x - X.x));    // Q2[k]
                }
            }
        }
        __syncthreads();   // the row buffer is reused by the next group
    }
    peer_store_fence(pushed);
}

// ---------------------------------------------------------------------------------------
// Inverse kernel (persistent).
// ---------------------------------------------------------------------------------------
template <int LOG2N>
__global__ void __launch_bounds__(FftLaunch<LOG2N>::THREADS, FftLaunch<LOG2N>::MINB)
k4_fft_inverse(const FftArgs a, int ngroups_per_member, int ngroups_total) {
    using L = FftLaunch<LOG2N>;
    using F = RowFft<LOG2N, +1>;
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR;
    const int lr = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    double2* s = fft_smem + (size_t)lr * N;
    F fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    constexpr int half = N >> 1;
    __shared__ double gsh[32];
    int gmember = -1;
    double gauge = 0.0;

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int row = (grp - member * ngroups_per_member) * L::RPB + lr;
        const bool live = row < P;
        if (member != gmember) {   // block-uniform
            gauge = load_gauge(a, member, gsh);
            gmember = member;
        }
        const double2* __restrict__ in =
            reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)(live ? row : 0) * a.pl.ncol);
        // Z[k] = U1[k] + i U2[k] (k < N/2), Z[N-k] = conj(U1[k]) + i conj(U2[k]); slot k holds
        // U1[k] and slot N-k holds U2[k]; slots 0 and N/2 hold (real, real) pairs = Z itself.
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int k = lt + t * TPR;
            if (k == 0 || k == half) {
                v[t] = __ldg(in + k);
            } else {
                const double2 X = __ldg(in + k), Y = __ldg(in + N - k);
                // k < N/2: X = U1[k], Y = U2[k] -> U1 + i U2
                // k > N/2: X = U2[k'], Y = U1[k'] (k' = N-k) -> conj(U1) + i conj(U2)
                v[t] = (k < half) ? make_double2(X.x - Y.y, X.y + Y.x) : make_double2(Y.x + X.y, X.x - Y.y);
            }
        }
        fft.template run<false>(v, s, lt);
        if (live) {
            double* __restrict__ p1 = a.psi1 + member * a.mstride;
            double* __restrict__ p2 = a.psi2 + member * a.mstride;
            // images of the edge rows: own array (periodic) or the ring neighbours' (NVLink peer memory)
            const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
            double* __restrict__ lo1 = a.pimg_lo + member * a.mstride;
            double* __restrict__ lo2 = lo1 + a.g.fstride;
            double* __restrict__ hi1 = a.pimg_hi + member * a.mstride;
            double* __restrict__ hi2 = hi1 + a.g.fstride;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = F::out_index(lt, e);
                const double2 z = v[e];
                const double t1 = z.x - gauge;            // pinned node: psi~1(0,0) = 0
                const double o1 = A0 * t1 + A1 * z.y;     // src/model.jl:196
                const double o2 = A2 * t1 + A3 * z.y;
                const int64_t o = a.g.at(n, row);
                const bool gl = n < GHOST, gr = n >= M - GHOST;
                p1[o] = o1; p2[o] = o2;
                if (gl) { p1[o + M] = o1; p2[o + M] = o2; }
                if (gr) { p1[o - M] = o1; p2[o - M] = o2; }
                if (gb) {
                    lo1[o + dyo] = o1; lo2[o + dyo] = o2;
                    if (gl) { lo1[o + dyo + M] = o1; lo2[o + dyo + M] = o2; }
                    if (gr) { lo1[o + dyo - M] = o1; lo2[o + dyo - M] = o2; }
                }
                if (gt) {
                    hi1[o - dyo] = o1; hi2[o - dyo] = o2;
                    if (gl) { hi1[o - dyo + M] = o1; hi2[o - dyo + M] = o2; }
                    if (gr) { hi1[o - dyo - M] = o1; hi2[o - dyo - M] = o2; }
                }
            }
            if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged before the thread moves on
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Radix-16 variant for N = 16^n (M = 4096, the headline grid, and M = 256): 16 points per
// thread, N/16 threads per row.  Both transforms are bound by shared-memory wavefronts (ncu:
// LSU data pipe 66 %, DRAM 44 %), and radix-16 needs one exchange less per row (3 passes
// instead of 4 at N = 4096).  Same Stockham indexing, same spectral layout.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int swz16(int i) { return i ^ ((i >> 4) & 7); }

// 16-point DFT in place, natural order output: X[u] = sum_t a[t] exp(SIGN * 2 pi i t u / 16)
template <int SIGN>
__device__ __forceinline__ void bfly16(double2 (&a)[16]) {
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    // step 1: 4-point DFTs over t = c + 4m  ->  y[c][q] stored at a[c + 4q]
#pragma unroll
    for (int c = 0; c < 4; ++c) bfly4<SIGN>(a[c], a[c + 4], a[c + 8], a[c + 12]);
    // step 2: twiddles W16^(c q), W16 = exp(SIGN * 2 pi i / 16)
    {
        const double2 W1 = make_double2(c1, SIGN * s1), W2 = make_double2(h, SIGN * h), W3 = make_double2(s1, SIGN * c1);
        a[1 + 4] = cmul(a[1 + 4], W1);                                  // c=1,q=1
        a[1 + 8] = cmul(a[1 + 8], W2);                                  // c=1,q=2
        a[1 + 12] = cmul(a[1 + 12], W3);                                // c=1,q=3
        a[2 + 4] = cmul(a[2 + 4], W2);                                  // c=2,q=1
        a[2 + 8] = muli<SIGN>(a[2 + 8]);                                // c=2,q=2: W16^4 = SIGN i
        a[2 + 12] = cmul(a[2 + 12], make_double2(-h, SIGN * h));        // c=2,q=3: W16^6
        a[3 + 4] = cmul(a[3 + 4], W3);                                  // c=3,q=1
        a[3 + 8] = cmul(a[3 + 8], make_double2(-h, SIGN * h));          // c=3,q=2: W16^6
        a[3 + 12] = cmul(a[3 + 12], make_double2(-c1, -SIGN * s1));     // c=3,q=3: W16^9 = -W16^1
    }
    // step 3: 4-point DFTs over c for each q: X[q + 4r] = sum_c y'[c][q] W4^(c r)
#pragma unroll
    for (int q = 0; q < 4; ++q) bfly4<SIGN>(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
    // now a[4q + r] = X[q + 4r]: transpose the 4x4 index block into natural order
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = q + 1; r < 4; ++r) {
            const double2 t = a[4 * q + r];
            a[4 * q + r] = a[4 * r + q];
            a[4 * r + q] = t;
        }
}

template <int LOG2N, int SIGN>
struct RowFft16 {
    static_assert(LOG2N % 4 == 0, "N must be a power of 16");
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPR = N / 16;
    static constexpr int NB = LOG2N / 4;
    double2 w16[NB > 1 ? NB - 1 : 1];   // base twiddle of pass p = 1 .. NB-1

    // `stride`: the table holds exp(-2 pi i n / (stride * N))
    __device__ __forceinline__ void init(const double2* __restrict__ tw, int lt, int stride = 1) {
        int Ns = 16;
#pragma unroll
        for (int p = 1; p < NB; ++p) {
            const int k = lt & (Ns - 1);
            w16[p - 1] = twid<SIGN>(tw, stride * (k * (N / (Ns * 16))));
            Ns *= 16;
        }
    }

    // On entry v[t] = x[lt + t*TPR].  TO_SMEM: on exit the transform sits in `s` (swizzled, natural
    // order) after a __syncthreads(); otherwise v[u] = X[lt + u*TPR].
    template <bool TO_SMEM>
    __device__ __forceinline__ void run(double2 (&v)[16], double2* s, int lt) {
        int Ns = 1;
#pragma unroll
        for (int p = 0; p < NB; ++p) {
            if (p > 0) {
                __syncthreads();
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = s[swz16(lt + t * TPR)];
                const double2 w1 = w16[p - 1];
                const double2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                const double2 w5 = cmul(w4, w1), w7 = cmul(w4, w3), w8 = cmul(w4, w4);
                v[5] = cmul(v[5], w5);
                v[6] = cmul(v[6], cmul(w3, w3));
                v[7] = cmul(v[7], w7);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], cmul(w8, w4));
                v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], cmul(w7, w7));
                v[15] = cmul(v[15], cmul(w8, w7));
            }
            bfly16<SIGN>(v);
            const bool last = (p == NB - 1);
            if (!last || TO_SMEM) {
                if (p > 0) __syncthreads();   // every thread has loaded its inputs of this pass
                const int k = lt & (Ns - 1);
                const int j0 = (lt - k) * 16 + k;
#pragma unroll
                for (int u = 0; u < 16; ++u) s[swz16(j0 + u * Ns)] = v[u];
            }
            Ns *= 16;
        }
        if (TO_SMEM) __syncthreads();
    }
};

template <int LOG2N>
struct Fft16Launch {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPR = N / 16;
    static constexpr int RPB = TPR >= 128 ? 1 : 128 / TPR;
    static constexpr int THREADS = TPR * RPB;
    static constexpr int MINB = 2;
    static constexpr size_t SMEM = (size_t)RPB * N * sizeof(double2);
};

template <int LOG2N>
__global__ void __launch_bounds__(Fft16Launch<LOG2N>::THREADS, Fft16Launch<LOG2N>::MINB)
k2_fft16_forward(const FftArgs a, int ngroups_per_member, int ngroups_total, int pf) {
    using L = Fft16Launch<LOG2N>;
    using F = RowFft16<LOG2N, -1>;
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR;
    const int lr = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    double2* s = fft_smem + (size_t)lr * N;
    F fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int row = (grp - member * ngroups_per_member) * L::RPB + lr;
        const bool live = row < a.pl.P;
        const double* __restrict__ q1 = a.q1 + member * a.mstride + a.g.at(0, live ? row : 0);
        const double* __restrict__ q2 = a.q2 + member * a.mstride + a.g.at(0, live ? row : 0);
        double2 v[16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            double x1[8], x2[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                x1[t] = __ldg(q1 + lt + (8 * hf + t) * TPR);
                x2[t] = __ldg(q2 + lt + (8 * hf + t) * TPR);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t)
                v[8 * hf + t] = make_double2(A0 * x1[t] + A1 * x2[t], A2 * x1[t] + A3 * x2[t]);   // src/model.jl:180
        }
        if (pf) {   // this CTA's next row group -> L2 (two resident CTAs of 8 warps cannot hide DRAM latency)
            const int gn = grp + gridDim.x;
            if (gn < ngroups_total) {
                const int mn = gn / ngroups_per_member;
                const int rn = min((gn - mn * ngroups_per_member) * L::RPB + lr, a.pl.P - 1);
                prefetch_row_l2(a.q1 + mn * a.mstride + a.g.at(0, rn), N * 8, lt, TPR);
                prefetch_row_l2(a.q2 + mn * a.mstride + a.g.at(0, rn), N * 8, lt, TPR);
            }
        }
        fft.template run<true>(v, s, lt);
        if (live) {
            double2* __restrict__ out =
                reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
            constexpr int half = N >> 1;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int k = lt + b * TPR;   // 0 .. N/2 - 1
                if (k == 0) {
                    const double2 z0 = s[swz16(0)];
                    out[0] = z0;
                    out[half] = s[swz16(half)];
                    store_col0(a, member, row, z0.x, pushed);   // Q1[0]: the Poisson k=0 column
                } else {
                    const double2 X = s[swz16(k)], Y = s[swz16(N - k)];
                    out[k] = make_double2(0.5 * (X.x + Y.x), 0.5 * (X.y - Y.y));        // Q1[k]
                    out[N - k] = make_double2(0.5 * (X.y + Y.y), 0.5 * (Y.x - X.x));    // Q2[k]
                }
            }
        }
        __syncthreads();   // the row buffer is reused by the next group
    }
    peer_store_fence(pushed);
}

template <int LOG2N>
__global__ void __launch_bounds__(Fft16Launch<LOG2N>::THREADS, Fft16Launch<LOG2N>::MINB)
k4_fft16_inverse(const FftArgs a, int ngroups_per_member, int ngroups_total, int pf) {
    using L = Fft16Launch<LOG2N>;
    using F = RowFft16<LOG2N, +1>;
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR;
    const int lr = threadIdx.x / TPR, lt = threadIdx.x % TPR;
    double2* s = fft_smem + (size_t)lr * N;
    F fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    constexpr int half = N >> 1;
    __shared__ double gsh[32];
    int gmember = -1;
    double gauge = 0.0;

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int row = (grp - member * ngroups_per_member) * L::RPB + lr;
        const bool live = row < P;
        if (member != gmember) {   // block-uniform
            gauge = load_gauge(a, member, gsh);
            gmember = member;
        }
        const double2* __restrict__ in =
            reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)(live ? row : 0) * a.pl.ncol);
        double2 v[16];   // see k4_fft_inverse for the re-tangling
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int k = lt + t * TPR;
            if (k == 0 || k == half) {
                v[t] = __ldg(in + k);
            } else {
                const double2 X = __ldg(in + k), Y = __ldg(in + N - k);
                v[t] = (k < half) ? make_double2(X.x - Y.y, X.y + Y.x) : make_double2(Y.x + X.y, X.x - Y.y);
            }
        }
        if (pf) {   // this CTA's next spectral row(s) -> L2
            const int gn = grp + gridDim.x;
            if (gn < ngroups_total) {
                const int mn = gn / ngroups_per_member;
                const int rn = min((gn - mn * ngroups_per_member) * L::RPB + lr, P - 1);
                prefetch_row_l2(a.S + mn * a.sstride + (int64_t)rn * a.pl.ncol, N * 16, lt, TPR);
            }
        }
        fft.template run<false>(v, s, lt);
        if (live) {
            double* __restrict__ p1 = a.psi1 + member * a.mstride;
            double* __restrict__ p2 = a.psi2 + member * a.mstride;
            // images of the edge rows: own array (periodic) or the ring neighbours' (NVLink peer memory)
            const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
            double* __restrict__ lo1 = a.pimg_lo + member * a.mstride;
            double* __restrict__ lo2 = lo1 + a.g.fstride;
            double* __restrict__ hi1 = a.pimg_hi + member * a.mstride;
            double* __restrict__ hi2 = hi1 + a.g.fstride;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = lt + e * TPR;
                const double2 z = v[e];
                const double t1 = z.x - gauge;            // pinned node: psi~1(0,0) = 0
                const double o1 = A0 * t1 + A1 * z.y;     // src/model.jl:196
                const double o2 = A2 * t1 + A3 * z.y;
                const int64_t o = a.g.at(n, row);
                const bool gl = n < GHOST, gr = n >= M - GHOST;
                p1[o] = o1; p2[o] = o2;
                if (gl) { p1[o + M] = o1; p2[o + M] = o2; }
                if (gr) { p1[o - M] = o1; p2[o - M] = o2; }
                if (gb) {
                    lo1[o + dyo] = o1; lo2[o + dyo] = o2;
                    if (gl) { lo1[o + dyo + M] = o1; lo2[o + dyo + M] = o2; }
                    if (gr) { lo1[o + dyo - M] = o1; lo2[o + dyo - M] = o2; }
                }
                if (gt) {
                    hi1[o - dyo] = o1; hi2[o - dyo] = o2;
                    if (gl) { hi1[o - dyo + M] = o1; hi2[o - dyo + M] = o2; }
                    if (gr) { hi1[o - dyo - M] = o1; hi2[o - dyo - M] = o2; }
                }
            }
            if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged before the thread moves on
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Ring-buffered, TMA-fed variant for N = 4096 (the headline grid).
//
// ncu on the kernels above (profiles/r01j): 128 registers and 64 KB of shared memory per 256-thread
// CTA allow two CTAs = 4 warps per scheduler, too few to hide a global-memory round trip, so a third
// of all issue slots were lost waiting for the row loads (long scoreboard 1.9 - 2.0 stalled warps per
// issue, + LG throttle) although the next row had been prefetched into L2.  Here the loads leave the
// instruction stream altogether:
//   * one persistent CTA per SM, 512 threads = two GROUPS of 256 threads; a group transforms one
//     row at a time exactly as above, synchronising with a named barrier of its own (bar.sync id, 256),
//     so the two groups run out of phase and overlap each other's exchange / butterfly / store phases;
//   * three 64 KB row buffers in a ring.  Row j of the CTA lives in buffer j mod 3: it ARRIVES there
//     by bulk asynchronous copies (cp.async.bulk global -> shared, SASS UBLKCP, completion counted on
//     an mbarrier) - the two 32 KB PV rows for the forward transform, the 64 KB spectral row for the
//     inverse - and is then transformed IN PLACE.  When a group has finished with its buffer, one
//     of its threads issues the copies of row j + 3 into it, which the OTHER group will consume: every
//     row is in flight for two full row periods before anybody waits for it.
// Arithmetic, operation order and spectral layout are those of k2_fft16_forward / k4_fft16_inverse:
// results are bit-identical (QG_FFT_RING=0 selects the kernels above).
// ---------------------------------------------------------------------------------------
constexpr int RING_N = 4096, RING_TPR = 256, RING_THREADS = 512, RING_NBUF = 3;
constexpr size_t RING_BUF_BYTES = (size_t)RING_N * sizeof(double2);                     // 64 KB
constexpr size_t RING_SMEM = RING_NBUF * RING_BUF_BYTES + 64 /* mbarriers */ + 2 * 32 * sizeof(double) /* gauge sums */;

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

// Which mbarrier announces row j of a CTA, and with which phase parity.  Buffer j mod 3 is used for the n-th time
// by row j, n = j / 3.  A parity wait cannot tell "phase n has completed" from "phase n - 1 has not": with ONE
// barrier per buffer a group waiting for row j while the copy of row j - 3 were still in flight would sail
// through (tests/test_ring_protocol.py finds that interleaving).  Two barriers per buffer, used alternately,
// remove the ambiguity: the previous use of barrier (b, n & 1) is row j - 6, which the waiting group has itself
// consumed, so that barrier is either in the phase of row j (wait) or past it (go).
__device__ __forceinline__ int ring_bar(int j) { return (j % RING_NBUF) * 2 + ((j / RING_NBUF) & 1); }
__device__ __forceinline__ uint32_t ring_parity(int j) { return (uint32_t)(j / (2 * RING_NBUF)) & 1u; }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One instruction pulls a whole row into L2 ahead of the bulk copy that will fetch it: the ring leaves a
// copy only about half a row period to land (one buffer of three is in flight while two are worked on).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// RowFft16<12>::run with the block barrier replaced by the group's named barrier
template <int SIGN>
struct RingFft {
    static constexpr int N = RING_N, TPR = RING_TPR;
    double2 w16[2];
    __device__ __forceinline__ void init(const double2* __restrict__ tw, int lt) {
        w16[0] = twid<SIGN>(tw, (lt & 15) * (N / 256));
        w16[1] = twid<SIGN>(tw, lt & 255);
    }
    template <bool TO_SMEM>
    __device__ __forceinline__ void run(double2 (&v)[16], double2* s, int lt, int bar) {
        int Ns = 1;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (p > 0) {
                group_sync(bar);
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = s[swz16(lt + t * TPR)];
                const double2 w1 = w16[p - 1];
                const double2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                const double2 w5 = cmul(w4, w1), w7 = cmul(w4, w3), w8 = cmul(w4, w4);
                v[5] = cmul(v[5], w5);
                v[6] = cmul(v[6], cmul(w3, w3));
                v[7] = cmul(v[7], w7);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], cmul(w8, w4));
                v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], cmul(w7, w7));
                v[15] = cmul(v[15], cmul(w8, w7));
            }
            bfly16<SIGN>(v);
            const bool last = (p == 2);
            if (!last || TO_SMEM) {
                group_sync(bar);   // every thread of the group has loaded its inputs of this pass
                const int k = lt & (Ns - 1);
                const int j0 = (lt - k) * 16 + k;
#pragma unroll
                for (int u = 0; u < 16; ++u) s[swz16(j0 + u * Ns)] = v[u];
            }
            Ns *= 16;
        }
        if (TO_SMEM) group_sync(bar);
    }
};

// block_sum restricted to one 256-thread group (same summation tree as a 256-thread block)
__device__ __forceinline__ double group_sum(double v, double* sh, int lt, int bar) {
    const int lane = lt & 31, w = lt >> 5;
    v = warp_sum(v);
    group_sync(bar);
    if (lane == 0) sh[w] = v;
    group_sync(bar);
    return warp_sum(lane < 8 ? sh[lane] : 0.0);
}

__global__ void __launch_bounds__(RING_THREADS, 1)
k2_fft16_ring(const FftArgs a, int rows_per_member, int total_rows, int pf) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(ring_raw + RING_NBUF * RING_BUF_BYTES);
    constexpr int N = RING_N, TPR = RING_TPR, half = N >> 1;
    const int grp = threadIdx.x >> 8, lt = threadIdx.x & 255, bar = 1 + grp;
    RingFft<-1> fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)
    const int G = gridDim.x;
    const int nmine = (total_rows - (int)blockIdx.x + G - 1) / G;   // this CTA's rows: blockIdx.x + j * G

    auto issue = [&](int j) {   // one thread: both PV rows of the CTA's j-th row -> buffer j mod 3
        const int gr = blockIdx.x + j * G, member = gr / rows_per_member, row = gr - member * rows_per_member;
        unsigned char* dst = ring_raw + (size_t)(j % RING_NBUF) * RING_BUF_BYTES;
        uint64_t* fb = &full[ring_bar(j)];
        mbar_expect_tx(fb, (uint32_t)RING_BUF_BYTES);
        bulk_load(dst, a.q1 + member * a.mstride + a.g.at(0, row), N * sizeof(double), fb);
        bulk_load(dst + N * sizeof(double), a.q2 + member * a.mstride + a.g.at(0, row), N * sizeof(double), fb);
        if (pf && j + pf < nmine) {   // the row `pf` turns of the ring later -> L2
            const int g2 = blockIdx.x + (j + pf) * G, m2 = g2 / rows_per_member, r2 = g2 - m2 * rows_per_member;
            bulk_prefetch_l2(a.q1 + m2 * a.mstride + a.g.at(0, r2), N * sizeof(double));
            bulk_prefetch_l2(a.q2 + m2 * a.mstride + a.g.at(0, r2), N * sizeof(double));
        }
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2 * RING_NBUF; ++b) mbar_init(&full[b], 1);
        for (int j = 0; j < RING_NBUF && j < nmine; ++j) issue(j);
    }
    __syncthreads();

    for (int j = grp; j < nmine; j += 2) {
        const int gr = blockIdx.x + j * G, member = gr / rows_per_member, row = gr - member * rows_per_member;
        double2* s = reinterpret_cast<double2*>(ring_raw + (size_t)(j % RING_NBUF) * RING_BUF_BYTES);
        mbar_wait(&full[ring_bar(j)], ring_parity(j));
        const double* x1s = reinterpret_cast<const double*>(s);
        const double* x2s = x1s + N;
        double2 v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const double x1 = x1s[lt + t * TPR], x2 = x2s[lt + t * TPR];
            v[t] = make_double2(A0 * x1 + A1 * x2, A2 * x1 + A3 * x2);   // src/model.jl:180
        }
        group_sync(bar);   // the raw rows have been consumed: pass 0 may overwrite them
        fft.template run<true>(v, s, lt, bar);
        {
            double2* __restrict__ out = reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int k = lt + b * TPR;   // 0 .. N/2 - 1
                if (k == 0) {
                    const double2 z0 = s[swz16(0)];
                    out[0] = z0;
                    out[half] = s[swz16(half)];
                    store_col0(a, member, row, z0.x, pushed);   // Q1[0]: the Poisson k=0 column
                } else {
                    const double2 X = s[swz16(k)], Y = s[swz16(N - k)];
                    out[k] = make_double2(0.5 * (X.x + Y.x), 0.5 * (X.y - Y.y));        // Q1[k]
                    out[N - k] = make_double2(0.5 * (X.y + Y.y), 0.5 * (Y.x - X.x));    // Q2[k]
                }
            }
        }
        // (the global stores above consumed every value loaded from the buffer, so behind this barrier the loads
        // have completed, not merely been issued, and the bulk copy may overwrite the buffer)
        group_sync(bar);   // the buffer is free
        if (lt == 0 && j + RING_NBUF < nmine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(j + RING_NBUF);
        }
    }
    peer_store_fence(pushed);
}

__global__ void __launch_bounds__(RING_THREADS, 1)
k4_fft16_ring(const FftArgs a, int rows_per_member, int total_rows) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(ring_raw + RING_NBUF * RING_BUF_BYTES);
    double* gsh = reinterpret_cast<double*>(ring_raw + RING_NBUF * RING_BUF_BYTES + 64);
    constexpr int N = RING_N, TPR = RING_TPR, half = N >> 1;
    const int grp = threadIdx.x >> 8, lt = threadIdx.x & 255, bar = 1 + grp;
    RingFft<+1> fft;
    fft.init(a.pl.tw, lt);
    const double A0 = a.A[0], A1 = a.A[1], A2 = a.A[2], A3 = a.A[3];
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    const int G = gridDim.x;
    const int nmine = (total_rows - (int)blockIdx.x + G - 1) / G;

    auto issue = [&](int j) {   // one thread: the spectral row of the CTA's j-th row -> buffer j mod 3
        const int gr = blockIdx.x + j * G, member = gr / rows_per_member, row = gr - member * rows_per_member;
        uint64_t* fb = &full[ring_bar(j)];
        mbar_expect_tx(fb, (uint32_t)RING_BUF_BYTES);
        bulk_load(ring_raw + (size_t)(j % RING_NBUF) * RING_BUF_BYTES, a.S + member * a.sstride + (int64_t)row * a.pl.ncol,
                  (uint32_t)RING_BUF_BYTES, fb);
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2 * RING_NBUF; ++b) mbar_init(&full[b], 1);
        for (int j = 0; j < RING_NBUF && j < nmine; ++j) issue(j);
    }
    __syncthreads();
    int gmember = -1;
    double gauge = 0.0;

    for (int j = grp; j < nmine; j += 2) {
        const int gr = blockIdx.x + j * G, member = gr / rows_per_member, row = gr - member * rows_per_member;
        if (member != gmember) {   // uniform over the group: psi~1(0,0), see load_gauge
            gauge = 0.0;
            if (a.use_gauge) {
                if (a.gpart == nullptr) {
                    gauge = a.scal[member * 4 + 1];
                } else {
                    double loc = 0.0;
                    for (int i = lt; i < a.ngp; i += TPR) loc += a.gpart[(int64_t)member * a.ngp + i];
                    gauge = group_sum(loc, gsh + 32 * grp, lt, bar);
                }
            }
            gmember = member;
        }
        double2* s = reinterpret_cast<double2*>(ring_raw + (size_t)(j % RING_NBUF) * RING_BUF_BYTES);
        mbar_wait(&full[ring_bar(j)], ring_parity(j));
        double2 v[16];   // see k4_fft_inverse for the re-tangling
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int k = lt + t * TPR;
            if (k == 0 || k == half) {
                v[t] = s[k];
            } else {
                const double2 X = s[k], Y = s[N - k];
                v[t] = (k < half) ? make_double2(X.x - Y.y, X.y + Y.x) : make_double2(Y.x + X.y, X.x - Y.y);
            }
        }
        group_sync(bar);   // the spectral row has been consumed: pass 0 may overwrite it
        fft.template run<false>(v, s, lt, bar);
        double* __restrict__ p1 = a.psi1 + member * a.mstride;
        double* __restrict__ p2 = a.psi2 + member * a.mstride;
        // images of the edge rows: own array (periodic) or the ring neighbours' (NVLink peer memory)
        const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
        double* __restrict__ lo1 = a.pimg_lo + member * a.mstride;
        double* __restrict__ lo2 = lo1 + a.g.fstride;
        double* __restrict__ hi1 = a.pimg_hi + member * a.mstride;
        double* __restrict__ hi2 = hi1 + a.g.fstride;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int n = lt + e * TPR;
            const double2 z = v[e];
            const double t1 = z.x - gauge;            // pinned node: psi~1(0,0) = 0
            const double o1 = A0 * t1 + A1 * z.y;     // src/model.jl:196
            const double o2 = A2 * t1 + A3 * z.y;
            const int64_t o = a.g.at(n, row);
            const bool gl = n < GHOST, gr_ = n >= M - GHOST;
            p1[o] = o1; p2[o] = o2;
            if (gl) { p1[o + M] = o1; p2[o + M] = o2; }
            if (gr_) { p1[o - M] = o1; p2[o - M] = o2; }
            if (gb) {
                lo1[o + dyo] = o1; lo2[o + dyo] = o2;
                if (gl) { lo1[o + dyo + M] = o1; lo2[o + dyo + M] = o2; }
                if (gr_) { lo1[o + dyo - M] = o1; lo2[o + dyo - M] = o2; }
            }
            if (gt) {
                hi1[o - dyo] = o1; hi2[o - dyo] = o2;
                if (gl) { hi1[o - dyo + M] = o1; hi2[o - dyo + M] = o2; }
                if (gr_) { hi1[o - dyo - M] = o1; hi2[o - dyo - M] = o2; }
            }
        }
        if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged before the buffer hand-over
        // The stores above depend on every value the last pass loaded from the buffer, and stores do not pass
        // the barrier: behind it all of the group's shared-memory loads have completed (a barrier alone orders
        // their issue only) and the bulk copy of row j + 3 may overwrite the buffer.
        group_sync(bar);
        if (lt == 0 && j + RING_NBUF < nmine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(j + RING_NBUF);
        }
    }
}

// ---------------------------------------------------------------------------------------
// M = 2N too long for one shared-memory row (M = 16384: 256 KB as a packed complex row).
// One CTA per (row, modal field) forward / (row, layer) inverse runs a real transform of
// length M as a complex transform of length N = M/2 on z[n] = x[2n] + i x[2n+1] plus the
// usual split  X[k] = E[k] + W^k O[k],  W = exp(-2 pi i / M).  The projections are linear,
// so they are applied in physical space before the forward transform and in spectral space
// before the inverse one (each CTA reads both inputs).  Same spectral layout as above.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// exp(-i pi t / 8), t = 0..7: W^(t * N/8) for M = 2N
__device__ __forceinline__ double2 w16th(int t) {
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    switch (t) {
        case 0: return make_double2(1.0, 0.0);
        case 1: return make_double2(c1, -s1);
        case 2: return make_double2(h, -h);
        case 3: return make_double2(s1, -c1);
        case 4: return make_double2(0.0, -1.0);
        case 5: return make_double2(-s1, -c1);
        case 6: return make_double2(-h, -h);
        default: return make_double2(-c1, -s1);
    }
}

template <int LOG2N>
__global__ void __launch_bounds__(FftLaunch<LOG2N>::THREADS, FftLaunch<LOG2N>::MINB)
k2_rfft_forward(const FftArgs a, int ngroups_per_member, int ngroups_total, int pf) {
    using L = FftLaunch<LOG2N>;
    using F = RowFft<LOG2N, -1, true>;
    static_assert(L::RPB == 1, "long-row path: one row per CTA");
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR, M = 2 * N;
    const int lt = threadIdx.x;
    double2* s = fft_smem;
    F fft;
    fft.init(a.pl.tw, lt, 2, fft_smem + N);   // half-length twiddles: exp(-2 pi i n / N) = tw[2n]
    const double2 w0 = __ldg(a.pl.tw + lt);   // W^lt, W = exp(-2 pi i / M)
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int rf = grp - member * ngroups_per_member;
        const int row = rf >> 1, field = rf & 1;
        const double A0 = a.A[2 * field], A1 = a.A[2 * field + 1];   // row `field` of P_inv
        const double2* __restrict__ q1 = reinterpret_cast<const double2*>(a.q1 + member * a.mstride + a.g.at(0, row));
        const double2* __restrict__ q2 = reinterpret_cast<const double2*>(a.q2 + member * a.mstride + a.g.at(0, row));
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const double2 x1 = __ldg(q1 + lt + t * TPR), x2 = __ldg(q2 + lt + t * TPR);
            v[t] = make_double2(A0 * x1.x + A1 * x2.x, A0 * x1.y + A1 * x2.y);   // (q~[2n], q~[2n+1])
        }
        if (pf) {   // next (row, field) of this CTA -> L2
            const int gn = grp + gridDim.x;
            if (gn < ngroups_total) {
                const int mn = gn / ngroups_per_member;
                const int rn = (gn - mn * ngroups_per_member) >> 1;
                prefetch_row_l2(a.q1 + mn * a.mstride + a.g.at(0, rn), M * 8, lt, TPR);
                prefetch_row_l2(a.q2 + mn * a.mstride + a.g.at(0, rn), M * 8, lt, TPR);
            }
        }
        fft.template run<true>(v, s, lt);
        double2* __restrict__ out =
            reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
        double* __restrict__ outs = reinterpret_cast<double*>(out);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int k = lt + b * TPR;   // 0 .. N/2 - 1
            if (k == 0) {
                const double2 Z0 = s[swz(0)], Zh = s[swz(N / 2)];
                outs[0 + field] = Z0.x + Z0.y;              // X[0]  -> slot 0, component `field`
                outs[2 * N + field] = Z0.x - Z0.y;          // X[N]  -> slot M/2
                const double2 Xh = cconj(Zh);               // X[N/2]
                if (field == 0) out[N / 2] = Xh; else out[M - N / 2] = Xh;
                if (field == 0) store_col0(a, member, row, Z0.x + Z0.y, pushed);
            } else {
                const double2 Za = s[swz(k)], Zb = s[swz(N - k)];
                const double2 E = make_double2(0.5 * (Za.x + Zb.x), 0.5 * (Za.y - Zb.y));
                const double2 O = make_double2(0.5 * (Za.y + Zb.y), 0.5 * (Zb.x - Za.x));
                const double2 T = cmul(cmul(w0, w16th(b)), O);   // W^k O, k = lt + b N/8 = lt + b M/16
                const double2 Xk = cadd(E, T), Xm = cconj(csub(E, T));   // X[k], X[N-k]
                if (field == 0) { out[k] = Xk; out[N - k] = Xm; }
                else { out[M - k] = Xk; out[M - N + k] = Xm; }
            }
        }
        __syncthreads();
    }
    peer_store_fence(pushed);
}

template <int LOG2N>
__global__ void __launch_bounds__(FftLaunch<LOG2N>::THREADS, FftLaunch<LOG2N>::MINB)
k4_rfft_inverse(const FftArgs a, int ngroups_per_member, int ngroups_total, int pf) {
    using L = FftLaunch<LOG2N>;
    using F = RowFft<LOG2N, +1, true>;
    extern __shared__ __align__(16) double2 fft_smem[];
    constexpr int N = L::N, TPR = L::TPR, M = 2 * N;
    const int lt = threadIdx.x;
    double2* s = fft_smem;
    F fft;
    fft.init(a.pl.tw, lt, 2, fft_smem + N);
    const double2 w0c = cconj(__ldg(a.pl.tw + lt));   // W^-lt
    const int P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    __shared__ double gsh[32];
    int gmember = -1;
    double gauge = 0.0;

    for (int grp = blockIdx.x; grp < ngroups_total; grp += gridDim.x) {
        const int member = grp / ngroups_per_member;
        const int rl = grp - member * ngroups_per_member;
        const int row = rl >> 1, layer = rl & 1;
        const double P0 = a.A[2 * layer], P1 = a.A[2 * layer + 1];   // row `layer` of P
        if (member != gmember) {
            gauge = load_gauge(a, member, gsh);
            gmember = member;
        }
        const double2* __restrict__ in =
            reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
        // Pre-processing by Hermitian pairs: the points k and N-k need the same four spectral values,
        // so a thread forms both V[k] = E + iO and V[N-k] = conj(E) + i conj(O) from one set of loads
        // (half the L2 traffic of loading per point) and stages them in shared memory for pass 0.
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = lt + t * TPR;   // 0 .. N/2-1
            if (k == 0) {
                const double2 s0 = __ldg(in), sN = __ldg(in + N);   // (U1[0], U2[0]), (U1[N], U2[N])
                const double X0 = P0 * (s0.x - gauge) + P1 * s0.y, XN = P0 * sN.x + P1 * sN.y;
                s[swz(0)] = make_double2(X0 + XN, X0 - XN);
                // the self-paired point N/2: V = 2 conj(X[N/2])
                const double2 h1 = __ldg(in + N / 2), h2 = __ldg(in + M - N / 2);
                s[swz(N / 2)] = make_double2(2.0 * (P0 * h1.x + P1 * h2.x), -2.0 * (P0 * h1.y + P1 * h2.y));
            } else {
                // X[k] = P0 U1[k] + P1 U2[k];  U1[k] = slot k, U2[k] = slot M-k
                const double2 a1 = __ldg(in + k), a2 = __ldg(in + M - k);
                const double2 b1 = __ldg(in + N - k), b2 = __ldg(in + N + k);
                const double2 Xk = make_double2(P0 * a1.x + P1 * a2.x, P0 * a1.y + P1 * a2.y);
                const double2 Xm = make_double2(P0 * b1.x + P1 * b2.x, P0 * b1.y + P1 * b2.y);   // X[N-k]
                const double2 E = make_double2(Xk.x + Xm.x, Xk.y - Xm.y);                        // X[k] + conj X[N-k]
                const double2 D = make_double2(Xk.x - Xm.x, Xk.y + Xm.y);                        // X[k] - conj X[N-k]
                const double2 O = cmul(cmul(w0c, cconj(w16th(t))), D);                           // W^-k (..)
                s[swz(k)] = make_double2(E.x - O.y, E.y + O.x);                                  // E + i O
                s[swz(N - k)] = make_double2(E.x + O.y, O.x - E.y);                              // conj(E) + i conj(O)
            }
        }
        __syncthreads();
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = s[swz(lt + t * TPR)];
        __syncthreads();   // pass 0 overwrites the buffer
        if (pf) {   // next (row, layer) of this CTA -> L2: a single resident CTA cannot hide the DRAM round trip
            const int gn = grp + gridDim.x;
            if (gn < ngroups_total) {
                const int mn = gn / ngroups_per_member;
                const int rn = (gn - mn * ngroups_per_member) >> 1;
                prefetch_row_l2(a.S + mn * a.sstride + (int64_t)rn * a.pl.ncol, M * 16, lt, TPR);
            }
        }
        fft.template run<false>(v, s, lt);
        double* __restrict__ p = (layer == 0 ? a.psi1 : a.psi2) + member * a.mstride;
        const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
        double* __restrict__ plo = a.pimg_lo + member * a.mstride + layer * a.g.fstride;
        double* __restrict__ phi = a.pimg_hi + member * a.mstride + layer * a.g.fstride;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = F::out_index(lt, e);        // z[n] = (psi[2n], psi[2n+1])
            const int64_t o = a.g.at(2 * n, row);
            const double2 z = v[e];
            *reinterpret_cast<double2*>(p + o) = z;
            const bool gl = n == 0, gr = n == N - 1;   // columns 0,1 / M-2,M-1 feed the x ghosts
            if (gl) *reinterpret_cast<double2*>(p + o + M) = z;
            if (gr) *reinterpret_cast<double2*>(p + o - M) = z;
            if (gb) {
                *reinterpret_cast<double2*>(plo + o + dyo) = z;
                if (gl) *reinterpret_cast<double2*>(plo + o + dyo + M) = z;
                if (gr) *reinterpret_cast<double2*>(plo + o + dyo - M) = z;
            }
            if (gt) {
                *reinterpret_cast<double2*>(phi + o - dyo) = z;
                if (gl) *reinterpret_cast<double2*>(phi + o - dyo + M) = z;
                if (gr) *reinterpret_cast<double2*>(phi + o - dyo - M) = z;
            }
        }
        if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// M = 16384, second generation: one (row, field) / (row, layer) transform per CLUSTER OF TWO CTAs.
//
// The single-CTA kernels above keep a whole 128 KB row in one CTA: one CTA of 1024 threads per SM,
// five block-wide exchanges per row and nothing to overlap them with (ncu, profiles/r01i: issue slots
// 30 % busy, 5.5 warps per issue stalled at the barrier, 5.8 on the row loads; 0.47 / 0.43 of the HBM
// roofline).  Here the length-N complex transform (N = M/2 = 8192, z[n] = x[2n] + i x[2n+1]) is split
// by ONE decimation-in-frequency step across a pair of CTAs,
//       y0[n] = z[n] + z[n+H],   y1[n] = (z[n] - z[n+H]) W_N^n,   n < H = N/2,
//       Z[2k] = FFT_H(y0)[k],    Z[2k+1] = FFT_H(y1)[k],
// so each CTA runs the 4096-point radix-16 engine of the headline grid on 64 KB of shared memory with
// 256 threads, two CTAs (of different clusters, i.e. different rows, out of phase) share an SM, and a
// row costs 3 exchanges of 64 KB per CTA instead of 5 of 128 KB.  The cross step never touches
// shared memory: CTA c loads z[n], z[n+H] for the n of ITS half of [0, H) straight into registers
// (so every input element is loaded exactly once per cluster), forms y0 and y1 there, keeps y_c and
// hands y_(1-c) to the SAME thread of the partner CTA through distributed shared memory (st.async into a
// 32 KB staging area, completion counted on the receiver's mbarrier; 32 KB each way per row) - after
// which every thread holds exactly the 16 inputs of its pass-0 butterfly.  The real-transform split
// X[k] = E[k] + W_M^k O[k] pairs Z[k] with Z[N-k], which have the same parity: it stays CTA-local.  The
// inverse is the mirror image: local pre-processing and inverse FFT_H, then z[n] = a[n] + W_N^-n b[n],
// z[n+H] = a[n] - W_N^-n b[n] after the same thread-to-thread exchange.  Same spectral layout, same
// projections (physical space before the forward, spectral space before the inverse transform).
// ---------------------------------------------------------------------------------------
constexpr int PAIR_H = 4096, PAIR_TPR = 256;
constexpr size_t PAIR_BUF_BYTES = (size_t)PAIR_H * sizeof(double2);              // 64 KB: the local 4096-point transform
constexpr size_t PAIR_STAGE_BYTES = (size_t)8 * PAIR_TPR * sizeof(double2);      // 32 KB: values handed in by the partner
constexpr size_t PAIR_SMEM = PAIR_BUF_BYTES + PAIR_STAGE_BYTES + 64;

// exp(-2 pi i b / 32), b = 0..7
__device__ __forceinline__ double2 w32nd(int b) {
    switch (b) {
        case 0: return make_double2(1.0, 0.0);
        case 1: return make_double2(0.98078528040323044913, -0.19509032201612826785);
        case 2: return make_double2(0.92387953251128675613, -0.38268343236508977173);
        case 3: return make_double2(0.83146961230254523708, -0.55557023301960222474);
        case 4: return make_double2(0.70710678118654752440, -0.70710678118654752440);
        case 5: return make_double2(0.55557023301960222474, -0.83146961230254523708);
        case 6: return make_double2(0.38268343236508977173, -0.92387953251128675613);
        default: return make_double2(0.19509032201612826785, -0.98078528040323044913);
    }
}

__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_c(uint32_t raddr, double2 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "l"(__double_as_longlong(v.x)), "l"(__double_as_longlong(v.y)), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The exchange of one unit: send the eight values `snd` to the same thread of the partner CTA, receive
// the partner's eight into `rcv`.  `it` = this CTA's unit counter (both CTAs of a cluster count alike).
struct PairLink {
    double2* stage;          // [8][256], written by the partner
    uint64_t* recv;          // counts the partner's bytes (one phase per unit)
    uint64_t* freeb;         // the partner has read what I sent for the previous unit (8 warp arrivals per phase)
    uint32_t r_stage, r_recv, r_free;   // the partner's copies, as shared::cluster addresses
    __device__ __forceinline__ void init(unsigned char* smem_tail, uint32_t partner) {
        stage = reinterpret_cast<double2*>(smem_tail);
        recv = reinterpret_cast<uint64_t*>(smem_tail + PAIR_STAGE_BYTES);
        freeb = recv + 1;
        if (threadIdx.x == 0) {
            mbar_init(recv, 1);
            mbar_init(freeb, PAIR_TPR / 32);
        }
        r_stage = cluster_map(smem_u32(stage), partner);
        r_recv = cluster_map(smem_u32(recv), partner);
        r_free = cluster_map(smem_u32(freeb), partner);
    }
    __device__ __forceinline__ void arm() {   // thread 0, once per unit, before waiting
        mbar_expect_tx(recv, (uint32_t)PAIR_STAGE_BYTES);
    }
    __device__ __forceinline__ void wait_partner_ready(uint32_t it) {   // before the first send of unit `it`
        if (it > 0) mbar_wait(freeb, (it - 1) & 1u);
    }
    __device__ __forceinline__ void send(int j, int lt, double2 v) {
        st_async_c(r_stage + (uint32_t)((j * PAIR_TPR + lt) * sizeof(double2)), v, r_recv);
    }
    __device__ __forceinline__ void receive(uint32_t it, int lt, double2 (&rcv)[8]) {
        mbar_wait(recv, it & 1u);
#pragma unroll
        for (int j = 0; j < 8; ++j) rcv[j] = stage[j * PAIR_TPR + lt];
        __syncwarp();
        if ((lt & 31) == 0) remote_arrive(r_free);   // this warp's slots of my staging area may be overwritten
    }
};

// C = rank of the CTA in its cluster, a compile-time constant so that the register arrays are indexed statically
template <int C>
__device__ __forceinline__ void pair_forward_body(const FftArgs& a, int units_per_member, int units_total, int pf,
                                                  unsigned char* pair_raw) {
    using F = RowFft16<12, -1>;
    constexpr int H = PAIR_H, TPR = PAIR_TPR, N = 2 * H, M = 2 * N;
    double2* s = reinterpret_cast<double2*>(pair_raw);
    const int lt = threadIdx.x;
    constexpr uint32_t c = C;
    const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
    PairLink link;
    link.init(pair_raw + PAIR_BUF_BYTES, c ^ 1u);
    F fft;
    fft.init(a.pl.tw, lt, 4);                               // exp(-2 pi i n / H) = tw[4n]
    const double2 wn = __ldg(a.pl.tw + 2 * lt);             // W_N^lt
    const double2 wdif = c ? make_double2(wn.y, -wn.x) : wn;   // W_N^(lt + c H/2) = (-i)^c W_N^lt
    const double2 wsp = __ldg(a.pl.tw + 2 * lt + c);        // W_M^(2 lt + c)
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)
    __syncthreads();
    cluster_sync_all();   // both CTAs' mbarriers exist before anybody sends

    uint32_t it = 0;
    for (int u = cid; u < units_total; u += ncl, ++it) {
        const int member = u / units_per_member;
        const int rf = u - member * units_per_member;
        const int row = rf >> 1, field = rf & 1;
        const double A0 = a.A[2 * field], A1 = a.A[2 * field + 1];   // row `field` of P_inv
        const double2* __restrict__ q1 = reinterpret_cast<const double2*>(a.q1 + member * a.mstride + a.g.at(0, row));
        const double2* __restrict__ q2 = reinterpret_cast<const double2*>(a.q2 + member * a.mstride + a.g.at(0, row));
        if (lt == 0) link.arm();
        double2 v[16];
        const int nb = (int)c * (H / 2) + lt;   // n_j = nb + 256 j
        link.wait_partner_ready(it);            // (long since true: the partner read my previous values a row ago)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = nb + j * TPR;
            const double2 x1l = __ldg(q1 + n), x2l = __ldg(q2 + n), x1h = __ldg(q1 + n + H), x2h = __ldg(q2 + n + H);
            const double2 zl = make_double2(A0 * x1l.x + A1 * x2l.x, A0 * x1l.y + A1 * x2l.y);   // (q~[2n], q~[2n+1])
            const double2 zh = make_double2(A0 * x1h.x + A1 * x2h.x, A0 * x1h.y + A1 * x2h.y);
            const double2 y0 = cadd(zl, zh);
            const double2 y1 = cmul(csub(zl, zh), cmul(wdif, w32nd(j)));
            if (c == 0) { v[j] = y0; link.send(j, lt, y1); } else { v[8 + j] = y1; link.send(j, lt, y0); }
        }
        if (pf) {   // the next unit of this cluster -> L2
            const int un = u + ncl;
            if (un < units_total) {
                const int mn = un / units_per_member;
                const int rn = (un - mn * units_per_member) >> 1;
                const char* b1 = reinterpret_cast<const char*>(a.q1 + mn * a.mstride + a.g.at(0, rn)) + (size_t)c * (H / 2) * 16;
                const char* b2 = reinterpret_cast<const char*>(a.q2 + mn * a.mstride + a.g.at(0, rn)) + (size_t)c * (H / 2) * 16;
                prefetch_row_l2(b1, (H / 2) * 16, lt, TPR);
                prefetch_row_l2(b1 + (size_t)H * 16, (H / 2) * 16, lt, TPR);
                prefetch_row_l2(b2, (H / 2) * 16, lt, TPR);
                prefetch_row_l2(b2 + (size_t)H * 16, (H / 2) * 16, lt, TPR);
            }
        }
        {
            double2 rcv[8];
            link.receive(it, lt, rcv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (c == 0) v[8 + j] = rcv[j]; else v[j] = rcv[j];
            }
        }
        fft.template run<true>(v, s, lt);   // Z[2k + c] = s[k]

        // real-transform split on the pairs (kappa, N - kappa), kappa = 2k + c: local partner k' = H - k - c
        double2* __restrict__ out = reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
        double* __restrict__ outs = reinterpret_cast<double*>(out);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = lt + b * TPR;            // 0 .. H/2 - 1
            const int kap = 2 * k + (int)c;
            if (c == 0 && k == 0) {
                const double2 Z0 = s[swz16(0)], Zh = s[swz16(H / 2)];   // Z[0], Z[N/2]
                outs[0 + field] = Z0.x + Z0.y;              // X[0]  -> slot 0, component `field`
                outs[2 * N + field] = Z0.x - Z0.y;          // X[N]  -> slot M/2
                const double2 Xh = cconj(Zh);               // X[N/2]
                if (field == 0) out[N / 2] = Xh; else out[M - N / 2] = Xh;
                if (field == 0) store_col0(a, member, row, Z0.x + Z0.y, pushed);
            } else {
                const double2 Za = s[swz16(k)], Zb = s[swz16(H - k - (int)c)];
                const double2 E = make_double2(0.5 * (Za.x + Zb.x), 0.5 * (Za.y - Zb.y));
                const double2 O = make_double2(0.5 * (Za.y + Zb.y), 0.5 * (Zb.x - Za.x));
                const double2 T = cmul(cmul(wsp, w32nd(b)), O);          // W_M^kappa O
                const double2 Xk = cadd(E, T), Xm = cconj(csub(E, T));   // X[kappa], X[N - kappa]
                if (field == 0) { out[kap] = Xk; out[N - kap] = Xm; }
                else { out[M - kap] = Xk; out[M - N + kap] = Xm; }
            }
        }
        __syncthreads();   // the row buffer is reused by the next unit
    }
    peer_store_fence(pushed);
    cluster_sync_all();   // nobody leaves while the partner may still write to it
}

__global__ void __launch_bounds__(PAIR_TPR, 2)
k2_rfft_pair(const FftArgs a, int units_per_member, int units_total, int pf) {
    extern __shared__ __align__(128) unsigned char pair_raw[];
    if (cluster_rank() == 0) pair_forward_body<0>(a, units_per_member, units_total, pf, pair_raw);
    else pair_forward_body<1>(a, units_per_member, units_total, pf, pair_raw);
}

template <int C>
__device__ __forceinline__ void pair_inverse_body(const FftArgs& a, int units_per_member, int units_total, int pf,
                                                  unsigned char* pair_raw, double* gsh) {
    using F = RowFft16<12, +1>;
    constexpr int H = PAIR_H, TPR = PAIR_TPR, N = 2 * H, M = 2 * N;
    double2* s = reinterpret_cast<double2*>(pair_raw);
    const int lt = threadIdx.x;
    constexpr uint32_t c = C;
    const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
    PairLink link;
    link.init(pair_raw + PAIR_BUF_BYTES, c ^ 1u);
    F fft;
    fft.init(a.pl.tw, lt, 4);
    const double2 wnc = cconj(__ldg(a.pl.tw + 2 * lt));         // W_N^-lt
    const double2 wout = c ? make_double2(-wnc.y, wnc.x) : wnc;  // W_N^-(lt + c H/2) = i^c W_N^-lt
    const double2 wpre = cconj(__ldg(a.pl.tw + 2 * lt + c));    // W_M^-(2 lt + c)
    const int P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    int gmember = -1;
    double gauge = 0.0;
    __syncthreads();
    cluster_sync_all();

    uint32_t it = 0;
    for (int u = cid; u < units_total; u += ncl, ++it) {
        const int member = u / units_per_member;
        const int rl = u - member * units_per_member;
        const int row = rl >> 1, layer = rl & 1;
        const double P0 = a.A[2 * layer], P1 = a.A[2 * layer + 1];   // row `layer` of P
        if (member != gmember) {
            gauge = load_gauge(a, member, gsh);
            gmember = member;
        }
        if (lt == 0) link.arm();
        const double2* __restrict__ in =
            reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
        // pre-processing by Hermitian pairs (kappa, N - kappa), kappa = 2k + c, see k4_rfft_inverse:
        // V[kappa] -> s[k], V[N - kappa] -> s[H - k - c]
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int k = lt + b * TPR;
            const int kap = 2 * k + (int)c;
            if (c == 0 && k == 0) {
                const double2 s0 = __ldg(in), sN = __ldg(in + N);   // (U1[0], U2[0]), (U1[N], U2[N])
                const double X0 = P0 * (s0.x - gauge) + P1 * s0.y, XN = P0 * sN.x + P1 * sN.y;
                s[swz16(0)] = make_double2(X0 + XN, X0 - XN);
                const double2 h1 = __ldg(in + N / 2), h2 = __ldg(in + M - N / 2);   // the self-paired point N/2
                s[swz16(H / 2)] = make_double2(2.0 * (P0 * h1.x + P1 * h2.x), -2.0 * (P0 * h1.y + P1 * h2.y));
            } else {
                const double2 a1 = __ldg(in + kap), a2 = __ldg(in + M - kap);
                const double2 b1 = __ldg(in + N - kap), b2 = __ldg(in + N + kap);
                const double2 Xk = make_double2(P0 * a1.x + P1 * a2.x, P0 * a1.y + P1 * a2.y);
                const double2 Xm = make_double2(P0 * b1.x + P1 * b2.x, P0 * b1.y + P1 * b2.y);   // X[N - kappa]
                const double2 E = make_double2(Xk.x + Xm.x, Xk.y - Xm.y);
                const double2 D = make_double2(Xk.x - Xm.x, Xk.y + Xm.y);
                const double2 O = cmul(cmul(wpre, cconj(w32nd(b))), D);                          // W_M^-kappa (..)
                s[swz16(k)] = make_double2(E.x - O.y, E.y + O.x);                                // E + i O
                s[swz16(H - k - (int)c)] = make_double2(E.x + O.y, O.x - E.y);                   // conj(E) + i conj(O)
            }
        }
        __syncthreads();
        double2 v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = s[swz16(lt + t * TPR)];
        __syncthreads();   // pass 0 overwrites the buffer
        if (pf) {   // the next unit of this cluster -> L2 (both CTAs need every line of the spectral row)
            const int un = u + ncl;
            if (un < units_total) {
                const int mn = un / units_per_member;
                const int rn = (un - mn * units_per_member) >> 1;
                const char* base = reinterpret_cast<const char*>(a.S + mn * a.sstride + (int64_t)rn * a.pl.ncol);
                prefetch_row_l2(base + (size_t)c * (M * 8), M * 8, lt, TPR);   // each CTA asks for half of the lines
            }
        }
        fft.template run<false>(v, s, lt);   // v[t] = a[lt + 256 t] (c = 0) or b[lt + 256 t] (c = 1)
        // CTA 0 finishes n = lt + 256 j (j < 8), CTA 1 n = lt + 256 (8 + j): hand the other half over
        link.wait_partner_ready(it);
#pragma unroll
        for (int j = 0; j < 8; ++j) link.send(j, lt, c == 0 ? v[8 + j] : v[j]);
        double2 rcv[8];
        link.receive(it, lt, rcv);
        double* __restrict__ p = (layer == 0 ? a.psi1 : a.psi2) + member * a.mstride;
        const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
        double* __restrict__ plo = a.pimg_lo + member * a.mstride + layer * a.g.fstride;
        double* __restrict__ phi = a.pimg_hi + member * a.mstride + layer * a.g.fstride;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 av = c == 0 ? v[j] : rcv[j];
            const double2 bv = c == 0 ? rcv[j] : v[8 + j];
            const double2 wb = cmul(cmul(wout, cconj(w32nd(j))), bv);   // W_N^-n b[n]
            const int n = (int)c * (H / 2) + lt + j * TPR;              // < H
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int nn = n + hh * H;                               // z[nn] = (psi[2 nn], psi[2 nn + 1])
                const double2 z = hh == 0 ? cadd(av, wb) : csub(av, wb);
                const int64_t o = a.g.at(2 * nn, row);
                *reinterpret_cast<double2*>(p + o) = z;
                const bool gl = nn == 0, gr = nn == N - 1;   // columns 0,1 / M-2,M-1 feed the x ghosts
                if (gl) *reinterpret_cast<double2*>(p + o + M) = z;
                if (gr) *reinterpret_cast<double2*>(p + o - M) = z;
                if (gb) {
                    *reinterpret_cast<double2*>(plo + o + dyo) = z;
                    if (gl) *reinterpret_cast<double2*>(plo + o + dyo + M) = z;
                    if (gr) *reinterpret_cast<double2*>(plo + o + dyo - M) = z;
                }
                if (gt) {
                    *reinterpret_cast<double2*>(phi + o - dyo) = z;
                    if (gl) *reinterpret_cast<double2*>(phi + o - dyo + M) = z;
                    if (gr) *reinterpret_cast<double2*>(phi + o - dyo - M) = z;
                }
            }
        }
        if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged
        // no block barrier needed here: the next unit's pre-processing writes the buffer, whose last
        // readers (the pass-2 loads) are separated from it by the barriers inside run()
    }
    cluster_sync_all();
}

__global__ void __launch_bounds__(PAIR_TPR, 2)
k4_rfft_pair(const FftArgs a, int units_per_member, int units_total, int pf) {
    extern __shared__ __align__(128) unsigned char pair_raw[];
    __shared__ double gsh[32];
    if (cluster_rank() == 0) pair_inverse_body<0>(a, units_per_member, units_total, pf, pair_raw, gsh);
    else pair_inverse_body<1>(a, units_per_member, units_total, pf, pair_raw, gsh);
}

// ---------------------------------------------------------------------------------------
// Direct DFT path for M that is not a power of two (or < 8).  One row per block.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k2_dft_forward(const FftArgs a) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int N = a.pl.M, row = blockIdx.x, member = blockIdx.y;
    double2* z = fft_smem;
    double2* Z = fft_smem + N;
    bool pushed = false;   // this thread stored into other ranks' memory (y-slab peer mode)
    const double* __restrict__ q1 = a.q1 + member * a.mstride + a.g.at(0, row);
    const double* __restrict__ q2 = a.q2 + member * a.mstride + a.g.at(0, row);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double x1 = q1[n], x2 = q2[n];
        z[n] = make_double2(a.A[0] * x1 + a.A[1] * x2, a.A[2] * x1 + a.A[3] * x2);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        int idx = 0;
        for (int n = 0; n < N; ++n) {
            acc = cadd(acc, cmul(z[n], twid<-1>(a.pl.tw, idx)));
            idx += k;
            if (idx >= N) idx -= N;
        }
        Z[k] = acc;
    }
    __syncthreads();
    double2* __restrict__ out = reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
    for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
        if (k == 0 || 2 * k == N) {
            out[k] = Z[k];
            if (k == 0) store_col0(a, member, row, Z[0].x, pushed);
        } else {
            const double2 A = Z[k], B = Z[N - k];
            out[k] = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
            out[N - k] = make_double2(0.5 * (A.y + B.y), 0.5 * (B.x - A.x));
        }
    }
    peer_store_fence(pushed);
}

__global__ void __launch_bounds__(256)
k4_dft_inverse(const FftArgs a) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int N = a.pl.M, row = blockIdx.x, member = blockIdx.y;
    double2* Z = fft_smem;
    double2* z = fft_smem + N;
    const double2* __restrict__ in =
        reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
    for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
        if (k == 0 || 2 * k == N) {
            Z[k] = in[k];
        } else {
            const double2 U1 = in[k], U2 = in[N - k];
            Z[k] = make_double2(U1.x - U2.y, U1.y + U2.x);
            Z[N - k] = make_double2(U1.x + U2.y, U2.x - U1.y);
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        int idx = 0;
        for (int k = 0; k < N; ++k) {
            acc = cadd(acc, cmul(Z[k], twid<+1>(a.pl.tw, idx)));
            idx += n;
            if (idx >= N) idx -= N;
        }
        z[n] = acc;
    }
    __syncthreads();
    __shared__ double gsh[32];
    const double gauge = load_gauge(a, member, gsh);
    double* __restrict__ p1 = a.psi1 + member * a.mstride;
    double* __restrict__ p2 = a.psi2 + member * a.mstride;
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    const bool gb = a.pimg_lo != nullptr && row < GHOST, gt = a.pimg_hi != nullptr && row >= P - GHOST;
    double* __restrict__ lo1 = a.pimg_lo + member * a.mstride;
    double* __restrict__ lo2 = lo1 + a.g.fstride;
    double* __restrict__ hi1 = a.pimg_hi + member * a.mstride;
    double* __restrict__ hi2 = hi1 + a.g.fstride;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double t1 = z[n].x - gauge;
        const double o1 = a.A[0] * t1 + a.A[1] * z[n].y;
        const double o2 = a.A[2] * t1 + a.A[3] * z[n].y;
        const int64_t o = a.g.at(n, row);
        const bool gl = n < GHOST, gr = n >= M - GHOST;
        p1[o] = o1; p2[o] = o2;
        if (gl) { p1[o + M] = o1; p2[o + M] = o2; }
        if (gr) { p1[o - M] = o1; p2[o - M] = o2; }
        if (gb) {
            lo1[o + dyo] = o1; lo2[o + dyo] = o2;
            if (gl) { lo1[o + dyo + M] = o1; lo2[o + dyo + M] = o2; }
            if (gr) { lo1[o + dyo - M] = o1; lo2[o + dyo - M] = o2; }
        }
        if (gt) {
            hi1[o - dyo] = o1; hi2[o - dyo] = o2;
            if (gl) { hi1[o - dyo + M] = o1; hi2[o - dyo + M] = o2; }
            if (gr) { hi1[o - dyo - M] = o1; hi2[o - dyo - M] = o2; }
        }
    }
    if ((gb | gt) && !a.periodic_y) __threadfence_system();   // peer stores acknowledged
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int LOG2N, bool FWD>
static cudaError_t launch_pow2(Handle* h, const FftArgs& a) {
    using L = FftLaunch<LOG2N>;
    auto kern = FWD ? k2_fft_forward<LOG2N> : k4_fft_inverse<LOG2N>;
    static bool configured_dev[QG_MAX_DEVICES] = {};
    static int blocks_per_sm_dev[QG_MAX_DEVICES] = {};
    bool& configured = configured_dev[dev_slot(h)];
    int& blocks_per_sm = blocks_per_sm_dev[dev_slot(h)];
    if (!configured) {
        if (L::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, L::THREADS, L::SMEM);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const int gpm = (h->plan.P + L::RPB - 1) / L::RPB;
    const int total = gpm * h->nm;
    int grid = num_sms() * blocks_per_sm;
    if (grid > total) grid = total;
    kern<<<grid, L::THREADS, L::SMEM, h->stream>>>(a, gpm, total);
    return cudaGetLastError();
}

template <int LOG2N, bool FWD>
static cudaError_t launch_r16(Handle* h, const FftArgs& a) {
    using L = Fft16Launch<LOG2N>;
    auto kern = FWD ? k2_fft16_forward<LOG2N> : k4_fft16_inverse<LOG2N>;
    static bool configured_dev[QG_MAX_DEVICES] = {};
    static int blocks_per_sm_dev[QG_MAX_DEVICES] = {};
    bool& configured = configured_dev[dev_slot(h)];
    int& blocks_per_sm = blocks_per_sm_dev[dev_slot(h)];
    if (!configured) {
        if (L::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, L::THREADS, L::SMEM);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const int gpm = (h->plan.P + L::RPB - 1) / L::RPB;
    const int total = gpm * h->nm;
    int grid = num_sms() * blocks_per_sm;
    if (grid > total) grid = total;
    static const int pf = getenv("QG_FFT_PF") ? atoi(getenv("QG_FFT_PF")) : 1;
    kern<<<grid, L::THREADS, L::SMEM, h->stream>>>(a, gpm, total, pf);
    return cudaGetLastError();
}

template <int LOG2N, bool FWD>
static cudaError_t launch_long(Handle* h, const FftArgs& a) {
    using L = FftLaunch<LOG2N>;
    auto kern = FWD ? k2_rfft_forward<LOG2N> : k4_rfft_inverse<LOG2N>;
    // row buffer + the per-thread base twiddles of the radix-8 passes (RowFft<.., W8S = true>)
    constexpr size_t smem = L::SMEM + (size_t)(LOG2N / 3 - 1) * L::THREADS * sizeof(double2);
    static bool configured_dev[QG_MAX_DEVICES] = {};
    static int blocks_per_sm_dev[QG_MAX_DEVICES] = {};
    bool& configured = configured_dev[dev_slot(h)];
    int& blocks_per_sm = blocks_per_sm_dev[dev_slot(h)];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, L::THREADS, smem);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const int gpm = 2 * h->plan.P;   // (row, field) or (row, layer)
    const int total = gpm * h->nm;
    int grid = num_sms() * blocks_per_sm;
    if (grid > total) grid = total;
    static const int pf = getenv("QG_FFT_PF") ? atoi(getenv("QG_FFT_PF")) : 1;
    kern<<<grid, L::THREADS, smem, h->stream>>>(a, gpm, total, pf);
    return cudaGetLastError();
}

template <bool FWD>
static cudaError_t launch_pair(Handle* h, const FftArgs& a) {
    auto kern = FWD ? k2_rfft_pair : k4_rfft_pair;
    static bool configured_dev[QG_MAX_DEVICES] = {};
    static int clusters_dev[QG_MAX_DEVICES] = {};
    bool& configured = configured_dev[dev_slot(h)];
    int& nclusters = clusters_dev[dev_slot(h)];
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(PAIR_TPR, 1, 1);
    cfg.dynamicSmemBytes = PAIR_SMEM;
    cfg.stream = h->stream;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM);
        if (e != cudaSuccess) return e;
        cfg.gridDim = dim3(2 * 2 * num_sms(), 1, 1);
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < 1) { (void)cudaGetLastError(); ncl = num_sms(); }
        nclusters = ncl;
        configured = true;
        if (getenv("QG_VERBOSE")) fprintf(stderr, "qgb200: long-row transforms: %d resident clusters of 2 CTAs\n", ncl);
    }
    const int upm = 2 * h->plan.P;   // (row, field) or (row, layer)
    const int total = upm * h->nm;
    const int ncl = nclusters < total ? nclusters : total;
    cfg.gridDim = dim3(2 * ncl, 1, 1);
    static const int pf = getenv("QG_FFT_PF") ? atoi(getenv("QG_FFT_PF")) : 1;
    return cudaLaunchKernelEx(&cfg, kern, a, upm, total, pf);
}

template <bool FWD>
static cudaError_t launch_ring(Handle* h, const FftArgs& a) {
    static bool configured_dev[QG_MAX_DEVICES] = {};
    bool& configured = configured_dev[dev_slot(h)];
    if (!configured) {
        cudaError_t e = FWD ? cudaFuncSetAttribute(k2_fft16_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RING_SMEM)
                            : cudaFuncSetAttribute(k4_fft16_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RING_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int total = h->plan.P * h->nm;
    int grid = num_sms();
    if (grid > (total + 1) / 2) grid = (total + 1) / 2;   // both groups of every CTA get a row
    // QG_RING_PF = n: also pull the row n turns of the ring ahead into L2 with one cp.async.bulk.prefetch.L2.
    // Measured at 4096^2: 102.8 us without, 105.4 us with n = 3, 124 us with n = 6 - off by default.
    static const int pf = getenv("QG_RING_PF") ? atoi(getenv("QG_RING_PF")) : 0;
    if (FWD) k2_fft16_ring<<<grid, RING_THREADS, RING_SMEM, h->stream>>>(a, h->plan.P, total, pf);
    else k4_fft16_ring<<<grid, RING_THREADS, RING_SMEM, h->stream>>>(a, h->plan.P, total);
    return cudaGetLastError();
}

template <bool FWD>
static cudaError_t dispatch_pow2(Handle* h, const FftArgs& a) {
    static const bool radix8_only = getenv("QG_FFT_RADIX8") != nullptr;
    // QG_FFT_RING: 0 = round-1 kernels for both directions, 1 (default) = ring-buffered forward transform,
    // 2 = ring-buffered in both directions.  Measured at 4096^2: forward 118 -> 103 us; the inverse is
    // slower in the ring (117 vs 111 us): re-tangling the Hermitian pairs out of shared memory costs 128 KB
    // of extra shared-memory reads per row, which the direct (L2-prefetched) global loads do not.
    static const int ring = getenv("QG_FFT_RING") ? atoi(getenv("QG_FFT_RING")) : 1;
    if (!radix8_only) {
        if (h->plan.log2M == 12 && (ring >= 2 || (ring == 1 && FWD))) return launch_ring<FWD>(h, a);
        if (h->plan.log2M == 12) return launch_r16<12, FWD>(h, a);
        if (h->plan.log2M == 8) return launch_r16<8, FWD>(h, a);
    }
    switch (h->plan.log2M) {
        case 3: return launch_pow2<3, FWD>(h, a);
        case 4: return launch_pow2<4, FWD>(h, a);
        case 5: return launch_pow2<5, FWD>(h, a);
        case 6: return launch_pow2<6, FWD>(h, a);
        case 7: return launch_pow2<7, FWD>(h, a);
        case 8: return launch_pow2<8, FWD>(h, a);
        case 9: return launch_pow2<9, FWD>(h, a);
        case 10: return launch_pow2<10, FWD>(h, a);
        case 11: return launch_pow2<11, FWD>(h, a);
        case 12: return launch_pow2<12, FWD>(h, a);
        case 13: return launch_pow2<13, FWD>(h, a);
        case 14: {   // M = 16384: half-length real transform per single CTA; QG_FFT_PAIR=1: per cluster pair
            // (measured on 16384 x 2048, profiles/r02/ab_r02c.log: pair 370 / 371 us vs 338 / 366 us - the pair
            // halves the shared-memory traffic and the barrier stalls, but its 16 warps per SM hide the row
            // loads worse than the 32 of the single-CTA kernel and the two CTAs meet once per row)
            static const bool pair = getenv("QG_FFT_PAIR") && atoi(getenv("QG_FFT_PAIR")) != 0;
            return pair ? launch_pair<FWD>(h, a) : launch_long<13, FWD>(h, a);
        }
        default: return cudaErrorInvalidValue;
    }
}

static cudaError_t set_smem(const void* fn, size_t bytes) {
    if (bytes > 48 * 1024)
        return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

cudaError_t launch_fft_forward(Handle* h, const double* q_fields, int /*which*/) {
    FftArgs a{};
    a.g = h->g;
    a.pl = h->plan;
    a.q1 = q_fields;
    a.q2 = q_fields + h->g.fstride;
    a.S = h->S;
    a.sstride = (int64_t)h->plan.P * h->plan.ncol;
    a.mstride = 2 * h->g.fstride;
    for (int i = 0; i < 4; ++i) a.A[i] = h->prm.Pinv[i];
    a.scal = h->scal;
    a.col0 = h->col0;
    if (h->peer_ok) {   // y-slab peer mode: the k=0 column is gathered by the stores themselves
        a.col0_n = h->dist_n;
        a.col0_off = h->dist_rank * h->plan.P;
        for (int r = 0; r < h->dist_n; ++r) a.col0_peer[r] = h->peer_mail[r] + 256;
    }
    KernelTimer t(h, QG_K_FFT_FWD);
    if (h->plan.pow2) return dispatch_pow2<true>(h, a);
    const size_t smem = 2 * (size_t)h->plan.M * sizeof(double2);
    cudaError_t e = set_smem((const void*)k2_dft_forward, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(h->plan.P, h->nm);
    k2_dft_forward<<<grid, 256, smem, h->stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fft_inverse(Handle* h, double* psi_fields, int use_gauge) {
    FftArgs a{};
    a.g = h->g;
    a.pl = h->plan;
    a.psi1 = psi_fields;
    a.psi2 = psi_fields + h->g.fstride;
    a.S = h->S;
    a.sstride = (int64_t)h->plan.P * h->plan.ncol;
    a.mstride = 2 * h->g.fstride;
    for (int i = 0; i < 4; ++i) a.A[i] = h->prm.Pfwd[i];
    a.scal = h->scal;
    a.use_gauge = use_gauge;
    a.periodic_y = h->dist_n > 1 ? 0 : 1;
    if (h->dist_n == 1) {
        a.pimg_lo = a.pimg_hi = a.psi1;
    } else if (h->peer_ok && psi_fields >= h->psi && psi_fields < h->psi + (int64_t)h->nfields * h->g.fstride) {
        // rows [0,2) -> the rank below, rows [P-2,P) -> the rank above (periodic ring)
        const int64_t off = a.psi1 - h->psi;
        a.pimg_lo = h->peer_psi[(h->dist_rank + h->dist_n - 1) % h->dist_n] + off;
        a.pimg_hi = h->peer_psi[(h->dist_rank + 1) % h->dist_n] + off;
    }
    if (h->gauge_parts) {   // K3 left per-slab shares of the gauge instead of running k3_gauge
        a.gpart = h->gpart;
        a.ngp = h->plan.ngp;
    }
    KernelTimer t(h, QG_K_FFT_INV);
    if (h->plan.pow2) return dispatch_pow2<false>(h, a);
    const size_t smem = 2 * (size_t)h->plan.M * sizeof(double2);
    cudaError_t e = set_smem((const void*)k4_dft_inverse, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(h->plan.P, h->nm);
    k4_dft_inverse<<<grid, 256, smem, h->stream>>>(a);
    return cudaGetLastError();
}

}  // namespace qg
