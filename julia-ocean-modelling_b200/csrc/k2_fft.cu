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
// through XOR-swizzled shared memory, the first pass fed straight from global memory.
// Twiddles come from a table evaluated in extended precision on the host.
// Non-power-of-two M (the reference benchmarks M = 8:8:128) takes a direct O(M^2) DFT
// path with the same spectral layout.
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
    // odd half: multiply by W8^t, W8 = exp(SIGN * i pi / 4)
    {
        // W8^1 = h (1 + SIGN i)
        const double2 t5 = muli<SIGN>(b5);
        b5 = make_double2(h * (b5.x + t5.x), h * (b5.y + t5.y));
        b6 = muli<SIGN>(b6);
        // W8^3 = h (-1 + SIGN i)
        const double2 t7 = muli<SIGN>(b7);
        b7 = make_double2(h * (t7.x - b7.x), h * (t7.y - b7.y));
    }
    bfly4<SIGN>(b0, b1, b2, b3);   // even outputs X[0], X[2], X[4], X[6]
    bfly4<SIGN>(b4, b5, b6, b7);   // odd outputs  X[1], X[3], X[5], X[7]
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = b4; a[3] = b5; a[5] = b6; a[7] = b7;
}

// Stockham FFT of one row of N = 8 * tpr points held 8 per thread: on entry
// v[t] = x[lt + t * tpr]; on exit the transform is in shared memory `s` (swizzled,
// natural order) and a __syncthreads() has been executed.
template <int SIGN>
__device__ __forceinline__ void fft_row(double2 (&v)[8], double2* s, int N, int log2N, int tpr, int lt,
                                        const double2* __restrict__ tw) {
    const int nb8 = log2N / 3;
    const int rem = log2N - 3 * nb8;
    int Ns = 1;
    for (int p = 0; p < nb8; ++p) {
        if (p > 0) {
            __syncthreads();
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = s[swz(lt + t * tpr)];
            __syncthreads();
            const int k = lt & (Ns - 1);
            const int step = N / (Ns * 8);
#pragma unroll
            for (int t = 1; t < 8; ++t) v[t] = cmul(v[t], twid<SIGN>(tw, t * k * step));
        }
        bfly8<SIGN>(v);
        const int k = lt & (Ns - 1);
        const int j0 = (lt - k) * 8 + k;
#pragma unroll
        for (int u = 0; u < 8; ++u) s[swz(j0 + u * Ns)] = v[u];
        Ns *= 8;
    }
    if (rem == 2) {   // one radix-4 pass, two butterflies per thread
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lt + b * tpr;
#pragma unroll
            for (int t = 0; t < 4; ++t) v[b * 4 + t] = s[swz(j + t * 2 * tpr)];
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lt + b * tpr;
            const int k = j & (Ns - 1);
            const int step = N / (Ns * 4);
#pragma unroll
            for (int t = 1; t < 4; ++t) v[b * 4 + t] = cmul(v[b * 4 + t], twid<SIGN>(tw, t * k * step));
            bfly4<SIGN>(v[b * 4 + 0], v[b * 4 + 1], v[b * 4 + 2], v[b * 4 + 3]);
            const int j0 = (j - k) * 4 + k;
#pragma unroll
            for (int u = 0; u < 4; ++u) s[swz(j0 + u * Ns)] = v[b * 4 + u];
        }
    } else if (rem == 1) {   // one radix-2 pass, four butterflies per thread
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = lt + b * tpr;
            v[b * 2 + 0] = s[swz(j)];
            v[b * 2 + 1] = s[swz(j + 4 * tpr)];
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = lt + b * tpr;
            const int k = j & (Ns - 1);
            const int step = N / (Ns * 2);
            const double2 w = cmul(v[b * 2 + 1], twid<SIGN>(tw, k * step));
            const double2 x0 = cadd(v[b * 2], w), x1 = csub(v[b * 2], w);
            const int j0 = (j - k) * 2 + k;
            s[swz(j0)] = x0;
            s[swz(j0 + Ns)] = x1;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Forward kernel, power-of-two M >= 8.  Block = rpb rows x tpr threads, grid = (rows, members).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k2_fft_forward(const FftArgs a) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int N = a.pl.M, tpr = a.pl.tpr;
    const int lr = threadIdx.x / tpr, lt = threadIdx.x - lr * tpr;
    const int row = blockIdx.x * a.pl.rpb + lr;
    const bool live = row < a.pl.P;
    const int member = blockIdx.y;
    double2* s = fft_smem + (size_t)lr * N;
    const double* __restrict__ q1 = a.q1 + member * a.mstride + a.g.at(0, live ? row : 0);
    const double* __restrict__ q2 = a.q2 + member * a.mstride + a.g.at(0, live ? row : 0);

    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int n = lt + t * tpr;
        const double x1 = live ? __ldg(q1 + n) : 0.0, x2 = live ? __ldg(q2 + n) : 0.0;
        v[t] = make_double2(a.A[0] * x1 + a.A[1] * x2, a.A[2] * x1 + a.A[3] * x2);   // src/model.jl:180
    }
    fft_row<-1>(v, s, N, a.pl.log2M, tpr, lt, a.pl.tw);

    if (!live) return;
    double2* __restrict__ out = reinterpret_cast<double2*>(a.S + member * a.sstride + (int64_t)row * a.pl.ncol);
    const int half = N >> 1;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int k = lt + b * tpr;   // 0 .. N/2 - 1
        if (k == 0) {
            out[0] = s[swz(0)];
            out[half] = s[swz(half)];
        } else {
            const double2 A = s[swz(k)], B = s[swz(N - k)];
            out[k] = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));        // Q1[k]
            out[N - k] = make_double2(0.5 * (A.y + B.y), 0.5 * (B.x - A.x));    // Q2[k]
        }
    }
}

// ---------------------------------------------------------------------------------------
// Inverse kernel, power-of-two M >= 8.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k4_fft_inverse(const FftArgs a) {
    extern __shared__ __align__(16) double2 fft_smem[];
    const int N = a.pl.M, tpr = a.pl.tpr;
    const int lr = threadIdx.x / tpr, lt = threadIdx.x - lr * tpr;
    const int row = blockIdx.x * a.pl.rpb + lr;
    const bool live = row < a.pl.P;
    const int member = blockIdx.y;
    double2* s = fft_smem + (size_t)lr * N;
    const double2* __restrict__ in =
        reinterpret_cast<const double2*>(a.S + member * a.sstride + (int64_t)(live ? row : 0) * a.pl.ncol);
    const int half = N >> 1;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int k = lt + b * tpr;
        if (k == 0) {
            s[swz(0)] = __ldg(in);
            s[swz(half)] = __ldg(in + half);
        } else {
            const double2 U1 = __ldg(in + k), U2 = __ldg(in + N - k);
            s[swz(k)] = make_double2(U1.x - U2.y, U1.y + U2.x);        // U1 + i U2
            s[swz(N - k)] = make_double2(U1.x + U2.y, U2.x - U1.y);    // conj(U1) + i conj(U2)
        }
    }
    __syncthreads();
    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = s[swz(lt + t * tpr)];
    __syncthreads();
    fft_row<+1>(v, s, N, a.pl.log2M, tpr, lt, a.pl.tw);

    if (!live) return;
    const double gauge = a.use_gauge ? a.scal[member * 4 + 1] : 0.0;
    double* __restrict__ p1 = a.psi1 + member * a.mstride;
    double* __restrict__ p2 = a.psi2 + member * a.mstride;
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    const bool gb = row < GHOST, gt = row >= P - GHOST;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int n = lt + t * tpr;
        const double2 z = s[swz(n)];
        const double t1 = z.x - gauge;                 // pinned node: psi~1(0,0) = 0
        const double o1 = a.A[0] * t1 + a.A[1] * z.y;  // src/model.jl:196
        const double o2 = a.A[2] * t1 + a.A[3] * z.y;
        const int64_t o = a.g.at(n, row);
        const bool gl = n < GHOST, gr = n >= M - GHOST;
        p1[o] = o1; p2[o] = o2;
        if (gl) { p1[o + M] = o1; p2[o + M] = o2; }
        if (gr) { p1[o - M] = o1; p2[o - M] = o2; }
        if (gb) {
            p1[o + dyo] = o1; p2[o + dyo] = o2;
            if (gl) { p1[o + dyo + M] = o1; p2[o + dyo + M] = o2; }
            if (gr) { p1[o + dyo - M] = o1; p2[o + dyo - M] = o2; }
        }
        if (gt) {
            p1[o - dyo] = o1; p2[o - dyo] = o2;
            if (gl) { p1[o - dyo + M] = o1; p2[o - dyo + M] = o2; }
            if (gr) { p1[o - dyo - M] = o1; p2[o - dyo - M] = o2; }
        }
    }
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
        } else {
            const double2 A = Z[k], B = Z[N - k];
            out[k] = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
            out[N - k] = make_double2(0.5 * (A.y + B.y), 0.5 * (B.x - A.x));
        }
    }
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
    const double gauge = a.use_gauge ? a.scal[member * 4 + 1] : 0.0;
    double* __restrict__ p1 = a.psi1 + member * a.mstride;
    double* __restrict__ p2 = a.psi2 + member * a.mstride;
    const int M = a.g.M, P = a.g.P;
    const int64_t dyo = (int64_t)P * a.g.pitch;
    const bool gb = row < GHOST, gt = row >= P - GHOST;
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
            p1[o + dyo] = o1; p2[o + dyo] = o2;
            if (gl) { p1[o + dyo + M] = o1; p2[o + dyo + M] = o2; }
            if (gr) { p1[o + dyo - M] = o1; p2[o + dyo - M] = o2; }
        }
        if (gt) {
            p1[o - dyo] = o1; p2[o - dyo] = o2;
            if (gl) { p1[o - dyo + M] = o1; p2[o - dyo + M] = o2; }
            if (gr) { p1[o - dyo - M] = o1; p2[o - dyo - M] = o2; }
        }
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
    KernelTimer t(h, QG_K_FFT_FWD);
    if (h->plan.pow2) {
        const size_t smem = (size_t)h->plan.rpb * h->plan.M * sizeof(double2);
        cudaError_t e = set_smem((const void*)k2_fft_forward, smem);
        if (e != cudaSuccess) return e;
        dim3 grid((h->plan.P + h->plan.rpb - 1) / h->plan.rpb, h->nm);
        k2_fft_forward<<<grid, h->plan.rpb * h->plan.tpr, smem, h->stream>>>(a);
    } else {
        const size_t smem = 2 * (size_t)h->plan.M * sizeof(double2);
        cudaError_t e = set_smem((const void*)k2_dft_forward, smem);
        if (e != cudaSuccess) return e;
        dim3 grid(h->plan.P, h->nm);
        k2_dft_forward<<<grid, 256, smem, h->stream>>>(a);
    }
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
    KernelTimer t(h, QG_K_FFT_INV);
    if (h->plan.pow2) {
        const size_t smem = (size_t)h->plan.rpb * h->plan.M * sizeof(double2);
        cudaError_t e = set_smem((const void*)k4_fft_inverse, smem);
        if (e != cudaSuccess) return e;
        dim3 grid((h->plan.P + h->plan.rpb - 1) / h->plan.rpb, h->nm);
        k4_fft_inverse<<<grid, h->plan.rpb * h->plan.tpr, smem, h->stream>>>(a);
    } else {
        const size_t smem = 2 * (size_t)h->plan.M * sizeof(double2);
        cudaError_t e = set_smem((const void*)k4_dft_inverse, smem);
        if (e != cudaSuccess) return e;
        dim3 grid(h->plan.P, h->nm);
        k4_dft_inverse<<<grid, 256, smem, h->stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace qg
