// Developer probe (not part of the library): HBM throughput of a read-modify-write sweep over a
// [P][NC] row-major array of doubles when each CTA owns a column slab of W doubles per row.
// Answers: how wide must the y-solve's slabs be for DRAM to run near its streaming rate?
#include <cstdio>
#include <cuda_runtime.h>

template <int W>   // W doubles per row segment; CTA = 256 threads, 64 KB per CTA
__global__ void __launch_bounds__(256, 3) probe(double* __restrict__ S, int P, int NC) {
    constexpr int ROWS = 8192 / W;          // rows per CTA tile
    constexpr int TPRW = W;                 // threads per row
    constexpr int RPI = 256 / TPRW;         // rows per iteration
    const int nrb = P / ROWS;
    const int slab = blockIdx.x / nrb, rb = blockIdx.x % nrb;
    const int tx = threadIdx.x % TPRW, ty = threadIdx.x / TPRW;
    double* base = S + (size_t)(rb * ROWS + ty) * NC + slab * W + tx;
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = base[(size_t)i * RPI * NC];
#pragma unroll
    for (int i = 0; i < 32; ++i) base[(size_t)i * RPI * NC] = v[i] + 1.0;
}

template <int W>
void run(double* S, int P, int NC) {
    const int ROWS = 8192 / W;
    const int grid = (NC / W) * (P / ROWS);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe<W><<<grid, 256>>>(S, P, NC);
    cudaEventRecord(a);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) probe<W><<<grid, 256>>>(S, P, NC);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = 2.0 * P * NC * 8;
    printf("W=%4d doubles (%5d B/row)  grid %6d  %.1f us  %.0f GB/s  %s\n", W, W * 8, grid, ms / reps * 1e3,
           bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int P = 4096, NC = 8192;
    double* S;
    cudaMalloc(&S, (size_t)P * NC * 8);
    cudaMemset(S, 0, (size_t)P * NC * 8);
    run<16>(S, P, NC);
    run<32>(S, P, NC);
    run<64>(S, P, NC);
    run<128>(S, P, NC);
    run<256>(S, P, NC);
    return 0;
}
