// Developer probe (not part of the library): in-place read-modify-write of a [P][NC] array of
// doubles by column slabs, tile staged in shared memory by TMA.  Variants:
//   ST=0: TMA load -> smem -> per-thread STG;  ST=1: TMA load -> smem update -> TMA store.
// RB = rows per CTA tile (tile bytes = RB * W * 8), box = 32 rows x W columns.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* src, const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <int W, int RB, int ST, int NT>
__global__ void __launch_bounds__(NT) probe(const __grid_constant__ CUtensorMap tm, double* __restrict__ S, int P, int NC) {
    extern __shared__ __align__(128) unsigned char raw[];
    double* tile = reinterpret_cast<double*>(raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + RB * W);
    const int nrb = P / RB;
    const int slab = blockIdx.x / nrb, rb = blockIdx.x % nrb;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, RB * W * 8);
        for (int k = 0; k < RB / 32; ++k) tma_load_2d(tile + k * 32 * W, &tm, slab * W, rb * RB + k * 32, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    if (ST == 0) {
        double* base = S + (size_t)(rb * RB) * NC + slab * W;
        for (int e = tid; e < RB * W; e += NT) {
            const int r = e / W, c = e % W;
            base[(size_t)r * NC + c] = tile[e] + 1.0;
        }
    } else {
        for (int e = tid; e < RB * W; e += NT) tile[e] += 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            for (int k = 0; k < RB / 32; ++k) tma_store_2d(tile + k * 32 * W, &tm, slab * W, rb * RB + k * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

template <int W, int RB, int ST, int NT>
void run(double* S, int P, int NC) {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)NC, (cuuint64_t)P};
    const cuuint64_t strides[1] = {(cuuint64_t)NC * 8};
    const cuuint32_t box[2] = {W, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, S, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int grid = (NC / W) * (P / RB);
    const size_t smem = (size_t)RB * W * 8 + 16;
    auto k = probe<W, RB, ST, NT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) k<<<grid, NT, smem>>>(tm, S, P, NC);
    cudaEventRecord(a);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k<<<grid, NT, smem>>>(tm, S, P, NC);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = 2.0 * P * NC * 8;
    printf("W=%3d RB=%4d (%3zu KB tile) %s NT=%4d occ=%d grid %6d  %.1f us  %.0f GB/s  %s\n", W, RB, smem / 1024,
           ST ? "TMAstore" : "STG     ", NT, occ, grid, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int P = 4096, NC = 8192;
    double* S;
    cudaMalloc(&S, (size_t)P * NC * 8);
    cudaMemset(S, 0, (size_t)P * NC * 8);
    run<16, 512, 0, 256>(S, P, NC);
    run<16, 512, 1, 256>(S, P, NC);
    run<16, 256, 0, 256>(S, P, NC);
    run<16, 256, 1, 256>(S, P, NC);
    run<16, 128, 0, 128>(S, P, NC);
    run<16, 128, 1, 128>(S, P, NC);
    run<16, 64, 0, 128>(S, P, NC);
    run<32, 256, 0, 256>(S, P, NC);
    run<32, 256, 1, 256>(S, P, NC);
    run<32, 128, 0, 256>(S, P, NC);
    run<32, 128, 1, 256>(S, P, NC);
    run<64, 128, 0, 256>(S, P, NC);
    run<64, 128, 1, 256>(S, P, NC);
    run<64, 64, 0, 256>(S, P, NC);
    run<64, 64, 1, 256>(S, P, NC);
    run<128, 32, 0, 256>(S, P, NC);
    run<128, 32, 1, 256>(S, P, NC);
    return 0;
}
