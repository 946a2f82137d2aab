// Internal declarations shared by the CUDA translation units of libqgb200.
// Public surface: include/qgb200.h.  Everything here is sm_100a-only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "qgb200.h"

namespace qg {

// ---- device layout of one field ------------------------------------------------------
// A field is a (P + 2*YPAD) x pitch array of doubles; interior node (i, j) lives at
// (j + YPAD) * pitch + XPAD + i.  XPAD = 16 doubles keeps interior rows 128-byte aligned;
// two ghost columns / rows on every side hold the periodic images (the biharmonic term
// reaches +-2, the reference gets away with one ghost because it refreshes the ghosts of
// the intermediate Laplacian, src/schemes/laplacian.jl:25).
constexpr int XPAD = 16;
constexpr int YPAD = 2;
constexpr int GHOST = 2;

// K1 tile geometry (k1_zeta.cu); the TMA box is (K1_BX, K1_BY, 1).
constexpr int K1_TX = 128;
constexpr int K1_THREADS = 256;

struct Geom {
    int M, P;
    int pitch;        // doubles per row, multiple of 16
    int rows;         // P + 2*YPAD
    int64_t fstride;  // doubles per field = pitch * rows
    __host__ __device__ int64_t at(int i, int j) const {
        return (int64_t)(j + YPAD) * pitch + XPAD + i;
    }
};

// ---- zeta step (K1) -------------------------------------------------------------------
struct ZetaArgs {
    Geom g;
    const double* f1;   // RHS of step n-1, field 0 of this launch (member 0, layer 0)
    const double* f2;   // RHS of step n-2
    double* fn;         // RHS of step n (output)
    double* qn;         // new PV (output)
    int zq, zpsi;       // first field index (tensor-map z coordinate) of q / psi inputs
    int euler;          // 1: steps 1-2 (no history read)
    int periodic_y;     // 1: write the y ghost images locally; 0: y-slab mode
    // Where the images of the first / last GHOST rows of q+ go (field 0 of the output slot, same
    // layout): this rank's own array when the run is periodic in y on one GPU, the ring
    // neighbours' arrays (NVLink peer memory) in y-slab mode, nullptr when NCCL fills the ghosts.
    double* qimg_lo;    // receives rows [0, GHOST) as its rows [P, P+GHOST)
    double* qimg_hi;    // receives rows [P-GHOST, P) as its rows [-GHOST, 0)
    double idx2;        // (1/dx)^2
    double hdx;         // 0.5*(1/dx)
    double i12dx2;      // 1 / (3*4*dx^2)
    double visc, dt;
    double beta[2];     // beta_1, beta_2
    double U, r;
    double c1, c2, c3;  // 23/12, 16/12, 5/12
};

// ---- spectral plan ------------------------------------------------------------------
// Spectral rows hold both modal fields as M complex slots (2M real columns):
// slot 0 = (Q1[0], Q2[0]), slot M/2 = (Q1[M/2], Q2[M/2]) (M even), slot k = Q1[k] and
// slot M-k = Q2[k] for 0 < k < M/2.  Field 1 = Poisson (barotropic), field 2 = Helmholtz.
struct Plan {
    int M, P, ncol;          // ncol = 2M real columns
    int pow2;                // 1: radix-8 Stockham path, 0: direct DFT path
    int log2M;
    int tpr, rpb;            // threads per row, rows per block (pow2 path)
    int C;                   // 32-row chunks in y
    int lenLast;             // rows in the last chunk
    int wpc, CS, m;          // warps per CTA, cluster size, chunks per warp (register variant)
    int ts_ok, ts_CS, ts_nchunk;   // TMA-staged variant: usable, cluster size, chunks per CTA
    int tp_ncl;              // persistent variant: clusters resident at once (0 = not queried yet)
    int tp_ok, tp_boxrows;   // persistent variant usable (P % 32 == 0, <= 8192 rows); rows per TMA box (<= 256)
    int tp_CS, tp_nchunk;    // its cluster size (<= 16, non-portable above 8) and chunks per CTA
    double* coltab;          // [72][ncol] per-column constants of the persistent y-solve (k3_ysolve.cu)
    double2* tw;             // exp(-2 pi i n / M), n < M
    double *rtab, *kap, *rho32, *h32, *rhoL, *hL, *inv1mrP, *pinw, *gw;   // per real column
    double *logr, *g1mr2;    // ln r and 1 / (1 - r^2) per real column (from the extended-precision r): k3_rank_correct
    int2* corr_work;         // y-slab mode: (32-column tile, 32-row segment) pairs k3_rank_correct has to touch
    int ncorr;
    int ngp;                 // gauge partial sums per member (= slabs of the TMA y-solve)
    double k0scale;          // dx^2 / M
};

struct FftArgs {
    Geom g;
    Plan pl;
    const double* q1;  // forward: layer-1 / layer-2 PV fields of member 0
    const double* q2;
    double* psi1;      // inverse: output fields of member 0
    double* psi2;
    double* S;         // spectral rows, member 0
    int64_t sstride;   // doubles per member in S
    int64_t mstride;   // doubles per member in q / psi (= 2 * fstride)
    double A[4];       // forward: P_inv; inverse: P (row-major)
    const double* scal;  // per member: [0] = sum of the Poisson k=0 column, [1] = gauge
    int use_gauge;
    int periodic_y;      // 1: write the y ghost images locally; 0: y-slab mode (halo exchange fills them)
    double* pimg_lo;     // inverse: where the images of the first / last GHOST rows of psi go (layer-1 field
    double* pimg_hi;     // of the output slot): own array, ring neighbours' peer memory, or nullptr (NCCL)
    double* col0_peer[8];   // forward, y-slab peer mode: every rank's gathered k=0 column; col0_n = 0 otherwise
    int col0_n, col0_off;   // number of ranks, offset of this rank's rows in the gathered column
    const double* gpart; // inverse: per-slab shares of the gauge (ngp per member), or nullptr -> scal[1]
    int ngp;
    double* col0;        // forward: compact copy of the Poisson k=0 column, [member * P + row]
};

struct YArgs {
    Plan pl;
    double* S;
    int64_t sstride;
    double* k0sol;      // per member: P doubles, solution of the singular k=0 Poisson column
    double* scal;       // per member: 4 doubles
    int pinned;         // 1: apply the reference's node-(0,0) pin (src/schemes/laplacian.jl:66-75)
    // k=0 Poisson column, compact: col0[member * preP + j]; k0sol has the same indexing
    const double* col0;
    int preP;           // rows of the column handed to k3_pre (global row count in y-slab mode)
    int row0;           // global row index of local row 0 (0 unless y-slab mode)
    // y-slab mode (one run split over ranks): mode 1 = sweep + write the rank-level carry
    // aggregates (FF, RR, X, Y)[ncol] and stop; mode 2 = solve with the carries entering the
    // rank from its neighbours given (Ain = forward carry into the first local row, Bin =
    // backward carry into the last local row).  mode 0 = cyclic over the local rows.
    int mode;
    double* aggr;       // [4][ncol]
    double* gpart;      // [member][ngp] per-slab shares of psi~1(0,0); nullptr: k3_gauge -> scal[1]
    int k0_external;    // persistent kernel: k3_pre has solved the k = 0 column (P > 4096), do not redo it
    int ngp;
    const double* Ain;  // [ncol]
    const double* Bin;  // [ncol]
    // y-slab peer mode: mode 1 writes its aggregates straight into every rank's aggr_all[rank]
    // and rank 0 its gauge into every rank's scal (NVLink peer stores); peer_n = 0 otherwise
    double* aggr_peer[8];
    double* scal_peer[8];
    int peer_n, peer_rank;
};

struct Handle;

// kernels / launchers (one per translation unit)
cudaError_t launch_zeta(Handle* h, int timestep);
cudaError_t launch_fft_forward(Handle* h, const double* q_fields, int which_pinv);
cudaError_t launch_fft_inverse(Handle* h, double* psi_fields, int use_gauge);
cudaError_t launch_ysolve(Handle* h, int pinned, int do_poisson_only);
cudaError_t dist_halo_exchange(Handle* h, double* base, int slot);   // qg_dist.cu
cudaError_t dist_allgather(Handle* h, const double* send, double* recv, size_t count);
cudaError_t dist_broadcast(Handle* h, double* buf, size_t count, int root);
cudaError_t dist_allreduce_sum(Handle* h, double* buf, size_t count);
cudaError_t dist_allreduce_max(Handle* h, double* buf, size_t count);
void dist_destroy(Handle* h);
cudaError_t dist_barrier(Handle* h);                       // cross-GPU flag barrier on the handle's stream (peer mode)
int dist_ipc_export(Handle* h, void* out256);
int dist_ipc_import(Handle* h, const void* all);
int dist_blobs_share_device(const void* all, int nranks);  // two exports from one GPU?
int dist_poll_error(Handle* h);                            // after a stream sync: did a flag barrier give up?
int dist_init(Handle* h, int rank, int nranks, const void* id128);
int dist_unique_id(void* out128, std::string* err);
cudaError_t launch_diag(Handle* h);
cudaError_t build_plan(Handle* h);
void free_plan(Handle* h);

struct Handle {
    qg_params prm;
    int device = 0;
    int nm = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Geom g;
    int nfields = 0;                 // 3 * nm * 2
    double* q = nullptr;             // [3 slots][nm][2] fields
    double* psi = nullptr;
    double* f = nullptr;
    int qcur = 0;                    // slot of the newest level of q and f_store
    int pcur = 0;                    // slot of the newest level of psi
    bool have_state = false;
    CUtensorMap tm_q, tm_psi, tm_S;
    CUtensorMap tm_S2, tm_T;         // persistent y-solve: big tile boxes, column table
    int k1_ty = 16;                  // K1 tile height (8, 12, 16 or 24; env QG_K1_TY)
    Plan plan;
    bool plan_ok = false;
    double* S = nullptr;             // spectral scratch [nm][P][2M]
    double* k0sol = nullptr;         // [nm][P]
    double* scal = nullptr;          // [nm][4]
    double* solve_tmp = nullptr;     // qg_solve: four padded fields (two in, two out)
    double* snap_stage = nullptr;    // snapshot staging: level 1 of zeta and of psi (qg_snapshot_begin)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_ev = nullptr;
    bool snap_pending = false;
    double* col0 = nullptr;          // [nm][P] compact Poisson k=0 column (written by K2)
    double* gpart = nullptr;         // [nm][ngp] gauge partial sums (written by K3, summed by K4)
    bool gauge_parts = false;        // the last y-solve left gpart (else scal[1] holds the gauge)
    // ---- y-slab decomposition of one run over several GPUs (qg_dist_init) ----
    int dist_n = 1, dist_rank = 0;
    int Pglob = 0;                   // global row count (= P when not distributed)
    void* nccl = nullptr;            // ncclComm_t
    double* col0_full = nullptr;     // [Pglob] gathered k=0 column
    double* k0sol_full = nullptr;    // [Pglob]
    double* aggr = nullptr;          // [4][ncol] this rank's carry aggregates
    double* aggr_all = nullptr;      // [dist_n][4][ncol]
    double* carry_in = nullptr;      // [2][ncol]  Ain, Bin
    // peer-memory exchange (qg_dist_ipc_import): the per-step halo rows, k=0 column, carry aggregates
    // and gauge are written by the producing kernels straight into the other ranks' memory and
    // ordered by flag barriers (k_xgpu_barrier) instead of NCCL calls
    bool peer_ok = false;
    double* mailbox = nullptr;       // [flags 256][col0_full Pglob][aggr_all n*4*ncol][scal 4]
    size_t mailbox_doubles = 0;
    double* peer_q[8] = {};          // every rank's q / psi / mailbox (own pointers at own rank)
    double* peer_psi[8] = {};
    double* peer_mail[8] = {};
    double* own_scal = nullptr;      // the scal array of qg_create (h->scal moves into the mailbox)
    unsigned long long epoch = 0;    // barrier counter
    bool q_halo_pending = false;     // K1 has pushed q rows to the neighbours and no barrier has run since
    double* diag_part = nullptr;     // partial sums for diagnostics
    double* ext_part = nullptr;      // partial extrema (qg_extrema)
    int diag_blocks = 0;
    int64_t launches = 0;
    int profiling = 0;
    // CUDA graphs of the steady state: three AB3 steps return every rotating slot to where it
    // started, so one captured 3-step cycle per (qcur, pcur) phase replays for the rest of the run
    cudaGraphExec_t gexec[3][3] = {};
    int64_t glaunches[3][3] = {};       // kernel launches inside each graph
    int64_t gkcount[3][3][QG_NKERNELS] = {};
    bool use_graph = true;
    bool warm = false;
    std::vector<cudaEvent_t> evpool;   // start/stop pairs
    std::vector<int> evkernel;         // kernel id of each recorded pair
    int ev_next(int id) {
        const size_t n = evkernel.size();
        if (2 * n + 2 > evpool.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
            evpool.push_back(a);
            evpool.push_back(b);
        }
        evkernel.push_back(id);
        return (int)n;
    }
    double kms[QG_NKERNELS] = {0};
    int64_t kcount[QG_NKERNELS] = {0};
    std::string err;

    double* field(double* base, int slot, int member, int layer) const {
        return base + ((int64_t)(slot * nm + member) * 2 + layer) * g.fstride;
    }
    int zindex(int slot, int member, int layer) const { return (slot * nm + member) * 2 + layer; }
};

// Function attributes (dynamic shared-memory opt-in, cluster size) are per device: the launchers
// cache "already configured" per device index so that one process may hold handles on several GPUs.
constexpr int QG_MAX_DEVICES = 64;
inline int dev_slot(const Handle* h) { return h->device >= 0 && h->device < QG_MAX_DEVICES ? h->device : 0; }

// Wraps a kernel launch.  With profiling on, a start/stop CUDA-event pair is recorded around
// the launch on the handle's stream (no host synchronisation: the events are read back in
// qg_kernel_times), so per-kernel durations can be taken inside a timed region.
struct KernelTimer {
    Handle* h;
    int id;
    int slot = -1;
    KernelTimer(Handle* h_, int id_) : h(h_), id(id_) {
        if (h->profiling) {
            slot = h->ev_next(id);
            if (slot >= 0) cudaEventRecord(h->evpool[2 * slot], h->stream);
        }
    }
    ~KernelTimer() {
        h->launches++;
        h->kcount[id]++;
        if (slot >= 0) cudaEventRecord(h->evpool[2 * slot + 1], h->stream);
    }
};

#ifdef __CUDACC__
// ---- mbarrier / TMA primitives (inline PTX) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- block-wide helpers (blockDim.x == 1024) -----------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum, result broadcast to all threads (every warp reduces the same
// per-warp totals in the same order)
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    return warp_sum(lane < nw ? sh[lane] : 0.0);
}

// exclusive prefix over threads of per-thread totals
__device__ __forceinline__ double block_exclusive_scan(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    double tot = lane < nw ? sh[lane] : 0.0;   // warp totals, scanned by every warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, tot, o);
        if (lane >= o) tot += n;
    }
    const double base = w > 0 ? __shfl_sync(0xffffffffu, tot, w - 1) : 0.0;
    return base + inc - v;
}
#endif

}  // namespace qg

struct qg_handle : public qg::Handle {};
