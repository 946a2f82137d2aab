// y-slab decomposition of one run over several GPUs: NCCL plumbing.
//
// The reference is a single process (no collective of any kind), so everything
// here is new.  One process per GPU; rank g owns P_global / G consecutive rows with every x (the
// x-FFT stays local).  Per time step the ranks exchange
//   - two halo rows of q after K1 and of psi after K4, up and down the periodic ring
//     (ncclSend / ncclRecv inside one group; the biharmonic term reaches +-2 rows),
//   - the k = 0 Poisson column (all-gather of P_local doubles) and the rank-level carry
//     aggregates of the y-solve (all-gather of 4 * 2M doubles) instead of the all-to-all
//     transpose a tridiagonal solver would need,
//   - the gauge constant (broadcast of 4 doubles from rank 0).
// All calls are enqueued on the handle's stream, so the step loop never synchronises the host.
// NCCL is loaded with dlopen at qg_dist_init: the single-GPU path has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qg_internal.cuh"

namespace qg {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static const char* load_nccl() {
    if (g_nccl.lib) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return "libnccl.so.2 not found (dlopen)";
#define QG_SYM(field, name)                                           \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                      \
    if (!g_nccl.field) return "libnccl is missing symbol " name;
    QG_SYM(GetUniqueId, "ncclGetUniqueId")
    QG_SYM(CommInitRank, "ncclCommInitRank")
    QG_SYM(CommDestroy, "ncclCommDestroy")
    QG_SYM(AllGather, "ncclAllGather")
    QG_SYM(Broadcast, "ncclBroadcast")
    QG_SYM(AllReduce, "ncclAllReduce")
    QG_SYM(Send, "ncclSend")
    QG_SYM(Recv, "ncclRecv")
    QG_SYM(GroupStart, "ncclGroupStart")
    QG_SYM(GroupEnd, "ncclGroupEnd")
    QG_SYM(GetErrorString, "ncclGetErrorString")
#undef QG_SYM
    g_nccl.lib = lib;
    return nullptr;
}

static cudaError_t nccl_check(Handle* h, ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return cudaSuccess;
    char b[256];
    snprintf(b, sizeof(b), "NCCL %s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->err = b;
    return cudaErrorUnknown;
}

// Developer builds only (-DQG_DEV): QG_DIST_SKIP=<mask> drops collectives (1 halo ring, 2 all-gather,
// 4 broadcast) to expose their cost in a timing run; results are then wrong by construction, so the
// switch does not exist in the shipped library.
#ifdef QG_DEV
static int skip_mask() {
    static const int m = getenv("QG_DIST_SKIP") ? atoi(getenv("QG_DIST_SKIP")) : 0;
    return m;
}
#else
static constexpr int skip_mask() { return 0; }
#endif

cudaError_t dist_allgather(Handle* h, const double* send, double* recv, size_t count) {
    if (skip_mask() & 2) return cudaSuccess;
    return nccl_check(h, g_nccl.AllGather(send, recv, count, ncclFloat64, (ncclComm_t)h->nccl, h->stream), "AllGather");
}

cudaError_t dist_broadcast(Handle* h, double* buf, size_t count, int root) {
    if (skip_mask() & 4) return cudaSuccess;
    return nccl_check(h, g_nccl.Broadcast(buf, buf, count, ncclFloat64, root, (ncclComm_t)h->nccl, h->stream), "Broadcast");
}

cudaError_t dist_allreduce_sum(Handle* h, double* buf, size_t count) {
    return nccl_check(h, g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)h->nccl, h->stream), "AllReduce");
}

cudaError_t dist_allreduce_max(Handle* h, double* buf, size_t count) {
    return nccl_check(h, g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclMax, (ncclComm_t)h->nccl, h->stream), "AllReduce");
}

// Fill the two ghost rows above and below the local rows of every field of one slot with the
// neighbours' boundary rows (periodic ring).  Whole padded rows travel, so the x ghosts and the
// corners arrive with them.
cudaError_t dist_halo_exchange(Handle* h, double* base, int slot) {
    if (skip_mask() & 1) return cudaSuccess;
    const Geom& g = h->g;
    const int up = (h->dist_rank + 1) % h->dist_n, down = (h->dist_rank + h->dist_n - 1) % h->dist_n;
    const size_t n = (size_t)GHOST * g.pitch;
    ncclComm_t comm = (ncclComm_t)h->nccl;
    ncclResult_t r = g_nccl.GroupStart();
    for (int m = 0; m < h->nm && r == ncclSuccess; ++m)
        for (int l = 0; l < 2 && r == ncclSuccess; ++l) {
            double* f = h->field(base, slot, m, l);
            double* first = f + (int64_t)YPAD * g.pitch;                     // local rows 0, 1
            double* last = f + (int64_t)(YPAD + g.P - GHOST) * g.pitch;      // local rows P-2, P-1
            double* ghost_lo = f;                                            // rows -2, -1
            double* ghost_hi = f + (int64_t)(YPAD + g.P) * g.pitch;          // rows P, P+1
            r = g_nccl.Send(first, n, ncclFloat64, down, comm, h->stream);
            if (r == ncclSuccess) r = g_nccl.Send(last, n, ncclFloat64, up, comm, h->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(ghost_hi, n, ncclFloat64, up, comm, h->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(ghost_lo, n, ncclFloat64, down, comm, h->stream);
        }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_check(h, r, "Send/Recv");
    return nccl_check(h, r2, "GroupEnd");
}

// ---- peer-memory exchange ----------------------------------------------------------------------
// One process per GPU, so the other ranks' arrays are reached through CUDA IPC handles; NVSwitch
// gives every pair full NVLink bandwidth.  Producers store straight into the consumers' memory
// (K1 / K4: two halo rows into each ring neighbour's ghost rows; K2: its part of the k=0 column
// into every rank's gathered column; y-solve mode 1: its carry aggregates into every rank's
// aggr_all; rank 0: the gauge) and the ranks order these stores with a flag barrier:
// rank r writes the barrier's epoch into slot r of every rank's flag array (st.release.sys after a
// system-scope fence) and spins until all slots of its own array carry that epoch.  Kernels on
// a stream run in order, so a rank signals only after its producing kernel has completed, and
// its consumer starts only after everyone has signalled.
struct PeerFlags {
    unsigned long long* f[8];
};

// Layout of the first 256 doubles of a mailbox, as 64-bit words: word r * 16 = the epoch rank r has
// reached (one 128-byte line per writer), word MB_ABORT = sticky abort word (0 = healthy).
constexpr int MB_ABORT = 200;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The wait is bounded: a rank that does not see every peer's epoch within `timeout_ns` (a dead peer, or
// ranks whose call sequences diverged) writes the abort word of EVERY rank and leaves; a set abort word
// ends this and every later barrier at once, so no GPU is ever left in an unkillable spin.  The host
// reads the word after its next stream synchronisation (dist_poll_error) and reports QG_ERR_CUDA.
// One rank per GPU is a precondition (qg_dist_ipc_import refuses anything else): kernels that wait on
// one another must be resident at the same time.
__global__ void k_xgpu_barrier(PeerFlags p, int rank, int n, unsigned long long epoch, unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long* mine = p.f[rank];
    if (ld_acquire_sys(mine + MB_ABORT) != 0ull) return;
    __threadfence_system();
    st_release_sys(p.f[i] + rank * 16, epoch);
    const unsigned long long t0 = global_ns();
    unsigned int backoff = 32;
    while (ld_acquire_sys(mine + i * 16) < epoch) {
        if (ld_acquire_sys(mine + MB_ABORT) != 0ull) return;
        if (global_ns() - t0 > timeout_ns) {
            // word = epoch that failed (never 0: epochs start at 1), high bits = the rank that never arrived
            const unsigned long long w = epoch | ((unsigned long long)(i + 1) << 56);
            for (int r = 0; r < n; ++r) st_release_sys(p.f[r] + MB_ABORT, w);
            return;
        }
        __nanosleep(backoff);
        if (backoff < 1024) backoff <<= 1;
    }
}

static unsigned long long barrier_timeout_ns() {
    static const unsigned long long t = [] {
        const char* v = getenv("QG_BARRIER_TIMEOUT_S");
        double s = v ? atof(v) : 120.0;
        if (!(s > 0.0)) s = 120.0;
        return (unsigned long long)(s * 1e9);
    }();
    return t;
}

cudaError_t dist_barrier(Handle* h) {
    PeerFlags p{};
    for (int r = 0; r < h->dist_n; ++r) p.f[r] = reinterpret_cast<unsigned long long*>(h->peer_mail[r]);
    ++h->epoch;
    h->launches++;
    h->q_halo_pending = false;
    k_xgpu_barrier<<<1, 32, 0, h->stream>>>(p, h->dist_rank, h->dist_n, h->epoch, barrier_timeout_ns());
    return cudaGetLastError();
}

// After a stream synchronisation: has a cross-GPU barrier of this run given up?  (peer mode only)
int dist_poll_error(Handle* h) {
    if (!h->peer_ok || !h->mailbox) return QG_OK;
    unsigned long long w = 0;
    cudaError_t e = cudaMemcpy(&w, reinterpret_cast<unsigned long long*>(h->mailbox) + MB_ABORT, sizeof(w), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { h->err = std::string("reading the barrier status: ") + cudaGetErrorString(e); return QG_ERR_CUDA; }
    if (w == 0) return QG_OK;
    char b[256];
    snprintf(b, sizeof(b), "cross-GPU barrier %llu timed out waiting for rank %d (a peer died, or the ranks did not make the "
             "same sequence of collective calls); the state of this run is invalid", w & 0x00ffffffffffffffull, (int)(w >> 56) - 1);
    h->err = b;
    return QG_ERR_CUDA;
}

// ---- IPC blobs -----------------------------------------------------------------------------------
// One export = 256 bytes: three CUDA IPC handles (q, psi, mailbox; 3 x 64 bytes), the 16-byte UUID of the
// exporting GPU, 48 bytes reserved (zero).
constexpr int IPC_BLOB = 256;

// 1 if two of the `nranks` exports in `all` come from the same GPU (equal UUIDs), else 0.  Two ranks on
// one GPU must never use the flag barrier: their barrier kernels are not guaranteed to be resident at
// the same time (B200_PROFILING.md: Xid 109).
int dist_blobs_share_device(const void* all, int nranks) {
    const unsigned char* b = static_cast<const unsigned char*>(all);
    for (int i = 0; i < nranks; ++i)
        for (int j = i + 1; j < nranks; ++j)
            if (memcmp(b + (size_t)i * IPC_BLOB + 192, b + (size_t)j * IPC_BLOB + 192, 16) == 0) return 1;
    return 0;
}

int dist_ipc_export(Handle* h, void* out256) {
    if (h->dist_n < 2 || !h->mailbox) { h->err = "qg_dist_ipc_export: call qg_dist_init first"; return QG_ERR_STATE; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hd[3];
    void* ptr[3] = {h->q, h->psi, h->mailbox};
    for (int i = 0; i < 3; ++i) {
        cudaError_t e = cudaIpcGetMemHandle(&hd[i], ptr[i]);
        if (e != cudaSuccess) { h->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return QG_ERR_CUDA; }
    }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, h->device);
    if (e != cudaSuccess) { h->err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e); return QG_ERR_CUDA; }
    static_assert(sizeof(prop.uuid) == 16, "cudaUUID_t is 16 bytes");
    unsigned char* out = static_cast<unsigned char*>(out256);
    memset(out, 0, IPC_BLOB);
    memcpy(out, hd, sizeof(hd));
    memcpy(out + 192, &prop.uuid, 16);
    return QG_OK;
}

int dist_ipc_import(Handle* h, const void* all) {
    if (h->dist_n < 2 || !h->mailbox) { h->err = "qg_dist_ipc_import: call qg_dist_init first"; return QG_ERR_STATE; }
    if (h->peer_ok) return QG_OK;
    if (dist_blobs_share_device(all, h->dist_n)) {
        h->err = "qg_dist_ipc_import: two ranks of this run share a GPU; the peer-memory flag barrier needs one rank per "
                 "GPU (the run stays on the NCCL exchange path)";
        return QG_ERR_INVALID;
    }
    const unsigned char* blobs = static_cast<const unsigned char*>(all);
    auto close_all = [&]() {   // undo a partial import
        for (int r = 0; r < h->dist_n; ++r) {
            if (r == h->dist_rank) continue;
            if (h->peer_q[r]) cudaIpcCloseMemHandle(h->peer_q[r]);
            if (h->peer_psi[r]) cudaIpcCloseMemHandle(h->peer_psi[r]);
            if (h->peer_mail[r]) cudaIpcCloseMemHandle(h->peer_mail[r]);
        }
        for (int r = 0; r < 8; ++r) h->peer_q[r] = h->peer_psi[r] = h->peer_mail[r] = nullptr;
    };
    for (int r = 0; r < h->dist_n; ++r) {
        if (r == h->dist_rank) {
            h->peer_q[r] = h->q; h->peer_psi[r] = h->psi; h->peer_mail[r] = h->mailbox;
            continue;
        }
        cudaIpcMemHandle_t hd[3];
        memcpy(hd, blobs + (size_t)r * IPC_BLOB, sizeof(hd));
        double** dst[3] = {&h->peer_q[r], &h->peer_psi[r], &h->peer_mail[r]};
        for (int i = 0; i < 3; ++i) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, hd[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                h->err = std::string("cudaIpcOpenMemHandle (is NVLink peer access available?): ") + cudaGetErrorString(e);
                cudaGetLastError();
                close_all();
                return QG_ERR_CUDA;
            }
            *dst[i] = static_cast<double*>(p);
        }
    }
    h->peer_ok = true;
    if (getenv("QG_VERBOSE"))
        fprintf(stderr, "qgb200: rank %d/%d exchanges over NVLink peer memory (CUDA IPC), NCCL off the step path\n",
                h->dist_rank, h->dist_n);
    return QG_OK;
}

void dist_destroy(Handle* h) {
    if (h->peer_ok) {
        cudaStreamSynchronize(h->stream);
        for (int r = 0; r < h->dist_n; ++r)
            if (r != h->dist_rank) {
                cudaIpcCloseMemHandle(h->peer_q[r]); cudaIpcCloseMemHandle(h->peer_psi[r]); cudaIpcCloseMemHandle(h->peer_mail[r]);
            }
        h->peer_ok = false;
    }
    if (h->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->nccl);
    h->nccl = nullptr;
    if (h->own_scal) { h->scal = h->own_scal; h->own_scal = nullptr; }
    cudaFree(h->mailbox); cudaFree(h->k0sol_full); cudaFree(h->aggr); cudaFree(h->carry_in);
    h->mailbox = h->col0_full = h->k0sol_full = h->aggr = h->aggr_all = h->carry_in = nullptr;
}

int dist_unique_id(void* out128, std::string* err) {
    const char* e = load_nccl();
    if (e) { *err = e; return QG_ERR_CUDA; }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return QG_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return QG_OK;
}

int dist_init(Handle* h, int rank, int nranks, const void* id128) {
    const char* e = load_nccl();
    if (e) { h->err = e; return QG_ERR_CUDA; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return QG_ERR_CUDA; }
    h->nccl = comm;
    h->dist_n = nranks;
    h->dist_rank = rank;
    h->Pglob = h->g.P * nranks;
    const size_t ncol = h->plan.ncol;
    // everything another rank may write lives in one allocation (one IPC handle): barrier flags,
    // the gathered k=0 column, every rank's carry aggregates, the scalars (pin total, gauge)
    const size_t pg = ((size_t)h->Pglob + 15) / 16 * 16;
    h->mailbox_doubles = 256 + pg + (size_t)nranks * 4 * ncol + 16;
    cudaError_t ce = cudaMalloc((void**)&h->mailbox, h->mailbox_doubles * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemset(h->mailbox, 0, h->mailbox_doubles * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->k0sol_full, (size_t)h->Pglob * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->aggr, 4 * ncol * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->carry_in, 2 * ncol * sizeof(double));
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { h->err = std::string("qg_dist_init: ") + cudaGetErrorString(ce); return QG_ERR_NOMEM; }
    h->col0_full = h->mailbox + 256;
    h->aggr_all = h->col0_full + pg;
    h->own_scal = h->scal;
    h->scal = h->aggr_all + (size_t)nranks * 4 * ncol;
    h->epoch = 0;
    return QG_OK;
}

}  // namespace qg
