/* qgb200.h — C ABI of the B200-native Phillips two-layer QG time stepper.
 *
 * Drop-in boundary for the reference's per-timestep hot path.  The reference
 * (JSLeadbetter/julia-ocean-modelling) has no FFI of its own: its boundary is the Julia
 * function API of src/model.jl and src/schemes/laplacian.jl.  Every entry point below
 * names the reference function (file:line, relative to the reference repo root) it
 * replaces.  The Julia host shim (julia-ocean-modelling_b200/julia/src/model.jl) keeps the
 * reference's names and signatures and forwards to these symbols with `ccall`; the Python
 * ctypes twin (julia-ocean-modelling_b200/python/qgb200) binds the same symbols and is the
 * one exercised by tests/ and bench.py (no Julia toolchain in the image).
 *
 * Conventions
 *   - Plain C: pointers and sizes only, no exceptions cross the boundary.  Every function
 *     returns 0 on success or a negative qg_status; qg_last_error() gives the message.
 *   - Host state arrays use exactly the reference layout: Float64, column-major
 *     (M+2, P+2, 2 layers, 3 time levels), one ghost ring, level 1 (offset 0) newest
 *     (src/model.jl:53-59,102-106).  Offset of [i,j,z,t] (0-based) =
 *     i + (M+2)*(j + (P+2)*(z + 2*t)).  With nmembers > 1 the members are concatenated.
 *     Host arrays are borrowed for the duration of a call only.
 *   - Device state is authoritative between calls; a handle is bound to one GPU and is
 *     not thread-safe.  Separate handles are independent (ensemble members, ranks).
 *   - All derived constants (beta_1, beta_2, S_eig, P, P_inv, ...) are computed by the host
 *     shim with the reference's own formulas (src/model.jl:83-121) and passed as data, so
 *     the reference's behaviour — including evolve_psi!'s use of P_matrix(H_1, H_1) at
 *     src/model.jl:173 — is reproduced as it runs.
 */
#ifndef QGB200_H
#define QGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_ABI_VERSION 2

typedef enum qg_status {
    QG_OK = 0,
    QG_ERR_INVALID = -1,   /* bad argument (null pointer, M or P < 3, unsupported size)   */
    QG_ERR_CUDA = -2,      /* CUDA runtime / driver failure (message has the CUDA string) */
    QG_ERR_NOMEM = -3,     /* device or host allocation failed                            */
    QG_ERR_NODEVICE = -4,  /* no usable sm_100 GPU: there is no CPU fallback               */
    QG_ERR_STATE = -5      /* call order violated (e.g. stepping before qg_upload_state)   */
} qg_status;

/* Parameter block: BaroclinicModel (src/model.jl:12-34) reduced to what the hot path reads,
 * plus the derived constants of src/model.jl:83-121 evaluated by the host. */
typedef struct qg_params {
    int32_t M;        /* nodes in x (contiguous index), src/model.jl:23                     */
    int32_t P;        /* nodes in y, src/model.jl:24                                        */
    double dx;        /* grid spacing, src/model.jl:25                                      */
    double dt;        /* time step, src/model.jl:20                                         */
    double visc;      /* Laplacian viscosity nu, src/model.jl:26                            */
    double r;         /* bottom friction, src/model.jl:27                                   */
    double U;         /* layer-1 mean flow, src/model.jl:22                                 */
    double beta1;     /* beta_1(model), src/model.jl:117                                    */
    double beta2;     /* beta_2(model), src/model.jl:118                                    */
    double alpha;     /* S_eig(model) = -1/R_d^2, src/model.jl:121                          */
    double Pinv[4];   /* P_inv_matrix(model), row-major 2x2, src/model.jl:90-99             */
    double Pfwd[4];   /* the P actually used by evolve_psi!, row-major, src/model.jl:173    */
    double H1, H2;    /* layer depths (diagnostics only), src/model.jl:13-14                */
    double S1;        /* S1_plus(model) (diagnostics only), src/model.jl:113                */
} qg_params;

typedef struct qg_handle qg_handle;

/* Library identification: returns QG_ABI_VERSION. */
int qg_abi_version(void);

/* Message of the most recent failure on `h` (or of the most recent failed qg_create when
 * `h` is NULL).  Never NULL; valid until the next call on the same handle/thread. */
const char* qg_last_error(const qg_handle* h);

/* Build the solver plan and allocate device state for `nmembers` independent runs that
 * share `params`.  Replaces the two factorisations get_poisson_cholesky
 * (src/schemes/laplacian.jl:66-75) and get_helmholtz_cholesky (:60-64) and the f_store
 * allocation of src/run_model_no_output.jl:5-8: the "factor" is a spectral plan
 * (x-FFT twiddles + per-wavenumber y-recurrence coefficients), O(M) numbers.
 * `device` is the CUDA ordinal.  `stream` is a cudaStream_t to launch on (NULL: the
 * handle creates its own non-blocking stream). */
int qg_create(const qg_params* params, int device, int nmembers, void* stream, qg_handle** out);

/* Free everything owned by the handle (Julia side: finalizer). */
int qg_destroy(qg_handle* h);

/* Host -> device: zeta, psi, f_store as built by initialise_model (src/model.jl:37-62) and
 * src/run_model_no_output.jl:8.  Ghost cells are regenerated from the interior
 * (update_doubly_periodic_bc!, src/schemes/boundary_conditions.jl:2-13).  Any pointer may
 * be NULL to leave that array untouched on the device (f_store NULL at first upload =
 * zeros). */
int qg_upload_state(qg_handle* h, const double* zeta, const double* psi, const double* f_store);

/* Same, for the state a run starts from: only time level 1 of zeta and psi is read from the
 * host arrays (same (M+2, P+2, 2, 3) layout); levels 2-3 and f_store are zeroed on the device,
 * which is exactly what initialise_model (src/model.jl:53-59) and
 * src/run_model_no_output.jl:8 hand to the loop.  Moves 4.5x fewer bytes over PCIe. */
int qg_upload_initial_state(qg_handle* h, const double* zeta, const double* psi);

/* initialise_model (src/model.jl:37-62) on the device, no host arrays involved: psi_l =
 * amplitude * uniform[0,1) at every node (amplitude = initial_kick * U * Ly, :41-42), periodic
 * ghosts (:44-45), q_1 = lap(psi_1) + S1 (psi_2 - psi_1), q_2 = lap(psi_2) + S2 (psi_1 - psi_2)
 * (:47-48) in level 1; history levels and f_store zero.  The reference draws from Julia's
 * unseeded global RNG; here the stream is Philox4x32-10 with counter (j*M + i, member*2 + layer)
 * and key `seed`, so a run is reproducible from (seed, parameters) alone.  S1, S2 are
 * S1_plus / S2_minus of src/model.jl:113-115, evaluated by the host shim.  On a y-slab handle
 * (qg_dist_init) j is the GLOBAL row index: every rank draws exactly its rows of the single-GPU
 * field, its neighbours' rows into the ghost rows included, without any exchange (collective
 * only in that all ranks must call it, and in peer mode it ends with a flag barrier). */
int qg_init_state(qg_handle* h, uint64_t seed, double amplitude, double S1, double S2);

/* Device -> host, all three time levels, ghosts included, reference layout.  Any pointer
 * may be NULL. */
int qg_download_state(qg_handle* h, double* zeta, double* psi, double* f_store);

/* Snapshot of the newest time level only: zeta[:,:,:,1] and psi[:,:,:,1], what run_model
 * writes at timestep 0 and every sample_timestep (src/run_model.jl:70-73, 86-90).  Host layout
 * (M+2, P+2, 2) per member, ghosts included; either pointer may be NULL.  qg_snapshot_begin
 * orders the snapshot after the work already queued on the handle and returns at once: the
 * device -> host copy runs on its own stream, so steps queued afterwards overlap it.  The host
 * buffers (pinned memory for a truly asynchronous copy) must stay valid, and must not be read,
 * until qg_snapshot_end returns.  One snapshot in flight per handle: a second begin waits for
 * the first. */
int qg_snapshot_begin(qg_handle* h, double* zeta_level1, double* psi_level1);
int qg_snapshot_end(qg_handle* h);

/* evolve_zeta!(model, zeta, psi, timestep, f_store), src/model.jl:155-170: Arakawa
 * Jacobian + biharmonic viscosity + beta / mean-flow / friction terms, Euler for
 * timestep 1 and 2, AB3 afterwards; pushes the new RHS into f_store and the new PV into
 * zeta (history shift of src/model.jl:102-106).  `timestep` is 1-based. */
int qg_evolve_zeta(qg_handle* h, int timestep);

/* evolve_psi!(model, zeta, psi, poisson_cholesky, helmholtz_cholesky), src/model.jl:172-199:
 * modal projection, pinned Poisson solve, modified-Helmholtz solve, back-projection,
 * history shift of psi. */
int qg_evolve_psi(qg_handle* h);

/* Loop body of src/run_model_no_output.jl:10-13 for timesteps first_timestep ..
 * first_timestep+nsteps-1, enqueued back to back; returns after the work has been
 * queued (qg_sync or any download waits for it). */
int qg_step(qg_handle* h, int first_timestep, int nsteps);

/* Block until all queued work of the handle has finished; reports asynchronous errors. */
int qg_sync(qg_handle* h);

/* Domain-integrated energy and enstrophy of the newest level, one value per member
 * (definition: DESIGN.md "Diagnostics"; the reference has none). */
int qg_diagnostics(qg_handle* h, double* energy, double* enstrophy);

/* Extrema of the newest level over the interior, per member: out[8 * member + ..] = {max q_1, min q_1,
 * max q_2, min q_2, max psi_1, min psi_1, max psi_2, min psi_2}.  Device-side counterpart of the
 * reference's update_max / update_min (src/run_model.jl:41-53), which scan a host matrix; the host
 * shims keep the running maximum / minimum.  y-slab mode: extrema of the whole domain (collective). */
int qg_extrema(qg_handle* h, double* out);

/* Single-use solves of src/schemes/laplacian.jl:78-111 on the plan of this handle:
 * sp_solve_poisson (pinned = 1, alpha ignored: uses the Poisson plan) or
 * sp_solve_modified_helmholtz with the handle's alpha (pinned = 0).
 * `f` and `u` are host (M+2, P+2) fields with ghosts; member 0's scratch is used. */
int qg_solve(qg_handle* h, int pinned, const double* f, double* u);

/* Per-kernel device times of the most recent qg_step with profiling enabled
 * (qg_set_profiling(h, 1)): for each kernel id < QG_NKERNELS the summed CUDA-event
 * milliseconds and the launch count, measured on the handle's stream. */
#define QG_NKERNELS 8
enum { QG_K_ZETA = 0, QG_K_FFT_FWD = 1, QG_K_YPRE = 2, QG_K_YSOLVE = 3, QG_K_GAUGE = 4,
       QG_K_FFT_INV = 5, QG_K_DIAG = 6, QG_K_PACK = 7 };
int qg_set_profiling(qg_handle* h, int enabled);
int qg_kernel_times(qg_handle* h, double* ms, int64_t* launches);
const char* qg_kernel_name(int kernel_id);

/* Number of kernels launched by this handle since creation (all streams). */
int64_t qg_launch_count(const qg_handle* h);

/* y-slab domain decomposition of ONE run over `nranks` GPUs of a node, one process per GPU
 * (new: the reference is a single process).  Create the handle with params.P = the rank's
 * LOCAL row count (global P / nranks, a multiple of 32, <= 8192); rank r owns global rows
 * [r * P, (r+1) * P).  After qg_dist_init every call works on the local slab: host arrays are
 * (M+2, P_local+2, 2, 3) with the neighbours' rows in the ghost rows on download.  Per step the
 * ranks exchange two halo rows of q and of psi with their ring neighbours, the k = 0 Poisson
 * column, the carry aggregates of the y-solve (4 x 2M doubles per rank, in place of the all-to-all
 * transpose a tridiagonal solver would need) and the gauge constant.  After qg_dist_init alone these
 * travel as NCCL calls on the handle's stream (send/recv ring, all-gathers, one broadcast); after
 * qg_dist_ipc_import - the default of the host shims - the producing kernels store them straight
 * into the other ranks' memory over NVLink and no collective call remains on the step path.
 * qg_nccl_unique_id fills a 128-byte NCCL id on one rank; the host distributes it (e.g.
 * torch.distributed / MPI broadcast) and every rank passes the same bytes.  Collective: all ranks
 * must call qg_dist_init, qg_upload_state, qg_init_state, qg_step, qg_download_state and
 * qg_diagnostics together. */
int qg_nccl_unique_id(void* out128);
int qg_dist_init(qg_handle* h, int rank, int nranks, const void* unique_id128);

/* Peer-memory exchange (optional, after qg_dist_init).  qg_dist_ipc_export fills 256 bytes: three
 * CUDA IPC handles (q, psi, the rank's mailbox; 3 x 64 bytes), the 16-byte UUID of the rank's GPU,
 * 48 reserved bytes.  The host gathers the exports of all ranks in rank order (nranks x 256 bytes)
 * and hands them to qg_dist_ipc_import on every rank.  From then on K1 / K4 store their two edge
 * rows straight into the ring neighbours' ghost rows, K2 its part of the k = 0 column into every
 * rank's gathered column, the y-solve its carry aggregates into every rank's table, rank 0 the
 * gauge, and a flag barrier (one tiny kernel, four per step) orders those stores.  Needs peer access
 * between all GPUs of the run (NVSwitch) and ONE RANK PER GPU: kernels that wait on one another must
 * be resident together, so qg_dist_ipc_import returns QG_ERR_INVALID when two exports carry the same
 * GPU UUID (qg_dist_ipc_blobs_share_device is that test, callable without a GPU) and the run stays
 * on the NCCL path.  The barrier's wait is bounded (QG_BARRIER_TIMEOUT_S, default 120 s): if a peer
 * never arrives every rank leaves the barrier, and the next qg_sync / qg_download_state /
 * qg_diagnostics returns QG_ERR_CUDA with the stalled rank in qg_last_error. */
int qg_dist_ipc_export(qg_handle* h, void* out256);
int qg_dist_ipc_import(qg_handle* h, const void* all_ranks);
int qg_dist_ipc_blobs_share_device(const void* all_ranks, int nranks);

/* Host-side view of the spectral plan, computed without a GPU (inspection and tests; the handle builds its
 * plan with the same routines).  For every real column c < 2M of the packed spectral layout: r[c], the decay
 * per row of the y-recurrences (0 for the singular Poisson k = 0 column), and kappa[c] = -r dx^2 / M, the scale
 * that turns the second sweep into the solution.  With worklist_len != NULL also the (32-column tile, 32-row
 * segment) pairs the y-slab edge correction touches for a slab of `rows_local` rows (a multiple of 32):
 * worklist[2i], worklist[2i+1], at most worklist_capacity pairs are written, *worklist_len is the full count.
 * Any output pointer may be NULL. */
int qg_plan_probe(const qg_params* params, int rows_local, double* r, double* kappa, int32_t* worklist,
                  int worklist_capacity, int* worklist_len);

/* Raw device pointers for zero-copy interop (multi-GPU plumbing, torch tensors):
 * which = 0: q, 1: psi, 2: f_store, 3: spectral scratch.  Returns the base pointer, the
 * row pitch in doubles, the left padding (x offset of interior column 0), the ghost-row
 * count above row 0, and the per-field stride in doubles. */
int qg_device_layout(qg_handle* h, int which, void** base, int64_t* pitch, int64_t* xpad,
                     int64_t* ypad, int64_t* field_stride);

#ifdef __cplusplus
}
#endif
#endif /* QGB200_H */
