// K1 — fused potential-vorticity step.
//
// One launch replaces evolve_zeta! (reference src/model.jl:155-170) for both layers and
// all ensemble members: biharmonic viscosity nu*Lap(Lap(psi)) (src/schemes/laplacian.jl:15-27
// applied twice, src/model.jl:140,148), Arakawa Jacobian J(q, psi)
// (src/schemes/arakawa.jl:7-62), beta / mean-flow / bottom-friction terms
// (src/model.jl:142-143,150-151), Euler or AB3 update (src/model.jl:123-136) and the
// history pushes (src/model.jl:102-106, done by slot rotation on the host).
//
// Data movement: each CTA owns a TX x TY tile; psi and q tiles with a 2-cell halo are
// brought into shared memory by two TMA tensor copies (cp.async.bulk.tensor.3d) that
// complete on one mbarrier; ghost cells in the padded device layout make the periodic
// wrap a plain halo read.  f history is read and q+/f written straight from/to global
// memory (pointwise, fully coalesced).  Algorithmic traffic: 8 reads + 4 writes of one
// double per cell per step = 96 B (64 B on the two Euler steps).
#include "qg_internal.cuh"

namespace qg {

constexpr int K1_THREADS = 256;
constexpr int K1_ROWS_PER_THREAD = K1_TY / (K1_THREADS / K1_TX);   // 8
constexpr int K1_TILE_BYTES = ((K1_BX * K1_BY * 8 + 127) / 128) * 128;
constexpr int K1_LAP_BYTES = (((K1_TX + 2) * (K1_TY + 2) * 8 + 127) / 128) * 128;
constexpr int K1_SMEM_BYTES = 2 * K1_TILE_BYTES + K1_LAP_BYTES + 128;

__global__ void __launch_bounds__(K1_THREADS)
k1_zeta_step(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_psi,
             const ZetaArgs a) {
    extern __shared__ __align__(128) unsigned char k1_smem[];
    typedef double TileRow[K1_BX];
    typedef double LapRow[K1_TX + 2];
    TileRow* s_psi = reinterpret_cast<TileRow*>(k1_smem);
    TileRow* s_q = reinterpret_cast<TileRow*>(k1_smem + K1_TILE_BYTES);
    LapRow* s_lap = reinterpret_cast<LapRow*>(k1_smem + 2 * K1_TILE_BYTES);
    uint64_t* barp = reinterpret_cast<uint64_t*>(k1_smem + 2 * K1_TILE_BYTES + K1_LAP_BYTES);
#define bar (*barp)

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * K1_TX;
    const int y0 = blockIdx.y * K1_TY;
    const int fz = blockIdx.z;          // member * 2 + layer
    const int layer = fz & 1;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, 2u * K1_BX * K1_BY * sizeof(double));
        tma_load_3d(&s_psi[0][0], &tm_psi, XPAD + x0 - GHOST, YPAD + y0 - GHOST, a.zpsi + fz, &bar);
        tma_load_3d(&s_q[0][0], &tm_q, XPAD + x0 - GHOST, YPAD + y0 - GHOST, a.zq + fz, &bar);
    }
    mbar_wait(&bar, 0);

    // Laplacian of psi on the tile plus a one-cell rim (same summation order as
    // src/schemes/laplacian.jl:21).
    for (int e = tid; e < (K1_TY + 2) * (K1_TX + 2); e += K1_THREADS) {
        const int ly = e / (K1_TX + 2);
        const int lx = e - ly * (K1_TX + 2);
        const int sy = ly + 1, sx = lx + 1;
        s_lap[ly][lx] = (s_psi[sy][sx - 1] + s_psi[sy][sx + 1] - 4.0 * s_psi[sy][sx] +
                         s_psi[sy - 1][sx] + s_psi[sy + 1][sx]) * a.idx2;
    }
    __syncthreads();

    const int lx = tid & (K1_TX - 1);
    const int ly0 = (tid / K1_TX) * K1_ROWS_PER_THREAD;
    const int x = x0 + lx;
    if (x >= a.g.M) return;
    const int64_t foff = (int64_t)fz * a.g.fstride;
    const double* __restrict__ f1 = a.f1 + foff;
    const double* __restrict__ f2 = a.f2 + foff;
    double* __restrict__ fn = a.fn + foff;
    double* __restrict__ qn = a.qn + foff;
    const double beta = a.beta[layer];
    const int sx = lx + GHOST;
    const int px = lx + 1;

#pragma unroll
    for (int i = 0; i < K1_ROWS_PER_THREAD; ++i) {
        const int ly = ly0 + i;
        const int y = y0 + ly;
        if (y >= a.g.P) break;
        const int sy = ly + GHOST;
        const int py = ly + 1;
        // psi neighbourhood (x is the reference's first index i, y its second index j)
        const double p_w = s_psi[sy][sx - 1], p_e = s_psi[sy][sx + 1];
        const double p_s = s_psi[sy - 1][sx], p_n = s_psi[sy + 1][sx];
        const double p_sw = s_psi[sy - 1][sx - 1], p_se = s_psi[sy - 1][sx + 1];
        const double p_nw = s_psi[sy + 1][sx - 1], p_ne = s_psi[sy + 1][sx + 1];
        const double q_c = s_q[sy][sx];
        const double q_w = s_q[sy][sx - 1], q_e = s_q[sy][sx + 1];
        const double q_s = s_q[sy - 1][sx], q_n = s_q[sy + 1][sx];
        const double q_sw = s_q[sy - 1][sx - 1], q_se = s_q[sy - 1][sx + 1];
        const double q_nw = s_q[sy + 1][sx - 1], q_ne = s_q[sy + 1][sx + 1];

        // src/schemes/arakawa.jl:13-15, 28-33, 46-51, 59
        const double jpp = (q_e - q_w) * (p_n - p_s) - (q_n - q_s) * (p_e - p_w);
        const double jpt = q_e * (p_ne - p_se) - q_w * (p_nw - p_sw) - q_n * (p_ne - p_nw) +
                           q_s * (p_se - p_sw);
        const double jtp = q_ne * (p_n - p_e) - q_sw * (p_w - p_s) - q_nw * (p_n - p_w) +
                           q_se * (p_e - p_s);
        const double jac = ((jpp + jpt) + jtp) * a.i12dx2;

        const double l_c = s_lap[py][px];
        const double lap2 = (s_lap[py][px - 1] + s_lap[py][px + 1] - 4.0 * l_c + s_lap[py - 1][px] +
                             s_lap[py + 1][px]) * a.idx2;
        const double dpsi = a.hdx * (p_e - p_w);
        // src/model.jl:144 / :152, evaluated left to right
        double rhs = (a.visc * lap2 - jac) - beta * dpsi;
        if (layer == 0)
            rhs -= a.U * (a.hdx * (q_e - q_w));
        else
            rhs -= a.r * l_c;

        const int64_t o = a.g.at(x, y);
        double qnew;
        if (a.euler) {
            qnew = q_c + a.dt * rhs;   // src/model.jl:126
        } else {
            const double h1 = __ldg(f1 + o), h2 = __ldg(f2 + o);
            qnew = q_c + a.dt * ((a.c1 * rhs - a.c2 * h1) + a.c3 * h2);   // src/model.jl:134-135
        }
        fn[o] = rhs;
        qn[o] = qnew;
        // periodic images (update_doubly_periodic_bc!, src/schemes/boundary_conditions.jl:2-13,
        // widened to two ghost cells)
        const bool gl = x < GHOST, gr = x >= a.g.M - GHOST;
        const bool gb = y < GHOST, gt = y >= a.g.P - GHOST;
        if (gl) qn[o + a.g.M] = qnew;
        if (gr) qn[o - a.g.M] = qnew;
        if (gb | gt) {
            const int64_t dyo = (int64_t)a.g.P * a.g.pitch;
            if (gb) {
                qn[o + dyo] = qnew;
                if (gl) qn[o + dyo + a.g.M] = qnew;
                if (gr) qn[o + dyo - a.g.M] = qnew;
            }
            if (gt) {
                qn[o - dyo] = qnew;
                if (gl) qn[o - dyo + a.g.M] = qnew;
                if (gr) qn[o - dyo - a.g.M] = qnew;
            }
        }
    }
}

cudaError_t launch_zeta(Handle* h, int timestep) {
    ZetaArgs a;
    a.g = h->g;
    const int cur = h->qcur, nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
    a.f1 = h->field(h->f, cur, 0, 0);
    a.f2 = h->field(h->f, prv, 0, 0);
    a.fn = h->field(h->f, nxt, 0, 0);
    a.qn = h->field(h->q, nxt, 0, 0);
    a.zq = h->zindex(cur, 0, 0);
    a.zpsi = h->zindex(h->pcur, 0, 0);
    a.euler = (timestep == 1 || timestep == 2) ? 1 : 0;   // src/model.jl:161
    const double inv = 1.0 / h->prm.dx;
    a.idx2 = inv * inv;
    a.hdx = 0.5 * inv;
    a.i12dx2 = 1.0 / (3 * 4 * (h->prm.dx * h->prm.dx));
    a.visc = h->prm.visc;
    a.dt = h->prm.dt;
    a.beta[0] = h->prm.beta1;
    a.beta[1] = h->prm.beta2;
    a.U = h->prm.U;
    a.r = h->prm.r;
    a.c1 = 23.0 / 12.0;
    a.c2 = 16.0 / 12.0;
    a.c3 = 5.0 / 12.0;
    dim3 grid((h->g.M + K1_TX - 1) / K1_TX, (h->g.P + K1_TY - 1) / K1_TY, h->nm * 2);
    {
        KernelTimer t(h, QG_K_ZETA);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k1_zeta_step, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
            attr_done = true;
        }
        k1_zeta_step<<<grid, K1_THREADS, K1_SMEM_BYTES, h->stream>>>(h->tm_q, h->tm_psi, a);
    }
    h->qcur = nxt;
    return cudaGetLastError();
}

}  // namespace qg
