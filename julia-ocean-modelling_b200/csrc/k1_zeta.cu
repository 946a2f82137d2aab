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


template <int TY>
struct K1Cfg {
    static constexpr int BXP = K1_TX + 2 * GHOST;   // psi tile width  (halo 2)
    static constexpr int BYP = TY + 2 * GHOST;
    static constexpr int BXQ = K1_TX + 2 * GHOST;   // q tile width: halo 1 is needed, but the TMA box must start
                                                    // on a 16-byte boundary in global memory, so it spans halo 2 in x
    static constexpr int BXL = K1_TX + 2;           // lap tile width  (halo 1)
    static constexpr int BYQ = TY + 2;
    static constexpr int RPT = TY / (K1_THREADS / K1_TX);   // rows per thread
    static constexpr int PSI_BYTES = ((BXP * BYP * 8 + 127) / 128) * 128;
    static constexpr int Q_BYTES = ((BXQ * BYQ * 8 + 127) / 128) * 128;
    static constexpr int LAP_BYTES = ((BXL * BYQ * 8 + 127) / 128) * 128;
    static constexpr int SMEM = PSI_BYTES + Q_BYTES + LAP_BYTES + 128;
    static constexpr int MINB = TY >= 24 ? 2 : (TY >= 16 ? 3 : 4);
};

template <int TY>
__global__ void __launch_bounds__(K1_THREADS, K1Cfg<TY>::MINB)
k1_zeta_step(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_psi,
             const ZetaArgs a) {
    using Cfg = K1Cfg<TY>;
    extern __shared__ __align__(128) unsigned char k1_smem[];
    typedef double PsiRow[Cfg::BXP];
    typedef double QRow[Cfg::BXQ];
    typedef double LapRow[Cfg::BXL];
    PsiRow* s_psi = reinterpret_cast<PsiRow*>(k1_smem);
    QRow* s_q = reinterpret_cast<QRow*>(k1_smem + Cfg::PSI_BYTES);
    LapRow* s_lap = reinterpret_cast<LapRow*>(k1_smem + Cfg::PSI_BYTES + Cfg::Q_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(k1_smem + Cfg::PSI_BYTES + Cfg::Q_BYTES + Cfg::LAP_BYTES);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * K1_TX;
    const int y0 = blockIdx.y * TY;
    const int fz = blockIdx.z;          // member * 2 + layer
    const int layer = fz & 1;

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)((Cfg::BXP * Cfg::BYP + Cfg::BXQ * Cfg::BYQ) * sizeof(double)));
        tma_load_3d(&s_psi[0][0], &tm_psi, XPAD + x0 - GHOST, YPAD + y0 - GHOST, a.zpsi + fz, bar);
        tma_load_3d(&s_q[0][0], &tm_q, XPAD + x0 - GHOST, YPAD + y0 - 1, a.zq + fz, bar);
    }

    // While the tiles are in flight: fetch this thread's RHS history (pointwise, coalesced).
    const int lx = tid & (K1_TX - 1);
    const int ly0 = (tid / K1_TX) * Cfg::RPT;
    const int x = x0 + lx;
    const bool xin = x < a.g.M;
    const int64_t foff = (int64_t)fz * a.g.fstride;
    const double* __restrict__ f1 = a.f1 + foff;
    const double* __restrict__ f2 = a.f2 + foff;
    double h1[Cfg::RPT], h2[Cfg::RPT];
#pragma unroll
    for (int i = 0; i < Cfg::RPT; ++i) {
        const int y = y0 + ly0 + i;
        const bool in = xin && y < a.g.P && !a.euler;
        h1[i] = in ? __ldg(f1 + a.g.at(x, y)) : 0.0;
        h2[i] = in ? __ldg(f2 + a.g.at(x, y)) : 0.0;
    }
    mbar_wait(bar, 0);

    // Laplacian of psi on the tile plus a one-cell rim (same summation order as
    // src/schemes/laplacian.jl:21).
    for (int e = tid; e < Cfg::BYQ * Cfg::BXL; e += K1_THREADS) {
        const int ly = e / Cfg::BXL;
        const int lxx = e - ly * Cfg::BXL;
        const int sy = ly + 1, sx = lxx + 1;
        s_lap[ly][lxx] = (s_psi[sy][sx - 1] + s_psi[sy][sx + 1] - 4.0 * s_psi[sy][sx] +
                          s_psi[sy - 1][sx] + s_psi[sy + 1][sx]) * a.idx2;
    }
    __syncthreads();
    if (!xin) return;

    double* __restrict__ fn = a.fn + foff;
    double* __restrict__ qn = a.qn + foff;
    const double beta = a.beta[layer];
    const int sx = lx + GHOST;   // column in the psi tile
    const int qx = lx + GHOST;   // column in the q tile
    const int px = lx + 1;       // column in the lap tile

    // Rolling 3x3 windows down the column: rows (s, c, n) = (y-1, y, y+1).
    double p_sw = s_psi[ly0 + 1][sx - 1], p_s = s_psi[ly0 + 1][sx], p_se = s_psi[ly0 + 1][sx + 1];
    double p_w = s_psi[ly0 + 2][sx - 1], p_c = s_psi[ly0 + 2][sx], p_e = s_psi[ly0 + 2][sx + 1];
    double q_sw = s_q[ly0][qx - 1], q_s = s_q[ly0][qx], q_se = s_q[ly0][qx + 1];
    double q_w = s_q[ly0 + 1][qx - 1], q_c = s_q[ly0 + 1][qx], q_e = s_q[ly0 + 1][qx + 1];
    double l_s = s_lap[ly0][px], l_c = s_lap[ly0 + 1][px];
    (void)p_c;

#pragma unroll
    for (int i = 0; i < Cfg::RPT; ++i) {
        const int ly = ly0 + i;
        const int y = y0 + ly;
        if (y >= a.g.P) break;
        const double p_nw = s_psi[ly + 3][sx - 1], p_n = s_psi[ly + 3][sx], p_ne = s_psi[ly + 3][sx + 1];
        const double q_nw = s_q[ly + 2][qx - 1], q_n = s_q[ly + 2][qx], q_ne = s_q[ly + 2][qx + 1];
        const double l_n = s_lap[ly + 2][px];
        const double l_w = s_lap[ly + 1][px - 1], l_e = s_lap[ly + 1][px + 1];

        // src/schemes/arakawa.jl:13-15, 28-33, 46-51, 59 (x = first index i, y = second index j)
        const double jpp = (q_e - q_w) * (p_n - p_s) - (q_n - q_s) * (p_e - p_w);
        const double jpt = q_e * (p_ne - p_se) - q_w * (p_nw - p_sw) - q_n * (p_ne - p_nw) +
                           q_s * (p_se - p_sw);
        const double jtp = q_ne * (p_n - p_e) - q_sw * (p_w - p_s) - q_nw * (p_n - p_w) +
                           q_se * (p_e - p_s);
        const double jac = ((jpp + jpt) + jtp) * a.i12dx2;

        const double lap2 = (l_w + l_e - 4.0 * l_c + l_s + l_n) * a.idx2;
        const double dpsi = a.hdx * (p_e - p_w);
        // src/model.jl:144 / :152, evaluated left to right
        double rhs = (a.visc * lap2 - jac) - beta * dpsi;
        if (layer == 0)
            rhs -= a.U * (a.hdx * (q_e - q_w));
        else
            rhs -= a.r * l_c;

        const int64_t o = a.g.at(x, y);
        double qnew;
        if (a.euler)
            qnew = q_c + a.dt * rhs;   // src/model.jl:126
        else
            qnew = q_c + a.dt * ((a.c1 * rhs - a.c2 * h1[i]) + a.c3 * h2[i]);   // src/model.jl:134-135
        fn[o] = rhs;
        qn[o] = qnew;
        // periodic images (update_doubly_periodic_bc!, src/schemes/boundary_conditions.jl:2-13,
        // widened to two ghost cells)
        const bool gl = x < GHOST, gr = x >= a.g.M - GHOST;
        const bool gb = a.qimg_lo != nullptr && y < GHOST, gt = a.qimg_hi != nullptr && y >= a.g.P - GHOST;
        if (gl) qn[o + a.g.M] = qnew;
        if (gr) qn[o - a.g.M] = qnew;
        if (gb | gt) {   // own array (periodic in y) or the ring neighbour's, over NVLink (y-slab mode)
            const int64_t dyo = (int64_t)a.g.P * a.g.pitch;
            if (gb) {
                double* __restrict__ im = a.qimg_lo + foff;
                im[o + dyo] = qnew;
                if (gl) im[o + dyo + a.g.M] = qnew;
                if (gr) im[o + dyo - a.g.M] = qnew;
            }
            if (gt) {
                double* __restrict__ im = a.qimg_hi + foff;
                im[o - dyo] = qnew;
                if (gl) im[o - dyo + a.g.M] = qnew;
                if (gr) im[o - dyo - a.g.M] = qnew;
            }
            // y-slab peer mode: these were posted writes over NVLink.  The thread waits for their acknowledgement
            // (at most twice: it owns at most two edge rows), so that the flag barrier enqueued behind this kernel
            // can never overtake them - the barrier kernel's own fence covers only its own thread's writes.  (Placed
            // here and not behind the loop: code behind the loop costs the main path 7 registers.)
            if (!a.periodic_y) __threadfence_system();
        }
        // roll the windows
        p_sw = p_w; p_s = p_c; p_se = p_e; p_w = p_nw; p_c = p_n; p_e = p_ne;
        q_sw = q_w; q_s = q_c; q_se = q_e; q_w = q_nw; q_c = q_n; q_e = q_ne;
        l_s = l_c; l_c = l_n;
    }
}

template <int TY>
static cudaError_t launch_zeta_ty(Handle* h, const ZetaArgs& a) {
    using Cfg = K1Cfg<TY>;
    static bool attr_done_dev[QG_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[dev_slot(h)];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k1_zeta_step<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((h->g.M + K1_TX - 1) / K1_TX, (h->g.P + TY - 1) / TY, h->nm * 2);
    k1_zeta_step<TY><<<grid, K1_THREADS, Cfg::SMEM, h->stream>>>(h->tm_q, h->tm_psi, a);
    return cudaGetLastError();
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
    a.periodic_y = h->dist_n > 1 ? 0 : 1;
    a.qimg_lo = a.qimg_hi = nullptr;
    if (h->dist_n == 1) {
        a.qimg_lo = a.qimg_hi = a.qn;
    } else if (h->peer_ok) {   // rows [0,2) -> the rank below, rows [P-2,P) -> the rank above (periodic ring)
        const int64_t off = a.qn - h->q;
        a.qimg_lo = h->peer_q[(h->dist_rank + h->dist_n - 1) % h->dist_n] + off;
        a.qimg_hi = h->peer_q[(h->dist_rank + 1) % h->dist_n] + off;
    }
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
    cudaError_t e;
    {
        KernelTimer t(h, QG_K_ZETA);
        switch (h->k1_ty) {
            case 8: e = launch_zeta_ty<8>(h, a); break;
            case 12: e = launch_zeta_ty<12>(h, a); break;
            case 24: e = launch_zeta_ty<24>(h, a); break;
            default: e = launch_zeta_ty<16>(h, a); break;
        }
    }
    h->qcur = nxt;
    return e;
}

}  // namespace qg
