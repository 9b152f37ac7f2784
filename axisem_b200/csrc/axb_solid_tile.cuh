// axb_solid_tile.cuh — S_A, the solid element kernel (included by axb_kernels.cuh).
//
// Replaces, per (sub)step: the Newmark predictor / symplectic drift of the solid
// displacement (time_evol_wave.F90:359-364 / :599-602), the axis masks
// (apply_masks.f90:55-100), glob_stiffness_{mono,di,quad}_4 (stiffness_mono.f90:60-157,
// stiffness_di.f90:60-256, stiffness_quad.f90:238-412), glob_anel_stiffness_*_cg4
// (stiffness_mono.f90:511-589, stiffness_di.f90:727-825, stiffness_quad.f90:672-774) and
// time_step_memvars_cg4 (attenuation.f90:81-202, :471-535).
//
// Design: the kernel is HBM-bound (about 1.6 flop/byte), so the
// job is to keep enough bytes in flight.  Elements are processed in tiles of TE elements
// (TP = 25*TE points).  Every array the tile needs is contiguous in HBM for that tile
// (state planes are point-major; the 15/24/26 coefficient planes are stored as one slab
// per tile), and one producer lane streams tiles into a ring of shared-memory stages with
// 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) that complete on an mbarrier.  The ring
// depth — not registers or occupancy — sets the bytes in flight.  NCW consumer warps map
// thread t to point t of the tile (all lanes busy, all global stores fully coalesced);
// the 5x5 contractions of unrolled_loops.f90:164-188 read their operands from shared
// memory (conflict-free: strides 1 and 5 words), and the first-stage results S1*/S2* are
// exchanged through the stage slots of velo/acc0, which are dead after the predictor.
//
// Arithmetic mirrors oracle/axisem_oracle.c statement by statement; with -fmad=false the
// results are bit-identical to it.
#pragma once

namespace axb {

#ifndef AXB_TES
#define AXB_TES 8
#endif
#ifndef AXB_SOLID_CTAS
#define AXB_SOLID_CTAS 3
#endif
constexpr int TES = AXB_TES;             // elements per solid tile (multiple of 4: 16-byte TMA granules)
constexpr int TPS = TES * NPT;           // points per solid tile
constexpr int NCWS = (TPS + 31) / 32;    // consumer warps
constexpr int NCTS = NCWS * 32;          // consumer threads
constexpr int SOLID_THREADS = NCTS + 32; // + one producer warp
constexpr int SOLID_CTAS_PER_SM = AXB_SOLID_CTAS;   // resident CTAs per SM the ring is sized for
constexpr int MAX_STAGES = 8;
constexpr int NCG = 12;                  // planes of the coarse-grained attenuation slab
// fluid tiles (axb_fluid_tile.cuh) and the layout-conversion kernels
constexpr int TE = 16;
constexpr int TP = TE * NPT;
constexpr int NCW = (TP + 31) / 32;
constexpr int NCT = NCW * 32;
constexpr int FLUID_THREADS = NCT + 32;
static_assert(TES % 4 == 0 && TE % 4 == 0, "tiles must be multiples of 4 elements");

// plane order inside the coefficient slab [tile][plane][TPS]
enum {
    C_M11s = 0, C_M21s, C_M41s, C_M12s, C_M22s, C_M32s, C_M42s, C_M11z, C_M21z, C_M41z,
    C_M_1, C_M_2, C_M_3, C_M_4, C_M_w1,                       // 15: every source order
    C_M13s = 15, C_M33s = 16, C_M43s = 17,                    // dipole
    C_M1phi = 15, C_M2phi = 16, C_M4phi = 17,                 // quadrupole
    C_M_5 = 18, C_M_6, C_M_7, C_M_8, C_M_w2, C_M_w3,          // dipole + quadrupole
    C_M_w4 = 24, C_M_w5 = 25                                  // quadrupole
};
// plane order inside the attenuation slab [tile][plane][TES*4]
enum { G_Y = 0, G_Vse, G_Vsx, G_Vze, G_Vzx, G_Dse, G_Dze, G_Dsx, G_Dzx, G_dmu, G_dka,
       G_invs };   // G_invs: inv_s_solid sampled at the four coarse points

__host__ __device__ constexpr int solid_ncomp(int order) { return order == 0 ? 2 : 3; }
__host__ __device__ constexpr int solid_nplanes(int order) { return order == 0 ? 15 : (order == 1 ? 24 : 26); }

// float offsets of one ring stage
struct SolidTileLayout {
    int u, coef, meta, cg, sdev, str, mv, floats;
    size_t stage_bytes, hdr_bytes;
};
__host__ __device__ constexpr SolidTileLayout solid_tile_layout(int order, bool anel, int n_sls) {
    SolidTileLayout L{};
    int o = 0;
    L.u = o; o += solid_ncomp(order) * 3 * TPS;       // [comp][disp|velo|acc0][TPS]
    L.coef = o; o += solid_nplanes(order) * TPS;
    L.meta = o; o += (3 * TES + 3) / 4 * 4;           // ints: axis, qidx_mu, qidx_ka
    L.cg = L.sdev = L.str = L.mv = o;
    if (anel) {
        L.cg = o; o += NCG * TES * 4;
        L.sdev = o; o += TES * 24;
        L.str = o; o += TES * 4;
        L.mv = o; o += TES * 24 * n_sls;
    }
    L.floats = o;
    L.stage_bytes = ((size_t)o * 4 + 127) / 128 * 128;
    L.hdr_bytes = ((size_t)640 + (size_t)TES * 88 * 4 + 127) / 128 * 128;
    return L;
}

struct SolidTileArgs {
    int ntiles;
    int mode;                 // 0: Newmark predictor, 1: symplectic drift, 2: none (op test),
                              // 3: lean Newmark (velo holds v + dt/2 a: disp += dt * velo, acc0 not read)
    int do_stiff;             // 0: skip elastic stiffness (anel-only op test keeps acc1)
    int anel;                 // 0 none, 1 cg4 stiffness only, 2 stiffness + memvar update, 3 update only
    int nst;                  // ring depth
    int n_sls;
    int dbg;                  // developer diagnostics: 1 = consumers only drain the ring (pure load rate)
    double dt, half_dt_sq;    // Newmark: dt, dt^2/2 ; symplectic: coefd in dt
    float *disp, *velo, *acc0, *acc1;
    size_t cs;                // component stride = 25 * padded element count
    const float *coef;        // [tile][plane][TPS]
    const int *meta;          // [tile][3][TES]
    const float *M0_w[10];    // axial vectors (5, nel_pad); index = number - 1
    const float *cg;          // [tile][NCG][TES*4]
    // per distinct Q and SLS: {ts_fac_t * a_j, ts_fac_tm1 * a_j} (attenuation.f90:162-175 evaluates
    // ts_fac_t(j) * a_j_mu(j) * src left to right, so the first product can be formed once)
    const double2 *c_mu_tab, *c_ka_tab;
    // the tables are small (one row per distinct Q): when they fit they are staged in shared
    // memory behind the header, and the ring starts at ring_off
    int tab_smem, ntab_mu, ntab_ka;
    int ring_off;             // byte offset of the ring in dynamic shared memory
    int emask;                // mode 2 only: axis masks on the input copy and on K u (energy, time_evol_wave.F90:1459-1471)
    double exp_w[8];          // exp(-w_j deltat) per SLS (constant bank)
    float *memvar, *src_dev_tm1, *src_tr_tm1;
};

// ---- PTX wrappers: mbarrier + bulk async copy ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or
// the hint runs out) instead of re-issuing the probe every few hundred cycles — the probes of
// waiting warps otherwise take issue slots from the warps that work (they were 8 % of all
// instructions of S_A: profiles/r01k vs r02c)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void bar_consumers() {
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

// ---- contractions over shared memory (unrolled_loops.f90:164-207, k ascending) ----------
// sum_k c[k] * p[k]        p -> val(0, j): the xi-line through this point
__device__ __forceinline__ float cxi(const float *p, const float (&c)[NP]) {
    float s = c[0] * p[0];
    s = s + c[1] * p[1];
    s = s + c[2] * p[2];
    s = s + c[3] * p[3];
    s = s + c[4] * p[4];
    return s;
}
// sum_k p[5k] * c[k]       p -> val(i, 0): the eta-line through this point
__device__ __forceinline__ float ceta(const float *p, const float (&c)[NP]) {
    float s = p[0] * c[0];
    s = s + p[5] * c[1];
    s = s + p[10] * c[2];
    s = s + p[15] * c[3];
    s = s + p[20] * c[4];
    return s;
}

struct PointG {               // derivative-matrix rows/columns of this thread's point
    float g2t_row[NP];        // G2T(i,k)  first stage, xi, non-axial
    float g2_col[NP];         // G2(k,j)   first stage, eta
    float g2_row[NP];         // G2(i,k)   second stage, xi, non-axial
    float g2t_col[NP];        // G2T(k,j)  second stage, eta
};
// axial rows are rare (two element columns of the mesh): read from shared memory on demand
__device__ __forceinline__ void axial_rows(const GMat &sG, int i, float (&g1t_row)[NP],
                                           float (&g1_row)[NP], float (&g0)[NP]) {
#pragma unroll
    for (int k = 0; k < NP; k++) {
        g1t_row[k] = sG.G1T[i + NP * k];
        g1_row[k] = sG.G1[i + NP * k];
        g0[k] = sG.G0[k];
    }
}

// =======================================================================================
// NSLS: number of standard linear solids compiled in (0: elastic, -1: run-time a.n_sls)
template <int ORDER, int NSLS>
__global__ void __launch_bounds__(SOLID_THREADS, SOLID_CTAS_PER_SM)
k_solid_tile(const __grid_constant__ GMat G, const __grid_constant__ SolidTileArgs a) {
    constexpr int NC = solid_ncomp(ORDER);
    constexpr int NPL = solid_nplanes(ORDER);
    const int n_sls = NSLS >= 0 ? NSLS : a.n_sls;
    const int anel = NSLS == 0 ? 0 : a.anel;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + MAX_STAGES;
    GMat &sG = *reinterpret_cast<GMat *>(smem + 128);
    float *x_rsum = reinterpret_cast<float *>(smem + 640);     // [TES][24]
    float *x_anS = x_rsum + TES * 24;                            // [TES][36]
    float *x_src = x_anS + TES * 36;                             // [TES][28]
    const SolidTileLayout Ly = solid_tile_layout(ORDER, NSLS != 0, n_sls);
    unsigned char *ring = smem + a.ring_off;
    double2 *s_tab = reinterpret_cast<double2 *>(smem + Ly.hdr_bytes);   // [mu rows | kappa rows][n_sls]

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    if (t == 0) {
        for (int s = 0; s < a.nst; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCWS); }
        fence_mbar_init();
    }
    {   // stage the derivative matrices
        const float *src = reinterpret_cast<const float *>(&G);
        float *dst = reinterpret_cast<float *>(&sG);
        for (int k = t; k < (int)(sizeof(GMat) / sizeof(float)); k += blockDim.x) dst[k] = src[k];
        if (NSLS != 0 && a.tab_smem) {
            const int nmu = a.ntab_mu * n_sls, nka = a.ntab_ka * n_sls;
            for (int k = t; k < nmu; k += blockDim.x) s_tab[k] = a.c_mu_tab[k];
            for (int k = t; k < nka; k += blockDim.x) s_tab[nmu + k] = a.c_ka_tab[k];
        }
    }
    __syncthreads();

    const bool anel_stiff = anel == 1 || anel == 2;
    const bool anel_update = anel >= 2;

    // ---------------------------------------------------------------- producer warp ----
    if (warp == NCWS) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t plane_b = TPS * 4;
            uint32_t bytes = NC * plane_b * (a.mode == 0 ? 3 : (a.mode == 2 ? 1 : 2)) + 3 * TES * 4;
            if (a.do_stiff) bytes += NPL * plane_b;
            if (anel) {
                bytes += NCG * TES * 16 + TES * 96 * n_sls;
                if (anel_update) bytes += TES * 96 + TES * 16;
            }
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                float *S = reinterpret_cast<float *>(ring + (size_t)s * Ly.stage_bytes);
                uint64_t *bar = &full[s];
                mbar_expect_tx(bar, bytes);
                const size_t pg = (size_t)tile * TPS;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const size_t off = (size_t)((ORDER == 0) ? 2 * c : c) * a.cs + pg;
                    float *d = S + Ly.u + c * 3 * TPS;
                    bulk_g2s(d, a.disp + off, plane_b, bar);
                    if (a.mode != 2) bulk_g2s(d + TPS, a.velo + off, plane_b, bar);
                    if (a.mode == 0) bulk_g2s(d + 2 * TPS, a.acc0 + off, plane_b, bar);
                }
                if (a.do_stiff) bulk_g2s(S + Ly.coef, a.coef + (size_t)tile * NPL * TPS, NPL * plane_b, bar);
                bulk_g2s(S + Ly.meta, a.meta + (size_t)tile * 3 * TES, 3 * TES * 4, bar);
                if (anel) {
                    bulk_g2s(S + Ly.cg, a.cg + (size_t)tile * NCG * TES * 4, NCG * TES * 16, bar);
                    bulk_g2s(S + Ly.mv, a.memvar + (size_t)tile * TES * 24 * n_sls, TES * 96 * n_sls, bar);
                    if (anel_update) {
                        bulk_g2s(S + Ly.sdev, a.src_dev_tm1 + (size_t)tile * TES * 24, TES * 96, bar);
                        bulk_g2s(S + Ly.str, a.src_tr_tm1 + (size_t)tile * TES * 4, TES * 16, bar);
                    }
                }
                if (++s == a.nst) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // --------------------------------------------------------------- consumer warps ----
    const bool pt = t < TPS;                     // this thread owns point t of the tile
    const int el = pt ? t / NPT : 0;            // element inside the tile
    const int q = pt ? t - el * NPT : 0;
    const int i = q % NP, j = q / NP;
    const int e25 = el * NPT;
    PointG L;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        L.g2t_row[k] = sG.G2T[i + NP * k];
        L.g2_col[k] = sG.G2[k + NP * j];
        L.g2_row[k] = sG.G2[i + NP * k];
        L.g2t_col[k] = sG.G2T[k + NP * j];
    }
    const float g0_i = sG.G0[i];
    const bool rowa = (i == 1) || (i == 3), colb = (j == 1) || (j == 3);
    const bool cgpt = pt && rowa && colb;
    const int cgk = (i == 1 ? 0 : 2) + (j == 1 ? 0 : 1);       // coarse index of (i,j)
    // memory-variable role: t < 24*TES  <->  (element t/24, l = t%24 = 4*v + k)
    const bool mvt = t < TES * 24;
    const int mel = mvt ? t / 24 : 0;
    const int ml = mvt ? t - mel * 24 : 0;
    const int mv_v = ml >> 2, mv_k = ml & 3;
    const bool mv_lane = mvt && !(ORDER == 0 && (mv_v == 3 || mv_v == 5));

    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        mbar_wait(&full[s], ph);
        if (a.dbg == 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == a.nst) { s = 0; ph ^= 1; }
            continue;
        }
        float *S = reinterpret_cast<float *>(ring + (size_t)s * Ly.stage_bytes);
        const int *meta = reinterpret_cast<const int *>(S + Ly.meta);
        const bool ax = meta[el] != 0;
        const size_t pg = (size_t)tile * TPS + t;
        const int eg = tile * TES + el;
        float *Ub = S + Ly.u;                   // slot (c, k) at Ub + (3c + k) * TPS
        const float *Cf = S + Ly.coef + t;      // coefficient n of this point: Cf[n * TPS]

        // ---- phase 1: predictor / drift, axis mask -> U in shared memory + disp in HBM ----
        float u1 = 0.f, u2 = 0.f, u3 = 0.f;
        if (pt) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const float *sl = Ub + c * 3 * TPS + t;
                float x = sl[0];
                if (a.mode == 0)
                    x = d2f(f2d(x) + a.dt * f2d(sl[TPS]) + a.half_dt_sq * f2d(sl[2 * TPS]));
                else if (a.mode == 1)
                    x = d2f(f2d(x) + f2d(sl[TPS]) * a.dt);
                else if (a.mode == 3)
                    x = d2f(f2d(x) + a.dt * f2d(sl[TPS]));
                if (ORDER == 0) { if (c == 0) u1 = x; else u3 = x; }
                else { if (c == 0) u1 = x; else if (c == 1) u2 = x; else u3 = x; }
            }
            // apply_axis_mask_{one,two,three}comp (apply_masks.f90:55-100)
            if (ax && i == 0 && (a.mode != 2 || a.emask)) {
                if (ORDER == 0) u1 = 0.f;
                else if (ORDER == 1) { u2 = 0.f; u3 = 0.f; }
                else { u1 = 0.f; u2 = 0.f; u3 = 0.f; }
            }
            if (a.mode != 2 || a.emask) {
                if (ORDER == 0) { Ub[t] = u1; Ub[3 * TPS + t] = u3; }
                else { Ub[t] = u1; Ub[3 * TPS + t] = u2; Ub[6 * TPS + t] = u3; }
            }
            if (a.mode != 2) {
                if (ORDER == 0) { a.disp[pg] = u1; a.disp[pg + 2 * a.cs] = u3; }
                else { a.disp[pg] = u1; a.disp[pg + a.cs] = u2; a.disp[pg + 2 * a.cs] = u3; }
            }
        }
        if (anel && mvt) {
            // r(v)(k) = sum over the standard linear solids (stiffness_mono.f90:545-549)
            float rsum = 0.0f;
            if (mv_lane) {
                const float *mv = S + Ly.mv + mel * 24 * n_sls + ml;
#pragma unroll
                for (int sl = 0; sl < n_sls; sl++) rsum = rsum + mv[24 * sl];
            }
            x_rsum[t] = rsum;
        }
        bar_consumers<NCTS>();

        // ---- phase 2: first-stage contractions, point-wise combinations -> S planes ----
        const int cu3 = (ORDER == 0) ? 1 : 2;   // slot row of the z component
        const float *U1xi = Ub + e25 + 5 * j, *U1et = Ub + e25 + i;
        const float *U2xi = U1xi + 3 * TPS, *U2et = U1et + 3 * TPS;              // ORDER != 0
        const float *U3xi = U1xi + cu3 * 3 * TPS, *U3et = U1et + cu3 * 3 * TPS;
        float l1 = 0.f, l2 = 0.f, l3 = 0.f;
        float X[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (pt) {
            float g1t_row[NP], g1_row[NP], g0[NP];
            if (ax) axial_rows(sG, i, g1t_row, g1_row, g0);
            if (a.do_stiff || anel_update) {
                if (!ax) { X[0] = cxi(U1xi, L.g2t_row); X[2] = cxi(U3xi, L.g2t_row); }
                else     { X[0] = cxi(U1xi, g1t_row);   X[2] = cxi(U3xi, g1t_row); }
                X[3] = ceta(U1et, L.g2_col);
                X[5] = ceta(U3et, L.g2_col);
                if (ORDER != 0) {
                    X[1] = ax ? cxi(U2xi, g1t_row) : cxi(U2xi, L.g2t_row);
                    X[4] = ceta(U2et, L.g2_col);
                }
            }
            if (a.do_stiff) {
                const float m11s = Cf[C_M11s * TPS], m21s = Cf[C_M21s * TPS], m41s = Cf[C_M41s * TPS];
                const float m12s = Cf[C_M12s * TPS], m22s = Cf[C_M22s * TPS], m32s = Cf[C_M32s * TPS];
                const float m42s = Cf[C_M42s * TPS];
                const float m11z = Cf[C_M11z * TPS], m21z = Cf[C_M21z * TPS], m41z = Cf[C_M41z * TPS];
                const float m_1 = Cf[C_M_1 * TPS], m_2 = Cf[C_M_2 * TPS], m_3 = Cf[C_M_3 * TPS], m_4 = Cf[C_M_4 * TPS];
                const float m_w1 = Cf[C_M_w1 * TPS];
                float *Sb = Ub + t;              // S planes: slots (c,1) and (c,2)
                if (ORDER == 0) {
                    // stiffness_mono.f90:60-157
                    const float X1 = X[0], X2 = X[2], X3 = X[3], X4 = X[5], us = u1;
                    l1 = m_4 * X4 + m_2 * X3 + m_1 * X1 + m_3 * X2 + us * m_w1;
                    const float S1s = m11s * X3 + m21s * X1 + m12s * X4 + m22s * X2 + m_1 * us;
                    const float S2s = m11s * X1 + m41s * X3 + m32s * X2 + m42s * X4 + m_2 * us;
                    const float S1z = m11z * X4 + m21z * X2 + m32s * X3 + m22s * X1 + m_3 * us;
                    const float S2z = m11z * X2 + m41z * X4 + m12s * X1 + m42s * X3 + m_4 * us;
                    Sb[1 * TPS] = S1s; Sb[2 * TPS] = S2s; Sb[4 * TPS] = S1z; Sb[5 * TPS] = S2z;
                } else if (ORDER == 1) {
                    // stiffness_di.f90:60-256
                    const float m13s = Cf[C_M13s * TPS], m23s = m32s, m33s = Cf[C_M33s * TPS], m43s = Cf[C_M43s * TPS];
                    const float m_5 = Cf[C_M_5 * TPS], m_6 = Cf[C_M_6 * TPS], m_7 = Cf[C_M_7 * TPS], m_8 = Cf[C_M_8 * TPS];
                    const float m_w2 = Cf[C_M_w2 * TPS], m_w3 = Cf[C_M_w3 * TPS];
                    const float X1 = X[0], X2 = X[1], X3 = X[2], X4 = X[3], X5 = X[4], X6 = X[5];
                    const float X7 = X1 + X2;
                    const float X8 = X4 + X5;
                    l2 = m_8 * X6 + m_7 * X3 + m_1 * X1 + m_5 * X2 + m_2 * X4 + m_6 * X5 + m_w1 * u2 + m_w2 * u3;
                    l3 = m_4 * X4 - m_4 * X5 + m_3 * X1 - m_3 * X2 + m_w2 * u2 + m_w3 * u3;
                    float c1 = m13s * X6, c2 = m23s * X3, c3 = m_3 * u3;
                    const float S1p = c1 + c2 + c3 + m11s * X4 + m21s * X1 + m12s * X5 + m22s * X2 + m_1 * u2;
                    const float S1m = c1 + c2 - c3 + m11s * X5 + m21s * X2 + m12s * X4 + m22s * X1 + m_5 * u2;
                    c1 = m33s * X3; c2 = m43s * X6; c3 = m_4 * u3;
                    const float S2p = c1 + c2 + c3 + m11s * X1 + m41s * X4 + m12s * X2 + m42s * X5 + m_2 * u2;
                    const float S2m = c1 + c2 - c3 + m11s * X2 + m41s * X5 + m12s * X1 + m42s * X4 + m_6 * u2;
                    const float S1z = m33s * X8 + m23s * X7 + m11z * X6 + m21z * X3 + m_7 * u2;
                    const float S2z = m13s * X7 + m43s * X8 + m11z * X3 + m41z * X6 + m_8 * u2;
                    Sb[1 * TPS] = S1p; Sb[2 * TPS] = S2p; Sb[4 * TPS] = S1m; Sb[5 * TPS] = S2m;
                    Sb[7 * TPS] = S1z; Sb[8 * TPS] = S2z;
                } else {
                    // stiffness_quad.f90:238-412
                    const float m1phi = Cf[C_M1phi * TPS], m2phi = Cf[C_M2phi * TPS], m4phi = Cf[C_M4phi * TPS];
                    const float m_5 = Cf[C_M_5 * TPS], m_6 = Cf[C_M_6 * TPS], m_7 = Cf[C_M_7 * TPS], m_8 = Cf[C_M_8 * TPS];
                    const float m_w2 = Cf[C_M_w2 * TPS], m_w3 = Cf[C_M_w3 * TPS];
                    const float m_w4 = Cf[C_M_w4 * TPS], m_w5 = Cf[C_M_w5 * TPS];
                    const float X1 = X[0], X2 = X[1], X3 = X[2], X4 = X[3], X5 = X[4], X6 = X[5];
                    const float us = u1, up = u2, uz = u3;
                    const float c1 = m_2 * X4, c2 = m_1 * X1, c3 = m_6 * X5, c4 = m_5 * X2, c5 = m_4 * X6, c6 = m_3 * X3;
                    l1 = c1 + c2 + 2 * (c3 + c4) + c5 + c6 + m_w1 * us + m_w2 * up + 2 * m_w3 * uz;
                    l2 = -2 * (c1 + c2 + c5 + c6) - (c3 + c4) + m_w2 * us + m_w4 * up - m_w3 * uz;
                    l3 = 2 * (m_8 * X5 + m_7 * X2) + m_w3 * (2 * us - up) + m_w5 * uz;
                    const float S1s = m11s * X4 + m21s * X1 + m12s * X6 + m22s * X3 + m_1 * (us - 2 * up);
                    const float S2s = m11s * X1 + m41s * X4 + m32s * X3 + m42s * X6 + m_2 * (us - 2 * up);
                    const float S1z = m11z * X6 + m21z * X3 + m32s * X4 + m22s * X1 + m_3 * (us - 2 * up);
                    const float S2z = m11z * X3 + m41z * X6 + m12s * X1 + m42s * X4 + m_4 * (us - 2 * up);
                    const float S1p = m1phi * X5 + m2phi * X2 + m_5 * (2 * us - up) + 2 * m_7 * uz;
                    const float S2p = m1phi * X2 + m4phi * X5 + m_6 * (2 * us - up) + 2 * m_8 * uz;
                    Sb[1 * TPS] = S1s; Sb[2 * TPS] = S2s; Sb[4 * TPS] = S1p; Sb[5 * TPS] = S2p;
                    Sb[7 * TPS] = S1z; Sb[8 * TPS] = S2z;
                }
            } else {
                // anelastic-only operator test: start from the stored acc1
                l1 = a.acc1[pg];
                if (ORDER != 0) l2 = a.acc1[pg + a.cs];
                l3 = a.acc1[pg + 2 * a.cs];
            }
            // ---- strain at the coarse points (compute_strain_att_el_cg4, attenuation.f90:471-535)
            if (anel_update && cgpt) {
                const float *cg = S + Ly.cg + el * 4 + cgk;
                const float dzdeta = cg[G_Dze * TES * 4], dzdxi = cg[G_Dzx * TES * 4];
                const float dsdeta = cg[G_Dse * TES * 4], dsdxi = cg[G_Dsx * TES * 4];
                const float is = cg[G_invs * TES * 4];
                float g1, g2, g3, g4 = 0.f, g5, g6 = 0.f;
                // gradient of f: ds = dzdeta*m1 + dzdxi*m2 ; dz = dsdeta*m1 + dsdxi*m2
                const float b2s = dzdeta * X[2] + dzdxi * X[5];     // d_s u3
                const float b2z = dsdeta * X[2] + dsdxi * X[5];     // d_z u3
                if (ORDER == 0) {
                    const float b1s = dzdeta * X[0] + dzdxi * X[3];
                    const float b1z = dsdeta * X[0] + dsdxi * X[3];
                    g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                    g2 = is * u1;
                } else if (ORDER == 1) {
                    // the gradients of (u1+u2) and (u1-u2) are contracted separately in the
                    // reference (attenuation.f90:489, :515); the bit-exact build does the same.
                    // The product build uses the linearity of the contraction, d(u1 +/- u2) =
                    // d u1 +/- d u2, with the derivatives every point already has: four extra
                    // 5-term contractions per warp less (they were 6 % of this kernel's
                    // instructions, executed at 4/25 lane efficiency); differs by rounding only
                    float Xp1, Xm1, Xp2, Xm2;
#ifdef AXB_STRICT
                    float up[NP], um[NP];
#pragma unroll
                    for (int k = 0; k < NP; k++) { up[k] = U1xi[k] + U2xi[k]; um[k] = U1xi[k] - U2xi[k]; }
                    if (!ax) { Xp1 = cxi(up, L.g2t_row); Xm1 = cxi(um, L.g2t_row); }
                    else     { Xp1 = cxi(up, g1t_row);   Xm1 = cxi(um, g1t_row); }
                    float vp[21], vm[21];
#pragma unroll
                    for (int k = 0; k < NP; k++) { vp[5 * k] = U1et[5 * k] + U2et[5 * k]; vm[5 * k] = U1et[5 * k] - U2et[5 * k]; }
                    Xp2 = ceta(vp, L.g2_col);
                    Xm2 = ceta(vm, L.g2_col);
#else
                    Xp1 = X[0] + X[1]; Xm1 = X[0] - X[1];
                    Xp2 = X[3] + X[4]; Xm2 = X[3] - X[4];
#endif
                    const float b1s = dzdeta * Xp1 + dzdxi * Xp2;
                    const float b1z = dsdeta * Xp1 + dsdxi * Xp2;
                    g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                    g2 = 2 * (is * u2);
                    const float c1s = dzdeta * Xm1 + dzdxi * Xm2;
                    const float c1z = dsdeta * Xm1 + dsdxi * Xm2;
                    g4 = -(is * u3) - c1z;
                    g6 = -g2 - c1s;
                } else {
                    const float b1s = dzdeta * X[0] + dzdxi * X[3];
                    const float b1z = dsdeta * X[0] + dsdxi * X[3];
                    g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                    g2 = is * (u1 - 2 * u2);
                    const float c1s = dzdeta * X[1] + dzdxi * X[4];   // gradient of u2
                    const float c1z = dsdeta * X[1] + dsdxi * X[4];
                    g4 = -2 * (is * u3) - c1z;
                    g6 = is * (u2 - 2 * u1) - c1s;
                }
                float trace = g1 + g2;
                trace = trace + g3;
                const float dmu = cg[G_dmu * TES * 4], dka = cg[G_dka * TES * 4];
                const double third = 1.0 / 3.0;
                const double dm2 = f2d(dmu * 2);
                float *src = x_src + el * 28 + cgk;
                src[0] = d2f(dm2 * (f2d(g1) - f2d(trace) * third));
                src[4] = d2f(dm2 * (f2d(g2) - f2d(trace) * third));
                src[8] = d2f(dm2 * (f2d(g3) - f2d(trace) * third));
                src[12] = (ORDER == 0) ? 0.0f : dmu * g4;
                src[16] = dmu * g5;
                src[20] = (ORDER == 0) ? 0.0f : dmu * g6;
                src[24] = dka * trace;                      // src_tr_t
            }
        }
        // ---- anelastic S terms at the four coarse points (glob_anel_stiffness_*_cg4) ----
        if (anel_stiff && t < TES * 4) {
            const int ce = t >> 2, ck = t & 3;
            const float *cg = S + Ly.cg + t;                // [plane][TES*4], index ce*4+ck = t
            const float yl = cg[G_Y * TES * 4];
            const float vse = cg[G_Vse * TES * 4], vsx = cg[G_Vsx * TES * 4];
            const float vze = cg[G_Vze * TES * 4], vzx = cg[G_Vzx * TES * 4];
            const float *r = x_rsum + ce * 24 + ck;
            const float r1 = r[0], r2 = r[4], r3 = r[8], r4 = r[12], r5 = r[16], r6 = r[20];
            float *Sa = x_anS + ce * 36 + ck;               // Sa[4a], a = 0..5 ; extras at 24,28,32
            if (ORDER == 0) {
                Sa[0] = vze * r1 + vse * r5;      // S1s
                Sa[4] = vzx * r1 + vsx * r5;      // S2s
                Sa[16] = vze * r5 + vse * r3;     // S1z
                Sa[20] = vzx * r5 + vsx * r3;     // S2z
                Sa[24] = yl * r2;
            } else if (ORDER == 1) {
                Sa[0] = vze * (r1 - r6) + vse * (r5 - r4);   // S1p
                Sa[4] = vzx * (r1 - r6) + vsx * (r5 - r4);   // S2p
                Sa[8] = vze * (r1 + r6) + vse * (r5 + r4);   // S1m
                Sa[12] = vzx * (r1 + r6) + vsx * (r5 + r4);  // S2m
                Sa[16] = vze * r5 + vse * r3;
                Sa[20] = vzx * r5 + vsx * r3;
                Sa[24] = 2 * yl * (r2 - r6);
                Sa[28] = yl * r4;
            } else {
                Sa[0] = vze * r1 + vse * r5;      // S1s
                Sa[4] = vzx * r1 + vsx * r5;      // S2s
                Sa[8] = vze * r6 + vse * r4;      // S1p
                Sa[12] = vzx * r6 + vsx * r4;     // S2p
                Sa[16] = vze * r5 + vse * r3;
                Sa[20] = vzx * r5 + vsx * r3;
                Sa[24] = yl * (r2 - 2 * r6);
                Sa[28] = yl * (r6 - 2 * r2);
                Sa[32] = 2 * yl * r4;
            }
        }
        bar_consumers<NCTS>();

        // ---- phase 3: second-stage contractions, axial terms, anelastic correction ----
        if (pt) {
            float g1t_row[NP], g1_row[NP], g0[NP];
            if (ax) axial_rows(sG, i, g1t_row, g1_row, g0);
            if (a.do_stiff) {
                const float *Sxi = Ub + e25 + 5 * j, *Set = Ub + e25 + i;
                const float Y1 = ax ? cxi(Sxi + 1 * TPS, g1_row) : cxi(Sxi + 1 * TPS, L.g2_row);
                const float Y2 = ceta(Set + 2 * TPS, L.g2t_col);
                const float Y3 = ax ? cxi(Sxi + 4 * TPS, g1_row) : cxi(Sxi + 4 * TPS, L.g2_row);
                const float Y4 = ceta(Set + 5 * TPS, L.g2t_col);
                if (ORDER == 0) {
                    l1 = l1 + Y1 + Y2;
                    l3 = Y3 + Y4;
                    if (ax) {
                        const size_t a0 = j + NP * (size_t)eg, b0 = NP * (size_t)eg;
                        const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0];
                        const float V1 = cxi(U1xi, g0);                    // vxm_4(G0, us)
                        const float V2 = ceta(Ub + 3 * TPS + e25, L.g2_col); // vxm_4(uz0, G2): uz(0,k)
                        float V4 = w1 * V1 + w3 * V2;
                        const float V3 = cxi(U3xi, g0);                    // vxm_4(G0, uz)
                        V4 = V4 + w2 * V3;
                        float X2a = g0_i * (w2 * V1);                      // outerprod_4(G0, m0_w2*V1)
                        if (i == 0) {
                            // vxm_4(V2, G2T) with V2(k) = m0_w3(k) * vxm_4(G0, us)(k)
                            float vb[NP];
#pragma unroll
                            for (int k = 0; k < NP; k++) vb[k] = a.M0_w[2][b0 + k] * cxi(Ub + e25 + 5 * k, g0);
                            X2a = X2a + cxi(vb, L.g2t_col);
                        }
                        l1 = l1 + g0_i * V4;
                        l3 = X2a + l3;
                    }
                } else {
                    const float Y5 = ax ? cxi(Sxi + 7 * TPS, g1_row) : cxi(Sxi + 7 * TPS, L.g2_row);
                    const float Y6 = ceta(Set + 8 * TPS, L.g2t_col);
                    if (ORDER == 1) {
                        l1 = Y1 + Y2;
                        l2 = Y3 + Y4 + l2;
                        l3 = Y5 + Y6 + l3;
                        if (ax) {
                            const size_t a0 = j + NP * (size_t)eg, b0 = NP * (size_t)eg;
                            const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0], w4 = a.M0_w[3][a0];
                            const float w6 = a.M0_w[5][a0], w7 = a.M0_w[6][a0], w8 = a.M0_w[7][a0], w9 = a.M0_w[8][a0];
                            const float w10 = a.M0_w[9][a0];
                            const float V1 = cxi(U1xi, g0), V2 = cxi(U2xi, g0), V3 = cxi(U3xi, g0);
                            const float V4 = ceta(Ub + e25, L.g2_col);      // vxm_4(u10, G2)
                            float s1p = g0_i * (w1 * V2 + w3 * V3);
                            const float s1m = g0_i * (w1 * V1 + (w2 + w6) * V4 + w9 * V2 + w10 * V3);
                            const float s1z = g0_i * (w3 * V1 + (w4 + w8) * V4 + w7 * V3 + w10 * V2);
                            if (i == 0) {
                                // vxm_4(V4, G2T), V4(k) = (w2+w6)(k) V2(k) + (w4+w8)(k) V3(k)
                                float vb[NP];
#pragma unroll
                                for (int k = 0; k < NP; k++) {
                                    const float k2 = a.M0_w[1][b0 + k], k6 = a.M0_w[5][b0 + k];
                                    const float k4 = a.M0_w[3][b0 + k], k8 = a.M0_w[7][b0 + k];
                                    vb[k] = (k2 + k6) * cxi(Ub + 3 * TPS + e25 + 5 * k, g0)
                                          + (k4 + k8) * cxi(Ub + 6 * TPS + e25 + 5 * k, g0);
                                }
                                s1p = s1p + cxi(vb, L.g2t_col);
                            }
                            l1 = l1 + s1p;
                            l2 = l2 + s1m;
                            l3 = l3 + s1z;
                        }
                    } else {
                        l1 = l1 + Y1 + Y2;
                        l2 = l2 + Y3 + Y4;
                        l3 = l3 + Y5 + Y6;
                        if (ax) {
                            const size_t a0 = j + NP * (size_t)eg;
                            const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0];
                            const float w4 = a.M0_w[3][a0], w5 = a.M0_w[4][a0], w6 = a.M0_w[5][a0];
                            const float V1 = cxi(U1xi, g0), V2 = cxi(U2xi, g0), V3 = cxi(U3xi, g0);
                            l1 = l1 + g0_i * (w1 * V1 + w2 * V2 + w3 * V3);
                            l2 = l2 + g0_i * (w2 * V1 + w4 * V2 + w5 * V3);
                            l3 = l3 + g0_i * (w3 * V1 + w5 * V2 + w6 * V3);
                        }
                    }
                }
            }
            if (anel_stiff) {
                const float *Sa = x_anS + el * 36;
                const float ga1 = ax ? g1_row[1] : L.g2_row[1];  // GA(i,1), GA(i,3)
                const float ga3 = ax ? g1_row[3] : L.g2_row[3];
                // mxm_cg4_sparse_b(GA, S1): c(i,1) = GA(i,1) S1(1) + GA(i,3) S1(3);
                //                           c(i,3) = GA(i,1) S1(2) + GA(i,3) S1(4)
                // mxm_cg4_sparse_a(S2, G2T): c(1,j) = S2(1) G2T(1,j) + S2(2) G2T(3,j);
                //                            c(3,j) = S2(3) G2T(1,j) + S2(4) G2T(3,j)
                const int kb = (j == 1) ? 0 : 1;
                const int ka = (i == 1) ? 0 : 2;
                float Xb[3], Xa[3];
#pragma unroll
                for (int m = 0; m < 3; m++) {
                    const float *S1 = Sa + 8 * m, *S2 = Sa + 8 * m + 4;
                    Xb[m] = colb ? (ga1 * S1[kb] + ga3 * S1[kb + 2]) : 0.0f;
                    Xa[m] = rowa ? (S2[ka] * L.g2t_col[1] + S2[ka + 1] * L.g2t_col[3]) : 0.0f;
                }
                const bool cg2 = rowa && colb;
                if (ORDER == 0) {
                    float ls = Xb[0] + Xa[0];
                    const float lz = Xb[2] + Xa[2];
                    if (cg2) ls = ls + Sa[24 + cgk];
                    l1 = l1 - ls;
                    l3 = l3 - lz;
                } else if (ORDER == 1) {
                    const float lp = Xb[0] + Xa[0];
                    float lm = Xb[1] + Xa[1];
                    float lz = Xb[2] + Xa[2];
                    if (cg2) { lm = lm + Sa[24 + cgk]; lz = lz - Sa[28 + cgk]; }
                    l1 = l1 - lp; l2 = l2 - lm; l3 = l3 - lz;
                } else {
                    float ls = Xb[0] + Xa[0];
                    float lp = -Xb[1] - Xa[1];
                    float lz = Xb[2] + Xa[2];
                    if (cg2) { ls = ls + Sa[24 + cgk]; lp = lp + Sa[28 + cgk]; lz = lz - Sa[32 + cgk]; }
                    l1 = l1 - ls; l2 = l2 - lp; l3 = l3 - lz;
                }
            }
            if (a.do_stiff || anel_stiff) {
                // apply_axis_mask_*(acc1) (time_evol_wave.F90:438-447); k_bdry2solid re-applies
                // it to the few points the S/F term touches afterwards
                if (ax && i == 0 && (a.mode != 2 || a.emask)) {
                    if (ORDER == 0) l1 = 0.f;
                    else if (ORDER == 1) { l2 = 0.f; l3 = 0.f; }
                    else { l1 = 0.f; l2 = 0.f; l3 = 0.f; }
                }
                a.acc1[pg] = l1;
                if (ORDER != 0) a.acc1[pg + a.cs] = l2;
                a.acc1[pg + 2 * a.cs] = l3;
            }
        }
        // ---- memory-variable update (time_step_memvars_cg4, attenuation.f90:136-200) ----
        if (anel_update && mvt) {
            const float src_dev_t = x_src[mel * 28 + ml];
            const float s_dev_tm1 = S[Ly.sdev + t];
            const size_t meg = (size_t)tile * TES + mel;
            if (mv_lane) {
                const double src_tr_t = f2d(x_src[mel * 28 + 24 + mv_k]);
                const double s_tr_tm1 = f2d(S[Ly.str + mel * 4 + mv_k]);
                const double dsrc_t = f2d(src_dev_t), dsrc_tm1 = f2d(s_dev_tm1);
                const double2 *c_mu = a.tab_smem ? s_tab + n_sls * meta[TES + mel]
                                                 : a.c_mu_tab + (size_t)n_sls * meta[TES + mel];
                const double2 *c_ka = a.tab_smem ? s_tab + n_sls * (a.ntab_mu + meta[2 * TES + mel])
                                                 : a.c_ka_tab + (size_t)n_sls * meta[2 * TES + mel];
                const float *mv = S + Ly.mv + mel * 24 * n_sls + ml;
                float *out = a.memvar + meg * 24 * n_sls + ml;
#pragma unroll
                for (int sl = 0; sl < n_sls; sl++) {
                    const double2 cm = c_mu[sl];
                    const double dev_buf = rnd32(cm.x * dsrc_t + cm.y * dsrc_tm1);
                    float nv;
                    if (mv_v < 3) {
                        const double2 ck = c_ka[sl];
                        const double tr_buf = rnd32(ck.x * src_tr_t + ck.y * s_tr_tm1);
                        nv = d2f(a.exp_w[sl] * f2d(mv[24 * sl]) + dev_buf + tr_buf);
                    } else {
                        nv = d2f(a.exp_w[sl] * f2d(mv[24 * sl]) + dev_buf);
                    }
                    out[24 * sl] = nv;
                }
            }
            a.src_dev_tm1[meg * 24 + ml] = src_dev_t;
            if (t < TES * 4) a.src_tr_tm1[(size_t)tile * TES * 4 + t] = x_src[(t >> 2) * 28 + 24 + (t & 3)];
        }
        // release the stage: generic-proxy accesses are ordered before the next TMA write
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == a.nst) { s = 0; ph ^= 1; }
    }
}

// ---- layout conversion (set-up only): host planes -> tile slabs -------------------------
// plane (25*nel) -> slab[tile][pl][TP]
__global__ void k_plane_to_slab(const float *src, float *slab, int pl, int npl, int nel, int te) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)NPT * nel) return;
    const int e = (int)(p / NPT), q = (int)(p - (size_t)e * NPT);
    slab[((size_t)(e / te) * npl + pl) * (te * NPT) + (e % te) * NPT + q] = src[p];
}
// (4, nel) -> slab[tile][pl][TE*4]
__global__ void k_cg_to_slab(const float *src, float *slab, int pl, int nel, int te) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)4 * nel) return;
    const int e = (int)(p / 4), k = (int)(p & 3);
    slab[((size_t)(e / te) * NCG + pl) * te * 4 + (e % te) * 4 + k] = src[p];
}

// inv_s_solid (25*nel) at the four coarse points -> plane G_invs of the attenuation slab
__global__ void k_invs_to_slab(const float *inv_s, float *slab, int nel, int te) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)4 * nel) return;
    const int e = (int)(p / 4), k = (int)(p & 3);
    const int i = (k & 2) ? 3 : 1, j = (k & 1) ? 3 : 1;      // coarse index = 2*(i==3) + (j==3)
    slab[((size_t)(e / te) * NCG + G_invs) * te * 4 + (e % te) * 4 + k] = inv_s[(size_t)e * NPT + i + NP * j];
}

}  // namespace axb
