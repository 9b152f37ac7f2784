// axb_fluid_tile.cuh — F_A, the fluid element kernel (included by axb_kernels.cuh).
//
// Replaces, per (sub)step: the predictor / drift of the potential chi
// (time_evol_wave.F90:357 / :597), apply_axis_mask_scal (apply_masks.f90:40-52),
// glob_fluid_stiffness_4 (stiffness_fluid.f90:139-216), add_source_fl
// (time_evol_wave.F90:1062-1076), bdry_copy2fluid (:1532-1571) and the free-surface mask
// (:383).  Same design as k_solid_tile (axb_solid_tile.cuh): tiles of TE elements streamed
// by one producer lane into a ring of shared-memory stages with 1-D TMA bulk copies,
// thread t of the NCW consumer warps owns point t of the tile, contractions over shared
// memory, results stored fully coalesced.
#pragma once

namespace axb {

constexpr int FLUID_MAX_STAGES = 8;

// stage layout (floats): [chi | dchi | ddchi0][TP], coef[npl][TP], meta ints [3][TE]
struct FluidTileArgs {
    int ntiles;
    int mode;                 // 0 Newmark, 1 symplectic drift, 2 none (op test),
                              // 3 lean Newmark (dchi holds dchi + dt/2 ddchi: chi += dt * dchi)
    int order;                // source order (monopole: no M_w term / no axis masks)
    int full;                 // 1: source, S/F coupling and masks (time loop); 0: bare stiffness
    int npl;                  // coefficient planes in the slab: M1chi, M2chi, M4chi [, M_w_fl] [, fs_mask]
    int mask_plane;           // slab index of the free-surface mask, -1: all ones
    int nst;                  // ring depth
    double dt, half_dt_sq;
    float *chi, *ddchi1;
    const float *dchi, *ddchi0;
    const float *coef;        // [tile][npl][TP]
    const int *meta;          // [tile][3][TE]: axis, S/F boundary index of the jpol=0 / jpol=4 row (1-based)
    const float *M0_w_fl;     // (5, nel_pad)
    const int *bdry_sel, *bdry_js;
    const float *bdry_matr;   // (5, nel_bdry, 2)
    int nel_bdry;
    const float *disp;        // solid displacement (already predicted)
    size_t cs_solid;
    int nelsrc;
    int ielsrc[8];
    const float *src_term;    // (5,5,8)
    const float *stf;
    int iter;
    const int *dyn;           // graph replay: iter = dyn[DYN_ITER]
    int use_mask;             // Newmark multiplies by the free-surface mask, symplectic does not
    int emask;                // mode 2 only: axis mask on the input copy and on the result (energy)
};

__host__ __device__ constexpr size_t fluid_stage_bytes(int npl) {
    return ((size_t)((3 + npl) * TP + (3 * TE + 3) / 4 * 4) * 4 + 127) / 128 * 128;
}
constexpr size_t FLUID_HDR_BYTES = 640;

__global__ void __launch_bounds__(FLUID_THREADS, 2)
k_fluid_tile(const __grid_constant__ GMat G, const __grid_constant__ FluidTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + FLUID_MAX_STAGES;
    GMat &sG = *reinterpret_cast<GMat *>(smem + 128);
    unsigned char *ring = smem + FLUID_HDR_BYTES;
    const size_t stage_bytes = fluid_stage_bytes(a.npl);
    const int off_coef = 3 * TP, off_meta = (3 + a.npl) * TP;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    if (t == 0) {
        for (int s = 0; s < a.nst; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        fence_mbar_init();
    }
    {
        const float *src = reinterpret_cast<const float *>(&G);
        float *dst = reinterpret_cast<float *>(&sG);
        for (int k = t; k < (int)(sizeof(GMat) / sizeof(float)); k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    if (warp == NCW) {                       // ---- producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t plane_b = TP * 4;
            const uint32_t bytes = plane_b * (a.mode == 0 ? 3 : (a.mode == 2 ? 1 : 2)) + a.npl * plane_b + 3 * TE * 4;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                float *S = reinterpret_cast<float *>(ring + (size_t)s * stage_bytes);
                uint64_t *bar = &full[s];
                mbar_expect_tx(bar, bytes);
                const size_t pg = (size_t)tile * TP;
                bulk_g2s(S, a.chi + pg, plane_b, bar);
                if (a.mode != 2) bulk_g2s(S + TP, a.dchi + pg, plane_b, bar);
                if (a.mode == 0) bulk_g2s(S + 2 * TP, a.ddchi0 + pg, plane_b, bar);
                bulk_g2s(S + off_coef, a.coef + (size_t)tile * a.npl * TP, a.npl * plane_b, bar);
                bulk_g2s(S + off_meta, a.meta + (size_t)tile * 3 * TE, 3 * TE * 4, bar);
                if (++s == a.nst) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    const bool pt = t < TP;
    const int el = pt ? t / NPT : 0;
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

    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        mbar_wait(&full[s], ph);
        float *S = reinterpret_cast<float *>(ring + (size_t)s * stage_bytes);
        const int *meta = reinterpret_cast<const int *>(S + off_meta);
        const bool ax = meta[el] != 0;
        const size_t pg = (size_t)tile * TP + t;
        const int eg = tile * TE + el;
        const float *Cf = S + off_coef + t;

        // ---- phase 1: predictor / drift + axis mask
        float c = 0.f;
        if (pt) {
            c = S[t];
            if (a.mode == 0)
                c = d2f(f2d(c) + a.dt * f2d(S[TP + t]) + a.half_dt_sq * f2d(S[2 * TP + t]));
            else if (a.mode == 1)
                c = d2f(f2d(c) + f2d(S[TP + t]) * a.dt);
            else if (a.mode == 3)
                c = d2f(f2d(c) + a.dt * f2d(S[TP + t]));
            if ((a.full || a.emask) && a.order != 0 && ax && i == 0) c = 0.f;     // apply_axis_mask_scal(chi)
            if (a.mode != 2) { S[t] = c; a.chi[pg] = c; }
            else if (a.emask) S[t] = c;
        }
        bar_consumers<NCT>();

        // ---- phase 2: first-stage contractions -> S1, S2 (slots of dchi / ddchi0)
        float l = 0.f;
        float g1t_row[NP], g1_row[NP], g0[NP];
        if (pt) {
            if (ax) axial_rows(sG, i, g1t_row, g1_row, g0);
            const float X1 = ax ? cxi(S + e25 + 5 * j, g1t_row) : cxi(S + e25 + 5 * j, L.g2t_row);
            const float X2 = ceta(S + e25 + i, L.g2_col);
            const float m1 = Cf[0], m2 = Cf[TP], m4 = Cf[2 * TP];
            S[TP + t] = m1 * X2 + m2 * X1;
            S[2 * TP + t] = m1 * X1 + m4 * X2;
        }
        bar_consumers<NCT>();

        // ---- phase 3: second stage, M_w / axial term, source, S/F term, masks
        if (pt) {
            const float X1 = ax ? cxi(S + TP + e25 + 5 * j, g1_row) : cxi(S + TP + e25 + 5 * j, L.g2_row);
            const float X2 = ceta(S + 2 * TP + e25 + i, L.g2t_col);
            l = X1 + X2;
            if (a.order != 0) {
                l = l + Cf[3 * TP] * c;
                if (ax) {
                    const float m0 = a.M0_w_fl[j + NP * (size_t)eg];
                    const float V1 = cxi(S + e25 + 5 * j, g0);
                    l = l + g0_i * (m0 * V1);
                }
            }
            if (a.full) {
                // add_source_fl (time_evol_wave.F90:1062-1076)
                if (a.nelsrc > 0) {
                    const float stf1 = a.stf[a.dyn ? a.dyn[DYN_ITER] : a.iter];
                    if (stf1 != 0.f)
                        for (int k = 0; k < a.nelsrc; k++)
                            if (a.ielsrc[k] - 1 == eg) l = l - a.src_term[q + NPT * k] * stf1;
                }
                // bdry_copy2fluid (time_evol_wave.F90:1532-1571)
                if (a.nel_bdry > 0 && (j == 0 || j == 4)) {
                    const int b = meta[(j == 0 ? TE : 2 * TE) + el] - 1;
                    if (b >= 0) {
                        const size_t ps = i + NP * a.bdry_js[b] + (size_t)NPT * (a.bdry_sel[b] - 1);
                        const float B1 = a.bdry_matr[i + NP * (size_t)b];
                        const float B2 = a.bdry_matr[i + NP * ((size_t)b + a.nel_bdry)];
                        const float us = a.disp[ps], uz = a.disp[ps + 2 * a.cs_solid];
                        if (a.order == 1) l = l - B1 * (us + a.disp[ps + a.cs_solid]) - B2 * uz;
                        else l = l - B1 * us - B2 * uz;
                    }
                }
                if (a.order != 0 && ax && i == 0) l = 0.f;            // apply_axis_mask_scal(ddchi1)
                if (a.use_mask && a.mask_plane >= 0) l = l * Cf[a.mask_plane * TP];
            }
            if (a.emask && a.order != 0 && ax && i == 0) l = 0.f;
            a.ddchi1[pg] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == a.nst) { s = 0; ph ^= 1; }
    }
}

}  // namespace axb
