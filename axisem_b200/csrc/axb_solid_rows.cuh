// axb_solid_rows.cuh — S_A, row-per-thread variant of the solid element kernel.
//
// Same job and same arithmetic as k_solid_tile (axb_solid_tile.cuh: predictor, masks,
// glob_stiffness_{mono,di,quad}_4, glob_anel_stiffness_*_cg4, time_step_memvars_cg4), but a
// different mapping, chosen because k_solid_tile turned out instruction-issue bound
// (profiles/r01d, r01f): here one thread owns the five points (0..4, j) of one element row.
//   * contractions along xi (first index) are done in registers, the derivative matrix
//     entries coming straight from the constant bank (kernel parameters) — no shared-memory
//     reads, no shuffles;
//   * contractions along eta exchange rows through a per-warp shared-memory buffer whose rows
//     are padded to 8 floats, so a row is read with one 128-bit and one 32-bit load;
//   * only the eta-contracted halves (u and S2*) are exchanged; S1* never leave registers.
// A warp processes 6 whole elements (30 lanes), so the exchange needs __syncwarp only: there
// is no block-level barrier in the kernel.  Two warps share a 12-element tile (1200-byte
// plane chunks keep every TMA transfer 16-byte granular).  Each pair owns one shared-memory
// stage: lane 0 of its even warp issues the tile's 1-D TMA bulk loads (UBLKCP) onto the
// pair's `full` mbarrier; results are written in place into the stage and leave through TMA
// bulk stores once both warps have arrived on the pair's `done` mbarrier.  Up to 8 pairs per
// CTA (one CTA per SM) are in different phases at any time, which is what keeps HBM busy.
#pragma once

namespace axb {

constexpr int TB = 12;                   // elements per pair tile
constexpr int TPB = TB * NPT;            // points per pair tile
constexpr int ROWS_MAX_PAIRS = 8;
constexpr int EXW = 720;                 // exchange floats per warp: 3 planes x 6 elements x 5 rows x 8

struct SolidRowsLayout {
    int u, coef, meta, cg, invs, sdev, str, mv, floats;
    size_t stage_bytes;
};
__host__ __device__ constexpr SolidRowsLayout solid_rows_layout(int order, bool anel, int n_sls) {
    SolidRowsLayout L{};
    int o = 0;
    L.u = o; o += solid_ncomp(order) * 3 * TPB;      // [comp][disp|velo|acc0][TPB]
    L.coef = o; o += solid_nplanes(order) * TPB;
    L.meta = o; o += 3 * TB;                         // ints: axis, qidx_mu, qidx_ka (36 ints = 144 B)
    L.cg = L.invs = L.sdev = L.str = L.mv = o;
    if (anel) {
        L.cg = o; o += NCG * TB * 4;
        L.invs = o; o += TPB;
        L.sdev = o; o += TB * 24;
        L.str = o; o += TB * 4;
        L.mv = o; o += TB * 24 * n_sls;
    }
    L.floats = o;
    L.stage_bytes = ((size_t)o * 4 + 2 * EXW * 4 + 127) / 128 * 128;   // + the two warps' exchange buffers
    return L;
}
constexpr size_t ROWS_HDR_BYTES = 256;   // full[8], done[8] mbarriers

// TMA store: shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- in-register contraction along xi: out[i] = sum_k M(i,k) in[k], M(i,k) = Mf[i + 5k] ----
#define AXB_XI(out, Mf, in)                                                              \
    _Pragma("unroll") for (int i_ = 0; i_ < NP; i_++) {                                  \
        float s_ = (Mf)[i_] * (in)[0];                                                   \
        s_ = s_ + (Mf)[i_ + 5] * (in)[1];                                                \
        s_ = s_ + (Mf)[i_ + 10] * (in)[2];                                               \
        s_ = s_ + (Mf)[i_ + 15] * (in)[3];                                               \
        s_ = s_ + (Mf)[i_ + 20] * (in)[4];                                               \
        (out)[i_] = s_;                                                                  \
    }

// write my row of 5 values into the padded exchange plane
__device__ __forceinline__ void ex_put(float *row, const float (&v)[NP]) {
    *reinterpret_cast<float4 *>(row) = make_float4(v[0], v[1], v[2], v[3]);
    row[4] = v[4];
}
// out[i] = sum_k rows[k][i] * c[k]     (rows of my element, 8 floats apart)
__device__ __forceinline__ void ex_eta(const float *el_rows, const float (&c)[NP], float (&out)[NP]) {
    float r[NP][NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const float4 q = *reinterpret_cast<const float4 *>(el_rows + 8 * k);
        r[k][0] = q.x; r[k][1] = q.y; r[k][2] = q.z; r[k][3] = q.w;
        r[k][4] = el_rows[8 * k + 4];
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        float s = r[0][i] * c[0];
        s = s + r[1][i] * c[1];
        s = s + r[2][i] * c[2];
        s = s + r[3][i] * c[3];
        s = s + r[4][i] * c[4];
        out[i] = s;
    }
}
// sum_k c[k] * p[k]  over shared memory (axial terms only)
__device__ __forceinline__ float dot5(const float *p, int stride, const float *c) {
    float s = p[0] * c[0];
    s = s + p[stride] * c[1];
    s = s + p[2 * stride] * c[2];
    s = s + p[3 * stride] * c[3];
    s = s + p[4 * stride] * c[4];
    return s;
}

// =======================================================================================
template <int ORDER, int NSLS>
__global__ void __launch_bounds__(ROWS_MAX_PAIRS * 64, 1)
k_solid_rows(const __grid_constant__ GMat G, const __grid_constant__ SolidTileArgs a) {
    constexpr int NC = solid_ncomp(ORDER);
    constexpr int NPL = solid_nplanes(ORDER);
    const int n_sls = NSLS >= 0 ? NSLS : a.n_sls;
    const int anel = NSLS == 0 ? 0 : a.anel;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *done = full + ROWS_MAX_PAIRS;
    const SolidRowsLayout Ly = solid_rows_layout(ORDER, NSLS != 0, n_sls);

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int pair = warp >> 1, half = warp & 1;
    const int npair = blockDim.x >> 6;
    if (t == 0) {
        for (int s = 0; s < npair; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], 2); }
        fence_mbar_init();
    }
    __syncthreads();

    float *S = reinterpret_cast<float *>(smem + ROWS_HDR_BYTES + (size_t)pair * Ly.stage_bytes);
    float *ex = S + Ly.floats + half * EXW;           // this warp's exchange / scratch buffer
    const bool anel_stiff = anel == 1 || anel == 2;
    const bool anel_update = anel >= 2;
    const bool issuer = half == 0 && lane == 0;

    // ---- this lane's row
    const bool rt = lane < 30;
    const int el6 = rt ? lane / NP : 0;               // element inside the warp's half tile
    const int j = rt ? lane - NP * el6 : 0;
    const int el = half * 6 + el6;                    // element inside the pair tile
    const int r0 = el * NPT + NP * j;                 // point (0, j) of my element inside the tile
    float g2_col[NP], g2t_col[NP];                    // G2(k,j), G2T(k,j)
#pragma unroll
    for (int k = 0; k < NP; k++) { g2_col[k] = G.G2[k + NP * j]; g2t_col[k] = G.G2T[k + NP * j]; }
    float *exrow = ex + el6 * 40 + j * 8;             // my row in exchange plane 0 (planes 240 floats apart)
    const float *exel = ex + el6 * 40;                // my element's rows

    auto issue_loads = [&](int tile) {
        const uint32_t plane_b = TPB * 4;
        uint32_t bytes = NC * plane_b * (a.mode == 0 ? 3 : (a.mode == 1 ? 2 : 1)) + 3 * TB * 4;
        if (a.do_stiff) bytes += NPL * plane_b;
        if (anel) {
            bytes += NCG * TB * 16 + TB * 96 * n_sls;
            if (anel_update) bytes += plane_b + TB * 96 + TB * 16;
        }
        uint64_t *bar = &full[pair];
        mbar_expect_tx(bar, bytes);
        const size_t pg = (size_t)tile * TPB;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const size_t off = (size_t)((ORDER == 0) ? 2 * c : c) * a.cs + pg;
            float *d = S + Ly.u + c * 3 * TPB;
            bulk_g2s(d, a.disp + off, plane_b, bar);
            if (a.mode != 2) bulk_g2s(d + TPB, a.velo + off, plane_b, bar);
            if (a.mode == 0) bulk_g2s(d + 2 * TPB, a.acc0 + off, plane_b, bar);
        }
        if (a.do_stiff) bulk_g2s(S + Ly.coef, a.coef + (size_t)tile * NPL * TPB, NPL * plane_b, bar);
        bulk_g2s(S + Ly.meta, a.meta + (size_t)tile * 3 * TB, 3 * TB * 4, bar);
        if (anel) {
            bulk_g2s(S + Ly.cg, a.cg + (size_t)tile * NCG * TB * 4, NCG * TB * 16, bar);
            bulk_g2s(S + Ly.mv, a.memvar + (size_t)tile * TB * 24 * n_sls, TB * 96 * n_sls, bar);
            if (anel_update) {
                bulk_g2s(S + Ly.invs, a.inv_s + pg, plane_b, bar);
                bulk_g2s(S + Ly.sdev, a.src_dev_tm1 + (size_t)tile * TB * 24, TB * 96, bar);
                bulk_g2s(S + Ly.str, a.src_tr_tm1 + (size_t)tile * TB * 4, TB * 16, bar);
            }
        }
    };
    auto issue_stores = [&](int tile) {
        const uint32_t plane_b = TPB * 4;
        const size_t pg = (size_t)tile * TPB;
        const bool acc_out = a.do_stiff || anel_stiff;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const size_t off = (size_t)((ORDER == 0) ? 2 * c : c) * a.cs + pg;
            float *d = S + Ly.u + c * 3 * TPB;
            if (a.mode != 2) bulk_s2g(a.disp + off, d, plane_b);
            if (acc_out) bulk_s2g(a.acc1 + off, d + TPB, plane_b);
        }
        if (anel_update) {
            bulk_s2g(a.memvar + (size_t)tile * TB * 24 * n_sls, S + Ly.mv, TB * 96 * n_sls);
            bulk_s2g(a.src_dev_tm1 + (size_t)tile * TB * 24, S + Ly.sdev, TB * 96);
            bulk_s2g(a.src_tr_tm1 + (size_t)tile * TB * 4, S + Ly.str, TB * 16);
        }
        bulk_commit();
    };

    const int tstride = gridDim.x * npair;
    int tile = blockIdx.x * npair + pair;
    if (tile >= a.ntiles) return;
    if (issuer) issue_loads(tile);
    uint32_t ph = 0;
    for (; tile < a.ntiles; tile += tstride, ph ^= 1) {
        mbar_wait(&full[pair], ph);
        if (a.dbg != 1) {
        const int *meta = reinterpret_cast<const int *>(S + Ly.meta);
        const bool ax = meta[el] != 0;
        const int eg = tile * TB + el;
        float *Ub = S + Ly.u;                         // slot (c, k) at Ub + (3c + k) * TPB
        const float *Cf = S + Ly.coef + r0;           // coefficient n at point i of my row: Cf[n * TPB + i]

        float u[3][NP], X[6][NP], l[3][NP];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < NP; i++) { u[c][i] = 0.f; l[c][i] = 0.f; X[c][i] = 0.f; X[c + 3][i] = 0.f; }
        // cg-point strain sources of my row (rows j = 1, 3 own coarse points i = 1, 3)
        float srcv[2][7];
        const bool cgrow = rt && (j == 1 || j == 3);

        if (rt) {
            // ---- predictor / drift + axis mask (time_evol_wave.F90:359-364; apply_masks.f90:55-100)
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int cc = (ORDER == 0) ? 2 * c : c;
                const float *sl = Ub + c * 3 * TPB + r0;
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    float x = sl[i];
                    if (a.mode == 0)
                        x = d2f(f2d(x) + a.dt * f2d(sl[TPB + i]) + a.half_dt_sq * f2d(sl[2 * TPB + i]));
                    else if (a.mode == 1)
                        x = d2f(f2d(x) + f2d(sl[TPB + i]) * a.dt);
                    u[cc][i] = x;
                }
            }
            if (ax && a.mode != 2) {
                if (ORDER == 0) u[0][0] = 0.f;
                else if (ORDER == 1) { u[1][0] = 0.f; u[2][0] = 0.f; }
                else { u[0][0] = 0.f; u[1][0] = 0.f; u[2][0] = 0.f; }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int cc = (ORDER == 0) ? 2 * c : c;
                if (a.mode != 2) {
                    float *sl = Ub + c * 3 * TPB + r0;
#pragma unroll
                    for (int i = 0; i < NP; i++) sl[i] = u[cc][i];
                }
                ex_put(exrow + c * 240, u[cc]);
            }
        }
        __syncwarp();
        const bool need_x = a.do_stiff || anel_update;
        if (rt && need_x) {
            // ---- first-stage contractions: X[c] = d/dxi-type, X[3+c] = d/deta-type
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int cc = (ORDER == 0) ? 2 * c : c;
                if (!ax) { AXB_XI(X[cc], G.G2T, u[cc]) } else { AXB_XI(X[cc], G.G1T, u[cc]) }
                ex_eta(exel + c * 240, g2_col, X[3 + cc]);
            }
            // ---- strain at the coarse points (compute_strain_att_el_cg4, attenuation.f90:471-535)
            if (anel_update && cgrow) {
                float Xp1[NP], Xm1[NP], Xp2[NP], Xm2[NP];
                if (ORDER == 1) {
                    float up[NP], um[NP];
#pragma unroll
                    for (int k = 0; k < NP; k++) { up[k] = u[0][k] + u[1][k]; um[k] = u[0][k] - u[1][k]; }
                    if (!ax) { AXB_XI(Xp1, G.G2T, up) AXB_XI(Xm1, G.G2T, um) }
                    else     { AXB_XI(Xp1, G.G1T, up) AXB_XI(Xm1, G.G1T, um) }
                    // eta contraction of (u1 +- u2): rows from the exchange planes 0 and 1
#pragma unroll
                    for (int i = 1; i < NP; i += 2) {
                        float sp = 0.f, sm = 0.f;
#pragma unroll
                        for (int k = 0; k < NP; k++) {
                            const float a1 = exel[8 * k + i], a2 = exel[240 + 8 * k + i];
                            const float vp = a1 + a2, vm = a1 - a2;
                            if (k == 0) { sp = vp * g2_col[0]; sm = vm * g2_col[0]; }
                            else { sp = sp + vp * g2_col[k]; sm = sm + vm * g2_col[k]; }
                        }
                        Xp2[i] = sp; Xm2[i] = sm;
                    }
                }
#pragma unroll
                for (int ii = 0; ii < 2; ii++) {
                    const int i = 1 + 2 * ii;
                    const int cgk = (i == 1 ? 0 : 2) + (j == 1 ? 0 : 1);
                    const float *cg = S + Ly.cg + el * 4 + cgk;
                    const float dzdeta = cg[G_Dze * TB * 4], dzdxi = cg[G_Dzx * TB * 4];
                    const float dsdeta = cg[G_Dse * TB * 4], dsdxi = cg[G_Dsx * TB * 4];
                    const float is = S[Ly.invs + r0 + i];
                    const float u1 = u[0][i], u2 = u[1][i], u3 = u[2][i];
                    float g1, g2, g3, g4 = 0.f, g5, g6 = 0.f;
                    const float b2s = dzdeta * X[2][i] + dzdxi * X[5][i];     // d_s u3
                    const float b2z = dsdeta * X[2][i] + dsdxi * X[5][i];     // d_z u3
                    if (ORDER == 0) {
                        const float b1s = dzdeta * X[0][i] + dzdxi * X[3][i];
                        const float b1z = dsdeta * X[0][i] + dsdxi * X[3][i];
                        g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                        g2 = is * u1;
                    } else if (ORDER == 1) {
                        const float b1s = dzdeta * Xp1[i] + dzdxi * Xp2[i];
                        const float b1z = dsdeta * Xp1[i] + dsdxi * Xp2[i];
                        g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                        g2 = 2 * (is * u2);
                        const float c1s = dzdeta * Xm1[i] + dzdxi * Xm2[i];
                        const float c1z = dsdeta * Xm1[i] + dsdxi * Xm2[i];
                        g4 = -(is * u3) - c1z;
                        g6 = -g2 - c1s;
                    } else {
                        const float b1s = dzdeta * X[0][i] + dzdxi * X[3][i];
                        const float b1z = dsdeta * X[0][i] + dsdxi * X[3][i];
                        g1 = b1s; g3 = b2z; g5 = b1z + b2s;
                        g2 = is * (u1 - 2 * u2);
                        const float c1s = dzdeta * X[1][i] + dzdxi * X[4][i];   // gradient of u2
                        const float c1z = dsdeta * X[1][i] + dsdxi * X[4][i];
                        g4 = -2 * (is * u3) - c1z;
                        g6 = is * (u2 - 2 * u1) - c1s;
                    }
                    float trace = g1 + g2;
                    trace = trace + g3;
                    const float dmu = cg[G_dmu * TB * 4], dka = cg[G_dka * TB * 4];
                    const double third = 1.0 / 3.0;
                    const double dm2 = f2d(dmu * 2);
                    srcv[ii][0] = d2f(dm2 * (f2d(g1) - f2d(trace) * third));
                    srcv[ii][1] = d2f(dm2 * (f2d(g2) - f2d(trace) * third));
                    srcv[ii][2] = d2f(dm2 * (f2d(g3) - f2d(trace) * third));
                    srcv[ii][3] = (ORDER == 0) ? 0.0f : dmu * g4;
                    srcv[ii][4] = dmu * g5;
                    srcv[ii][5] = (ORDER == 0) ? 0.0f : dmu * g6;
                    srcv[ii][6] = dka * trace;                      // src_tr_t
                }
            }
        }
        __syncwarp();                                 // everybody is done reading u from the exchange
        float S1[3][NP], S2[3][NP];
        if (rt && a.do_stiff) {
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    const float m11s = Cf[C_M11s * TPB + i], m21s = Cf[C_M21s * TPB + i], m41s = Cf[C_M41s * TPB + i];
                    const float m12s = Cf[C_M12s * TPB + i], m22s = Cf[C_M22s * TPB + i], m32s = Cf[C_M32s * TPB + i];
                    const float m42s = Cf[C_M42s * TPB + i];
                    const float m11z = Cf[C_M11z * TPB + i], m21z = Cf[C_M21z * TPB + i], m41z = Cf[C_M41z * TPB + i];
                    const float m_1 = Cf[C_M_1 * TPB + i], m_2 = Cf[C_M_2 * TPB + i], m_3 = Cf[C_M_3 * TPB + i];
                    const float m_4 = Cf[C_M_4 * TPB + i], m_w1 = Cf[C_M_w1 * TPB + i];
                    if (ORDER == 0) {
                        // stiffness_mono.f90:60-157
                        const float X1 = X[0][i], X2 = X[2][i], X3 = X[3][i], X4 = X[5][i], us = u[0][i];
                        l[0][i] = m_4 * X4 + m_2 * X3 + m_1 * X1 + m_3 * X2 + us * m_w1;
                        S1[0][i] = m11s * X3 + m21s * X1 + m12s * X4 + m22s * X2 + m_1 * us;
                        S2[0][i] = m11s * X1 + m41s * X3 + m32s * X2 + m42s * X4 + m_2 * us;
                        S1[2][i] = m11z * X4 + m21z * X2 + m32s * X3 + m22s * X1 + m_3 * us;
                        S2[2][i] = m11z * X2 + m41z * X4 + m12s * X1 + m42s * X3 + m_4 * us;
                    } else if (ORDER == 1) {
                        // stiffness_di.f90:60-256
                        const float m13s = Cf[C_M13s * TPB + i], m23s = m32s, m33s = Cf[C_M33s * TPB + i], m43s = Cf[C_M43s * TPB + i];
                        const float m_5 = Cf[C_M_5 * TPB + i], m_6 = Cf[C_M_6 * TPB + i], m_7 = Cf[C_M_7 * TPB + i], m_8 = Cf[C_M_8 * TPB + i];
                        const float m_w2 = Cf[C_M_w2 * TPB + i], m_w3 = Cf[C_M_w3 * TPB + i];
                        const float X1 = X[0][i], X2 = X[1][i], X3 = X[2][i], X4 = X[3][i], X5 = X[4][i], X6 = X[5][i];
                        const float u2 = u[1][i], u3 = u[2][i];
                        const float X7 = X1 + X2;
                        const float X8 = X4 + X5;
                        l[1][i] = m_8 * X6 + m_7 * X3 + m_1 * X1 + m_5 * X2 + m_2 * X4 + m_6 * X5 + m_w1 * u2 + m_w2 * u3;
                        l[2][i] = m_4 * X4 - m_4 * X5 + m_3 * X1 - m_3 * X2 + m_w2 * u2 + m_w3 * u3;
                        float c1 = m13s * X6, c2 = m23s * X3, c3 = m_3 * u3;
                        S1[0][i] = c1 + c2 + c3 + m11s * X4 + m21s * X1 + m12s * X5 + m22s * X2 + m_1 * u2;    // S1p
                        S1[1][i] = c1 + c2 - c3 + m11s * X5 + m21s * X2 + m12s * X4 + m22s * X1 + m_5 * u2;    // S1m
                        c1 = m33s * X3; c2 = m43s * X6; c3 = m_4 * u3;
                        S2[0][i] = c1 + c2 + c3 + m11s * X1 + m41s * X4 + m12s * X2 + m42s * X5 + m_2 * u2;    // S2p
                        S2[1][i] = c1 + c2 - c3 + m11s * X2 + m41s * X5 + m12s * X1 + m42s * X4 + m_6 * u2;    // S2m
                        S1[2][i] = m33s * X8 + m23s * X7 + m11z * X6 + m21z * X3 + m_7 * u2;                   // S1z
                        S2[2][i] = m13s * X7 + m43s * X8 + m11z * X3 + m41z * X6 + m_8 * u2;                   // S2z
                    } else {
                        // stiffness_quad.f90:238-412
                        const float m1phi = Cf[C_M1phi * TPB + i], m2phi = Cf[C_M2phi * TPB + i], m4phi = Cf[C_M4phi * TPB + i];
                        const float m_5 = Cf[C_M_5 * TPB + i], m_6 = Cf[C_M_6 * TPB + i], m_7 = Cf[C_M_7 * TPB + i], m_8 = Cf[C_M_8 * TPB + i];
                        const float m_w2 = Cf[C_M_w2 * TPB + i], m_w3 = Cf[C_M_w3 * TPB + i];
                        const float m_w4 = Cf[C_M_w4 * TPB + i], m_w5 = Cf[C_M_w5 * TPB + i];
                        const float X1 = X[0][i], X2 = X[1][i], X3 = X[2][i], X4 = X[3][i], X5 = X[4][i], X6 = X[5][i];
                        const float us = u[0][i], up = u[1][i], uz = u[2][i];
                        const float c1 = m_2 * X4, c2 = m_1 * X1, c3 = m_6 * X5, c4 = m_5 * X2, c5 = m_4 * X6, c6 = m_3 * X3;
                        l[0][i] = c1 + c2 + 2 * (c3 + c4) + c5 + c6 + m_w1 * us + m_w2 * up + 2 * m_w3 * uz;
                        l[1][i] = -2 * (c1 + c2 + c5 + c6) - (c3 + c4) + m_w2 * us + m_w4 * up - m_w3 * uz;
                        l[2][i] = 2 * (m_8 * X5 + m_7 * X2) + m_w3 * (2 * us - up) + m_w5 * uz;
                        S1[0][i] = m11s * X4 + m21s * X1 + m12s * X6 + m22s * X3 + m_1 * (us - 2 * up);       // S1s
                        S2[0][i] = m11s * X1 + m41s * X4 + m32s * X3 + m42s * X6 + m_2 * (us - 2 * up);       // S2s
                        S1[2][i] = m11z * X6 + m21z * X3 + m32s * X4 + m22s * X1 + m_3 * (us - 2 * up);       // S1z
                        S2[2][i] = m11z * X3 + m41z * X6 + m12s * X1 + m42s * X4 + m_4 * (us - 2 * up);       // S2z
                        S1[1][i] = m1phi * X5 + m2phi * X2 + m_5 * (2 * us - up) + 2 * m_7 * uz;              // S1p
                        S2[1][i] = m1phi * X2 + m4phi * X5 + m_6 * (2 * us - up) + 2 * m_8 * uz;              // S2p
                    }
                }
                // S2* rows into the exchange (u is no longer needed there)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    if (!(ORDER == 0 && c == 1)) ex_put(exrow + c * 240, S2[c]);
        }
        __syncwarp();
        if (rt) {
            if (a.do_stiff) {
                // ---- second stage: Y1 = GA . S1 (xi, in registers), Y2 = S2 . G2T (eta, exchanged)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (ORDER == 0 && c == 1) continue;
                    float Y1[NP], Y2[NP];
                    if (!ax) { AXB_XI(Y1, G.G2, S1[c]) } else { AXB_XI(Y1, G.G1, S1[c]) }
                    ex_eta(exel + c * 240, g2t_col, Y2);
#pragma unroll
                    for (int i = 0; i < NP; i++) {
                        if (ORDER == 0) {
                            if (c == 0) l[0][i] = l[0][i] + Y1[i] + Y2[i];
                            else l[2][i] = Y1[i] + Y2[i];
                        } else if (ORDER == 1) {
                            if (c == 0) l[0][i] = Y1[i] + Y2[i];
                            else l[c][i] = Y1[i] + Y2[i] + l[c][i];
                        } else {
                            l[c][i] = l[c][i] + Y1[i] + Y2[i];
                        }
                    }
                }
                // ---- axial rank-1 terms (stiffness_mono.f90:127-150, stiffness_di.f90:203-243,
                // stiffness_quad.f90:380-400); displacement of other rows from the stage
                if (ax) {
                    const size_t a0 = j + NP * (size_t)eg, b0 = NP * (size_t)eg;
                    const float *U1 = Ub + el * NPT, *U2 = Ub + 3 * TPB + el * NPT;
                    const float *U3 = Ub + ((ORDER == 0) ? 3 : 6) * TPB + el * NPT;
                    if (ORDER == 0) {
                        const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0];
                        const float V1 = dot5(U1 + NP * j, 1, G.G0);          // vxm_4(G0, us)
                        const float V2 = dot5(U3, NP, g2_col);                // vxm_4(uz0, G2)
                        float V4 = w1 * V1 + w3 * V2;
                        const float V3 = dot5(U3 + NP * j, 1, G.G0);          // vxm_4(G0, uz)
                        V4 = V4 + w2 * V3;
                        float vb[NP];
#pragma unroll
                        for (int k = 0; k < NP; k++) vb[k] = a.M0_w[2][b0 + k] * dot5(U1 + NP * k, 1, G.G0);
                        const float V1b = dot5(vb, 1, g2t_col);               // vxm_4(V2, G2T)
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            float X2a = G.G0[i] * (w2 * V1);
                            if (i == 0) X2a = X2a + V1b;
                            l[0][i] = l[0][i] + G.G0[i] * V4;
                            l[2][i] = X2a + l[2][i];
                        }
                    } else if (ORDER == 1) {
                        const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0], w4 = a.M0_w[3][a0];
                        const float w6 = a.M0_w[5][a0], w7 = a.M0_w[6][a0], w8 = a.M0_w[7][a0], w9 = a.M0_w[8][a0];
                        const float w10 = a.M0_w[9][a0];
                        const float V1 = dot5(U1 + NP * j, 1, G.G0), V2 = dot5(U2 + NP * j, 1, G.G0);
                        const float V3 = dot5(U3 + NP * j, 1, G.G0);
                        const float V4 = dot5(U1, NP, g2_col);                // vxm_4(u10, G2)
                        float vb[NP];
#pragma unroll
                        for (int k = 0; k < NP; k++) {
                            const float k2 = a.M0_w[1][b0 + k], k6 = a.M0_w[5][b0 + k];
                            const float k4 = a.M0_w[3][b0 + k], k8 = a.M0_w[7][b0 + k];
                            vb[k] = (k2 + k6) * dot5(U2 + NP * k, 1, G.G0) + (k4 + k8) * dot5(U3 + NP * k, 1, G.G0);
                        }
                        const float V1b = dot5(vb, 1, g2t_col);
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            float s1p = G.G0[i] * (w1 * V2 + w3 * V3);
                            const float s1m = G.G0[i] * (w1 * V1 + (w2 + w6) * V4 + w9 * V2 + w10 * V3);
                            const float s1z = G.G0[i] * (w3 * V1 + (w4 + w8) * V4 + w7 * V3 + w10 * V2);
                            if (i == 0) s1p = s1p + V1b;
                            l[0][i] = l[0][i] + s1p;
                            l[1][i] = l[1][i] + s1m;
                            l[2][i] = l[2][i] + s1z;
                        }
                    } else {
                        const float w1 = a.M0_w[0][a0], w2 = a.M0_w[1][a0], w3 = a.M0_w[2][a0];
                        const float w4 = a.M0_w[3][a0], w5 = a.M0_w[4][a0], w6 = a.M0_w[5][a0];
                        const float V1 = dot5(U1 + NP * j, 1, G.G0), V2 = dot5(U2 + NP * j, 1, G.G0);
                        const float V3 = dot5(U3 + NP * j, 1, G.G0);
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            l[0][i] = l[0][i] + G.G0[i] * (w1 * V1 + w2 * V2 + w3 * V3);
                            l[1][i] = l[1][i] + G.G0[i] * (w2 * V1 + w4 * V2 + w5 * V3);
                            l[2][i] = l[2][i] + G.G0[i] * (w3 * V1 + w5 * V2 + w6 * V3);
                        }
                    }
                }
            } else {
                // anelastic-only operator test: start from the stored acc1
                const size_t pg = (size_t)tile * TPB + r0;
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    l[0][i] = a.acc1[pg + i];
                    if (ORDER != 0) l[1][i] = a.acc1[pg + a.cs + i];
                    l[2][i] = a.acc1[pg + 2 * a.cs + i];
                }
            }
        }
        // ---- anelastic part; the exchange buffer becomes scratch: rsum[6][24] | anS[6][36] | src[6][28]
        float *x_rsum = ex, *x_anS = ex + 144, *x_src = ex + 360;
        if (anel) {
            __syncwarp();
            // r(v)(k) = sum over the standard linear solids (stiffness_mono.f90:545-549)
            for (int it = lane; it < 144; it += 32) {
                const int e6 = it / 24, ml = it - 24 * e6, v = ml >> 2;
                float rsum = 0.0f;
                if (!(ORDER == 0 && (v == 3 || v == 5))) {
                    const float *mv = S + Ly.mv + (half * 6 + e6) * 24 * n_sls + ml;
#pragma unroll
                    for (int sl = 0; sl < n_sls; sl++) rsum = rsum + mv[24 * sl];
                }
                x_rsum[it] = rsum;
            }
            if (anel_update && cgrow) {
#pragma unroll
                for (int ii = 0; ii < 2; ii++) {
                    const int cgk = (ii == 0 ? 0 : 2) + (j == 1 ? 0 : 1);
                    float *src = x_src + el6 * 28 + cgk;
#pragma unroll
                    for (int v = 0; v < 7; v++) src[4 * v] = srcv[ii][v];
                }
            }
            __syncwarp();
        }
        if (anel_stiff) {
            if (lane < 24) {
                // S terms at the four coarse points (glob_anel_stiffness_*_cg4)
                const int ce = lane >> 2, ck = lane & 3;
                const float *cg = S + Ly.cg + (half * 6 + ce) * 4 + ck;
                const float yl = cg[G_Y * TB * 4];
                const float vse = cg[G_Vse * TB * 4], vsx = cg[G_Vsx * TB * 4];
                const float vze = cg[G_Vze * TB * 4], vzx = cg[G_Vzx * TB * 4];
                const float *r = x_rsum + ce * 24 + ck;
                const float r1 = r[0], r2 = r[4], r3 = r[8], r4 = r[12], r5 = r[16], r6 = r[20];
                float *Sa = x_anS + ce * 36 + ck;
                if (ORDER == 0) {
                    Sa[0] = vze * r1 + vse * r5;
                    Sa[4] = vzx * r1 + vsx * r5;
                    Sa[16] = vze * r5 + vse * r3;
                    Sa[20] = vzx * r5 + vsx * r3;
                    Sa[24] = yl * r2;
                } else if (ORDER == 1) {
                    Sa[0] = vze * (r1 - r6) + vse * (r5 - r4);
                    Sa[4] = vzx * (r1 - r6) + vsx * (r5 - r4);
                    Sa[8] = vze * (r1 + r6) + vse * (r5 + r4);
                    Sa[12] = vzx * (r1 + r6) + vsx * (r5 + r4);
                    Sa[16] = vze * r5 + vse * r3;
                    Sa[20] = vzx * r5 + vsx * r3;
                    Sa[24] = 2 * yl * (r2 - r6);
                    Sa[28] = yl * r4;
                } else {
                    Sa[0] = vze * r1 + vse * r5;
                    Sa[4] = vzx * r1 + vsx * r5;
                    Sa[8] = vze * r6 + vse * r4;
                    Sa[12] = vzx * r6 + vsx * r4;
                    Sa[16] = vze * r5 + vse * r3;
                    Sa[20] = vzx * r5 + vsx * r3;
                    Sa[24] = yl * (r2 - 2 * r6);
                    Sa[28] = yl * (r6 - 2 * r2);
                    Sa[32] = 2 * yl * r4;
                }
            }
            __syncwarp();
            if (rt) {
                const float *Sa = x_anS + el6 * 36;
                const bool colb = (j == 1) || (j == 3);
                const int kb = (j == 1) ? 0 : 1;
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    const bool rowa = (i == 1) || (i == 3);
                    const int ka = (i == 1) ? 0 : 2;
                    const float ga1 = ax ? G.G1[i + 5] : G.G2[i + 5];       // GA(i,1), GA(i,3)
                    const float ga3 = ax ? G.G1[i + 15] : G.G2[i + 15];
                    float Xb[3], Xa[3];
#pragma unroll
                    for (int m = 0; m < 3; m++) {
                        const float *T1 = Sa + 8 * m, *T2 = Sa + 8 * m + 4;
                        Xb[m] = colb ? (ga1 * T1[kb] + ga3 * T1[kb + 2]) : 0.0f;
                        Xa[m] = rowa ? (T2[ka] * g2t_col[1] + T2[ka + 1] * g2t_col[3]) : 0.0f;
                    }
                    const bool cg2 = rowa && colb;
                    const int cgk = (i == 1 ? 0 : 2) + (j == 1 ? 0 : 1);
                    if (ORDER == 0) {
                        float ls = Xb[0] + Xa[0];
                        const float lz = Xb[2] + Xa[2];
                        if (cg2) ls = ls + Sa[24 + cgk];
                        l[0][i] = l[0][i] - ls;
                        l[2][i] = l[2][i] - lz;
                    } else if (ORDER == 1) {
                        const float lp = Xb[0] + Xa[0];
                        float lm = Xb[1] + Xa[1];
                        float lz = Xb[2] + Xa[2];
                        if (cg2) { lm = lm + Sa[24 + cgk]; lz = lz - Sa[28 + cgk]; }
                        l[0][i] = l[0][i] - lp; l[1][i] = l[1][i] - lm; l[2][i] = l[2][i] - lz;
                    } else {
                        float ls = Xb[0] + Xa[0];
                        float lp = -Xb[1] - Xa[1];
                        float lz = Xb[2] + Xa[2];
                        if (cg2) { ls = ls + Sa[24 + cgk]; lp = lp + Sa[28 + cgk]; lz = lz - Sa[32 + cgk]; }
                        l[0][i] = l[0][i] - ls; l[1][i] = l[1][i] - lp; l[2][i] = l[2][i] - lz;
                    }
                }
            }
        }
        if (rt && (a.do_stiff || anel_stiff)) {
            // apply_axis_mask_*(acc1) (time_evol_wave.F90:438-447), then acc1 in place of velo
            if (ax && a.mode != 2) {
                if (ORDER == 0) l[0][0] = 0.f;
                else if (ORDER == 1) { l[1][0] = 0.f; l[2][0] = 0.f; }
                else { l[0][0] = 0.f; l[1][0] = 0.f; l[2][0] = 0.f; }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int cc = (ORDER == 0) ? 2 * c : c;
                float *sl = Ub + (c * 3 + 1) * TPB + r0;
#pragma unroll
                for (int i = 0; i < NP; i++) sl[i] = l[cc][i];
            }
        }
        // ---- memory-variable update in place (time_step_memvars_cg4, attenuation.f90:136-200)
        if (anel_update) {
            for (int it = lane; it < 144; it += 32) {
                const int e6 = it / 24, ml = it - 24 * e6, v = ml >> 2, k = ml & 3;
                const int me = half * 6 + e6;
                const float src_dev_t = x_src[e6 * 28 + ml];
                float *sdev = S + Ly.sdev + me * 24 + ml;
                const float s_dev_tm1 = *sdev;
                if (!(ORDER == 0 && (v == 3 || v == 5))) {
                    const double src_tr_t = f2d(x_src[e6 * 28 + 24 + k]);
                    const double s_tr_tm1 = f2d(S[Ly.str + me * 4 + k]);
                    const double dsrc_t = f2d(src_dev_t), dsrc_tm1 = f2d(s_dev_tm1);
                    const double2 *c_mu = a.c_mu_tab + (size_t)n_sls * meta[TB + me];
                    const double2 *c_ka = a.c_ka_tab + (size_t)n_sls * meta[2 * TB + me];
                    float *mv = S + Ly.mv + me * 24 * n_sls + ml;
#pragma unroll
                    for (int sl = 0; sl < n_sls; sl++) {
                        const double2 cm = c_mu[sl];
                        const double dev_buf = rnd32(cm.x * dsrc_t + cm.y * dsrc_tm1);
                        float nv;
                        if (v < 3) {
                            const double2 ck = c_ka[sl];
                            const double tr_buf = rnd32(ck.x * src_tr_t + ck.y * s_tr_tm1);
                            nv = d2f(a.exp_w[sl] * f2d(mv[24 * sl]) + dev_buf + tr_buf);
                        } else {
                            nv = d2f(a.exp_w[sl] * f2d(mv[24 * sl]) + dev_buf);
                        }
                        mv[24 * sl] = nv;
                    }
                }
                *sdev = src_dev_t;
            }
            __syncwarp();                             // src_tr_tm1 of the previous step has been consumed
            if (lane < 24) S[Ly.str + (half * 6 + (lane >> 2)) * 4 + (lane & 3)] = x_src[(lane >> 2) * 28 + 24 + (lane & 3)];
        }
        }  // dbg
        // ---- hand the tile to the TMA stores, then fetch the pair's next tile
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[pair]);
        if (half == 0) {
            if (lane == 0) {
                mbar_wait(&done[pair], ph);
                if (a.dbg != 1) issue_stores(tile);
                bulk_wait_read();
                if (tile + tstride < a.ntiles) issue_loads(tile + tstride);
            }
            __syncwarp();
        }
    }
    if (issuer) bulk_wait_all();
}

}  // namespace axb
