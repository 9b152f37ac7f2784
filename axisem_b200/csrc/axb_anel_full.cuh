// axb_anel_full.cuh — attenuation with memory variables at all 25 GLL points of an element
// (COARSE_GRAINED false in inparam_advanced).
//
// Replaces glob_anel_stiffness_{mono,di,quad}_4 (stiffness_mono.f90:409-505,
// stiffness_di.f90:604-721, stiffness_quad.f90:555-666), time_step_memvars_4
// (attenuation.f90:210-334), compute_strain_att_el_4 (attenuation.f90:542-606) and the
// pointwise operators below it: axisym_gradient_solid_el_4, f_over_s_solid_el_4 with its
// L'Hopital branch on the axis (pointwise_derivatives.f90:252-286, 104-126, 419-448).
//
// This option moves 25/4 times the memory-variable bytes of the coarse-grained default
// (3 kB read + 3 kB written per element and step for 5 SLS), so it runs as its own
// HBM-bound kernel behind S_A instead of riding in S_A's ring: one thread per GLL point,
// TEA elements per CTA, every global access a run of 25 consecutive floats per element,
// the displacement / r-sum / S planes exchanged through shared memory for the 5x5
// contractions.  S_A has already written the predicted, masked displacement and the elastic
// K u to HBM when this kernel starts.
//
// Arithmetic mirrors oracle/axisem_oracle.c (glob_anel_stiffness_4, compute_strain_att_el_4,
// time_step_memvars); bit-identical with -fmad=false.
#pragma once

namespace axb {

#ifndef AXB_ANEL_MINB
#define AXB_ANEL_MINB 4                   // resident CTAs per SM the register budget is set for
#endif
#ifndef AXB_ANEL_TEA
#define AXB_ANEL_TEA 5                    // 125 points on 128 lanes; 8 (200 on 224) measured 9 % slower
#endif
constexpr int TEA = AXB_ANEL_TEA;         // elements per CTA
constexpr int TPA = TEA * NPT;            // points per CTA
constexpr int ANEL_THREADS = (TPA + 31) / 32 * 32;

struct AnelFullArgs {
    int nel, n_sls;
    int do_stiff, do_update, mask;        // mask: re-apply the axis mask to acc1
    const float *disp; float *acc1; size_t cs;
    const int *axis;                      // (nel)
    const float *Y, *Vse, *Vsx, *Vze, *Vzx;          // (25 nel)
    const float *Y0, *V0se, *V0sx, *V0ze, *V0zx;     // (5 nel)
    const float *Dse, *Dze, *Dsx, *Dzx, *inv_s;      // (25 nel)
    const float *dmu, *dka;                          // (25 nel)
    const int *qidx_mu, *qidx_ka;                    // (nel) rows of the a_j tables
    const double2 *c_mu_tab, *c_ka_tab;
    double exp_w[8];
    float *memvar;                        // (25, 6, n_sls, nel)
    float *src_dev_tm1;                   // (25, 6, nel)
    float *src_tr_tm1;                    // (25, nel)
};

// NSLS > 0: the number of SLS at compile time — the 6 x NSLS memory variables of the point are
// loaded once, all loads in flight together, and stay in registers from the K term to the update
// (the run-time version read them twice, one dependent round trip per SLS: ncu showed 18 warps
// per issue waiting on the long scoreboard at 26 % issue utilisation).  NSLS = 0: any n_sls.
template <int ORDER, int NSLS>
__global__ void __launch_bounds__(ANEL_THREADS, NSLS > 0 ? AXB_ANEL_MINB : 1)
k_anel_full(const __grid_constant__ GMat G, const __grid_constant__ AnelFullArgs a) {
    __shared__ GMat sG;
    __shared__ float sU[3][TPA];          // u1, u2, u3
    __shared__ float sT[2][TPA];          // dipole: u1+u2, u1-u2 ; quadrupole: u1-2u2, u2-2u1
    __shared__ float sR[6][TPA];          // r(v) = sum over the SLS of the memory variables
    __shared__ float sS[6][TPA];          // S1a, S2a, S1b, S2b, S1z, S2z
    stage_g(G, sG);
    const int t = threadIdx.x;
    const int el = t / NPT, q = t - el * NPT;
    const int i = q % NP, j = q / NP;
    const int e = blockIdx.x * TEA + el;
    const bool pt = t < TPA && e < a.nel;
    const int e25 = el * NPT;
    const size_t p = (size_t)NPT * e + q;
    const bool ax = pt && a.axis[e] != 0;
    constexpr bool CT = NSLS > 0;
    const int n_sls = CT ? NSLS : a.n_sls;

    float g2t_row[NP], g2_col[NP], g2_row[NP], g2t_col[NP], gat_row[NP], ga_row[NP], g0[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        g2t_row[k] = sG.G2T[i + NP * k];
        g2_col[k] = sG.G2[k + NP * j];
        g2_row[k] = sG.G2[i + NP * k];
        g2t_col[k] = sG.G2T[k + NP * j];
        gat_row[k] = ax ? sG.G1T[i + NP * k] : g2t_row[k];    // first stage, xi
        ga_row[k] = ax ? sG.G1[i + NP * k] : g2_row[k];       // GA(i,k), second stage
        g0[k] = sG.G0[k];
    }
    const float g0_i = sG.G0[i];

    float u1 = 0.f, u2 = 0.f, u3 = 0.f;
    if (pt) {
        u1 = a.disp[p];
        if (ORDER != 0) u2 = a.disp[p + a.cs];
        u3 = a.disp[p + 2 * a.cs];
    }
    if (t < TPA) {
        sU[0][t] = u1; sU[1][t] = u2; sU[2][t] = u3;
        if (ORDER == 1) { sT[0][t] = u1 + u2; sT[1][t] = u1 - u2; }
        if (ORDER == 2) { sT[0][t] = u1 - 2 * u2; sT[1][t] = u2 - 2 * u1; }
    }
    const float *mv = a.memvar + (size_t)NPT * 6 * n_sls * e + q;    // mv[25 * (v + 6 j)]
    float mreg[CT ? NSLS : 1][6];
    // everything the update needs from HBM, requested before the first barrier
    float dzdeta = 0.f, dzdxi = 0.f, dsdeta = 0.f, dsdxi = 0.f, is = 0.f, dmu = 0.f, dka = 0.f, tr_old = 0.f;
    float dev_old[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (CT && pt) {
#pragma unroll
        for (int s = 0; s < (CT ? NSLS : 1); s++)
#pragma unroll
            for (int v = 0; v < 6; v++)
                mreg[s][v] = (ORDER == 0 && (v == 3 || v == 5)) ? 0.f : mv[NPT * (v + 6 * s)];
    }
    if (a.do_update && pt) {
        dzdeta = a.Dze[p]; dzdxi = a.Dzx[p]; dsdeta = a.Dse[p]; dsdxi = a.Dsx[p];
        is = a.inv_s[p]; dmu = a.dmu[p]; dka = a.dka[p];
        tr_old = a.src_tr_tm1[p];
#pragma unroll
        for (int v = 0; v < 6; v++) dev_old[v] = a.src_dev_tm1[(size_t)NPT * 6 * e + q + NPT * v];
    }

    // ---------------------------------------------------------------- anelastic K term ----
    float r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f, r5 = 0.f, r6 = 0.f;
    float yl = 0.f;
    if (a.do_stiff) {
        if (pt) {
            if (CT) {
#pragma unroll
                for (int s = 0; s < (CT ? NSLS : 1); s++) {
                    r1 = r1 + mreg[s][0]; r2 = r2 + mreg[s][1]; r3 = r3 + mreg[s][2];
                    if (ORDER != 0) r4 = r4 + mreg[s][3];
                    r5 = r5 + mreg[s][4];
                    if (ORDER != 0) r6 = r6 + mreg[s][5];
                }
            } else {
                for (int s = 0; s < n_sls; s++) {
                    const float *m = mv + NPT * 6 * s;
                    r1 = r1 + m[0]; r2 = r2 + m[NPT]; r3 = r3 + m[2 * NPT];
                    if (ORDER != 0) r4 = r4 + m[3 * NPT];
                    r5 = r5 + m[4 * NPT];
                    if (ORDER != 0) r6 = r6 + m[5 * NPT];
                }
            }
            const float vse = a.Vse[p], vsx = a.Vsx[p], vze = a.Vze[p], vzx = a.Vzx[p];
            yl = a.Y[p];
            if (ORDER == 0) {
                sS[0][t] = vze * r1 + vse * r5;
                sS[1][t] = vzx * r1 + vsx * r5;
            } else if (ORDER == 1) {
                sS[0][t] = vze * (r1 - r6) + vse * (r5 - r4);
                sS[1][t] = vzx * (r1 - r6) + vsx * (r5 - r4);
                sS[2][t] = vze * (r1 + r6) + vse * (r5 + r4);
                sS[3][t] = vzx * (r1 + r6) + vsx * (r5 + r4);
            } else {
                sS[0][t] = vze * r1 + vse * r5;
                sS[1][t] = vzx * r1 + vsx * r5;
                sS[2][t] = vze * r6 + vse * r4;
                sS[3][t] = vzx * r6 + vsx * r4;
            }
            sS[4][t] = vze * r5 + vse * r3;
            sS[5][t] = vzx * r5 + vsx * r3;
            sR[0][t] = r1; sR[1][t] = r2; sR[2][t] = r3; sR[3][t] = r4; sR[4][t] = r5; sR[5][t] = r6;
        }
    }
    __syncthreads();
    if (a.do_stiff && pt) {
        const int xi0 = e25 + 5 * j, et0 = e25 + i;
        const float X1 = cxi(&sS[0][xi0], ga_row), X2 = ceta(&sS[1][et0], g2t_col);
        const float X5 = cxi(&sS[4][xi0], ga_row), X6 = ceta(&sS[5][et0], g2t_col);
        float X3 = 0.f, X4 = 0.f;
        if (ORDER != 0) { X3 = cxi(&sS[2][xi0], ga_row); X4 = ceta(&sS[3][et0], g2t_col); }
        float la, lb = 0.f, lz;
        // r at (0, j) and (0, k) of this element
        const int a0 = e25 + 5 * j;
        const size_t v0 = (size_t)NP * e;
        if (ORDER == 0) {
            la = X1 + X2 + yl * r2;
            lz = X5 + X6;
            if (ax) {
                const float v0ze = a.V0ze[v0 + j], v0se = a.V0se[v0 + j], y0 = a.Y0[v0 + j];
                const float V1 = v0ze * sR[0][a0] + v0se * sR[4][a0] + y0 * sR[1][a0];
                la = la + g0_i * V1;
                const float V2 = v0ze * sR[4][a0] + v0se * sR[2][a0];
                lz = lz + g0_i * V2;
                if (i == 0) {
                    float V3[NP];
#pragma unroll
                    for (int k = 0; k < NP; k++)
                        V3[k] = a.V0zx[v0 + k] * sR[4][e25 + 5 * k] + a.V0sx[v0 + k] * sR[2][e25 + 5 * k];
                    lz = lz + cxi(V3, g2t_col);
                }
            }
        } else if (ORDER == 1) {
            la = X1 + X2;
            lb = X3 + X4 + 2 * yl * (r2 - r6);
            lz = X5 + X6 - yl * r4;
            if (ax) {
                const float v0ze = a.V0ze[v0 + j], v0se = a.V0se[v0 + j], y0 = a.Y0[v0 + j];
                const float v0zx = a.V0zx[v0 + j], v0sx = a.V0sx[v0 + j];
                const float q1 = sR[0][a0], q2 = sR[1][a0], q4 = sR[3][a0], q5 = sR[4][a0], q6 = sR[5][a0];
                const float V1 = v0ze * (q1 - q6) + v0se * (q5 - q4);
                const float V2 = v0zx * (q1 - q6) + v0sx * (q5 - q4);
                la = la + g0_i * V1;
                if (i == 0) {
                    float Vk[NP];
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        const int ak = e25 + 5 * k;
                        Vk[k] = a.V0zx[v0 + k] * (sR[0][ak] - sR[5][ak]) + a.V0sx[v0 + k] * (sR[4][ak] - sR[3][ak]);
                    }
                    la = la + cxi(Vk, g2t_col);
                }
                const float V1b = v0ze * (q1 + q6) + v0se * (q5 + q4) + y0 * 2 * (q2 - q6);
                lb = lb + g0_i * V1b;
                // the reference adds outerprod(G0, V2) here (stiffness_di.f90:710-711)
                lz = lz + g0_i * V2;
            }
        } else {
            la = X1 + X2 + yl * (r2 - 2 * r6);
            lb = -X3 - X4 + yl * (r6 - 2 * r2);
            lz = X5 + X6 - 2 * yl * r4;
            if (ax) {
                const float v0ze = a.V0ze[v0 + j], v0se = a.V0se[v0 + j], y0 = a.Y0[v0 + j];
                const float q1 = sR[0][a0], q2 = sR[1][a0], q3 = sR[2][a0], q6 = sR[5][a0];
                la = la + g0_i * (v0ze * q1 + y0 * (q2 - 2 * q6));
                lb = lb + g0_i * (-v0ze * q6 + y0 * (q6 - 2 * q2));
                lz = lz + g0_i * (v0se * q3);
            }
        }
        float c1 = a.acc1[p] - la;
        float c2 = (ORDER != 0) ? a.acc1[p + a.cs] - lb : 0.f;
        float c3 = a.acc1[p + 2 * a.cs] - lz;
        if (a.mask && ax && i == 0) {
            if (ORDER == 0) c1 = 0.f;
            else if (ORDER == 1) { c2 = 0.f; c3 = 0.f; }
            else { c1 = 0.f; c2 = 0.f; c3 = 0.f; }
        }
        a.acc1[p] = c1;
        if (ORDER != 0) a.acc1[p + a.cs] = c2;
        a.acc1[p + 2 * a.cs] = c3;
    }
    if (!a.do_update || !pt) return;

    // ---------------------------------------------- strain (compute_strain_att_el_4) ----
    const int xi0 = e25 + 5 * j, et0 = e25 + i;
    // axisym_gradient_solid_el_4 of the plane f
#define AXB_GRAD(f, ds, dz)                                                               \
    {                                                                                     \
        const float m1_ = cxi(&(f)[xi0], gat_row), m2_ = ceta(&(f)[et0], g2_col);         \
        ds = dzdeta * m1_ + dzdxi * m2_;                                                  \
        dz = dsdeta * m1_ + dsdxi * m2_;                                                  \
    }
    const bool lhop = ax && i == 0;       // f/s -> d_s f on the axis
    float g1, g2, g3, g4 = 0.f, g5, g6 = 0.f;
    float b1s, b1z, b2s, b2z;
    AXB_GRAD(sU[2], b2s, b2z);
    if (ORDER == 0) {
        AXB_GRAD(sU[0], b1s, b1z);
        g2 = lhop ? b1s : is * u1;
    } else if (ORDER == 1) {
        AXB_GRAD(sT[0], b1s, b1z);
        float fs = is * u2;
        if (lhop) { float ds, dz; AXB_GRAD(sU[1], ds, dz); (void)dz; fs = ds; }
        g2 = 2 * fs;
        float c1s, c1z;
        AXB_GRAD(sT[1], c1s, c1z);
        const float fs3 = lhop ? b2s : is * u3;
        g4 = -fs3 - c1z;
        g6 = -g2 - c1s;
    } else {
        AXB_GRAD(sU[0], b1s, b1z);
        float fs = is * sT[0][t];
        if (lhop) { float ds, dz; AXB_GRAD(sT[0], ds, dz); (void)dz; fs = ds; }
        g2 = fs;
        float c1s, c1z;
        AXB_GRAD(sU[1], c1s, c1z);
        const float fs3 = lhop ? b2s : is * u3;
        float fs2 = is * sT[1][t];
        if (lhop) { float ds, dz; AXB_GRAD(sT[1], ds, dz); (void)dz; fs2 = ds; }
        g4 = -2 * fs3 - c1z;
        g6 = fs2 - c1s;
    }
#undef AXB_GRAD
    g1 = b1s; g3 = b2z; g5 = b1z + b2s;

    // ------------------------------------ memory variables (time_step_memvars_4) ----
    float trace = g1 + g2;
    trace = trace + g3;
    const double third = 1.0 / 3.0;
    const double dm2 = f2d(dmu * 2);
    float src[6];
    src[0] = d2f(dm2 * (f2d(g1) - f2d(trace) * third));
    src[1] = d2f(dm2 * (f2d(g2) - f2d(trace) * third));
    src[2] = d2f(dm2 * (f2d(g3) - f2d(trace) * third));
    src[3] = (ORDER == 0) ? 0.0f : dmu * g4;
    src[4] = dmu * g5;
    src[5] = (ORDER == 0) ? 0.0f : dmu * g6;
    const float src_tr = dka * trace;
    float *dev_tm1 = a.src_dev_tm1 + (size_t)NPT * 6 * e + q;
    float *tr_tm1 = a.src_tr_tm1 + p;
    const double d_tr_t = f2d(src_tr), d_tr_tm1 = f2d(tr_old);
    double d_t[6], d_tm1[6];
#pragma unroll
    for (int v = 0; v < 6; v++) { d_t[v] = f2d(src[v]); d_tm1[v] = f2d(dev_old[v]); }
    const double2 *c_mu = a.c_mu_tab + (size_t)n_sls * a.qidx_mu[e];
    const double2 *c_ka = a.c_ka_tab + (size_t)n_sls * a.qidx_ka[e];
    float *mvw = a.memvar + (size_t)NPT * 6 * n_sls * e + q;
#pragma unroll
    for (int s = 0; s < (CT ? NSLS : n_sls); s++) {
        const double2 cm = c_mu[s], ck = c_ka[s];
        const double ew = a.exp_w[s];
        const double tr_buf = rnd32(ck.x * d_tr_t + ck.y * d_tr_tm1);
        float *m = mvw + NPT * 6 * s;
#pragma unroll
        for (int v = 0; v < 6; v++) {
            if (ORDER == 0 && (v == 3 || v == 5)) continue;
            const double dev_buf = rnd32(cm.x * d_t[v] + cm.y * d_tm1[v]);
            const float old = CT ? mreg[CT ? s : 0][v] : m[NPT * v];
            if (v < 3) m[NPT * v] = d2f(ew * f2d(old) + dev_buf + tr_buf);
            else m[NPT * v] = d2f(ew * f2d(old) + dev_buf);
        }
    }
    *tr_tm1 = src_tr;
#pragma unroll
    for (int v = 0; v < 6; v++) dev_tm1[NPT * v] = src[v];
}

}  // namespace axb
