// axb_kernels.cuh — sm_100a kernels of the AxiSEM time loop.
//
// Mapping (round 1): one warp per spectral element, lane q = ipol + 5*jpol (< 25) owns
// one GLL point; the 5x5 contractions of unrolled_loops.f90:164-188 are done with warp
// shuffles (no shared-memory round trip, no block barrier); all HBM loads/stores are
// 100-byte contiguous runs per (plane, element), consecutive warps touch consecutive
// elements.  Pointwise kernels (correctors with assembly) are one thread per GLL point.
//
// Arithmetic mirrors oracle/axisem_oracle.c statement by statement (same association,
// real(8) promotion where the Fortran promotes).  Built with -fmad=false the results are
// bit-identical to the oracle (libaxisem_b200_strict.so, used by the parity tests); the
// product build keeps FMA contraction on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace axb {

constexpr int NP = 5;
constexpr int NPT = 25;
constexpr unsigned FULL = 0xffffffffu;

struct GMat {            // Fortran order: G(i,j) at [i + 5*j]
    float G0[NP];
    float G1[NPT], G1T[NPT], G2[NPT], G2T[NPT];
};

// ---------------------------------------------------------------------------------------
struct SolidPlanes {
    const float *M11s, *M21s, *M41s, *M12s, *M22s, *M32s, *M42s, *M11z, *M21z, *M41z;
    const float *M13s, *M33s, *M43s, *M1phi, *M2phi, *M4phi;
    const float *M_1, *M_2, *M_3, *M_4, *M_5, *M_6, *M_7, *M_8;
    const float *M_w1, *M_w2, *M_w3, *M_w4, *M_w5;
    const float *M0_w1, *M0_w2, *M0_w3, *M0_w4, *M0_w5, *M0_w6, *M0_w7, *M0_w8, *M0_w9, *M0_w10;
};

struct AttCg {            // coarse-grained attenuation inputs (device pointers)
    int n_sls;
    const float *Ycg, *Vse, *Vsx, *Vze, *Vzx;      // (4,nel)
    const float *Dse, *Dze, *Dsx, *Dzx;            // (4,nel)
    const float *dmu, *dka;                        // (4,nel)
    const float *inv_s;                            // (5,5,nel)
    const int *qidx_mu, *qidx_ka;                  // (nel) index into a_j tables
    const double *a_mu_tab, *a_ka_tab;             // (ntab, n_sls)
    const double *exp_w, *ts_t, *ts_tm1;           // (n_sls)
    float *memvar;                                 // (4,6,n_sls,nel)
    float *src_dev_tm1;                            // (4,6,nel)
    float *src_tr_tm1;                             // (4,nel)
};

struct SolidStepArgs {
    int nel;
    int mode;                 // 0: Newmark predictor, 1: symplectic drift, 2: none (op test)
    double dt, half_dt_sq;    // Newmark: dt, dt^2/2 ; symplectic: coefd in dt
    float *disp, *velo, *acc0, *acc1;
    const int *axis;          // (nel) 0/1
    int anel;                 // 0 none, 1 cg4 stiffness only, 2 cg4 stiffness + memvar update,
                              // 3 memvar update only (time_step_memvars as its own pass)
    int do_stiff;             // 0: skip elastic stiffness (anel-only op test keeps acc1)
};

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float shfl(float v, int src) { return __shfl_sync(FULL, v, src); }

// sum_k coef[k] * val(k, j)   (first index contracted; val(k,j) lives on lane 5*j + k)
__device__ __forceinline__ float contract_xi(float val, const float (&coef)[NP], int j5) {
    float s = coef[0] * shfl(val, j5 + 0);
    s = s + coef[1] * shfl(val, j5 + 1);
    s = s + coef[2] * shfl(val, j5 + 2);
    s = s + coef[3] * shfl(val, j5 + 3);
    s = s + coef[4] * shfl(val, j5 + 4);
    return s;
}
// sum_k val(i, k) * coef[k]   (second index contracted; val(i,k) lives on lane i + 5*k)
__device__ __forceinline__ float contract_eta(float val, const float (&coef)[NP], int i) {
    float s = shfl(val, i + 0) * coef[0];
    s = s + shfl(val, i + 5) * coef[1];
    s = s + shfl(val, i + 10) * coef[2];
    s = s + shfl(val, i + 15) * coef[3];
    s = s + shfl(val, i + 20) * coef[4];
    return s;
}
// vxm_4(a, b)(j) = sum_k a(k) b(k,j) where a(k) is held by lane `base + stride*k`
__device__ __forceinline__ float contract_vec(float aval, int base, int stride,
                                              const float (&bcol)[NP]) {
    float s = shfl(aval, base) * bcol[0];
    s = s + shfl(aval, base + stride) * bcol[1];
    s = s + shfl(aval, base + 2 * stride) * bcol[2];
    s = s + shfl(aval, base + 3 * stride) * bcol[3];
    s = s + shfl(aval, base + 4 * stride) * bcol[4];
    return s;
}

// Per-lane rows/columns of the derivative matrices.
struct LaneG {
    float g2t_row[NP];   // G2T(i,k)   first-stage xi contraction, non-axial
    float g1t_row[NP];   // G1T(i,k)   ... axial
    float g2_col[NP];    // G2(k,j)    first-stage eta contraction
    float g2_row[NP];    // G2(i,k)    second stage, non-axial
    float g1_row[NP];    // G1(i,k)    second stage, axial
    float g2t_col[NP];   // G2T(k,j)   second stage eta
    float g0_i;          // G0(i)
    float g0[NP];        // G0(k)
};

// G is first staged in shared memory by the whole block (uniform parameter-space reads),
// then each lane picks its rows/columns (lane-dependent indices would serialise on the
// constant bank).
__device__ __forceinline__ void stage_g(const GMat &G, GMat &S) {
    const float *src = reinterpret_cast<const float *>(&G);
    float *dst = reinterpret_cast<float *>(&S);
    for (int t = threadIdx.x; t < (int)(sizeof(GMat) / sizeof(float)); t += blockDim.x) dst[t] = src[t];
    __syncthreads();
}
__device__ __forceinline__ void load_lane_g(const GMat &G, int i, int j, LaneG &L) {
#pragma unroll
    for (int k = 0; k < NP; k++) {
        L.g2t_row[k] = G.G2T[i + NP * k];
        L.g1t_row[k] = G.G1T[i + NP * k];
        L.g2_col[k] = G.G2[k + NP * j];
        L.g2_row[k] = G.G2[i + NP * k];
        L.g1_row[k] = G.G1[i + NP * k];
        L.g2t_col[k] = G.G2T[k + NP * j];
        L.g0[k] = G.G0[k];
    }
    L.g0_i = G.G0[i];
}

#define LDP(plane) ((plane)[pe])
#define LD0(plane) ((plane)[j + NP * (size_t)e])

// ---------------------------------------------------------------------------------------
// Elastic stiffness, one element per warp.  u*: displacement at this lane's point.
// Returns the local stiffness in l1,l2,l3 (component 2 unused for the monopole).
// stiffness_mono.f90:60-157
__device__ __forceinline__ void stiff_mono(const SolidPlanes &P, const LaneG &L, size_t pe,
                                           int e, int i, int j, bool ax, float us, float uz,
                                           float &ls, float &lz, float &X1o, float &X2o,
                                           float &X3o, float &X4o) {
    const int j5 = 5 * j;
    float X1, X2, X3, X4;
    if (!ax) { X1 = contract_xi(us, L.g2t_row, j5); X2 = contract_xi(uz, L.g2t_row, j5); }
    else     { X1 = contract_xi(us, L.g1t_row, j5); X2 = contract_xi(uz, L.g1t_row, j5); }
    X3 = contract_eta(us, L.g2_col, i);
    X4 = contract_eta(uz, L.g2_col, i);
    X1o = X1; X2o = X2; X3o = X3; X4o = X4;
    const float m_1 = LDP(P.M_1), m_2 = LDP(P.M_2), m_3 = LDP(P.M_3), m_4 = LDP(P.M_4);
    const float m_w1 = LDP(P.M_w1);
    const float m11s = LDP(P.M11s), m21s = LDP(P.M21s), m41s = LDP(P.M41s);
    const float m12s = LDP(P.M12s), m22s = LDP(P.M22s), m32s = LDP(P.M32s), m42s = LDP(P.M42s);
    const float m11z = LDP(P.M11z), m21z = LDP(P.M21z), m41z = LDP(P.M41z);
    ls = m_4 * X4 + m_2 * X3 + m_1 * X1 + m_3 * X2 + us * m_w1;
    float S1s = m11s * X3 + m21s * X1 + m12s * X4 + m22s * X2 + m_1 * us;
    float S2s = m11s * X1 + m41s * X3 + m32s * X2 + m42s * X4 + m_2 * us;
    float S1z = m11z * X4 + m21z * X2 + m32s * X3 + m22s * X1 + m_3 * us;
    float S2z = m11z * X2 + m41z * X4 + m12s * X1 + m42s * X3 + m_4 * us;
    X2 = contract_eta(S2s, L.g2t_col, i);
    X4 = contract_eta(S2z, L.g2t_col, i);
    if (!ax) { X1 = contract_xi(S1s, L.g2_row, j5); X3 = contract_xi(S1z, L.g2_row, j5); }
    else     { X1 = contract_xi(S1s, L.g1_row, j5); X3 = contract_xi(S1z, L.g1_row, j5); }
    ls = ls + X1 + X2;
    lz = X3 + X4;
    if (ax) {
        const float w1 = LD0(P.M0_w1), w2 = LD0(P.M0_w2), w3 = LD0(P.M0_w3);
        float V1 = contract_vec(us, j5, 1, L.g0);            // vxm_4(G0, us)(j): a(k)=G0(k), b=us
        // note: vxm_4(a,b) = sum_k a(k) b(k,j); here a = G0 (uniform), b(k,j) on lane 5j+k
        float V2 = contract_vec(uz, 0, 5, L.g2_col);         // vxm_4(uz0, G2): uz0(k)=uz(0,k) on lane 5k
        float V4 = w1 * V1 + w3 * V2;
        float V3 = contract_vec(uz, j5, 1, L.g0);            // vxm_4(G0, uz)
        V4 = V4 + w2 * V3;
        float X2a = L.g0_i * (w2 * V1);                      // outerprod_4(G0, m0_w2*V1)
        float V2b = w3 * V1;                                 // per column j
        float V1b = contract_vec(V2b, 0, 5, L.g2t_col);      // vxm_4(V2, G2T): V2(k) on lane 5k
        if (i == 0) X2a = X2a + V1b;
        ls = ls + L.g0_i * V4;
        lz = X2a + lz;
    }
}

// stiffness_di.f90:60-256
__device__ __forceinline__ void stiff_di(const SolidPlanes &P, const LaneG &L, size_t pe, int e,
                                         int i, int j, bool ax, float u1, float u2, float u3,
                                         float &l1, float &l2, float &l3, float (&Xo)[6]) {
    const int j5 = 5 * j;
    float X1, X2, X3, X4, X5, X6;
    X4 = contract_eta(u1, L.g2_col, i);
    X5 = contract_eta(u2, L.g2_col, i);
    X6 = contract_eta(u3, L.g2_col, i);
    if (!ax) { X1 = contract_xi(u1, L.g2t_row, j5); X2 = contract_xi(u2, L.g2t_row, j5); X3 = contract_xi(u3, L.g2t_row, j5); }
    else     { X1 = contract_xi(u1, L.g1t_row, j5); X2 = contract_xi(u2, L.g1t_row, j5); X3 = contract_xi(u3, L.g1t_row, j5); }
    Xo[0] = X1; Xo[1] = X2; Xo[2] = X3; Xo[3] = X4; Xo[4] = X5; Xo[5] = X6;
    const float m_1 = LDP(P.M_1), m_2 = LDP(P.M_2), m_3 = LDP(P.M_3), m_4 = LDP(P.M_4);
    const float m_5 = LDP(P.M_5), m_6 = LDP(P.M_6), m_7 = LDP(P.M_7), m_8 = LDP(P.M_8);
    const float m_w1 = LDP(P.M_w1), m_w2 = LDP(P.M_w2), m_w3 = LDP(P.M_w3);
    const float m11s = LDP(P.M11s), m21s = LDP(P.M21s), m41s = LDP(P.M41s);
    const float m12s = LDP(P.M12s), m22s = LDP(P.M22s), m42s = LDP(P.M42s);
    const float m13s = LDP(P.M13s), m23s = LDP(P.M32s), m33s = LDP(P.M33s), m43s = LDP(P.M43s);
    const float m11z = LDP(P.M11z), m21z = LDP(P.M21z), m41z = LDP(P.M41z);
    const float X7 = X1 + X2;
    const float X8 = X4 + X5;
    const float ls2 = m_8 * X6 + m_7 * X3 + m_1 * X1 + m_5 * X2 + m_2 * X4 + m_6 * X5 + m_w1 * u2 + m_w2 * u3;
    const float ls3 = m_4 * X4 - m_4 * X5 + m_3 * X1 - m_3 * X2 + m_w2 * u2 + m_w3 * u3;
    float c1 = m13s * X6, c2 = m23s * X3, c3 = m_3 * u3;
    float S1p = c1 + c2 + c3 + m11s * X4 + m21s * X1 + m12s * X5 + m22s * X2 + m_1 * u2;
    float S1m = c1 + c2 - c3 + m11s * X5 + m21s * X2 + m12s * X4 + m22s * X1 + m_5 * u2;
    c1 = m33s * X3; c2 = m43s * X6; c3 = m_4 * u3;
    float S2p = c1 + c2 + c3 + m11s * X1 + m41s * X4 + m12s * X2 + m42s * X5 + m_2 * u2;
    float S2m = c1 + c2 - c3 + m11s * X2 + m41s * X5 + m12s * X1 + m42s * X4 + m_6 * u2;
    float S1z = m33s * X8 + m23s * X7 + m11z * X6 + m21z * X3 + m_7 * u2;
    float S2z = m13s * X7 + m43s * X8 + m11z * X3 + m41z * X6 + m_8 * u2;
    if (!ax) { X1 = contract_xi(S1p, L.g2_row, j5); X3 = contract_xi(S1m, L.g2_row, j5); X5 = contract_xi(S1z, L.g2_row, j5); }
    else     { X1 = contract_xi(S1p, L.g1_row, j5); X3 = contract_xi(S1m, L.g1_row, j5); X5 = contract_xi(S1z, L.g1_row, j5); }
    X2 = contract_eta(S2p, L.g2t_col, i);
    X4 = contract_eta(S2m, L.g2t_col, i);
    X6 = contract_eta(S2z, L.g2t_col, i);
    l1 = X1 + X2;
    l2 = X3 + X4 + ls2;
    l3 = X5 + X6 + ls3;
    if (ax) {
        const float w1 = LD0(P.M0_w1), w2 = LD0(P.M0_w2), w3 = LD0(P.M0_w3), w4 = LD0(P.M0_w4);
        const float w6 = LD0(P.M0_w6), w7 = LD0(P.M0_w7), w8 = LD0(P.M0_w8), w9 = LD0(P.M0_w9);
        const float w10 = LD0(P.M0_w10);
        float V1 = contract_vec(u1, j5, 1, L.g0);
        float V2 = contract_vec(u2, j5, 1, L.g0);
        float V3 = contract_vec(u3, j5, 1, L.g0);
        float V4 = contract_vec(u1, 0, 5, L.g2_col);       // vxm_4(u10, G2)
        float s1p = L.g0_i * (w1 * V2 + w3 * V3);
        float s1m = L.g0_i * (w1 * V1 + (w2 + w6) * V4 + w9 * V2 + w10 * V3);
        float s1z = L.g0_i * (w3 * V1 + (w4 + w8) * V4 + w7 * V3 + w10 * V2);
        float V4b = (w2 + w6) * V2 + (w4 + w8) * V3;
        float V1b = contract_vec(V4b, 0, 5, L.g2t_col);    // vxm_4(V4, G2T)
        if (i == 0) s1p = s1p + V1b;
        l1 = l1 + s1p;
        l2 = l2 + s1m;
        l3 = l3 + s1z;
    }
}

// stiffness_quad.f90:238-412
__device__ __forceinline__ void stiff_quad(const SolidPlanes &P, const LaneG &L, size_t pe, int e,
                                           int i, int j, bool ax, float us, float up, float uz,
                                           float &ls, float &lp, float &lz, float (&Xo)[6]) {
    const int j5 = 5 * j;
    float X1, X2, X3, X4, X5, X6;
    if (!ax) { X1 = contract_xi(us, L.g2t_row, j5); X2 = contract_xi(up, L.g2t_row, j5); X3 = contract_xi(uz, L.g2t_row, j5); }
    else     { X1 = contract_xi(us, L.g1t_row, j5); X2 = contract_xi(up, L.g1t_row, j5); X3 = contract_xi(uz, L.g1t_row, j5); }
    X4 = contract_eta(us, L.g2_col, i);
    X5 = contract_eta(up, L.g2_col, i);
    X6 = contract_eta(uz, L.g2_col, i);
    Xo[0] = X1; Xo[1] = X2; Xo[2] = X3; Xo[3] = X4; Xo[4] = X5; Xo[5] = X6;
    const float m_1 = LDP(P.M_1), m_2 = LDP(P.M_2), m_3 = LDP(P.M_3), m_4 = LDP(P.M_4);
    const float m_5 = LDP(P.M_5), m_6 = LDP(P.M_6), m_7 = LDP(P.M_7), m_8 = LDP(P.M_8);
    const float m_w1 = LDP(P.M_w1), m_w2 = LDP(P.M_w2), m_w3 = LDP(P.M_w3), m_w4 = LDP(P.M_w4), m_w5 = LDP(P.M_w5);
    const float m11s = LDP(P.M11s), m21s = LDP(P.M21s), m41s = LDP(P.M41s);
    const float m12s = LDP(P.M12s), m22s = LDP(P.M22s), m32s = LDP(P.M32s), m42s = LDP(P.M42s);
    const float m11z = LDP(P.M11z), m21z = LDP(P.M21z), m41z = LDP(P.M41z);
    const float m1phi = LDP(P.M1phi), m2phi = LDP(P.M2phi), m4phi = LDP(P.M4phi);
    const float c1 = m_2 * X4, c2 = m_1 * X1, c3 = m_6 * X5, c4 = m_5 * X2, c5 = m_4 * X6, c6 = m_3 * X3;
    ls = c1 + c2 + 2 * (c3 + c4) + c5 + c6 + m_w1 * us + m_w2 * up + 2 * m_w3 * uz;
    lp = -2 * (c1 + c2 + c5 + c6) - (c3 + c4) + m_w2 * us + m_w4 * up - m_w3 * uz;
    lz = 2 * (m_8 * X5 + m_7 * X2) + m_w3 * (2 * us - up) + m_w5 * uz;
    float S1s = m11s * X4 + m21s * X1 + m12s * X6 + m22s * X3 + m_1 * (us - 2 * up);
    float S2s = m11s * X1 + m41s * X4 + m32s * X3 + m42s * X6 + m_2 * (us - 2 * up);
    float S1z = m11z * X6 + m21z * X3 + m32s * X4 + m22s * X1 + m_3 * (us - 2 * up);
    float S2z = m11z * X3 + m41z * X6 + m12s * X1 + m42s * X4 + m_4 * (us - 2 * up);
    float S1p = m1phi * X5 + m2phi * X2 + m_5 * (2 * us - up) + 2 * m_7 * uz;
    float S2p = m1phi * X2 + m4phi * X5 + m_6 * (2 * us - up) + 2 * m_8 * uz;
    X2 = contract_eta(S2s, L.g2t_col, i);
    X4 = contract_eta(S2p, L.g2t_col, i);
    X6 = contract_eta(S2z, L.g2t_col, i);
    if (!ax) { X1 = contract_xi(S1s, L.g2_row, j5); X3 = contract_xi(S1p, L.g2_row, j5); X5 = contract_xi(S1z, L.g2_row, j5); }
    else     { X1 = contract_xi(S1s, L.g1_row, j5); X3 = contract_xi(S1p, L.g1_row, j5); X5 = contract_xi(S1z, L.g1_row, j5); }
    ls = ls + X1 + X2;
    lp = lp + X3 + X4;
    lz = lz + X5 + X6;
    if (ax) {
        const float w1 = LD0(P.M0_w1), w2 = LD0(P.M0_w2), w3 = LD0(P.M0_w3);
        const float w4 = LD0(P.M0_w4), w5 = LD0(P.M0_w5), w6 = LD0(P.M0_w6);
        float V1 = contract_vec(us, j5, 1, L.g0);
        float V2 = contract_vec(up, j5, 1, L.g0);
        float V3 = contract_vec(uz, j5, 1, L.g0);
        ls = ls + L.g0_i * (w1 * V1 + w2 * V2 + w3 * V3);
        lp = lp + L.g0_i * (w2 * V1 + w4 * V2 + w5 * V3);
        lz = lz + L.g0_i * (w3 * V1 + w5 * V2 + w6 * V3);
    }
}

// ---------------------------------------------------------------------------------------
// Coarse-grained anelastic stiffness + memory-variable update for one element (one warp).
// `scr` is this warp's shared scratch (>= 64 floats).
// stiffness_{mono,di,quad}.f90 `glob_anel_stiffness_*_cg4`; attenuation.f90:81-202,471-535.
//
// Lane roles:  lane l < 24  <->  (k = l % 4, v = l / 4): memory variable component v+1 at
// coarse point k+1;   lanes with (i,j) in {1,3}x{1,3}: the coarse points themselves.
template <int ORDER>
__device__ __forceinline__ void anel_cg4(const AttCg &A, const LaneG &L, int e, int lane, int i,
                                         int j, bool ax, bool do_stiff, bool do_update,
                                         float u1, float u2, float u3, const float (&X)[6],
                                         float &l1, float &l2, float &l3, float *scr) {
    const int n_sls = A.n_sls;
    const int k = lane & 3, v = lane >> 2;                 // valid for lane < 24
    const bool mv_lane = lane < 24 && !(ORDER == 0 && (v == 3 || v == 5));
    const size_t mv_base = (size_t)24 * n_sls * e;         // (4,6,n_sls,nel)
    float R[8];                                            // n_sls <= 8
    float rsum = 0.0f;
    if (mv_lane) {
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (s < n_sls) { R[s] = A.memvar[mv_base + lane + 24 * s]; rsum = rsum + R[s]; }
    }
    // r(v)(k) for all v,k -> scratch [0..23]
    if (lane < 24) scr[lane] = rsum;
    __syncwarp();
    if (do_stiff) {
        // the four coarse points compute the S terms
        if (lane < 4) {
            const float yl = A.Ycg[lane + 4 * (size_t)e];
            const float vse = A.Vse[lane + 4 * (size_t)e], vsx = A.Vsx[lane + 4 * (size_t)e];
            const float vze = A.Vze[lane + 4 * (size_t)e], vzx = A.Vzx[lane + 4 * (size_t)e];
            const float r1 = scr[lane], r2 = scr[4 + lane], r3 = scr[8 + lane];
            const float r4 = scr[12 + lane], r5 = scr[16 + lane], r6 = scr[20 + lane];
            float *S = scr + 24;            // S[a*4 + k], a = 0..5 ; extra[a*4+k] at 48..59
            if (ORDER == 0) {
                S[0 + lane] = vze * r1 + vse * r5;      // S1s
                S[4 + lane] = vzx * r1 + vsx * r5;      // S2s
                S[16 + lane] = vze * r5 + vse * r3;     // S1z
                S[20 + lane] = vzx * r5 + vsx * r3;     // S2z
                scr[48 + lane] = yl * r2;
            } else if (ORDER == 1) {
                S[0 + lane] = vze * (r1 - r6) + vse * (r5 - r4);   // S1p
                S[4 + lane] = vzx * (r1 - r6) + vsx * (r5 - r4);   // S2p
                S[8 + lane] = vze * (r1 + r6) + vse * (r5 + r4);   // S1m
                S[12 + lane] = vzx * (r1 + r6) + vsx * (r5 + r4);  // S2m
                S[16 + lane] = vze * r5 + vse * r3;
                S[20 + lane] = vzx * r5 + vsx * r3;
                scr[48 + lane] = 2 * yl * (r2 - r6);
                scr[52 + lane] = yl * r4;
            } else {
                S[0 + lane] = vze * r1 + vse * r5;      // S1s
                S[4 + lane] = vzx * r1 + vsx * r5;      // S2s
                S[8 + lane] = vze * r6 + vse * r4;      // S1p
                S[12 + lane] = vzx * r6 + vsx * r4;     // S2p
                S[16 + lane] = vze * r5 + vse * r3;
                S[20 + lane] = vzx * r5 + vsx * r3;
                scr[48 + lane] = yl * (r2 - 2 * r6);
                scr[52 + lane] = yl * (r6 - 2 * r2);
                scr[56 + lane] = 2 * yl * r4;
            }
        }
        __syncwarp();
        if (lane < NPT) {
            const float *S = scr + 24;
            const float *ga = ax ? L.g1_row : L.g2_row;      // GA(i,k)
            // mxm_cg4_sparse_b(GA, S1): c(i,1) = GA(i,1) S1(1) + GA(i,3) S1(3);
            //                           c(i,3) = GA(i,1) S1(2) + GA(i,3) S1(4)
            // mxm_cg4_sparse_a(S2, G2T): c(1,j) = S2(1) G2T(1,j) + S2(2) G2T(3,j);
            //                            c(3,j) = S2(3) G2T(1,j) + S2(4) G2T(3,j)
            const int kb = (j == 1) ? 0 : 1;                 // column 1 -> S(1),S(3); column 3 -> S(2),S(4)
            const bool colb = (j == 1) || (j == 3);
            const int ka = (i == 1) ? 0 : 2;                 // row 1 -> S(1),S(2); row 3 -> S(3),S(4)
            const bool rowa = (i == 1) || (i == 3);
            const int cgk = (i == 1 ? 0 : 2) + (j == 1 ? 0 : 1);   // coarse index of (i,j)
            const bool cgpt = rowa && colb;
            float Xb[3], Xa[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float *S1 = S + 8 * a, *S2 = S + 8 * a + 4;
                Xb[a] = colb ? (ga[1] * S1[kb] + ga[3] * S1[kb + 2]) : 0.0f;
                Xa[a] = rowa ? (S2[ka] * L.g2t_col[1] + S2[ka + 1] * L.g2t_col[3]) : 0.0f;
            }
            if (ORDER == 0) {
                float ls = Xb[0] + Xa[0];
                float lz = Xb[2] + Xa[2];
                if (cgpt) ls = ls + scr[48 + cgk];
                l1 = l1 - ls;
                l3 = l3 - lz;
            } else if (ORDER == 1) {
                float lp = Xb[0] + Xa[0];
                float lm = Xb[1] + Xa[1];
                float lz = Xb[2] + Xa[2];
                if (cgpt) { lm = lm + scr[48 + cgk]; lz = lz - scr[52 + cgk]; }
                l1 = l1 - lp; l2 = l2 - lm; l3 = l3 - lz;
            } else {
                float ls = Xb[0] + Xa[0];
                float lp = -Xb[1] - Xa[1];
                float lz = Xb[2] + Xa[2];
                if (cgpt) { ls = ls + scr[48 + cgk]; lp = lp + scr[52 + cgk]; lz = lz - scr[56 + cgk]; }
                l1 = l1 - ls; l2 = l2 - lp; l3 = l3 - lz;
            }
        }
        __syncwarp();
    }
    if (!do_update) return;
    // ---- strain at the coarse points (compute_strain_att_el_cg4) ----
    // X[0..2] = d/dxi-type contraction (mxm1) of comps 1,2,3 ; X[3..5] = eta-type (mxm2)
    // for the dipole the gradient of (u1+u2) and (u1-u2) is contracted separately, as in
    // the reference (attenuation.f90:489, :515).
    float Xp1 = 0.f, Xp2 = 0.f, Xm1 = 0.f, Xm2 = 0.f;
    const int j5 = 5 * j;
    if (ORDER == 1) {
        const float up = u1 + u2, um = u1 - u2;
        if (!ax) { Xp1 = contract_xi(up, L.g2t_row, j5); Xm1 = contract_xi(um, L.g2t_row, j5); }
        else     { Xp1 = contract_xi(up, L.g1t_row, j5); Xm1 = contract_xi(um, L.g1t_row, j5); }
        Xp2 = contract_eta(up, L.g2_col, i);
        Xm2 = contract_eta(um, L.g2_col, i);
    }
    const bool rowa = (i == 1) || (i == 3), colb = (j == 1) || (j == 3);
    if (lane < NPT && rowa && colb) {
        const int cgk = (i == 1 ? 0 : 2) + (j == 1 ? 0 : 1);
        const size_t c4 = cgk + 4 * (size_t)e;
        const float dzdeta = A.Dze[c4], dzdxi = A.Dzx[c4], dsdeta = A.Dse[c4], dsdxi = A.Dsx[c4];
        const float is = A.inv_s[lane + NPT * (size_t)e];
        float g1, g2, g3, g4 = 0.f, g5, g6 = 0.f;
        // gradient of f: ds = dzdeta*m1 + dzdxi*m2 ; dz = dsdeta*m1 + dsdxi*m2
        float b2s = dzdeta * X[2] + dzdxi * X[5];     // d_s u3
        float b2z = dsdeta * X[2] + dsdxi * X[5];     // d_z u3
        if (ORDER == 0) {
            float b1s = dzdeta * X[0] + dzdxi * X[3];
            float b1z = dsdeta * X[0] + dsdxi * X[3];
            g1 = b1s; g3 = b2z; g5 = b1z + b2s;
            g2 = is * u1;
        } else if (ORDER == 1) {
            float b1s = dzdeta * Xp1 + dzdxi * Xp2;
            float b1z = dsdeta * Xp1 + dsdxi * Xp2;
            g1 = b1s; g3 = b2z; g5 = b1z + b2s;
            g2 = 2 * (is * u2);
            float c1s = dzdeta * Xm1 + dzdxi * Xm2;
            float c1z = dsdeta * Xm1 + dsdxi * Xm2;
            g4 = -(is * u3) - c1z;
            g6 = -g2 - c1s;
        } else {
            float b1s = dzdeta * X[0] + dzdxi * X[3];
            float b1z = dsdeta * X[0] + dsdxi * X[3];
            g1 = b1s; g3 = b2z; g5 = b1z + b2s;
            g2 = is * (u1 - 2 * u2);
            float c1s = dzdeta * X[1] + dzdxi * X[4];   // gradient of u2
            float c1z = dsdeta * X[1] + dsdxi * X[4];
            g4 = -2 * (is * u3) - c1z;
            g6 = is * (u2 - 2 * u1) - c1s;
        }
        float trace = g1 + g2;
        trace = trace + g3;
        const float dmu = A.dmu[c4], dka = A.dka[c4];
        const double third = 1.0 / 3.0;
        const double dm2 = (double)(dmu * 2);
        scr[0 + cgk] = (float)(dm2 * ((double)g1 - (double)trace * third));
        scr[4 + cgk] = (float)(dm2 * ((double)g2 - (double)trace * third));
        scr[8 + cgk] = (float)(dm2 * ((double)g3 - (double)trace * third));
        scr[12 + cgk] = (ORDER == 0) ? 0.0f : dmu * g4;
        scr[16 + cgk] = dmu * g5;
        scr[20 + cgk] = (ORDER == 0) ? 0.0f : dmu * g6;
        scr[24 + cgk] = dka * trace;                    // src_tr_t
    }
    __syncwarp();
    if (lane < 24) {
        const float src_dev_t = scr[lane];
        const size_t sb = lane + 24 * (size_t)e;
        const float s_dev_tm1 = A.src_dev_tm1[sb];
        if (mv_lane) {
            const float src_tr_t = scr[24 + k];
            const float s_tr_tm1 = A.src_tr_tm1[k + 4 * (size_t)e];
            const double *a_mu = A.a_mu_tab + (size_t)n_sls * A.qidx_mu[e];
            const double *a_ka = A.a_ka_tab + (size_t)n_sls * A.qidx_ka[e];
#pragma unroll
            for (int s = 0; s < 8; s++) {
                if (s < n_sls) {
                    const float dev_buf = (float)(A.ts_t[s] * a_mu[s] * (double)src_dev_t
                                                  + A.ts_tm1[s] * a_mu[s] * (double)s_dev_tm1);
                    float nv;
                    if (v < 3) {
                        const float tr_buf = (float)(A.ts_t[s] * a_ka[s] * (double)src_tr_t
                                                     + A.ts_tm1[s] * a_ka[s] * (double)s_tr_tm1);
                        nv = (float)(A.exp_w[s] * (double)R[s] + (double)dev_buf + (double)tr_buf);
                    } else {
                        nv = (float)(A.exp_w[s] * (double)R[s] + (double)dev_buf);
                    }
                    A.memvar[mv_base + lane + 24 * s] = nv;
                }
            }
        }
        A.src_dev_tm1[sb] = src_dev_t;
    }
    __syncwarp();
    if (lane < 4) A.src_tr_tm1[lane + 4 * (size_t)e] = scr[24 + lane];
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// S_A: solid predictor + axis mask + elastic stiffness [+ anelastic stiffness + memvars].
// Replaces time_evol_wave.F90:359-364, 392-422 (+453-457) / :599-602, 621-654.
template <int ORDER>
__global__ void __launch_bounds__(256)
k_solid_element(const __grid_constant__ GMat G, const __grid_constant__ SolidPlanes P,
                const __grid_constant__ AttCg A, const __grid_constant__ SolidStepArgs a) {
    __shared__ float s_scr[8][64];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const bool active = lane < NPT;
    const int q = active ? lane : 0;
    const int i = q % NP, j = q / NP;
    __shared__ GMat sG;
    stage_g(G, sG);
    LaneG L;
    load_lane_g(sG, i, j, L);
    const size_t cs = (size_t)NPT * a.nel;
    for (int e = blockIdx.x * warps_per_block + wib; e < a.nel; e += gridDim.x * warps_per_block) {
        const size_t pe = (size_t)NPT * e + q;
        const bool ax = a.axis[e] != 0;
        float u1 = 0.f, u2 = 0.f, u3 = 0.f;
        if (active) {
            u1 = a.disp[pe];
            if (ORDER != 0) u2 = a.disp[pe + cs];
            u3 = a.disp[pe + 2 * cs];
            if (a.mode == 0) {
                u1 = (float)((double)u1 + a.dt * (double)a.velo[pe] + a.half_dt_sq * (double)a.acc0[pe]);
                if (ORDER != 0)
                    u2 = (float)((double)u2 + a.dt * (double)a.velo[pe + cs] + a.half_dt_sq * (double)a.acc0[pe + cs]);
                u3 = (float)((double)u3 + a.dt * (double)a.velo[pe + 2 * cs] + a.half_dt_sq * (double)a.acc0[pe + 2 * cs]);
            } else if (a.mode == 1) {
                u1 = (float)((double)u1 + (double)a.velo[pe] * a.dt);
                if (ORDER != 0) u2 = (float)((double)u2 + (double)a.velo[pe + cs] * a.dt);
                u3 = (float)((double)u3 + (double)a.velo[pe + 2 * cs] * a.dt);
            }
            // apply_axis_mask_{one,two,three}comp (apply_masks.f90:55-100)
            if (ax && i == 0 && a.mode != 2) {
                if (ORDER == 0) u1 = 0.f;
                else if (ORDER == 1) { u2 = 0.f; u3 = 0.f; }
                else { u1 = 0.f; u2 = 0.f; u3 = 0.f; }
            }
            if (a.mode != 2) {
                a.disp[pe] = u1;
                if (ORDER != 0) a.disp[pe + cs] = u2;
                a.disp[pe + 2 * cs] = u3;
            }
        }
        float l1 = 0.f, l2 = 0.f, l3 = 0.f;
        float X[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (a.do_stiff) {
            if (ORDER == 0) {
                float x1, x2, x3, x4;
                stiff_mono(P, L, pe, e, i, j, ax, u1, u3, l1, l3, x1, x2, x3, x4);
                X[0] = x1; X[2] = x2; X[3] = x3; X[5] = x4;
            } else if (ORDER == 1) {
                stiff_di(P, L, pe, e, i, j, ax, u1, u2, u3, l1, l2, l3, X);
            } else {
                stiff_quad(P, L, pe, e, i, j, ax, u1, u2, u3, l1, l2, l3, X);
            }
        } else {
            // anelastic-only operator test: start from the stored acc1, contractions for
            // the strain are still needed when updating the memory variables
            if (active) { l1 = a.acc1[pe]; if (ORDER != 0) l2 = a.acc1[pe + cs]; l3 = a.acc1[pe + 2 * cs]; }
            if (a.anel >= 2) {
                const int j5 = 5 * j;
                const float (&gt)[NP] = ax ? L.g1t_row : L.g2t_row;
                X[0] = contract_xi(u1, gt, j5); X[1] = contract_xi(u2, gt, j5); X[2] = contract_xi(u3, gt, j5);
                X[3] = contract_eta(u1, L.g2_col, i); X[4] = contract_eta(u2, L.g2_col, i); X[5] = contract_eta(u3, L.g2_col, i);
            }
        }
        if (a.anel)
            anel_cg4<ORDER>(A, L, e, lane, i, j, ax, a.anel != 3, a.anel >= 2, u1, u2, u3, X, l1, l2, l3, s_scr[wib]);
        if (active && (a.do_stiff || a.anel == 1 || a.anel == 2)) {
            // apply_axis_mask_*(acc1) (time_evol_wave.F90:438-447); k_bdry2solid re-applies
            // it to the few points the S/F term touches afterwards
            if (ax && i == 0 && a.mode != 2) {
                if (ORDER == 0) l1 = 0.f;
                else if (ORDER == 1) { l2 = 0.f; l3 = 0.f; }
                else { l1 = 0.f; l2 = 0.f; l3 = 0.f; }
            }
            a.acc1[pe] = l1;
            if (ORDER != 0) a.acc1[pe + cs] = l2;
            a.acc1[pe + 2 * cs] = l3;
        }
    }
}

// ---------------------------------------------------------------------------------------
struct FluidStepArgs {
    int nel;
    int mode;                 // 0 Newmark, 1 symplectic drift, 2 none (op test)
    int order;                // source order (monopole: no M_w term / no axis masks)
    int full;                 // 1: apply source, S/F coupling and masks (time loop); 0: bare stiffness
    double dt, half_dt_sq;
    float *chi, *ddchi1;
    const float *dchi, *ddchi0;
    const int *axis;
    const float *M1chi, *M2chi, *M4chi, *M_w_fl, *M0_w_fl;
    const float *fs_mask;     // may be null
    // S/F coupling seen from the fluid: per fluid element the boundary index (1-based, 0 =
    // none) of its jpol=0 row and of its jpol=4 row
    const int2 *bdry_of_el;
    const int *bdry_sel, *bdry_js;
    const float *bdry_matr;   // (5, nel_bdry, 2)
    int nel_bdry;
    const float *disp;        // solid displacement (already predicted)
    size_t cs_solid;
    // fluid source
    int nelsrc;
    int ielsrc[8];
    const float *src_term;    // (5,5,8)
    const float *stf;         // stf(niter) ; index *iter
    const int *iter;
    int use_mask;             // Newmark multiplies by the free-surface mask, symplectic does not
};

// F_A: fluid predictor + stiffness + source + S/F term + masks.
// Replaces time_evol_wave.F90:357, 366-383 / :597, 604-614 and stiffness_fluid.f90:139-216.
__global__ void __launch_bounds__(256)
k_fluid_element(const __grid_constant__ GMat G, const __grid_constant__ FluidStepArgs a) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const bool active = lane < NPT;
    const int q = active ? lane : 0;
    const int i = q % NP, j = q / NP, j5 = 5 * j;
    __shared__ GMat sG;
    stage_g(G, sG);
    LaneG L;
    load_lane_g(sG, i, j, L);
    for (int e = blockIdx.x * warps_per_block + wib; e < a.nel; e += gridDim.x * warps_per_block) {
        const size_t pe = (size_t)NPT * e + q;
        const bool ax = a.axis[e] != 0;
        float c = 0.f;
        if (active) {
            c = a.chi[pe];
            if (a.mode == 0)
                c = (float)((double)c + a.dt * (double)a.dchi[pe] + a.half_dt_sq * (double)a.ddchi0[pe]);
            else if (a.mode == 1)
                c = (float)((double)c + (double)a.dchi[pe] * a.dt);
            if (a.full && a.order != 0 && ax && i == 0) c = 0.f;     // apply_axis_mask_scal(chi)
            if (a.mode != 2) a.chi[pe] = c;
        }
        float X1 = ax ? contract_xi(c, L.g1t_row, j5) : contract_xi(c, L.g2t_row, j5);
        float X2 = contract_eta(c, L.g2_col, i);
        float m1 = 0.f, m2 = 0.f, m4 = 0.f;
        if (active) { m1 = a.M1chi[pe]; m2 = a.M2chi[pe]; m4 = a.M4chi[pe]; }
        float S1 = m1 * X2 + m2 * X1;
        float S2 = m1 * X1 + m4 * X2;
        X1 = ax ? contract_xi(S1, L.g1_row, j5) : contract_xi(S1, L.g2_row, j5);
        X2 = contract_eta(S2, L.g2t_col, i);
        float l = X1 + X2;
        if (a.order != 0) {
            const float mw = active ? a.M_w_fl[pe] : 0.f;
            l = l + mw * c;
            if (ax) {
                const float m0 = a.M0_w_fl[j + NP * (size_t)e];
                float V1 = contract_vec(c, j5, 1, L.g0);
                l = l + L.g0_i * (m0 * V1);
            }
        }
        if (a.full && active) {
            // add_source_fl (time_evol_wave.F90:1062-1076)
            if (a.nelsrc > 0) {
                const float stf1 = a.stf[*a.iter];
                if (stf1 != 0.f)
                    for (int k = 0; k < a.nelsrc; k++)
                        if (a.ielsrc[k] - 1 == e) l = l - a.src_term[q + NPT * k] * stf1;
            }
            // bdry_copy2fluid (time_evol_wave.F90:1532-1571)
            if (a.nel_bdry > 0 && (j == 0 || j == 4)) {
                const int2 bd = a.bdry_of_el[e];
                const int b = (j == 0 ? bd.x : bd.y) - 1;
                if (b >= 0) {
                    const size_t ps = i + NP * a.bdry_js[b] + (size_t)NPT * (a.bdry_sel[b] - 1);
                    const float B1 = a.bdry_matr[i + NP * (size_t)b];
                    const float B2 = a.bdry_matr[i + NP * ((size_t)b + a.nel_bdry)];
                    const float us = a.disp[ps], uz = a.disp[ps + 2 * a.cs_solid];
                    if (a.order == 1) l = l - B1 * (us + a.disp[ps + a.cs_solid]) - B2 * uz;
                    else l = l - B1 * us - B2 * uz;
                }
            }
            if (a.order != 0 && ax && i == 0) l = 0.f;                // apply_axis_mask_scal(ddchi1)
            if (a.use_mask && a.fs_mask) l = l * a.fs_mask[pe];
        }
        if (active) a.ddchi1[pe] = l;
    }
}

// ---------------------------------------------------------------------------------------
// Assembly group table (pull-style direct stiffness summation):
// for every element-local point p, gid[p] < 0  -> not shared;
// else grp[gid[p]] = nloc, grp[..+1] = nrem, then nloc local point addresses in ascending
// element order (commun.F90:101-128), then nrem receive-slab slots in message order
// (commpi.F90:469-477).
struct AsmTable {
    const int *gid;
    const int *grp;
};

__device__ __forceinline__ float assembled(const float *vec, const AsmTable &T, int g,
                                           const float *recv, size_t recv_cs, int c) {
    const int nloc = T.grp[g], nrem = T.grp[g + 1];
    float s = 0.0f;
    for (int m = 0; m < nloc; m++) s = s + vec[T.grp[g + 2 + m]];
    for (int m = 0; m < nrem; m++) s = s + recv[T.grp[g + 2 + nloc + m] + recv_cs * c];
    return s;
}

struct FluidCorrArgs {
    int npts;
    int mode;                 // 0 Newmark, 1 symplectic
    double half_dt;           // Newmark: dt/2 ; symplectic: coefv
    float *ddchi1, *ddchi0, *dchi;
    const float *chi;
    const float *inv_mass_fluid, *gamma;   // gamma may be null
    AsmTable T;
    const float *recv; size_t recv_cs;
    int assemble_only;
};

// F_B: pdistsum_fluid + mass inversion + sponge + velocity-potential update.
// Replaces commun.F90:180-283 (+commpi.F90:587-637) and time_evol_wave.F90:430-434, 459-460.
__global__ void __launch_bounds__(256) k_fluid_corrector(const __grid_constant__ FluidCorrArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    float v = a.ddchi1[p];
    const int g = a.T.gid[p];
    if (g >= 0) v = assembled(a.ddchi1, a.T, g, a.recv, a.recv_cs, 0);
    if (a.assemble_only) { a.ddchi0[p] = v; return; }   // op test: result staged in ddchi0
    if (a.mode == 0) v = -a.inv_mass_fluid[p] * v;
    else v = -v * a.inv_mass_fluid[p];
    const float dc = a.dchi[p];
    if (a.gamma) {
        const float gm = a.gamma[p];
        v = v - 2 * gm * dc - (gm * gm) * a.chi[p];
    }
    if (a.mode == 0) {
        a.dchi[p] = (float)((double)dc + a.half_dt * (double)(a.ddchi0[p] + v));
    } else {
        a.dchi[p] = (float)((double)dc + a.half_dt * (double)v);
    }
    a.ddchi0[p] = v;          // ddchi0 = ddchi1 (Newmark); also where S_bdry reads it
}

struct BdrySolidArgs {
    int nel_bdry, order;
    const int *bdry_sel, *bdry_fel, *bdry_js, *bdry_jf;
    const float *bdry_matr;
    const int *axis_solid;
    const float *uflu;        // assembled, mass-inverted ddchi1
    float *acc1; size_t cs;
};
// bdry_copy2solid + axis mask of the touched points (time_evol_wave.F90:1577-1611, 438-447)
__global__ void k_bdry2solid(const __grid_constant__ BdrySolidArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nel_bdry * NP) return;
    const int b = t / NP, i = t % NP;
    const int es = a.bdry_sel[b] - 1, ef = a.bdry_fel[b] - 1;
    const size_t ps = i + NP * a.bdry_js[b] + (size_t)NPT * es;
    const size_t pf = i + NP * a.bdry_jf[b] + (size_t)NPT * ef;
    const float B1 = a.bdry_matr[i + NP * (size_t)b], B2 = a.bdry_matr[i + NP * ((size_t)b + a.nel_bdry)];
    const float f = a.uflu[pf];
    float a1 = a.acc1[ps] + B1 * f;
    float a2 = 0.f;
    if (a.order == 1) a2 = a.acc1[ps + a.cs] + B1 * f;
    float a3 = a.acc1[ps + 2 * a.cs] + B2 * f;
    if (a.axis_solid[es] && i == 0) {
        if (a.order == 0) a1 = 0.f;
        else if (a.order == 1) { a2 = 0.f; a3 = 0.f; }
        else { a1 = 0.f; a3 = 0.f; }
    }
    a.acc1[ps] = a1;
    if (a.order == 1) a.acc1[ps + a.cs] = a2;
    a.acc1[ps + 2 * a.cs] = a3;
}
// the reverse coupling as a stand-alone operator (op test only; the loop fuses it in F_A)
__global__ void k_bdry2fluid(int nel_bdry, int order, const int *bsel, const int *bfel,
                             const int *bjs, const int *bjf, const float *bm,
                             const float *usol, size_t cs, float *uflu) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nel_bdry * NP) return;
    const int b = t / NP, i = t % NP;
    const size_t ps = i + NP * bjs[b] + (size_t)NPT * (bsel[b] - 1);
    const size_t pf = i + NP * bjf[b] + (size_t)NPT * (bfel[b] - 1);
    const float B1 = bm[i + NP * (size_t)b], B2 = bm[i + NP * ((size_t)b + nel_bdry)];
    if (order == 1) uflu[pf] = uflu[pf] - B1 * (usol[ps] + usol[ps + cs]) - B2 * usol[ps + 2 * cs];
    else uflu[pf] = uflu[pf] - B1 * usol[ps] - B2 * usol[ps + 2 * cs];
}

struct SolidCorrArgs {
    int npts;                 // 25 * nel
    int order, mode;          // mode 0 Newmark, 1 symplectic
    double half_dt;           // dt/2 or coefv
    float *acc1, *acc0, *velo;
    const float *disp;
    const float *inv_mass_rho, *gamma;
    AsmTable T;
    const float *recv; size_t recv_cs;
    int nelsrc;
    int ielsrc[8];
    const float *src_term;    // (5,5,8,3)
    const float *stf;         // Newmark: stf(niter), symplectic: stf_symp(nstages, niter)
    const int *iter;
    int stf_stride, stf_off;  // index = iter*stride + off
    int assemble_only;
};

// S_B: pdistsum_solid + source + mass inversion + sponge + velocity update.
// Replaces commun.F90:69-171 (+commpi.F90:453-500) and time_evol_wave.F90:466-494 / :689-715.
template <int ORDER>
__global__ void __launch_bounds__(256) k_solid_corrector(const __grid_constant__ SolidCorrArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    const size_t cs = (size_t)a.npts;
    const int g = a.T.gid[p];
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (ORDER == 0 && c == 1) { v[c] = 0.f; continue; }
        v[c] = a.acc1[p + cs * c];
    }
    if (g >= 0) {
        const int nloc = a.T.grp[g], nrem = a.T.grp[g + 1];
        float s[3] = {0.f, 0.f, 0.f};
        for (int m = 0; m < nloc; m++) {
            const int ad = a.T.grp[g + 2 + m];
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (!(ORDER == 0 && c == 1)) s[c] = s[c] + a.acc1[ad + cs * c];
        }
        for (int m = 0; m < nrem; m++) {
            const int sl = a.T.grp[g + 2 + nloc + m];
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (!(ORDER == 0 && c == 1)) s[c] = s[c] + a.recv[sl + a.recv_cs * c];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = s[c];
    }
    if (a.assemble_only) {
        // op test: stage the assembled field in acc0 (acc1 must stay intact while other
        // threads still pull from it)
#pragma unroll
        for (int c = 0; c < 3; c++) if (!(ORDER == 0 && c == 1)) a.acc0[p + cs * c] = v[c];
        return;
    }
    // add_source_el (time_evol_wave.F90:1082-1097)
    if (a.nelsrc > 0) {
        const float stf1 = a.stf[(size_t)(*a.iter) * a.stf_stride + a.stf_off];
        if (stf1 != 0.f) {
            const int e = p / NPT, q = p - e * NPT;
            for (int k = 0; k < a.nelsrc; k++)
                if (a.ielsrc[k] - 1 == e) {
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        if (!(ORDER == 0 && c == 1)) v[c] = v[c] - a.src_term[q + NPT * (k + 8 * c)] * stf1;
                }
        }
    }
    const float im = a.inv_mass_rho[p];
    const float gm = a.gamma ? a.gamma[p] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (ORDER == 0 && c == 1) continue;
        float x = v[c];
        const float vel = a.velo[p + cs * c];
        if (a.mode == 0) {
            if (ORDER == 1 && c == 2) x = (float)(-2.0 * (double)im * (double)x);
            else x = -im * x;
            if (a.gamma) x = x - 2 * gm * vel - (gm * gm) * a.disp[p + cs * c];
            a.velo[p + cs * c] = (float)((double)vel + a.half_dt * (double)(a.acc0[p + cs * c] + x));
            a.acc0[p + cs * c] = x;
        } else {
            x = -im * x;
            if (a.gamma) x = x - 2 * gm * vel - (gm * gm) * a.disp[p + cs * c];
            if (ORDER == 1 && c == 2) a.velo[p + cs * c] = (float)((double)vel + 2.0 * (double)x * a.half_dt);
            else a.velo[p + cs * c] = (float)((double)vel + (double)x * a.half_dt);
            a.acc0[p + cs * c] = x;     // keeps the reference's `acc` available to get_state
        }
    }
}

// final drift of a symplectic step (time_evol_wave.F90:720-725)
__global__ void k_drift(int n, float *x, const float *v, double cd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) x[p] = (float)((double)x[p] + (double)v[p] * cd);
}

// ---------------------------------------------------------------------------------------
// Halo pack: partial sums of the shared points in glob2el order, written straight into the
// neighbour's receive slab (commpi.F90:371-404, 408-449).  One thread per (entry, comp).
struct PackArgs {
    int nentries, nc;
    const int *start;         // CSR over entries
    const int *addr;          // local point addresses, glob2el order
    const float *vec; size_t cs;
    const int *dst_msg;       // message index of each entry
    const int *dst_slot;      // slot inside the peer's slab
    float *dst_base[8];       // per message: peer slab base for the current parity
    size_t dst_cs[8];         // per message: component stride of the peer slab
};
__global__ void k_halo_pack(const __grid_constant__ PackArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nentries * a.nc) return;
    const int en = t % a.nentries, c = t / a.nentries;
    float s = 0.0f;
    for (int m = a.start[en]; m < a.start[en + 1]; m++) s = s + a.vec[a.addr[m] + a.cs * c];
    const int msg = a.dst_msg[en];
    a.dst_base[msg][a.dst_slot[en] + a.dst_cs[msg] * c] = s;
}
// release: all packed data of this kernel-ordered stream is visible before the flag
struct FlagArgs { int n; volatile int *flag[8]; int value; };
__global__ void k_halo_signal(const __grid_constant__ FlagArgs a) {
    if (threadIdx.x < a.n) {
        __threadfence_system();
        *a.flag[threadIdx.x] = a.value;
    }
}
__global__ void k_halo_wait(const __grid_constant__ FlagArgs a) {
    if (threadIdx.x < a.n) {
        while (*a.flag[threadIdx.x] < a.value) { __nanosleep(200); }
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------
struct RecArgs {
    int num_rec, order, seis_it, nseismo_max;
    const int *recfile_el;    // (num_rec,3)
    const float *disp; size_t cs;
    float *recdump;           // (3, num_rec, nseismo_max)
    int *counters;            // [0] iter, [1] iseismo, [2] istrain
};
// nc_compute_recfile_seis_bare (seismograms.f90:783-820) every seis_it steps
__global__ void k_sample_receivers(const __grid_constant__ RecArgs a) {
    const int iter = a.counters[0];
    if (iter % a.seis_it != 0) return;
    const int is = a.counters[1];
    if (is >= a.nseismo_max) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.num_rec) return;
    const int iel = a.recfile_el[r], ip = a.recfile_el[r + a.num_rec], jp = a.recfile_el[r + 2 * a.num_rec];
    const size_t p = ip + NP * jp + (size_t)NPT * (iel - 1);
    const float d1 = a.disp[p], d2 = a.disp[p + a.cs], d3 = a.disp[p + 2 * a.cs];
    float *out = a.recdump + (size_t)3 * a.num_rec * is + 3 * r;
    if (a.order == 0) { out[0] = d1; out[1] = 0.f; out[2] = d3; }
    else if (a.order == 1) { out[0] = d1 + d2; out[1] = d1 - d2; out[2] = d3; }
    else { out[0] = d1; out[1] = d2; out[2] = d3; }
}

struct DumpArgs {
    int nel_s, nel_f, order, strain_it, nstrain_max;
    const int *kwf_mask, *kwf_map;
    const float *disp, *chi; size_t cs;
    const int *axis_f;
    const float *inv_rho, *Dse, *Dze, *Dsx, *Dzx;
    float *snap; size_t npts;
    int *counters;
};
// dump_disp_global, solid part (wavefields_io.f90:1041-1052, 743-762)
__global__ void k_dump_solid(const __grid_constant__ DumpArgs a) {
    const int iter = a.counters[0];
    if (iter % a.strain_it != 0) return;
    const int is = a.counters[2];
    if (is >= a.nstrain_max) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NPT * a.nel_s) return;
    if (!a.kwf_mask[p]) return;
    const int ct = a.kwf_map[p] - 1;
    const size_t vs = a.npts * a.nstrain_max;
    float *base = a.snap + a.npts * is;
    const float u1 = a.disp[p], u2 = a.disp[p + a.cs], u3 = a.disp[p + 2 * a.cs];
    float f1 = u1, f2 = u2;
    if (a.order == 1) { f1 = u1 + u2; f2 = u1 - u2; }
    base[ct] = f1;
    if (a.order != 0) base[ct + vs] = f2;
    base[ct + 2 * vs] = u3;
}
// dump_disp_global, fluid part: u = 1/rho grad(chi) (wavefields_io.f90:1073-1090,
// pointwise_derivatives.f90:509-546); one warp per fluid element
__global__ void __launch_bounds__(256)
k_dump_fluid(const __grid_constant__ GMat G, const __grid_constant__ DumpArgs a) {
    const int iter = a.counters[0];
    if (iter % a.strain_it != 0) return;
    const int is = a.counters[2];
    if (is >= a.nstrain_max) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const bool active = lane < NPT;
    const int q = active ? lane : 0, i = q % NP, j = q / NP, j5 = 5 * j;
    __shared__ GMat sG;
    stage_g(G, sG);
    LaneG L;
    load_lane_g(sG, i, j, L);
    const size_t vs = a.npts * a.nstrain_max;
    float *base = a.snap + a.npts * is;
    for (int e = blockIdx.x * wpb + wib; e < a.nel_f; e += gridDim.x * wpb) {
        const size_t pe = (size_t)NPT * e + q;
        const bool ax = a.axis_f[e] != 0;
        const float c = active ? a.chi[pe] : 0.f;
        const float m1 = ax ? contract_xi(c, L.g1t_row, j5) : contract_xi(c, L.g2t_row, j5);
        const float m2 = contract_eta(c, L.g2_col, i);
        if (!active) continue;
        const size_t pk = pe + (size_t)NPT * a.nel_s;
        if (!a.kwf_mask[pk]) continue;
        const int ct = a.kwf_map[pk] - 1;
        const float dsdf = a.Dze[pe] * m1 + a.Dzx[pe] * m2;
        const float dzdf = a.Dse[pe] * m1 + a.Dsx[pe] * m2;
        base[ct] = a.inv_rho[pe] * dsdf;
        if (a.order != 0) base[ct + vs] = 0.f;
        base[ct + 2 * vs] = a.inv_rho[pe] * dzdf;
    }
}

// end of step: iter += 1 ; sample counters advance where a dump happened
__global__ void k_advance(int *counters, int seis_it, int strain_it, int num_rec, int have_kwf,
                          int nseismo_max, int nstrain_max, int pre) {
    // pre = 1: account for the dumps of the current iter (called after the dump kernels)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int iter = counters[0];
        if (num_rec > 0 && iter % seis_it == 0 && counters[1] < nseismo_max) counters[1] += 1;
        if (have_kwf && strain_it > 0 && iter % strain_it == 0 && counters[2] < nstrain_max) counters[2] += 1;
        (void)pre;
    }
}
__global__ void k_next_iter(int *counters) {
    if (threadIdx.x == 0 && blockIdx.x == 0) counters[0] += 1;
}

}  // namespace axb
