// axb_kernels.cuh — sm_100a kernels of the AxiSEM time loop.
//
// Solid elements (S_A): TMA-bulk ring + thread-per-point tiles, see axb_solid_tile.cuh.
// Fluid elements (F_A): the same design, axb_fluid_tile.cuh.  The wavefield-dump kernel of
// the fluid keeps the warp-per-element mapping with shuffle contractions.  Pointwise kernels
// (correctors with assembly) are one thread per GLL point.
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

// real(4) <-> real(8) conversions.  The reference runs with flush-to-zero (SOLVER/ftz.c:44-48)
// and --ftz=true makes nvcc emulate that on conversions as well, at three instructions each
// on sm_100a.  The bit-exact build (-DAXB_STRICT) keeps that; the product build converts with
// a plain cvt, which is identical unless a value is subnormal (below 1.2e-38).
// rnd32(x): the value a real(8) expression takes when the Fortran assigns it to a real(4)
// variable that is then promoted again.  The bit-exact build rounds; the product build keeps
// the real(8) value (one rounding less: error below 6e-8 relative, two conversions saved).
#ifdef AXB_STRICT
__device__ __forceinline__ double f2d(float x) { return (double)x; }
__device__ __forceinline__ float d2f(double x) { return (float)x; }
__device__ __forceinline__ double rnd32(double x) { return (double)(float)x; }
#else
__device__ __forceinline__ double rnd32(double x) { return x; }
__device__ __forceinline__ double f2d(float x) { double d; asm("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x)); return d; }
__device__ __forceinline__ float d2f(double x) { float f; asm("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x)); return f; }
#endif

struct GMat {            // Fortran order: G(i,j) at [i + 5*j]
    float G0[NP];
    float G1[NPT], G1T[NPT], G2[NPT], G2T[NPT];
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

// ---------------------------------------------------------------------------------------
// Assembly tables (pull-style direct stiffness summation).  Only the 16
// edge points of an element can be shared (commun.F90:110-120); each has one int4 entry
// cp[16*e + slot]:
//   x == -1        not shared: the point keeps its own value;
//   x >= 0         2..4 local copies and no remote ones: x,y,z,w (-1 padded) are their point
//                  addresses in ascending element order — the order commun.F90:101-128 sums in;
//   x == -2        general group (more copies, or halo partners): grp[y] = nloc, grp[y+1] =
//                  nrem, then nloc local addresses, then nrem receive-slab slots in message
//                  order (commpi.F90:469-477).
// One coalesced 16-byte load then tells a thread everything, and the value loads that
// follow are independent of each other.
// Warm the L2 for the block that will run `ahead` points further on: the correctors are bound by
// the latency of two dependent DRAM round trips per thread (assembly entry -> gather), not by
// bandwidth; a prefetch one wave ahead turns the first of them into an L2 hit.
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
struct AsmTable {
    const int4 *cp;
    const int *grp;
};
// slot of element-local point q among the 16 edge points (-1: interior): row j = 0 -> 0..4,
// (i = 0 | 4, j = 1..3) -> 5..10, row j = 4 -> 11..15.  Looked up in two packed constants (four
// bits per point) instead of the division and four branches it takes to compute.
__device__ __forceinline__ int edge_slot(int q) {
    constexpr unsigned INTERIOR = (7u << 6) | (7u << 11) | (7u << 16);
    constexpr unsigned long long LO = 0x9800076000543210ull;   // q = 0..15
    constexpr unsigned long long HI = 0x0000000fedcba000ull;   // q = 16..24
    const unsigned long long w = (q & 16) ? HI : LO;
    const int slot = (int)((w >> ((q & 15) * 4)) & 15ull);
    return ((INTERIOR >> q) & 1u) ? -1 : slot;
}

// Halo receive side.  Every value a neighbour GPU delivers travels as one 8-byte word
// {value bits, exchange number} written with a single store into this rank's receive slab (the
// "LL" protocol of collective libraries): the word is its own arrival flag, so the sender needs
// no fence, no block counter and no separate flag store, and the receiver needs no fence between
// flag and data.  Only the threads of cut points read the slab, spinning on their own words, so
// the rest of a corrector overlaps the exchange (the reference overlaps it with the next
// stiffness call, time_evol_wave.F90:386-427).
// A neighbour that never delivers (a dead peer process, a rank that was not stepped) must not
// hang the GPU: the spin is bounded by `timeout_ns`, and the first thread to give up raises
// the handle's abort flag, which every later wait honours at once.  axb_synchronize reports it
// the way the reference's pcheck stops a run (commpi.F90:64-111).
struct HaloArrival {
    int nmsg, value;             // value: exchange number the words of this exchange carry
    int off[8];                  // first slot of each message (to name the late message)
    volatile int *abort;         // counters[2]: 0 fine, m + 1 = message m timed out
    unsigned long long timeout_ns;
    const int *dyn;              // graph replay: value = dyn[DYN_SEQ] + value
};
// device-resident step counters, used instead of by-value arguments when a step is replayed
// from a CUDA graph (axb_api.cu)
enum { DYN_ITER = 0, DYN_SEQ = 1, DYN_ISEISMO = 2 };
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int2 ld_volatile_v2(const int2 *p) {
    int2 w;
    asm volatile("ld.volatile.global.v2.s32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p));
    return w;
}
__device__ __forceinline__ void st_volatile_v2(int2 *p, int x, int y) {
    asm volatile("st.volatile.global.v2.s32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ int halo_expected(const HaloArrival &h) {
    return h.dyn ? h.dyn[DYN_SEQ] + h.value : h.value;
}
// raise the abort flag for the message that owns `slot`
__device__ __forceinline__ void halo_give_up(const HaloArrival &h, int slot) {
    int m = 0;
    while (m + 1 < h.nmsg && slot >= h.off[m + 1]) m++;
    *h.abort = m + 1;
    __threadfence();
}
// the NC values (one per field component, component stride cs) in slot `slot` of the slab once the
// neighbour's words of exchange `want` are there; component 1 is skipped for SKIP1 (monopole)
template <int NC, bool SKIP1>
__device__ __forceinline__ void halo_read(const int2 *slab, size_t cs, int slot, int want, const HaloArrival &h,
                                          float (&out)[NC]) {
    int2 w[NC];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        if (SKIP1 && c == 1) { w[c] = make_int2(0, want); continue; }
        w[c] = ld_volatile_v2(slab + cs * c + slot);
        ok = ok && w[c].y == want;
    }
    if (!ok) {
        const unsigned long long t0 = global_ns();
        while (true) {
            ok = true;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if (SKIP1 && c == 1) continue;
                w[c] = ld_volatile_v2(slab + cs * c + slot);
                ok = ok && w[c].y == want;
            }
            if (ok || *h.abort) break;
            if (global_ns() - t0 > h.timeout_ns) { halo_give_up(h, slot); break; }
            __nanosleep(100);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) out[c] = ok ? __int_as_float(w[c].x) : 0.f;
}

}  // namespace axb
#include "axb_solid_tile.cuh"
#include "axb_fluid_tile.cuh"
#include "axb_anel_full.cuh"
namespace axb {

#ifndef AXB_CORR_MINB
#define AXB_CORR_MINB 6      // resident 256-thread blocks per SM the correctors are compiled for
#endif
struct FluidCorrArgs {
    int npts;
    int mode;                 // 0 Newmark, 1 symplectic, 2 lean Newmark (dchi holds dchi + dt/2 ddchi)
    double half_dt;           // Newmark: dt/2 ; symplectic: coefv ; lean Newmark: dt
    float *ddchi1, *ddchi0, *dchi;
    const float *chi;
    const float *inv_mass_fluid, *gamma;   // gamma may be null
    AsmTable T;
    const int2 *recv; size_t recv_cs;    // both parities: [2][nc][recv_cs] words; recv_parity picks one
    int recv_parity;
    HaloArrival arrival;
    int assemble_only;
};
// the receive slab of the current exchange (graph replay: parity from the device counter)
__device__ __forceinline__ const int2 *recv_slab(const int2 *recv, size_t cs, int nc, int parity, const int *dyn) {
    if (dyn) parity = dyn[DYN_SEQ] & 1;
    return recv + (size_t)parity * nc * cs;
}

// F_B: pdistsum_fluid + mass inversion + sponge + velocity-potential update.
// Replaces commun.F90:180-283 (+commpi.F90:587-637) and time_evol_wave.F90:430-434, 459-460.
#ifndef AXB_FCORR_MINB
#define AXB_FCORR_MINB 8     // the fluid corrector fits 32 registers
#endif
__global__ void __launch_bounds__(256, AXB_FCORR_MINB) k_fluid_corrector(const __grid_constant__ FluidCorrArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    float v = a.ddchi1[p];
    const int e = p / NPT, slot = edge_slot(p - e * NPT);
    const float imf = a.assemble_only ? 0.f : a.inv_mass_fluid[p];
    const float dc = a.assemble_only ? 0.f : a.dchi[p];
    const float dd0 = (a.assemble_only || a.mode != 0) ? 0.f : a.ddchi0[p];
    if (slot >= 0) {
        const int4 c = a.T.cp[16 * (size_t)e + slot];
        if (c.x >= 0) {
            float s = 0.0f;
            s = s + a.ddchi1[c.x];
            s = s + a.ddchi1[c.y];
            if (c.z >= 0) s = s + a.ddchi1[c.z];
            if (c.w >= 0) s = s + a.ddchi1[c.w];
            v = s;
        } else if (c.x == -2) {
            const int g = c.y;
            const int nloc = a.T.grp[g], nrem = a.T.grp[g + 1];
            float s = 0.0f;
            for (int m = 0; m < nloc; m++) s = s + a.ddchi1[a.T.grp[g + 2 + m]];
            if (nrem > 0) {
                const int2 *recv = recv_slab(a.recv, a.recv_cs, 1, a.recv_parity, a.arrival.dyn);
                const int want = halo_expected(a.arrival);
                for (int m = 0; m < nrem; m++) {
                    float r[1];
                    halo_read<1, false>(recv, 0, a.T.grp[g + 2 + nloc + m], want, a.arrival, r);
                    s = s + r[0];
                }
            }
            v = s;
        }
    }
    if (a.assemble_only) { a.ddchi0[p] = v; return; }   // op test: result staged in ddchi0
    if (a.mode != 1) v = -imf * v;
    else v = -v * imf;
    if (a.gamma) {
        const float gm = a.gamma[p];
        v = v - 2 * gm * dc - (gm * gm) * a.chi[p];
    }
    if (a.mode == 0) {
        a.dchi[p] = d2f(f2d(dc) + a.half_dt * f2d(dd0 + v));
    } else {
        // symplectic kick, or lean Newmark: (dchi + dt/2 ddchi)_new = (dchi + dt/2 ddchi)_old + dt ddchi_new
        a.dchi[p] = d2f(f2d(dc) + a.half_dt * f2d(v));
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
    size_t cs;                // component stride (25 * padded element count)
    int order, mode;          // mode 0 Newmark, 1 symplectic, 2 lean Newmark (velo holds v + dt/2 a),
                              // 3 back from lean: acc0 = a, velo = velo - dt/2 a
    double half_dt;           // dt/2 or coefv ; lean: dt ; back from lean: dt/2
    float *acc1, *acc0, *velo;
    const float *disp;
    const float *inv_mass_rho, *gamma;
    AsmTable T;
    const int2 *recv; size_t recv_cs;    // both parities: [2][3][recv_cs] words
    int recv_parity;
    HaloArrival arrival;
    const int *dyn;           // graph replay: iter = dyn[DYN_ITER]
    int ahead;                // points to prefetch ahead (0: off)
    int nelsrc;
    int src_emin, src_emax;   // 0-based range of the source elements: one compare pair per point in
                              // front of the eight-way search (which was a quarter of this
                              // kernel's instructions, profiles/r01i)
    int ielsrc[8];
    const float *src_term;    // (5,5,8,3)
    const float *stf;         // Newmark: stf(niter), symplectic: stf_symp(nstages, niter)
    int iter;                 // time steps completed before this one (index into stf)
    int stf_stride, stf_off;  // index = iter*stride + off
    int assemble_only;
};

// Assembled value of a point whose copies do not fit the int4 entry: more than four local copies,
// or partial sums of neighbour ranks (cut points).  A fraction of a percent of the points; the
// corrector forms it before it requests anything else, so that nothing but the point's address is
// live across the spin (at 40 registers anything more spills in the common path).
template <int ORDER>
__device__ __forceinline__ float3 solid_group_sum(const SolidCorrArgs &a, int g) {
    const size_t cs = a.cs;
    const int nloc = a.T.grp[g], nrem = a.T.grp[g + 1];
    float s[3] = {0.f, 0.f, 0.f};
    for (int m = 0; m < nloc; m++) {
        const int ad = a.T.grp[g + 2 + m];
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (!(ORDER == 0 && c == 1)) s[c] = s[c] + a.acc1[ad + cs * c];
    }
    if (nrem > 0) {
        const int2 *recv = recv_slab(a.recv, a.recv_cs, 3, a.recv_parity, a.arrival.dyn);
        const int want = halo_expected(a.arrival);
        for (int m = 0; m < nrem; m++) {
            float r[3];
            halo_read<3, ORDER == 0>(recv, a.recv_cs, a.T.grp[g + 2 + nloc + m], want, a.arrival, r);
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (!(ORDER == 0 && c == 1)) s[c] = s[c] + r[c];
        }
    }
    return make_float3(s[0], s[1], s[2]);
}

// S_B: pdistsum_solid + source + mass inversion + sponge + velocity update.
// Replaces commun.F90:69-171 (+commpi.F90:453-500) and time_evol_wave.F90:466-494 / :689-715.
// MODE (compile time, so that each variant carries only its own loads, stores and registers):
// 0 Newmark, 1 symplectic, 2 lean Newmark, 3 back from lean, 4 assemble only (operator test).
template <int ORDER, int MODE>
__global__ void __launch_bounds__(256, MODE == 0 ? AXB_CORR_MINB - 1 : AXB_CORR_MINB)
k_solid_corrector(const __grid_constant__ SolidCorrArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    const size_t cs = a.cs;
    const int e = p / NPT, q = p - e * NPT, slot = edge_slot(q);
    // the assembly entry first: it heads the only dependent chain (entry -> gather)
    int4 cp = make_int4(-1, -1, -1, -1);
    if (slot >= 0) cp = a.T.cp[16 * (size_t)e + slot];
    if (a.ahead > 0 && p + a.ahead < a.npts) {
        const int pn = p + a.ahead;
        const int en = pn / NPT, sn = edge_slot(pn - en * NPT);
        if (sn >= 0) prefetch_l2(a.T.cp + 16 * (size_t)en + sn);
        if ((threadIdx.x & 7) == 0) {           // one request per 32-byte sector
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (ORDER == 0 && c == 1) continue;
                prefetch_l2(a.acc1 + pn + cs * c);
                if (MODE != 4) prefetch_l2(a.velo + pn + cs * c);
                if (MODE == 0) prefetch_l2(a.acc0 + pn + cs * c);
            }
            if (MODE != 4) prefetch_l2(a.inv_mass_rho + pn);
        }
    }
    float3 grp_sum = make_float3(0.f, 0.f, 0.f);
    if (cp.x == -2) grp_sum = solid_group_sum<ORDER>(a, cp.y);
    // every other input of this point is requested up front (the stores below would otherwise
    // order the per-component loads behind them)
    constexpr bool upd = MODE != 4;
    const bool sponge = upd && a.gamma != nullptr;
    float v[3], vel[3], a0[3], dsp[3];
    const float im = upd ? a.inv_mass_rho[p] : 0.f;
    const float gm = sponge ? a.gamma[p] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        v[c] = vel[c] = a0[c] = dsp[c] = 0.f;
        if (ORDER == 0 && c == 1) continue;
        v[c] = a.acc1[p + cs * c];
        if (upd) {
            vel[c] = a.velo[p + cs * c];
            if (MODE == 0) a0[c] = a.acc0[p + cs * c];
            if (sponge) dsp[c] = a.disp[p + cs * c];
        }
    }
    if (cp.x >= 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (ORDER == 0 && c == 1) continue;
            const float *vec = a.acc1 + cs * c;
            float s = 0.0f;
            s = s + vec[cp.x];
            s = s + vec[cp.y];
            if (cp.z >= 0) s = s + vec[cp.z];
            if (cp.w >= 0) s = s + vec[cp.w];
            v[c] = s;
        }
    } else if (cp.x == -2) {
        v[0] = grp_sum.x; v[1] = grp_sum.y; v[2] = grp_sum.z;
    }
    if (MODE == 4) {
        // op test: stage the assembled field in acc0 (acc1 must stay intact while other
        // threads still pull from it)
#pragma unroll
        for (int c = 0; c < 3; c++) if (!(ORDER == 0 && c == 1)) a.acc0[p + cs * c] = v[c];
        return;
    }
    // add_source_el (time_evol_wave.F90:1082-1097)
    if (a.nelsrc > 0 && e >= a.src_emin && e <= a.src_emax) {
        const int iter = a.dyn ? a.dyn[DYN_ITER] : a.iter;
        const float stf1 = a.stf[(size_t)iter * a.stf_stride + a.stf_off];
        if (stf1 != 0.f) {
            for (int k = 0; k < a.nelsrc; k++)
                if (a.ielsrc[k] - 1 == e) {
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        if (!(ORDER == 0 && c == 1)) v[c] = v[c] - a.src_term[q + NPT * (k + 8 * c)] * stf1;
                }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (ORDER == 0 && c == 1) continue;
        float x = v[c];
        if (MODE != 1) {
            if (ORDER == 1 && c == 2) x = d2f(-2.0 * f2d(im) * f2d(x));
            else x = -im * x;
            if (sponge) x = x - 2 * gm * vel[c] - (gm * gm) * dsp[c];
            if (MODE == 0) {
                a.velo[p + cs * c] = d2f(f2d(vel[c]) + a.half_dt * f2d(a0[c] + x));
                a.acc0[p + cs * c] = x;
            } else if (MODE == 2) {
                // lean Newmark: w = v + dt/2 a is the stored quantity; w_new = w_old + dt a_new
                a.velo[p + cs * c] = d2f(f2d(vel[c]) + a.half_dt * f2d(x));
            } else {
                // back to the reference's state: a_new from the still intact acc1, v = w - dt/2 a
                a.velo[p + cs * c] = d2f(f2d(vel[c]) - a.half_dt * f2d(x));
                a.acc0[p + cs * c] = x;
            }
        } else {
            x = -im * x;
            if (sponge) x = x - 2 * gm * vel[c] - (gm * gm) * dsp[c];
            if (ORDER == 1 && c == 2) a.velo[p + cs * c] = d2f(f2d(vel[c]) + 2.0 * f2d(x) * a.half_dt);
            else a.velo[p + cs * c] = d2f(f2d(vel[c]) + f2d(x) * a.half_dt);
            a.acc0[p + cs * c] = x;     // keeps the reference's `acc` available to get_state
        }
    }
}

// final drift of a symplectic step (time_evol_wave.F90:720-725)
__global__ void k_drift(int n, float *x, const float *v, double cd) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) x[p] = d2f(f2d(x[p]) + f2d(v[p]) * cd);
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
    int2 *dst_base[8];        // per message: peer slab base (parity 0), words {value, exchange number}
    size_t dst_cs[8];         // per message: component stride of the peer slab (= its slot count)
    int parity;               // which half of the peers' slabs this exchange writes
    int value;                // exchange number the words carry
    const int *dyn;           // graph replay: parity = dyn[DYN_SEQ] & 1, value = dyn[DYN_SEQ] + 1
};
// feed_buffer + ISEND (commpi.F90:371-452): one thread per (send entry, component) sums the local
// copies of the cut point and stores {sum, exchange number} into the neighbour's slab
__global__ void k_halo_pack(const __grid_constant__ PackArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nentries * a.nc) return;
    const int seq = a.dyn ? a.dyn[DYN_SEQ] : 0;
    const int parity = a.dyn ? (seq & 1) : a.parity;
    const int value = a.dyn ? seq + 1 : a.value;
    const int en = t % a.nentries, c = t / a.nentries;
    float s = 0.0f;
    for (int m = a.start[en]; m < a.start[en + 1]; m++) s = s + a.vec[a.addr[m] + a.cs * c];
    const int msg = a.dst_msg[en];
    st_volatile_v2(a.dst_base[msg] + (size_t)parity * a.nc * a.dst_cs[msg] + a.dst_slot[en] + a.dst_cs[msg] * c,
                   __float_as_int(s), value);
}

// ---------------------------------------------------------------------------------------
struct RecArgs {
    int num_rec, order, seis_it, nseismo_max;
    const int *recfile_el;    // (num_rec,3)
    const float *disp; size_t cs;
    float *recdump;           // (3, num_rec, nseismo_max)
    int iter, iseismo;        // host-tracked: time step just completed, next sample slot
    int *dyn;                 // graph replay (k_step_end): the step being completed is dyn[DYN_ITER] + 1
};
__device__ __forceinline__ void sample_receiver(const RecArgs &a, int r, int is) {
    const int iel = a.recfile_el[r], ip = a.recfile_el[r + a.num_rec], jp = a.recfile_el[r + 2 * a.num_rec];
    const size_t p = ip + NP * jp + (size_t)NPT * (iel - 1);
    const float d1 = a.disp[p], d2 = a.disp[p + a.cs], d3 = a.disp[p + 2 * a.cs];
    float *out = a.recdump + (size_t)3 * a.num_rec * is + 3 * r;
    if (a.order == 0) { out[0] = d1; out[1] = 0.f; out[2] = d3; }
    else if (a.order == 1) { out[0] = d1 + d2; out[1] = d1 - d2; out[2] = d3; }
    else { out[0] = d1; out[1] = d2; out[2] = d3; }
}
// nc_compute_recfile_seis_bare (seismograms.f90:783-820) every seis_it steps
__global__ void k_sample_receivers(const __grid_constant__ RecArgs a) {
    const int iter = a.iter;
    if (iter % a.seis_it != 0) return;
    const int is = a.iseismo;
    if (is >= a.nseismo_max) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.num_rec) return;
    sample_receiver(a, r, is);
}
// Last node of a graph-replayed Newmark step (one block): receiver sampling of the step just
// completed, then the device-resident step counters move on.
__global__ void k_step_end(const __grid_constant__ RecArgs a) {
    const int iter = a.dyn[DYN_ITER] + 1, is = a.dyn[DYN_ISEISMO];
    const bool sample = a.num_rec > 0 && iter % a.seis_it == 0 && is < a.nseismo_max;
    if (sample)
        for (int r = threadIdx.x; r < a.num_rec; r += blockDim.x) sample_receiver(a, r, is);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.dyn[DYN_ITER] = iter;
        a.dyn[DYN_SEQ] = a.dyn[DYN_SEQ] + 1;
        if (sample) a.dyn[DYN_ISEISMO] = is + 1;
    }
}
__global__ void k_set_dyn(int *dyn, int iter, int seq, int iseismo) {
    dyn[DYN_ITER] = iter; dyn[DYN_SEQ] = seq; dyn[DYN_ISEISMO] = iseismo;
}
// axb_set_stf_values for a handful of samples: the values travel as kernel arguments, so the
// caller's buffer is free again when the call returns
struct StfPatch { int first, n; float v[32]; };
__global__ void k_patch_stf(float *stf, const __grid_constant__ StfPatch a) {
    if (threadIdx.x < a.n) stf[a.first + threadIdx.x] = a.v[threadIdx.x];
}
// classic <-> lean Newmark state of the fluid: dchi +/- dt/2 ddchi0
__global__ void k_axpy(int n, float *x, const float *y, double c) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) x[p] = d2f(f2d(x[p]) + c * f2d(y[p]));
}

struct DumpArgs {
    int nel_s, nel_f, order, strain_it, nstrain_max;
    const int *kwf_mask, *kwf_map;
    const float *disp, *chi; size_t cs;
    const int *axis_f;
    const float *inv_rho, *Dse, *Dze, *Dsx, *Dzx;
    float *snap; size_t npts;
    int iter, istrain;        // host-tracked: time step just completed, next snapshot slot
};
// dump_disp_global, solid part (wavefields_io.f90:1041-1052, 743-762)
__global__ void k_dump_solid(const __grid_constant__ DumpArgs a) {
    const int iter = a.iter;
    if (iter % a.strain_it != 0) return;
    const int is = a.istrain;
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
    const int iter = a.iter;
    if (iter % a.strain_it != 0) return;
    const int is = a.istrain;
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

// runtime_info (time_evol_wave.F90:998-1056), the part that matters on the device: the run is
// declared blown up when max |disp(1,1,:,:)| exceeds 10 |magnitude| (:1042) or is not finite.
// counters[3] keeps the first offending iteration (0 = fine).
__global__ void k_blowup_check(const float *disp, size_t cs, int nel, float thresh, int iter, int *counters) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    const size_t p = (size_t)NPT * e + NP + 1;          // point (ipol, jpol) = (1, 1)
    // written so that NaN counts as blown up (fmaxf would drop it)
    const bool ok = fabsf(disp[p]) <= thresh && fabsf(disp[p + cs]) <= thresh && fabsf(disp[p + 2 * cs]) <= thresh;
    if (!ok) atomicCAS(&counters[3], 0, iter > 0 ? iter : 1);
}

// ---------------------------------------------------------------------------------------
// energy (time_evol_wave.F90:1424-1526): the four sums of one sample, accumulated in real(8)
// (the reference's sum() runs over real(4) arrays in an order the compiler chooses).
struct EnergyArgs {
    int npts; size_t cs; int order;
    const float *stiff, *u, *v, *mass;     // solid: K u (masked), disp, velo, unassem_mass_rho_solid
    const int *axis;                       // fluid: K dchi (masked), dchi, ddchi, unassem_mass_lam_fluid
    double *out;                           // [0] potential-type sum, [1] kinetic-type sum of this domain
};
__device__ __forceinline__ void energy_block_sum(double a, double b, double *out) {
    __shared__ double sa[8], sb[8];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(FULL, a, o); b += __shfl_down_sync(FULL, b, o); }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) { a += sa[k]; b += sb[k]; }
        atomicAdd(out, a);
        atomicAdd(out + 1, b);
    }
}
__global__ void __launch_bounds__(256) k_energy_solid(const __grid_constant__ EnergyArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double epot = 0.0, ekin = 0.0;
    if (p < a.npts) {
        const int e = p / NPT, i = (p - e * NPT) % NP;
        const bool ax0 = a.axis[e] != 0 && i == 0;
        const float m = a.mass[p];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            // monopole: stiff(:,:,:,2) stays zero, vel(:,:,:,2) is summed like the other two
            if (!(a.order == 0 && c == 1)) {
                const bool masked = ax0 && (a.order == 0 ? c == 0 : (a.order == 1 ? c != 0 : true));
                const float d = masked ? 0.f : a.u[p + a.cs * c];
                epot += (double)(a.stiff[p + a.cs * c] * d);         // stiff = stiff * disp
            }
            const float v = a.v[p + a.cs * c];
            float x = v * v * m;                                      // vel**2 * unassem_mass_rho_solid
            if (a.order == 1 && c == 2) x = 2.0f * x;
            ekin += (double)x;
        }
    }
    energy_block_sum(epot, ekin, a.out);
}
__global__ void __launch_bounds__(256) k_energy_fluid(const __grid_constant__ EnergyArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double epot = 0.0, ekin = 0.0;
    if (p < a.npts) {
        const int e = p / NPT, i = (p - e * NPT) % NP;
        const bool masked = a.order != 0 && a.axis[e] != 0 && i == 0;
        const float dd = a.v[p];
        epot = (double)(dd * dd * a.mass[p]);                        // ddchi**2 * unassem_mass_lam_fluid
        const float d = masked ? 0.f : a.u[p];
        ekin = (double)(a.stiff[p] * d);                             // stiff_flu * dchi
    }
    energy_block_sum(epot, ekin, a.out);
}

}  // namespace axb
#include "axb_dump_fields.cuh"
