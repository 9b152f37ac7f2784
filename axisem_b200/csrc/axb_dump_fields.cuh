// axb_dump_fields.cuh — wavefield dumps of dump_type strain_only / fullfields (included by
// axb_kernels.cuh).
//
// Replaces compute_strain (time_evol_wave.F90:1264-1410: the six strain components Ekk, E11, E13,
// E22, E12, E23 of both domains, dipole fields taken to the (s, phi, z) system) and, for
// fullfields, dump_velo_global (wavefields_io.f90:932-1015), with the point selection of
// dump_field_1d (:787-836): the kwf mapping for strain_only, the packed block
// ibeg:iend x jbeg:jend of every element for fullfields.  Runs every strain_it steps only, so the
// mapping is the simple one: a warp per element, lane q = point (i, j), the 5x5 contractions of
// pointwise_derivatives.f90:329-365 by shuffles (k ascending, as mxm does).
//
// dump_type 3 = the xdmf snapshots of glob_snapshot_xdmf (wavefields_io.f90:119-203): u_s, u_p,
// u_z, calc_straintrace (:630-686) and calc_curlinplane (:601-627) at the plot points of
// dump_xdmf_grid (meshes_io.F90:110-437; `xmap` = mapping_ijel_iplot where plotting_mask is set,
// per element-local point, fluid elements first).
#pragma once

namespace axb {

struct FieldDumpArgs {
    int nel_s, nel_f, order, dump_type;      // dump_type 1 strain_only, 2 fullfields, 3 xdmf
    const int *xmap;                         // xdmf: plot point (1-based) of (q, element), 0 = none
    int ibeg, iend, jbeg, jend;
    int nstrain_max, istrain;
    const int *kwf_mask, *kwf_map;
    const int *axis_s, *axis_f;
    const float *disp, *velo; size_t cs;
    const float *chi, *dchi;
    const float *Dse, *Dze, *Dsx, *Dzx, *inv_s;            // solid, (5,5,nel_solid)
    const float *Dse_f, *Dze_f, *Dsx_f, *Dzx_f, *inv_s_f;  // fluid
    const float *inv_rho;
    float *snap; size_t npts;
};

struct Grad2 { float ds, dz; };
// axisym_gradient_{solid,fluid} at this lane's point; every lane of the warp takes part
__device__ __forceinline__ Grad2 lane_gradient(float f, bool ax, const LaneG &L, int i, int j5,
                                               float dse, float dze, float dsx, float dzx) {
    const float m1 = ax ? contract_xi(f, L.g1t_row, j5) : contract_xi(f, L.g2t_row, j5);
    const float m2 = contract_eta(f, L.g2_col, i);
    Grad2 g;
    g.ds = dze * m1 + dzx * m2;
    g.dz = dse * m1 + dsx * m2;
    return g;
}
// f_over_s_{solid,fluid}: f / s, on the axis (i = 0 of an axial element) the s-derivative
__device__ __forceinline__ float lane_f_over_s(float f, bool ax, const LaneG &L, int i, int j5,
                                               float inv_s, float dze, float dzx) {
    const float m1 = contract_xi(f, L.g1t_row, j5);
    const float m2 = contract_eta(f, L.g2_col, i);
    const float fs = inv_s * f;
    return (ax && i == 0) ? dze * m1 + dzx * m2 : fs;
}
__device__ __forceinline__ long dump_slot(const FieldDumpArgs &a, bool fluid, int e, int q) {
    if (a.dump_type == 3) return (long)a.xmap[q + (size_t)NPT * ((size_t)e + (fluid ? 0 : a.nel_f))] - 1;
    if (a.dump_type == 1) {
        const size_t pk = q + (size_t)NPT * ((size_t)e + (fluid ? a.nel_s : 0));
        return a.kwf_mask[pk] ? (long)a.kwf_map[pk] - 1 : -1;
    }
    const int i = q % NP, j = q / NP, ni = a.iend - a.ibeg + 1, nj = a.jend - a.jbeg + 1;
    if (i < a.ibeg || i > a.iend || j < a.jbeg || j > a.jend) return -1;
    const long base = fluid ? (long)ni * nj * a.nel_s : 0;
    return base + ((long)e * nj + (j - a.jbeg)) * ni + (i - a.ibeg);
}

__global__ void __launch_bounds__(256)
k_dump_fields_solid(const __grid_constant__ GMat G, const __grid_constant__ FieldDumpArgs a) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const bool active = lane < NPT;
    const int q = active ? lane : 0, i = q % NP, j = q / NP, j5 = 5 * j;
    __shared__ GMat sG;
    stage_g(G, sG);
    LaneG L;
    load_lane_g(sG, i, j, L);
    const bool mono = a.order == 0, di = a.order == 1, full = a.dump_type == 2;
    const int V_TR = mono ? 3 : 5, V_VS = mono ? 4 : 6, V_VZ = mono ? 5 : 8;
    const size_t vs = a.npts * a.nstrain_max;
    float *base = a.snap + a.npts * a.istrain;
    const float two = 2.0f;
    for (int e = blockIdx.x * wpb + wib; e < a.nel_s; e += gridDim.x * wpb) {
        const size_t p = (size_t)NPT * e + q;
        const bool ax = a.axis_s[e] != 0;
        const float u1 = active ? a.disp[p] : 0.f, u2 = active ? a.disp[p + a.cs] : 0.f;
        const float u3 = active ? a.disp[p + 2 * a.cs] : 0.f;
        const float dse = a.Dse[p], dze = a.Dze[p], dsx = a.Dsx[p], dzx = a.Dzx[p], is = a.inv_s[p];
        float E_dsus, E_dsuz, E_dpup, E_dsup = 0.f, E_dzup = 0.f, E_tr;
        Grad2 g = lane_gradient(di ? u1 + u2 : u1, ax, L, i, j5, dse, dze, dsx, dzx);
        E_dsus = g.ds;
        const Grad2 h3 = lane_gradient(u3, ax, L, i, j5, dse, dze, dsx, dzx);
        // axisym_gradient_solid_add: grad(1) = old(2) + dsdf, grad(2) = old(1) + dzdf
        float g1 = g.dz + h3.ds;
        const float g2 = g.ds + h3.dz;
        g1 = g1 / two;
        E_dsuz = g1;
        if (mono) {
            const float buff = lane_f_over_s(u1, ax, L, i, j5, is, dze, dzx);
            E_dpup = buff; E_tr = buff + g2;
        } else if (di) {
            const float fs = lane_f_over_s(u2, ax, L, i, j5, is, dze, dzx);
            const float buff = two * fs;
            E_dpup = buff; E_tr = buff + g2;
            const Grad2 hm = lane_gradient(u1 - u2, ax, L, i, j5, dse, dze, dsx, dzx);
            const float fs3 = lane_f_over_s(u3, ax, L, i, j5, is, dze, dzx);
            E_dsup = -fs - hm.ds / two;
            E_dzup = -(fs3 + hm.dz) / two;
        } else {
            const float buff = lane_f_over_s(u1 - two * u2, ax, L, i, j5, is, dze, dzx);
            E_dpup = buff; E_tr = buff + g2;
            const Grad2 hp = lane_gradient(u2, ax, L, i, j5, dse, dze, dsx, dzx);
            const float fs = lane_f_over_s(u1 + u2 / two, ax, L, i, j5, is, dze, dzx);
            const float fs3 = lane_f_over_s(u3, ax, L, i, j5, is, dze, dzx);
            E_dsup = -fs - hp.ds / two;
            E_dzup = -fs3 - hp.dz / two;
        }
        if (!active) continue;
        const long ct = dump_slot(a, false, e, q);
        if (ct < 0) continue;
        if (a.dump_type == 3) {
            base[ct] = di ? u1 + u2 : u1; base[ct + vs] = di ? u1 - u2 : u2; base[ct + 2 * vs] = u3;
            base[ct + 3 * vs] = E_tr;
            base[ct + 4 * vs] = g.dz - h3.ds;        // d_z u_s - d_s u_z
            continue;
        }
        base[ct] = E_dsus; base[ct + vs] = E_dsuz; base[ct + 2 * vs] = E_dpup;
        if (!mono) { base[ct + 3 * vs] = E_dsup; base[ct + 4 * vs] = E_dzup; }
        base[ct + V_TR * vs] = E_tr;
        if (full) {
            const float v1 = a.velo[p], v2 = a.velo[p + a.cs], v3 = a.velo[p + 2 * a.cs];
            if (di) { base[ct + V_VS * vs] = v1 + v2; base[ct + 7 * vs] = v1 - v2; }
            else { base[ct + V_VS * vs] = v1; if (!mono) base[ct + 7 * vs] = v2; }
            base[ct + V_VZ * vs] = v3;
        }
    }
}

__global__ void __launch_bounds__(256)
k_dump_fields_fluid(const __grid_constant__ GMat G, const __grid_constant__ FieldDumpArgs a) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const bool active = lane < NPT;
    const int q = active ? lane : 0, i = q % NP, j = q / NP, j5 = 5 * j;
    __shared__ GMat sG;
    stage_g(G, sG);
    LaneG L;
    load_lane_g(sG, i, j, L);
    const bool mono = a.order == 0, di = a.order == 1, full = a.dump_type == 2;
    const int V_TR = mono ? 3 : 5, V_VS = mono ? 4 : 6, V_VZ = mono ? 5 : 8;
    const size_t vs = a.npts * a.nstrain_max;
    float *base = a.snap + a.npts * a.istrain;
    const float two = 2.0f;
    for (int e = blockIdx.x * wpb + wib; e < a.nel_f; e += gridDim.x * wpb) {
        const size_t p = (size_t)NPT * e + q;
        const bool ax = a.axis_f[e] != 0;
        const float dse = a.Dse_f[p], dze = a.Dze_f[p], dsx = a.Dsx_f[p], dzx = a.Dzx_f[p];
        const float is = a.inv_s_f[p], ir = a.inv_rho[p];
        // displacement in the fluid: 1/rho grad(chi)
        const Grad2 gc = lane_gradient(active ? a.chi[p] : 0.f, ax, L, i, j5, dse, dze, dsx, dzx);
        const float us = active ? gc.ds * ir : 0.f, uz = active ? gc.dz * ir : 0.f;
        const Grad2 g = lane_gradient(us, ax, L, i, j5, dse, dze, dsx, dzx);
        const float E_dsus = g.ds;
        const Grad2 hz = lane_gradient(uz, ax, L, i, j5, dse, dze, dsx, dzx);
        float g1 = g.dz + hz.ds;
        const float g2 = g.ds + hz.dz;
        g1 = g1 / two;
        const float fs = lane_f_over_s(us, ax, L, i, j5, is, dze, dzx);
        const float fz = lane_f_over_s(uz, ax, L, i, j5, is, dze, dzx);
        Grad2 w;
        w.ds = w.dz = 0.f;
        float ws_shift = 0.f;
        if (full) {
            w = lane_gradient(active ? a.dchi[p] : 0.f, ax, L, i, j5, dse, dze, dsx, dzx);
            // wavefields_io.f90:992-993 takes the s component at first index jbeg:jend where it
            // means ibeg:iend: the value stored for point i is that of i - ibeg + jbeg
            const int iq = min(max(i - a.ibeg + a.jbeg, 0), NP - 1);
            ws_shift = shfl(w.ds, iq + j5);
        }
        if (!active) continue;
        const long ct = dump_slot(a, true, e, q);
        if (ct < 0) continue;
        if (a.dump_type == 3) {
            base[ct] = us; base[ct + vs] = 0.f; base[ct + 2 * vs] = uz;
            base[ct + 3 * vs] = fs + g2; base[ct + 4 * vs] = 0.f;
            continue;
        }
        base[ct] = E_dsus; base[ct + vs] = g1; base[ct + 2 * vs] = fs;
        if (!mono) {
            base[ct + 3 * vs] = di ? (-fs) / two : -fs;
            base[ct + 4 * vs] = di ? fz / two : -fz;
        }
        base[ct + V_TR * vs] = fs + g2;
        if (full) {
            base[ct + V_VS * vs] = ir * ws_shift;
            if (!mono) base[ct + 7 * vs] = 0.f;
            base[ct + V_VZ * vs] = ir * w.dz;
        }
    }
}

}  // namespace axb
