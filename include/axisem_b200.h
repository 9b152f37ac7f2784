/*
 * axisem_b200.h — C ABI of the B200-native AxiSEM SOLVER time loop.
 *
 * Drop-in seam: the single `call time_loop` of the reference
 * (SOLVER/main.f90:92 -> SOLVER/time_evol_wave.F90:231-245).  Everything above that call
 * (mesh ingest, pre-computed terms, source/receiver set-up) stays on the host and hands
 * its module arrays to this library once; `axb_run` then replaces
 * `sf_time_loop_newmark` (time_evol_wave.F90:264-502) / `symplectic_time_loop`
 * (:516-741) with device-resident stepping.  INTEGRATION.md shows the Fortran
 * `iso_c_binding` interface a maintainer would add.
 *
 * Conventions (identical to the reference's own C interop, SOLVER/nc_routines.F90:290-300,
 * SOLVER/pthread.c:44-57):
 *   - all arrays are caller-owned HOST memory, contiguous, Fortran (column-major) order,
 *     borrowed for the duration of the call and copied to the device — never retained;
 *   - `float` = real(kind=realkind)=sp (global_parameters.f90:40), `double` = dp,
 *     `int32_t` = default integer, logicals are 4-byte integers (0 = .false.);
 *   - every index stored in an integer map is 1-based, exactly as the Fortran holds it;
 *     pol indices (ipol, jpol) are 0-based as in the reference (`0:npol`);
 *   - element-local linear point index: ipt = (iel-1)*25 + jpol*5 + ipol + 1
 *     (SOLVER/commun.F90:303);
 *   - every function returns 0 on success; on failure a non-zero code is returned and
 *     `axb_last_error()` describes it (the reference would `stop`).
 *   - npol must be 4 (the `_4` routines of the reference).
 *
 * The same header is implemented twice with different prefixes:
 *   libaxisem_b200.so      axb_*   CUDA sm_100a (the product)
 *   oracle/libaxisem_oracle.so  axo_*   CPU restatement of the reference (test oracle)
 * AXB_PREFIX selects the prefix; signatures are identical.
 */
#ifndef AXISEM_B200_H
#define AXISEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef AXB_PREFIX
#define AXB_PREFIX axb_
#endif
#define AXB_CAT_(a, b) a##b
#define AXB_CAT(a, b) AXB_CAT_(a, b)
#define AXB(name) AXB_CAT(AXB_PREFIX, name)

typedef struct axb_handle_s *axb_handle;

/* src_type(1): monopole / dipole / quadpole (SOLVER/data_source.f90:38) */
enum { AXB_MONOPOLE = 0, AXB_DIPOLE = 1, AXB_QUADPOLE = 2 };

/* time_scheme (SOLVER/time_evol_wave.F90:237, :760-941) */
enum {
    AXB_NEWMARK2 = 0, AXB_SYMPLEC4 = 1, AXB_ML_SO4M5 = 2, AXB_ML_SO6M7 = 3,
    AXB_KL_O8M17 = 4, AXB_SS_35O10 = 5
};

/* stf_type for the point-wise STF of the symplectic schemes: every case of compute_stf_t
 * (SOLVER/source.f90:206-233; gauss_t :818, gauss_d_t :835, gauss_dd_t :853, errorf_t :873,
 * delta_src_t :890, quasiheavi_t :908) */
enum {
    AXB_STF_GAUSS_0 = 0, AXB_STF_GAUSS_1 = 1, AXB_STF_GAUSS_2 = 2,
    AXB_STF_ERRORF = 3, AXB_STF_DIRAC_0 = 4, AXB_STF_QUHEAVI = 5
};

enum { AXB_DOMAIN_SOLID = 0, AXB_DOMAIN_FLUID = 1 };

/* state arrays for axb_get_state / axb_set_state */
enum {
    AXB_F_DISP = 0, AXB_F_VELO = 1, AXB_F_ACC0 = 2, AXB_F_ACC1 = 3,   /* (5,5,nel_solid,3) */
    AXB_F_CHI = 4, AXB_F_DCHI = 5, AXB_F_DDCHI0 = 6, AXB_F_DDCHI1 = 7, /* (5,5,nel_fluid)   */
    AXB_F_MEMVAR = 8,       /* (4,6,n_sls,nel_solid) cg4 or (5,5,6,n_sls,nel_solid)          */
    AXB_F_SRC_DEV_TM1 = 9,  /* (4,6,nel_solid) or (5,5,6,nel_solid)                          */
    AXB_F_SRC_TR_TM1 = 10   /* (4,nel_solid)   or (5,5,nel_solid)                            */
};

/* single operators, for per-routine parity tests (each cites the routine it replaces) */
enum {
    AXB_OP_SOLID_STIFFNESS = 0,  /* acc1 = K(disp): glob_stiffness_{mono,di,quad}_4           */
    AXB_OP_ANEL_STIFFNESS = 1,   /* acc1 -= anelastic term: glob_anel_stiffness_*             */
    AXB_OP_FLUID_STIFFNESS = 2,  /* ddchi1 = K(chi): glob_fluid_stiffness_4                   */
    AXB_OP_PDISTSUM_SOLID = 3,   /* pdistsum_solid(acc1)  (commun.F90:69-171), all phases     */
    AXB_OP_PDISTSUM_FLUID = 4,   /* pdistsum_fluid(ddchi1) (commun.F90:180-283)               */
    AXB_OP_MEMVARS = 5,          /* time_step_memvars(memvar, disp) (attenuation.f90:55-334)  */
    AXB_OP_BDRY2FLUID = 6,       /* bdry_copy2fluid(ddchi1, disp) (time_evol_wave.F90:1532)   */
    AXB_OP_BDRY2SOLID = 7        /* bdry_copy2solid(acc1, ddchi1) (time_evol_wave.F90:1577)   */
};

/* Solid pre-computed planes, SOLVER/data_matr.f90:46-76.  Each non-NULL pointer is a
 * (0:4,0:4,nel_solid) array; M0_w* are (0:4,nel_solid).  Planes not allocated for the
 * source order (def_precomp_terms.f90:1216-1284) are NULL. */
typedef struct {
    const float *M11s, *M21s, *M41s, *M12s, *M22s, *M32s, *M42s, *M11z, *M21z, *M41z;
    const float *M13s, *M33s, *M43s;                 /* dipole */
    const float *M1phi, *M2phi, *M4phi;              /* quadrupole */
    const float *M_1, *M_2, *M_3, *M_4, *M_5, *M_6, *M_7, *M_8;
    const float *M_w1, *M_w2, *M_w3, *M_w4, *M_w5;
    const float *M0_w1, *M0_w2, *M0_w3, *M0_w4, *M0_w5, *M0_w6, *M0_w7, *M0_w8, *M0_w9, *M0_w10;
} axb_solid_terms;

/* Attenuation inputs: SOLVER/attenuation.f90:38-49, data_matr.f90:95-112,
 * data_pointwise.f90.  Coarse-grained (`att_coarse_grained`, the default) uses the *_cg4
 * members (1:4,nel_solid); otherwise the full (0:4,0:4,nel_solid) planes, the axial
 * (0:4,nel_solid) vectors and inv_s_solid are required.  The SLS fit itself is host
 * work (attenuation.f90:1183-1339, unseeded RNG); only its results are passed. */
typedef struct {
    int32_t coarse_grained;
    int32_t n_sls;
    int32_t do_corr_lowq;
    const double *y_j;              /* (n_sls) */
    const double *exp_w_j_deltat;   /* (n_sls) */
    const double *ts_fac_t;         /* (n_sls) */
    const double *ts_fac_tm1;       /* (n_sls) */
    const float *Q_mu, *Q_kappa;    /* (nel_solid) */
    const float *delta_mu_cg4, *delta_kappa_cg4;
    const float *Y_cg4, *V_s_eta_cg4, *V_s_xi_cg4, *V_z_eta_cg4, *V_z_xi_cg4;
    const float *DsDeta_over_J_sol_cg4, *DzDeta_over_J_sol_cg4;
    const float *DsDxi_over_J_sol_cg4, *DzDxi_over_J_sol_cg4;
    const float *delta_mu, *delta_kappa;
    const float *Y, *V_s_eta, *V_s_xi, *V_z_eta, *V_z_xi;
    const float *Y0, *V0_s_eta, *V0_s_xi, *V0_z_eta, *V0_z_xi;
    const float *DsDeta_over_J_sol, *DzDeta_over_J_sol, *DsDxi_over_J_sol, *DzDxi_over_J_sol;
    const float *inv_s_solid;       /* (0:4,0:4,nel_solid); needed by both variants */
} axb_attenuation;

const char *AXB(last_error)(void);

/* mynum, nproc: SOLVER/data_proc.f90.  device = CUDA ordinal (ignored by the oracle). */
int AXB(create)(axb_handle *h, int32_t device, int32_t rank, int32_t nranks);
int AXB(destroy)(axb_handle h);

/* data_mesh.f90:58-66 (igloc_*, nglob_*), def_grid.f90:59-77 (axis flags, ax_el_*),
 * data_spec.f90:38-39 (G0(0:4), G1,G1T,G2,G2T(0:4,0:4)). */
int AXB(set_mesh)(axb_handle h, int32_t npol, int32_t nel_solid, int32_t nel_fluid,
                  int32_t nglob_solid, int32_t nglob_fluid,
                  const int32_t *igloc_solid, const int32_t *igloc_fluid,
                  const int32_t *axis_solid, const int32_t *axis_fluid,
                  const int32_t *ax_el_solid, int32_t naxel_solid,
                  const int32_t *ax_el_fluid, int32_t naxel_fluid,
                  const float *G0, const float *G1, const float *G1T,
                  const float *G2, const float *G2T);

int AXB(set_solid_terms)(axb_handle h, int32_t src_order, const axb_solid_terms *t);

/* data_matr.f90:79-82, :37, data_mesh.f90:181.  M_w_fl / M0_w_fl may be NULL for
 * monopole sources; fluid_free_surface_mask may be NULL (= all ones). */
int AXB(set_fluid_terms)(axb_handle h, const float *M1chi_fl, const float *M2chi_fl,
                         const float *M4chi_fl, const float *M_w_fl, const float *M0_w_fl,
                         const float *inv_mass_fluid, const float *fluid_free_surface_mask);

/* inv_mass_rho(0:4,0:4,nel_solid), data_matr.f90:36 (dipole: 1/2 folded in). */
int AXB(set_mass)(axb_handle h, const float *inv_mass_rho);

/* {solid,fluid}_absorbing_gamma, data_mesh.f90:185-186; NULL,NULL = have_absorbing_bc false */
int AXB(set_sponge)(axb_handle h, const float *solid_gamma, const float *fluid_gamma);

/* data_mesh.f90:106-110, data_matr.f90:88: bdry_matr(0:4,nel_bdry,2). */
int AXB(set_sf_boundary)(axb_handle h, int32_t nel_bdry, const int32_t *bdry_solid_el,
                         const int32_t *bdry_fluid_el, const int32_t *bdry_jpol_solid,
                         const int32_t *bdry_jpol_fluid, const float *bdry_matr);

int AXB(set_attenuation)(axb_handle h, const axb_attenuation *a);

/* data_source.f90:38,50-52: source_term_el(0:4,0:4,8,3), ielsrc(8), stf(niter).
 * fluid_src != 0 passes source_term_fl(0:4,0:4,8) instead. */
int AXB(set_source)(axb_handle h, int32_t fluid_src, int32_t nelsrc, const int32_t *ielsrc,
                    const float *source_term, const float *stf, int32_t niter);

/* overwrite stf(first_iter+1 : first_iter+n) (0-based first_iter) — lets a host stream the
 * source time function while the loop runs (Newmark only) */
int AXB(set_stf_values)(axb_handle h, int32_t first_iter, int32_t n, const float *values);

/* point-wise STF for the symplectic schemes, compute_stf_t (source.f90:206-233) */
int AXB(set_stf_params)(axb_handle h, int32_t stf_type, double decay, double t_0,
                        double shift_fact, double magnitude);

/* out(nstages, n): stf_symp of steps first_iter+1 .. first_iter+n (0-based first_iter) as the
 * symplectic loop applies it, real(stf_symp(i), kind=realkind) (time_evol_wave.F90:592-593, :689);
 * valid after axb_finalize_setup of a symplectic scheme */
int AXB(get_stf_symp)(axb_handle h, int32_t first_iter, int32_t n, float *out);

/* recfile_el(num_rec,3) = (iel, ipol, jpol), data_mesh.f90:138 */
int AXB(set_receivers)(axb_handle h, int32_t num_rec, const int32_t *recfile_el);

/* displ_only wavefield dump (wavefields_io.f90:743-783, 1019-1115; meshes_io.F90:489-640):
 * kwf_mask, mapping_ijel_ikwf are (0:4,0:4,nel_solid+nel_fluid); the fluid pointwise
 * planes (data_pointwise.f90) and inv_rho_fluid are (0:4,0:4,nel_fluid). */
int AXB(set_kwf)(axb_handle h, const int32_t *kwf_mask, const int32_t *mapping_ijel_ikwf,
                 int32_t npoint_solid_kwf, int32_t npoint_fluid_kwf,
                 const float *inv_rho_fluid, const float *DsDeta_over_J_flu,
                 const float *DzDeta_over_J_flu, const float *DsDxi_over_J_flu,
                 const float *DzDxi_over_J_flu);

/* dump_type of the wavefield dumps (data_io.f90; SOLVER/inparam_advanced KERNEL_DUMPTYPE):
 * displ_only (default; dump_disp_global), strain_only and fullfields.  The latter two replace
 * compute_strain (time_evol_wave.F90:1264-1410) and, for fullfields, dump_velo_global
 * (wavefields_io.f90:932-1015), evaluated on the device every strain_it steps:
 *   strain_only: 6 (monopole: 4) fields on the kwf point set of axb_set_kwf (kwf_mapping_sol/flu,
 *                wavefields_io.f90:743-783);
 *   fullfields:  the same fields + 3 (2) velocity fields on the block ibeg:iend x jbeg:jend of
 *                every element, packed in Fortran order (i, j, element), solid then fluid.
 * Needs the pointwise-derivative planes of the solid (data_pointwise: D*D*_over_J_sol), inv_s_solid
 * and inv_s_fluid, all (0:4,0:4,nel); the fluid planes and inv_rho_fluid come from axb_set_kwf,
 * which must be called too.  Before axb_finalize_setup.  Not restated: the zeroing of the
 * source elements for src_dump_type == 'mask' — the reference never assigns that variable. */
enum { AXB_DUMP_DISPL_ONLY = 0, AXB_DUMP_STRAIN_ONLY = 1, AXB_DUMP_FULLFIELDS = 2 };
int AXB(set_dump)(axb_handle h, int32_t dump_type, int32_t ibeg, int32_t iend, int32_t jbeg, int32_t jend,
                  const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                  const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol,
                  const float *inv_s_solid, const float *inv_s_fluid);
/* shape of the snapshot buffer: npoints (npts_sol + npts_flu) and nvars = nvar/2 of
 * nc_routines.F90:943-1050, in that order: displ_only (s, p, z); strain_only (strain_dsus, _dsuz,
 * _dpup, [_dsup, _dzup,] straintrace); fullfields: the same followed by velo_s, [velo_p,] velo_z */
int AXB(snapshot_layout)(axb_handle h, int32_t *npoints, int32_t *nvars);

/* XDMF snapshots (data_io.f90:36, 67-70: dump_xdmf, i_arr_xdmf, j_arr_xdmf; inparam_advanced
 * XDMF_GLL_I / XDMF_GLL_J / XDMF_RMIN.. XDMF_COLAT_MAX).  The plot-point maps are the host's
 * (dump_xdmf_grid, meshes_io.F90:110-437): plotting_mask and mapping_ijel_iplot are
 * (i_n_xdmf, j_n_xdmf, nelem) Fortran order with the FLUID elements first (iel = 1..nel_fluid,
 * then nel_fluid + 1..nelem), mapping 1-based into the npoint_plot plot points; i_arr_xdmf /
 * j_arr_xdmf hold the GLL indices (0..npol) of the rows / columns that are plotted.  Every
 * snap_it steps, the first time at iter 0 (dump_stuff, time_evol_wave.F90:1167-1176), the loop
 * forms what glob_snapshot_xdmf (wavefields_io.f90:119-203) hands to its writers: the
 * displacement in (s, phi, z) — 1/rho grad(chi) in the fluid — calc_straintrace (:630-686) and
 * calc_curlinplane (:601-627).  The derivative planes are those of data_pointwise
 * (DsDeta_over_J_sol, ..., inv_s_solid; *_flu, inv_s_fluid, inv_rho_fluid); fluid planes may be
 * NULL when nel_fluid = 0.  Call before axb_finalize_setup. */
int AXB(set_xdmf)(axb_handle h, int32_t snap_it, int32_t i_n_xdmf, int32_t j_n_xdmf,
                  const int32_t *i_arr_xdmf, const int32_t *j_arr_xdmf,
                  const int32_t *plotting_mask, const int32_t *mapping_ijel_iplot, int32_t npoint_plot,
                  const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                  const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol, const float *inv_s_solid,
                  const float *DsDeta_over_J_flu, const float *DzDeta_over_J_flu,
                  const float *DsDxi_over_J_flu, const float *DzDxi_over_J_flu, const float *inv_s_fluid,
                  const float *inv_rho_fluid);
/* number of xdmf snapshots taken so far (isnap of the reference) */
int AXB(xdmf_count)(axb_handle h, int32_t *nsnap);
/* out(npoint_plot, nsnap, 5): u_s, u_p, u_z, straintrace, curlinplane of snapshots
 * first..first+nsnap-1 (0-based) — the records of xdmf_snap_{s,p,z,trace,curlip}_NNNN.dat
 * (wavefields_io.f90:195-199; u_p is not written for monopole sources, it is zero here) */
int AXB(fetch_xdmf)(axb_handle h, int32_t first, int32_t nsnap, float *out);

/* data_comm.f90:36-71.  glocal_index_msg is (maxmsg, nmsg) Fortran order; send and
 * receive lists coincide (get_mesh.f90:303-310).  glob2el is (num_comm_gll,3) =
 * (ipol, jpol, iel). */
int AXB(set_halo)(axb_handle h, int32_t domain, int32_t nmsg, const int32_t *list_peer,
                  const int32_t *sizemsg, const int32_t *glocal_index_msg, int32_t maxmsg,
                  int32_t num_comm_gll, const int32_t *glob2el);

/* data_time.f90: time_scheme, deltat, niter, seis_it, strain_it (0 = no wavefield dump) */
int AXB(set_time)(axb_handle h, int32_t scheme, double deltat, int32_t niter,
                  int32_t seis_it, int32_t strain_it);

/* dump_energy (time_evol_wave.F90:1424-1526, called from dump_stuff :1150 at iter 0 and after
 * every step): unassem_mass_rho_solid(0:4,0:4,nel_solid) (dipole: factor two folded in,
 * def_precomp_terms.f90:745-751) and unassem_mass_lam_fluid(0:4,0:4,nel_fluid) (:812-815;
 * NULL without a fluid).  Enables the diagnostic. */
int AXB(set_energy)(axb_handle h, const float *unassem_mass_rho_solid,
                    const float *unassem_mass_lam_fluid);

/* builds the derived (device) structures; must follow all axb_set_* calls */
int AXB(finalize_setup)(axb_handle h);

/* Multi-rank runs: every rank's handle must be connected to its peers before axb_run.
 * In-process (oracle, single-process multi-GPU): axb_connect_local with all handles.
 * One process per GPU: every rank exports one opaque blob of AXB_IPC_BLOB_BYTES bytes
 * (CUDA IPC handles of its receive slabs plus its message lists;
 * axb_ipc_blob_bytes() returns the same number at run time), the blobs travel once over any
 * channel (MPI_Allgather in the Fortran host, torch.distributed here), and each rank imports
 * the blobs of the ranks its halo lists name.  A smaller buffer is rejected with an error.
 * The mappings an import opens are closed by axb_destroy.
 *
 * A neighbour that never delivers does not hang the GPU: every wait for a neighbour's value is
 * bounded (10 s; AXB_HALO_TIMEOUT_MS overrides), the rank then raises its abort flag and
 * axb_synchronize fails with "HALO EXCHANGE TIMED OUT ..." — the counterpart of the
 * reference's pcheck, which stops all ranks when one fails (commpi.F90:64-111). */
#define AXB_IPC_BLOB_BYTES 1024
int32_t AXB(ipc_blob_bytes)(void);
int AXB(connect_local)(axb_handle *handles, int32_t n);
int AXB(ipc_export)(axb_handle h, void *blob, int32_t blob_bytes);
int AXB(ipc_import)(axb_handle h, int32_t peer_rank, const void *blob, int32_t blob_bytes);

/* Launch on an existing CUDA stream (a cudaStream_t, e.g. torch's current stream) instead
 * of the handle's own; kernels are only enqueued by axb_run — axb_synchronize (or any
 * fetch/get call) waits for them.  No-ops in the oracle. */
int AXB(set_stream)(axb_handle h, void *cuda_stream);
int AXB(synchronize)(axb_handle h);

/* Advance `nsteps` full time steps (collective over connected handles: every rank
 * must call it with the same nsteps).  State stays on the device. */
int AXB(run)(axb_handle h, int32_t nsteps);
/* in-process lockstep variant for connect_local groups */
int AXB(run_group)(axb_handle *handles, int32_t n, int32_t nsteps);

/* Per-kernel device timing with CUDA events on the launching stream (off by default).
 * Classes: 0 solid element kernel (S_A), 1 fluid element kernel (F_A), 2 fluid corrector
 * (F_B), 3 S/F coupling, 4 solid corrector (S_B), 5 halo pack/signal/wait, 6 sampling and
 * dumps, 7 other.  get_profile synchronises, returns accumulated milliseconds and launch
 * counts per class (arrays of 8) and resets the accumulators.  enable: 0 off, 1 events around
 * every launch, 2 events around the solid element kernel only (two records per step, so the
 * step time is not disturbed). */
int AXB(profile)(axb_handle h, int32_t enable);
int AXB(get_profile)(axb_handle h, double *ms, int64_t *launches);

int32_t AXB(iter)(axb_handle h);         /* time steps done so far                     */
int32_t AXB(nseismo)(axb_handle h);      /* seismogram samples recorded so far         */
int32_t AXB(nstrain)(axb_handle h);      /* wavefield snapshots recorded so far        */
int64_t AXB(gpu_launches)(axb_handle h); /* kernels launched by this handle (0: oracle) */

/* recdumpvar slice: out(3, num_rec, nsamples) Fortran order, samples first..first+n-1
 * (0-based; sample 0 is the dump at iter 0).  nc_dump_rec, nc_routines.F90:530-540 */
int AXB(fetch_seismograms)(axb_handle h, int32_t first, int32_t nsamples, float *out);
/* oneddumpvar slice: out(npoints, nsnap, nvars), shape and variable order as axb_snapshot_layout
 * reports (displ_only: 3 variables s, p, z); nc_routines.F90:248,275 */
int AXB(fetch_snapshots)(axb_handle h, int32_t first, int32_t nsnap, float *out);

/* energy samples first..first+n-1 (0-based; sample k belongs to iter k): out(4, n) =
 * sum(stiff*disp), sum(vel**2*mass) of the solid, sum(ddchi**2*mass) and sum(stiff*dchi) of the
 * fluid — this rank's sums; the host applies psum and two*pi as time_evol_wave.F90:1474-1510 */
int AXB(fetch_energy)(axb_handle h, int32_t first, int32_t n, float *out);

int AXB(get_state)(axb_handle h, int32_t field, float *out);
int AXB(set_state)(axb_handle h, int32_t field, const float *in);

/* run one operator of the loop on the current state (testing) */
int AXB(apply_op)(axb_handle h, int32_t op);

#ifdef __cplusplus
}
#endif
#endif /* AXISEM_B200_H */
