/*
 * axisem_oracle.c — CPU restatement of the AxiSEM SOLVER time loop.  TEST ORACLE ONLY.
 *
 * This file is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libaxisem_b200.so)
 * never links, loads or calls it.
 *
 * PARITY: PINNED END TO END, UNPINNED AT ARRAY LEVEL.  The reference is Fortran 2003 + MPI and
 * cannot be compiled in this image (no Fortran compiler), and its own tests hold no
 * array-level vectors for this path.  What its tests do hold are the
 * golden seismograms of the nightly regression (TESTING/nightly/test_0{1,2,3}/ref_data/
 * axisem.mseed: explosion, mtr, mtp in elastic prem_ani) and the tabulated background model of
 * TEST04.  Both are committed as fixtures (tests/golden/nightly_ref_seismograms.npz,
 * prem_ani_model_bm.npz) and this restatement reproduces them: the model to print precision,
 * the seismograms of all three source orders with median waveform correlation 0.9985 /
 * 0.9997 / 0.9998 (explosion / mtr / mtp; minimum 0.92 / 0.99 / 0.996) and median amplitude
 * ratio 1.003-1.006 on a synthetic mesh that differs from the reference's
 * (tests/test_nightly_reference.py, tests/nightly_compare.py); on a whole Earth built as the
 * mesher builds it (inner square, fluid core, coarsening layers; tests/test_nightly_full_sphere.py)
 * the minimum over all traces is 0.9974 / 0.986 / 0.9984 and every amplitude within 2.5 %, the level
 * at which the reference's traces agree with the YSPEC solution it ships.  Sample-level identity with an
 * execution of the Fortran is not established; beyond the two fixtures the restatement is
 * pinned by the analytic / self-consistency checks the reference itself uses
 * (tests/test_oracle_physics.py) and by committed fixtures of its own output.
 *
 * Arithmetic rules reproduced from the Fortran:
 *   - fields and pre-computed planes are real(4); `sum(a(i,:)*b(:,j))` is evaluated
 *     left to right in real(4) (unrolled_loops.f90:164-188); compiled with
 *     -ffp-contract=off so no FMA contraction happens;
 *   - expressions containing the real(8) scalars deltat, half_dt, half_dt_sq, two,
 *     third, coefd/coefv, ts_fac_*, exp_w_j_deltat, a_j_* are promoted to real(8) and
 *     rounded once on assignment (time_evol_wave.F90:357-364,459,472,477;
 *     attenuation.f90:144-192);
 *   - flush-to-zero is on (SOLVER/ftz.c:44-48): axo_run sets MXCSR FTZ|DAZ.
 *
 * Layout: exactly the Fortran memory order.  u(0:4,0:4,nel,3) -> u[i + 5*j + 25*e + 25*nel*c].
 */
#define AXB_PREFIX axo_
#include "../include/axisem_b200.h"

#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <xmmintrin.h>
#include <pmmintrin.h>
#endif

#define NP 5
#define NPT 25
#define MAXMSG 8

typedef struct {
    int nmsg;
    int peer[MAXMSG];
    int size[MAXMSG];
    int *glocal[MAXMSG];      /* 1-based glocal ids */
    int ncomm;
    int *glob2el;             /* (ncomm,3) Fortran order */
    float *sendbuf[MAXMSG];   /* (size, nc) */
    float *recvbuf[MAXMSG];
    /* one-process-per-rank mode (axo_ipc_*): receive slabs live in POSIX shared memory */
    float *shm_recv[2][MAXMSG];    /* [parity][msg] in my segment */
    float *peer_recv[2][MAXMSG];   /* where my message m goes (peer's segment) */
    volatile int *my_flag[MAXMSG], *peer_flag[MAXMSG];
    int seq;
} halo_t;

struct axb_handle_s {
    int rank, nranks;
    int nel_s, nel_f, nglob_s, nglob_f;
    int *igloc_s, *igloc_f, *axis_s, *axis_f, *ax_el_s, *ax_el_f;
    int naxel_s, naxel_f;
    float G0[NP], G1[NPT], G1T[NPT], G2[NPT], G2T[NPT];
    int src_order;
    /* solid planes (copies) */
    float *M11s, *M21s, *M41s, *M12s, *M22s, *M32s, *M42s, *M11z, *M21z, *M41z;
    float *M13s, *M33s, *M43s, *M1phi, *M2phi, *M4phi;
    float *M_1, *M_2, *M_3, *M_4, *M_5, *M_6, *M_7, *M_8;
    float *M_w1, *M_w2, *M_w3, *M_w4, *M_w5;
    float *M0_w1, *M0_w2, *M0_w3, *M0_w4, *M0_w5, *M0_w6, *M0_w7, *M0_w8, *M0_w9, *M0_w10;
    /* fluid */
    float *M1chi, *M2chi, *M4chi, *M_w_fl, *M0_w_fl, *inv_mass_fluid, *fs_mask;
    float *inv_mass_rho;
    float *gamma_s, *gamma_f;
    int have_abc;
    /* S/F boundary */
    int nel_bdry;
    int *bdry_sel, *bdry_fel, *bdry_js, *bdry_jf;
    float *bdry_matr;
    /* attenuation */
    int anel, cg, n_sls, corr_lowq;
    double *y_j, *exp_w, *ts_t, *ts_tm1;
    float *Q_mu, *Q_kappa;
    float *dmu_cg, *dka_cg, *Ycg, *Vse_cg, *Vsx_cg, *Vze_cg, *Vzx_cg;
    float *Dse_cg, *Dze_cg, *Dsx_cg, *Dzx_cg;
    float *dmu, *dka, *Y, *Vse, *Vsx, *Vze, *Vzx, *Y0, *V0se, *V0sx, *V0ze, *V0zx;
    float *Dse, *Dze, *Dsx, *Dzx, *inv_s;
    /* source */
    int fluid_src, nelsrc, ielsrc[8], niter_stf;
    float *src_term;     /* (5,5,8,3) or (5,5,8) */
    float *stf;
    int stf_type;
    double decay, t_0, shift_fact, magnitude;
    /* receivers */
    int num_rec;
    int *recfile_el;     /* (num_rec,3) Fortran order */
    /* kwf */
    int have_kwf, npt_s_kwf, npt_f_kwf;
    int *kwf_mask, *kwf_map;
    float *inv_rho_fluid, *Dse_f, *Dze_f, *Dsx_f, *Dzx_f;
    halo_t halo[2];
    /* time */
    int scheme, niter, seis_it, strain_it;
    double deltat, half_dt, half_dt_sq, t;
    int nstages;
    double coefd[40], coefv[40], coeff[40];
    /* state */
    float *disp, *velo, *acc0, *acc1, *chi, *dchi, *ddchi0, *ddchi1;
    float *memvar, *src_dev_tm1, *src_tr_tm1;
    float *gvec_s, *gvec_f;
    int iter, iseismo, istrain;
    float *recdump;      /* (3, num_rec, nseismo_max) */
    int dump_energy;     /* time_evol_wave.F90:1150 */
    float *um_rho_s, *um_lam_f;   /* unassem_mass_rho_solid, unassem_mass_lam_fluid */
    float *energy;       /* (4, niter + 1) */
    int nseismo_max;
    float *snapdump;     /* (npoints, nstrain_max, nvars) */
    int nstrain_max;
    /* dump_type strain_only / fullfields (axo_set_dump) */
    int dump_type, ibeg, iend, jbeg, jend;
    float *dDse, *dDze, *dDsx, *dDzx, *d_inv_s, *d_inv_s_f;   /* data_pointwise planes of the dumps */
    /* xdmf snapshots (axo_set_xdmf) */
    int have_xdmf, snap_it, isnap, nsnap_max, x_in, x_jn, npoint_plot;
    int *x_iarr, *x_jarr, *x_mask, *x_map;
    float *xs_Dse, *xs_Dze, *xs_Dsx, *xs_Dzx, *xs_inv_s;        /* solid planes */
    float *xf_Dse, *xf_Dze, *xf_Dsx, *xf_Dzx, *xf_inv_s, *xf_inv_rho;
    float *xsnap;        /* (npoint_plot, nsnap_max, 5) */
    int finalized;
    struct axb_handle_s **group;
    int ngroup;
    /* shared-memory segment of this rank (one-process-per-rank mode) */
    char shm_name[64];
    void *shm_base;
    size_t shm_bytes;
    int ipc_mode;
};

typedef struct axb_handle_s axo_t;

static char g_err[512] = "";
static int fail(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return 1; }
const char *axo_last_error(void) { return g_err; }

static float *dupf(const float *p, size_t n) {
    if (!p) return NULL;
    float *q = (float *)malloc((n ? n : 1) * sizeof(float));
    memcpy(q, p, n * sizeof(float));
    return q;
}
static int *dupi(const int32_t *p, size_t n) {
    if (!p) return NULL;
    int *q = (int *)malloc((n ? n : 1) * sizeof(int));
    memcpy(q, p, n * sizeof(int));
    return q;
}
static double *dupd(const double *p, size_t n) {
    if (!p) return NULL;
    double *q = (double *)malloc((n ? n : 1) * sizeof(double));
    memcpy(q, p, n * sizeof(double));
    return q;
}
static float *zerosf(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

int axo_create(axb_handle *h, int32_t device, int32_t rank, int32_t nranks) {
    (void)device;
    axo_t *o = (axo_t *)calloc(1, sizeof(axo_t));
    if (!o) return fail("out of memory");
    o->rank = rank;
    o->nranks = nranks;
    o->seis_it = 1;
    *h = o;
    return 0;
}

int axo_destroy(axb_handle h) {
    /* test infrastructure: arrays are released with the process; free the big ones */
    if (!h) return 0;
    free(h->disp); free(h->velo); free(h->acc0); free(h->acc1);
    free(h->chi); free(h->dchi); free(h->ddchi0); free(h->ddchi1);
    free(h->memvar); free(h->src_dev_tm1); free(h->src_tr_tm1);
    free(h->gvec_s); free(h->gvec_f); free(h->recdump); free(h->snapdump);
    free(h->x_iarr); free(h->x_jarr); free(h->x_mask); free(h->x_map); free(h->xsnap);
    free(h->xs_Dse); free(h->xs_Dze); free(h->xs_Dsx); free(h->xs_Dzx); free(h->xs_inv_s);
    free(h->xf_Dse); free(h->xf_Dze); free(h->xf_Dsx); free(h->xf_Dzx); free(h->xf_inv_s); free(h->xf_inv_rho);
    if (h->shm_base) { munmap(h->shm_base, h->shm_bytes); shm_unlink(h->shm_name); }
    free(h);
    return 0;
}

int axo_set_mesh(axb_handle h, int32_t npol, int32_t nel_solid, int32_t nel_fluid,
                 int32_t nglob_solid, int32_t nglob_fluid, const int32_t *igloc_solid,
                 const int32_t *igloc_fluid, const int32_t *axis_solid,
                 const int32_t *axis_fluid, const int32_t *ax_el_solid, int32_t naxel_solid,
                 const int32_t *ax_el_fluid, int32_t naxel_fluid, const float *G0,
                 const float *G1, const float *G1T, const float *G2, const float *G2T) {
    if (npol != 4) return fail("npol must be 4");
    h->nel_s = nel_solid; h->nel_f = nel_fluid;
    h->nglob_s = nglob_solid; h->nglob_f = nglob_fluid;
    h->igloc_s = dupi(igloc_solid, (size_t)NPT * nel_solid);
    h->igloc_f = dupi(igloc_fluid, (size_t)NPT * nel_fluid);
    h->axis_s = dupi(axis_solid, nel_solid);
    h->axis_f = dupi(axis_fluid, nel_fluid);
    h->ax_el_s = dupi(ax_el_solid, naxel_solid); h->naxel_s = naxel_solid;
    h->ax_el_f = dupi(ax_el_fluid, naxel_fluid); h->naxel_f = naxel_fluid;
    memcpy(h->G0, G0, sizeof h->G0);
    memcpy(h->G1, G1, sizeof h->G1); memcpy(h->G1T, G1T, sizeof h->G1T);
    memcpy(h->G2, G2, sizeof h->G2); memcpy(h->G2T, G2T, sizeof h->G2T);
    return 0;
}

int axo_set_solid_terms(axb_handle h, int32_t src_order, const axb_solid_terms *t) {
    size_t n = (size_t)NPT * h->nel_s, n0 = (size_t)NP * h->nel_s;
    h->src_order = src_order;
#define CP(x) h->x = dupf(t->x, n)
#define CP0(x) h->x = dupf(t->x, n0)
    CP(M11s); CP(M21s); CP(M41s); CP(M12s); CP(M22s); CP(M32s); CP(M42s);
    CP(M11z); CP(M21z); CP(M41z); CP(M13s); CP(M33s); CP(M43s);
    CP(M1phi); CP(M2phi); CP(M4phi);
    CP(M_1); CP(M_2); CP(M_3); CP(M_4); CP(M_5); CP(M_6); CP(M_7); CP(M_8);
    CP(M_w1); CP(M_w2); CP(M_w3); CP(M_w4); CP(M_w5);
    CP0(M0_w1); CP0(M0_w2); CP0(M0_w3); CP0(M0_w4); CP0(M0_w5);
    CP0(M0_w6); CP0(M0_w7); CP0(M0_w8); CP0(M0_w9); CP0(M0_w10);
#undef CP
#undef CP0
    return 0;
}

int axo_set_fluid_terms(axb_handle h, const float *M1chi_fl, const float *M2chi_fl,
                        const float *M4chi_fl, const float *M_w_fl, const float *M0_w_fl,
                        const float *inv_mass_fluid, const float *fluid_free_surface_mask) {
    size_t n = (size_t)NPT * h->nel_f;
    h->M1chi = dupf(M1chi_fl, n); h->M2chi = dupf(M2chi_fl, n); h->M4chi = dupf(M4chi_fl, n);
    h->M_w_fl = dupf(M_w_fl, n); h->M0_w_fl = dupf(M0_w_fl, (size_t)NP * h->nel_f);
    h->inv_mass_fluid = dupf(inv_mass_fluid, n);
    h->fs_mask = dupf(fluid_free_surface_mask, n);
    return 0;
}

int axo_set_mass(axb_handle h, const float *inv_mass_rho) {
    h->inv_mass_rho = dupf(inv_mass_rho, (size_t)NPT * h->nel_s);
    return 0;
}

int axo_set_energy(axb_handle h, const float *um_rho, const float *um_lam) {
    if (!um_rho && h->nel_s > 0) return fail("axo_set_energy: NULL unassem_mass_rho_solid");
    if (!um_lam && h->nel_f > 0) return fail("axo_set_energy: NULL unassem_mass_lam_fluid");
    h->um_rho_s = dupf(um_rho, (size_t)NPT * h->nel_s);
    h->um_lam_f = dupf(um_lam, (size_t)NPT * h->nel_f);
    h->dump_energy = 1;
    return 0;
}
int axo_set_sponge(axb_handle h, const float *solid_gamma, const float *fluid_gamma) {
    h->have_abc = (solid_gamma != NULL) || (fluid_gamma != NULL);
    h->gamma_s = dupf(solid_gamma, (size_t)NPT * h->nel_s);
    h->gamma_f = dupf(fluid_gamma, (size_t)NPT * h->nel_f);
    if (h->have_abc && !h->gamma_s) h->gamma_s = zerosf((size_t)NPT * h->nel_s);
    if (h->have_abc && !h->gamma_f) h->gamma_f = zerosf((size_t)NPT * h->nel_f);
    return 0;
}

int axo_set_sf_boundary(axb_handle h, int32_t nel_bdry, const int32_t *bdry_solid_el,
                        const int32_t *bdry_fluid_el, const int32_t *bdry_jpol_solid,
                        const int32_t *bdry_jpol_fluid, const float *bdry_matr) {
    h->nel_bdry = nel_bdry;
    h->bdry_sel = dupi(bdry_solid_el, nel_bdry); h->bdry_fel = dupi(bdry_fluid_el, nel_bdry);
    h->bdry_js = dupi(bdry_jpol_solid, nel_bdry); h->bdry_jf = dupi(bdry_jpol_fluid, nel_bdry);
    h->bdry_matr = dupf(bdry_matr, (size_t)NP * nel_bdry * 2);
    return 0;
}

int axo_set_attenuation(axb_handle h, const axb_attenuation *a) {
    size_t n4 = (size_t)4 * h->nel_s, n = (size_t)NPT * h->nel_s, n0 = (size_t)NP * h->nel_s;
    h->anel = 1; h->cg = a->coarse_grained; h->n_sls = a->n_sls; h->corr_lowq = a->do_corr_lowq;
    h->y_j = dupd(a->y_j, a->n_sls); h->exp_w = dupd(a->exp_w_j_deltat, a->n_sls);
    h->ts_t = dupd(a->ts_fac_t, a->n_sls); h->ts_tm1 = dupd(a->ts_fac_tm1, a->n_sls);
    h->Q_mu = dupf(a->Q_mu, h->nel_s); h->Q_kappa = dupf(a->Q_kappa, h->nel_s);
    h->inv_s = dupf(a->inv_s_solid, n);
    if (h->cg) {
        h->dmu_cg = dupf(a->delta_mu_cg4, n4); h->dka_cg = dupf(a->delta_kappa_cg4, n4);
        h->Ycg = dupf(a->Y_cg4, n4); h->Vse_cg = dupf(a->V_s_eta_cg4, n4);
        h->Vsx_cg = dupf(a->V_s_xi_cg4, n4); h->Vze_cg = dupf(a->V_z_eta_cg4, n4);
        h->Vzx_cg = dupf(a->V_z_xi_cg4, n4);
        h->Dse_cg = dupf(a->DsDeta_over_J_sol_cg4, n4); h->Dze_cg = dupf(a->DzDeta_over_J_sol_cg4, n4);
        h->Dsx_cg = dupf(a->DsDxi_over_J_sol_cg4, n4); h->Dzx_cg = dupf(a->DzDxi_over_J_sol_cg4, n4);
        if (!h->dmu_cg || !h->Ycg || !h->Dse_cg || !h->inv_s) return fail("cg4 attenuation arrays missing");
    } else {
        h->dmu = dupf(a->delta_mu, n); h->dka = dupf(a->delta_kappa, n);
        h->Y = dupf(a->Y, n); h->Vse = dupf(a->V_s_eta, n); h->Vsx = dupf(a->V_s_xi, n);
        h->Vze = dupf(a->V_z_eta, n); h->Vzx = dupf(a->V_z_xi, n);
        h->Y0 = dupf(a->Y0, n0); h->V0se = dupf(a->V0_s_eta, n0); h->V0sx = dupf(a->V0_s_xi, n0);
        h->V0ze = dupf(a->V0_z_eta, n0); h->V0zx = dupf(a->V0_z_xi, n0);
        h->Dse = dupf(a->DsDeta_over_J_sol, n); h->Dze = dupf(a->DzDeta_over_J_sol, n);
        h->Dsx = dupf(a->DsDxi_over_J_sol, n); h->Dzx = dupf(a->DzDxi_over_J_sol, n);
        if (!h->dmu || !h->Y || !h->Y0 || !h->Dse || !h->inv_s) return fail("full attenuation arrays missing");
    }
    return 0;
}

int axo_set_source(axb_handle h, int32_t fluid_src, int32_t nelsrc, const int32_t *ielsrc,
                   const float *source_term, const float *stf, int32_t niter) {
    if (nelsrc > 8) return fail("nelsrc > 8");
    h->fluid_src = fluid_src; h->nelsrc = nelsrc;
    for (int k = 0; k < 8; k++) h->ielsrc[k] = (ielsrc && k < nelsrc) ? ielsrc[k] : 0;
    h->src_term = dupf(source_term, (size_t)NPT * 8 * (fluid_src ? 1 : 3));
    h->stf = dupf(stf, niter); h->niter_stf = niter;
    return 0;
}

int axo_set_stf_params(axb_handle h, int32_t stf_type, double decay, double t_0,
                       double shift_fact, double magnitude) {
    h->stf_type = stf_type; h->decay = decay; h->t_0 = t_0;
    h->shift_fact = shift_fact; h->magnitude = magnitude;
    return 0;
}

int axo_set_receivers(axb_handle h, int32_t num_rec, const int32_t *recfile_el) {
    h->num_rec = num_rec;
    h->recfile_el = dupi(recfile_el, (size_t)3 * num_rec);
    return 0;
}

int axo_set_kwf(axb_handle h, const int32_t *kwf_mask, const int32_t *mapping_ijel_ikwf,
                int32_t npoint_solid_kwf, int32_t npoint_fluid_kwf, const float *inv_rho_fluid,
                const float *DsDeta_over_J_flu, const float *DzDeta_over_J_flu,
                const float *DsDxi_over_J_flu, const float *DzDxi_over_J_flu) {
    size_t n = (size_t)NPT * (h->nel_s + h->nel_f), nf = (size_t)NPT * h->nel_f;
    h->have_kwf = 1; h->npt_s_kwf = npoint_solid_kwf; h->npt_f_kwf = npoint_fluid_kwf;
    h->kwf_mask = dupi(kwf_mask, n); h->kwf_map = dupi(mapping_ijel_ikwf, n);
    h->inv_rho_fluid = dupf(inv_rho_fluid, nf);
    h->Dse_f = dupf(DsDeta_over_J_flu, nf); h->Dze_f = dupf(DzDeta_over_J_flu, nf);
    h->Dsx_f = dupf(DsDxi_over_J_flu, nf); h->Dzx_f = dupf(DzDxi_over_J_flu, nf);
    return 0;
}

/* dump_type of the wavefield dumps (data_io.f90; parameters.F90:400-403 for ibeg..jend) */
int axo_set_dump(axb_handle h, int32_t dump_type, int32_t ibeg, int32_t iend, int32_t jbeg, int32_t jend,
                 const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                 const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol,
                 const float *inv_s_solid, const float *inv_s_fluid) {
    size_t n = (size_t)NPT * h->nel_s, nf = (size_t)NPT * h->nel_f;
    if (dump_type < AXB_DUMP_DISPL_ONLY || dump_type > AXB_DUMP_FULLFIELDS) return fail("unknown dump_type");
    if (ibeg < 0 || iend > 4 || ibeg > iend || jbeg < 0 || jend > 4 || jbeg > jend) return fail("bad ibeg..jend");
    h->dump_type = dump_type; h->ibeg = ibeg; h->iend = iend; h->jbeg = jbeg; h->jend = jend;
    if (dump_type == AXB_DUMP_DISPL_ONLY) return 0;
    if (!DsDeta_over_J_sol || !DzDeta_over_J_sol || !DsDxi_over_J_sol || !DzDxi_over_J_sol || !inv_s_solid ||
        (h->nel_f > 0 && !inv_s_fluid))
        return fail("axo_set_dump: NULL plane");
    h->dDse = dupf(DsDeta_over_J_sol, n); h->dDze = dupf(DzDeta_over_J_sol, n);
    h->dDsx = dupf(DsDxi_over_J_sol, n); h->dDzx = dupf(DzDxi_over_J_sol, n);
    h->d_inv_s = dupf(inv_s_solid, n); h->d_inv_s_f = dupf(inv_s_fluid, nf);
    return 0;
}
/* xdmf snapshots: the maps of dump_xdmf_grid (meshes_io.F90:110-437), fluid elements first */
int axo_set_xdmf(axb_handle h, int32_t snap_it, int32_t i_n_xdmf, int32_t j_n_xdmf,
                 const int32_t *i_arr_xdmf, const int32_t *j_arr_xdmf,
                 const int32_t *plotting_mask, const int32_t *mapping_ijel_iplot, int32_t npoint_plot,
                 const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                 const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol, const float *inv_s_solid,
                 const float *DsDeta_over_J_flu, const float *DzDeta_over_J_flu,
                 const float *DsDxi_over_J_flu, const float *DzDxi_over_J_flu, const float *inv_s_fluid,
                 const float *inv_rho_fluid) {
    size_t n = (size_t)NPT * h->nel_s, nf = (size_t)NPT * h->nel_f;
    size_t nm = (size_t)i_n_xdmf * j_n_xdmf * (h->nel_s + h->nel_f);
    if (snap_it < 1) return fail("axo_set_xdmf: snap_it must be positive");
    if (i_n_xdmf < 1 || i_n_xdmf > NP || j_n_xdmf < 1 || j_n_xdmf > NP) return fail("axo_set_xdmf: bad i_n_xdmf / j_n_xdmf");
    for (int k = 0; k < i_n_xdmf; k++) if (i_arr_xdmf[k] < 0 || i_arr_xdmf[k] >= NP) return fail("axo_set_xdmf: i_arr_xdmf out of range");
    for (int k = 0; k < j_n_xdmf; k++) if (j_arr_xdmf[k] < 0 || j_arr_xdmf[k] >= NP) return fail("axo_set_xdmf: j_arr_xdmf out of range");
    if (!DsDeta_over_J_sol || !DzDeta_over_J_sol || !DsDxi_over_J_sol || !DzDxi_over_J_sol || !inv_s_solid)
        return fail("axo_set_xdmf: NULL solid plane");
    if (h->nel_f > 0 && (!DsDeta_over_J_flu || !DzDeta_over_J_flu || !DsDxi_over_J_flu || !DzDxi_over_J_flu ||
                         !inv_s_fluid || !inv_rho_fluid))
        return fail("axo_set_xdmf: NULL fluid plane");
    for (size_t k = 0; k < nm; k++)
        if (plotting_mask[k] && (mapping_ijel_iplot[k] < 1 || mapping_ijel_iplot[k] > npoint_plot))
            return fail("axo_set_xdmf: mapping_ijel_iplot out of range");
    h->have_xdmf = 1; h->snap_it = snap_it; h->x_in = i_n_xdmf; h->x_jn = j_n_xdmf; h->npoint_plot = npoint_plot;
    h->x_iarr = dupi(i_arr_xdmf, i_n_xdmf); h->x_jarr = dupi(j_arr_xdmf, j_n_xdmf);
    h->x_mask = dupi(plotting_mask, nm); h->x_map = dupi(mapping_ijel_iplot, nm);
    h->xs_Dse = dupf(DsDeta_over_J_sol, n); h->xs_Dze = dupf(DzDeta_over_J_sol, n);
    h->xs_Dsx = dupf(DsDxi_over_J_sol, n); h->xs_Dzx = dupf(DzDxi_over_J_sol, n); h->xs_inv_s = dupf(inv_s_solid, n);
    if (h->nel_f > 0) {
        h->xf_Dse = dupf(DsDeta_over_J_flu, nf); h->xf_Dze = dupf(DzDeta_over_J_flu, nf);
        h->xf_Dsx = dupf(DsDxi_over_J_flu, nf); h->xf_Dzx = dupf(DzDxi_over_J_flu, nf);
        h->xf_inv_s = dupf(inv_s_fluid, nf); h->xf_inv_rho = dupf(inv_rho_fluid, nf);
    }
    return 0;
}
int axo_xdmf_count(axb_handle h, int32_t *nsnap) { *nsnap = h->isnap; return 0; }
int axo_fetch_xdmf(axb_handle h, int32_t first, int32_t nsnap, float *out) {
    size_t npts = (size_t)h->npoint_plot;
    if (!h->have_xdmf) return fail("xdmf snapshots not enabled (axo_set_xdmf)");
    if (first < 0 || nsnap < 0 || first + nsnap > h->isnap) return fail("xdmf snapshot range");
    for (int v = 0; v < 5; v++)
        for (int s = 0; s < nsnap; s++)
            memcpy(out + npts * (s + (size_t)nsnap * v),
                   h->xsnap + npts * ((first + s) + (size_t)h->nsnap_max * v), sizeof(float) * npts);
    return 0;
}
static int snapshot_nvars(const axo_t *o) {
    const int mono = o->src_order == AXB_MONOPOLE;
    if (o->dump_type == AXB_DUMP_STRAIN_ONLY) return mono ? 4 : 6;
    if (o->dump_type == AXB_DUMP_FULLFIELDS) return mono ? 6 : 9;
    return 3;
}
static size_t snapshot_npoints(const axo_t *o) {
    if (o->dump_type == AXB_DUMP_FULLFIELDS)
        return (size_t)(o->iend - o->ibeg + 1) * (o->jend - o->jbeg + 1) * ((size_t)o->nel_s + o->nel_f);
    return (size_t)o->npt_s_kwf + o->npt_f_kwf;
}
int axo_snapshot_layout(axb_handle h, int32_t *npoints, int32_t *nvars) {
    *npoints = (int32_t)snapshot_npoints(h); *nvars = snapshot_nvars(h);
    return 0;
}

int axo_set_halo(axb_handle h, int32_t domain, int32_t nmsg, const int32_t *list_peer,
                 const int32_t *sizemsg, const int32_t *glocal_index_msg, int32_t maxmsg,
                 int32_t num_comm_gll, const int32_t *glob2el) {
    if (nmsg > MAXMSG) return fail("too many neighbours");
    halo_t *H = &h->halo[domain];
    int nc = domain == AXB_DOMAIN_SOLID ? 3 : 1;
    H->nmsg = nmsg;
    for (int m = 0; m < nmsg; m++) {
        H->peer[m] = list_peer[m];
        H->size[m] = sizemsg[m];
        H->glocal[m] = (int *)malloc(sizeof(int) * (sizemsg[m] ? sizemsg[m] : 1));
        for (int ip = 0; ip < sizemsg[m]; ip++) H->glocal[m][ip] = glocal_index_msg[ip + (size_t)maxmsg * m];
        H->sendbuf[m] = zerosf((size_t)sizemsg[m] * nc);
        H->recvbuf[m] = zerosf((size_t)sizemsg[m] * nc);
    }
    H->ncomm = num_comm_gll;
    H->glob2el = dupi(glob2el, (size_t)3 * num_comm_gll);
    return 0;
}

int axo_set_time(axb_handle h, int32_t scheme, double deltat, int32_t niter, int32_t seis_it,
                 int32_t strain_it) {
    h->scheme = scheme; h->deltat = deltat; h->niter = niter;
    h->seis_it = seis_it > 0 ? seis_it : 1; h->strain_it = strain_it;
    /* parameters.F90:1124-1125 */
    h->half_dt = 0.5 * deltat;
    h->half_dt_sq = 0.5 * deltat * deltat;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* symplectic_coefficients, time_evol_wave.F90:749-967; SS_scheme :972-992 (including
 * the integer division 1/2 == 0 at :981) */
static void ss_scheme(int n, int nstages, double *a, double *b, const double *g) {
    double s = 0.0;
    a[0] = g[0] / 2.0;
    for (int i = 1; i < n; i++) a[i] = (g[i - 1] + g[i]) / 2.0;
    for (int i = 0; i < n; i++) s += a[i];
    a[n] = (double)(1 / 2) - s;                       /* reference quirk: 1/2 == 0 */
    for (int i = n + 2; i <= 2 * n + 2; i++) a[i - 1] = a[2 * n + 3 - i - 1];
    for (int i = 0; i < n; i++) b[i] = g[i];
    s = 0.0;
    for (int i = 0; i < n; i++) s += g[i];
    b[n] = 1.0 - 2.0 * s;
    for (int i = n + 2; i <= 2 * n + 1; i++) b[i - 1] = b[2 * n + 2 - i - 1];
    (void)nstages;
}

static int symplectic_coefficients(axo_t *o) {
    double *d = o->coefd, *v = o->coefv;
    int n, ns = 0;
    switch (o->scheme) {
    case AXB_SYMPLEC4: {
        /* the reference's literals are default-real (single precision) constants */
        double zeta = (double)0.1786178958448091f, iota = (double)-0.2123418310626054f,
               kappa = (double)-0.06626458266981849f;
        ns = 4;
        d[0] = zeta; d[1] = kappa; d[2] = 1.0 - 2.0 * (zeta + kappa); d[3] = kappa; d[4] = zeta;
        v[0] = 0.5 - iota; v[1] = iota; v[2] = iota; v[3] = 0.5 - iota;
        break; }
    case AXB_ML_SO4M5: {
        double rho = (14.0 - sqrt(19.0)) / 108.0, theta = (20.0 - 7.0 * sqrt(19.0)) / 108.0;
        double nu = 2.0 / 5.0, lambda = -1.0 / 10.0;
        ns = 5;
        d[0] = rho; d[1] = theta; d[2] = 0.5 - rho - theta; d[3] = 0.5 - rho - theta;
        d[4] = theta; d[5] = rho;
        v[0] = nu; v[1] = lambda; v[2] = 1.0 - 2.0 * (nu + lambda); v[3] = lambda; v[4] = nu;
        break; }
    case AXB_ML_SO6M7: {
        ns = 7;
        d[0] = (double)-1.01308797891717472981f; d[1] = (double)1.18742957373254270702f;
        d[2] = (double)-0.01833585209646059034f; d[3] = (double)0.34399425728109261313f;
        for (int i = 5; i <= 8; i++) d[i - 1] = d[ns + 2 - i - 1];
        v[0] = (double)0.00016600692650009894f; v[1] = (double)-0.37962421426377360608f;
        v[2] = (double)0.68913741185181063674f; v[3] = (double)0.38064159097092574080f;
        for (int i = 5; i <= 7; i++) v[i - 1] = v[ns + 1 - i - 1];
        break; }
    case AXB_KL_O8M17: {
        static const float gf[8] = {0.13020248308889008088f, 0.56116298177510838456f,
            -0.38947496264484728641f, 0.15884190655515560090f, -0.39590389413323757734f,
            0.18453964097831570709f, 0.25837438768632204729f, 0.29501172360931029887f};
        double g[8];
        n = 8; ns = 2 * n + 1;
        for (int i = 0; i < n; i++) g[i] = (double)gf[i];
        ss_scheme(n, ns, d, v, g);
        break; }
    case AXB_SS_35O10: {
        static const float gf[17] = {0.078795722521686419263907679337684f,
            0.31309610341510852776481247192647f, 0.027918383235078066109520273275299f,
            -0.22959284159390709415121339679655f, 0.13096206107716486317465685927961f,
            -0.26973340565451071434460973222411f, 0.074973343155891435666137105641410f,
            0.11199342399981020488957508073640f, 0.36613344954622675119314812353150f,
            -0.39910563013603589787862981058340f, 0.10308739852747107731580277001372f,
            0.41143087395589023782070411897608f, -0.0048663605831352617621956593099771f,
            -0.39203335370863990644808193642610f, 0.051942502962449647037182904015976f,
            0.050665090759924496335874344156866f, 0.049674370639729879054568800279461f};
        double g[17];
        n = 17; ns = 2 * n + 1;
        for (int i = 0; i < n; i++) g[i] = (double)gf[i];
        ss_scheme(n, ns, d, v, g);
        break; }
    default:
        return fail("unknown time scheme");
    }
    o->nstages = ns;
    for (int i = 0; i <= ns; i++) d[i] *= o->deltat;
    for (int i = 0; i < ns; i++) v[i] *= o->deltat;
    for (int i = 0; i < ns; i++) {
        double s = 0.0;
        for (int k = 0; k <= i; k++) s += d[k];
        o->coeff[i] = s;
    }
    return 0;
}

int axo_finalize_setup(axb_handle h) {
    size_t ns = (size_t)NPT * h->nel_s * 3, nf = (size_t)NPT * h->nel_f;
    h->disp = zerosf(ns); h->velo = zerosf(ns); h->acc0 = zerosf(ns); h->acc1 = zerosf(ns);
    h->chi = zerosf(nf); h->dchi = zerosf(nf); h->ddchi0 = zerosf(nf); h->ddchi1 = zerosf(nf);
    h->gvec_s = zerosf((size_t)h->nglob_s * 3); h->gvec_f = zerosf((size_t)h->nglob_f);
    if (h->anel) {
        size_t per = h->cg ? 4 : NPT;
        h->memvar = zerosf(per * 6 * h->n_sls * h->nel_s);
        h->src_dev_tm1 = zerosf(per * 6 * h->nel_s);
        h->src_tr_tm1 = zerosf(per * h->nel_s);
    }
    if (h->scheme != AXB_NEWMARK2) { if (symplectic_coefficients(h)) return 1; }
    /* nseismo = floor(niter/seis_it) + 1 (parameters.F90:929) */
    h->nseismo_max = h->niter / h->seis_it + 1;
    h->recdump = zerosf((size_t)3 * h->num_rec * h->nseismo_max);
    if (h->dump_energy) h->energy = zerosf((size_t)4 * (h->niter + 1));
    if (h->strain_it > 0 && h->have_kwf) {
        h->nstrain_max = h->niter / h->strain_it + 1;
        if (h->dump_type != AXB_DUMP_DISPL_ONLY && !h->dDse) return fail("axo_set_dump: planes missing");
        h->snapdump = zerosf(snapshot_npoints(h) * h->nstrain_max * snapshot_nvars(h));
    }
    h->iter = 0; h->iseismo = 0; h->istrain = 0; h->t = 0.0;
    h->finalized = 1;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* unrolled_loops.f90:164-188: c(i,j) = sum(a(i,:)*b(:,j)), k ascending, real(4) */
static void mxm_4(const float *a, const float *b, float *c) {
    for (int j = 0; j < NP; j++)
        for (int i = 0; i < NP; i++) {
            float s = a[i + NP * 0] * b[0 + NP * j];
            s = s + a[i + NP * 1] * b[1 + NP * j];
            s = s + a[i + NP * 2] * b[2 + NP * j];
            s = s + a[i + NP * 3] * b[3 + NP * j];
            s = s + a[i + NP * 4] * b[4 + NP * j];
            c[i + NP * j] = s;
        }
}
/* unrolled_loops.f90:194-207 */
static void vxm_4(const float *a, const float *b, float *c) {
    for (int j = 0; j < NP; j++) {
        float s = a[0] * b[0 + NP * j];
        s = s + a[1] * b[1 + NP * j];
        s = s + a[2] * b[2 + NP * j];
        s = s + a[3] * b[3 + NP * j];
        s = s + a[4] * b[4 + NP * j];
        c[j] = s;
    }
}
/* unrolled_loops.f90:222-229: outerprod(a,b)(i,j) = a(i)*b(j) */
static void outerprod_4(const float *a, const float *b, float *c) {
    for (int j = 0; j < NP; j++)
        for (int i = 0; i < NP; i++) c[i + NP * j] = a[i] * b[j];
}
/* unrolled_loops.f90:76-96 */
static void mxm_cg4_sparse_a(const float *a, const float *b, float *c) {
    memset(c, 0, NPT * sizeof(float));
    for (int j = 0; j < NP; j++) {
        c[1 + NP * j] = a[0] * b[1 + NP * j] + a[1] * b[3 + NP * j];
        c[3 + NP * j] = a[2] * b[1 + NP * j] + a[3] * b[3 + NP * j];
    }
}
/* unrolled_loops.f90:100-120 */
static void mxm_cg4_sparse_b(const float *a, const float *b, float *c) {
    memset(c, 0, NPT * sizeof(float));
    for (int i = 0; i < NP; i++) {
        c[i + NP * 1] = a[i + NP * 1] * b[0] + a[i + NP * 3] * b[2];
        c[i + NP * 3] = a[i + NP * 1] * b[1] + a[i + NP * 3] * b[3];
    }
}
/* unrolled_loops.f90:124-158 */
static void mxm_cg4_sparse_c(const float *a, const float *b, float *c) {
    static const int ii[4] = {1, 1, 3, 3}, jj[4] = {1, 3, 1, 3};
    for (int q = 0; q < 4; q++) {
        int i = ii[q], j = jj[q];
        float s = a[i + NP * 0] * b[0 + NP * j];
        s = s + a[i + NP * 1] * b[1 + NP * j];
        s = s + a[i + NP * 2] * b[2 + NP * j];
        s = s + a[i + NP * 3] * b[3 + NP * j];
        s = s + a[i + NP * 4] * b[4 + NP * j];
        c[q] = s;
    }
}

#define EL(p, e) ((p) + (size_t)NPT * (e))
#define EL0(p, e) ((p) + (size_t)NP * (e))
#define FOR25 for (int q = 0; q < NPT; q++)

/* ------------------------------------------------------------------------------------ */
/* stiffness_mono.f90:60-157 */
static void glob_stiffness_mono_4(const axo_t *o, float *glob, const float *u) {
    const size_t cs = (size_t)NPT * o->nel_s;
    float us[NPT], uz[NPT], X1[NPT], X2[NPT], X3[NPT], X4[NPT];
    float S1s[NPT], S2s[NPT], S1z[NPT], S2z[NPT], ls[NPT], lz[NPT], T[NPT];
    float V1[NP], V2[NP], V3[NP], V4[NP], uz0[NP], W[NP];
    for (int e = 0; e < o->nel_s; e++) {
        const float *m_1 = EL(o->M_1, e), *m_2 = EL(o->M_2, e), *m_3 = EL(o->M_3, e), *m_4 = EL(o->M_4, e);
        const float *m_w1 = EL(o->M_w1, e);
        const float *m11s = EL(o->M11s, e), *m21s = EL(o->M21s, e), *m41s = EL(o->M41s, e);
        const float *m12s = EL(o->M12s, e), *m22s = EL(o->M22s, e), *m32s = EL(o->M32s, e), *m42s = EL(o->M42s, e);
        const float *m11z = EL(o->M11z, e), *m21z = EL(o->M21z, e), *m41z = EL(o->M41z, e);
        memcpy(us, EL(u, e), sizeof us);
        memcpy(uz, EL(u + 2 * cs, e), sizeof uz);
        if (!o->axis_s[e]) { mxm_4(o->G2T, us, X1); mxm_4(o->G2T, uz, X2); }
        else               { mxm_4(o->G1T, us, X1); mxm_4(o->G1T, uz, X2); }
        mxm_4(us, o->G2, X3);
        mxm_4(uz, o->G2, X4);
        FOR25 {
            ls[q] = m_4[q] * X4[q] + m_2[q] * X3[q] + m_1[q] * X1[q] + m_3[q] * X2[q] + us[q] * m_w1[q];
            S1s[q] = m11s[q] * X3[q] + m21s[q] * X1[q] + m12s[q] * X4[q] + m22s[q] * X2[q] + m_1[q] * us[q];
            S2s[q] = m11s[q] * X1[q] + m41s[q] * X3[q] + m32s[q] * X2[q] + m42s[q] * X4[q] + m_2[q] * us[q];
            S1z[q] = m11z[q] * X4[q] + m21z[q] * X2[q] + m32s[q] * X3[q] + m22s[q] * X1[q] + m_3[q] * us[q];
            S2z[q] = m11z[q] * X2[q] + m41z[q] * X4[q] + m12s[q] * X1[q] + m42s[q] * X3[q] + m_4[q] * us[q];
        }
        mxm_4(S2s, o->G2T, X2);
        mxm_4(S2z, o->G2T, X4);
        if (!o->axis_s[e]) { mxm_4(o->G2, S1s, X1); mxm_4(o->G2, S1z, X3); }
        else               { mxm_4(o->G1, S1s, X1); mxm_4(o->G1, S1z, X3); }
        FOR25 { ls[q] = ls[q] + X1[q] + X2[q]; lz[q] = X3[q] + X4[q]; }
        if (o->axis_s[e]) {
            const float *m0_w1 = EL0(o->M0_w1, e), *m0_w2 = EL0(o->M0_w2, e), *m0_w3 = EL0(o->M0_w3, e);
            for (int j = 0; j < NP; j++) uz0[j] = uz[0 + NP * j];
            vxm_4(o->G0, us, V1);
            vxm_4(uz0, o->G2, V2);
            for (int j = 0; j < NP; j++) V4[j] = m0_w1[j] * V1[j] + m0_w3[j] * V2[j];
            vxm_4(o->G0, uz, V3);
            for (int j = 0; j < NP; j++) V4[j] = V4[j] + m0_w2[j] * V3[j];
            for (int j = 0; j < NP; j++) W[j] = m0_w2[j] * V1[j];
            outerprod_4(o->G0, W, X2);
            for (int j = 0; j < NP; j++) V2[j] = m0_w3[j] * V1[j];
            vxm_4(V2, o->G2T, V1);
            for (int j = 0; j < NP; j++) X2[0 + NP * j] = X2[0 + NP * j] + V1[j];
            outerprod_4(o->G0, V4, T);
            FOR25 { ls[q] = ls[q] + T[q]; lz[q] = X2[q] + lz[q]; }
        }
        memcpy(EL(glob, e), ls, sizeof ls);
        memcpy(EL(glob + 2 * cs, e), lz, sizeof lz);
    }
}

/* stiffness_di.f90:60-256 */
static void glob_stiffness_di_4(const axo_t *o, float *glob, const float *u) {
    const size_t cs = (size_t)NPT * o->nel_s;
    float u1[NPT], u2[NPT], u3[NPT];
    float X1[NPT], X2[NPT], X3[NPT], X4[NPT], X5[NPT], X6[NPT], X7[NPT], X8[NPT];
    float S1p[NPT], S1m[NPT], S2p[NPT], S2m[NPT], S1z[NPT], S2z[NPT];
    float ls2[NPT], ls3[NPT], l1[NPT], l2[NPT], l3[NPT];
    float V1[NP], V2[NP], V3[NP], V4[NP], V5[NP], u10[NP], u20[NP], W[NP];
    for (int e = 0; e < o->nel_s; e++) {
        const float *m_1 = EL(o->M_1, e), *m_2 = EL(o->M_2, e), *m_3 = EL(o->M_3, e), *m_4 = EL(o->M_4, e);
        const float *m_5 = EL(o->M_5, e), *m_6 = EL(o->M_6, e), *m_7 = EL(o->M_7, e), *m_8 = EL(o->M_8, e);
        const float *m_w1 = EL(o->M_w1, e), *m_w2 = EL(o->M_w2, e), *m_w3 = EL(o->M_w3, e);
        const float *m11s = EL(o->M11s, e), *m21s = EL(o->M21s, e), *m41s = EL(o->M41s, e);
        const float *m12s = EL(o->M12s, e), *m22s = EL(o->M22s, e), *m42s = EL(o->M42s, e);
        const float *m13s = EL(o->M13s, e), *m23s = EL(o->M32s, e) /* stiffness_di.f90:128 */;
        const float *m33s = EL(o->M33s, e), *m43s = EL(o->M43s, e);
        const float *m11z = EL(o->M11z, e), *m21z = EL(o->M21z, e), *m41z = EL(o->M41z, e);
        memcpy(u1, EL(u, e), sizeof u1);
        memcpy(u2, EL(u + cs, e), sizeof u2);
        memcpy(u3, EL(u + 2 * cs, e), sizeof u3);
        mxm_4(u1, o->G2, X4); mxm_4(u2, o->G2, X5); mxm_4(u3, o->G2, X6);
        if (!o->axis_s[e]) { mxm_4(o->G2T, u1, X1); mxm_4(o->G2T, u2, X2); mxm_4(o->G2T, u3, X3); }
        else               { mxm_4(o->G1T, u1, X1); mxm_4(o->G1T, u2, X2); mxm_4(o->G1T, u3, X3); }
        FOR25 {
            float c1, c2, c3;
            X7[q] = X1[q] + X2[q];
            X8[q] = X4[q] + X5[q];
            ls2[q] = m_8[q] * X6[q] + m_7[q] * X3[q] + m_1[q] * X1[q] + m_5[q] * X2[q]
                   + m_2[q] * X4[q] + m_6[q] * X5[q] + m_w1[q] * u2[q] + m_w2[q] * u3[q];
            ls3[q] = m_4[q] * X4[q] - m_4[q] * X5[q] + m_3[q] * X1[q] - m_3[q] * X2[q]
                   + m_w2[q] * u2[q] + m_w3[q] * u3[q];
            c1 = m13s[q] * X6[q]; c2 = m23s[q] * X3[q]; c3 = m_3[q] * u3[q];
            S1p[q] = c1 + c2 + c3 + m11s[q] * X4[q] + m21s[q] * X1[q] + m12s[q] * X5[q] + m22s[q] * X2[q] + m_1[q] * u2[q];
            S1m[q] = c1 + c2 - c3 + m11s[q] * X5[q] + m21s[q] * X2[q] + m12s[q] * X4[q] + m22s[q] * X1[q] + m_5[q] * u2[q];
            c1 = m33s[q] * X3[q]; c2 = m43s[q] * X6[q]; c3 = m_4[q] * u3[q];
            S2p[q] = c1 + c2 + c3 + m11s[q] * X1[q] + m41s[q] * X4[q] + m12s[q] * X2[q] + m42s[q] * X5[q] + m_2[q] * u2[q];
            S2m[q] = c1 + c2 - c3 + m11s[q] * X2[q] + m41s[q] * X5[q] + m12s[q] * X1[q] + m42s[q] * X4[q] + m_6[q] * u2[q];
            S1z[q] = m33s[q] * X8[q] + m23s[q] * X7[q] + m11z[q] * X6[q] + m21z[q] * X3[q] + m_7[q] * u2[q];
            S2z[q] = m13s[q] * X7[q] + m43s[q] * X8[q] + m11z[q] * X3[q] + m41z[q] * X6[q] + m_8[q] * u2[q];
        }
        if (!o->axis_s[e]) { mxm_4(o->G2, S1p, X1); mxm_4(o->G2, S1m, X3); mxm_4(o->G2, S1z, X5); }
        else               { mxm_4(o->G1, S1p, X1); mxm_4(o->G1, S1m, X3); mxm_4(o->G1, S1z, X5); }
        mxm_4(S2p, o->G2T, X2); mxm_4(S2m, o->G2T, X4); mxm_4(S2z, o->G2T, X6);
        FOR25 {
            l1[q] = X1[q] + X2[q];
            l2[q] = X3[q] + X4[q] + ls2[q];
            l3[q] = X5[q] + X6[q] + ls3[q];
        }
        if (o->axis_s[e]) {
            const float *w1 = EL0(o->M0_w1, e), *w2 = EL0(o->M0_w2, e), *w3 = EL0(o->M0_w3, e);
            const float *w4 = EL0(o->M0_w4, e), *w6 = EL0(o->M0_w6, e), *w7 = EL0(o->M0_w7, e);
            const float *w8 = EL0(o->M0_w8, e), *w9 = EL0(o->M0_w9, e), *w10 = EL0(o->M0_w10, e);
            for (int j = 0; j < NP; j++) { u10[j] = u1[0 + NP * j]; u20[j] = u2[0 + NP * j]; }
            vxm_4(o->G0, u1, V1); vxm_4(o->G0, u2, V2); vxm_4(o->G0, u3, V3);
            vxm_4(u10, o->G2, V4); vxm_4(u20, o->G2, V5);
            for (int j = 0; j < NP; j++) W[j] = w1[j] * V2[j] + w3[j] * V3[j];
            outerprod_4(o->G0, W, S1p);
            for (int j = 0; j < NP; j++)
                W[j] = w1[j] * V1[j] + (w2[j] + w6[j]) * V4[j] + w9[j] * V2[j] + w10[j] * V3[j];
            outerprod_4(o->G0, W, S1m);
            for (int j = 0; j < NP; j++)
                W[j] = w3[j] * V1[j] + (w4[j] + w8[j]) * V4[j] + w7[j] * V3[j] + w10[j] * V2[j];
            outerprod_4(o->G0, W, S1z);
            for (int j = 0; j < NP; j++) V4[j] = (w2[j] + w6[j]) * V2[j] + (w4[j] + w8[j]) * V3[j];
            vxm_4(V4, o->G2T, V1);
            for (int j = 0; j < NP; j++) S1p[0 + NP * j] = S1p[0 + NP * j] + V1[j];
            FOR25 { l1[q] = l1[q] + S1p[q]; l2[q] = l2[q] + S1m[q]; l3[q] = l3[q] + S1z[q]; }
            (void)V5;
        }
        memcpy(EL(glob, e), l1, sizeof l1);
        memcpy(EL(glob + cs, e), l2, sizeof l2);
        memcpy(EL(glob + 2 * cs, e), l3, sizeof l3);
    }
}

/* stiffness_quad.f90:238-412 */
static void glob_stiffness_quad_4(const axo_t *o, float *glob, const float *u) {
    const size_t cs = (size_t)NPT * o->nel_s;
    float us[NPT], up[NPT], uz[NPT];
    float X1[NPT], X2[NPT], X3[NPT], X4[NPT], X5[NPT], X6[NPT];
    float S1s[NPT], S2s[NPT], S1p[NPT], S2p[NPT], S1z[NPT], S2z[NPT];
    float ls[NPT], lp[NPT], lz[NPT];
    float V1[NP], V2[NP], V3[NP], W[NP];
    for (int e = 0; e < o->nel_s; e++) {
        const float *m_1 = EL(o->M_1, e), *m_2 = EL(o->M_2, e), *m_3 = EL(o->M_3, e), *m_4 = EL(o->M_4, e);
        const float *m_5 = EL(o->M_5, e), *m_6 = EL(o->M_6, e), *m_7 = EL(o->M_7, e), *m_8 = EL(o->M_8, e);
        const float *m_w1 = EL(o->M_w1, e), *m_w2 = EL(o->M_w2, e), *m_w3 = EL(o->M_w3, e);
        const float *m_w4 = EL(o->M_w4, e), *m_w5 = EL(o->M_w5, e);
        const float *m11s = EL(o->M11s, e), *m21s = EL(o->M21s, e), *m41s = EL(o->M41s, e);
        const float *m12s = EL(o->M12s, e), *m22s = EL(o->M22s, e), *m32s = EL(o->M32s, e), *m42s = EL(o->M42s, e);
        const float *m11z = EL(o->M11z, e), *m21z = EL(o->M21z, e), *m41z = EL(o->M41z, e);
        const float *m1phi = EL(o->M1phi, e), *m2phi = EL(o->M2phi, e), *m4phi = EL(o->M4phi, e);
        memcpy(us, EL(u, e), sizeof us);
        memcpy(up, EL(u + cs, e), sizeof up);
        memcpy(uz, EL(u + 2 * cs, e), sizeof uz);
        if (!o->axis_s[e]) { mxm_4(o->G2T, us, X1); mxm_4(o->G2T, up, X2); mxm_4(o->G2T, uz, X3); }
        else               { mxm_4(o->G1T, us, X1); mxm_4(o->G1T, up, X2); mxm_4(o->G1T, uz, X3); }
        mxm_4(us, o->G2, X4); mxm_4(up, o->G2, X5); mxm_4(uz, o->G2, X6);
        FOR25 {
            float c1 = m_2[q] * X4[q], c2 = m_1[q] * X1[q], c3 = m_6[q] * X5[q];
            float c4 = m_5[q] * X2[q], c5 = m_4[q] * X6[q], c6 = m_3[q] * X3[q];
            ls[q] = c1 + c2 + 2 * (c3 + c4) + c5 + c6 + m_w1[q] * us[q] + m_w2[q] * up[q] + 2 * m_w3[q] * uz[q];
            lp[q] = -2 * (c1 + c2 + c5 + c6) - (c3 + c4) + m_w2[q] * us[q] + m_w4[q] * up[q] - m_w3[q] * uz[q];
            lz[q] = 2 * (m_8[q] * X5[q] + m_7[q] * X2[q]) + m_w3[q] * (2 * us[q] - up[q]) + m_w5[q] * uz[q];
            S1s[q] = m11s[q] * X4[q] + m21s[q] * X1[q] + m12s[q] * X6[q] + m22s[q] * X3[q] + m_1[q] * (us[q] - 2 * up[q]);
            S2s[q] = m11s[q] * X1[q] + m41s[q] * X4[q] + m32s[q] * X3[q] + m42s[q] * X6[q] + m_2[q] * (us[q] - 2 * up[q]);
            S1z[q] = m11z[q] * X6[q] + m21z[q] * X3[q] + m32s[q] * X4[q] + m22s[q] * X1[q] + m_3[q] * (us[q] - 2 * up[q]);
            S2z[q] = m11z[q] * X3[q] + m41z[q] * X6[q] + m12s[q] * X1[q] + m42s[q] * X4[q] + m_4[q] * (us[q] - 2 * up[q]);
            S1p[q] = m1phi[q] * X5[q] + m2phi[q] * X2[q] + m_5[q] * (2 * us[q] - up[q]) + 2 * m_7[q] * uz[q];
            S2p[q] = m1phi[q] * X2[q] + m4phi[q] * X5[q] + m_6[q] * (2 * us[q] - up[q]) + 2 * m_8[q] * uz[q];
        }
        mxm_4(S2s, o->G2T, X2); mxm_4(S2p, o->G2T, X4); mxm_4(S2z, o->G2T, X6);
        if (!o->axis_s[e]) { mxm_4(o->G2, S1s, X1); mxm_4(o->G2, S1p, X3); mxm_4(o->G2, S1z, X5); }
        else               { mxm_4(o->G1, S1s, X1); mxm_4(o->G1, S1p, X3); mxm_4(o->G1, S1z, X5); }
        FOR25 {
            ls[q] = ls[q] + X1[q] + X2[q];
            lp[q] = lp[q] + X3[q] + X4[q];
            lz[q] = lz[q] + X5[q] + X6[q];
        }
        if (o->axis_s[e]) {
            const float *w1 = EL0(o->M0_w1, e), *w2 = EL0(o->M0_w2, e), *w3 = EL0(o->M0_w3, e);
            const float *w4 = EL0(o->M0_w4, e), *w5 = EL0(o->M0_w5, e), *w6 = EL0(o->M0_w6, e);
            vxm_4(o->G0, us, V1); vxm_4(o->G0, up, V2); vxm_4(o->G0, uz, V3);
            for (int j = 0; j < NP; j++) W[j] = w1[j] * V1[j] + w2[j] * V2[j] + w3[j] * V3[j];
            outerprod_4(o->G0, W, S1s);
            for (int j = 0; j < NP; j++) W[j] = w2[j] * V1[j] + w4[j] * V2[j] + w5[j] * V3[j];
            outerprod_4(o->G0, W, S1p);
            for (int j = 0; j < NP; j++) W[j] = w3[j] * V1[j] + w5[j] * V2[j] + w6[j] * V3[j];
            outerprod_4(o->G0, W, S1z);
            FOR25 { ls[q] = ls[q] + S1s[q]; lp[q] = lp[q] + S1p[q]; lz[q] = lz[q] + S1z[q]; }
        }
        memcpy(EL(glob, e), ls, sizeof ls);
        memcpy(EL(glob + cs, e), lp, sizeof lp);
        memcpy(EL(glob + 2 * cs, e), lz, sizeof lz);
    }
}

/* stiffness_fluid.f90:139-216 */
static void glob_fluid_stiffness_4(const axo_t *o, float *glob, const float *chi) {
    float c[NPT], X1[NPT], X2[NPT], S1[NPT], S2[NPT], l[NPT], T[NPT], V1[NP], W[NP];
    for (int e = 0; e < o->nel_f; e++) {
        const float *m1 = EL(o->M1chi, e), *m2 = EL(o->M2chi, e), *m4 = EL(o->M4chi, e);
        memcpy(c, EL(chi, e), sizeof c);
        if (o->axis_f[e]) mxm_4(o->G1T, c, X1); else mxm_4(o->G2T, c, X1);
        mxm_4(c, o->G2, X2);
        FOR25 { S1[q] = m1[q] * X2[q] + m2[q] * X1[q]; S2[q] = m1[q] * X1[q] + m4[q] * X2[q]; }
        if (o->axis_f[e]) mxm_4(o->G1, S1, X1); else mxm_4(o->G2, S1, X1);
        mxm_4(S2, o->G2T, X2);
        FOR25 l[q] = X1[q] + X2[q];
        if (o->src_order != AXB_MONOPOLE) {
            const float *mw = EL(o->M_w_fl, e);
            FOR25 l[q] = l[q] + mw[q] * c[q];
            if (o->axis_f[e]) {
                const float *m0 = EL0(o->M0_w_fl, e);
                vxm_4(o->G0, c, V1);
                for (int j = 0; j < NP; j++) W[j] = m0[j] * V1[j];
                outerprod_4(o->G0, W, T);
                FOR25 l[q] = l[q] + T[q];
            }
        }
        memcpy(EL(glob, e), l, sizeof l);
    }
}

/* ------------------------------------------------------------------------------------ */
/* anelastic stiffness: stiffness_mono.f90:409-589, stiffness_di.f90:604-825,
 * stiffness_quad.f90:555-774 */
#define MV_CG(R, k, v, j, e) (R)[(k) + 4 * ((v) + 6 * ((j) + (size_t)o->n_sls * (e)))]
#define MV_FULL(R, q, v, j, e) (R)[(q) + NPT * ((v) + 6 * ((j) + (size_t)o->n_sls * (e)))]

static void glob_anel_stiffness_cg4(const axo_t *o, float *glob, const float *R) {
    const size_t cs = (size_t)NPT * o->nel_s;
    static const int pidx[4] = {1 + NP * 1, 1 + NP * 3, 3 + NP * 1, 3 + NP * 3}; /* (1,1),(1,3),(3,1),(3,3) */
    float X1[NPT], X2[NPT], X3[NPT], X4[NPT], X5[NPT], X6[NPT];
    for (int e = 0; e < o->nel_s; e++) {
        const float *yl = o->Ycg + 4 * (size_t)e, *vse = o->Vse_cg + 4 * (size_t)e, *vsx = o->Vsx_cg + 4 * (size_t)e;
        const float *vze = o->Vze_cg + 4 * (size_t)e, *vzx = o->Vzx_cg + 4 * (size_t)e;
        float r[6][4];
        const float *GA = o->axis_s[e] ? o->G1 : o->G2;
        for (int v = 0; v < 6; v++) for (int k = 0; k < 4; k++) r[v][k] = 0.0f;
        for (int j = 0; j < o->n_sls; j++)
            for (int v = 0; v < 6; v++) {
                if (o->src_order == AXB_MONOPOLE && (v == 3 || v == 5)) continue;
                for (int k = 0; k < 4; k++) r[v][k] = r[v][k] + MV_CG(R, k, v, j, e);
            }
        const float *r1 = r[0], *r2 = r[1], *r3 = r[2], *r4 = r[3], *r5 = r[4], *r6 = r[5];
        float S1a[4], S2a[4], S1b[4], S2b[4], S1z[4], S2z[4];
        float *g1 = EL(glob, e), *g2 = EL(glob + cs, e), *g3 = EL(glob + 2 * cs, e);
        if (o->src_order == AXB_MONOPOLE) {
            for (int k = 0; k < 4; k++) {
                S1a[k] = vze[k] * r1[k] + vse[k] * r5[k];
                S2a[k] = vzx[k] * r1[k] + vsx[k] * r5[k];
                S1z[k] = vze[k] * r5[k] + vse[k] * r3[k];
                S2z[k] = vzx[k] * r5[k] + vsx[k] * r3[k];
            }
            mxm_cg4_sparse_b(GA, S1a, X1); mxm_cg4_sparse_b(GA, S1z, X3);
            mxm_cg4_sparse_a(S2a, o->G2T, X2); mxm_cg4_sparse_a(S2z, o->G2T, X4);
            float ls[NPT], lz[NPT];
            FOR25 { ls[q] = X1[q] + X2[q]; lz[q] = X3[q] + X4[q]; }
            for (int k = 0; k < 4; k++) ls[pidx[k]] = ls[pidx[k]] + yl[k] * r2[k];
            FOR25 { g1[q] = g1[q] - ls[q]; g3[q] = g3[q] - lz[q]; }
        } else if (o->src_order == AXB_DIPOLE) {
            for (int k = 0; k < 4; k++) {
                S1a[k] = vze[k] * (r1[k] - r6[k]) + vse[k] * (r5[k] - r4[k]);
                S2a[k] = vzx[k] * (r1[k] - r6[k]) + vsx[k] * (r5[k] - r4[k]);
                S1b[k] = vze[k] * (r1[k] + r6[k]) + vse[k] * (r5[k] + r4[k]);
                S2b[k] = vzx[k] * (r1[k] + r6[k]) + vsx[k] * (r5[k] + r4[k]);
                S1z[k] = vze[k] * r5[k] + vse[k] * r3[k];
                S2z[k] = vzx[k] * r5[k] + vsx[k] * r3[k];
            }
            mxm_cg4_sparse_b(GA, S1a, X1); mxm_cg4_sparse_b(GA, S1b, X3); mxm_cg4_sparse_b(GA, S1z, X5);
            mxm_cg4_sparse_a(S2a, o->G2T, X2); mxm_cg4_sparse_a(S2b, o->G2T, X4); mxm_cg4_sparse_a(S2z, o->G2T, X6);
            float lp[NPT], lm[NPT], lz[NPT];
            FOR25 { lp[q] = X1[q] + X2[q]; lm[q] = X3[q] + X4[q]; lz[q] = X5[q] + X6[q]; }
            for (int k = 0; k < 4; k++) {
                lm[pidx[k]] = lm[pidx[k]] + 2 * yl[k] * (r2[k] - r6[k]);
                lz[pidx[k]] = lz[pidx[k]] - yl[k] * r4[k];
            }
            FOR25 { g1[q] = g1[q] - lp[q]; g2[q] = g2[q] - lm[q]; g3[q] = g3[q] - lz[q]; }
        } else {
            for (int k = 0; k < 4; k++) {
                S1a[k] = vze[k] * r1[k] + vse[k] * r5[k];
                S2a[k] = vzx[k] * r1[k] + vsx[k] * r5[k];
                S1b[k] = vze[k] * r6[k] + vse[k] * r4[k];
                S2b[k] = vzx[k] * r6[k] + vsx[k] * r4[k];
                S1z[k] = vze[k] * r5[k] + vse[k] * r3[k];
                S2z[k] = vzx[k] * r5[k] + vsx[k] * r3[k];
            }
            mxm_cg4_sparse_b(GA, S1a, X1); mxm_cg4_sparse_b(GA, S1b, X3); mxm_cg4_sparse_b(GA, S1z, X5);
            mxm_cg4_sparse_a(S2a, o->G2T, X2); mxm_cg4_sparse_a(S2b, o->G2T, X4); mxm_cg4_sparse_a(S2z, o->G2T, X6);
            float ls[NPT], lp[NPT], lz[NPT];
            FOR25 { ls[q] = X1[q] + X2[q]; lp[q] = -X3[q] - X4[q]; lz[q] = X5[q] + X6[q]; }
            for (int k = 0; k < 4; k++) {
                ls[pidx[k]] = ls[pidx[k]] + yl[k] * (r2[k] - 2 * r6[k]);
                lp[pidx[k]] = lp[pidx[k]] + yl[k] * (r6[k] - 2 * r2[k]);
                lz[pidx[k]] = lz[pidx[k]] - 2 * yl[k] * r4[k];
            }
            FOR25 { g1[q] = g1[q] - ls[q]; g2[q] = g2[q] - lp[q]; g3[q] = g3[q] - lz[q]; }
        }
    }
}

static void glob_anel_stiffness_4(const axo_t *o, float *glob, const float *R) {
    const size_t cs = (size_t)NPT * o->nel_s;
    float r[6][NPT], X1[NPT], X2[NPT], X3[NPT], X4[NPT], X5[NPT], X6[NPT];
    float S1a[NPT], S2a[NPT], S1b[NPT], S2b[NPT], S1z[NPT], S2z[NPT], T[NPT];
    float V1[NP], V2[NP], V3[NP], V4[NP];
    for (int e = 0; e < o->nel_s; e++) {
        const float *yl = EL(o->Y, e), *vse = EL(o->Vse, e), *vsx = EL(o->Vsx, e);
        const float *vze = EL(o->Vze, e), *vzx = EL(o->Vzx, e);
        const float *GA = o->axis_s[e] ? o->G1 : o->G2;
        for (int v = 0; v < 6; v++) FOR25 r[v][q] = 0.0f;
        for (int j = 0; j < o->n_sls; j++)
            for (int v = 0; v < 6; v++) {
                if (o->src_order == AXB_MONOPOLE && (v == 3 || v == 5)) continue;
                FOR25 r[v][q] = r[v][q] + MV_FULL(R, q, v, j, e);
            }
        const float *r1 = r[0], *r2 = r[1], *r3 = r[2], *r4 = r[3], *r5 = r[4], *r6 = r[5];
        float *g1 = EL(glob, e), *g2 = EL(glob + cs, e), *g3 = EL(glob + 2 * cs, e);
        const float *y0 = EL0(o->Y0, e), *v0se = EL0(o->V0se, e), *v0sx = EL0(o->V0sx, e);
        const float *v0ze = EL0(o->V0ze, e), *v0zx = EL0(o->V0zx, e);
#define R0(a, j) (a)[0 + NP * (j)]
        if (o->src_order == AXB_MONOPOLE) {
            float ls[NPT], lz[NPT];
            FOR25 {
                S1a[q] = vze[q] * r1[q] + vse[q] * r5[q];
                S2a[q] = vzx[q] * r1[q] + vsx[q] * r5[q];
                S1z[q] = vze[q] * r5[q] + vse[q] * r3[q];
                S2z[q] = vzx[q] * r5[q] + vsx[q] * r3[q];
            }
            mxm_4(GA, S1a, X1); mxm_4(GA, S1z, X3);
            mxm_4(S2a, o->G2T, X2); mxm_4(S2z, o->G2T, X4);
            FOR25 { ls[q] = X1[q] + X2[q] + yl[q] * r2[q]; lz[q] = X3[q] + X4[q]; }
            if (o->axis_s[e]) {
                for (int j = 0; j < NP; j++)
                    V1[j] = v0ze[j] * R0(r1, j) + v0se[j] * R0(r5, j) + y0[j] * R0(r2, j);
                outerprod_4(o->G0, V1, T);
                FOR25 ls[q] = ls[q] + T[q];
                for (int j = 0; j < NP; j++) {
                    V2[j] = v0ze[j] * R0(r5, j) + v0se[j] * R0(r3, j);
                    V3[j] = v0zx[j] * R0(r5, j) + v0sx[j] * R0(r3, j);
                }
                vxm_4(V3, o->G2T, V4);
                outerprod_4(o->G0, V2, T);
                FOR25 lz[q] = lz[q] + T[q];
                for (int j = 0; j < NP; j++) lz[0 + NP * j] = lz[0 + NP * j] + V4[j];
            }
            FOR25 { g1[q] = g1[q] - ls[q]; g3[q] = g3[q] - lz[q]; }
        } else if (o->src_order == AXB_DIPOLE) {
            float lp[NPT], lm[NPT], lz[NPT];
            FOR25 {
                S1a[q] = vze[q] * (r1[q] - r6[q]) + vse[q] * (r5[q] - r4[q]);
                S2a[q] = vzx[q] * (r1[q] - r6[q]) + vsx[q] * (r5[q] - r4[q]);
                S1b[q] = vze[q] * (r1[q] + r6[q]) + vse[q] * (r5[q] + r4[q]);
                S2b[q] = vzx[q] * (r1[q] + r6[q]) + vsx[q] * (r5[q] + r4[q]);
                S1z[q] = vze[q] * r5[q] + vse[q] * r3[q];
                S2z[q] = vzx[q] * r5[q] + vsx[q] * r3[q];
            }
            mxm_4(GA, S1a, X1); mxm_4(GA, S1b, X3); mxm_4(GA, S1z, X5);
            mxm_4(S2a, o->G2T, X2); mxm_4(S2b, o->G2T, X4); mxm_4(S2z, o->G2T, X6);
            FOR25 {
                lp[q] = X1[q] + X2[q];
                lm[q] = X3[q] + X4[q] + 2 * yl[q] * (r2[q] - r6[q]);
                lz[q] = X5[q] + X6[q] - yl[q] * r4[q];
            }
            if (o->axis_s[e]) {
                for (int j = 0; j < NP; j++) {
                    V1[j] = v0ze[j] * (R0(r1, j) - R0(r6, j)) + v0se[j] * (R0(r5, j) - R0(r4, j));
                    V2[j] = v0zx[j] * (R0(r1, j) - R0(r6, j)) + v0sx[j] * (R0(r5, j) - R0(r4, j));
                }
                vxm_4(V2, o->G2T, V3);
                outerprod_4(o->G0, V1, T);
                FOR25 lp[q] = lp[q] + T[q];
                for (int j = 0; j < NP; j++) lp[0 + NP * j] = lp[0 + NP * j] + V3[j];
                for (int j = 0; j < NP; j++)
                    V1[j] = v0ze[j] * (R0(r1, j) + R0(r6, j)) + v0se[j] * (R0(r5, j) + R0(r4, j))
                          + y0[j] * 2 * (R0(r2, j) - R0(r6, j));
                outerprod_4(o->G0, V1, T);
                FOR25 lm[q] = lm[q] + T[q];
                /* reference quirk (stiffness_di.f90:710-711): V1 is computed but V2 is added */
                outerprod_4(o->G0, V2, T);
                FOR25 lz[q] = lz[q] + T[q];
            }
            FOR25 { g1[q] = g1[q] - lp[q]; g2[q] = g2[q] - lm[q]; g3[q] = g3[q] - lz[q]; }
        } else {
            float ls[NPT], lp[NPT], lz[NPT];
            FOR25 {
                S1a[q] = vze[q] * r1[q] + vse[q] * r5[q];
                S2a[q] = vzx[q] * r1[q] + vsx[q] * r5[q];
                S1b[q] = vze[q] * r6[q] + vse[q] * r4[q];
                S2b[q] = vzx[q] * r6[q] + vsx[q] * r4[q];
                S1z[q] = vze[q] * r5[q] + vse[q] * r3[q];
                S2z[q] = vzx[q] * r5[q] + vsx[q] * r3[q];
            }
            mxm_4(GA, S1a, X1); mxm_4(GA, S1b, X3); mxm_4(GA, S1z, X5);
            mxm_4(S2a, o->G2T, X2); mxm_4(S2b, o->G2T, X4); mxm_4(S2z, o->G2T, X6);
            FOR25 {
                ls[q] = X1[q] + X2[q] + yl[q] * (r2[q] - 2 * r6[q]);
                lp[q] = -X3[q] - X4[q] + yl[q] * (r6[q] - 2 * r2[q]);
                lz[q] = X5[q] + X6[q] - 2 * yl[q] * r4[q];
            }
            if (o->axis_s[e]) {
                for (int j = 0; j < NP; j++) V1[j] = v0ze[j] * R0(r1, j) + y0[j] * (R0(r2, j) - 2 * R0(r6, j));
                outerprod_4(o->G0, V1, T);
                FOR25 ls[q] = ls[q] + T[q];
                for (int j = 0; j < NP; j++) V1[j] = -v0ze[j] * R0(r6, j) + y0[j] * (R0(r6, j) - 2 * R0(r2, j));
                outerprod_4(o->G0, V1, T);
                FOR25 lp[q] = lp[q] + T[q];
                for (int j = 0; j < NP; j++) V1[j] = v0se[j] * R0(r3, j);
                outerprod_4(o->G0, V1, T);
                FOR25 lz[q] = lz[q] + T[q];
            }
            FOR25 { g1[q] = g1[q] - ls[q]; g2[q] = g2[q] - lp[q]; g3[q] = g3[q] - lz[q]; }
        }
#undef R0
    }
}

/* ------------------------------------------------------------------------------------ */
/* attenuation.f90:1139-1155 */
static void fast_correct(int n, const double *y, double *yp) {
    double dy[32];
    dy[0] = 1 + .5 * y[0];
    for (int k = 1; k < n; k++) dy[k] = dy[k - 1] + (dy[k - 1] - .5) * y[k - 1] + .5 * y[k];
    for (int k = 0; k < n; k++) yp[k] = y[k] * dy[k];
}
static void a_j_of_Q(const axo_t *o, float Q, double *a_j) {
    double yq[32], yp[32], s = 0.0;
    for (int k = 0; k < o->n_sls; k++) yq[k] = o->y_j[k] / Q;
    if (o->corr_lowq) fast_correct(o->n_sls, yq, yp);
    else memcpy(yp, yq, sizeof(double) * o->n_sls);
    for (int k = 0; k < o->n_sls; k++) s += yp[k];
    for (int k = 0; k < o->n_sls; k++) a_j[k] = yp[k] / s;
}

/* pointwise_derivatives.f90:216-248 */
static void gradient_cg4(const axo_t *o, const float *f, float *ds, float *dz, int e) {
    float m1[4], m2[4];
    const float *dzdeta = o->Dze_cg + 4 * (size_t)e, *dzdxi = o->Dzx_cg + 4 * (size_t)e;
    const float *dsdeta = o->Dse_cg + 4 * (size_t)e, *dsdxi = o->Dsx_cg + 4 * (size_t)e;
    mxm_cg4_sparse_c(o->axis_s[e] ? o->G1T : o->G2T, f, m1);
    mxm_cg4_sparse_c(f, o->G2, m2);
    for (int k = 0; k < 4; k++) {
        ds[k] = dzdeta[k] * m1[k] + dzdxi[k] * m2[k];
        dz[k] = dsdeta[k] * m1[k] + dsdxi[k] * m2[k];
    }
}
/* pointwise_derivatives.f90:55-73 */
static void f_over_s_cg4(const axo_t *o, const float *f, float *out, int e) {
    const float *is = EL(o->inv_s, e);
    out[0] = is[1 + NP * 1] * f[1 + NP * 1];
    out[1] = is[1 + NP * 3] * f[1 + NP * 3];
    out[2] = is[3 + NP * 1] * f[3 + NP * 1];
    out[3] = is[3 + NP * 3] * f[3 + NP * 3];
}

/* attenuation.f90:471-535 */
static void compute_strain_att_el_cg4(const axo_t *o, const float *u1, const float *u2,
                                      const float *u3, float g[6][4], int e) {
    float b1s[4], b1z[4], b2s[4], b2z[4], T[NPT], fs[4], fs2[4];
    for (int v = 0; v < 6; v++) for (int k = 0; k < 4; k++) g[v][k] = 0.0f;
    if (o->src_order == AXB_DIPOLE) { FOR25 T[q] = u1[q] + u2[q]; gradient_cg4(o, T, b1s, b1z, e); }
    else gradient_cg4(o, u1, b1s, b1z, e);
    gradient_cg4(o, u3, b2s, b2z, e);
    for (int k = 0; k < 4; k++) { g[0][k] = b1s[k]; g[2][k] = b2z[k]; g[4][k] = b1z[k] + b2s[k]; }
    if (o->src_order == AXB_MONOPOLE) {
        f_over_s_cg4(o, u1, fs, e);
        for (int k = 0; k < 4; k++) g[1][k] = fs[k];
    } else if (o->src_order == AXB_DIPOLE) {
        f_over_s_cg4(o, u2, fs, e);
        for (int k = 0; k < 4; k++) g[1][k] = 2 * fs[k];
        FOR25 T[q] = u1[q] - u2[q];
        gradient_cg4(o, T, b1s, b1z, e);
        f_over_s_cg4(o, u3, fs, e);
        for (int k = 0; k < 4; k++) { g[3][k] = -fs[k] - b1z[k]; g[5][k] = -g[1][k] - b1s[k]; }
    } else {
        FOR25 T[q] = u1[q] - 2 * u2[q];
        f_over_s_cg4(o, T, fs, e);
        for (int k = 0; k < 4; k++) g[1][k] = fs[k];
        gradient_cg4(o, u2, b1s, b1z, e);
        f_over_s_cg4(o, u3, fs, e);
        FOR25 T[q] = u2[q] - 2 * u1[q];
        f_over_s_cg4(o, T, fs2, e);
        for (int k = 0; k < 4; k++) { g[3][k] = -2 * fs[k] - b1z[k]; g[5][k] = fs2[k] - b1s[k]; }
    }
}

/* pointwise_derivatives.f90:252-286 */
static void gradient_4(const axo_t *o, const float *f, float *ds, float *dz, int e) {
    float m1[NPT], m2[NPT];
    const float *dzdeta = EL(o->Dze, e), *dzdxi = EL(o->Dzx, e);
    const float *dsdeta = EL(o->Dse, e), *dsdxi = EL(o->Dsx, e);
    mxm_4(o->axis_s[e] ? o->G1T : o->G2T, f, m1);
    mxm_4(f, o->G2, m2);
    FOR25 {
        ds[q] = dzdeta[q] * m1[q] + dzdxi[q] * m2[q];
        dz[q] = dsdeta[q] * m1[q] + dsdxi[q] * m2[q];
    }
}
/* pointwise_derivatives.f90:104-126 (+ dsdf_elem_solid :419-446) */
static void f_over_s_4(const axo_t *o, const float *f, float *out, int e) {
    const float *is = EL(o->inv_s, e);
    FOR25 out[q] = is[q] * f[q];
    if (o->axis_s[e]) {
        float ds[NPT], dz[NPT];
        gradient_4(o, f, ds, dz, e);
        for (int j = 0; j < NP; j++) out[0 + NP * j] = ds[0 + NP * j];
    }
}
/* attenuation.f90:542-606 */
static void compute_strain_att_el_4(const axo_t *o, const float *u1, const float *u2,
                                    const float *u3, float g[6][NPT], int e) {
    float b1s[NPT], b1z[NPT], b2s[NPT], b2z[NPT], T[NPT], fs[NPT], fs2[NPT];
    for (int v = 0; v < 6; v++) FOR25 g[v][q] = 0.0f;
    if (o->src_order == AXB_DIPOLE) { FOR25 T[q] = u1[q] + u2[q]; gradient_4(o, T, b1s, b1z, e); }
    else gradient_4(o, u1, b1s, b1z, e);
    gradient_4(o, u3, b2s, b2z, e);
    FOR25 { g[0][q] = b1s[q]; g[2][q] = b2z[q]; g[4][q] = b1z[q] + b2s[q]; }
    if (o->src_order == AXB_MONOPOLE) {
        f_over_s_4(o, u1, fs, e);
        FOR25 g[1][q] = fs[q];
    } else if (o->src_order == AXB_DIPOLE) {
        f_over_s_4(o, u2, fs, e);
        FOR25 g[1][q] = 2 * fs[q];
        FOR25 T[q] = u1[q] - u2[q];
        gradient_4(o, T, b1s, b1z, e);
        f_over_s_4(o, u3, fs, e);
        FOR25 { g[3][q] = -fs[q] - b1z[q]; g[5][q] = -g[1][q] - b1s[q]; }
    } else {
        FOR25 T[q] = u1[q] - 2 * u2[q];
        f_over_s_4(o, T, fs, e);
        FOR25 g[1][q] = fs[q];
        gradient_4(o, u2, b1s, b1z, e);
        f_over_s_4(o, u3, fs, e);
        FOR25 T[q] = u2[q] - 2 * u1[q];
        f_over_s_4(o, T, fs2, e);
        FOR25 { g[3][q] = -2 * fs[q] - b1z[q]; g[5][q] = fs2[q] - b1s[q]; }
    }
}

/* attenuation.f90:81-202 (cg4) and :210-334 (_4): the two differ only in the number of
 * points per element (np = 4 or 25). */
static void time_step_memvars(axo_t *o) {
    const size_t cs = (size_t)NPT * o->nel_s;
    const double third = 1.0 / 3.0;
    double a_mu[32], a_ka[32];
    float Qmu_last = -1.0f, Qka_last = -1.0f;   /* cg4 leaves them uninitialised; treat as -1 */
    const int np = o->cg ? 4 : NPT;
    for (int e = 0; e < o->nel_s; e++) {
        const float *u1 = EL(o->disp, e), *u2 = EL(o->disp + cs, e), *u3 = EL(o->disp + 2 * cs, e);
        float gr[6][NPT];
        if (o->cg) {
            float g4[6][4];
            compute_strain_att_el_cg4(o, u1, u2, u3, g4, e);
            for (int v = 0; v < 6; v++) for (int k = 0; k < 4; k++) gr[v][k] = g4[v][k];
        } else {
            compute_strain_att_el_4(o, u1, u2, u3, gr, e);
        }
        if (o->Q_mu[e] != Qmu_last) { Qmu_last = o->Q_mu[e]; a_j_of_Q(o, Qmu_last, a_mu); }
        if (o->Q_kappa[e] != Qka_last) { Qka_last = o->Q_kappa[e]; a_j_of_Q(o, Qka_last, a_ka); }
        const float *dmu = o->cg ? o->dmu_cg + 4 * (size_t)e : EL(o->dmu, e);
        const float *dka = o->cg ? o->dka_cg + 4 * (size_t)e : EL(o->dka, e);
        float *tr_tm1 = o->src_tr_tm1 + (size_t)np * e;
        float *dev_tm1 = o->src_dev_tm1 + (size_t)np * 6 * e;
        for (int k = 0; k < np; k++) {
            /* trace: sum(grad(:,1:3), dim=2) in real(4), left to right */
            float trace = gr[0][k] + gr[1][k];
            trace = trace + gr[2][k];
            float src_tr_t = dka[k] * trace;
            float src_dev_t[6] = {0, 0, 0, 0, 0, 0};
            for (int v = 0; v < 3; v++)
                src_dev_t[v] = (float)((double)(dmu[k] * 2) * ((double)gr[v][k] - (double)trace * third));
            src_dev_t[4] = dmu[k] * gr[4][k];
            if (o->src_order != AXB_MONOPOLE) {
                src_dev_t[3] = dmu[k] * gr[3][k];
                src_dev_t[5] = dmu[k] * gr[5][k];
            }
            float s_tr_tm1 = tr_tm1[k];
            float s_dev_tm1[6];
            for (int v = 0; v < 6; v++) s_dev_tm1[v] = dev_tm1[k + np * v];
            for (int j = 0; j < o->n_sls; j++) {
                float tr_buf = (float)(o->ts_t[j] * a_ka[j] * (double)src_tr_t
                                       + o->ts_tm1[j] * a_ka[j] * (double)s_tr_tm1);
                float dev_buf[6];
                for (int v = 0; v < 6; v++)
                    dev_buf[v] = (float)(o->ts_t[j] * a_mu[j] * (double)src_dev_t[v]
                                         + o->ts_tm1[j] * a_mu[j] * (double)s_dev_tm1[v]);
                size_t base = (size_t)np * 6 * ((size_t)j + (size_t)o->n_sls * e);
                float *mv = o->memvar + base;
                for (int v = 0; v < 3; v++)
                    mv[k + np * v] = (float)(o->exp_w[j] * (double)mv[k + np * v]
                                             + (double)dev_buf[v] + (double)tr_buf);
                mv[k + np * 4] = (float)(o->exp_w[j] * (double)mv[k + np * 4] + (double)dev_buf[4]);
                if (o->src_order != AXB_MONOPOLE) {
                    mv[k + np * 3] = (float)(o->exp_w[j] * (double)mv[k + np * 3] + (double)dev_buf[3]);
                    mv[k + np * 5] = (float)(o->exp_w[j] * (double)mv[k + np * 5] + (double)dev_buf[5]);
                }
            }
            tr_tm1[k] = src_tr_t;
            for (int v = 0; v < 6; v++) dev_tm1[k + np * v] = src_dev_t[v];
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* apply_masks.f90:40-100 */
static void axis_mask(float *u, size_t cs, const int *ax, int nax, int c_lo, int c_hi) {
    for (int a = 0; a < nax; a++)
        for (int c = c_lo; c <= c_hi; c++)
            for (int j = 0; j < NP; j++) u[0 + NP * j + (size_t)NPT * (ax[a] - 1) + cs * c] = 0.0f;
}
static void mask_solid(const axo_t *o, float *u) {
    size_t cs = (size_t)NPT * o->nel_s;
    if (o->src_order == AXB_MONOPOLE) axis_mask(u, cs, o->ax_el_s, o->naxel_s, 0, 0);
    else if (o->src_order == AXB_DIPOLE) axis_mask(u, cs, o->ax_el_s, o->naxel_s, 1, 2);
    else axis_mask(u, cs, o->ax_el_s, o->naxel_s, 0, 2);
}

/* time_evol_wave.F90:1532-1571 */
static void bdry_copy2fluid(const axo_t *o, float *uflu, const float *usol) {
    size_t cs = (size_t)NPT * o->nel_s;
    for (int b = 0; b < o->nel_bdry; b++) {
        int js = o->bdry_js[b], jf = o->bdry_jf[b], es = o->bdry_sel[b] - 1, ef = o->bdry_fel[b] - 1;
        const float *B1 = o->bdry_matr + (size_t)NP * b, *B2 = o->bdry_matr + (size_t)NP * (b + o->nel_bdry);
        for (int i = 0; i < NP; i++) {
            size_t ps = i + NP * js + (size_t)NPT * es, pf = i + NP * jf + (size_t)NPT * ef;
            if (o->src_order == AXB_DIPOLE)
                uflu[pf] = uflu[pf] - B1[i] * (usol[ps] + usol[ps + cs]) - B2[i] * usol[ps + 2 * cs];
            else
                uflu[pf] = uflu[pf] - B1[i] * usol[ps] - B2[i] * usol[ps + 2 * cs];
        }
    }
}
/* time_evol_wave.F90:1577-1611 */
static void bdry_copy2solid(const axo_t *o, float *usol, const float *uflu) {
    size_t cs = (size_t)NPT * o->nel_s;
    for (int b = 0; b < o->nel_bdry; b++) {
        int js = o->bdry_js[b], jf = o->bdry_jf[b], es = o->bdry_sel[b] - 1, ef = o->bdry_fel[b] - 1;
        const float *B1 = o->bdry_matr + (size_t)NP * b, *B2 = o->bdry_matr + (size_t)NP * (b + o->nel_bdry);
        for (int i = 0; i < NP; i++) {
            size_t ps = i + NP * js + (size_t)NPT * es, pf = i + NP * jf + (size_t)NPT * ef;
            usol[ps] = usol[ps] + B1[i] * uflu[pf];
            if (o->src_order == AXB_DIPOLE) usol[ps + cs] = usol[ps + cs] + B1[i] * uflu[pf];
            usol[ps + 2 * cs] = usol[ps + 2 * cs] + B2[i] * uflu[pf];
        }
    }
}

/* time_evol_wave.F90:1062-1097 */
static void add_source_el(const axo_t *o, float *acc, float stf1) {
    size_t cs = (size_t)NPT * o->nel_s;
    if (stf1 == 0.0f) return;
    for (int k = 0; k < o->nelsrc; k++)
        for (int c = 0; c < 3; c++)
            FOR25 {
                size_t p = q + (size_t)NPT * (o->ielsrc[k] - 1) + cs * c;
                acc[p] = acc[p] - o->src_term[q + NPT * (k + 8 * c)] * stf1;
            }
}
static void add_source_fl(const axo_t *o, float *ddchi, float stf1) {
    if (stf1 == 0.0f) return;
    for (int k = 0; k < o->nelsrc; k++)
        FOR25 {
            size_t p = q + (size_t)NPT * (o->ielsrc[k] - 1);
            ddchi[p] = ddchi[p] - o->src_term[q + NPT * k] * stf1;
        }
}

/* ------------------------------------------------------------------------------------ */
/* commun.F90:69-283 + commpi.F90:371-637.  nc = 3 for the solid (the time loop always
 * passes the 3-wide array: size(vec,dim=4)), 1 for the fluid. */
static const int edge_q[16] = {0, 1, 2, 3, 4, 5, 9, 10, 14, 15, 19, 20, 21, 22, 23, 24};

static void feed_buffer(axo_t *o, int dom, const float *vec) {
    halo_t *H = &o->halo[dom];
    int nc = dom == AXB_DOMAIN_SOLID ? 3 : 1;
    int nel = dom == AXB_DOMAIN_SOLID ? o->nel_s : o->nel_f;
    int nglob = dom == AXB_DOMAIN_SOLID ? o->nglob_s : o->nglob_f;
    const int *igloc = dom == AXB_DOMAIN_SOLID ? o->igloc_s : o->igloc_f;
    float *gvec = dom == AXB_DOMAIN_SOLID ? o->gvec_s : o->gvec_f;
    size_t cs = (size_t)NPT * nel;
    memset(gvec, 0, sizeof(float) * (size_t)nglob * nc);
    for (int ip = 0; ip < H->ncomm; ip++) {
        int ipol = H->glob2el[ip], jpol = H->glob2el[ip + H->ncomm], iel = H->glob2el[ip + 2 * H->ncomm];
        size_t ipt = (size_t)(iel - 1) * NPT + jpol * NP + ipol;
        int ipg = igloc[ipt] - 1;
        for (int c = 0; c < nc; c++)
            gvec[ipg + (size_t)nglob * c] = gvec[ipg + (size_t)nglob * c] + vec[ipt + cs * c];
    }
    for (int m = 0; m < H->nmsg; m++)
        for (int ip = 0; ip < H->size[m]; ip++) {
            int ipg = H->glocal[m][ip] - 1;
            for (int c = 0; c < nc; c++)
                H->sendbuf[m][ip + (size_t)H->size[m] * c] = gvec[ipg + (size_t)nglob * c];
        }
}

static void gather_scatter(axo_t *o, int dom, float *vec) {
    int nc = dom == AXB_DOMAIN_SOLID ? 3 : 1;
    int nel = dom == AXB_DOMAIN_SOLID ? o->nel_s : o->nel_f;
    int nglob = dom == AXB_DOMAIN_SOLID ? o->nglob_s : o->nglob_f;
    const int *igloc = dom == AXB_DOMAIN_SOLID ? o->igloc_s : o->igloc_f;
    float *gvec = dom == AXB_DOMAIN_SOLID ? o->gvec_s : o->gvec_f;
    size_t cs = (size_t)NPT * nel;
    memset(gvec, 0, sizeof(float) * (size_t)nglob * nc);
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < 16; k++) {
            size_t ipt = (size_t)e * NPT + edge_q[k];
            int id = igloc[ipt] - 1;
            for (int c = 0; c < nc; c++)
                gvec[id + (size_t)nglob * c] = gvec[id + (size_t)nglob * c] + vec[ipt + cs * c];
        }
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < 16; k++) {
            size_t ipt = (size_t)e * NPT + edge_q[k];
            int id = igloc[ipt] - 1;
            for (int c = 0; c < nc; c++) vec[ipt + cs * c] = gvec[id + (size_t)nglob * c];
        }
}

/* "MPI": copy my send buffers into the peers' receive buffers (message m of the peer
 * that lists me as its neighbour) */
static int exchange(axo_t *o, int dom) {
    halo_t *H = &o->halo[dom];
    int nc = dom == AXB_DOMAIN_SOLID ? 3 : 1;
    for (int m = 0; m < H->nmsg; m++) {
        axo_t *p = NULL;
        for (int g = 0; g < o->ngroup; g++) if (o->group[g]->rank == H->peer[m]) p = o->group[g];
        if (!p) return fail("halo peer not connected");
        halo_t *Hp = &p->halo[dom];
        int mm = -1;
        for (int q = 0; q < Hp->nmsg; q++) if (Hp->peer[q] == o->rank) mm = q;
        if (mm < 0 || Hp->size[mm] != H->size[m]) return fail("halo lists inconsistent");
        memcpy(Hp->recvbuf[mm], H->sendbuf[m], sizeof(float) * (size_t)H->size[m] * nc);
    }
    return 0;
}

static void extract_from_buffer(axo_t *o, int dom, float *vec) {
    halo_t *H = &o->halo[dom];
    int nc = dom == AXB_DOMAIN_SOLID ? 3 : 1;
    int nel = dom == AXB_DOMAIN_SOLID ? o->nel_s : o->nel_f;
    int nglob = dom == AXB_DOMAIN_SOLID ? o->nglob_s : o->nglob_f;
    const int *igloc = dom == AXB_DOMAIN_SOLID ? o->igloc_s : o->igloc_f;
    float *gvec = dom == AXB_DOMAIN_SOLID ? o->gvec_s : o->gvec_f;
    size_t cs = (size_t)NPT * nel;
    for (int m = 0; m < H->nmsg; m++)
        for (int ip = 0; ip < H->size[m]; ip++) {
            int ipg = H->glocal[m][ip] - 1;
            for (int c = 0; c < nc; c++)
                gvec[ipg + (size_t)nglob * c] = gvec[ipg + (size_t)nglob * c]
                                                + H->recvbuf[m][ip + (size_t)H->size[m] * c];
        }
    for (int ip = 0; ip < H->ncomm; ip++) {
        int ipol = H->glob2el[ip], jpol = H->glob2el[ip + H->ncomm], iel = H->glob2el[ip + 2 * H->ncomm];
        size_t ipt = (size_t)(iel - 1) * NPT + jpol * NP + ipol;
        int ipg = igloc[ipt] - 1;
        for (int c = 0; c < nc; c++) vec[ipt + cs * c] = gvec[ipg + (size_t)nglob * c];
    }
}

/* ------------------------------------------------------------------------------------ */
/* seismograms.f90:783-820 */
static void sample_receivers(axo_t *o) {
    size_t cs = (size_t)NPT * o->nel_s;
    float *out = o->recdump + (size_t)3 * o->num_rec * o->iseismo;
    for (int r = 0; r < o->num_rec; r++) {
        int iel = o->recfile_el[r], ip = o->recfile_el[r + o->num_rec], jp = o->recfile_el[r + 2 * o->num_rec];
        size_t p = ip + NP * jp + (size_t)NPT * (iel - 1);
        float d1 = o->disp[p], d2 = o->disp[p + cs], d3 = o->disp[p + 2 * cs];
        if (o->src_order == AXB_MONOPOLE)      { out[3 * r] = d1; out[3 * r + 1] = 0.0f; out[3 * r + 2] = d3; }
        else if (o->src_order == AXB_DIPOLE)   { out[3 * r] = d1 + d2; out[3 * r + 1] = d1 - d2; out[3 * r + 2] = d3; }
        else                                   { out[3 * r] = d1; out[3 * r + 1] = d2; out[3 * r + 2] = d3; }
    }
    o->iseismo++;
}

/* wavefields_io.f90:1019-1115 (displ_only + netcdf branch), :743-783 */
static void dump_disp_global(axo_t *o) {
    size_t cs = (size_t)NPT * o->nel_s;
    size_t npts = (size_t)o->npt_s_kwf + o->npt_f_kwf;
    size_t vs = npts * o->nstrain_max;
    float *base = o->snapdump + npts * o->istrain;
    for (int e = 0; e < o->nel_s; e++)
        FOR25 {
            size_t p = q + (size_t)NPT * e;
            if (!o->kwf_mask[p]) continue;
            int ct = o->kwf_map[p] - 1;
            float u1 = o->disp[p], u2 = o->disp[p + cs], u3 = o->disp[p + 2 * cs];
            float f1, f2;
            if (o->src_order == AXB_DIPOLE) { f1 = u1 + u2; f2 = u1 - u2; } else { f1 = u1; f2 = u2; }
            base[ct] = f1;
            if (o->src_order != AXB_MONOPOLE) base[ct + vs] = f2;
            base[ct + 2 * vs] = u3;
        }
    for (int e = 0; e < o->nel_f; e++) {
        float m1[NPT], m2[NPT];
        const float *f = EL(o->chi, e);
        mxm_4(o->axis_f[e] ? o->G1T : o->G2T, f, m1);
        mxm_4(f, o->G2, m2);
        FOR25 {
            size_t p = q + (size_t)NPT * e, pk = p + cs;
            if (!o->kwf_mask[pk]) continue;
            int ct = o->kwf_map[pk] - 1;   /* 1-based into [solid | fluid] */
            float dsdf = o->Dze_f[p] * m1[q] + o->Dzx_f[p] * m2[q];
            float dzdf = o->Dse_f[p] * m1[q] + o->Dsx_f[p] * m2[q];
            base[ct] = o->inv_rho_fluid[p] * dsdf;
            if (o->src_order != AXB_MONOPOLE) base[ct + vs] = 0.0f;   /* wavefields_io.f90:1080 */
            base[ct + 2 * vs] = o->inv_rho_fluid[p] * dzdf;
        }
    }
    o->istrain++;
}

/* ---- dump_type strain_only / fullfields ------------------------------------------------ */
/* axisym_gradient_solid / _fluid of one element (pointwise_derivatives.f90:329-365, :509-544) */
static void grad_el(const float *G1T, const float *G2T, const float *G2, int axial, const float *f,
                    const float *dse, const float *dze, const float *dsx, const float *dzx,
                    float *ds, float *dz) {
    float m1[NPT], m2[NPT];
    mxm_4(axial ? G1T : G2T, f, m1);
    mxm_4(f, G2, m2);
    FOR25 {
        ds[q] = dze[q] * m1[q] + dzx[q] * m2[q];
        dz[q] = dse[q] * m1[q] + dsx[q] * m2[q];
    }
}
/* f_over_s_solid / f_over_s_fluid of one element (:130-178): f / s, on the axis the s-derivative */
static void f_over_s_el(const float *G1T, const float *G2, int axial, const float *f, const float *inv_s,
                        const float *dze, const float *dzx, float *out) {
    FOR25 out[q] = inv_s[q] * f[q];
    if (axial) {
        float m1[NPT], m2[NPT];
        mxm_4(G1T, f, m1);
        mxm_4(f, G2, m2);
        for (int j = 0; j < NP; j++) out[NP * j] = dze[NP * j] * m1[NP * j] + dzx[NP * j] * m2[NP * j];
    }
}
/* where element-local point q of element e (domain offset eoff = 0 solid, nel_s fluid) goes in
 * the dump buffer: kwf mapping for strain_only (wavefields_io.f90:743-783), the packed block
 * ibeg:iend x jbeg:jend for fullfields (:811-812); -1 = not dumped */
static long dump_slot(const axo_t *o, int fluid, int e, int q) {
    if (o->dump_type == AXB_DUMP_STRAIN_ONLY) {
        size_t pk = q + (size_t)NPT * ((size_t)e + (fluid ? o->nel_s : 0));
        return o->kwf_mask[pk] ? (long)o->kwf_map[pk] - 1 : -1;
    }
    int i = q % NP, j = q / NP, ni = o->iend - o->ibeg + 1, nj = o->jend - o->jbeg + 1;
    if (i < o->ibeg || i > o->iend || j < o->jbeg || j > o->jend) return -1;
    long base = fluid ? (long)ni * nj * o->nel_s : 0;
    return base + ((long)e * nj + (j - o->jbeg)) * ni + (i - o->ibeg);
}
/* compute_strain (time_evol_wave.F90:1264-1410) [+ dump_velo_global, wavefields_io.f90:932-1015,
 * for fullfields].  Variable planes in the order of nc_routines.F90:976-1045. */
static void compute_strain_dump(axo_t *o) {
    const size_t cs = (size_t)NPT * o->nel_s;
    const size_t npts = snapshot_npoints(o), vs = npts * o->nstrain_max;
    float *base = o->snapdump + npts * o->istrain;
    const int mono = o->src_order == AXB_MONOPOLE, di = o->src_order == AXB_DIPOLE;
    const int full = o->dump_type == AXB_DUMP_FULLFIELDS;
    const int V_DSUS = 0, V_DSUZ = 1, V_DPUP = 2, V_DSUP = 3, V_DZUP = 4, V_TR = mono ? 3 : 5;
    const int V_VS = mono ? 4 : 6, V_VP = 7, V_VZ = mono ? 5 : 8;
    const float two = 2.0f;
    for (int e = 0; e < o->nel_s; e++) {
        const float *u1 = EL(o->disp, e), *u2 = EL(o->disp + cs, e), *u3 = EL(o->disp + 2 * cs, e);
        const float *dse = EL(o->dDse, e), *dze = EL(o->dDze, e), *dsx = EL(o->dDsx, e), *dzx = EL(o->dDzx, e);
        const float *is = EL(o->d_inv_s, e);
        const int ax = o->axis_s[e];
        float T[NPT], g1[NPT], g2[NPT], hs[NPT], hz[NPT], buff[NPT], fs[NPT], fs3[NPT];
        float E[6][NPT];
        if (di) { FOR25 T[q] = u1[q] + u2[q]; grad_el(o->G1T, o->G2T, o->G2, ax, T, dse, dze, dsx, dzx, g1, g2); }
        else grad_el(o->G1T, o->G2T, o->G2, ax, u1, dse, dze, dsx, dzx, g1, g2);
        FOR25 E[V_DSUS][q] = g1[q];
        grad_el(o->G1T, o->G2T, o->G2, ax, u3, dse, dze, dsx, dzx, hs, hz);
        /* axisym_gradient_solid_add: grad(1) = old(2) + dsdf, grad(2) = old(1) + dzdf */
        FOR25 { float o1 = g1[q], o2 = g2[q]; g1[q] = o2 + hs[q]; g2[q] = o1 + hz[q]; }
        FOR25 { g1[q] = g1[q] / two; E[V_DSUZ][q] = g1[q]; }
        if (mono) {
            f_over_s_el(o->G1T, o->G2, ax, u1, is, dze, dzx, buff);
            FOR25 { E[V_DPUP][q] = buff[q]; E[V_TR][q] = buff[q] + g2[q]; }
        } else if (di) {
            f_over_s_el(o->G1T, o->G2, ax, u2, is, dze, dzx, fs);
            FOR25 { buff[q] = two * fs[q]; E[V_DPUP][q] = buff[q]; E[V_TR][q] = buff[q] + g2[q]; }
            FOR25 T[q] = u1[q] - u2[q];
            grad_el(o->G1T, o->G2T, o->G2, ax, T, dse, dze, dsx, dzx, hs, hz);
            f_over_s_el(o->G1T, o->G2, ax, u3, is, dze, dzx, fs3);
            FOR25 {
                E[V_DSUP][q] = -fs[q] - hs[q] / two;
                E[V_DZUP][q] = -(fs3[q] + hz[q]) / two;
            }
        } else {
            FOR25 T[q] = u1[q] - two * u2[q];
            f_over_s_el(o->G1T, o->G2, ax, T, is, dze, dzx, buff);
            FOR25 { E[V_DPUP][q] = buff[q]; E[V_TR][q] = buff[q] + g2[q]; }
            grad_el(o->G1T, o->G2T, o->G2, ax, u2, dse, dze, dsx, dzx, hs, hz);
            FOR25 T[q] = u1[q] + u2[q] / two;
            f_over_s_el(o->G1T, o->G2, ax, T, is, dze, dzx, fs);
            f_over_s_el(o->G1T, o->G2, ax, u3, is, dze, dzx, fs3);
            FOR25 {
                E[V_DSUP][q] = -fs[q] - hs[q] / two;
                E[V_DZUP][q] = -fs3[q] - hz[q] / two;
            }
        }
        const float *v1 = EL(o->velo, e), *v2 = EL(o->velo + cs, e), *v3 = EL(o->velo + 2 * cs, e);
        FOR25 {
            long ct = dump_slot(o, 0, e, q);
            if (ct < 0) continue;
            for (int v = 0; v < (mono ? 4 : 6); v++) base[ct + vs * v] = E[v][q];
            if (full) {
                if (di) { base[ct + vs * V_VS] = v1[q] + v2[q]; base[ct + vs * V_VP] = v1[q] - v2[q]; }
                else { base[ct + vs * V_VS] = v1[q]; if (!mono) base[ct + vs * V_VP] = v2[q]; }
                base[ct + vs * V_VZ] = v3[q];
            }
        }
    }
    for (int e = 0; e < o->nel_f; e++) {
        const float *dse = EL(o->Dse_f, e), *dze = EL(o->Dze_f, e), *dsx = EL(o->Dsx_f, e), *dzx = EL(o->Dzx_f, e);
        const float *is = EL(o->d_inv_s_f, e), *ir = EL(o->inv_rho_fluid, e);
        const int ax = o->axis_f[e];
        float us[NPT], uz[NPT], g1[NPT], g2[NPT], hs[NPT], hz[NPT], fs[NPT], fz[NPT], ws[NPT], wz[NPT];
        float E[6][NPT];
        grad_el(o->G1T, o->G2T, o->G2, ax, EL(o->chi, e), dse, dze, dsx, dzx, us, uz);
        FOR25 { us[q] = us[q] * ir[q]; uz[q] = uz[q] * ir[q]; }
        grad_el(o->G1T, o->G2T, o->G2, ax, us, dse, dze, dsx, dzx, g1, g2);
        FOR25 E[V_DSUS][q] = g1[q];
        grad_el(o->G1T, o->G2T, o->G2, ax, uz, dse, dze, dsx, dzx, hs, hz);
        FOR25 { float o1 = g1[q], o2 = g2[q]; g1[q] = o2 + hs[q]; g2[q] = o1 + hz[q]; }
        FOR25 { g1[q] = g1[q] / two; E[V_DSUZ][q] = g1[q]; }
        f_over_s_el(o->G1T, o->G2, ax, us, is, dze, dzx, fs);
        FOR25 { E[V_DPUP][q] = fs[q]; E[V_TR][q] = fs[q] + g2[q]; }
        if (!mono) {
            f_over_s_el(o->G1T, o->G2, ax, uz, is, dze, dzx, fz);
            if (di) FOR25 { E[V_DSUP][q] = (-fs[q]) / two; E[V_DZUP][q] = fz[q] / two; }
            else FOR25 { E[V_DSUP][q] = -fs[q]; E[V_DZUP][q] = -fz[q]; }
        }
        if (full) grad_el(o->G1T, o->G2T, o->G2, ax, EL(o->dchi, e), dse, dze, dsx, dzx, ws, wz);
        FOR25 {
            long ct = dump_slot(o, 1, e, q);
            if (ct < 0) continue;
            for (int v = 0; v < (mono ? 4 : 6); v++) base[ct + vs * v] = E[v][q];
            if (full) {
                /* wavefields_io.f90:992-993 reads the s component at first index jbeg:jend where
                 * it means ibeg:iend: the value stored for point i is that of i - ibeg + jbeg */
                int i = q % NP, j = q / NP, iq = i - o->ibeg + o->jbeg;
                base[ct + vs * V_VS] = ir[q] * ws[iq + NP * j];
                if (!mono) base[ct + vs * V_VP] = 0.0f;
                base[ct + vs * V_VZ] = ir[q] * wz[q];
            }
        }
    }
    o->istrain++;
}

/* xdmf_mapping (wavefields_io.f90:690-738) of one element: the corners of the plot cells, in
 * the reference's loop order, where plotting_mask says this element owns the plot point */
static void xdmf_map_el(const axo_t *o, int iel, const float *v, float *plane) {
    const int in = o->x_in, jn = o->x_jn;
    const int *mask = o->x_mask + (size_t)in * jn * iel, *map = o->x_map + (size_t)in * jn * iel;
    for (int i = 0; i < in - 1; i++) {
        const int ipol = o->x_iarr[i], ipol1 = o->x_iarr[i + 1];
        for (int j = 0; j < jn - 1; j++) {
            const int jpol = o->x_jarr[j], jpol1 = o->x_jarr[j + 1];
            if (mask[i + in * j]) plane[map[i + in * j] - 1] = v[ipol + NP * jpol];
            if (mask[i + 1 + in * j]) plane[map[i + 1 + in * j] - 1] = v[ipol1 + NP * jpol];
            if (mask[i + 1 + in * (j + 1)]) plane[map[i + 1 + in * (j + 1)] - 1] = v[ipol1 + NP * jpol1];
            if (mask[i + in * (j + 1)]) plane[map[i + in * (j + 1)] - 1] = v[ipol + NP * jpol1];
        }
    }
}
/* glob_snapshot_xdmf (wavefields_io.f90:119-203) with calc_straintrace (:630-686) and
 * calc_curlinplane (:601-627): planes u_s, u_p, u_z, straintrace, curlinplane */
static void glob_snapshot_xdmf(axo_t *o) {
    const size_t cs = (size_t)NPT * o->nel_s;
    const size_t npts = (size_t)o->npoint_plot;
    const int mono = o->src_order == AXB_MONOPOLE, di = o->src_order == AXB_DIPOLE;
    const float two_rk = 2.0f;
    if (!o->xsnap) {
        o->nsnap_max = o->niter / o->snap_it + 1;        /* parameters.F90:946 */
        o->xsnap = zerosf(npts * o->nsnap_max * 5);
    }
    float *P[5];
    for (int v = 0; v < 5; v++) P[v] = o->xsnap + npts * (o->isnap + (size_t)o->nsnap_max * v);
    for (int e = 0; e < o->nel_f; e++) {
        const float *dse = EL(o->xf_Dse, e), *dze = EL(o->xf_Dze, e), *dsx = EL(o->xf_Dsx, e), *dzx = EL(o->xf_Dzx, e);
        const float *is = EL(o->xf_inv_s, e), *ir = EL(o->xf_inv_rho, e);
        const int ax = o->axis_f[e];
        float us[NPT], uz[NPT], g1[NPT], g2[NPT], hs[NPT], hz[NPT], fs[NPT], zero[NPT], tr[NPT];
        grad_el(o->G1T, o->G2T, o->G2, ax, EL(o->chi, e), dse, dze, dsx, dzx, us, uz);
        FOR25 { us[q] = us[q] * ir[q]; uz[q] = uz[q] * ir[q]; zero[q] = 0.0f; }
        grad_el(o->G1T, o->G2T, o->G2, ax, us, dse, dze, dsx, dzx, g1, g2);     /* 1: dsus, 2: dzus */
        grad_el(o->G1T, o->G2T, o->G2, ax, uz, dse, dze, dsx, dzx, hs, hz);
        /* axisym_gradient_fluid_add: grad(2) = dsus + dzuz */
        FOR25 g2[q] = g1[q] + hz[q];
        f_over_s_el(o->G1T, o->G2, ax, us, is, dze, dzx, fs);
        FOR25 tr[q] = fs[q] + g2[q];
        xdmf_map_el(o, e, us, P[0]); xdmf_map_el(o, e, zero, P[1]); xdmf_map_el(o, e, uz, P[2]);
        xdmf_map_el(o, e, tr, P[3]); xdmf_map_el(o, e, zero, P[4]);
    }
    for (int e = 0; e < o->nel_s; e++) {
        const float *u1 = EL(o->disp, e), *u2 = EL(o->disp + cs, e), *u3 = EL(o->disp + 2 * cs, e);
        const float *dse = EL(o->xs_Dse, e), *dze = EL(o->xs_Dze, e), *dsx = EL(o->xs_Dsx, e), *dzx = EL(o->xs_Dzx, e);
        const float *is = EL(o->xs_inv_s, e);
        const int ax = o->axis_s[e];
        float S[NPT], Pp[NPT], T[NPT], g1[NPT], g2[NPT], hs[NPT], hz[NPT], buff[NPT], tr[NPT], curl[NPT];
        /* f_sol_spz: the (+, -) pair of a dipole source becomes (s, phi) */
        if (di) FOR25 { S[q] = u1[q] + u2[q]; Pp[q] = u1[q] - u2[q]; }
        else FOR25 { S[q] = u1[q]; Pp[q] = u2[q]; }
        grad_el(o->G1T, o->G2T, o->G2, ax, S, dse, dze, dsx, dzx, g1, g2);      /* 1: dsus, 2: dzus */
        grad_el(o->G1T, o->G2T, o->G2, ax, u3, dse, dze, dsx, dzx, hs, hz);     /* 1: dsuz, 2: dzuz */
        FOR25 curl[q] = g2[q] - hs[q];
        if (mono) f_over_s_el(o->G1T, o->G2, ax, u1, is, dze, dzx, buff);
        else if (di) { f_over_s_el(o->G1T, o->G2, ax, u2, is, dze, dzx, buff); FOR25 buff[q] = two_rk * buff[q]; }
        else { FOR25 T[q] = u1[q] - two_rk * u2[q]; f_over_s_el(o->G1T, o->G2, ax, T, is, dze, dzx, buff); }
        FOR25 tr[q] = buff[q] + (g1[q] + hz[q]);
        const int iel = o->nel_f + e;
        xdmf_map_el(o, iel, S, P[0]); xdmf_map_el(o, iel, Pp, P[1]); xdmf_map_el(o, iel, u3, P[2]);
        xdmf_map_el(o, iel, tr, P[3]); xdmf_map_el(o, iel, curl, P[4]);
    }
    o->isnap++;
}

/* time_evol_wave.F90:1104-1251, the parts on the hot path */
static void solid_stiffness(axo_t *o, float *acc, const float *u);
/* time_evol_wave.F90:1424-1526.  sum() is taken in array order with a real(4) accumulator;
 * psum and the factor two*pi are left to the caller (see the header). */
static void energy(axo_t *o, int iter) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    float *out = o->energy + (size_t)4 * iter;
    float ekin_sol = 0.0f, epot_sol = 0.0f, ekin_flu = 0.0f, epot_flu = 0.0f;
    float *disp = (float *)malloc(sizeof(float) * (3 * ns + 1));
    float *stiff = zerosf(3 * ns);
    memcpy(disp, o->disp, sizeof(float) * 3 * ns);
    mask_solid(o, disp);
    solid_stiffness(o, stiff, disp);
    mask_solid(o, stiff);
    for (size_t k = 0; k < 3 * ns; k++) { stiff[k] = stiff[k] * disp[k]; epot_sol = epot_sol + stiff[k]; }
    for (int c = 0; c < 3; c++)
        for (size_t p = 0; p < ns; p++) {
            float v = o->velo[p + ns * c];
            float x = v * v * o->um_rho_s[p];
            if (c == 2 && o->src_order == AXB_DIPOLE) x = 2.0f * x;
            stiff[p + ns * c] = x;
        }
    for (size_t k = 0; k < 3 * ns; k++) ekin_sol = ekin_sol + stiff[k];
    free(disp); free(stiff);
    if (o->nel_f > 0) {
        /* the reference passes ddchi0 (Newmark) / ddchi (symplectic: our ddchi1) */
        const float *ddchi = o->scheme == AXB_NEWMARK2 ? o->ddchi0 : o->ddchi1;
        float *dchi = (float *)malloc(sizeof(float) * (nf + 1));
        float *sf = zerosf(nf);
        for (size_t p = 0; p < nf; p++) { float x = ddchi[p] * ddchi[p] * o->um_lam_f[p]; epot_flu = epot_flu + x; }
        memcpy(dchi, o->dchi, sizeof(float) * nf);
        if (o->src_order != AXB_MONOPOLE) axis_mask(dchi, 0, o->ax_el_f, o->naxel_f, 0, 0);
        glob_fluid_stiffness_4(o, sf, dchi);
        if (o->src_order != AXB_MONOPOLE) axis_mask(sf, 0, o->ax_el_f, o->naxel_f, 0, 0);
        for (size_t p = 0; p < nf; p++) { sf[p] = sf[p] * dchi[p]; ekin_flu = ekin_flu + sf[p]; }
        free(dchi); free(sf);
    }
    out[0] = epot_sol; out[1] = ekin_sol; out[2] = epot_flu; out[3] = ekin_flu;
}

static void dump_stuff(axo_t *o, int iter) {
    if (o->dump_energy) energy(o, iter);
    if (o->num_rec > 0 && iter % o->seis_it == 0 && o->iseismo < o->nseismo_max) sample_receivers(o);
    if (o->strain_it > 0 && o->have_kwf && iter % o->strain_it == 0 && o->istrain < o->nstrain_max) {
        if (o->dump_type == AXB_DUMP_DISPL_ONLY) dump_disp_global(o);
        else compute_strain_dump(o);
    }
    /* time_evol_wave.F90:1167-1176 */
    if (o->have_xdmf && iter % o->snap_it == 0 && (!o->xsnap || o->isnap < o->nsnap_max)) glob_snapshot_xdmf(o);
}

/* ------------------------------------------------------------------------------------ */
static void solid_stiffness(axo_t *o, float *acc, const float *u) {
    if (o->src_order == AXB_MONOPOLE) glob_stiffness_mono_4(o, acc, u);
    else if (o->src_order == AXB_DIPOLE) glob_stiffness_di_4(o, acc, u);
    else glob_stiffness_quad_4(o, acc, u);
}
static void anel_stiffness(axo_t *o, float *acc) {
    if (o->cg) glob_anel_stiffness_cg4(o, acc, o->memvar);
    else glob_anel_stiffness_4(o, acc, o->memvar);
}

/* the loop body is split at the two MPI phases so that a group of in-process ranks can
 * be advanced in lockstep; the order of operations inside is the reference's. */
typedef struct { int stage; double cd, cv; float stf; } stage_t;

static void newmark_part1(axo_t *o) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    const double dt = o->deltat, hsq = o->half_dt_sq;
    o->t += dt;
    for (size_t p = 0; p < nf; p++)
        o->chi[p] = (float)((double)o->chi[p] + dt * (double)o->dchi[p] + hsq * (double)o->ddchi0[p]);
    for (int c = 0; c < 3; c++) {
        if (c == 1 && o->src_order == AXB_MONOPOLE) continue;
        float *d = o->disp + ns * c; const float *v = o->velo + ns * c, *a = o->acc0 + ns * c;
        for (size_t p = 0; p < ns; p++)
            d[p] = (float)((double)d[p] + dt * (double)v[p] + hsq * (double)a[p]);
    }
    if (o->src_order != AXB_MONOPOLE) axis_mask(o->chi, 0, o->ax_el_f, o->naxel_f, 0, 0);
    glob_fluid_stiffness_4(o, o->ddchi1, o->chi);
    if (o->fluid_src) add_source_fl(o, o->ddchi1, o->stf[o->iter]);
    bdry_copy2fluid(o, o->ddchi1, o->disp);
    if (o->src_order != AXB_MONOPOLE) axis_mask(o->ddchi1, 0, o->ax_el_f, o->naxel_f, 0, 0);
    if (o->fs_mask) for (size_t p = 0; p < nf; p++) o->ddchi1[p] = o->ddchi1[p] * o->fs_mask[p];
    feed_buffer(o, AXB_DOMAIN_FLUID, o->ddchi1);            /* pdistsum_fluid phase 1 */
}
static void fluid_local_and_solid_stiffness(axo_t *o, float *ddchi, float *acc) {
    gather_scatter(o, AXB_DOMAIN_FLUID, ddchi);
    mask_solid(o, o->disp);
    solid_stiffness(o, acc, o->disp);
    if (o->anel) anel_stiffness(o, acc);
}
static void newmark_part2(axo_t *o) {
    const size_t nf = (size_t)NPT * o->nel_f;
    fluid_local_and_solid_stiffness(o, o->ddchi1, o->acc1);
    extract_from_buffer(o, AXB_DOMAIN_FLUID, o->ddchi1);    /* pdistsum_fluid phase 2 */
    for (size_t p = 0; p < nf; p++) o->ddchi1[p] = -o->inv_mass_fluid[p] * o->ddchi1[p];
    if (o->have_abc)
        for (size_t p = 0; p < nf; p++)
            o->ddchi1[p] = o->ddchi1[p] - 2 * o->gamma_f[p] * o->dchi[p]
                           - (o->gamma_f[p] * o->gamma_f[p]) * o->chi[p];
    bdry_copy2solid(o, o->acc1, o->ddchi1);
    mask_solid(o, o->acc1);
    feed_buffer(o, AXB_DOMAIN_SOLID, o->acc1);              /* pdistsum_solid phase 1 */
}
static void newmark_part3(axo_t *o) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    const double hdt = o->half_dt;
    gather_scatter(o, AXB_DOMAIN_SOLID, o->acc1);
    if (o->anel) time_step_memvars(o);
    for (size_t p = 0; p < nf; p++) {
        o->dchi[p] = (float)((double)o->dchi[p] + hdt * (double)(o->ddchi0[p] + o->ddchi1[p]));
        o->ddchi0[p] = o->ddchi1[p];
    }
    extract_from_buffer(o, AXB_DOMAIN_SOLID, o->acc1);      /* pdistsum_solid phase 2 */
    if (!o->fluid_src) add_source_el(o, o->acc1, o->stf[o->iter]);
    for (int c = 0; c < 3; c++) {
        if (c == 1 && o->src_order == AXB_MONOPOLE) continue;
        float *a1 = o->acc1 + ns * c, *a0 = o->acc0 + ns * c, *v = o->velo + ns * c;
        const float *d = o->disp + ns * c;
        for (size_t e = 0, p = 0; p < ns; p++) {
            (void)e;
            float im = o->inv_mass_rho[p];
            if (c == 2 && o->src_order == AXB_DIPOLE) a1[p] = (float)(-2.0 * (double)im * (double)a1[p]);
            else a1[p] = -im * a1[p];
            if (o->have_abc)
                a1[p] = a1[p] - 2 * o->gamma_s[p] * v[p] - (o->gamma_s[p] * o->gamma_s[p]) * d[p];
            v[p] = (float)((double)v[p] + hdt * (double)(a0[p] + a1[p]));
            a0[p] = a1[p];
        }
    }
    o->iter++;
    dump_stuff(o, o->iter);
}

/* source.f90:662-692: the reference's erf (Numerical Recipes erfc; default-real literals) */
static double erf_nr(double x) {
    static const float c[10] = {-1.26551223f, 1.00002368f, 0.37409196f, 0.09678418f, -0.18628806f,
                                0.27886807f, -1.13520398f, 1.48851587f, -0.82215223f, 0.17087277f};
    double z = fabs(x), t = 1.0 / (1.0 + 0.5 * z), poly = (double)c[9], erfcc;
    for (int k = 8; k >= 0; k--) poly = t * poly + (double)c[k];
    erfcc = t * exp(-z * z + poly);
    if (x < 0.0) erfcc = 2.0 - erfcc;
    return 1.0 - erfcc;
}

/* compute_stf_t, source.f90:206-233: gauss_t :818-831, gauss_d_t :835-849, gauss_dd_t :853-869,
 * errorf_t :873-886, delta_src_t :890-904 (whole-array assignments switched by t(1)),
 * quasiheavi_t :908-917 (stf_t(seis_it:nstf_t) = magnitude) */
static void compute_stf_t(const axo_t *o, int nstf_t, const double *t, double *stf_t) {
    const double a = o->decay / o->t_0, pi = 3.1415926535898;
    int i;
    switch (o->stf_type) {
    case AXB_STF_GAUSS_0:
        for (i = 0; i < nstf_t; i++) {
            double x = a * (t[i] - o->shift_fact);
            stf_t[i] = exp(-(x * x)) * o->magnitude * a / sqrt(pi);
        }
        break;
    case AXB_STF_GAUSS_1:
        for (i = 0; i < nstf_t; i++) {
            double x = a * (t[i] - o->shift_fact);
            stf_t[i] = -2.0 * a * a * (t[i] - o->shift_fact) * exp(-(x * x)) / (a * sqrt(2.0) * exp(-0.5)) * o->magnitude;
        }
        break;
    case AXB_STF_GAUSS_2:
        for (i = 0; i < nstf_t; i++) {
            double x = a * (t[i] - o->shift_fact);
            stf_t[i] = a * a * (2.0 * a * a * (t[i] - o->shift_fact) * (t[i] - o->shift_fact) - 1.0)
                       * exp(-(x * x)) / (2.0 * a * a * exp(-1.5)) * o->magnitude;
        }
        break;
    case AXB_STF_ERRORF:
        for (i = 0; i < nstf_t; i++) stf_t[i] = (erf_nr(a * (t[i] - o->shift_fact)) * 0.5 + 0.5) * o->magnitude;
        break;
    case AXB_STF_DIRAC_0:
        for (i = 0; i < nstf_t; i++) stf_t[i] = 0.0;
        if (t[0] > (o->shift_fact - o->deltat) && t[0] <= o->shift_fact)
            for (i = 0; i < nstf_t; i++) stf_t[i] = (t[i] - t[0]) / o->deltat * o->magnitude / o->deltat;
        if (t[0] >= o->shift_fact && t[0] < (o->shift_fact + o->deltat))
            for (i = 0; i < nstf_t; i++) stf_t[i] = (1. - (t[i] - t[0]) / o->deltat) * o->magnitude / o->deltat;
        break;
    case AXB_STF_QUHEAVI:
        for (i = 0; i < nstf_t; i++) stf_t[i] = 0.0;
        for (i = (o->seis_it > 1 ? o->seis_it : 1); i <= nstf_t; i++) stf_t[i - 1] = o->magnitude;
        break;
    default:
        fprintf(stderr, " source time function non existant: %d\n", o->stf_type);
        abort();
    }
}

/* time_evol_wave.F90:584-739 */
static void symp_drift(axo_t *o, double cd) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    for (size_t p = 0; p < nf; p++) o->chi[p] = (float)((double)o->chi[p] + (double)o->dchi[p] * cd);
    for (int c = 0; c < 3; c++) {
        if (c == 1 && o->src_order == AXB_MONOPOLE) continue;
        float *d = o->disp + ns * c; const float *v = o->velo + ns * c;
        for (size_t p = 0; p < ns; p++) d[p] = (float)((double)d[p] + (double)v[p] * cd);
    }
}
static void symp_part1(axo_t *o, int i) {
    symp_drift(o, o->coefd[i]);
    if (o->src_order != AXB_MONOPOLE) axis_mask(o->chi, 0, o->ax_el_f, o->naxel_f, 0, 0);
    glob_fluid_stiffness_4(o, o->ddchi1, o->chi);
    bdry_copy2fluid(o, o->ddchi1, o->disp);
    if (o->src_order != AXB_MONOPOLE) axis_mask(o->ddchi1, 0, o->ax_el_f, o->naxel_f, 0, 0);
    feed_buffer(o, AXB_DOMAIN_FLUID, o->ddchi1);
}
static void symp_part2(axo_t *o) {
    const size_t nf = (size_t)NPT * o->nel_f;
    fluid_local_and_solid_stiffness(o, o->ddchi1, o->acc1);
    extract_from_buffer(o, AXB_DOMAIN_FLUID, o->ddchi1);
    for (size_t p = 0; p < nf; p++) o->ddchi1[p] = -o->ddchi1[p] * o->inv_mass_fluid[p];
    if (o->have_abc)
        for (size_t p = 0; p < nf; p++)
            o->ddchi1[p] = o->ddchi1[p] - 2 * o->gamma_f[p] * o->dchi[p]
                           - (o->gamma_f[p] * o->gamma_f[p]) * o->chi[p];
    bdry_copy2solid(o, o->acc1, o->ddchi1);
    mask_solid(o, o->acc1);
    feed_buffer(o, AXB_DOMAIN_SOLID, o->acc1);
}
static void symp_part3(axo_t *o, int i, double stf_i) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    const double cv = o->coefv[i];
    gather_scatter(o, AXB_DOMAIN_SOLID, o->acc1);
    for (size_t p = 0; p < nf; p++) o->dchi[p] = (float)((double)o->dchi[p] + cv * (double)o->ddchi1[p]);
    extract_from_buffer(o, AXB_DOMAIN_SOLID, o->acc1);
    add_source_el(o, o->acc1, (float)stf_i);
    for (int c = 0; c < 3; c++) {
        if (c == 1 && o->src_order == AXB_MONOPOLE) continue;
        float *a = o->acc1 + ns * c, *v = o->velo + ns * c; const float *d = o->disp + ns * c;
        for (size_t p = 0; p < ns; p++) {
            a[p] = -o->inv_mass_rho[p] * a[p];
            if (o->have_abc)
                a[p] = a[p] - 2 * o->gamma_s[p] * v[p] - (o->gamma_s[p] * o->gamma_s[p]) * d[p];
            if (c == 2 && o->src_order == AXB_DIPOLE) v[p] = (float)((double)v[p] + 2.0 * (double)a[p] * cv);
            else v[p] = (float)((double)v[p] + (double)a[p] * cv);
        }
    }
}
static void symp_finish(axo_t *o) {
    symp_drift(o, o->coefd[o->nstages]);
    if (o->anel) time_step_memvars(o);
    o->iter++;
    dump_stuff(o, o->iter);
}

#define PAR_RANKS

static void set_ftz(void) {
#if defined(__x86_64__)
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);        /* SOLVER/ftz.c:44-48 */
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
#endif
}

int axo_connect_local(axb_handle *handles, int32_t n) {
    axo_t **g = (axo_t **)malloc(sizeof(axo_t *) * n);
    for (int i = 0; i < n; i++) g[i] = handles[i];
    for (int i = 0; i < n; i++) { handles[i]->group = g; handles[i]->ngroup = n; }
    return 0;
}
/* ---- one process per rank: the "MPI" of the oracle is a POSIX shared-memory segment per
 * rank holding its receive slabs [domain][parity][message] and one arrival counter per
 * message; peers write straight into it (commpi.F90:408-449 ISEND/IRECV) and the owner
 * spins on the counters (MPI_WAITALL, :453, :587).  Blob = segment name + slab layout. */
typedef struct {
    int rank;
    char name[64];
    int nmsg[2], peer[2][MAXMSG], size[2][MAXMSG];
    long off[2][2][MAXMSG];       /* float offset of [dom][parity][msg] */
    long bytes;
} ipc_blob_t;

static void ipc_layout(const axo_t *o, ipc_blob_t *b) {
    memset(b, 0, sizeof *b);
    b->rank = o->rank;
    long off = 1024;              /* first 4 KiB: counters, int[2][MAXMSG] */
    for (int d = 0; d < 2; d++) {
        const halo_t *H = &o->halo[d];
        int nc = d == AXB_DOMAIN_SOLID ? 3 : 1;
        b->nmsg[d] = H->nmsg;
        for (int m = 0; m < H->nmsg; m++) { b->peer[d][m] = H->peer[m]; b->size[d][m] = H->size[m]; }
        for (int par = 0; par < 2; par++)
            for (int m = 0; m < H->nmsg; m++) { b->off[d][par][m] = off; off += (long)H->size[m] * nc; }
    }
    b->bytes = off * (long)sizeof(float);
}

int32_t axo_ipc_blob_bytes(void) { return AXB_IPC_BLOB_BYTES; }

int axo_ipc_export(axb_handle h, void *blob, int32_t n) {
    static int counter = 0;
    ipc_blob_t b;
    if ((size_t)n < sizeof b) return fail("ipc blob too small");
    if (!h->finalized) return fail("ipc_export before finalize_setup");
    ipc_layout(h, &b);
    if (!h->shm_base) {
        snprintf(h->shm_name, sizeof h->shm_name, "/axo_%d_%d_%d", (int)getpid(), h->rank, counter++);
        int fd = shm_open(h->shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return fail("shm_open failed");
        if (ftruncate(fd, b.bytes) != 0) { close(fd); return fail("ftruncate failed"); }
        h->shm_base = mmap(NULL, b.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (h->shm_base == MAP_FAILED) { h->shm_base = NULL; return fail("mmap failed"); }
        h->shm_bytes = b.bytes;
        memset(h->shm_base, 0, b.bytes);
        for (int d = 0; d < 2; d++)
            for (int m = 0; m < h->halo[d].nmsg; m++) {
                h->halo[d].my_flag[m] = (volatile int *)h->shm_base + d * MAXMSG + m;
                for (int par = 0; par < 2; par++)
                    h->halo[d].shm_recv[par][m] = (float *)h->shm_base + b.off[d][par][m];
            }
    }
    snprintf(b.name, sizeof b.name, "%s", h->shm_name);
    memcpy(blob, &b, sizeof b);
    h->ipc_mode = 1;
    return 0;
}

int axo_ipc_import(axb_handle h, int32_t peer_rank, const void *blob, int32_t n) {
    ipc_blob_t b;
    if ((size_t)n < sizeof b) return fail("ipc blob too small");
    memcpy(&b, blob, sizeof b);
    if (b.rank != peer_rank) return fail("ipc blob does not belong to that rank");
    int fd = shm_open(b.name, O_RDWR, 0600);
    if (fd < 0) return fail("shm_open(peer) failed");
    void *base = mmap(NULL, b.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return fail("mmap(peer) failed");
    for (int d = 0; d < 2; d++) {
        halo_t *H = &h->halo[d];
        for (int m = 0; m < H->nmsg; m++) {
            if (H->peer[m] != peer_rank) continue;
            int mm = -1;
            for (int q = 0; q < b.nmsg[d]; q++) if (b.peer[d][q] == h->rank) mm = q;
            if (mm < 0 || b.size[d][mm] != H->size[m]) return fail("halo lists inconsistent between ranks");
            for (int par = 0; par < 2; par++) H->peer_recv[par][m] = (float *)base + b.off[d][par][mm];
            H->peer_flag[m] = (volatile int *)base + d * MAXMSG + mm;
        }
    }
    h->ipc_mode = 1;
    return 0;
}

/* send_recv_buffers + MPI_WAITALL between processes */
static int exchange_ipc(axo_t *o, int dom) {
    halo_t *H = &o->halo[dom];
    int nc = dom == AXB_DOMAIN_SOLID ? 3 : 1;
    int par = H->seq & 1;
    for (int m = 0; m < H->nmsg; m++) {
        if (!H->peer_recv[par][m]) return fail("halo peers not connected (axo_ipc_import)");
        memcpy(H->peer_recv[par][m], H->sendbuf[m], sizeof(float) * (size_t)H->size[m] * nc);
        __atomic_store_n((int *)H->peer_flag[m], H->seq + 1, __ATOMIC_RELEASE);
    }
    for (int m = 0; m < H->nmsg; m++) {
        long spins = 0;
        while (__atomic_load_n((int *)H->my_flag[m], __ATOMIC_ACQUIRE) < H->seq + 1) {
            if (++spins > 2000000000L) return fail("halo exchange timed out");
            if ((spins & 1023) == 0) sched_yield();
        }
        H->recvbuf[m] = H->shm_recv[par][m];
    }
    H->seq++;
    return 0;
}

static int run_ipc(axo_t *o, int nsteps) {
    for (int s = 0; s < nsteps; s++) {
        if (o->scheme == AXB_NEWMARK2) {
            newmark_part1(o);
            if (exchange_ipc(o, AXB_DOMAIN_FLUID)) return 1;
            newmark_part2(o);
            if (exchange_ipc(o, AXB_DOMAIN_SOLID)) return 1;
            newmark_part3(o);
        } else {
            double stf_symp[40];
            o->t += o->deltat;
            double subdt[40];
            for (int k = 0; k < o->nstages; k++) subdt[k] = o->t - o->deltat + o->coeff[k];
            compute_stf_t(o, o->nstages, subdt, stf_symp);
            for (int k = 0; k < o->nstages; k++) {
                symp_part1(o, k);
                if (exchange_ipc(o, AXB_DOMAIN_FLUID)) return 1;
                symp_part2(o);
                if (exchange_ipc(o, AXB_DOMAIN_SOLID)) return 1;
                symp_part3(o, k, stf_symp[k]);
            }
            symp_finish(o);
        }
    }
    return 0;
}

/* One thread per rank (= one MPI rank per core in the reference); barriers stand where
 * the reference has its MPI_WAITALLs. */
typedef struct { axo_t **hs; int n, i, nsteps; pthread_barrier_t *bar; int err; } worker_t;

static void *rank_worker(void *arg) {
    worker_t *w = (worker_t *)arg;
    axo_t *o = w->hs[w->i];
    set_ftz();
    for (int s = 0; s < w->nsteps; s++) {
        if (o->scheme == AXB_NEWMARK2) {
            newmark_part1(o);
            pthread_barrier_wait(w->bar);
            if (exchange(o, AXB_DOMAIN_FLUID)) w->err = 1;
            pthread_barrier_wait(w->bar);
            newmark_part2(o);
            pthread_barrier_wait(w->bar);
            if (exchange(o, AXB_DOMAIN_SOLID)) w->err = 1;
            pthread_barrier_wait(w->bar);
            newmark_part3(o);
        } else {
            double stf_symp[40];
            o->t += o->deltat;
            double subdt[40];
            for (int k = 0; k < o->nstages; k++) subdt[k] = o->t - o->deltat + o->coeff[k];
            compute_stf_t(o, o->nstages, subdt, stf_symp);
            for (int k = 0; k < o->nstages; k++) {
                symp_part1(o, k);
                pthread_barrier_wait(w->bar);
                if (exchange(o, AXB_DOMAIN_FLUID)) w->err = 1;
                pthread_barrier_wait(w->bar);
                symp_part2(o);
                pthread_barrier_wait(w->bar);
                if (exchange(o, AXB_DOMAIN_SOLID)) w->err = 1;
                pthread_barrier_wait(w->bar);
                symp_part3(o, k, stf_symp[k]);
            }
            symp_finish(o);
        }
    }
    return NULL;
}

static int run_group_threaded(axb_handle *hs, int n, int nsteps) {
    pthread_barrier_t bar;
    pthread_t th[64];
    worker_t w[64];
    if (n > 64) return fail("too many ranks");
    pthread_barrier_init(&bar, NULL, n);
    for (int i = 0; i < n; i++) {
        w[i].hs = hs; w[i].n = n; w[i].i = i; w[i].nsteps = nsteps; w[i].bar = &bar; w[i].err = 0;
        pthread_create(&th[i], NULL, rank_worker, &w[i]);
    }
    int err = 0;
    for (int i = 0; i < n; i++) { pthread_join(th[i], NULL); err |= w[i].err; }
    pthread_barrier_destroy(&bar);
    return err;
}

int axo_run_group(axb_handle *hs, int32_t n, int32_t nsteps) {
    set_ftz();
    for (int i = 0; i < n; i++) {
        if (!hs[i]->finalized) return fail("finalize_setup not called");
        if (n > 1 && hs[i]->ngroup != n) return fail("group not connected");
        if (hs[i]->iter + nsteps > hs[i]->niter) return fail("run beyond niter");
    }
    for (int i = 0; i < n; i++)
        if (hs[i]->iter == 0 && hs[i]->iseismo == 0 && hs[i]->istrain == 0)
            dump_stuff(hs[i], 0);                          /* time_evol_wave.F90:350 */
    if (n > 1) return run_group_threaded(hs, n, nsteps);
    for (int s = 0; s < nsteps; s++) {
        if (hs[0]->scheme == AXB_NEWMARK2) {
            /* one thread per rank (= one MPI rank per core in the reference); the end of
             * each parallel loop is the barrier at which the "messages" are exchanged */
            PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); newmark_part1(hs[i]); }
            for (int i = 0; i < n; i++) if (exchange(hs[i], AXB_DOMAIN_FLUID)) return 1;
            PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); newmark_part2(hs[i]); }
            for (int i = 0; i < n; i++) if (exchange(hs[i], AXB_DOMAIN_SOLID)) return 1;
            PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); newmark_part3(hs[i]); }
        } else {
            double stf_symp[40];
            for (int i = 0; i < n; i++) hs[i]->t += hs[i]->deltat;
            double subdt[40];
            for (int k = 0; k < hs[0]->nstages; k++) subdt[k] = hs[0]->t - hs[0]->deltat + hs[0]->coeff[k];
            compute_stf_t(hs[0], hs[0]->nstages, subdt, stf_symp);
            for (int k = 0; k < hs[0]->nstages; k++) {
                PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); symp_part1(hs[i], k); }
                for (int i = 0; i < n; i++) if (exchange(hs[i], AXB_DOMAIN_FLUID)) return 1;
                PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); symp_part2(hs[i]); }
                for (int i = 0; i < n; i++) if (exchange(hs[i], AXB_DOMAIN_SOLID)) return 1;
                PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); symp_part3(hs[i], k, stf_symp[k]); }
            }
            PAR_RANKS for (int i = 0; i < n; i++) { set_ftz(); symp_finish(hs[i]); }
        }
    }
    return 0;
}

int axo_run(axb_handle h, int32_t nsteps) {
    if (h->ipc_mode) {
        set_ftz();
        if (!h->finalized) return fail("finalize_setup not called");
        if (h->iter + nsteps > h->niter) return fail("run beyond niter");
        if (h->iter == 0 && h->iseismo == 0 && h->istrain == 0) dump_stuff(h, 0);
        return run_ipc(h, nsteps);
    }
    if (h->nranks > 1 && h->ngroup > 1) return fail("use run_group for in-process groups");
    axb_handle one[1] = {h};
    if (h->ngroup == 0) { h->group = (axo_t **)malloc(sizeof(axo_t *)); h->group[0] = h; h->ngroup = 1; }
    return axo_run_group(one, 1, nsteps);
}

int axo_set_stf_values(axb_handle h, int32_t first, int32_t n, const float *v) {
    if (first < 0 || first + n > h->niter_stf) return fail("stf range");
    memcpy(h->stf + first, v, sizeof(float) * n);
    return 0;
}
int axo_get_stf_symp(axb_handle h, int32_t first, int32_t n, float *out) {
    double t = 0.0, subdt[40], stf_symp[40];
    if (h->scheme == AXB_NEWMARK2 || h->nstages <= 0) return fail("axo_get_stf_symp: symplectic schemes only, after finalize_setup");
    if (first < 0 || n < 0 || first + n > h->niter) return fail("stf range");
    for (int it = 0; it < first + n; it++) {          /* t accumulated as in time_evol_wave.F90:586 */
        t += h->deltat;
        if (it < first) continue;
        for (int k = 0; k < h->nstages; k++) subdt[k] = t - h->deltat + h->coeff[k];
        compute_stf_t(h, h->nstages, subdt, stf_symp);
        for (int k = 0; k < h->nstages; k++) out[(size_t)(it - first) * h->nstages + k] = (float)stf_symp[k];
    }
    return 0;
}
int axo_profile(axb_handle h, int32_t e) { (void)h; (void)e; return 0; }
int axo_get_profile(axb_handle h, double *ms, int64_t *n) {
    (void)h; for (int i = 0; i < 8; i++) { ms[i] = 0.0; n[i] = 0; } return 0;
}
int axo_set_stream(axb_handle h, void *s) { (void)h; (void)s; return 0; }
int axo_synchronize(axb_handle h) { (void)h; return 0; }

int32_t axo_iter(axb_handle h) { return h->iter; }
int32_t axo_nseismo(axb_handle h) { return h->iseismo; }
int32_t axo_nstrain(axb_handle h) { return h->istrain; }
int64_t axo_gpu_launches(axb_handle h) { (void)h; return 0; }

int axo_fetch_seismograms(axb_handle h, int32_t first, int32_t nsamples, float *out) {
    if (first < 0 || first + nsamples > h->iseismo) return fail("seismogram range");
    memcpy(out, h->recdump + (size_t)3 * h->num_rec * first, sizeof(float) * 3 * h->num_rec * nsamples);
    return 0;
}
int axo_fetch_energy(axb_handle h, int32_t first, int32_t n, float *out) {
    if (!h->dump_energy) return fail("energy diagnostic not enabled (axo_set_energy)");
    if (first < 0 || n < 0 || first + n > h->iter + 1) return fail("fetch_energy: range beyond the computed samples");
    memcpy(out, h->energy + (size_t)4 * first, sizeof(float) * 4 * n);
    return 0;
}
int axo_fetch_snapshots(axb_handle h, int32_t first, int32_t nsnap, float *out) {
    size_t npts = snapshot_npoints(h);
    if (first < 0 || first + nsnap > h->istrain) return fail("snapshot range");
    for (int v = 0; v < snapshot_nvars(h); v++)
        for (int s = 0; s < nsnap; s++)
            memcpy(out + npts * (s + (size_t)nsnap * v),
                   h->snapdump + npts * ((first + s) + (size_t)h->nstrain_max * v), sizeof(float) * npts);
    return 0;
}

static float *field_ptr(axo_t *o, int f, size_t *n) {
    size_t ns = (size_t)NPT * o->nel_s * 3, nf = (size_t)NPT * o->nel_f;
    size_t per = o->cg ? 4 : NPT;
    switch (f) {
    case AXB_F_DISP: *n = ns; return o->disp;
    case AXB_F_VELO: *n = ns; return o->velo;
    case AXB_F_ACC0: *n = ns; return o->acc0;
    case AXB_F_ACC1: *n = ns; return o->acc1;
    case AXB_F_CHI: *n = nf; return o->chi;
    case AXB_F_DCHI: *n = nf; return o->dchi;
    case AXB_F_DDCHI0: *n = nf; return o->ddchi0;
    case AXB_F_DDCHI1: *n = nf; return o->ddchi1;
    case AXB_F_MEMVAR: *n = per * 6 * o->n_sls * o->nel_s; return o->memvar;
    case AXB_F_SRC_DEV_TM1: *n = per * 6 * o->nel_s; return o->src_dev_tm1;
    case AXB_F_SRC_TR_TM1: *n = per * o->nel_s; return o->src_tr_tm1;
    }
    return NULL;
}
int axo_get_state(axb_handle h, int32_t field, float *out) {
    size_t n; float *p = field_ptr(h, field, &n);
    if (!p) return fail("no such field");
    memcpy(out, p, n * sizeof(float));
    return 0;
}
int axo_set_state(axb_handle h, int32_t field, const float *in) {
    size_t n; float *p = field_ptr(h, field, &n);
    if (!p) return fail("no such field");
    memcpy(p, in, n * sizeof(float));
    return 0;
}

int axo_apply_op(axb_handle h, int32_t op) {
    set_ftz();
    if (!h->finalized) return fail("finalize_setup not called");
    switch (op) {
    case AXB_OP_SOLID_STIFFNESS: solid_stiffness(h, h->acc1, h->disp); return 0;
    case AXB_OP_ANEL_STIFFNESS: if (!h->anel) return fail("no attenuation"); anel_stiffness(h, h->acc1); return 0;
    case AXB_OP_FLUID_STIFFNESS: glob_fluid_stiffness_4(h, h->ddchi1, h->chi); return 0;
    case AXB_OP_PDISTSUM_SOLID:
        if (h->halo[0].nmsg) return fail("apply_op(pdistsum) is single-rank only");
        gather_scatter(h, AXB_DOMAIN_SOLID, h->acc1); return 0;
    case AXB_OP_PDISTSUM_FLUID:
        if (h->halo[1].nmsg) return fail("apply_op(pdistsum) is single-rank only");
        gather_scatter(h, AXB_DOMAIN_FLUID, h->ddchi1); return 0;
    case AXB_OP_MEMVARS: if (!h->anel) return fail("no attenuation"); time_step_memvars(h); return 0;
    case AXB_OP_BDRY2FLUID: bdry_copy2fluid(h, h->ddchi1, h->disp); return 0;
    case AXB_OP_BDRY2SOLID: bdry_copy2solid(h, h->acc1, h->ddchi1); return 0;
    }
    return fail("unknown op");
}
