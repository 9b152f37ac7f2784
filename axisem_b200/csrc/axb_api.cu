// axb_api.cu — C ABI (include/axisem_b200.h) of the device-resident AxiSEM time loop.
//
// Set-up calls copy the host's module arrays to HBM once; axb_run only enqueues kernels on
// the handle's stream (no host round trip per step: iteration counter, STF sample index
// and seismogram cursor live on the device).  There is no CPU fallback: every entry point
// either runs on the GPU or fails with an error.
#pragma GCC visibility push(default)
#include "../../include/axisem_b200.h"
#pragma GCC visibility pop
#include "axb_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace axb;

namespace {

thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return 1; }

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess)                                                           \
            return fail(std::string(#call) + ": " + cudaGetErrorString(_e));             \
    } while (0)

constexpr int MAXMSG = 8;
typedef void (*solid_kernel_t)(GMat, SolidTileArgs);

struct Halo {
    int nmsg = 0, nc = 1;
    int peer[MAXMSG] = {0}, size[MAXMSG] = {0}, offset[MAXMSG] = {0};
    std::vector<std::vector<int>> glocal;     // per message, 1-based glocal ids
    int ncomm = 0;
    std::vector<int> glob2el;                 // (ncomm,3) Fortran order
    int nslots = 0;                           // sum of sizes
    // device
    int nentries = 0;
    int *d_start = nullptr, *d_addr = nullptr, *d_dst_msg = nullptr, *d_dst_slot = nullptr;
    // receive slab (owned): [parity 2][nc][nslots] words {value, exchange number}, written by
    // the neighbours (axb_kernels.cuh: halo_read / k_halo_pack)
    int2 *recv = nullptr;
    // where my messages go (peer memory)
    int2 *peer_recv[MAXMSG] = {nullptr};
    int peer_nslots[MAXMSG] = {0}, peer_offset[MAXMSG] = {0};
    int seq = 0;                              // exchanges done
};

}  // namespace

struct axb_handle_s {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int nel_s = 0, nel_f = 0, nglob_s = 0, nglob_f = 0;
    std::vector<int> igloc_s, igloc_f, axis_s_h, axis_f_h;
    int *d_axis_s = nullptr, *d_axis_f = nullptr;
    GMat G;
    int order = 0;
    // solid element kernel inputs in their device layout (axb_solid_tile.cuh)
    int te_s = TES;                // elements per solid tile of the chosen variant
    int nel_pad_s = 0;             // nel_s rounded up to whole tiles
    size_t css = 0;                // component stride of disp/velo/acc* = 25 * nel_pad_s
    float *d_coef = nullptr;       // [tile][plane][TP]
    int *d_meta = nullptr;         // [tile][3][TE]: axis, qidx_mu, qidx_ka
    float *d_M0_w[10] = {nullptr};
    bool have_solid_terms = false;
    void (*solid_kernel)(GMat, SolidTileArgs) = nullptr;
    int nst = 0;                   // ring depth of k_solid_tile
    size_t smem_solid = 0;
    std::vector<void *> allocs;
    // fluid
    int nel_pad_f = 0;
    float *d_coef_f = nullptr;     // [tile][npl_f][TP]: M1chi, M2chi, M4chi [, M_w_fl] [, fs_mask]
    int npl_f = 0, mask_plane_f = -1;
    int *d_meta_f = nullptr;       // [tile][3][TE]
    float *M0_w_fl = nullptr;
    float *inv_mass_fluid = nullptr;
    int nst_f = 0;
    size_t smem_fluid = 0;
    float *inv_mass_rho = nullptr, *gamma_s = nullptr, *gamma_f = nullptr;
    // dump_energy
    bool dump_energy = false;
    float *um_rho_s = nullptr, *um_lam_f = nullptr;
    double *d_energy = nullptr;    // (4, niter + 1): epot_sol, ekin_sol, epot_flu, ekin_flu
    int ienergy = 0;               // next energy sample (== iter it belongs to)
    // boundary
    int nel_bdry = 0;
    std::vector<int> bdry_fel_h, bdry_jf_h;
    int *d_bsel = nullptr, *d_bfel = nullptr, *d_bjs = nullptr, *d_bjf = nullptr;
    float *d_bmatr = nullptr;
    // attenuation
    bool anel = false, cg = true;
    int n_sls = 0;
    float *d_cg = nullptr;         // [tile][NCG][TE*4]
    float *d_inv_s = nullptr;      // (25 * nel_pad_s)
    double2 *d_c_mu_tab = nullptr, *d_c_ka_tab = nullptr;
    int ntab_mu = 0, ntab_ka = 0;  // rows (distinct Q values) of the two tables
    int tab_smem = 0, ring_off = 0; // S_A: tables staged in shared memory, ring offset
    std::vector<double> ts_t_h, ts_tm1_h, exp_w_h;
    float *memvar = nullptr, *src_dev_tm1 = nullptr, *src_tr_tm1 = nullptr;
    // COARSE_GRAINED false: flat (25 nel) planes and (5 nel) axial vectors (axb_anel_full.cuh)
    const float *f_Y = nullptr, *f_Vse = nullptr, *f_Vsx = nullptr, *f_Vze = nullptr, *f_Vzx = nullptr;
    const float *f_Y0 = nullptr, *f_V0se = nullptr, *f_V0sx = nullptr, *f_V0ze = nullptr, *f_V0zx = nullptr;
    const float *f_Dse = nullptr, *f_Dze = nullptr, *f_Dsx = nullptr, *f_Dzx = nullptr;
    const float *f_dmu = nullptr, *f_dka = nullptr;
    int *d_qidx_mu = nullptr, *d_qidx_ka = nullptr;
    std::vector<float> Qmu_h, Qka_h;
    std::vector<double> y_j;
    int corr_lowq = 0;
    // source
    int fluid_src = 0, nelsrc = 0, ielsrc[8] = {0}, niter_stf = 0;
    int src_emin = 0, src_emax = -1;     // 0-based range spanned by ielsrc
    float *d_src_term = nullptr, *d_stf = nullptr;
    int stf_type = 0;
    double decay = 0, t_0 = 1, shift = 0, magnitude = 0;
    // receivers / dump
    int num_rec = 0;
    int *d_recfile = nullptr;
    float *d_recdump = nullptr;
    int nseismo_max = 0;
    bool have_kwf = false;
    int npt_s_kwf = 0, npt_f_kwf = 0, nstrain_max = 0;
    int *d_kwf_mask = nullptr, *d_kwf_map = nullptr;
    float *d_inv_rho = nullptr, *d_Dse_f = nullptr, *d_Dze_f = nullptr, *d_Dsx_f = nullptr, *d_Dzx_f = nullptr;
    float *d_snap = nullptr;
    // xdmf snapshots (axb_set_xdmf)
    bool have_xdmf = false;
    int snap_it = 0, isnap = 0, nsnap_max = 0, npoint_plot = 0;
    int *d_xmap = nullptr;
    float *d_xsnap = nullptr;
    float *xs_Dse = nullptr, *xs_Dze = nullptr, *xs_Dsx = nullptr, *xs_Dzx = nullptr, *xs_inv_s = nullptr;
    float *xf_Dse = nullptr, *xf_Dze = nullptr, *xf_Dsx = nullptr, *xf_Dzx = nullptr, *xf_inv_s = nullptr, *xf_inv_rho = nullptr;
    // dump_type strain_only / fullfields (axb_set_dump)
    int dump_type = AXB_DUMP_DISPL_ONLY, ibeg = 0, iend = 4, jbeg = 0, jend = 4;
    float *dDse = nullptr, *dDze = nullptr, *dDsx = nullptr, *dDzx = nullptr, *d_inv_s_dump = nullptr, *d_inv_s_f = nullptr;
    Halo halo[2];
    // time
    int scheme = 0, niter = 0, seis_it = 1, strain_it = 0, nstages = 0;
    double deltat = 0, half_dt = 0, half_dt_sq = 0;
    double coefd[40], coefv[40], coeff[40];
    float *d_stf_symp = nullptr;
    // state
    float *disp = nullptr, *velo = nullptr, *acc0 = nullptr, *acc1 = nullptr;
    float *chi = nullptr, *dchi = nullptr, *ddchi0 = nullptr, *ddchi1 = nullptr;
    int4 *d_asm_cp_s = nullptr, *d_asm_cp_f = nullptr;
    int *d_asm_grp_s = nullptr, *d_asm_grp_f = nullptr;
    int *d_counters = nullptr;     // [2] halo abort flag, [3] first blown-up iteration
    int iter = 0, iseismo = 0, istrain = 0;
    // lean Newmark: between steps the velo / dchi buffers hold v + dt/2 a and
    // dchi + dt/2 ddchi, acc0 is not maintained; the reference's state is re-formed on demand
    bool lean = false;             // formulation selected at finalize_setup
    bool lean_state = false;       // the buffers currently hold the lean quantities
    int lean_entry_iter = 0;       // iter at which they were formed
    // CUDA-graph replay of a Newmark step: device-resident step counters + the instantiated graph
    bool use_graph = false;
    int *d_dyn = nullptr;
    bool dyn_synced = false;
    cudaGraphExec_t step_graph = nullptr;
    int step_graph_nodes = 0;
    unsigned long long halo_timeout_ns = 10000000000ull;
    bool have_M_w_fl = false;
    std::vector<void *> ipc_opened;
    bool acc1_is_acc0 = false;     // after a full step acc1/ddchi1 == acc0/ddchi0 in the reference
    bool finalized = false;
    int64_t launches = 0;
    int grid_s = 0, grid_f = 0, grid_ft = 0, sms = 0;
    // per-kernel event timing (axb_profile)
    int prof = 0;                         // 0 off, 1 every launch, 2 the solid element kernel only
    int prof_cls = 7;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cls;              // class of each (start, stop) pair
    size_t ev_used = 0;
    double prof_ms[8] = {0};
    int64_t prof_n[8] = {0};
    std::vector<axb_handle_s *> group;
};

namespace {

template <class T>
int upload(axb_handle_s *h, T *&dst, const T *src, size_t n) {
    dst = nullptr;
    if (!src) return 0;
    CK(cudaMalloc((void **)&dst, std::max<size_t>(n, 1) * sizeof(T)));
    h->allocs.push_back(dst);
    if (n) CK(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <class T>
int dzeros(axb_handle_s *h, T *&dst, size_t n) {
    CK(cudaMalloc((void **)&dst, std::max<size_t>(n, 1) * sizeof(T)));
    h->allocs.push_back(dst);
    CK(cudaMemset(dst, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}
// upload n values into a zero-filled allocation of npad values
template <class T>
int upload_padded(axb_handle_s *h, T *&dst, const T *src, size_t n, size_t npad) {
    dst = nullptr;
    if (!src) return 0;
    CK(cudaMalloc((void **)&dst, std::max<size_t>(npad, 1) * sizeof(T)));
    h->allocs.push_back(dst);
    CK(cudaMemset(dst, 0, std::max<size_t>(npad, 1) * sizeof(T)));
    if (n) CK(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
int use(axb_handle_s *h) {
    CK(cudaSetDevice(h->device));
    return 0;
}
#define UP(dst, src, n) do { if (upload(h, dst, src, (size_t)(n))) return 1; } while (0)
#define UPC(dst, src, n) do { float *_t; if (upload(h, _t, src, (size_t)(n))) return 1; dst = _t; } while (0)

// symplectic_coefficients / SS_scheme: time_evol_wave.F90:749-992 (incl. 1/2 == 0 at :981)
void ss_scheme(int n, double *a, double *b, const double *g) {
    double s = 0.0;
    a[0] = g[0] / 2.0;
    for (int i = 1; i < n; i++) a[i] = (g[i - 1] + g[i]) / 2.0;
    for (int i = 0; i < n; i++) s += a[i];
    a[n] = (double)(1 / 2) - s;
    for (int i = n + 2; i <= 2 * n + 2; i++) a[i - 1] = a[2 * n + 3 - i - 1];
    for (int i = 0; i < n; i++) b[i] = g[i];
    s = 0.0;
    for (int i = 0; i < n; i++) s += g[i];
    b[n] = 1.0 - 2.0 * s;
    for (int i = n + 2; i <= 2 * n + 1; i++) b[i - 1] = b[2 * n + 2 - i - 1];
}

int symplectic_coefficients(axb_handle_s *o) {
    double *d = o->coefd, *v = o->coefv;
    int n, ns = 0;
    switch (o->scheme) {
    case AXB_SYMPLEC4: {
        // default-real literals in the Fortran source -> single-precision values
        double zeta = (double)0.1786178958448091f, iota = (double)-0.2123418310626054f,
               kappa = (double)-0.06626458266981849f;
        ns = 4;
        d[0] = zeta; d[1] = kappa; d[2] = 1.0 - 2.0 * (zeta + kappa); d[3] = kappa; d[4] = zeta;
        v[0] = 0.5 - iota; v[1] = iota; v[2] = iota; v[3] = 0.5 - iota;
        break; }
    case AXB_ML_SO4M5: {
        double rho = (14.0 - std::sqrt(19.0)) / 108.0, theta = (20.0 - 7.0 * std::sqrt(19.0)) / 108.0;
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
        ss_scheme(n, d, v, g);
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
        ss_scheme(n, d, v, g);
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

// the reference's own error function (source.f90:662-692; coefficients are default-real literals)
double erf_nr(double x) {
    static const float c[10] = {-1.26551223f, 1.00002368f, 0.37409196f, 0.09678418f, -0.18628806f,
                                0.27886807f, -1.13520398f, 1.48851587f, -0.82215223f, 0.17087277f};
    const double z = std::fabs(x), t = 1.0 / (1.0 + 0.5 * z);
    double poly = (double)c[9];
    for (int k = 8; k >= 0; k--) poly = t * poly + (double)c[k];
    double erfcc = t * std::exp(-z * z + poly);
    if (x < 0.0) erfcc = 2.0 - erfcc;
    return 1.0 - erfcc;
}
// compute_stf_t (source.f90:206-233) on the nstages sub-stage times `t` of one step.  The smooth
// types are point-wise; delta_src_t (:890-904) is a hat function switched by the FIRST sub-stage
// time of the step, quasiheavi_t (:908-917) sets the sub-stages seis_it..nstages (an index, not a
// time) to the magnitude.
void stf_t(const axb_handle_s *o, int n, const double *t, double *out) {
    const double a = o->decay / o->t_0, pi = 3.1415926535898, dt = o->deltat, sh = o->shift;
    for (int k = 0; k < n; k++) {
        const double x = a * (t[k] - sh);
        switch (o->stf_type) {
        case AXB_STF_GAUSS_0: out[k] = std::exp(-(x * x)) * o->magnitude * a / std::sqrt(pi); break;
        case AXB_STF_GAUSS_1: out[k] = -2.0 * a * a * (t[k] - sh) * std::exp(-(x * x))
                                       / (a * std::sqrt(2.0) * std::exp(-0.5)) * o->magnitude; break;
        case AXB_STF_GAUSS_2: out[k] = a * a * (2.0 * a * a * (t[k] - sh) * (t[k] - sh) - 1.0)
                                       * std::exp(-(x * x)) / (2.0 * a * a * std::exp(-1.5)) * o->magnitude; break;
        case AXB_STF_ERRORF: out[k] = (erf_nr(x) * 0.5 + 0.5) * o->magnitude; break;
        case AXB_STF_QUHEAVI: out[k] = (k + 1 >= o->seis_it) ? o->magnitude : 0.0; break;
        default: out[k] = 0.0; break;
        }
    }
    if (o->stf_type == AXB_STF_DIRAC_0) {
        if (t[0] > sh - dt && t[0] <= sh)
            for (int k = 0; k < n; k++) out[k] = (t[k] - t[0]) / dt * o->magnitude / dt;
        if (t[0] >= sh && t[0] < sh + dt)
            for (int k = 0; k < n; k++) out[k] = (1.0 - (t[k] - t[0]) / dt) * o->magnitude / dt;
    }
}

void fast_correct(int n, const double *y, double *yp) {      // attenuation.f90:1139-1155
    double dy[32];
    dy[0] = 1 + .5 * y[0];
    for (int k = 1; k < n; k++) dy[k] = dy[k - 1] + (dy[k - 1] - .5) * y[k - 1] + .5 * y[k];
    for (int k = 0; k < n; k++) yp[k] = y[k] * dy[k];
}
void a_j_of_Q(const axb_handle_s *o, float Q, double *a_j) { // attenuation.f90:116-134
    const int n = o->n_sls;
    double yq[32], yp[32], s = 0.0;
    for (int k = 0; k < n; k++) yq[k] = o->y_j[k] / Q;
    if (o->corr_lowq) fast_correct(n, yq, yp);
    else std::memcpy(yp, yq, sizeof(double) * n);
    for (int k = 0; k < n; k++) s += yp[k];
    for (int k = 0; k < n; k++) a_j[k] = yp[k] / s;
}

// Pull-style assembly table (see AsmTable).  Only the 16 edge points of an element take
// part (commun.F90:110-120); members are listed in ascending element order.
int build_asm(axb_handle_s *h, int nel, int nglob, const std::vector<int> &igloc, Halo &H,
              int4 *&d_cp, int *&d_grp) {
    static const int edge_q[16] = {0, 1, 2, 3, 4, 5, 9, 10, 14, 15, 19, 20, 21, 22, 23, 24};
    const size_t npts = (size_t)NPT * nel;
    std::vector<int> count(nglob + 1, 0);
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < 16; k++) count[igloc[(size_t)e * NPT + edge_q[k]] - 1 + 1]++;
    // remote contributions per glocal id: slots in message order
    std::vector<int> rcount(nglob, 0);
    for (int m = 0; m < H.nmsg; m++)
        for (int ip = 0; ip < H.size[m]; ip++) rcount[H.glocal[m][ip] - 1]++;
    std::vector<int64_t> start(nglob + 1, 0);
    for (int g = 0; g < nglob; g++) start[g + 1] = start[g] + count[g + 1];
    std::vector<int> members(start[nglob]);
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < 16; k++) {
            const size_t p = (size_t)e * NPT + edge_q[k];
            members[fill[igloc[p] - 1]++] = (int)p;
        }
    // group offsets
    std::vector<int64_t> goff(nglob, -1);
    int64_t total = 0;
    for (int g = 0; g < nglob; g++) {
        const int nloc = count[g + 1];
        if (nloc + rcount[g] >= 2 && nloc >= 1) { goff[g] = total; total += 2 + nloc + rcount[g]; }
    }
    if (total > 0x7fffffffLL) return fail("assembly table too large for 32-bit offsets");
    std::vector<int> grp(std::max<int64_t>(total, 1), 0);
    std::vector<int> rfill(nglob, 0);
    for (int g = 0; g < nglob; g++) {
        if (goff[g] < 0) continue;
        const int nloc = count[g + 1];
        grp[goff[g]] = nloc;
        grp[goff[g] + 1] = rcount[g];
        for (int m = 0; m < nloc; m++) grp[goff[g] + 2 + m] = members[start[g] + m];
    }
    for (int m = 0; m < H.nmsg; m++)
        for (int ip = 0; ip < H.size[m]; ip++) {
            const int g = H.glocal[m][ip] - 1;
            if (goff[g] < 0) return fail("halo point without a local edge copy");
            grp[goff[g] + 2 + count[g + 1] + rfill[g]++] = H.offset[m] + ip;
        }
    // per edge point: the int4 fast entry (see AsmTable)
    std::vector<int4> cp((size_t)std::max(nel, 1) * 16, make_int4(-1, -1, -1, -1));
    for (int e = 0; e < nel; e++)
        for (int k = 0; k < 16; k++) {
            const size_t p = (size_t)e * NPT + edge_q[k];
            const int g = igloc[p] - 1;
            const int64_t o = goff[g];
            if (o < 0) continue;
            const int nloc = count[g + 1];
            int4 &c = cp[(size_t)e * 16 + k];
            if (rcount[g] == 0 && nloc <= 4) {
                int m[4] = {-1, -1, -1, -1};
                for (int t = 0; t < nloc; t++) m[t] = members[start[g] + t];
                c = make_int4(m[0], m[1], m[2], m[3]);
            } else {
                c = make_int4(-2, (int)o, 0, 0);
            }
        }
    if (upload(h, d_cp, cp.data(), cp.size())) return 1;
    if (upload(h, d_grp, grp.data(), grp.size())) return 1;
    (void)npts;
    return 0;
}

// Halo send side: CSR of local copies per send entry in glob2el order (commpi.F90:383-391)
int build_halo_send(axb_handle_s *h, int nel, int nglob, const std::vector<int> &igloc, Halo &H) {
    if (H.nmsg == 0) return 0;
    std::vector<std::vector<int>> mem(nglob);
    for (int ip = 0; ip < H.ncomm; ip++) {
        const int ipol = H.glob2el[ip], jpol = H.glob2el[ip + H.ncomm], iel = H.glob2el[ip + 2 * H.ncomm];
        const size_t ipt = (size_t)(iel - 1) * NPT + jpol * NP + ipol;
        mem[igloc[ipt] - 1].push_back((int)ipt);
    }
    std::vector<int> start(1, 0), addr, dmsg, dslot;
    for (int m = 0; m < H.nmsg; m++)
        for (int ip = 0; ip < H.size[m]; ip++) {
            const auto &v = mem[H.glocal[m][ip] - 1];
            addr.insert(addr.end(), v.begin(), v.end());
            start.push_back((int)addr.size());
            dmsg.push_back(m);
            dslot.push_back(ip);           // + peer offset, added at connect time
        }
    H.nentries = (int)dmsg.size();
    if (addr.empty()) addr.push_back(0);
    if (upload(h, H.d_start, start.data(), start.size())) return 1;
    if (upload(h, H.d_addr, addr.data(), addr.size())) return 1;
    if (upload(h, H.d_dst_msg, dmsg.data(), dmsg.size())) return 1;
    if (upload(h, H.d_dst_slot, dslot.data(), dslot.size())) return 1;
    (void)nel;
    return 0;
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

// ======================================================================================
#pragma GCC visibility push(default)
extern "C" {

const char *axb_last_error(void) { return g_err.c_str(); }
static int snapshot_nvars(const axb_handle_s *h);
static size_t snapshot_npoints(const axb_handle_s *h);

int axb_create(axb_handle *out, int32_t device, int32_t rank, int32_t nranks) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("axisem_b200: no CUDA device (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("axb_create: bad device ordinal");
    axb_handle_s *h = new axb_handle_s();
    h->device = device; h->rank = rank; h->nranks = nranks;
    if (use(h)) { delete h; return 1; }
    cudaError_t se = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete h; return fail(cudaGetErrorString(se)); }
    h->own_stream = true;
    *out = h;
    return 0;
}

int axb_destroy(axb_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : h->allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int axb_set_stream(axb_handle h, void *cuda_stream) {
    if (use(h)) return 1;
    if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
    return 0;
}

int axb_set_mesh(axb_handle h, int32_t npol, int32_t nel_solid, int32_t nel_fluid,
                 int32_t nglob_solid, int32_t nglob_fluid, const int32_t *igloc_solid,
                 const int32_t *igloc_fluid, const int32_t *axis_solid,
                 const int32_t *axis_fluid, const int32_t *ax_el_solid, int32_t naxel_solid,
                 const int32_t *ax_el_fluid, int32_t naxel_fluid, const float *G0,
                 const float *G1, const float *G1T, const float *G2, const float *G2T) {
    if (npol != 4) return fail("axb_set_mesh: npol must be 4");
    if (use(h)) return 1;
    h->nel_s = nel_solid; h->nel_f = nel_fluid; h->nglob_s = nglob_solid; h->nglob_f = nglob_fluid;
    h->te_s = TES;
    h->nel_pad_s = (nel_solid + h->te_s - 1) / h->te_s * h->te_s;
    h->nel_pad_f = (nel_fluid + TE - 1) / TE * TE;
    h->css = (size_t)NPT * h->nel_pad_s;
    h->igloc_s.assign(igloc_solid, igloc_solid + (size_t)NPT * nel_solid);
    if (nel_fluid) h->igloc_f.assign(igloc_fluid, igloc_fluid + (size_t)NPT * nel_fluid);
    for (size_t p = 0; p < h->igloc_s.size(); p++)
        if (h->igloc_s[p] < 1 || h->igloc_s[p] > nglob_solid) return fail("igloc_solid out of range");
    for (size_t p = 0; p < h->igloc_f.size(); p++)
        if (h->igloc_f[p] < 1 || h->igloc_f[p] > nglob_fluid) return fail("igloc_fluid out of range");
    // the axial element lists must agree with the logical flags (def_grid.f90:59-77)
    {
        int n = 0;
        for (int e = 0; e < nel_solid; e++) n += axis_solid[e] != 0;
        if (n != naxel_solid) return fail("axis_solid / ax_el_solid mismatch");
        for (int a = 0; a < naxel_solid; a++)
            if (ax_el_solid[a] < 1 || ax_el_solid[a] > nel_solid || !axis_solid[ax_el_solid[a] - 1])
                return fail("ax_el_solid inconsistent with axis_solid");
        n = 0;
        for (int e = 0; e < nel_fluid; e++) n += axis_fluid[e] != 0;
        if (n != naxel_fluid) return fail("axis_fluid / ax_el_fluid mismatch");
        for (int a = 0; a < naxel_fluid; a++)
            if (ax_el_fluid[a] < 1 || ax_el_fluid[a] > nel_fluid || !axis_fluid[ax_el_fluid[a] - 1])
                return fail("ax_el_fluid inconsistent with axis_fluid");
    }
    h->axis_s_h.assign(axis_solid, axis_solid + nel_solid);
    h->axis_f_h.assign(axis_fluid, axis_fluid + nel_fluid);
    UP(h->d_axis_s, (const int *)axis_solid, nel_solid);
    UP(h->d_axis_f, (const int *)axis_fluid, nel_fluid);
    std::memcpy(h->G.G0, G0, sizeof h->G.G0);
    std::memcpy(h->G.G1, G1, sizeof h->G.G1); std::memcpy(h->G.G1T, G1T, sizeof h->G.G1T);
    std::memcpy(h->G.G2, G2, sizeof h->G.G2); std::memcpy(h->G.G2T, G2T, sizeof h->G.G2T);
    return 0;
}

// planes (0:4,0:4,nel_solid) -> one coefficient slab [tile][plane][TP] (axb_solid_tile.cuh)
static int plane_to_slab(axb_handle_s *h, const float *host_plane, float *d_tmp, int pl, int npl) {
    const size_t n = (size_t)NPT * h->nel_s;
    if (n == 0) return 0;
    CK(cudaMemcpy(d_tmp, host_plane, n * sizeof(float), cudaMemcpyHostToDevice));
    k_plane_to_slab<<<(unsigned)((n + 255) / 256), 256>>>(d_tmp, h->d_coef, pl, npl, h->nel_s, h->te_s);
    CK(cudaGetLastError());
    return 0;
}

int axb_set_solid_terms(axb_handle h, int32_t src_order, const axb_solid_terms *t) {
    if (use(h)) return 1;
    if (src_order < 0 || src_order > 2) return fail("bad src_order");
    h->order = src_order;
    if (h->nel_s == 0) { h->have_solid_terms = true; return 0; }
    // slab plane order: the enum in axb_solid_tile.cuh
    std::vector<const float *> pl = {t->M11s, t->M21s, t->M41s, t->M12s, t->M22s, t->M32s, t->M42s,
                                     t->M11z, t->M21z, t->M41z, t->M_1, t->M_2, t->M_3, t->M_4, t->M_w1};
    std::vector<const float *> w0 = {t->M0_w1, t->M0_w2, t->M0_w3, t->M0_w4, t->M0_w5,
                                     t->M0_w6, t->M0_w7, t->M0_w8, t->M0_w9, t->M0_w10};
    bool ok0 = t->M0_w1 && t->M0_w2 && t->M0_w3;
    if (src_order == AXB_DIPOLE) {
        for (const float *q : {t->M13s, t->M33s, t->M43s, t->M_5, t->M_6, t->M_7, t->M_8, t->M_w2, t->M_w3}) pl.push_back(q);
        ok0 = ok0 && t->M0_w4 && t->M0_w6 && t->M0_w7 && t->M0_w8 && t->M0_w9 && t->M0_w10;
    } else if (src_order == AXB_QUADPOLE) {
        for (const float *q : {t->M1phi, t->M2phi, t->M4phi, t->M_5, t->M_6, t->M_7, t->M_8, t->M_w2, t->M_w3,
                               t->M_w4, t->M_w5}) pl.push_back(q);
        ok0 = ok0 && t->M0_w4 && t->M0_w5 && t->M0_w6;
    }
    const int npl = solid_nplanes(src_order);
    if ((int)pl.size() != npl) return fail("internal: plane list");
    for (const float *q : pl)
        if (!q) return fail("axb_set_solid_terms: a plane required for this source order is NULL");
    if (!ok0) return fail("axb_set_solid_terms: an axial vector required for this source order is NULL");
    const size_t ntiles = h->nel_pad_s / h->te_s;
    if (dzeros(h, h->d_coef, ntiles * npl * h->te_s * NPT)) return 1;
    float *d_tmp = nullptr;
    CK(cudaMalloc((void **)&d_tmp, (size_t)NPT * h->nel_s * sizeof(float)));
    for (int k = 0; k < npl; k++)
        if (plane_to_slab(h, pl[k], d_tmp, k, npl)) { cudaFree(d_tmp); return 1; }
    CK(cudaDeviceSynchronize());
    cudaFree(d_tmp);
    for (int k = 0; k < 10; k++)
        if (upload_padded(h, h->d_M0_w[k], w0[k], (size_t)NP * h->nel_s, (size_t)NP * h->nel_pad_s)) return 1;
    h->have_solid_terms = true;
    return 0;
}

int axb_set_fluid_terms(axb_handle h, const float *M1chi_fl, const float *M2chi_fl,
                        const float *M4chi_fl, const float *M_w_fl, const float *M0_w_fl,
                        const float *inv_mass_fluid, const float *fluid_free_surface_mask) {
    if (use(h)) return 1;
    const size_t n = (size_t)NPT * h->nel_f;
    if (h->nel_f == 0) return 0;
    if (!M1chi_fl || !M2chi_fl || !M4chi_fl || !inv_mass_fluid) return fail("axb_set_fluid_terms: NULL plane");
    // the free-surface mask is all ones unless the fluid reaches the surface: skip the plane then
    bool mask_needed = false;
    if (fluid_free_surface_mask)
        for (size_t p = 0; p < n; p++) if (fluid_free_surface_mask[p] != 1.0f) { mask_needed = true; break; }
    std::vector<const float *> pl = {M1chi_fl, M2chi_fl, M4chi_fl};
    // slot 3 is always M_w_fl for non-monopole sources (axb_fluid_tile.cuh reads it there)
    h->have_M_w_fl = M_w_fl != nullptr;
    if (M_w_fl) pl.push_back(M_w_fl);
    h->mask_plane_f = -1;
    if (mask_needed) { h->mask_plane_f = (int)pl.size(); pl.push_back(fluid_free_surface_mask); }
    h->npl_f = (int)pl.size();
    const size_t ntiles = h->nel_pad_f / TE;
    if (dzeros(h, h->d_coef_f, ntiles * h->npl_f * TP)) return 1;
    float *d_tmp = nullptr;
    CK(cudaMalloc((void **)&d_tmp, n * sizeof(float)));
    for (int k = 0; k < h->npl_f; k++) {
        CK(cudaMemcpy(d_tmp, pl[k], n * sizeof(float), cudaMemcpyHostToDevice));
        k_plane_to_slab<<<(unsigned)((n + 255) / 256), 256>>>(d_tmp, h->d_coef_f, k, h->npl_f, h->nel_f, TE);
        CK(cudaGetLastError());
    }
    CK(cudaDeviceSynchronize());
    cudaFree(d_tmp);
    if (upload_padded(h, h->M0_w_fl, M0_w_fl, (size_t)NP * h->nel_f, (size_t)NP * h->nel_pad_f)) return 1;
    UP(h->inv_mass_fluid, inv_mass_fluid, n);
    return 0;
}

int axb_set_mass(axb_handle h, const float *inv_mass_rho) {
    if (use(h)) return 1;
    if (!inv_mass_rho && h->nel_s > 0) return fail("axb_set_mass: NULL inv_mass_rho");
    if (upload_padded(h, h->inv_mass_rho, inv_mass_rho, (size_t)NPT * h->nel_s, h->css)) return 1;
    return 0;
}

int axb_set_energy(axb_handle h, const float *unassem_mass_rho_solid, const float *unassem_mass_lam_fluid) {
    if (use(h)) return 1;
    if (!unassem_mass_rho_solid && h->nel_s > 0) return fail("axb_set_energy: NULL unassem_mass_rho_solid");
    if (!unassem_mass_lam_fluid && h->nel_f > 0) return fail("axb_set_energy: NULL unassem_mass_lam_fluid");
    if (h->nel_s > 0) UP(h->um_rho_s, unassem_mass_rho_solid, (size_t)NPT * h->nel_s);
    if (h->nel_f > 0) UP(h->um_lam_f, unassem_mass_lam_fluid, (size_t)NPT * h->nel_f);
    h->dump_energy = true;
    return 0;
}

int axb_set_sponge(axb_handle h, const float *solid_gamma, const float *fluid_gamma) {
    if (use(h)) return 1;
    if (!solid_gamma && !fluid_gamma) return 0;
    if (solid_gamma) UP(h->gamma_s, solid_gamma, (size_t)NPT * h->nel_s);
    else if (dzeros(h, h->gamma_s, (size_t)NPT * h->nel_s)) return 1;
    if (fluid_gamma) UP(h->gamma_f, fluid_gamma, (size_t)NPT * h->nel_f);
    else if (dzeros(h, h->gamma_f, (size_t)NPT * h->nel_f)) return 1;
    return 0;
}

int axb_set_sf_boundary(axb_handle h, int32_t nel_bdry, const int32_t *bdry_solid_el,
                        const int32_t *bdry_fluid_el, const int32_t *bdry_jpol_solid,
                        const int32_t *bdry_jpol_fluid, const float *bdry_matr) {
    if (use(h)) return 1;
    for (int b = 0; b < nel_bdry; b++) {
        if (bdry_solid_el[b] < 1 || bdry_solid_el[b] > h->nel_s || bdry_fluid_el[b] < 1 ||
            bdry_fluid_el[b] > h->nel_f)
            return fail("S/F boundary element out of range");
        if ((bdry_jpol_solid[b] != 0 && bdry_jpol_solid[b] != 4) ||
            (bdry_jpol_fluid[b] != 0 && bdry_jpol_fluid[b] != 4))
            return fail("S/F boundary jpol must be 0 or npol");
    }
    h->nel_bdry = nel_bdry;
    h->bdry_fel_h.assign(bdry_fluid_el, bdry_fluid_el + nel_bdry);
    h->bdry_jf_h.assign(bdry_jpol_fluid, bdry_jpol_fluid + nel_bdry);
    UP(h->d_bsel, (const int *)bdry_solid_el, nel_bdry); UP(h->d_bfel, (const int *)bdry_fluid_el, nel_bdry);
    UP(h->d_bjs, (const int *)bdry_jpol_solid, nel_bdry); UP(h->d_bjf, (const int *)bdry_jpol_fluid, nel_bdry);
    UP(h->d_bmatr, bdry_matr, (size_t)NP * nel_bdry * 2);
    return 0;
}

int axb_set_attenuation(axb_handle h, const axb_attenuation *a) {
    if (use(h)) return 1;
    if (a->n_sls < 1 || a->n_sls > 8) return fail("n_sls must be in 1..8");
    const size_t n4 = (size_t)4 * h->nel_s, n = (size_t)NPT * h->nel_s;
    h->anel = true; h->cg = a->coarse_grained != 0; h->corr_lowq = a->do_corr_lowq;
    h->n_sls = a->n_sls;
    h->y_j.assign(a->y_j, a->y_j + a->n_sls);
    h->exp_w_h.assign(a->exp_w_j_deltat, a->exp_w_j_deltat + a->n_sls);
    h->ts_t_h.assign(a->ts_fac_t, a->ts_fac_t + a->n_sls);
    h->ts_tm1_h.assign(a->ts_fac_tm1, a->ts_fac_tm1 + a->n_sls);
    if (!a->Q_mu || !a->Q_kappa) return fail("axb_set_attenuation: NULL Q array");
    h->Qmu_h.assign(a->Q_mu, a->Q_mu + h->nel_s);
    h->Qka_h.assign(a->Q_kappa, a->Q_kappa + h->nel_s);
    if (!h->cg) {
        // memory variables at all 25 points (attenuation.f90:210-334): flat planes
        const float *pl[11] = {a->Y, a->V_s_eta, a->V_s_xi, a->V_z_eta, a->V_z_xi, a->DsDeta_over_J_sol,
                               a->DzDeta_over_J_sol, a->DsDxi_over_J_sol, a->DzDxi_over_J_sol,
                               a->delta_mu, a->delta_kappa};
        const float **dst[11] = {&h->f_Y, &h->f_Vse, &h->f_Vsx, &h->f_Vze, &h->f_Vzx, &h->f_Dse,
                                 &h->f_Dze, &h->f_Dsx, &h->f_Dzx, &h->f_dmu, &h->f_dka};
        for (int k = 0; k < 11; k++) {
            if (!pl[k]) return fail("axb_set_attenuation: NULL (0:4,0:4,nel_solid) array");
            UPC(*dst[k], pl[k], n);
        }
        const float *p0[5] = {a->Y0, a->V0_s_eta, a->V0_s_xi, a->V0_z_eta, a->V0_z_xi};
        const float **d0[5] = {&h->f_Y0, &h->f_V0se, &h->f_V0sx, &h->f_V0ze, &h->f_V0zx};
        for (int k = 0; k < 5; k++) {
            if (!p0[k]) return fail("axb_set_attenuation: NULL axial (0:4,nel_solid) array");
            UPC(*d0[k], p0[k], (size_t)NP * h->nel_s);
        }
        if (!a->inv_s_solid) return fail("axb_set_attenuation: NULL inv_s_solid");
        if (upload_padded(h, h->d_inv_s, a->inv_s_solid, n, h->css)) return 1;
        return 0;
    }
    // slab plane order: enum G_* in axb_solid_tile.cuh
    constexpr int NCGH = 11;      // planes that come from the host as (4, nel) arrays
    const float *cgp[NCGH] = {a->Y_cg4, a->V_s_eta_cg4, a->V_s_xi_cg4, a->V_z_eta_cg4, a->V_z_xi_cg4,
                             a->DsDeta_over_J_sol_cg4, a->DzDeta_over_J_sol_cg4,
                             a->DsDxi_over_J_sol_cg4, a->DzDxi_over_J_sol_cg4,
                             a->delta_mu_cg4, a->delta_kappa_cg4};
    for (int k = 0; k < NCGH; k++)
        if (!cgp[k]) return fail("axb_set_attenuation: NULL cg4 array");
    if (!a->inv_s_solid) return fail("axb_set_attenuation: NULL inv_s_solid");
    const size_t ntiles = h->nel_pad_s / h->te_s;
    if (dzeros(h, h->d_cg, ntiles * NCG * h->te_s * 4)) return 1;
    if (n4) {
        float *d_tmp = nullptr;
        CK(cudaMalloc((void **)&d_tmp, n4 * sizeof(float)));
        for (int k = 0; k < NCGH; k++) {
            CK(cudaMemcpy(d_tmp, cgp[k], n4 * sizeof(float), cudaMemcpyHostToDevice));
            k_cg_to_slab<<<(unsigned)((n4 + 255) / 256), 256>>>(d_tmp, h->d_cg, k, h->nel_s, h->te_s);
            CK(cudaGetLastError());
        }
        CK(cudaDeviceSynchronize());
        cudaFree(d_tmp);
    }
    if (n4) {
        // the coarse-grained kernels only need 1/s at the four coarse points
        float *d_tmp = nullptr;
        CK(cudaMalloc((void **)&d_tmp, n * sizeof(float)));
        CK(cudaMemcpy(d_tmp, a->inv_s_solid, n * sizeof(float), cudaMemcpyHostToDevice));
        k_invs_to_slab<<<(unsigned)((n4 + 255) / 256), 256>>>(d_tmp, h->d_cg, h->nel_s, h->te_s);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        cudaFree(d_tmp);
    }
    return 0;
}

int axb_set_source(axb_handle h, int32_t fluid_src, int32_t nelsrc, const int32_t *ielsrc,
                   const float *source_term, const float *stf, int32_t niter) {
    if (use(h)) return 1;
    if (nelsrc < 0 || nelsrc > 8) return fail("nelsrc must be in 0..8");
    h->fluid_src = fluid_src; h->nelsrc = nelsrc;
    for (int k = 0; k < 8; k++) h->ielsrc[k] = (k < nelsrc) ? ielsrc[k] : 0;
    h->src_emin = 0x7fffffff; h->src_emax = -1;
    for (int k = 0; k < nelsrc; k++) {
        h->src_emin = std::min(h->src_emin, ielsrc[k] - 1);
        h->src_emax = std::max(h->src_emax, ielsrc[k] - 1);
    }
    for (int k = 0; k < nelsrc; k++)
        if (ielsrc[k] < 1 || ielsrc[k] > (fluid_src ? h->nel_f : h->nel_s)) return fail("ielsrc out of range");
    UP(h->d_src_term, source_term, (size_t)NPT * 8 * (fluid_src ? 1 : 3));
    UP(h->d_stf, stf, niter);
    h->niter_stf = niter;
    return 0;
}

int axb_set_stf_params(axb_handle h, int32_t stf_type, double decay, double t_0,
                       double shift_fact, double magnitude) {
    if (stf_type < AXB_STF_GAUSS_0 || stf_type > AXB_STF_QUHEAVI)
        return fail("axb_set_stf_params: source time function non existant (compute_stf_t, source.f90:206-233: "
                    "gauss_0, gauss_1, gauss_2, errorf, dirac_0, quheavi)");
    if (!(t_0 > 0)) return fail("axb_set_stf_params: t_0 must be positive");
    h->stf_type = stf_type; h->decay = decay; h->t_0 = t_0; h->shift = shift_fact;
    h->magnitude = magnitude;
    return 0;
}

int axb_set_receivers(axb_handle h, int32_t num_rec, const int32_t *recfile_el) {
    if (use(h)) return 1;
    for (int r = 0; r < num_rec; r++) {
        const int iel = recfile_el[r], ip = recfile_el[r + num_rec], jp = recfile_el[r + 2 * num_rec];
        if (iel < 1 || iel > h->nel_s || ip < 0 || ip > 4 || jp < 0 || jp > 4)
            return fail("recfile_el out of range");
    }
    h->num_rec = num_rec;
    UP(h->d_recfile, (const int *)recfile_el, (size_t)3 * num_rec);
    return 0;
}

int axb_set_kwf(axb_handle h, const int32_t *kwf_mask, const int32_t *mapping_ijel_ikwf,
                int32_t npoint_solid_kwf, int32_t npoint_fluid_kwf, const float *inv_rho_fluid,
                const float *DsDeta_over_J_flu, const float *DzDeta_over_J_flu,
                const float *DsDxi_over_J_flu, const float *DzDxi_over_J_flu) {
    if (use(h)) return 1;
    const size_t n = (size_t)NPT * (h->nel_s + h->nel_f), nf = (size_t)NPT * h->nel_f;
    const int npts = npoint_solid_kwf + npoint_fluid_kwf;
    for (size_t p = 0; p < n; p++)
        if (kwf_mask[p] && (mapping_ijel_ikwf[p] < 1 || mapping_ijel_ikwf[p] > npts))
            return fail("mapping_ijel_ikwf out of range");
    h->have_kwf = true; h->npt_s_kwf = npoint_solid_kwf; h->npt_f_kwf = npoint_fluid_kwf;
    UP(h->d_kwf_mask, (const int *)kwf_mask, n); UP(h->d_kwf_map, (const int *)mapping_ijel_ikwf, n);
    UP(h->d_inv_rho, inv_rho_fluid, nf);
    UP(h->d_Dse_f, DsDeta_over_J_flu, nf); UP(h->d_Dze_f, DzDeta_over_J_flu, nf);
    UP(h->d_Dsx_f, DsDxi_over_J_flu, nf); UP(h->d_Dzx_f, DzDxi_over_J_flu, nf);
    if (h->nel_f && (!h->d_inv_rho || !h->d_Dse_f || !h->d_Dze_f || !h->d_Dsx_f || !h->d_Dzx_f))
        return fail("axb_set_kwf: NULL fluid plane");
    return 0;
}

int axb_set_dump(axb_handle h, int32_t dump_type, int32_t ibeg, int32_t iend, int32_t jbeg, int32_t jend,
                 const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                 const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol,
                 const float *inv_s_solid, const float *inv_s_fluid) {
    if (use(h)) return 1;
    if (dump_type < AXB_DUMP_DISPL_ONLY || dump_type > AXB_DUMP_FULLFIELDS) return fail("unknown dump_type");
    if (ibeg < 0 || iend > 4 || ibeg > iend || jbeg < 0 || jend > 4 || jbeg > jend) return fail("bad ibeg..jend");
    h->dump_type = dump_type; h->ibeg = ibeg; h->iend = iend; h->jbeg = jbeg; h->jend = jend;
    if (dump_type == AXB_DUMP_DISPL_ONLY) return 0;
    if (!DsDeta_over_J_sol || !DzDeta_over_J_sol || !DsDxi_over_J_sol || !DzDxi_over_J_sol || !inv_s_solid ||
        (h->nel_f > 0 && !inv_s_fluid))
        return fail("axb_set_dump: NULL plane");
    const size_t n = (size_t)NPT * h->nel_s, nf = (size_t)NPT * h->nel_f;
    UP(h->dDse, DsDeta_over_J_sol, n); UP(h->dDze, DzDeta_over_J_sol, n);
    UP(h->dDsx, DsDxi_over_J_sol, n); UP(h->dDzx, DzDxi_over_J_sol, n);
    UP(h->d_inv_s_dump, inv_s_solid, n); UP(h->d_inv_s_f, inv_s_fluid, nf);
    return 0;
}
// xdmf snapshots: the host's plot-point maps (dump_xdmf_grid, meshes_io.F90:110-437) become one
// entry per element-local point, visited as xdmf_mapping does (wavefields_io.f90:690-738: the
// corners of the plot cells, so nothing is plotted with fewer than two rows or columns)
int axb_set_xdmf(axb_handle h, int32_t snap_it, int32_t i_n_xdmf, int32_t j_n_xdmf,
                 const int32_t *i_arr_xdmf, const int32_t *j_arr_xdmf,
                 const int32_t *plotting_mask, const int32_t *mapping_ijel_iplot, int32_t npoint_plot,
                 const float *DsDeta_over_J_sol, const float *DzDeta_over_J_sol,
                 const float *DsDxi_over_J_sol, const float *DzDxi_over_J_sol, const float *inv_s_solid,
                 const float *DsDeta_over_J_flu, const float *DzDeta_over_J_flu,
                 const float *DsDxi_over_J_flu, const float *DzDxi_over_J_flu, const float *inv_s_fluid,
                 const float *inv_rho_fluid) {
    if (use(h)) return 1;
    if (snap_it < 1) return fail("axb_set_xdmf: snap_it must be positive");
    if (i_n_xdmf < 1 || i_n_xdmf > NP || j_n_xdmf < 1 || j_n_xdmf > NP) return fail("axb_set_xdmf: bad i_n_xdmf / j_n_xdmf");
    for (int k = 0; k < i_n_xdmf; k++) if (i_arr_xdmf[k] < 0 || i_arr_xdmf[k] >= NP) return fail("axb_set_xdmf: i_arr_xdmf out of range");
    for (int k = 0; k < j_n_xdmf; k++) if (j_arr_xdmf[k] < 0 || j_arr_xdmf[k] >= NP) return fail("axb_set_xdmf: j_arr_xdmf out of range");
    if (!DsDeta_over_J_sol || !DzDeta_over_J_sol || !DsDxi_over_J_sol || !DzDxi_over_J_sol || !inv_s_solid)
        return fail("axb_set_xdmf: NULL solid plane");
    if (h->nel_f > 0 && (!DsDeta_over_J_flu || !DzDeta_over_J_flu || !DsDxi_over_J_flu || !DzDxi_over_J_flu ||
                         !inv_s_fluid || !inv_rho_fluid))
        return fail("axb_set_xdmf: NULL fluid plane");
    const int nelem = h->nel_s + h->nel_f;
    std::vector<int> xmap((size_t)NPT * std::max(nelem, 1), 0);
    for (int el = 0; el < nelem; el++)
        for (int j = 0; j < j_n_xdmf; j++)
            for (int i = 0; i < i_n_xdmf; i++) {
                const size_t k = i + (size_t)i_n_xdmf * (j + (size_t)j_n_xdmf * el);
                if (!plotting_mask[k] || i_n_xdmf < 2 || j_n_xdmf < 2) continue;
                if (mapping_ijel_iplot[k] < 1 || mapping_ijel_iplot[k] > npoint_plot)
                    return fail("axb_set_xdmf: mapping_ijel_iplot out of range");
                xmap[i_arr_xdmf[i] + NP * j_arr_xdmf[j] + (size_t)NPT * el] = mapping_ijel_iplot[k];
            }
    h->have_xdmf = true; h->snap_it = snap_it; h->npoint_plot = npoint_plot;
    UP(h->d_xmap, xmap.data(), xmap.size());
    const size_t n = (size_t)NPT * h->nel_s, nf = (size_t)NPT * h->nel_f;
    UP(h->xs_Dse, DsDeta_over_J_sol, n); UP(h->xs_Dze, DzDeta_over_J_sol, n);
    UP(h->xs_Dsx, DsDxi_over_J_sol, n); UP(h->xs_Dzx, DzDxi_over_J_sol, n); UP(h->xs_inv_s, inv_s_solid, n);
    if (h->nel_f > 0) {
        UP(h->xf_Dse, DsDeta_over_J_flu, nf); UP(h->xf_Dze, DzDeta_over_J_flu, nf);
        UP(h->xf_Dsx, DsDxi_over_J_flu, nf); UP(h->xf_Dzx, DzDxi_over_J_flu, nf);
        UP(h->xf_inv_s, inv_s_fluid, nf); UP(h->xf_inv_rho, inv_rho_fluid, nf);
    }
    return 0;
}
int axb_xdmf_count(axb_handle h, int32_t *nsnap) { *nsnap = h->isnap; return 0; }
int axb_fetch_xdmf(axb_handle h, int32_t first, int32_t nsnap, float *out) {
    if (use(h)) return 1;
    if (!h->have_xdmf) return fail("xdmf snapshots not enabled (axb_set_xdmf)");
    if (first < 0 || nsnap < 0 || first + nsnap > h->isnap) return fail("xdmf snapshot range");
    const size_t npts = (size_t)h->npoint_plot;
    for (int v = 0; v < 5; v++)
        CK(cudaMemcpyAsync(out + npts * (size_t)nsnap * v, h->d_xsnap + npts * (first + (size_t)h->nsnap_max * v),
                           sizeof(float) * npts * nsnap, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
static int snapshot_nvars(const axb_handle_s *h) {
    const bool mono = h->order == 0;
    if (h->dump_type == AXB_DUMP_STRAIN_ONLY) return mono ? 4 : 6;
    if (h->dump_type == AXB_DUMP_FULLFIELDS) return mono ? 6 : 9;
    return 3;
}
static size_t snapshot_npoints(const axb_handle_s *h) {
    if (h->dump_type == AXB_DUMP_FULLFIELDS)
        return (size_t)(h->iend - h->ibeg + 1) * (h->jend - h->jbeg + 1) * ((size_t)h->nel_s + h->nel_f);
    return (size_t)h->npt_s_kwf + h->npt_f_kwf;
}
int axb_snapshot_layout(axb_handle h, int32_t *npoints, int32_t *nvars) {
    *npoints = (int32_t)snapshot_npoints(h); *nvars = snapshot_nvars(h);
    return 0;
}

int axb_set_halo(axb_handle h, int32_t domain, int32_t nmsg, const int32_t *list_peer,
                 const int32_t *sizemsg, const int32_t *glocal_index_msg, int32_t maxmsg,
                 int32_t num_comm_gll, const int32_t *glob2el) {
    if (domain != AXB_DOMAIN_SOLID && domain != AXB_DOMAIN_FLUID) return fail("bad domain");
    if (nmsg > MAXMSG) return fail("too many neighbours (max 8)");
    Halo &H = h->halo[domain];
    const int nglob = domain == AXB_DOMAIN_SOLID ? h->nglob_s : h->nglob_f;
    const int nel = domain == AXB_DOMAIN_SOLID ? h->nel_s : h->nel_f;
    H.nmsg = nmsg;
    H.nc = domain == AXB_DOMAIN_SOLID ? 3 : 1;
    H.glocal.assign(nmsg, {});
    int off = 0;
    for (int m = 0; m < nmsg; m++) {
        H.peer[m] = list_peer[m]; H.size[m] = sizemsg[m]; H.offset[m] = off;
        off += sizemsg[m];
        H.glocal[m].resize(sizemsg[m]);
        for (int ip = 0; ip < sizemsg[m]; ip++) {
            const int g = glocal_index_msg[ip + (size_t)maxmsg * m];
            if (g < 1 || g > nglob) return fail("glocal_index_msg out of range");
            H.glocal[m][ip] = g;
        }
    }
    H.nslots = off;
    H.ncomm = num_comm_gll;
    H.glob2el.assign(glob2el, glob2el + (size_t)3 * num_comm_gll);
    for (int ip = 0; ip < num_comm_gll; ip++) {
        const int ipol = glob2el[ip], jpol = glob2el[ip + num_comm_gll], iel = glob2el[ip + 2 * num_comm_gll];
        if (ipol < 0 || ipol > 4 || jpol < 0 || jpol > 4 || iel < 1 || iel > nel) return fail("glob2el out of range");
    }
    return 0;
}

int axb_set_time(axb_handle h, int32_t scheme, double deltat, int32_t niter, int32_t seis_it,
                 int32_t strain_it) {
    if (scheme < 0 || scheme > AXB_SS_35O10) return fail("unknown time scheme");
    if (!(deltat > 0)) return fail("deltat must be positive");
    h->scheme = scheme; h->deltat = deltat; h->niter = niter;
    h->seis_it = seis_it > 0 ? seis_it : 1; h->strain_it = strain_it;
    h->half_dt = 0.5 * deltat;                 // parameters.F90:1124-1125
    h->half_dt_sq = 0.5 * deltat * deltat;
    return 0;
}

int axb_finalize_setup(axb_handle h) {
    if (use(h)) return 1;
    if (h->nel_s > 0 && !h->inv_mass_rho) return fail("axb_set_mass not called");
    if (h->nel_s > 0 && !h->have_solid_terms) return fail("axb_set_solid_terms not called");
    if (h->nel_f > 0 && !h->d_coef_f) return fail("axb_set_fluid_terms not called");
    if (h->nel_f > 0 && h->order != 0 && !h->have_M_w_fl) return fail("axb_set_fluid_terms: M_w_fl is required for dipole/quadrupole sources");
    if (h->nel_f > 0 && h->order != 0 && !h->M0_w_fl) return fail("axb_set_fluid_terms: M0_w_fl is required for dipole/quadrupole sources");
    const size_t ns = h->css * 3, nf = (size_t)NPT * h->nel_pad_f;
    if (ns > 0x7fffffffULL) return fail("too many solid points for 32-bit point addresses");
    if (dzeros(h, h->disp, ns) || dzeros(h, h->velo, ns) || dzeros(h, h->acc0, ns) || dzeros(h, h->acc1, ns)) return 1;
    if (dzeros(h, h->chi, nf) || dzeros(h, h->dchi, nf) || dzeros(h, h->ddchi0, nf) || dzeros(h, h->ddchi1, nf)) return 1;
    // halo slabs first (assembly tables reference slot numbers)
    for (int d = 0; d < 2; d++) {
        Halo &H = h->halo[d];
        if (H.nmsg == 0) continue;
        if (dzeros(h, H.recv, (size_t)2 * H.nc * H.nslots)) return 1;
    }
    if (build_asm(h, h->nel_s, h->nglob_s, h->igloc_s, h->halo[0], h->d_asm_cp_s, h->d_asm_grp_s)) return 1;
    if (build_asm(h, h->nel_f, h->nglob_f, h->igloc_f, h->halo[1], h->d_asm_cp_f, h->d_asm_grp_f)) return 1;
    if (build_halo_send(h, h->nel_s, h->nglob_s, h->igloc_s, h->halo[0])) return 1;
    if (build_halo_send(h, h->nel_f, h->nglob_f, h->igloc_f, h->halo[1])) return 1;
    // per-tile element metadata of the fluid kernel: axis flag, S/F boundary entry (1-based)
    // of the jpol=0 row and of the jpol=npol row
    if (h->nel_f) {
        std::vector<int> mf((size_t)h->nel_pad_f * 3, 0);
        auto at = [&](int e, int row) -> int & { return mf[((size_t)(e / TE) * 3 + row) * TE + e % TE]; };
        for (int e = 0; e < h->nel_f; e++) at(e, 0) = h->axis_f_h[e] != 0;
        for (int b = 0; b < h->nel_bdry; b++) {
            int &slot = at(h->bdry_fel_h[b] - 1, h->bdry_jf_h[b] == 0 ? 1 : 2);
            if (slot != 0) return fail("two S/F boundary entries on the same fluid edge");
            slot = b + 1;
        }
        UP(h->d_meta_f, mf.data(), mf.size());
    }
    // per-tile element metadata of the solid kernel: axis flag, a_j table rows
    const int te = h->te_s;
    std::vector<int> meta((size_t)std::max(h->nel_pad_s, te) * 3, 0);
    auto meta_at = [&](int e, int row) -> int & { return meta[((size_t)(e / te) * 3 + row) * te + e % te]; };
    for (int e = 0; e < h->nel_s; e++) meta_at(e, 0) = h->axis_s_h[e] != 0;
    if (h->anel) {
        // a_j tables per distinct Q (time_step_memvars_cg4 recomputes them whenever Q changes)
        std::map<float, int> idx_mu, idx_ka;
        std::vector<double2> tab_mu, tab_ka;
        double aj[32];
        auto push = [&](std::vector<double2> &tab) {
            for (int k = 0; k < h->n_sls; k++)
                tab.push_back(make_double2(h->ts_t_h[k] * aj[k], h->ts_tm1_h[k] * aj[k]));
        };
        for (int e = 0; e < h->nel_s; e++) {
            auto it = idx_mu.find(h->Qmu_h[e]);
            if (it == idx_mu.end()) {
                it = idx_mu.emplace(h->Qmu_h[e], (int)idx_mu.size()).first;
                a_j_of_Q(h, h->Qmu_h[e], aj);
                push(tab_mu);
            }
            meta_at(e, 1) = it->second;
            auto ik = idx_ka.find(h->Qka_h[e]);
            if (ik == idx_ka.end()) {
                ik = idx_ka.emplace(h->Qka_h[e], (int)idx_ka.size()).first;
                a_j_of_Q(h, h->Qka_h[e], aj);
                push(tab_ka);
            }
            meta_at(e, 2) = ik->second;
        }
        if (tab_mu.empty()) { tab_mu.assign(h->n_sls, make_double2(0, 0)); tab_ka.assign(h->n_sls, make_double2(0, 0)); }
        UP(h->d_c_mu_tab, tab_mu.data(), tab_mu.size());
        UP(h->d_c_ka_tab, tab_ka.data(), tab_ka.size());
        h->ntab_mu = (int)(tab_mu.size() / h->n_sls);
        h->ntab_ka = (int)(tab_ka.size() / h->n_sls);
        const size_t np = h->cg ? 4 : NPT;
        if (dzeros(h, h->memvar, np * 6 * h->n_sls * h->nel_pad_s)) return 1;
        if (dzeros(h, h->src_dev_tm1, np * 6 * h->nel_pad_s)) return 1;
        if (dzeros(h, h->src_tr_tm1, np * h->nel_pad_s)) return 1;
        if (!h->cg) {
            std::vector<int> qm(std::max(h->nel_s, 1), 0), qk(std::max(h->nel_s, 1), 0);
            for (int e = 0; e < h->nel_s; e++) { qm[e] = meta_at(e, 1); qk[e] = meta_at(e, 2); }
            UP(h->d_qidx_mu, qm.data(), qm.size());
            UP(h->d_qidx_ka, qk.data(), qk.size());
        }
    }
    UP(h->d_meta, meta.data(), meta.size());
    if (h->scheme != AXB_NEWMARK2 && h->fluid_src && h->nelsrc > 0)
        return fail("a source in the fluid is only defined for the Newmark scheme: the reference's "
                    "symplectic loop never calls add_source_fl (time_evol_wave.F90:584-739)");
    if (h->scheme != AXB_NEWMARK2) {
        if (symplectic_coefficients(h)) return 1;
        // stf at the sub-stage times of every step: subdt = t - deltat + coeff
        // (time_evol_wave.F90:592-593, :689), t accumulated as in :586
        std::vector<float> tab((size_t)h->nstages * std::max(h->niter, 1));
        double t = 0.0, subdt[40], stf_symp[40];
        for (int it = 0; it < h->niter; it++) {
            t += h->deltat;
            for (int k = 0; k < h->nstages; k++) subdt[k] = t - h->deltat + h->coeff[k];
            stf_t(h, h->nstages, subdt, stf_symp);
            for (int k = 0; k < h->nstages; k++) tab[(size_t)it * h->nstages + k] = (float)stf_symp[k];
        }
        UP(h->d_stf_symp, tab.data(), tab.size());
    } else if (h->nelsrc > 0 && (!h->d_stf || h->niter_stf < h->niter)) {
        return fail("stf(niter) missing or shorter than niter");
    }
    h->nseismo_max = h->niter / h->seis_it + 1;          // parameters.F90:929
    if (dzeros(h, h->d_recdump, (size_t)3 * std::max(h->num_rec, 1) * h->nseismo_max)) return 1;
    if (h->strain_it > 0 && h->have_kwf) {
        h->nstrain_max = h->niter / h->strain_it + 1;
        if (h->dump_type != AXB_DUMP_DISPL_ONLY && !h->dDse) return fail("axb_set_dump: planes missing");
        if (dzeros(h, h->d_snap, snapshot_npoints(h) * h->nstrain_max * snapshot_nvars(h))) return 1;
    }
    if (h->have_xdmf) {
        h->nsnap_max = h->niter / h->snap_it + 1;            // parameters.F90:946
        if (dzeros(h, h->d_xsnap, (size_t)std::max(h->npoint_plot, 1) * h->nsnap_max * 5)) return 1;
    }
    if (dzeros(h, h->d_counters, 4)) return 1;
    if (dzeros(h, h->d_dyn, 4)) return 1;
    {
        // Lean Newmark formulation: the product build's default; the bit-exact build keeps the
        // reference's statement order.  Not with a sponge (its terms need v and u at the old time
        // level) and not with the energy diagnostic (needs v after every step).  AXB_LEAN=0/1.
#ifdef AXB_STRICT
        bool lean = false;
#else
        bool lean = true;
#endif
        if (const char *ev = getenv("AXB_LEAN")) lean = atoi(ev) != 0;
        // ... and not with fullfields dumps (velo and dchi of the step are dumped)
        h->lean = lean && h->scheme == AXB_NEWMARK2 && !h->gamma_s && !h->gamma_f && !h->dump_energy &&
                  !(h->dump_type == AXB_DUMP_FULLFIELDS && h->strain_it > 0 && h->have_kwf);
        h->lean_state = false;
        bool graph = true;
        if (const char *ev = getenv("AXB_GRAPH")) graph = atoi(ev) != 0;
        h->use_graph = graph && h->scheme == AXB_NEWMARK2 && !h->dump_energy;
        if (const char *ev = getenv("AXB_HALO_TIMEOUT_MS")) h->halo_timeout_ns = (unsigned long long)std::max(1, atoi(ev)) * 1000000ull;
    }
    if (h->dump_energy && dzeros(h, h->d_energy, (size_t)4 * (h->niter + 1))) return 1;
    // persistent grids: whole multiples of the SM count
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    const int sms = prop.multiProcessorCount;
    h->sms = sms;
    h->grid_f = std::max(1, std::min(cdiv(h->nel_f, 8), sms * 8));     // k_dump_fluid
    if (h->nel_f > 0) {
        // F_A: two persistent CTAs per SM, each with its own ring
        const size_t cap = std::min<size_t>(prop.sharedMemPerBlockOptin, prop.sharedMemPerMultiprocessor / 2 - 1024);
        const size_t sb = fluid_stage_bytes(h->npl_f);
        int nst = (int)std::min<size_t>(FLUID_MAX_STAGES, (cap - FLUID_HDR_BYTES) / sb);
        if (nst < 2) return fail("not enough shared memory for the fluid tile ring");
        // test hooks: a shallow ring / a small grid make every CTA wrap its ring many times on a
        // small mesh (tests/test_gpu_scale.py)
        if (const char *ev = getenv("AXB_FLUID_STAGES")) nst = std::max(2, std::min(nst, atoi(ev)));
        h->nst_f = nst;
        h->smem_fluid = FLUID_HDR_BYTES + (size_t)nst * sb;
        h->grid_ft = std::max(1, std::min(h->nel_pad_f / TE, 2 * sms));
        if (const char *ev = getenv("AXB_FLUID_GRID")) h->grid_ft = std::max(1, std::min(h->grid_ft, atoi(ev)));
        CK(cudaFuncSetAttribute(k_fluid_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_fluid));
    }
    if (h->nel_s > 0) {
        // compiled variants: elastic, the reference default NR_LIN_SOLIDS 5, any other n_sls
        const int v = (!h->anel || !h->cg) ? 0 : (h->n_sls == 5 ? 1 : 2);
        const size_t optin = prop.sharedMemPerBlockOptin;
        {
            // S_A (tile): persistent CTAs; the ring takes all the shared memory it can get
            const SolidTileLayout Ly = solid_tile_layout(h->order, h->anel && h->cg, h->n_sls);
            const size_t cap = SOLID_CTAS_PER_SM == 1 ? optin
                             : std::min<size_t>(optin, prop.sharedMemPerMultiprocessor / SOLID_CTAS_PER_SM - 1024);
            if (cap < Ly.hdr_bytes + 2 * Ly.stage_bytes) return fail("not enough shared memory for the solid tile ring");
            int nst = (int)std::min<size_t>(MAX_STAGES, (cap - Ly.hdr_bytes) / Ly.stage_bytes);
            if (const char *ev = getenv("AXB_SOLID_STAGES")) nst = std::max(2, std::min(nst, atoi(ev)));
            h->nst = nst;
            // the a_j tables go to shared memory when that does not cost a ring stage
            size_t tab_bytes = 0;
            h->tab_smem = 0;
            if (h->anel && h->cg) {
                const size_t need = ((size_t)(h->ntab_mu + h->ntab_ka) * h->n_sls * sizeof(double2) + 127) / 128 * 128;
                const char *ev = getenv("AXB_TAB_SMEM");
                if (Ly.hdr_bytes + need + (size_t)nst * Ly.stage_bytes <= cap && !(ev && atoi(ev) == 0)) {
                    tab_bytes = need;
                    h->tab_smem = 1;
                }
            }
            h->ring_off = (int)(Ly.hdr_bytes + tab_bytes);
            h->smem_solid = Ly.hdr_bytes + tab_bytes + (size_t)nst * Ly.stage_bytes;
            h->grid_s = std::max(1, std::min(h->nel_pad_s / TES, sms * SOLID_CTAS_PER_SM));
            if (const char *ev = getenv("AXB_SOLID_GRID")) h->grid_s = std::max(1, std::min(h->grid_s, atoi(ev)));
            static solid_kernel_t const table[3][3] = {
                {k_solid_tile<0, 0>, k_solid_tile<0, 5>, k_solid_tile<0, -1>},
                {k_solid_tile<1, 0>, k_solid_tile<1, 5>, k_solid_tile<1, -1>},
                {k_solid_tile<2, 0>, k_solid_tile<2, 5>, k_solid_tile<2, -1>}};
            h->solid_kernel = table[h->order][v];
        }
        CK(cudaFuncSetAttribute(h->solid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_solid));
    }
    h->iter = h->iseismo = h->istrain = h->ienergy = 0;
    h->isnap = 0;
    h->finalized = true;
    CK(cudaDeviceSynchronize());
    // host copies no longer needed
    std::vector<int>().swap(h->igloc_s);
    std::vector<int>().swap(h->igloc_f);
    return 0;
}

// --------------------------------------------------------------------------------------
// halo wiring
static int wire(axb_handle_s *me, int d, int m, int2 *peer_recv, int peer_nslots, int peer_offset) {
    Halo &H = me->halo[d];
    H.peer_recv[m] = peer_recv;
    H.peer_nslots[m] = peer_nslots;
    H.peer_offset[m] = peer_offset;
    return 0;
}

int axb_connect_local(axb_handle *hs, int32_t n) {
    for (int a = 0; a < n; a++) {
        if (!hs[a]->finalized) return fail("connect_local before finalize_setup");
        hs[a]->group.assign(hs, hs + n);
    }
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            if (a == b || hs[a]->device == hs[b]->device) continue;
            if (use(hs[a])) return 1;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, hs[a]->device, hs[b]->device));
            if (!can) return fail("GPUs are not peer-accessible");
            cudaError_t e = cudaDeviceEnablePeerAccess(hs[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(cudaGetErrorString(e));
            cudaGetLastError();
        }
    for (int a = 0; a < n; a++)
        for (int d = 0; d < 2; d++) {
            Halo &H = hs[a]->halo[d];
            for (int m = 0; m < H.nmsg; m++) {
                axb_handle_s *p = nullptr;
                for (int b = 0; b < n; b++) if (hs[b]->rank == H.peer[m]) p = hs[b];
                if (!p) return fail("halo peer not in the group");
                Halo &Hp = p->halo[d];
                int mm = -1;
                for (int q = 0; q < Hp.nmsg; q++) if (Hp.peer[q] == hs[a]->rank) mm = q;
                if (mm < 0 || Hp.size[mm] != H.size[m]) return fail("halo lists inconsistent between ranks");
                wire(hs[a], d, m, Hp.recv, Hp.nslots, Hp.offset[mm]);
            }
        }
    return 0;
}

// IPC blob: memhandles of the solid and of the fluid receive slab, then per domain: nmsg,
// nslots, (peer, size, offset) x MAXMSG
struct IpcBlob {
    cudaIpcMemHandle_t recv[2];
    int rank, nmsg[2], nslots[2];
    int peer[2][MAXMSG], size[2][MAXMSG], offset[2][MAXMSG];
};

static_assert(sizeof(IpcBlob) <= AXB_IPC_BLOB_BYTES, "AXB_IPC_BLOB_BYTES too small");
int32_t axb_ipc_blob_bytes(void) { return AXB_IPC_BLOB_BYTES; }

int axb_ipc_export(axb_handle h, void *blob, int32_t blob_bytes) {
    if (use(h)) return 1;
    if (blob_bytes < AXB_IPC_BLOB_BYTES) return fail("ipc blob too small: AXB_IPC_BLOB_BYTES (1024) are required");
    if (!h->finalized) return fail("ipc_export before finalize_setup");
    IpcBlob b;
    std::memset(&b, 0, sizeof b);
    b.rank = h->rank;
    for (int d = 0; d < 2; d++) {
        Halo &H = h->halo[d];
        b.nmsg[d] = H.nmsg; b.nslots[d] = H.nslots;
        for (int m = 0; m < H.nmsg; m++) { b.peer[d][m] = H.peer[m]; b.size[d][m] = H.size[m]; b.offset[d][m] = H.offset[m]; }
        if (H.nmsg) CK(cudaIpcGetMemHandle(&b.recv[d], H.recv));
    }
    std::memcpy(blob, &b, sizeof b);
    return 0;
}

int axb_ipc_import(axb_handle h, int32_t peer_rank, const void *blob, int32_t blob_bytes) {
    if (use(h)) return 1;
    if (blob_bytes < AXB_IPC_BLOB_BYTES) return fail("ipc blob too small: AXB_IPC_BLOB_BYTES (1024) are required");
    IpcBlob b;
    std::memcpy(&b, blob, sizeof b);
    if (b.rank != peer_rank) return fail("ipc blob does not belong to that rank");
    for (int d = 0; d < 2; d++) {
        Halo &H = h->halo[d];
        for (int m = 0; m < H.nmsg; m++) {
            if (H.peer[m] != peer_rank) continue;
            int mm = -1;
            for (int q = 0; q < b.nmsg[d]; q++) if (b.peer[d][q] == h->rank) mm = q;
            if (mm < 0 || b.size[d][mm] != H.size[m]) return fail("halo lists inconsistent between ranks");
            void *pr = nullptr;
            CK(cudaIpcOpenMemHandle(&pr, b.recv[d], cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(pr);
            wire(h, d, m, (int2 *)pr, b.nslots[d], b.offset[d][mm]);
        }
    }
    return 0;
}

// --------------------------------------------------------------------------------------
// kernel launches
static void prof_mark(axb_handle_s *h, bool begin) {
    if (!h->prof || (h->prof == 2 && h->prof_cls != 0)) return;
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
    }
    cudaEventRecord(h->ev_pool[h->ev_used++], h->stream);
    if (begin) h->ev_cls.push_back(h->prof_cls);
}
#define LAUNCH(h, kern, grid, block, ...)                                                \
    do {                                                                                 \
        prof_mark((h), true);                                                            \
        kern<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);                         \
        prof_mark((h), false);                                                           \
        (h)->launches++;                                                                 \
    } while (0)
#define CLS(h, c) (h)->prof_cls = (c)

// dyn: the kernel is being captured into the step graph and takes the exchange number from
// the device-resident counters (value = dyn[DYN_SEQ] + 1)
static HaloArrival halo_arrival(const axb_handle_s *h, const Halo &H, bool dyn) {
    HaloArrival r;
    std::memset(&r, 0, sizeof r);
    r.nmsg = H.nmsg; r.value = dyn ? 1 : H.seq;
    for (int m = 0; m < H.nmsg; m++) r.off[m] = H.offset[m];
    r.abort = h->d_counters + 2; r.timeout_ns = h->halo_timeout_ns;
    r.dyn = dyn ? h->d_dyn : nullptr;
    return r;
}
static SolidTileArgs solid_args(axb_handle_s *h, int mode, double c0, double c1, int anel, int do_stiff) {
    SolidTileArgs a;
    std::memset(&a, 0, sizeof a);
    a.ntiles = h->nel_pad_s / h->te_s; a.mode = mode; a.do_stiff = do_stiff; a.anel = anel;
    a.nst = h->nst; a.n_sls = h->n_sls; a.dt = c0; a.half_dt_sq = c1;
    if (const char *ev = getenv("AXB_DEBUG_SOLID")) a.dbg = atoi(ev);
    a.disp = h->disp; a.velo = h->velo; a.acc0 = h->acc0; a.acc1 = h->acc1; a.cs = h->css;
    a.coef = h->d_coef; a.meta = h->d_meta;
    for (int k = 0; k < 10; k++) a.M0_w[k] = h->d_M0_w[k];
    a.cg = h->d_cg;
    a.c_mu_tab = h->d_c_mu_tab; a.c_ka_tab = h->d_c_ka_tab;
    a.tab_smem = h->tab_smem; a.ntab_mu = h->ntab_mu; a.ntab_ka = h->ntab_ka; a.ring_off = h->ring_off;
    for (int k = 0; k < 8; k++) a.exp_w[k] = k < h->n_sls ? h->exp_w_h[k] : 0.0;
    a.memvar = h->memvar; a.src_dev_tm1 = h->src_dev_tm1; a.src_tr_tm1 = h->src_tr_tm1;
    return a;
}
#define LAUNCH_SMEM(h, kern, grid, block, smem, ...)                                     \
    do {                                                                                 \
        prof_mark((h), true);                                                            \
        kern<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);                    \
        prof_mark((h), false);                                                           \
        (h)->launches++;                                                                 \
    } while (0)
static void launch_solid_element(axb_handle_s *h, const SolidTileArgs &a) {
    if (h->nel_s == 0) return;
    CLS(h, 0);
    LAUNCH_SMEM(h, h->solid_kernel, h->grid_s, SOLID_THREADS, h->smem_solid, h->G, a);
}
// COARSE_GRAINED false: the anelastic K term and/or the memory-variable update at all 25
// points, behind S_A (axb_anel_full.cuh)
static void launch_anel_full(axb_handle_s *h, int do_stiff, int do_update, int mask) {
    if (h->nel_s == 0) return;
    CLS(h, 0);
    AnelFullArgs a;
    std::memset(&a, 0, sizeof a);
    a.nel = h->nel_s; a.n_sls = h->n_sls; a.do_stiff = do_stiff; a.do_update = do_update; a.mask = mask;
    a.disp = h->disp; a.acc1 = h->acc1; a.cs = h->css; a.axis = h->d_axis_s;
    a.Y = h->f_Y; a.Vse = h->f_Vse; a.Vsx = h->f_Vsx; a.Vze = h->f_Vze; a.Vzx = h->f_Vzx;
    a.Y0 = h->f_Y0; a.V0se = h->f_V0se; a.V0sx = h->f_V0sx; a.V0ze = h->f_V0ze; a.V0zx = h->f_V0zx;
    a.Dse = h->f_Dse; a.Dze = h->f_Dze; a.Dsx = h->f_Dsx; a.Dzx = h->f_Dzx; a.inv_s = h->d_inv_s;
    a.dmu = h->f_dmu; a.dka = h->f_dka; a.qidx_mu = h->d_qidx_mu; a.qidx_ka = h->d_qidx_ka;
    a.c_mu_tab = h->d_c_mu_tab; a.c_ka_tab = h->d_c_ka_tab;
    for (int k = 0; k < 8; k++) a.exp_w[k] = k < h->n_sls ? h->exp_w_h[k] : 0.0;
    a.memvar = h->memvar; a.src_dev_tm1 = h->src_dev_tm1; a.src_tr_tm1 = h->src_tr_tm1;
    const int grid = cdiv(h->nel_s, TEA);
    if (h->n_sls == 5) {
        if (h->order == 0) LAUNCH(h, (k_anel_full<0, 5>), grid, ANEL_THREADS, h->G, a);
        else if (h->order == 1) LAUNCH(h, (k_anel_full<1, 5>), grid, ANEL_THREADS, h->G, a);
        else LAUNCH(h, (k_anel_full<2, 5>), grid, ANEL_THREADS, h->G, a);
    } else {
        if (h->order == 0) LAUNCH(h, (k_anel_full<0, 0>), grid, ANEL_THREADS, h->G, a);
        else if (h->order == 1) LAUNCH(h, (k_anel_full<1, 0>), grid, ANEL_THREADS, h->G, a);
        else LAUNCH(h, (k_anel_full<2, 0>), grid, ANEL_THREADS, h->G, a);
    }
}
// S_A of one (sub)step: predictor/drift + K u, with the anelastic work fused (coarse-grained)
// or in the kernel behind it (all 25 points).  anel: 0 none, 1 K term only, 2 K term + update
static void launch_solid_step(axb_handle_s *h, int mode, double c0, double c1, int anel) {
    if (!h->anel) anel = 0;
    if (h->anel && !h->cg) {
        launch_solid_element(h, solid_args(h, mode, c0, c1, 0, 1));
        launch_anel_full(h, 1, anel == 2, 1);
    } else {
        launch_solid_element(h, solid_args(h, mode, c0, c1, anel, 1));
    }
}
static void launch_fluid_element(axb_handle_s *h, int mode, double c0, double c1, int full, int use_mask, int energy = 0,
                                 bool dyn = false) {
    if (h->nel_f == 0) return;
    CLS(h, 1);
    FluidTileArgs a;
    std::memset(&a, 0, sizeof a);
    a.ntiles = h->nel_pad_f / TE; a.mode = mode; a.order = h->order; a.full = full;
    a.npl = h->npl_f; a.mask_plane = h->mask_plane_f; a.nst = h->nst_f; a.dt = c0; a.half_dt_sq = c1;
    a.chi = energy ? h->dchi : h->chi; a.ddchi1 = h->ddchi1; a.dchi = h->dchi; a.ddchi0 = h->ddchi0;
    a.emask = energy;
    a.coef = h->d_coef_f; a.meta = h->d_meta_f; a.M0_w_fl = h->M0_w_fl;
    a.bdry_sel = h->d_bsel; a.bdry_js = h->d_bjs; a.bdry_matr = h->d_bmatr; a.nel_bdry = h->nel_bdry;
    a.disp = h->disp; a.cs_solid = h->css;
    a.nelsrc = h->fluid_src ? h->nelsrc : 0;
    for (int k = 0; k < 8; k++) a.ielsrc[k] = h->ielsrc[k];
    a.src_term = h->d_src_term; a.stf = h->d_stf; a.iter = h->iter; a.use_mask = use_mask;
    a.dyn = dyn ? h->d_dyn : nullptr;
    LAUNCH_SMEM(h, k_fluid_tile, h->grid_ft, FLUID_THREADS, h->smem_fluid, h->G, a);
}
static void launch_fluid_corr(axb_handle_s *h, int mode, double c, int assemble_only, bool dyn = false) {
    if (h->nel_f == 0) return;
    CLS(h, 2);
    FluidCorrArgs a;
    a.npts = NPT * h->nel_f; a.mode = mode; a.half_dt = c;
    a.ddchi1 = h->ddchi1; a.ddchi0 = h->ddchi0; a.dchi = h->dchi; a.chi = h->chi;
    a.inv_mass_fluid = h->inv_mass_fluid; a.gamma = h->gamma_f;
    a.T.cp = h->d_asm_cp_f; a.T.grp = h->d_asm_grp_f;
    Halo &H = h->halo[1];
    a.recv = H.recv; a.recv_parity = (H.seq + 1) & 1;
    a.recv_cs = H.nslots; a.assemble_only = assemble_only;
    a.arrival = halo_arrival(h, H, dyn);
    LAUNCH(h, k_fluid_corrector, cdiv(a.npts, 256), 256, a);
}
static void launch_solid_corr(axb_handle_s *h, int mode, double c, int stf_stride, int stf_off, int assemble_only,
                              bool dyn = false) {
    if (h->nel_s == 0) return;
    CLS(h, 4);
    SolidCorrArgs a;
    a.npts = NPT * h->nel_s; a.cs = h->css; a.order = h->order; a.mode = mode; a.half_dt = c;
    a.acc1 = h->acc1; a.acc0 = h->acc0; a.velo = h->velo; a.disp = h->disp;
    a.inv_mass_rho = h->inv_mass_rho; a.gamma = h->gamma_s;
    a.T.cp = h->d_asm_cp_s; a.T.grp = h->d_asm_grp_s;
    Halo &H = h->halo[0];
    a.recv = H.recv; a.recv_parity = (H.seq + 1) & 1;
    a.recv_cs = H.nslots;
    a.arrival = halo_arrival(h, H, dyn);
    a.dyn = dyn ? h->d_dyn : nullptr;
    {
        // one wave of resident blocks ahead
        static const int ahead_env = [] { const char *e = getenv("AXB_CORR_AHEAD"); return e ? atoi(e) : -1; }();
        a.ahead = ahead_env >= 0 ? ahead_env : h->sms * AXB_CORR_MINB * 256;
    }
    a.nelsrc = h->fluid_src ? 0 : h->nelsrc;
    a.src_emin = h->src_emin; a.src_emax = h->src_emax;
    for (int k = 0; k < 8; k++) a.ielsrc[k] = h->ielsrc[k];
    a.src_term = h->d_src_term;
    a.stf = (mode != 1) ? h->d_stf : h->d_stf_symp;
    a.iter = h->iter; a.stf_stride = stf_stride; a.stf_off = stf_off;
    a.assemble_only = assemble_only;
    const int grid = cdiv(a.npts, 256);
    typedef void (*corr_kernel_t)(SolidCorrArgs);
    static corr_kernel_t const table[3][5] = {
        {k_solid_corrector<0, 0>, k_solid_corrector<0, 1>, k_solid_corrector<0, 2>, k_solid_corrector<0, 3>, k_solid_corrector<0, 4>},
        {k_solid_corrector<1, 0>, k_solid_corrector<1, 1>, k_solid_corrector<1, 2>, k_solid_corrector<1, 3>, k_solid_corrector<1, 4>},
        {k_solid_corrector<2, 0>, k_solid_corrector<2, 1>, k_solid_corrector<2, 2>, k_solid_corrector<2, 3>, k_solid_corrector<2, 4>}};
    LAUNCH(h, table[h->order][assemble_only ? 4 : mode], grid, 256, a);
}
static void launch_bdry2solid(axb_handle_s *h) {
    if (h->nel_bdry == 0) return;
    CLS(h, 3);
    BdrySolidArgs a;
    a.nel_bdry = h->nel_bdry; a.order = h->order; a.bdry_sel = h->d_bsel; a.bdry_fel = h->d_bfel;
    a.bdry_js = h->d_bjs; a.bdry_jf = h->d_bjf; a.bdry_matr = h->d_bmatr; a.axis_solid = h->d_axis_s;
    a.uflu = h->ddchi0; a.acc1 = h->acc1; a.cs = h->css;
    LAUNCH(h, k_bdry2solid, cdiv(h->nel_bdry * NP, 128), 128, a);
}
// phase 1 of pdistsum_*: pack the partial sums of the cut points into the neighbours' slabs
static int halo_send(axb_handle_s *h, int d, const float *vec, size_t cs, bool dyn = false) {
    Halo &H = h->halo[d];
    if (H.nmsg == 0) return 0;
    CLS(h, 5);
    const int parity = H.seq & 1;
    PackArgs a;
    std::memset(&a, 0, sizeof a);
    a.nentries = H.nentries; a.nc = H.nc; a.start = H.d_start; a.addr = H.d_addr; a.vec = vec; a.cs = cs;
    a.dst_msg = H.d_dst_msg; a.dst_slot = H.d_dst_slot;
    a.value = H.seq + 1;
    for (int m = 0; m < H.nmsg; m++) {
        if (!H.peer_recv[m]) return fail("halo peers not connected (axb_connect_local / axb_ipc_import)");
        a.dst_base[m] = H.peer_recv[m] + H.peer_offset[m];
        a.dst_cs[m] = H.peer_nslots[m];
    }
    a.parity = parity;
    a.dyn = dyn ? h->d_dyn : nullptr;
    // every word carries its own arrival flag: no fence, no signal launch
    LAUNCH(h, k_halo_pack, std::max(1, cdiv((long long)a.nentries * a.nc, 128)), 128, a);
    if (!dyn) H.seq++;
    return 0;
}
// phase 2 (WAITALL) has no launch of its own: the threads of cut points inside the correctors
// spin on the words they need (halo_read), the rest of the kernel overlaps the exchange.
static void halo_wait(axb_handle_s *, int) {}
// energy (time_evol_wave.F90:1424-1526) of the state after step `iter`.  K u and K dchi go
// to acc1 / ddchi1, which are dead between steps.
static void launch_energy(axb_handle_s *h) {
    if (!h->dump_energy || h->iter > h->niter || h->iter != h->ienergy) return;
    h->ienergy++;
    CLS(h, 6);
    double *out = h->d_energy + (size_t)4 * h->iter;
    if (h->nel_s > 0) {
        SolidTileArgs a = solid_args(h, 2, 0.0, 0.0, 0, 1);
        a.emask = 1;
        launch_solid_element(h, a);
        CLS(h, 6);
        EnergyArgs e;
        e.npts = NPT * h->nel_s; e.cs = h->css; e.order = h->order;
        e.stiff = h->acc1; e.u = h->disp; e.v = h->velo; e.mass = h->um_rho_s; e.axis = h->d_axis_s; e.out = out;
        LAUNCH(h, k_energy_solid, cdiv(e.npts, 256), 256, e);
    }
    if (h->nel_f > 0) {
        launch_fluid_element(h, 2, 0.0, 0.0, 0, 0, 1);
        CLS(h, 6);
        EnergyArgs e;
        e.npts = NPT * h->nel_f; e.cs = 0; e.order = h->order;
        e.stiff = h->ddchi1; e.u = h->dchi; e.v = h->ddchi0; e.mass = h->um_lam_f; e.axis = h->d_axis_f; e.out = out + 2;
        LAUNCH(h, k_energy_fluid, cdiv(e.npts, 256), 256, e);
    }
    h->acc1_is_acc0 = true;
}
static RecArgs rec_args(axb_handle_s *h) {
    RecArgs a;
    a.num_rec = h->num_rec; a.order = h->order; a.seis_it = h->seis_it; a.nseismo_max = h->nseismo_max;
    a.recfile_el = h->d_recfile; a.disp = h->disp; a.cs = h->css;
    a.recdump = h->d_recdump; a.iter = h->iter; a.iseismo = h->iseismo; a.dyn = nullptr;
    return a;
}
static bool receivers_due(const axb_handle_s *h) {
    return h->num_rec > 0 && h->iter % h->seis_it == 0 && h->iseismo < h->nseismo_max;
}
static void launch_wavefield_dump(axb_handle_s *h);
static void launch_dumps(axb_handle_s *h) {
    // dump_stuff (time_evol_wave.F90:1104-1251): receivers every seis_it, wavefield every strain_it
    launch_energy(h);
    CLS(h, 6);
    if (receivers_due(h)) {
        RecArgs a = rec_args(h);
        LAUNCH(h, k_sample_receivers, cdiv(h->num_rec, 128), 128, a);
        h->iseismo++;
    }
    launch_wavefield_dump(h);
}
static void launch_wavefield_dump(axb_handle_s *h) {
    CLS(h, 6);
    if (h->have_xdmf && h->iter % h->snap_it == 0 && h->isnap < h->nsnap_max) {
        // glob_snapshot_xdmf (time_evol_wave.F90:1167-1176)
        FieldDumpArgs a;
        std::memset(&a, 0, sizeof a);
        a.nel_s = h->nel_s; a.nel_f = h->nel_f; a.order = h->order; a.dump_type = 3;
        a.xmap = h->d_xmap; a.ibeg = 0; a.iend = 4; a.jbeg = 0; a.jend = 4;
        a.nstrain_max = h->nsnap_max; a.istrain = h->isnap;
        a.axis_s = h->d_axis_s; a.axis_f = h->d_axis_f;
        a.disp = h->disp; a.velo = h->velo; a.cs = h->css; a.chi = h->chi; a.dchi = h->dchi;
        a.Dse = h->xs_Dse; a.Dze = h->xs_Dze; a.Dsx = h->xs_Dsx; a.Dzx = h->xs_Dzx; a.inv_s = h->xs_inv_s;
        a.Dse_f = h->xf_Dse; a.Dze_f = h->xf_Dze; a.Dsx_f = h->xf_Dsx; a.Dzx_f = h->xf_Dzx;
        a.inv_s_f = h->xf_inv_s; a.inv_rho = h->xf_inv_rho;
        a.snap = h->d_xsnap; a.npts = (size_t)h->npoint_plot;
        if (h->nel_s) LAUNCH(h, k_dump_fields_solid, std::max(1, std::min(cdiv(h->nel_s, 8), h->sms * 8)), 256, h->G, a);
        if (h->nel_f) LAUNCH(h, k_dump_fields_fluid, h->grid_f, 256, h->G, a);
        h->isnap++;
    }
    if (h->have_kwf && h->strain_it > 0 && h->iter % h->strain_it == 0 && h->istrain < h->nstrain_max &&
        h->dump_type != AXB_DUMP_DISPL_ONLY) {
        // compute_strain [+ dump_velo_global] (axb_dump_fields.cuh)
        FieldDumpArgs a;
        std::memset(&a, 0, sizeof a);
        a.nel_s = h->nel_s; a.nel_f = h->nel_f; a.order = h->order; a.dump_type = h->dump_type;
        a.ibeg = h->ibeg; a.iend = h->iend; a.jbeg = h->jbeg; a.jend = h->jend;
        a.nstrain_max = h->nstrain_max; a.istrain = h->istrain;
        a.kwf_mask = h->d_kwf_mask; a.kwf_map = h->d_kwf_map; a.axis_s = h->d_axis_s; a.axis_f = h->d_axis_f;
        a.disp = h->disp; a.velo = h->velo; a.cs = h->css; a.chi = h->chi; a.dchi = h->dchi;
        a.Dse = h->dDse; a.Dze = h->dDze; a.Dsx = h->dDsx; a.Dzx = h->dDzx; a.inv_s = h->d_inv_s_dump;
        a.Dse_f = h->d_Dse_f; a.Dze_f = h->d_Dze_f; a.Dsx_f = h->d_Dsx_f; a.Dzx_f = h->d_Dzx_f;
        a.inv_s_f = h->d_inv_s_f; a.inv_rho = h->d_inv_rho;
        a.snap = h->d_snap; a.npts = snapshot_npoints(h);
        if (h->nel_s) LAUNCH(h, k_dump_fields_solid, std::max(1, std::min(cdiv(h->nel_s, 8), h->sms * 8)), 256, h->G, a);
        if (h->nel_f) LAUNCH(h, k_dump_fields_fluid, h->grid_f, 256, h->G, a);
        h->istrain++;
        return;
    }
    if (h->have_kwf && h->strain_it > 0 && h->iter % h->strain_it == 0 && h->istrain < h->nstrain_max) {
        DumpArgs a;
        a.nel_s = h->nel_s; a.nel_f = h->nel_f; a.order = h->order; a.strain_it = h->strain_it;
        a.nstrain_max = h->nstrain_max; a.kwf_mask = h->d_kwf_mask; a.kwf_map = h->d_kwf_map;
        a.disp = h->disp; a.chi = h->chi; a.cs = h->css; a.axis_f = h->d_axis_f;
        a.inv_rho = h->d_inv_rho; a.Dse = h->d_Dse_f; a.Dze = h->d_Dze_f; a.Dsx = h->d_Dsx_f; a.Dzx = h->d_Dzx_f;
        a.snap = h->d_snap; a.npts = (size_t)h->npt_s_kwf + h->npt_f_kwf; a.iter = h->iter; a.istrain = h->istrain;
        if (h->nel_s) LAUNCH(h, k_dump_solid, cdiv((long long)NPT * h->nel_s, 256), 256, a);
        if (h->nel_f) LAUNCH(h, k_dump_fluid, h->grid_f, 256, h->G, a);
        h->istrain++;
    }
}

// classic <-> lean Newmark state (see axb_handle_s::lean)
static int lean_steps_since(axb_handle_s *h) { return h->iter - h->lean_entry_iter; }
static void to_lean(axb_handle_s *h) {
    if (!h->lean || h->lean_state) return;
    CLS(h, 7);
    const int ns = (int)(3 * h->css), nf = NPT * h->nel_f;
    if (h->nel_s) LAUNCH(h, k_axpy, cdiv(ns, 256), 256, ns, h->velo, h->acc0, h->half_dt);
    if (nf) LAUNCH(h, k_axpy, cdiv(nf, 256), 256, nf, h->dchi, h->ddchi0, h->half_dt);
    h->lean_state = true;
    h->lean_entry_iter = h->iter;
}
static void to_classic(axb_handle_s *h) {
    if (!h->lean_state) return;
    CLS(h, 7);
    const int ns = (int)(3 * h->css), nf = NPT * h->nel_f;
    if (h->nel_s) {
        if (lean_steps_since(h) == 0) {
            LAUNCH(h, k_axpy, cdiv(ns, 256), 256, ns, h->velo, h->acc0, -h->half_dt);
        } else {
            // a of the last step from the still intact acc1 (same assembly, source and mass
            // terms as the corrector that ran), then v = (v + dt/2 a) - dt/2 a
            h->iter--;
            launch_solid_corr(h, 3, h->half_dt, 1, 0, 0, false);
            h->iter++;
        }
    }
    if (nf) LAUNCH(h, k_axpy, cdiv(nf, 256), 256, nf, h->dchi, h->ddchi0, -h->half_dt);
    h->lean_state = false;
}

// One Newmark step, split at the two exchange points so that in-process groups can be
// enqueued rank by rank (every send is enqueued before the matching wait of any rank).
// dyn: the launches are being captured into the step graph (device-resident counters).
static int newmark_a(axb_handle_s *h, bool dyn = false) {
    // S_A first: the fluid needs the *predicted* solid displacement on the S/F boundary
    const int mode = h->lean ? 3 : 0;
    launch_solid_step(h, mode, h->deltat, h->half_dt_sq, 2);
    launch_fluid_element(h, mode, h->deltat, h->half_dt_sq, 1, 1, 0, dyn);
    if (halo_send(h, 1, h->ddchi1, (size_t)NPT * h->nel_f, dyn)) return 1;
    return 0;
}
static int newmark_b(axb_handle_s *h, bool dyn = false) {
    halo_wait(h, 1);
    if (h->lean) launch_fluid_corr(h, 2, h->deltat, 0, dyn);
    else launch_fluid_corr(h, 0, h->half_dt, 0, dyn);
    launch_bdry2solid(h);
    if (halo_send(h, 0, h->acc1, h->css, dyn)) return 1;
    return 0;
}
// blow-up guard of runtime_info (time_evol_wave.F90:1042-1054), every 100 steps on the device
static void launch_runtime_info(axb_handle_s *h) {
    if (h->nel_s == 0 || h->magnitude == 0.0 || h->iter % 100 != 0) return;
    CLS(h, 7);
    LAUNCH(h, k_blowup_check, cdiv(h->nel_s, 256), 256, h->disp, h->css, h->nel_s,
           (float)(10.0 * std::fabs(h->magnitude)), h->iter, h->d_counters);
}
static int check_blowup(axb_handle_s *h) {
    int c[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (c[2] != 0) {
        // the reference's pcheck stops every rank when one fails (commpi.F90:64-111)
        return fail("HALO EXCHANGE TIMED OUT on rank " + std::to_string(h->rank) + ": message " +
                    std::to_string(c[2] - 1) + " of a neighbour never arrived (peer not stepped, or dead); "
                    "the state of this rank is no longer valid");
    }
    if (c[3] != 0) return fail("DISPLACEMENTS BLEW UP: |disp(1,1,:,:)| > 10 |magnitude| at or before time step " + std::to_string(c[3]));
    return 0;
}
static int newmark_c(axb_handle_s *h, bool dyn = false) {
    halo_wait(h, 0);
    if (h->lean) launch_solid_corr(h, 2, h->deltat, 1, 0, 0, dyn);
    else launch_solid_corr(h, 0, h->half_dt, 1, 0, 0, dyn);
    if (dyn) {
        // last node of the graph: receiver sampling + the device counters move on
        CLS(h, 6);
        RecArgs a = rec_args(h);
        a.dyn = h->d_dyn;
        LAUNCH(h, k_step_end, 1, 256, a);
        return 0;
    }
    h->iter++;
    launch_runtime_info(h);
    launch_dumps(h);
    return 0;
}
// The Newmark step as one CUDA graph (S_A, F_A, pack, F_B, coupling, pack, S_B, step end): one
// launch per step from the host, kernel-to-kernel hand-over on the device.
static int ensure_step_graph(axb_handle_s *h) {
    if (h->step_graph || !h->use_graph) return 0;
    const int64_t l0 = h->launches;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        h->use_graph = false;           // e.g. the legacy default stream: keep launching directly
        return 0;
    }
    int rc = newmark_a(h, true) || newmark_b(h, true) || newmark_c(h, true);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->step_graph_nodes = (int)(h->launches - l0);
    h->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (e != cudaSuccess || !g) { cudaGetLastError(); h->use_graph = false; return 0; }
    e = cudaGraphInstantiate(&h->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); h->step_graph = nullptr; h->use_graph = false; }
    return 0;
}
static int newmark_graph_step(axb_handle_s *h) {
    if (!h->dyn_synced) {
        CLS(h, 7);
        LAUNCH(h, k_set_dyn, 1, 1, h->d_dyn, h->iter, std::max(h->halo[0].seq, h->halo[1].seq), h->iseismo);
        h->dyn_synced = true;
    }
    CK(cudaGraphLaunch(h->step_graph, h->stream));
    h->launches += h->step_graph_nodes;
    // host mirrors of what the graph did on the device
    for (int d = 0; d < 2; d++) if (h->halo[d].nmsg) h->halo[d].seq++;
    h->iter++;
    if (receivers_due(h)) h->iseismo++;
    launch_runtime_info(h);
    launch_wavefield_dump(h);
    return 0;
}
static int symp_a(axb_handle_s *h, int k) {
    launch_solid_step(h, 1, h->coefd[k], 0.0, 1);
    launch_fluid_element(h, 1, h->coefd[k], 0.0, 1, 0);
    return halo_send(h, 1, h->ddchi1, (size_t)NPT * h->nel_f);
}
static int symp_b(axb_handle_s *h, int k) {
    halo_wait(h, 1);
    launch_fluid_corr(h, 1, h->coefv[k], 0);
    launch_bdry2solid(h);
    return halo_send(h, 0, h->acc1, h->css);
}
static int symp_c(axb_handle_s *h, int k) {
    halo_wait(h, 0);
    launch_solid_corr(h, 1, h->coefv[k], h->nstages, k, 0);
    return 0;
}
static int symp_finish(axb_handle_s *h) {
    const double cd = h->coefd[h->nstages];
    const int nf = NPT * h->nel_f;
    const size_t cs = h->css;
    CLS(h, 7);
    if (nf) LAUNCH(h, k_drift, cdiv(nf, 256), 256, nf, h->chi, h->dchi, cd);
    if (h->nel_s) {
        if (h->order == 0) {
            LAUNCH(h, k_drift, cdiv(cs, 256), 256, (int)cs, h->disp, h->velo, cd);
            LAUNCH(h, k_drift, cdiv(cs, 256), 256, (int)cs, h->disp + 2 * cs, h->velo + 2 * cs, cd);
        } else {
            LAUNCH(h, k_drift, cdiv(3 * cs, 256), 256, (int)(3 * cs), h->disp, h->velo, cd);
        }
    }
    if (h->anel && h->cg) launch_solid_element(h, solid_args(h, 2, 0.0, 0.0, 3, 0));
    else if (h->anel) launch_anel_full(h, 0, 1, 0);
    h->iter++;
    launch_runtime_info(h);
    launch_dumps(h);
    return 0;
}

int axb_run_group(axb_handle *hs, int32_t n, int32_t nsteps) {
    for (int i = 0; i < n; i++) {
        if (!hs[i]->finalized) return fail("axb_run before axb_finalize_setup");
        if (hs[i]->iter + nsteps > hs[i]->niter) return fail("axb_run beyond niter");
    }
    for (int i = 0; i < n; i++) {
        if (use(hs[i])) return 1;
        if (hs[i]->iter == 0 && hs[i]->iseismo == 0 && hs[i]->istrain == 0) launch_dumps(hs[i]);
        hs[i]->acc1_is_acc0 = true;
        to_lean(hs[i]);
    }
    // one handle, Newmark, no per-kernel events: the step is replayed from its CUDA graph
    const bool graph = n == 1 && hs[0]->use_graph && hs[0]->scheme == AXB_NEWMARK2 && !hs[0]->prof;
    if (graph && ensure_step_graph(hs[0])) return 1;
    for (int s = 0; s < nsteps; s++) {
        if (graph && hs[0]->step_graph) {
            if (newmark_graph_step(hs[0])) return 1;
            continue;
        }
        for (int i = 0; i < n; i++) hs[i]->dyn_synced = false;
        if (hs[0]->scheme == AXB_NEWMARK2) {
            for (int i = 0; i < n; i++) { if (use(hs[i]) || newmark_a(hs[i])) return 1; }
            for (int i = 0; i < n; i++) { if (use(hs[i]) || newmark_b(hs[i])) return 1; }
            for (int i = 0; i < n; i++) { if (use(hs[i]) || newmark_c(hs[i])) return 1; }
        } else {
            for (int k = 0; k < hs[0]->nstages; k++) {
                for (int i = 0; i < n; i++) { if (use(hs[i]) || symp_a(hs[i], k)) return 1; }
                for (int i = 0; i < n; i++) { if (use(hs[i]) || symp_b(hs[i], k)) return 1; }
                for (int i = 0; i < n; i++) { if (use(hs[i]) || symp_c(hs[i], k)) return 1; }
            }
            for (int i = 0; i < n; i++) { if (use(hs[i]) || symp_finish(hs[i])) return 1; }
        }
    }
    for (int i = 0; i < n; i++) {
        if (use(hs[i])) return 1;
        CK(cudaGetLastError());
    }
    return 0;
}

int axb_run(axb_handle h, int32_t nsteps) {
    axb_handle one[1] = {h};
    return axb_run_group(one, 1, nsteps);
}

int axb_synchronize(axb_handle h) {
    if (use(h)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (h->finalized && check_blowup(h)) return 1;
    return 0;
}

int axb_profile(axb_handle h, int32_t enable) {
    if (use(h)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    h->prof = enable < 0 ? 0 : (enable > 2 ? 1 : enable);
    h->ev_used = 0; h->ev_cls.clear();
    for (int i = 0; i < 8; i++) { h->prof_ms[i] = 0.0; h->prof_n[i] = 0; }
    return 0;
}
int axb_get_profile(axb_handle h, double *ms, int64_t *launches) {
    if (use(h)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    for (size_t k = 0; k < h->ev_cls.size(); k++) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev_pool[2 * k], h->ev_pool[2 * k + 1]));
        h->prof_ms[h->ev_cls[k]] += t;
        h->prof_n[h->ev_cls[k]] += 1;
    }
    for (int i = 0; i < 8; i++) { ms[i] = h->prof_ms[i]; launches[i] = h->prof_n[i]; h->prof_ms[i] = 0.0; h->prof_n[i] = 0; }
    h->ev_used = 0; h->ev_cls.clear();
    return 0;
}

int axb_set_stf_values(axb_handle h, int32_t first_iter, int32_t n, const float *values) {
    if (use(h)) return 1;
    if (!h->d_stf || first_iter < 0 || n < 0 || first_iter + n > h->niter_stf) return fail("stf range");
    // the caller's buffer is borrowed for the duration of the call only: a handful of samples
    // travel as kernel arguments, anything longer is copied synchronously
    if (n <= 32) {
        if (n == 0) return 0;
        StfPatch a;
        a.first = first_iter; a.n = n;
        for (int k = 0; k < 32; k++) a.v[k] = k < n ? values[k] : 0.f;
        CLS(h, 7);
        LAUNCH(h, k_patch_stf, 1, 32, h->d_stf, a);
        return 0;
    }
    CK(cudaMemcpyAsync(h->d_stf + first_iter, values, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int axb_get_stf_symp(axb_handle h, int32_t first_iter, int32_t n, float *out) {
    if (use(h)) return 1;
    if (h->scheme == AXB_NEWMARK2 || !h->d_stf_symp) return fail("axb_get_stf_symp: symplectic schemes only, after axb_finalize_setup");
    if (first_iter < 0 || n < 0 || first_iter + n > h->niter) return fail("stf range");
    if (n == 0) return 0;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->d_stf_symp + (size_t)first_iter * h->nstages, sizeof(float) * (size_t)n * h->nstages, cudaMemcpyDeviceToHost));
    return 0;
}

int32_t axb_iter(axb_handle h) { return h->iter; }
int32_t axb_nseismo(axb_handle h) { return h->iseismo; }
int32_t axb_nstrain(axb_handle h) { return h->istrain; }
int64_t axb_gpu_launches(axb_handle h) { return h->launches; }

int axb_fetch_seismograms(axb_handle h, int32_t first, int32_t nsamples, float *out) {
    if (use(h)) return 1;
    if (first < 0 || nsamples < 0 || first + nsamples > h->iseismo) return fail("seismogram range");
    CK(cudaMemcpyAsync(out, h->d_recdump + (size_t)3 * h->num_rec * first,
                       sizeof(float) * 3 * h->num_rec * nsamples, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int axb_fetch_energy(axb_handle h, int32_t first, int32_t n, float *out) {
    if (use(h)) return 1;
    if (!h->dump_energy || !h->d_energy) return fail("energy diagnostic not enabled (axb_set_energy)");
    if (first < 0 || n < 0 || first + n > h->iter + 1) return fail("fetch_energy: range beyond the computed samples");
    std::vector<double> tmp((size_t)4 * std::max(n, 1));
    CK(cudaMemcpyAsync(tmp.data(), h->d_energy + (size_t)4 * first, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t k = 0; k < (size_t)4 * n; k++) out[k] = (float)tmp[k];
    return 0;
}
int axb_fetch_snapshots(axb_handle h, int32_t first, int32_t nsnap, float *out) {
    if (use(h)) return 1;
    const size_t npts = snapshot_npoints(h);
    if (first < 0 || nsnap < 0 || first + nsnap > h->istrain) return fail("snapshot range");
    for (int v = 0; v < snapshot_nvars(h); v++)
        CK(cudaMemcpyAsync(out + npts * (size_t)nsnap * v,
                           h->d_snap + npts * (first + (size_t)h->nstrain_max * v),
                           sizeof(float) * npts * nsnap, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Solid fields are (5,5,nel_solid,3) on the host and 3 planes of pitch css on the device.
static float *field_ptr(axb_handle_s *o, int f, size_t *n, bool *planes, bool reading) {
    const size_t ns = (size_t)NPT * o->nel_s, nf = (size_t)NPT * o->nel_f;
    *planes = false;
    switch (f) {
    case AXB_F_DISP: *n = ns; *planes = true; return o->disp;
    case AXB_F_VELO: *n = ns; *planes = true; return o->velo;
    case AXB_F_ACC0: *n = ns; *planes = true; return o->acc0;
    case AXB_F_ACC1: *n = ns; *planes = true; return (reading && o->acc1_is_acc0) ? o->acc0 : o->acc1;
    case AXB_F_CHI: *n = nf; return o->chi;
    case AXB_F_DCHI: *n = nf; return o->dchi;
    case AXB_F_DDCHI0: *n = nf; return o->ddchi0;
    case AXB_F_DDCHI1: *n = nf; return (reading && o->acc1_is_acc0) ? o->ddchi0 : o->ddchi1;
    case AXB_F_MEMVAR: if (!o->anel) return nullptr; *n = (size_t)(o->cg ? 4 : NPT) * 6 * o->n_sls * o->nel_s; return o->memvar;
    case AXB_F_SRC_DEV_TM1: if (!o->anel) return nullptr; *n = (size_t)(o->cg ? 4 : NPT) * 6 * o->nel_s; return o->src_dev_tm1;
    case AXB_F_SRC_TR_TM1: if (!o->anel) return nullptr; *n = (size_t)(o->cg ? 4 : NPT) * o->nel_s; return o->src_tr_tm1;
    }
    return nullptr;
}
int axb_get_state(axb_handle h, int32_t field, float *out) {
    if (use(h)) return 1;
    if (!h->finalized) return fail("get_state before finalize_setup");
    to_classic(h);
    size_t n = 0; bool planes = false;
    float *p = field_ptr(h, field, &n, &planes, true);
    if (!p) return fail("no such field");
    if (n == 0) return 0;
    if (planes)
        CK(cudaMemcpy2DAsync(out, n * sizeof(float), p, h->css * sizeof(float), n * sizeof(float), 3,
                             cudaMemcpyDeviceToHost, h->stream));
    else
        CK(cudaMemcpyAsync(out, p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int axb_set_state(axb_handle h, int32_t field, const float *in) {
    if (use(h)) return 1;
    if (!h->finalized) return fail("set_state before finalize_setup");
    to_classic(h);
    if (field == AXB_F_ACC1 || field == AXB_F_DDCHI1) h->acc1_is_acc0 = false;
    size_t n = 0; bool planes = false;
    float *p = field_ptr(h, field, &n, &planes, false);
    if (!p) return fail("no such field");
    if (n == 0) return 0;
    if (planes)
        CK(cudaMemcpy2DAsync(p, h->css * sizeof(float), in, n * sizeof(float), n * sizeof(float), 3,
                             cudaMemcpyHostToDevice, h->stream));
    else
        CK(cudaMemcpyAsync(p, in, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int axb_apply_op(axb_handle h, int32_t op) {
    if (use(h)) return 1;
    if (!h->finalized) return fail("apply_op before finalize_setup");
    to_classic(h);
    h->acc1_is_acc0 = false;
    const size_t css = h->css;
    switch (op) {
    case AXB_OP_SOLID_STIFFNESS: launch_solid_element(h, solid_args(h, 2, 0, 0, 0, 1)); break;
    case AXB_OP_ANEL_STIFFNESS:
        if (!h->anel) return fail("no attenuation");
        if (h->cg) launch_solid_element(h, solid_args(h, 2, 0, 0, 1, 0));
        else launch_anel_full(h, 1, 0, 0);
        break;
    case AXB_OP_FLUID_STIFFNESS: launch_fluid_element(h, 2, 0, 0, 0, 0); break;
    case AXB_OP_PDISTSUM_SOLID:
        if (h->halo[0].nmsg) return fail("apply_op(pdistsum) is single-rank only");
        launch_solid_corr(h, 0, 0, 1, 0, 1);
        CK(cudaMemcpyAsync(h->acc1, h->acc0, css * 3 * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        break;
    case AXB_OP_PDISTSUM_FLUID:
        if (h->halo[1].nmsg) return fail("apply_op(pdistsum) is single-rank only");
        launch_fluid_corr(h, 0, 0, 1);
        CK(cudaMemcpyAsync(h->ddchi1, h->ddchi0, (size_t)NPT * h->nel_f * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        break;
    case AXB_OP_MEMVARS:
        if (!h->anel) return fail("no attenuation");
        if (h->cg) launch_solid_element(h, solid_args(h, 2, 0, 0, 3, 0));
        else launch_anel_full(h, 0, 1, 0);
        break;
    case AXB_OP_BDRY2FLUID:
        if (h->nel_bdry)
            LAUNCH(h, k_bdry2fluid, cdiv(h->nel_bdry * NP, 128), 128, h->nel_bdry, h->order, h->d_bsel,
                   h->d_bfel, h->d_bjs, h->d_bjf, h->d_bmatr, h->disp, css, h->ddchi1);
        break;
    case AXB_OP_BDRY2SOLID: {
        if (!h->nel_bdry) break;
        BdrySolidArgs a;
        a.nel_bdry = h->nel_bdry; a.order = h->order; a.bdry_sel = h->d_bsel; a.bdry_fel = h->d_bfel;
        a.bdry_js = h->d_bjs; a.bdry_jf = h->d_bjf; a.bdry_matr = h->d_bmatr;
        a.axis_solid = nullptr; a.uflu = h->ddchi1; a.acc1 = h->acc1; a.cs = css;
        // stand-alone operator: no axis mask (bdry_copy2solid itself has none)
        std::vector<int> z(std::max(h->nel_s, 1), 0);
        int *dz = nullptr;
        CK(cudaMalloc((void **)&dz, z.size() * sizeof(int)));
        CK(cudaMemcpy(dz, z.data(), z.size() * sizeof(int), cudaMemcpyHostToDevice));
        a.axis_solid = dz;
        LAUNCH(h, k_bdry2solid, cdiv(h->nel_bdry * NP, 128), 128, a);
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(dz);
        break; }
    default: return fail("unknown op");
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
