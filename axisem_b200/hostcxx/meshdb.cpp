#include "meshdb.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>
#include <unordered_set>

namespace axisem {
namespace {

// Fortran sequential unformatted: [int32 nbytes][payload][int32 nbytes]
class Unformatted {
public:
    explicit Unformatted(const std::string &path) : f_(std::fopen(path.c_str(), "rb")), path_(path) {
        if (!f_) throw SolverError("cannot open " + path);
    }
    ~Unformatted() { if (f_) std::fclose(f_); }
    // next record, whatever its length
    std::vector<unsigned char> next(const char *what) {
        int32_t n = 0, n2 = 0;
        if (std::fread(&n, 4, 1, f_) != 1) throw SolverError(path_ + ": end of file before " + what);
        if (n < 0) throw SolverError(path_ + ": bad record marker before " + what);
        std::vector<unsigned char> b((size_t)n);
        if (n && std::fread(b.data(), 1, (size_t)n, f_) != (size_t)n) throw SolverError(path_ + ": truncated record " + what);
        if (std::fread(&n2, 4, 1, f_) != 1 || n2 != n) throw SolverError(path_ + ": record markers disagree at " + what);
        return b;
    }
    std::vector<unsigned char> next(const char *what, size_t nbytes) {
        std::vector<unsigned char> b = next(what);
        if (b.size() != nbytes)
            throw SolverError(path_ + ": record " + what + " has " + std::to_string(b.size()) + " bytes, expected " +
                              std::to_string(nbytes));
        return b;
    }
    int32_t i32(const char *what) {
        const auto b = next(what, 4);
        int32_t v;
        std::memcpy(&v, b.data(), 4);
        return v;
    }
private:
    FILE *f_;
    std::string path_;
};

Array arr(Array::Type t, std::vector<uint64_t> dims, const void *src) {
    Array a;
    a.type = t;
    a.dims = std::move(dims);
    const size_t nb = a.count() * (t == Array::F64 ? 8 : 4);
    a.bytes.resize(nb);
    if (nb) std::memcpy(a.bytes.data(), src, nb);
    return a;
}
Array scalar_i(int32_t v) { return arr(Array::I32, {}, &v); }
Array scalar_d(double v) { return arr(Array::F64, {}, &v); }

// spectral arrays come in the mesher's kind: accept real(4) or real(8), keep/convert as asked
Array real_record(const std::vector<unsigned char> &b, size_t n, Array::Type want, std::vector<uint64_t> dims,
                  const char *what) {
    if (b.size() != 4 * n && b.size() != 8 * n) throw SolverError(std::string("record ") + what + " has the wrong length");
    const bool is8 = b.size() == 8 * n;
    if ((want == Array::F64) == is8) return arr(want, std::move(dims), b.data());
    if (want == Array::F32) {
        std::vector<float> v(n);
        for (size_t k = 0; k < n; k++) { double d; std::memcpy(&d, b.data() + 8 * k, 8); v[k] = (float)d; }
        return arr(Array::F32, std::move(dims), v.data());
    }
    std::vector<double> v(n);
    for (size_t k = 0; k < n; k++) { float f; std::memcpy(&f, b.data() + 4 * k, 4); v[k] = (double)f; }
    return arr(Array::F64, std::move(dims), v.data());
}

// def_grid.f90:95-180: element-local points (element, jpol, ipol order) whose global number
// takes part in any message -> glob2el(num_comm_gll, 3) = (ipol, jpol, iel)
void build_glob2el(Modules &m, const std::string &dom, int nel, int npol, const int32_t *igloc, int nmsg,
                   const int32_t *sizemsg, const int32_t *glocal, int maxmsg) {
    std::unordered_set<int32_t> sent;
    for (int im = 0; im < nmsg; im++)
        for (int ip = 0; ip < sizemsg[im]; ip++) sent.insert(glocal[(size_t)im * maxmsg + ip]);
    std::vector<int32_t> ipol_v, jpol_v, iel_v;
    const int n1 = npol + 1;
    for (int iel = 1; iel <= nel; iel++)
        for (int jpol = 0; jpol <= npol; jpol++)
            for (int ipol = 0; ipol <= npol; ipol++) {
                const size_t ipt = (size_t)(iel - 1) * n1 * n1 + (size_t)jpol * n1 + ipol;
                if (sent.count(igloc[ipt])) { ipol_v.push_back(ipol); jpol_v.push_back(jpol); iel_v.push_back(iel); }
            }
    const size_t nc = iel_v.size();
    std::vector<int32_t> g(3 * nc);
    for (size_t k = 0; k < nc; k++) { g[k] = ipol_v[k]; g[nc + k] = jpol_v[k]; g[2 * nc + k] = iel_v[k]; }
    m.put("data_comm%num_comm_gll_" + dom, scalar_i((int32_t)nc));
    m.put("data_comm%glob2el_" + dom, arr(Array::I32, {3, nc}, g.data()));     // Fortran (ncomm,3)
}

void read_messaging(Unformatted &u, Modules &m, const std::string &dom) {
    const int32_t nmsg = u.i32(("sizerecv_" + dom).c_str());
    m.put("data_comm%sizerecv_" + dom, scalar_i(nmsg));
    if (nmsg <= 0) return;
    const auto lst = u.next(("listrecv_" + dom).c_str(), 4 * (size_t)nmsg);
    const auto siz = u.next(("sizemsgrecv_" + dom).c_str(), 4 * (size_t)nmsg);
    m.put("data_comm%listrecv_" + dom, arr(Array::I32, {(uint64_t)nmsg}, lst.data()));
    m.put("data_comm%sizemsgrecv_" + dom, arr(Array::I32, {(uint64_t)nmsg}, siz.data()));
    const int32_t *sz = reinterpret_cast<const int32_t *>(siz.data());
    const int maxmsg = *std::max_element(sz, sz + nmsg);
    // glocal_index_msg_recv(1:sizemsgrecvmax, 1:sizerecv), one record per point (pdb.f90:2340-2346)
    std::vector<int32_t> gl((size_t)maxmsg * nmsg, 0);
    for (int im = 0; im < nmsg; im++)
        for (int ip = 0; ip < sz[im]; ip++) gl[(size_t)im * maxmsg + ip] = u.i32("glocal_index_msg_recv");
    m.put("data_comm%glocal_index_msg_recv_" + dom, arr(Array::I32, {(uint64_t)nmsg, (uint64_t)maxmsg}, gl.data()));
}

}  // namespace

Modules read_meshdb(const std::string &path, int mynum) {
    Unformatted u(path);
    Modules m;
    // ---- read_mesh_basics (data_mesh.f90:195-209)
    const char *basics[13] = {"nproc_mesh", "npol", "nelem", "npoint", "nel_solid", "nel_fluid", "npoint_solid",
                              "npoint_fluid", "nglob_solid", "nglob_fluid", "nel_bdry", "ndisc", "lfbkgrdmodel"};
    int32_t v[13];
    for (int k = 0; k < 13; k++) {
        v[k] = u.i32(basics[k]);
        m.put(std::string("data_mesh%") + basics[k], scalar_i(v[k]));
    }
    const int nproc = v[0], npol = v[1], nelem = v[2], nel_solid = v[4], nel_fluid = v[5];
    const int npoint_solid = v[6], npoint_fluid = v[7], nel_bdry = v[10], ndisc = v[11], lfbkgrdmodel = v[12];
    if (npol < 1 || nelem != nel_solid + nel_fluid || v[3] != nelem * (npol + 1) * (npol + 1))
        throw SolverError(path + ": inconsistent basic mesh parameters");
    m.put("data_proc%mynum", scalar_i(mynum));
    m.put("data_proc%nproc", scalar_i(nproc));
    // ---- read_mesh_advanced (data_mesh.f90:212-303)
    const size_t n1 = (size_t)npol + 1;
    for (const char *nm : {"xi_k", "eta", "dxi", "wt", "wt_axial_k"})
        m.put(std::string("data_spec%") + nm, real_record(u.next(nm), n1, Array::F64, {n1}, nm));
    m.put("data_spec%G0", real_record(u.next("G0"), n1, Array::F32, {n1}, "G0"));
    for (const char *nm : {"G1", "G1T", "G2", "G2T"})
        m.put(std::string("data_spec%") + nm, real_record(u.next(nm), n1 * n1, Array::F32, {n1 * n1}, nm));
    const int32_t npoin = u.i32("npoin");
    m.put("data_mesh%npoin", scalar_i(npoin));
    {
        const auto s = u.next("crd_nodes(:,1)", 8 * (size_t)npoin);
        auto z = u.next("crd_nodes(:,2)", 8 * (size_t)npoin);
        double *zd = reinterpret_cast<double *>(z.data());
        for (int k = 0; k < npoin; k++) if (std::abs(zd[k]) < 1.e-8) zd[k] = 0.0;      // data_mesh.f90:256-258
        std::vector<double> crd(2 * (size_t)npoin);
        std::memcpy(crd.data(), s.data(), 8 * (size_t)npoin);
        std::memcpy(crd.data() + npoin, zd, 8 * (size_t)npoin);
        m.put("data_mesh%crd_nodes", arr(Array::F64, {2, (uint64_t)npoin}, crd.data()));   // Fortran (npoin,2)
    }
    {
        std::vector<int32_t> lnods((size_t)8 * nelem);                                    // Fortran (nelem,8)
        for (int iel = 0; iel < nelem; iel++) {
            const auto b = u.next("lnods", 32);
            for (int k = 0; k < 8; k++) std::memcpy(&lnods[(size_t)k * nelem + iel], b.data() + 4 * k, 4);
        }
        m.put("data_mesh%lnods", arr(Array::I32, {8, (uint64_t)nelem}, lnods.data()));
    }
    m.put("data_mesh%nglob", scalar_i(u.i32("nglob")));
    {
        // eltype: character(len=6)(nelem) -> codes 0 curved, 1 linear, 2 semino, 3 semiso
        const auto b = u.next("eltype", 6 * (size_t)nelem);
        std::vector<int32_t> code(nelem);
        for (int e = 0; e < nelem; e++) {
            const std::string t(reinterpret_cast<const char *>(b.data()) + 6 * e, 6);
            if (t == "curved") code[e] = 0;
            else if (t == "linear") code[e] = 1;
            else if (t == "semino") code[e] = 2;
            else if (t == "semiso") code[e] = 3;
            else throw SolverError(path + ": unknown element type '" + t + "'");
        }
        m.put("data_mesh%eltype", arr(Array::I32, {(uint64_t)nelem}, code.data()));
        const auto c = u.next("coarsing", 4 * (size_t)nelem);
        m.put("data_mesh%coarsing", arr(Array::I32, {(uint64_t)nelem}, c.data()));
    }
    m.put("data_mesh%ielsolid", arr(Array::I32, {(uint64_t)nel_solid}, u.next("ielsolid", 4 * (size_t)nel_solid).data()));
    m.put("data_mesh%ielfluid", arr(Array::I32, {(uint64_t)nel_fluid}, u.next("ielfluid", 4 * (size_t)nel_fluid).data()));
    m.put("data_mesh%igloc_solid", arr(Array::I32, {(uint64_t)npoint_solid}, u.next("igloc_solid", 4 * (size_t)npoint_solid).data()));
    m.put("data_mesh%igloc_fluid", arr(Array::I32, {(uint64_t)npoint_fluid}, u.next("igloc_fluid", 4 * (size_t)npoint_fluid).data()));
    const int32_t have_bdry = u.i32("have_bdry_elem");
    m.put("data_mesh%have_bdry_elem", scalar_i(have_bdry != 0));
    if (have_bdry)
        for (const char *nm : {"bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid", "bdry_jpol_fluid"})
            m.put(std::string("data_mesh%") + nm, arr(Array::I32, {(uint64_t)nel_bdry}, u.next(nm, 4 * (size_t)nel_bdry).data()));
    // ---- read_db (get_mesh.f90:101-383)
    {
        const auto b = u.next("pts_wavelngth,period,courant,deltat", 32);
        double d[4];
        std::memcpy(d, b.data(), 32);
        m.put("data_mesh%pts_wavelngth", scalar_d(d[0]));
        m.put("data_time%period", scalar_d(d[1]));
        m.put("data_time%courant", scalar_d(d[2]));
        m.put("data_time%deltat", scalar_d(d[3]));
    }
    {
        const auto b = u.next("bkgrdmodel", (size_t)lfbkgrdmodel);
        std::vector<int32_t> chars(b.begin(), b.end());
        m.put("data_mesh%bkgrdmodel", arr(Array::I32, {(uint64_t)lfbkgrdmodel}, chars.data()));
        u.next("override_ext_q");
    }
    int32_t have_fluid = 0;
    {
        const auto b = u.next("router,have_fluid", 12);
        double router;
        std::memcpy(&router, b.data(), 8);
        std::memcpy(&have_fluid, b.data() + 8, 4);
        m.put("data_mesh%router", scalar_d(router));
        m.put("data_mesh%have_fluid", scalar_i(have_fluid != 0));
    }
    {
        std::vector<double> discont(ndisc);
        std::vector<int32_t> solid_domain(ndisc), idom_fluid(ndisc);
        for (int k = 0; k < ndisc; k++) {
            const auto b = u.next("discont,solid_domain,idom_fluid", 16);
            std::memcpy(&discont[k], b.data(), 8);
            std::memcpy(&solid_domain[k], b.data() + 8, 4);
            std::memcpy(&idom_fluid[k], b.data() + 12, 4);
        }
        m.put("data_mesh%discont", arr(Array::F64, {(uint64_t)ndisc}, discont.data()));
        m.put("data_mesh%solid_domain", arr(Array::I32, {(uint64_t)ndisc}, solid_domain.data()));
        m.put("data_mesh%idom_fluid", arr(Array::I32, {(uint64_t)ndisc}, idom_fluid.data()));
    }
    {
        const auto b = u.next("rmin,minh_ic,maxh_ic,maxh_icb", 32);
        double d[4];
        std::memcpy(d, b.data(), 32);
        m.put("data_mesh%rmin", scalar_d(d[0]));
    }
    {
        const auto b = u.next("hmin_glob,hmax_glob", 16);
        double d[2];
        std::memcpy(d, b.data(), 16);
        m.put("data_mesh%hmin_glob", scalar_d(d[0]));
        m.put("data_mesh%hmax_glob", scalar_d(d[1]));
        u.next("min_distance_dim,min_distance_nondim", 16);
    }
    for (int k = 0; k < 2; k++) {
        u.next("char_time,globel", 12);
        u.next("char_time_rad,theta", 16);
    }
    int32_t nax[3];
    {
        const auto b = u.next("naxel,naxel_solid,naxel_fluid", 12);
        std::memcpy(nax, b.data(), 12);
    }
    m.put("data_mesh%ax_el", arr(Array::I32, {(uint64_t)nax[0]}, u.next("ax_el", 4 * (size_t)nax[0]).data()));
    const auto axs = u.next("ax_el_solid", 4 * (size_t)nax[1]);
    const auto axf = u.next("ax_el_fluid", 4 * (size_t)nax[2]);
    m.put("data_mesh%ax_el_solid", arr(Array::I32, {(uint64_t)nax[1]}, axs.data()));
    m.put("data_mesh%ax_el_fluid", arr(Array::I32, {(uint64_t)nax[2]}, axf.data()));
    // def_grid.f90:59-77: the axial flags (the mesher's ax_el_* lists name the same elements
    // the SOLVER finds from scoord(0,npol,iel) == 0)
    {
        std::vector<int32_t> as(nel_solid, 0), af(nel_fluid, 0);
        const int32_t *s = reinterpret_cast<const int32_t *>(axs.data());
        const int32_t *f = reinterpret_cast<const int32_t *>(axf.data());
        for (int k = 0; k < nax[1]; k++) {
            if (s[k] < 1 || s[k] > nel_solid) throw SolverError(path + ": ax_el_solid out of range");
            as[s[k] - 1] = 1;
        }
        for (int k = 0; k < nax[2]; k++) {
            if (f[k] < 1 || f[k] > nel_fluid) throw SolverError(path + ": ax_el_fluid out of range");
            af[f[k] - 1] = 1;
        }
        m.put("data_mesh%axis_solid", arr(Array::I32, {(uint64_t)nel_solid}, as.data()));
        m.put("data_mesh%axis_fluid", arr(Array::I32, {(uint64_t)nel_fluid}, af.data()));
    }
    read_messaging(u, m, "solid");
    if (have_fluid) read_messaging(u, m, "fluid");
    else m.put("data_comm%sizerecv_fluid", scalar_i(0));
    for (const char *dom : {"solid", "fluid"}) {
        const std::string d = dom;
        const int nmsg = m.int_of("data_comm%sizerecv_" + d);
        if (nmsg <= 0) continue;
        const Array &gl = m.at("data_comm%glocal_index_msg_recv_" + d);
        build_glob2el(m, d, d == "solid" ? nel_solid : nel_fluid, npol, m.i("data_mesh%igloc_" + d), nmsg,
                      m.i("data_comm%sizemsgrecv_" + d), gl.i32(), (int)gl.dims[1]);
    }
    return m;
}

void write_container(const Modules &m, const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw SolverError("cannot write " + path);
    std::fwrite("AXBPROB1", 1, 8, f);
    const uint32_t n = (uint32_t)m.size();
    std::fwrite(&n, 4, 1, f);
    m.for_each([&](const std::string &name, const Array &a) {
        const uint16_t nl = (uint16_t)name.size();
        std::fwrite(&nl, 2, 1, f);
        std::fwrite(name.data(), 1, nl, f);
        const uint8_t t = (uint8_t)a.type, nd = (uint8_t)a.dims.size();
        std::fwrite(&t, 1, 1, f);
        std::fwrite(&nd, 1, 1, f);
        for (uint64_t d : a.dims) std::fwrite(&d, 8, 1, f);
        const uint64_t nb = a.bytes.size();
        std::fwrite(&nb, 8, 1, f);
        if (nb) std::fwrite(a.bytes.data(), 1, nb, f);
    });
    std::fclose(f);
}

}  // namespace axisem
