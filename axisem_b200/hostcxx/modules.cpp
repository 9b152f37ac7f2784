#include "modules.hpp"

#include <cstdio>
#include <cstring>

namespace axisem {

size_t Array::count() const {
    size_t n = 1;
    for (uint64_t d : dims) n *= (size_t)d;
    return n;
}
const float *Array::f32() const {
    if (type != F32) throw SolverError("array is not real(4)");
    return reinterpret_cast<const float *>(bytes.data());
}
const double *Array::f64() const {
    if (type != F64) throw SolverError("array is not real(8)");
    return reinterpret_cast<const double *>(bytes.data());
}
const int32_t *Array::i32() const {
    if (type != I32) throw SolverError("array is not integer(4)");
    return reinterpret_cast<const int32_t *>(bytes.data());
}

namespace {
struct File {
    FILE *f;
    std::string path;
    File(const std::string &p) : f(std::fopen(p.c_str(), "rb")), path(p) {
        if (!f) throw SolverError("cannot open " + p);
    }
    ~File() { if (f) std::fclose(f); }
    void get(void *dst, size_t n) {
        if (n && std::fread(dst, 1, n, f) != n) throw SolverError(path + ": truncated");
    }
    template <class T> T val() { T v; get(&v, sizeof v); return v; }
};
}  // namespace

Modules Modules::read(const std::string &path) {
    File in(path);
    char magic[8];
    in.get(magic, 8);
    if (std::memcmp(magic, "AXBPROB1", 8) != 0) throw SolverError(path + ": not an AXBPROB1 container");
    const uint32_t nrec = in.val<uint32_t>();
    Modules m;
    for (uint32_t r = 0; r < nrec; r++) {
        const uint16_t nl = in.val<uint16_t>();
        std::string name(nl, '\0');
        in.get(&name[0], nl);
        Array a;
        const uint8_t t = in.val<uint8_t>();
        if (t > 2) throw SolverError(path + ": bad type code in record " + name);
        a.type = (Array::Type)t;
        const uint8_t nd = in.val<uint8_t>();
        a.dims.resize(nd);
        for (uint8_t k = 0; k < nd; k++) a.dims[k] = in.val<uint64_t>();
        const uint64_t nb = in.val<uint64_t>();
        const size_t esz = a.type == Array::F64 ? 8 : 4;
        if (nb != a.count() * esz) throw SolverError(path + ": size mismatch in record " + name);
        a.bytes.resize(nb);
        in.get(a.bytes.data(), nb);
        m.vars_[name] = std::move(a);
    }
    return m;
}

const Array &Modules::at(const std::string &name) const {
    auto it = vars_.find(name);
    if (it == vars_.end()) throw SolverError("module variable " + name + " is not set");
    return it->second;
}
const float *Modules::f(const std::string &name) const { return has(name) ? at(name).f32() : nullptr; }
const double *Modules::d(const std::string &name) const { return has(name) ? at(name).f64() : nullptr; }
const int32_t *Modules::i(const std::string &name) const { return has(name) ? at(name).i32() : nullptr; }
int32_t Modules::int_of(const std::string &name) const {
    const Array &a = at(name);
    if (a.count() != 1) throw SolverError(name + " is not a scalar");
    return a.i32()[0];
}
int32_t Modules::int_of(const std::string &name, int32_t dflt) const { return has(name) ? int_of(name) : dflt; }
double Modules::real_of(const std::string &name) const {
    const Array &a = at(name);
    if (a.count() != 1) throw SolverError(name + " is not a scalar");
    return a.f64()[0];
}

}  // namespace axisem
