// modules.hpp — the Fortran module variables the SOLVER time loop works on, held by name.
//
// The reference's `time_loop` (SOLVER/time_evol_wave.F90:231-245) takes no arguments: it
// reads what `prepare_waves` left in the modules data_mesh, data_spec, data_matr,
// data_pointwise, data_source, data_time, data_comm and attenuation.  `Modules` is that
// state on the C++ side of the seam: every array is stored under "<module>%<variable>" in
// the memory order the Fortran holds it (column-major, 1-based index values), so that the
// pointers can be handed to the C ABI of include/axisem_b200.h unchanged.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace axisem {

struct Array {
    enum Type : uint8_t { F32 = 0, F64 = 1, I32 = 2 };
    Type type = F32;
    std::vector<uint64_t> dims;
    std::vector<unsigned char> bytes;

    size_t count() const;
    const float *f32() const;
    const double *f64() const;
    const int32_t *i32() const;
};

class Modules {
public:
    // read an AXBPROB1 container (axisem_b200/host/problem_bin.py)
    static Modules read(const std::string &path);

    bool has(const std::string &name) const { return vars_.count(name) != 0; }
    const Array &at(const std::string &name) const;
    // array accessors: nullptr when the variable is not allocated (as in the Fortran, where
    // e.g. M13s only exists for dipole sources, def_precomp_terms.f90:1216-1284)
    const float *f(const std::string &name) const;
    const double *d(const std::string &name) const;
    const int32_t *i(const std::string &name) const;
    // scalars
    int32_t int_of(const std::string &name) const;
    int32_t int_of(const std::string &name, int32_t dflt) const;
    double real_of(const std::string &name) const;
    void put(const std::string &name, Array a) { vars_[name] = std::move(a); }
    // take every variable of `other` that is not set here
    void merge_missing(const Modules &other) {
        for (const auto &kv : other.vars_) vars_.insert(kv);
    }
    size_t size() const { return vars_.size(); }
    void for_each(const std::function<void(const std::string &, const Array &)> &fn) const {
        for (const auto &kv : vars_) fn(kv.first, kv.second);
    }

private:
    std::map<std::string, Array> vars_;
};

struct SolverError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

}  // namespace axisem
