// modcholesky.hpp -- header-only C++17 host facade over the C-ABI (modchol_b200.h).
//
// The reference is compiled (Rust) code and its toolchain is absent from this image, so the
// host side above the C-ABI is mirrored here in C++ with the reference's names and semantics:
//
//   reference (argmin-rs/modcholesky v0.2.0)              this header
//   struct Decomposition<L,E,P>{l,e,p}   lib.rs:109-119    modcholesky::Decomposition<L,E,P>
//   trait ModCholeskyGMW81<L,E,P>        gmw81.rs:24-32    modcholesky::ModCholeskyGMW81<L,E,P>
//   trait ModCholeskySE90<L,E,P>         se90.rs:24-32     modcholesky::ModCholeskySE90<L,E,P>
//   trait ModCholeskySE99<L,E,P>         se99.rs:21-29     modcholesky::ModCholeskySE99<L,E,P>
//   trait GershgorinCircles              gershgorin.rs:15  modcholesky::GershgorinCircles
//   impl .. for ndarray::Array2<f64>     gmw81.rs:34-41 .. modcholesky::Array2 (row-major f64)
//   (new) batched entry point                              modcholesky::Array3
//
// Error behaviour: the reference panics (assert!(self.is_square()), se99.rs:38; default
// trait bodies panic!("Not implemented!")).  Here those are exceptions: std::logic_error for
// the assertion / "Not implemented!", std::runtime_error carrying modchol_last_error() for a
// negative C-ABI status.  There is no CPU fallback.
#ifndef MODCHOLESKY_HPP
#define MODCHOLESKY_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "modchol_b200.h"

namespace modcholesky {

enum class Mode : int { Strict = MODCHOL_MODE_STRICT, Fast = MODCHOL_MODE_FAST };

// lib.rs:109-119
template <class L, class E, class P>
struct Decomposition {
    L l;
    E e;
    P p;
    static Decomposition make(L l_, E e_, P p_) { return Decomposition{std::move(l_), std::move(e_), std::move(p_)}; }
};

// gmw81.rs:24-32 / se90.rs:24-32 / se99.rs:21-29: one method each, default body "panics".
template <class L, class E, class P>
struct ModCholeskyGMW81 {
    virtual ~ModCholeskyGMW81() = default;
    virtual Decomposition<L, E, P> mod_cholesky_gmw81() const { throw std::logic_error("Not implemented!"); }
};
template <class L, class E, class P>
struct ModCholeskySE90 {
    virtual ~ModCholeskySE90() = default;
    virtual Decomposition<L, E, P> mod_cholesky_se90() const { throw std::logic_error("Not implemented!"); }
};
template <class L, class E, class P>
struct ModCholeskySE99 {
    virtual ~ModCholeskySE99() = default;
    virtual Decomposition<L, E, P> mod_cholesky_se99() const { throw std::logic_error("Not implemented!"); }
};
// gershgorin.rs:15-18
struct GershgorinCircles {
    virtual ~GershgorinCircles() = default;
    virtual std::vector<std::pair<double, double>> gershgorin_circles() const = 0;
};

inline void check(int rc)
{
    if (rc != MODCHOL_OK)
        throw std::runtime_error("modchol error " + std::to_string(rc) + ": " + modchol_last_error());
}

struct Array1 {
    std::vector<double> data;
    std::size_t len() const { return data.size(); }
    double operator[](std::size_t i) const { return data[i]; }
};
struct Array1Usize {
    std::vector<std::uint64_t> data;
    std::size_t len() const { return data.size(); }
    std::uint64_t operator[](std::size_t i) const { return data[i]; }
};

// Stand-in for ndarray::Array2<f64> in standard (row-major) layout.
class Array2 : public ModCholeskyGMW81<Array2, Array1, Array1Usize>,
               public ModCholeskySE90<Array2, Array1, Array1Usize>,
               public ModCholeskySE99<Array2, Array1, Array1Usize>,
               public GershgorinCircles {
public:
    using Decomp = Decomposition<Array2, Array1, Array1Usize>;
    std::size_t rows = 0, cols = 0;
    std::vector<double> data;
    Mode mode = Mode::Strict;

    Array2() = default;
    Array2(std::size_t r, std::size_t c) : rows(r), cols(c), data(r * c, 0.0) {}
    Array2(std::size_t r, std::size_t c, std::vector<double> d) : rows(r), cols(c), data(std::move(d)) {}
    bool is_square() const { return rows == cols; }
    double& operator()(std::size_t i, std::size_t j) { return data[i * cols + j]; }
    double operator()(std::size_t i, std::size_t j) const { return data[i * cols + j]; }

    Decomp mod_cholesky_gmw81() const override { return factor(MODCHOL_GMW81); }  // gmw81.rs:39
    Decomp mod_cholesky_se90() const override { return factor(MODCHOL_SE90); }    // se90.rs:38
    Decomp mod_cholesky_se99() const override { return factor(MODCHOL_SE99); }    // se99.rs:35

    std::vector<std::pair<double, double>> gershgorin_circles() const override  // gershgorin.rs:20
    {
        if (!is_square())
            throw std::logic_error("assertion failed: self.is_square()");
        std::vector<double> out(2 * rows);
        check(modchol_gershgorin_circles_batched_f64(1, (int32_t)rows, data.data(), out.data(),
                                                     MODCHOL_MEM_HOST, nullptr));
        std::vector<std::pair<double, double>> v;
        for (std::size_t i = 0; i < rows; ++i)
            v.emplace_back(out[2 * i], out[2 * i + 1]);
        return v;
    }

private:
    Decomp factor(int algo) const
    {
        if (!is_square())  // se90.rs:41, se99.rs:38, gmw81.rs:43
            throw std::logic_error("assertion failed: self.is_square()");
        Decomp d{Array2(rows, cols), Array1{std::vector<double>(rows)},
                 Array1Usize{std::vector<std::uint64_t>(rows)}};
        check(modchol_factor_batched_f64(algo, (int)mode, 1, (int32_t)rows, data.data(), d.l.data.data(),
                                         d.e.data.data(), d.p.data.data(), nullptr, MODCHOL_MEM_HOST,
                                         nullptr));
        return d;
    }
};

// The new batched entry point: (batch, n, n) -> Decomposition<Array3, Array2, Array2<usize>>.
class Array3 {
public:
    struct Result {
        std::vector<double> l;              // batch * n * n
        std::vector<double> e;              // batch * n
        std::vector<std::uint64_t> p;       // batch * n
        std::vector<std::int32_t> phase2_start;  // batch
    };
    std::size_t batch = 0, n = 0;
    std::vector<double> data;
    Mode mode = Mode::Strict;

    Array3(std::size_t b, std::size_t n_) : batch(b), n(n_), data(b * n_ * n_, 0.0) {}
    double& operator()(std::size_t b, std::size_t i, std::size_t j) { return data[(b * n + i) * n + j]; }

    Result mod_cholesky_gmw81_batched(int ngpus = 1) const { return factor(MODCHOL_GMW81, ngpus); }
    Result mod_cholesky_se90_batched(int ngpus = 1) const { return factor(MODCHOL_SE90, ngpus); }
    Result mod_cholesky_se99_batched(int ngpus = 1) const { return factor(MODCHOL_SE99, ngpus); }

private:
    Result factor(int algo, int ngpus) const
    {
        Result r{std::vector<double>(batch * n * n), std::vector<double>(batch * n),
                 std::vector<std::uint64_t>(batch * n), std::vector<std::int32_t>(batch)};
        if (ngpus == 1)
            check(modchol_factor_batched_f64(algo, (int)mode, (int64_t)batch, (int32_t)n, data.data(),
                                             r.l.data(), r.e.data(), r.p.data(), r.phase2_start.data(),
                                             MODCHOL_MEM_HOST, nullptr));
        else
            check(modchol_factor_batched_multi_gpu_f64(algo, (int)mode, (int64_t)batch, (int32_t)n,
                                                       data.data(), r.l.data(), r.e.data(), r.p.data(),
                                                       r.phase2_start.data(), ngpus));
        return r;
    }
};

}  // namespace modcholesky
#endif
