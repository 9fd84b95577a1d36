// test_facade.cpp -- the reference's doctest (lib.rs:27-60) and se99 3x3 test
// (se99.rs:210-233) written against the C++ facade, plus the trait-default "panic" and the
// non-square assertion.  Built by tests/test_cpp_facade.py; needs a GPU to run.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "modcholesky.hpp"

using namespace modcholesky;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

// P A P^T + P E P^T == L L^T with P[i, p[i]] = 1 (utils.rs:94-113)
static double identity_residual(const Array2& a, const Array2::Decomp& d)
{
    const std::size_t n = a.rows;
    double worst = 0;
    for (std::size_t i = 0; i < n; ++i)
        for (std::size_t j = 0; j < n; ++j) {
            double llt = 0;
            for (std::size_t k = 0; k < n; ++k) llt += d.l(i, k) * d.l(j, k);
            double rhs = a(d.p[i], d.p[j]) + (i == j ? d.e[d.p[i]] : 0.0);
            worst = std::fmax(worst, std::fabs(llt - rhs));
        }
    return worst;
}

int main()
{
    if (modchol_device_count() < 1) { std::printf("SKIP: no CUDA device\n"); return 77; }
    Array2 a(3, 3, {1.0, 1.0, 2.0, 1.0, 1.0, 3.0, 2.0, 3.0, 1.0});
    for (Mode m : {Mode::Strict, Mode::Fast}) {
        a.mode = m;
        auto d = a.mod_cholesky_se99();
        const double res_l[9] = {1.732050807568877, 0.0, 0.0, 0.5773502691896257, 1.698920954907997, 0.0,
                                 1.154700538379251, 1.37342077428181, 0.006912871809428971};
        for (int i = 0; i < 9; ++i) CHECK(std::fabs(d.l.data[i] - res_l[i]) < 1e-12);  // lib.rs:54
        CHECK(identity_residual(a, d) < 1e-12);                                          // lib.rs:57-58
        CHECK(identity_residual(a, a.mod_cholesky_se90()) < 1e-12);                      // se90.rs:211
        CHECK(identity_residual(a, a.mod_cholesky_gmw81()) < 1e-12);                     // gmw81.rs:188
    }
    // gershgorin.rs:47-67
    Array2 g(4, 4, {10.0, -1.0, 0.0, 1.0, 0.2, 8.0, 0.2, 0.2, 1.0, 1.0, 2.0, 1.0, -1.0, -1.0, -1.0, -11.0});
    auto c = g.gershgorin_circles();
    const double want[4][2] = {{10.0, 2.0}, {8.0, 0.6}, {2.0, 1.2}, {-11.0, 2.2}};
    for (int i = 0; i < 4; ++i) {
        CHECK(std::fabs(c[i].first - want[i][0]) < 4.5e-16);
        CHECK(std::fabs(c[i].second - want[i][1]) < 4.5e-16);
    }
    // non-square -> the reference's assert!(self.is_square())
    bool threw = false;
    try { Array2(2, 3).mod_cholesky_se99(); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw);
    // default trait bodies panic!("Not implemented!") (se99.rs:26-28)
    struct Foo : ModCholeskySE99<int, int, int> {};
    threw = false;
    try { Foo().mod_cholesky_se99(); } catch (const std::logic_error& e) { threw = std::string(e.what()) == "Not implemented!"; }
    CHECK(threw);
    // batched Array3 entry point agrees with the per-matrix calls
    Array3 b(5, 3);
    for (std::size_t k = 0; k < 5; ++k)
        for (std::size_t i = 0; i < 3; ++i)
            for (std::size_t j = 0; j < 3; ++j) b(k, i, j) = a(i, j) + (i == j ? 0.25 * k : 0.0);
    auto rb = b.mod_cholesky_se99_batched();
    a.mode = Mode::Strict;
    auto d0 = a.mod_cholesky_se99();
    for (int i = 0; i < 9; ++i) CHECK(rb.l[i] == d0.l.data[i]);
    for (int i = 0; i < 3; ++i) CHECK(rb.p[i] == d0.p[i] && rb.e[i] == d0.e[i]);
    std::printf(fails ? "facade: %d failures\n" : "facade ok\n", fails);
    return fails ? 1 : 0;
}
