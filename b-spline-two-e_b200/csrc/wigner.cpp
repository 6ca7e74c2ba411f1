// wigner.cpp -- exact Wigner 3j(000) / 6j for integer angular momenta and the
// angular factor ang_k_LS of the coupled two-electron matrix elements.
//
// Replaces the GSL calls behind src/tools/wigner_tools.f90:30-60 and restates
// :107-112 (C_red_mat) and :126-138 (ang_k_LS).  Values are evaluated from the
// prime factorisation of the Racah formula: every term of the 6j sum is an
// exact integer (arbitrary precision), only the final square root is rounded,
// so structural zeros are exact zeros and the sparsity thresholds of
// hamiltonian.f90:174 / mat_els.f90:568 are never decided by rounding noise.
#include "wigner.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace bs2e {
namespace {

// ---- prime table and factorial exponents ----------------------------------
struct Primes {
    std::vector<int> p;
    explicit Primes(int nmax)
    {
        std::vector<char> sieve(nmax + 1, 1);
        for (int i = 2; i <= nmax; ++i) {
            if (!sieve[i]) continue;
            p.push_back(i);
            for (long long j = 1LL * i * i; j <= nmax; j += i) sieve[j] = 0;
        }
    }
};
const Primes& primes()
{
    static const Primes P(4096);
    return P;
}

// exponent vector of a product of factorials
struct Expo {
    std::vector<int> e;  // per prime index
    void ensure(size_t n) { if (e.size() < n) e.resize(n, 0); }
    void mul_fact(int n, int sign = +1)
    {
        if (n < 0) throw std::domain_error("negative factorial");
        const auto& pr = primes().p;
        for (size_t i = 0; i < pr.size() && pr[i] <= n; ++i) {
            int cnt = 0;
            for (long long q = pr[i]; q <= n; q *= pr[i]) cnt += n / (int)q;
            ensure(i + 1);
            e[i] += sign * cnt;
        }
    }
    void div_fact(int n) { mul_fact(n, -1); }
    void mul_int(int n, int sign = +1)
    {
        const auto& pr = primes().p;
        for (size_t i = 0; i < pr.size() && n > 1; ++i)
            while (n % pr[i] == 0) { ensure(i + 1); e[i] += sign; n /= pr[i]; }
    }
};

// ---- minimal unsigned big integer ----------------------------------------
struct BigU {
    std::vector<uint32_t> d;  // little endian
    bool zero() const { return d.empty(); }
    void mul(uint32_t m)
    {
        uint64_t carry = 0;
        for (auto& x : d) { uint64_t v = (uint64_t)x * m + carry; x = (uint32_t)v; carry = v >> 32; }
        if (carry) d.push_back((uint32_t)carry);
    }
    static BigU from_expo(const Expo& ex)
    {
        BigU r;
        r.d.push_back(1);
        const auto& pr = primes().p;
        for (size_t i = 0; i < ex.e.size(); ++i) {
            if (ex.e[i] < 0) throw std::logic_error("non-integer Racah term");
            for (int q = 0; q < ex.e[i]; ++q) r.mul((uint32_t)pr[i]);
        }
        return r;
    }
    static int cmp(const BigU& a, const BigU& b)
    {
        if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
        for (size_t i = a.d.size(); i-- > 0;)
            if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
        return 0;
    }
    void add(const BigU& b)
    {
        if (d.size() < b.d.size()) d.resize(b.d.size(), 0);
        uint64_t carry = 0;
        for (size_t i = 0; i < d.size(); ++i) {
            uint64_t v = (uint64_t)d[i] + (i < b.d.size() ? b.d[i] : 0) + carry;
            d[i] = (uint32_t)v;
            carry = v >> 32;
        }
        if (carry) d.push_back((uint32_t)carry);
    }
    void sub(const BigU& b)  // requires *this >= b
    {
        int64_t borrow = 0;
        for (size_t i = 0; i < d.size(); ++i) {
            int64_t v = (int64_t)d[i] - (i < b.d.size() ? b.d[i] : 0) - borrow;
            borrow = v < 0;
            if (v < 0) v += (int64_t)1 << 32;
            d[i] = (uint32_t)v;
        }
        while (!d.empty() && d.back() == 0) d.pop_back();
    }
    long double to_ld() const
    {
        long double v = 0.0L;
        for (size_t i = d.size(); i-- > 0;) v = v * 4294967296.0L + (long double)d[i];
        return v;
    }
};

// sqrt( prod p^e ) in extended precision
long double sqrt_expo(const Expo& ex)
{
    const auto& pr = primes().p;
    long double num = 1.0L, den = 1.0L;
    for (size_t i = 0; i < ex.e.size(); ++i) {
        int e = ex.e[i];
        if (e == 0) continue;
        int ae = std::abs(e);
        long double f = std::pow((long double)pr[i], (long double)(ae / 2));
        if (ae & 1) f *= std::sqrt((long double)pr[i]);
        if (e > 0) num *= f; else den *= f;
    }
    return num / den;
}

bool triangle(int a, int b, int c) { return a + b >= c && a + c >= b && b + c >= a; }

// Delta(a,b,c) = (a+b-c)!(a-b+c)!(-a+b+c)!/(a+b+c+1)!
void mul_delta(Expo& ex, int a, int b, int c)
{
    ex.mul_fact(a + b - c);
    ex.mul_fact(a - b + c);
    ex.mul_fact(-a + b + c);
    ex.div_fact(a + b + c + 1);
}

}  // namespace

double three_j0(int ja, int jb, int jc)
{
    if (ja < 0 || jb < 0 || jc < 0 || !triangle(ja, jb, jc)) return 0.0;
    const int J = ja + jb + jc;
    if (J & 1) return 0.0;
    const int g = J / 2;
    Expo ex;  // value^2 = Delta * (g!/((g-ja)!(g-jb)!(g-jc)!))^2
    mul_delta(ex, ja, jb, jc);
    for (int rep = 0; rep < 2; ++rep) {
        ex.mul_fact(g);
        ex.div_fact(g - ja);
        ex.div_fact(g - jb);
        ex.div_fact(g - jc);
    }
    const long double v = sqrt_expo(ex);
    return (double)((g & 1) ? -v : v);
}

double six_j(int j1, int j2, int j3, int j4, int j5, int j6)
{
    if (j1 < 0 || j2 < 0 || j3 < 0 || j4 < 0 || j5 < 0 || j6 < 0) return 0.0;
    if (!triangle(j1, j2, j3) || !triangle(j1, j5, j6) || !triangle(j4, j2, j6) ||
        !triangle(j4, j5, j3))
        return 0.0;
    const int a[4] = {j1 + j2 + j3, j1 + j5 + j6, j4 + j2 + j6, j4 + j5 + j3};
    const int b[3] = {j1 + j2 + j4 + j5, j2 + j3 + j5 + j6, j3 + j1 + j6 + j4};
    const int tmin = std::max(std::max(a[0], a[1]), std::max(a[2], a[3]));
    const int tmax = std::min(b[0], std::min(b[1], b[2]));
    if (tmax < tmin) return 0.0;
    BigU pos, neg;
    for (int t = tmin; t <= tmax; ++t) {
        Expo ex;  // (t+1)! / [prod (t-a_i)! prod (b_j-t)!]  -- an integer
        ex.mul_fact(t + 1);
        for (int q = 0; q < 4; ++q) ex.div_fact(t - a[q]);
        for (int q = 0; q < 3; ++q) ex.div_fact(b[q] - t);
        BigU term = BigU::from_expo(ex);
        if (t & 1) neg.add(term); else pos.add(term);
    }
    const int c = BigU::cmp(pos, neg);
    if (c == 0) return 0.0;
    long double s;
    if (c > 0) { pos.sub(neg); s = pos.to_ld(); }
    else { neg.sub(pos); s = -neg.to_ld(); }
    Expo dl;
    mul_delta(dl, j1, j2, j3);
    mul_delta(dl, j1, j5, j6);
    mul_delta(dl, j4, j2, j6);
    mul_delta(dl, j4, j5, j3);
    return (double)(s * sqrt_expo(dl));
}

// wigner_tools.f90:107-112
// General 3j symbol for integer j, m (wigner_tools.f90:30-45 calls gsl_sf_coupling_3j):
// Racah's single sum in long double; the arguments of this path are small (the dipole
// Wigner-Eckart factor (L 1 L'; -M q M')), so factorials stay far inside the range.
double three_j(int ja, int jb, int jc, int ma, int mb, int mc)
{
    auto fact = [](int n) { long double f = 1.0L; for (int q = 2; q <= n; ++q) f *= (long double)q; return f; };
    auto tri = [](int a, int b, int c) { return a + b >= c && a + c >= b && b + c >= a; };
    if (ja < 0 || jb < 0 || jc < 0 || ma + mb + mc != 0) return 0.0;
    if (std::abs(ma) > ja || std::abs(mb) > jb || std::abs(mc) > jc || !tri(ja, jb, jc)) return 0.0;
    const int t1 = jb - jc - ma, t2 = ja + mb - jc, t3 = ja + jb - jc, t4 = ja - ma, t5 = jb + mb;
    const int tmin = std::max(0, std::max(t1, t2)), tmax = std::min(t3, std::min(t4, t5));
    if (tmax < tmin) return 0.0;
    long double sum = 0.0L;
    for (int t = tmin; t <= tmax; ++t) {
        const long double d = fact(t) * fact(t - t1) * fact(t - t2) * fact(t3 - t) * fact(t4 - t) * fact(t5 - t);
        sum += ((t & 1) ? -1.0L : 1.0L) / d;
    }
    const long double delta = fact(ja + jb - jc) * fact(ja - jb + jc) * fact(-ja + jb + jc) / fact(ja + jb + jc + 1);
    long double v = sqrtl(delta * fact(ja + ma) * fact(ja - ma) * fact(jb + mb) * fact(jb - mb) * fact(jc + mc) *
                          fact(jc - mc)) * sum;
    if ((ja - jb - mc) & 1) v = -v;
    return (double)v;
}

double C_red_mat(int k, int a, int b)
{
    const double sgn = (a & 1) ? -1.0 : 1.0;
    return sgn * std::sqrt((double)((2 * a + 1) * (2 * b + 1))) * three_j0(a, k, b);
}

// wigner_tools.f90:126-138
double ang_k_LS(int k, int la, int lb, int lc, int ld, int L)
{
    if (((la + k + lc) & 1) || ((lb + k + ld) & 1)) return 0.0;
    const double sgn = ((lb + lc + L) & 1) ? -1.0 : 1.0;
    return sgn * six_j(la, lb, L, ld, lc, k) * C_red_mat(k, la, lc) * C_red_mat(k, lb, ld);
}

}  // namespace bs2e
