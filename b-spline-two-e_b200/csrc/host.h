// host.h -- host-side companions of the hot path (see host.cpp)
#pragma once
#include <complex>
#include <cstdint>
#include <vector>

namespace bs2e {
namespace host {

struct SymLabel { int l, m; bool pi; };  // orbital_tools.f90:10-13 (type sym)

std::vector<double> generate_grid(int k, int m, int Z, double h_max, double r_max);
void gauss_legendre(int N, double a, double b, double* x, double* w);
int find_max_n_b(int k, const std::vector<double>& knots, double x);
void setup_S(int ks, const std::vector<double>& knots, int k_GL, std::complex<double>* S);
void setup_H_one_particle(int ks, const std::vector<double>& knots, int Z, int l, int CAP_order,
                          double CAP_r_0, std::complex<double> CAP_eta, int k_GL,
                          std::complex<double>* H);
void setup_radial_dip(int ks, const std::vector<double>& knots, int k_GL, int gauge, std::complex<double>* A,
                      std::complex<double>* B);
void count_configs(int L, bool pi, int max_l_1p, int n_b, int k_spline, int max_n_b, int n_all_l,
                   int l_2_max, std::vector<int64_t>& conf_n, std::vector<int64_t>& conf_l,
                   std::vector<int64_t>& conf_eqv);
std::vector<SymLabel> basis_syms(int max_L, bool z_pol);

}  // namespace host
}  // namespace bs2e
