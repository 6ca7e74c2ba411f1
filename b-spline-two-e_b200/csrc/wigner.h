// wigner.h -- exact angular-momentum coupling coefficients (host side).
// See wigner.cpp; mirrors src/tools/wigner_tools.f90 of the reference.
#pragma once

namespace bs2e {

double three_j0(int ja, int jb, int jc);                          // (ja jb jc; 0 0 0)
double three_j(int ja, int jb, int jc, int ma, int mb, int mc);   // wigner_tools.f90:30-45 (integer j, m)
double six_j(int j1, int j2, int j3, int j4, int j5, int j6);      // {j1 j2 j3; j4 j5 j6}
double C_red_mat(int k, int a, int b);                             // wigner_tools.f90:107-112
double ang_k_LS(int k, int la, int lb, int lc, int ld, int L);     // wigner_tools.f90:126-138

}  // namespace bs2e
