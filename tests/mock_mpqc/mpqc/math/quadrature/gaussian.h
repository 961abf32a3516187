// MOCK of mpqc/math/quadrature/gaussian.h: declarations of what ccsd_t.h's (uninstantiated) Laplace code names.
#pragma once
#include "mpqc/math/external/eigen/eigen.h"
namespace mpqc {
namespace math {
void gauss_legendre(int N, Eigen::VectorXd& w, Eigen::VectorXd& x);
}  // namespace math
}  // namespace mpqc
