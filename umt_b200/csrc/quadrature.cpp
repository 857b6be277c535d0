// Product quadrature sets (Y2): what rt/quadProduct.F90:97-182 (3-D) and the product
// branch of rt/quadrz.F90 (RZ) build from the tabulated polar/azimuthal rules, the
// normalisation and starting/finishing flags of rt/rtquad.F90:95-127 and the RZ
// angular-derivative coefficients of rt/AngleCoef2D.F90 + mods/AngleSet_mod.F90:337-347.
#include <cmath>

#include "umt_internal.h"
#include "quad_tables.inc"

namespace {
const double kPi = 3.14159265358979323846;
struct Rule { const double *x, *w; int n; };
// row N of a triangular table starts at N(N-1)/2 (QuadratureData_mod.F90:1144-1145)
Rule row(const double *x, const double *w, int N) { return {x + N * (N - 1) / 2, w + N * (N - 1) / 2, N}; }
}  // namespace

int umt_host_finish_quadrature(int ndim, std::vector<double> &omega, std::vector<double> &weight, std::vector<unsigned char> &start,
                               std::vector<unsigned char> &finish, std::vector<double> &angDerivFac, std::vector<double> &w1,
                               std::vector<double> &w2);

int umt_host_product_quadrature(int ndim, int npolar, int nazimuthal, int polaraxis, std::vector<double> &omega,
                                std::vector<double> &weight, std::vector<unsigned char> &start,
                                std::vector<unsigned char> &finish, std::vector<double> &angDerivFac,
                                std::vector<double> &w1, std::vector<double> &w2) {
  if (npolar < 1 || npolar > 32 || nazimuthal < 1 || nazimuthal > 32 || polaraxis < 1 || polaraxis > 3) return 1;
  const Rule polar = row(UMT_QT_cosTheta, UMT_QT_weightTheta, npolar);
  omega.clear(); weight.clear();
  if (ndim == 3) {
    const Rule azi = row(UMT_QT_cosPhiXYZ, UMT_QT_weightPhiXYZ, nazimuthal);
    // base ordinates of the first octant: azimuth ascending, polar cosine ascending (table is descending)
    for (int ia = 0; ia < azi.n; ia++)
      for (int ip = polar.n - 1; ip >= 0; ip--) {
        const double ct = polar.x[ip], st = std::sqrt(1.0 - ct * ct);
        double v[3];   // (polar-axis component, next axis, third axis)
        v[0] = ct;
        v[1] = st * azi.x[ia];
        // keeps |omega| = 1 to rounding; subtraction order as quadProduct.F90:97-111 (x, y, z order of the two known components)
        v[2] = polaraxis == 3 ? std::sqrt(1.0 - v[1] * v[1] - v[0] * v[0]) : std::sqrt(1.0 - v[0] * v[0] - v[1] * v[1]);
        double o[3];
        for (int k = 0; k < 3; k++) o[(polaraxis - 1 + k) % 3] = v[k];
        const double w = polar.w[ip] * azi.w[ia];
        // eight octant images, numbered consecutively (quadProduct.F90:122-182)
        for (int oct = 0; oct < 8; oct++) {
          const bool nx = (oct & 3) == 1 || (oct & 3) == 2, ny = (oct & 3) >= 2, nzg = oct >= 4;
          omega.push_back(nx ? -o[0] : o[0]);
          omega.push_back(ny ? -o[1] : o[1]);
          omega.push_back(nzg ? -o[2] : o[2]);
          weight.push_back(w);
        }
      }
  } else {
    const Rule azi = row(UMT_QT_cosPhiRZ, UMT_QT_weightPhiRZ, nazimuthal);
    for (int ip = polar.n - 1; ip >= 0; ip--) {
      const double xi = polar.x[ip], st = std::sqrt(1.0 - xi * xi);
      for (int half = 0; half < 2; half++) {          // xi < 0 level first, then xi > 0
        const double sxi = half == 0 ? -xi : xi;
        auto push = [&](double mu, double w) { omega.push_back(mu); omega.push_back(sxi); weight.push_back(w); };
        push(-std::sqrt(1.0 - xi * xi), 0.0);           // starting direction
        for (int ia = azi.n - 1; ia >= 0; ia--) push(-st * azi.x[ia], polar.w[ip] * azi.w[ia]);
        for (int ia = 0; ia < azi.n; ia++) push(st * azi.x[ia], polar.w[ip] * azi.w[ia]);
        push(std::sqrt(1.0 - xi * xi), 0.0);            // finishing direction
      }
    }
  }
  return umt_host_finish_quadrature(ndim, omega, weight, start, finish, angDerivFac, w1, w2);
}

// the GTA angle set in r-z: level-symmetric S2 (rt/quadrz.F90:82-160 with norder = 2): per xi-level (xi = -+1/sqrt 3) a starting
// direction, mu = -1/sqrt 3, mu = +1/sqrt 3 (weight pi/2 each) and a finishing direction
int umt_host_gta_quadrature_rz(std::vector<double> &omega, std::vector<double> &weight, std::vector<unsigned char> &start,
                               std::vector<unsigned char> &finish, std::vector<double> &angDerivFac, std::vector<double> &w1,
                               std::vector<double> &w2) {
  const double xilev = 0.577350269189625, mu = 0.577350269189625;   // QuadratureData_mod.F90:711-713
  omega.clear(); weight.clear();
  for (int half = 0; half < 2; half++) {
    const double sxi = half == 0 ? -xilev : xilev;
    auto push = [&](double m, double w) { omega.push_back(m); omega.push_back(sxi); weight.push_back(w); };
    push(-std::sqrt(1.0 - xilev * xilev), 0.0);
    push(-mu, 0.5 * kPi);
    push(mu, 0.5 * kPi);
    push(std::sqrt(1.0 - xilev * xilev), 0.0);
  }
  return umt_host_finish_quadrature(2, omega, weight, start, finish, angDerivFac, w1, w2);
}

// normalisation, starting/finishing flags (rtquad.F90:95-127) and, in r-z, the weighted-diamond coefficients of every xi-level
int umt_host_finish_quadrature(int ndim, std::vector<double> &omega, std::vector<double> &weight, std::vector<unsigned char> &start,
                               std::vector<unsigned char> &finish, std::vector<double> &angDerivFac, std::vector<double> &w1,
                               std::vector<double> &w2) {
  const int NA = (int)weight.size();
  // sum of weights * wtiso = 1 (rtquad.F90:95-105; wtiso = 1/4pi in xyz, 1/2pi in rz, Size_mod.F90:278-281)
  const double wtiso = ndim == 3 ? 1.0 / (4.0 * kPi) : 1.0 / (2.0 * kPi);
  double sum = 0.0;
  for (double w : weight) sum += w;
  const double fac = 1.0 / (wtiso * sum);
  for (double &w : weight) w = fac * w;
  start.assign(NA, 0); finish.assign(NA, 0);
  angDerivFac.assign(NA, 0.0); w1.assign(NA, 1.0); w2.assign(NA, 0.0);
  if (ndim == 3) return 0;
  // zero-weight directions alternate starting / finishing (rtquad.F90:107-127)
  bool expectStart = true;
  for (int a = 0; a < NA; a++)
    if (weight[a] < 2.220446049250313e-16) { (expectStart ? start : finish)[a] = 1; expectStart = !expectStart; }
  // weighted-diamond coefficients per xi-level (AngleCoef2D.F90)
  std::vector<double> alpha(NA, 0.0), tau(NA, 0.0);
  int a = 0;
  while (a < NA) {
    int b = a + 1;
    while (b < NA && !start[b]) b++;                   // level = [a, b)
    double wl = 0.0;
    for (int k = a; k < b; k++) wl += weight[k];
    double phim = kPi, mum = omega[2 * a];
    for (int k = a; k < b; k++) {
      if (start[k] || finish[k]) continue;
      alpha[k] = alpha[k - 1] - weight[k] * omega[2 * k];
      const double phip = phim - weight[k] * kPi / wl;
      const double mup = std::sqrt(1.0 - omega[2 * k + 1] * omega[2 * k + 1]) * std::cos(phip);
      if (omega[2 * k] < mum || omega[2 * k] > mup) return 2;   // "Mu not between limits"
      tau[k] = (omega[2 * k] - mum) / (mup - mum);
      phim = phip; mum = mup;
      angDerivFac[k] = omega[2 * k] + alpha[k] / (weight[k] * tau[k]);
      w1[k] = 1.0 / tau[k];
      w2[k] = (1.0 - tau[k]) / tau[k];
    }
    a = b;
  }
  return 0;
}
