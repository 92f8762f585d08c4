// Equation of state of HYCOM (stmt_fns.h) as device functions: sig(t,s) and tofsig(r,s).
//
// The reference compiles exactly one family in (-DEOS_SIG0|-DEOS_SIG2 with -DEOS_7T|9T|12T|17T,
// stmt_fns.h:2-22 defines `sigver` 1..8 from them); here the family is the run-time scalar
// `sigver` of hycom_tsadvc_params so one library serves every build of the host model.
// The translation unit is compiled with -fmad=false, divisions and square roots are IEEE
// round-to-nearest, and every expression keeps the Fortran grouping (x**2 = x*x,
// x**3 = (x*x)*x), so sig() of all families and tofsig() of the 12-term family reproduce a
// non-contracting CPU evaluation bit for bit; tofsig() of the 7/9-term families goes through
// atan2/cos/sin whose last bit is library specific (tests allow 1e-12 there).
#pragma once
#include <cuda_runtime.h>

#include "march_common.cuh"

namespace tsadvc {
namespace eos {

// aone/x, IEEE round-to-nearest: the short division sequence with its exact fallback
__device__ __forceinline__ double rcp_exact(double x) { return div_rn(1.0, x); }

struct Poly79 { double c1, c2, c3, c4, c5, c6, c7, c8, c9; };

__device__ __forceinline__ Poly79 poly79(int sigver) {
  switch (sigver) {
    case 1:   // 7-term sigma-0, stmt_fns.h:53-61
      return {-1.36471E-01, 4.68181E-02, 8.07004E-01, -7.45353E-03, -2.94418E-03, 3.43570E-05,
              3.48658E-05, 0.0, 0.0};
    case 2:   // 7-term sigma-2, :65-73
      return {9.77093E+00, -2.26493E-02, 7.89879E-01, -6.43205E-03, -2.62983E-03, 2.75835E-05,
              3.15235E-05, 0.0, 0.0};
    case 3:   // 9-term sigma-0, :86-96
      return {-4.311829E-02, 5.429948E-02, 8.011774E-01, -7.641336E-03, -3.258442E-03,
              3.757643E-05, 3.630361E-05, 8.675546E-05, 3.995086E-06};
    default:  // 9-term sigma-2, :100-110
      return {9.903308E+00, -1.618075E-02, 7.819166E-01, -6.593939E-03, -2.896464E-03,
              3.038697E-05, 3.266933E-05, 1.180109E-04, 3.399511E-06};
  }
}

// root of t**3 + a2*t**2 + a1*t + a0 = 0 the way stmt_fns.h:311-323 / :348-361 + :379 take it
__device__ __forceinline__ double cubic_root(double a0, double a1, double a2) {
  const double a3rd = 1.0 / 3.0;
  const double a2t = a3rd * a2;
  const double q = a3rd * a1 - a2t * a2t;                                    // cubq
  const double r = a3rd * (0.5 * a1 * a2 - 1.5 * a0) - a2t * a2t * a2t;      // cubr
  const double disc = -(q * q * q + r * r);
  const double an = a3rd * atan2(sqrt(0.0 > disc ? 0.0 : disc), r);          // cuban
  const double sq = sqrt(-q);
  return -(sq * cos(an)) + sqrt(3.0) * (sq * sin(an)) - a2t;
}

// 12-term rational function at the reference pressure (:129-172)
struct Rat12 { double n1, n2, n3, n4, n5, n6, d1, d2, d3, d4, d5, d6; };
__device__ __forceinline__ Rat12 rat12(int sigver) {
  const double rpdb = (sigver == 7 ? 0.0 : 2000.e4) * 1.e-4;
  Rat12 c;
  c.n1 = -1.4627567840659594e-01 + rpdb * 5.0879498675039621e-03;   // c101
  c.n2 = 6.4247392832635697e-02 + rpdb * 1.6333913018305079e-05;    // c102
  c.n3 = 8.1213979591704621e-01 + rpdb * 4.3899924880543972e-06;    // c103
  c.n4 = -8.1321489441909698e-03;                                    // c004
  c.n5 = 4.5199845091090296e-03;                                     // c005
  c.n6 = 4.6347888132781394e-04;                                     // c006
  c.d1 = 1.0000000000000000e+00 + rpdb * 1.1995545126831476e-05;    // c111
  c.d2 = 1.0316374535350838e-02 + rpdb * 5.5234008384648383e-08;    // c112
  c.d3 = 8.9521792365142522e-04 + rpdb * 8.4310335919950873e-09;    // c113
  c.d4 = -2.8438341552142710e-05;                                    // c014
  c.d5 = -1.1887778959461776e-05;                                    // c015
  c.d6 = -4.0163964812921489e-06;                                    // c016
  return c;
}

// 17-term rational function (Jackett et al. 2006) at the reference pressure (:216-290)
__device__ __forceinline__ double sig17(int sigver, double t, double s) {
  const double rpdb = (sigver == 5 ? 0.0 : 2000.e4) * 1.e-4;
  const double c101 = 9.9984085444849347e+02 + (1.1798263740430364e-02 - 2.5862187075154352e-08 * rpdb) * rpdb;
  const double c002 = 7.3471625860981584e+00;
  const double c103 = -5.3211231792841769e-02 + (9.8920219266399117e-08 - 3.2921414007960662e-12 * rpdb) * rpdb;
  const double c004 = 3.6492439109814549e-04;
  const double c105 = 2.5880571023991390e+00 + 4.6996642771754730e-06 * rpdb;
  const double c006 = 6.7168282786692355e-03, c007 = 1.9203202055760151e-03;
  const double c108 = 1.0000000000000000e+00 + 6.7103246285651894e-06 * rpdb;
  const double c109 = 7.2815210113327091e-03 - 9.1534417604289062e-18 * (rpdb * rpdb * rpdb);
  const double c010 = -4.4787265461983921e-05;
  const double c111 = 3.3851002965802430e-07 - 2.4461698007024582e-17 * (rpdb * rpdb);
  const double c012 = 1.3651202389758572e-10, c013 = 1.7632126669040377e-03;
  const double c014 = 8.8066583251206474e-06, c015 = 1.8832689434804897e-10;
  const double c016 = 5.7463776745432097e-06, c017 = 1.4716275472242334e-09;
  const double num = c101 + t * (c002 + t * (c103 + t * c004)) + s * (c105 - t * c006 + s * c007);   // :503-504
  const double sp = 0.0 > s ? 0.0 : s;                                                               // max(sqrmin,s)
  const double den = c108 + t * (c109 + t * (c010 + t * (c111 + t * c012))) +
                     s * (c013 - t * (c014 + t * t * c015) + sqrt(sp) * (c016 + t * t * c017));       // :505-507
  return num * rcp_exact(den) - 1000.0;                                                                 // :508-509
}

// sigma(t,s)
__device__ __forceinline__ double sig(int sigver, double t, double s) {
  if (sigver == 1 || sigver == 2) {
    const Poly79 c = poly79(sigver);   // :332
    return (c.c1 + c.c3 * s + t * (c.c2 + c.c5 * s + t * (c.c4 + c.c7 * s + c.c6 * t)));
  }
  if (sigver == 3 || sigver == 4) {
    const Poly79 c = poly79(sigver);   // :368-369
    return (c.c1 + s * (c.c3 + s * c.c8) + t * (c.c2 + s * (c.c5 + s * c.c9) + t * (c.c4 + s * c.c7 + t * c.c6)));
  }
  if (sigver == 7 || sigver == 8) {
    const Rat12 c = rat12(sigver);     // :419-424
    const double num = c.n1 + (c.n2 + c.n4 * t + c.n5 * s) * t + (c.n3 + c.n6 * s) * s;
    const double den = c.d1 + (c.d2 + c.d4 * t + c.d5 * s) * t + (c.d3 + c.d6 * s) * s;
    return num * rcp_exact(den);
  }
  return sig17(sigver, t, s);
}

// temperature from sigma and salinity
__device__ __forceinline__ double tofsig(int sigver, double r, double s) {
  if (sigver >= 1 && sigver <= 4) {
    const Poly79 c = poly79(sigver);
    const double rc6 = 1.0 / c.c6;
    if (sigver <= 2)                   // :308-310, :323
      return cubic_root((c.c1 + c.c3 * s - r) * rc6, (c.c2 + c.c5 * s) * rc6, (c.c4 + c.c7 * s) * rc6);
    return cubic_root((c.c1 + s * (c.c3 + s * c.c8) - r) * rc6, (c.c2 + s * (c.c5 + s * c.c9)) * rc6,
                      (c.c4 + s * c.c7) * rc6);   // :349-351, :379
  }
  if (sigver == 7 || sigver == 8) {    // :441-449
    const Rat12 c = rat12(sigver);
    const double qa = (c.n4 - r * c.d4);
    const double qb = ((c.n2 + c.n5 * s) - r * (c.d2 + c.d5 * s));
    const double qc = ((c.n1 + (c.n3 + c.n6 * s) * s) - r * (c.d1 + (c.d3 + c.d6 * s) * s));
    const double disc = qb * qb - 4.0 * qa * qc;
    return (-qb - sqrt(0.0 > disc ? 0.0 : disc)) / (2.0 * qa);
  }
  return 99.0;                         // :533 "NOT AVAILABLE AS AN EXPRESSION" for the 17-term fit
}

}  // namespace eos
}  // namespace tsadvc
