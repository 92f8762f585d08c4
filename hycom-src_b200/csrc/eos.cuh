// Equation of state of HYCOM (stmt_fns.h): sig(t,s) and tofsig(r,s).
//
// The reference compiles exactly one family in (-DEOS_SIG0|-DEOS_SIG2 with -DEOS_7T|9T|12T|17T,
// stmt_fns.h:2-22 defines `sigver` 1..8 from them); here the family is the run-time scalar
// `sigver` of hycom_tsadvc_params so one library serves every build of the host model.
// The coefficients - including the ones stmt_fns.h derives at compile time for the reference
// pressure (c101.., rc6) - are formed once on the host with the same double operations
// gfortran's constant folder uses and travel in the kernel parameter block, so the device code
// reads them as constant-bank operands instead of materialising 64-bit immediates.
// Everything is compiled with -fmad=false / -ffp-contract=off, divisions and square roots are
// IEEE round-to-nearest, and every expression keeps the Fortran grouping (x**2 = x*x,
// x**3 = (x*x)*x), so sig() of all families and tofsig() of the 12-term family reproduce a
// non-contracting CPU evaluation bit for bit; tofsig() of the 7/9-term families goes through
// atan2/cos/sin whose last bit is library specific (tests allow 1e-12 there).
#pragma once
#include <cuda_runtime.h>

#include "march_common.cuh"

namespace tsadvc {
namespace eos {

struct Coef {
  int family;     // 7, 9, 12 or 17 terms
  int pad;
  double c[18];   // meaning per family, see fill()
};

// stmt_fns.h:53-110 (7/9-term), :129-172 (12-term), :216-290 (17-term)
inline void fill(int sigver, Coef& e) {
  for (double& x : e.c) x = 0.0;
  e.pad = 0;
  const double rpdb = ((sigver % 2 == 1) ? 0.0 : 2000.e4) * 1.e-4;   // pref*prs2pdb
  if (sigver >= 1 && sigver <= 4) {
    static const double k[4][9] = {
        {-1.36471E-01, 4.68181E-02, 8.07004E-01, -7.45353E-03, -2.94418E-03, 3.43570E-05, 3.48658E-05, 0.0, 0.0},
        {9.77093E+00, -2.26493E-02, 7.89879E-01, -6.43205E-03, -2.62983E-03, 2.75835E-05, 3.15235E-05, 0.0, 0.0},
        {-4.311829E-02, 5.429948E-02, 8.011774E-01, -7.641336E-03, -3.258442E-03, 3.757643E-05, 3.630361E-05,
         8.675546E-05, 3.995086E-06},
        {9.903308E+00, -1.618075E-02, 7.819166E-01, -6.593939E-03, -2.896464E-03, 3.038697E-05, 3.266933E-05,
         1.180109E-04, 3.399511E-06}};
    e.family = sigver <= 2 ? 7 : 9;
    for (int q = 0; q < 9; ++q) e.c[q] = k[sigver - 1][q];   // c1..c9
    e.c[9] = 1.0 / e.c[5];                                    // rc6
    e.c[10] = 1.0 / 3.0;                                      // a3rd
  } else if (sigver == 7 || sigver == 8) {
    e.family = 12;
    e.c[0] = -1.4627567840659594e-01 + rpdb * 5.0879498675039621e-03;   // c101
    e.c[1] = 6.4247392832635697e-02 + rpdb * 1.6333913018305079e-05;    // c102
    e.c[2] = 8.1213979591704621e-01 + rpdb * 4.3899924880543972e-06;    // c103
    e.c[3] = -8.1321489441909698e-03;                                    // c004
    e.c[4] = 4.5199845091090296e-03;                                     // c005
    e.c[5] = 4.6347888132781394e-04;                                     // c006
    e.c[6] = 1.0000000000000000e+00 + rpdb * 1.1995545126831476e-05;    // c111
    e.c[7] = 1.0316374535350838e-02 + rpdb * 5.5234008384648383e-08;    // c112
    e.c[8] = 8.9521792365142522e-04 + rpdb * 8.4310335919950873e-09;    // c113
    e.c[9] = -2.8438341552142710e-05;                                    // c014
    e.c[10] = -1.1887778959461776e-05;                                   // c015
    e.c[11] = -4.0163964812921489e-06;                                   // c016
  } else {
    e.family = 17;
    e.c[0] = 9.9984085444849347e+02 + (1.1798263740430364e-02 - 2.5862187075154352e-08 * rpdb) * rpdb;    // c101
    e.c[1] = 7.3471625860981584e+00;                                                                      // c002
    e.c[2] = -5.3211231792841769e-02 + (9.8920219266399117e-08 - 3.2921414007960662e-12 * rpdb) * rpdb;   // c103
    e.c[3] = 3.6492439109814549e-04;                                                                      // c004
    e.c[4] = 2.5880571023991390e+00 + 4.6996642771754730e-06 * rpdb;                                      // c105
    e.c[5] = 6.7168282786692355e-03;                                                                      // c006
    e.c[6] = 1.9203202055760151e-03;                                                                      // c007
    e.c[7] = 1.0000000000000000e+00 + 6.7103246285651894e-06 * rpdb;                                      // c108
    e.c[8] = 7.2815210113327091e-03 - 9.1534417604289062e-18 * (rpdb * rpdb * rpdb);                      // c109
    e.c[9] = -4.4787265461983921e-05;                                                                     // c010
    e.c[10] = 3.3851002965802430e-07 - 2.4461698007024582e-17 * (rpdb * rpdb);                            // c111
    e.c[11] = 1.3651202389758572e-10;                                                                     // c012
    e.c[12] = 1.7632126669040377e-03;                                                                     // c013
    e.c[13] = 8.8066583251206474e-06;                                                                     // c014
    e.c[14] = 1.8832689434804897e-10;                                                                     // c015
    e.c[15] = 5.7463776745432097e-06;                                                                     // c016
    e.c[16] = 1.4716275472242334e-09;                                                                     // c017
  }
}

// aone/x, IEEE round-to-nearest: the short division sequence with its exact fallback
__device__ __forceinline__ double rcp_exact(double x) { return div_rn(1.0, x); }

// root of t**3 + a2*t**2 + a1*t + a0 = 0 the way stmt_fns.h:311-323 / :348-361 + :379 take it
__device__ __forceinline__ double cubic_root(double a3rd, double a0, double a1, double a2) {
  const double a2t = a3rd * a2;
  const double q = a3rd * a1 - a2t * a2t;                                    // cubq
  const double r = a3rd * (0.5 * a1 * a2 - 1.5 * a0) - a2t * a2t * a2t;      // cubr
  const double disc = -(q * q * q + r * r);
  const double an = a3rd * atan2(sqrt(0.0 > disc ? 0.0 : disc), r);          // cuban
  const double sq = sqrt(-q);
  return -(sq * cos(an)) + sqrt(3.0) * (sq * sin(an)) - a2t;
}

// sigma(t,s)
__device__ __forceinline__ double sig(const Coef& e, double t, double s) {
  const double* c = e.c;
  if (e.family == 17) {
    const double num = c[0] + t * (c[1] + t * (c[2] + t * c[3])) + s * (c[4] - t * c[5] + s * c[6]);   // :503-504
    const double sp = 0.0 > s ? 0.0 : s;                                                              // max(sqrmin,s)
    const double den = c[7] + t * (c[8] + t * (c[9] + t * (c[10] + t * c[11]))) +
                       s * (c[12] - t * (c[13] + t * t * c[14]) + sqrt(sp) * (c[15] + t * t * c[16]));  // :505-507
    return num * rcp_exact(den) - 1000.0;                                                             // :508-509
  }
  if (e.family == 7)    // :332
    return (c[0] + c[2] * s + t * (c[1] + c[4] * s + t * (c[3] + c[6] * s + c[5] * t)));
  if (e.family == 9)    // :368-369
    return (c[0] + s * (c[2] + s * c[7]) + t * (c[1] + s * (c[4] + s * c[8]) + t * (c[3] + s * c[6] + t * c[5])));
  // 12-term, :419-424
  const double num = c[0] + (c[1] + c[3] * t + c[4] * s) * t + (c[2] + c[5] * s) * s;
  const double den = c[6] + (c[7] + c[9] * t + c[10] * s) * t + (c[8] + c[11] * s) * s;
  return num * rcp_exact(den);
}

// temperature from sigma and salinity
__device__ __forceinline__ double tofsig(const Coef& e, double r, double s) {
  const double* c = e.c;
  if (e.family == 7)    // :308-310, :323
    return cubic_root(c[10], (c[0] + c[2] * s - r) * c[9], (c[1] + c[4] * s) * c[9], (c[3] + c[6] * s) * c[9]);
  if (e.family == 9)    // :349-351, :379
    return cubic_root(c[10], (c[0] + s * (c[2] + s * c[7]) - r) * c[9], (c[1] + s * (c[4] + s * c[8])) * c[9],
                      (c[3] + s * c[6]) * c[9]);
  if (e.family == 12) {  // :441-449
    const double qa = (c[3] - r * c[9]);
    const double qb = ((c[1] + c[4] * s) - r * (c[7] + c[10] * s));
    const double qc = ((c[0] + (c[2] + c[5] * s) * s) - r * (c[6] + (c[8] + c[11] * s) * s));
    const double disc = qb * qb - 4.0 * qa * qc;
    return (-qb - sqrt(0.0 > disc ? 0.0 : disc)) / (2.0 * qa);
  }
  return 99.0;          // :533 "NOT AVAILABLE AS AN EXPRESSION" for the 17-term fit
}

}  // namespace eos
}  // namespace tsadvc
