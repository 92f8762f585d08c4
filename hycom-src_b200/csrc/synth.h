// Synthetic mod_cb_arrays state for the tsadvc path.
//
// Every value is a pure function of the GLOBAL indices (ig,jg,k), a field id and
// a seed, so every tiling sees identical data (the reference's own test idea:
// mod_pipe compares 1 tile against N tiles bit for bit).  Only +,-,*,/ and floor
// on doubles plus integer hashing are used, so the host (g++) and device (nvcc
// -fmad=false) evaluations are bit-identical.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD inline
#endif

namespace synth {

struct Cfg {
  int32_t itdm, jtdm, kdm, nreg;
  int32_t ntracr, pad;
  uint64_t seed;
  double dx0, dy0;  // grid spacing (m)
  double delt1;     // leapfrog step (s)
};

// tile placement of the array being filled
struct Tile {
  int32_t idm, jdm, nbdy, ii, jj, i0, j0, pad;
};

enum Field : int {
  F_TEMP = 0, F_SALN = 1, F_TH3D = 2, F_DP = 3, F_UFLX = 4, F_VFLX = 5, F_TRACER = 6,
  F_SCPX = 10, F_SCPY = 11, F_SCUX = 12, F_SCUY = 13, F_SCVX = 14, F_SCVY = 15,
  F_ONETA = 16
};

constexpr double kOnem = 9806.0;                       // mod_cb_arrays.F90:842-846
constexpr double kHugel = 1.2676506002282294e30;       // 2.0**100, "land" (mod_cb_arrays.F90:853)

SYN_HD uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// uniform [0,1) keyed on (seed, stream, ig, jg, k)
SYN_HD double u01(uint64_t seed, int stream, int ig, int jg, int k) {
  uint64_t h = mix64(seed + 0x48594330ull + (uint64_t)stream * 0x100000001B3ull);
  h = mix64(h ^ ((uint64_t)(uint32_t)ig * 0x9E3779B1ull));
  h = mix64(h ^ ((uint64_t)(uint32_t)jg * 0x85EBCA77ull));
  h = mix64(h ^ ((uint64_t)(uint32_t)k * 0xC2B2AE3Dull));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

SYN_HD double ffloor(double t) {
  double f = (double)(long long)t;
  return (f > t) ? f - 1.0 : f;
}

// period-1 "sine" made of two parabolas: C1, range [-1,1], exact arithmetic
SYN_HD double wave(double t) {
  t = t - ffloor(t);
  return (t < 0.5) ? 16.0 * t * (0.5 - t) : -16.0 * (t - 0.5) * (1.0 - t);
}

SYN_HD double dmax(double a, double b) { return a > b ? a : b; }

struct XYZ { double x, y, z; };
SYN_HD XYZ coords(const Cfg& c, int ig, int jg, int k) {
  XYZ p;
  p.x = ((double)ig - 0.5) / (double)c.itdm;
  p.y = ((double)jg - 0.5) / (double)c.jtdm;
  p.z = ((double)k - 0.5) / (double)c.kdm;
  return p;
}

// smooth layer thickness in metres; exactly zero in outcrop patches
SYN_HD double hsmooth(const Cfg& c, int ig, int jg, int k) {
  const XYZ p = coords(c, ig, jg, k);
  const double a = 15.0 + 120.0 * p.z * (1.0 + 0.5 * wave(3.0 * p.x + p.z) * wave(2.0 * p.y + 0.37 * k));
  const double o = dmax(0.0, wave(1.3 * p.x + 0.21 * k) + wave(1.7 * p.y + 0.4) - 1.3);
  return dmax(0.0, a - 90.0 * o);
}

SYN_HD double scpx_of(const Cfg& c, double x, double y) {
  return c.dx0 * (0.25 + 3.0 * y * (1.0 - y)) * (1.0 + 0.05 * wave(2.0 * x));
}
SYN_HD double scpy_of(const Cfg& c, double x, double y) {
  return c.dy0 * (1.0 + 0.03 * wave(3.0 * y + x));
}

// wrap a global index into 1..n for periodic directions; returns 0 when the
// point lies outside a closed domain
SYN_HD int wrap(int g, int n, bool periodic) {
  if (g >= 1 && g <= n) return g;
  if (!periodic) return 0;
  int w = (g - 1) % n;
  if (w < 0) w += n;
  return w + 1;
}

SYN_HD bool is_sea(const Cfg& c, const uint8_t* sea, int ig, int jg) {
  const bool per_i = !(c.nreg == 0 || c.nreg == 4), per_j = c.nreg > 2;
  const int wi = wrap(ig, c.itdm, per_i), wj = wrap(jg, c.jtdm, per_j);
  if (wi == 0 || wj == 0) return false;
  return sea[(size_t)(wj - 1) * c.itdm + (wi - 1)] != 0;
}

// value of `field` at global (ig,jg,k); lev = 0 old time level (slot n),
// 1 centre time level (slot m).  (ig,jg) may lie in a periodic image.
SYN_HD double value(const Cfg& c, const uint8_t* sea, int field, int ktr, int lev, int ig, int jg,
                    int k) {
  const bool per_i = !(c.nreg == 0 || c.nreg == 4), per_j = c.nreg > 2;
  const int wi = wrap(ig, c.itdm, per_i), wj = wrap(jg, c.jtdm, per_j);
  if (wi == 0 || wj == 0) return 0.0;  // beyond a closed edge: vland
  const XYZ p = coords(c, wi, wj, k);
  switch (field) {
    case F_SCPX: return scpx_of(c, p.x, p.y);
    case F_SCPY: return scpy_of(c, p.x, p.y);
    case F_SCUX: return scpx_of(c, p.x - 0.5 / c.itdm, p.y);
    case F_SCUY: return scpy_of(c, p.x - 0.5 / c.itdm, p.y);
    case F_SCVX: return scpx_of(c, p.x, p.y - 0.5 / c.jtdm);
    case F_SCVY: return scpy_of(c, p.x, p.y - 0.5 / c.jtdm);
    case F_ONETA: return 1.0 + 0.002 * wave(2.0 * p.x + 3.0 * p.y + 0.25 * lev);
    default: break;
  }
  const bool sea_c = sea[(size_t)(wj - 1) * c.itdm + (wi - 1)] != 0;
  if (field == F_UFLX || field == F_VFLX) {
    // mass fluxes: zero on land faces, Courant-like
    // amplitude <= ~0.15 of the smooth face thickness, both signs
    const int di = (field == F_UFLX) ? 1 : 0, dj = 1 - di;
    if (!sea_c || !is_sea(c, sea, wi - di, wj - dj)) return 0.0;
    const int ni = wrap(wi - di, c.itdm, per_i), nj = wrap(wj - dj, c.jtdm, per_j);
    const double hf = 0.5 * (hsmooth(c, wi, wj, k) + hsmooth(c, ni, nj, k));
    double cu;
    if (field == F_UFLX)
      cu = 0.10 * wave(2.5 * p.y + 0.13 * k + 0.25) * wave(1.5 * p.x + 0.3) +
           0.04 * wave(7.0 * p.x + 5.0 * p.y + 0.1 * k) + 0.01 * (2.0 * u01(c.seed, 40, wi, wj, k) - 1.0);
    else
      cu = 0.10 * wave(2.0 * p.x + 0.17 * k) * wave(1.5 * p.y + 0.55) +
           0.04 * wave(5.0 * p.x - 6.0 * p.y + 0.2 * k) + 0.01 * (2.0 * u01(c.seed, 41, wi, wj, k) - 1.0);
    const double area = scpx_of(c, p.x, p.y) * scpy_of(c, p.x, p.y);
    return cu * (area / c.delt1) * (kOnem * hf);
  }
  if (!sea_c) return kHugel;  // land cells hold a sentinel that must never be read
  switch (field) {
    case F_DP: {
      double h = hsmooth(c, wi, wj, k) * (1.0 + 0.1 * (2.0 * u01(c.seed, 30, wi, wj, k) - 1.0));
      if (u01(c.seed, 31, wi, wj, k) < 0.01) h = h * 0.02;  // thin: dp+flxdiv may go negative
      return kOnem * h;
    }
    case F_TEMP: {
      double t = 2.0 + 26.0 * (1.0 - p.z) * (0.5 + 0.5 * wave(0.9 * p.y + 0.1)) +
                 1.5 * wave(6.0 * p.x + 4.0 * p.y + 0.3 * k);
      if (wave(2.0 * p.x + 3.0 * p.y + 0.11 * k) > 0.2) t = t + 3.0;  // front
      if (lev == 1) t = t + 0.2 * (2.0 * u01(c.seed, 11, wi, wj, k) - 1.0);
      return t - 2.0;
    }
    case F_SALN: {
      double s = 34.0 + 3.0 * wave(0.7 * p.x + 1.3 * p.y + 0.1 * k) + 0.02 * (2.0 * u01(c.seed, 20, wi, wj, k) - 1.0);
      if (wave(9.0 * p.x - 2.0 * p.y) > 0.0) s = s + 0.5;
      if (lev == 1) s = s + 0.02 * (2.0 * u01(c.seed, 21, wi, wj, k) - 1.0);
      return s;
    }
    case F_TH3D: {
      double d = 25.0 + 2.0 * p.z + 0.8 * wave(0.7 * p.x + 1.3 * p.y + 0.1 * k) - 0.5 * wave(0.9 * p.y + 0.1);
      if (lev == 1) d = d + 0.01 * (2.0 * u01(c.seed, 23, wi, wj, k) - 1.0);
      return d;
    }
    case F_TRACER: {
      const double q = wave((1.0 + 0.3 * ktr) * p.x + (2.0 - 0.2 * ktr) * p.y + 0.1 * k);
      double tr = (q > 0.3) ? 0.5 + 0.5 * u01(c.seed, 50 + ktr, wi, wj, k) : 0.0;  // step fronts
      if (lev == 1 && tr > 0.0) tr = tr * (1.0 - 0.01 * u01(c.seed, 70 + ktr, wi, wj, k));
      return tr;
    }
    default: break;
  }
  return 0.0;
}

// value for local (i,j) of a tile (Fortran indices, halo included)
// halo_mode 0: cells outside 1..ii x 1..jj receive `fill`; 1: the global function
SYN_HD double tile_value(const Cfg& c, const Tile& t, const uint8_t* sea, int field, int ktr,
                         int lev, int i, int j, int k, int halo_mode, double fill) {
  const bool interior = (i >= 1 && i <= t.ii && j >= 1 && j <= t.jj);
  if (!interior && halo_mode == 0) return fill;
  return value(c, sea, field, ktr, lev, t.i0 + i, t.j0 + j, k);
}

}  // namespace synth
