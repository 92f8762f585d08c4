/*
 * oracle/tsadvc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * PARITY PINNED AGAINST THE REFERENCE'S SOURCE TEXT executed by oracle/fortran_exec.py
 * (see tsadvc_oracle.h; no reference binary: no Fortran compiler exists here).
 *
 * Sweep-by-sweep C99 restatement of /root/reference/mod_tsadvc.F90 and the
 * pieces of bigrid.F90 / mod_xc_sm.h / mod_xc_mp.h / geopar.F90 that path
 * needs.  Every sweep keeps the Fortran loop ranges ("margin"), the Fortran
 * operation order (left-to-right, parentheses honoured), the module scratch
 * slabs and the OpenMP schedule (static,jblk over j).  Build with
 * -ffp-contract=off for the parity library (unfused IEEE double arithmetic);
 * the timed CPU baseline is the same source with -O2 -march=native -fopenmp
 * (config/xc40-gnu-relo_omp:22 analogue).
 *
 * All reals are double (-fdefault-real-8, config/generic-gnu-relo_one:22), all
 * integers 32-bit.
 */
#include "tsadvc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX3(a, b, c) MAX2(MAX2(a, b), c)
#define MIN3(a, b, c) MIN2(MIN2(a, b), c)
#define MAX5(a, b, c, d, e) MAX2(MAX2(MAX2(MAX2(a, b), c), d), e)
#define MIN5(a, b, c, d, e) MIN2(MIN2(MIN2(MIN2(a, b), c), d), e)

/* Fortran (i,j) with lower bounds 1-nbdy on a slab of leading dimension ld */
#define IX(i, j) ((size_t)((i) + nb - 1) + ld * (size_t)((j) + nb - 1))

#define GEOM(t)                                   \
  const int nb = (t)->nbdy;                       \
  const size_t ld = (size_t)((t)->idm + 2 * nb);  \
  const int ii = (t)->ii, jj = (t)->jj;           \
  (void)ii; (void)jj; (void)ld

static char g_err[256] = "";
const char *orc_last_error(void) { return g_err; }

/* ---- taps ---------------------------------------------------------------- */
#define ORC_NTAP 32
static struct { char tag[16]; double *buf; } g_tap[ORC_NTAP];
static int g_ntap = 0;
void orc_clear_taps(void) { g_ntap = 0; }
void orc_set_tap(const char *tag, double *buf) {
  for (int q = 0; q < g_ntap; q++)
    if (!strcmp(g_tap[q].tag, tag)) { g_tap[q].buf = buf; return; }
  if (g_ntap < ORC_NTAP) {
    strncpy(g_tap[g_ntap].tag, tag, 15);
    g_tap[g_ntap].tag[15] = 0;
    g_tap[g_ntap].buf = buf;
    g_ntap++;
  }
}
static void tap(const orc_tile *t, const char *tag, const double *a) {
  for (int q = 0; q < g_ntap; q++)
    if (g_tap[q].buf && !strcmp(g_tap[q].tag, tag))
      memcpy(g_tap[q].buf, a, sizeof(double) * (size_t)orc_slab(t));
}

/* ---- allocation ---------------------------------------------------------- */
int64_t orc_slab(const orc_tile *t) {
  return (int64_t)(t->idm + 2 * t->nbdy) * (int64_t)(t->jdm + 2 * t->nbdy);
}

static double *alloc_r(size_t n) {
  /* mod_tsadvc.F90:130-145, mod_dimensions.F90:265-289: r_init = quiet NaN */
  double *p = (double *)malloc(sizeof(double) * n);
  if (p) for (size_t q = 0; q < n; q++) p[q] = NAN;
  return p;
}
static int *alloc_i(size_t n) { return (int *)calloc(n, sizeof(int)); }

orc_tile *orc_tile_create(int idm, int jdm, int kdm, int nbdy, int ii, int jj,
                          int i0, int j0, int itdm, int jtdm, int nreg,
                          int ntracr) {
  orc_tile *t = (orc_tile *)calloc(1, sizeof(orc_tile));
  if (!t) return NULL;
  t->idm = idm; t->jdm = jdm; t->kdm = kdm; t->nbdy = nbdy; t->ms = 0;
  t->ii = ii; t->jj = jj; t->kk = kdm; t->i0 = i0; t->j0 = j0;
  t->itdm = itdm; t->jtdm = jtdm; t->nreg = nreg; t->mnproc = 1;
  t->ntracr = ntracr;
  size_t P = (size_t)orc_slab(t), K = (size_t)kdm;
  t->ip = alloc_i(P); t->iu = alloc_i(P); t->iv = alloc_i(P); t->iq = alloc_i(P);
  t->ipim1 = alloc_i(P); t->ipip1 = alloc_i(P);
  t->ipjm1 = alloc_i(P); t->ipjp1 = alloc_i(P);
  t->scp2 = alloc_r(P); t->scp2i = alloc_r(P); t->scuy = alloc_r(P);
  t->scvx = alloc_r(P); t->aspux = alloc_r(P); t->aspvy = alloc_r(P);
  t->temp = alloc_r(P * K * 2); t->saln = alloc_r(P * K * 2);
  t->th3d = alloc_r(P * K * 2); t->dp = alloc_r(P * K * 2);
  t->tracer = ntracr > 0 ? alloc_r(P * K * 2 * (size_t)ntracr) : NULL;
  t->uflx = alloc_r(P * K); t->vflx = alloc_r(P * K);
  t->theta = alloc_r(P * K);
  /* mod_asselin.F90 operands */
  t->dpo = alloc_r(P * K * 2); t->onetao = alloc_r(P * 2); t->pbavg = alloc_r(P * 3); t->pbot = alloc_r(P);
  t->otemp = alloc_r(P * K); t->osaln = alloc_r(P * K); t->oth3d = alloc_r(P * K);
  t->otracer = ntracr > 0 ? alloc_r(P * K * (size_t)ntracr) : NULL;
  t->oq2 = alloc_r(P * (K + 2)); t->oq2l = alloc_r(P * (K + 2));
  t->ra2fac = 0.125; t->oneta0 = 0.01;
  t->q2 = alloc_r(P * (K + 2) * 2); t->q2l = alloc_r(P * (K + 2) * 2); /* (0:kk+1,2), mod_cb_arrays.F90:513-514 */
  t->oneta = alloc_r(P * 2); t->onetamas = alloc_r(P * 2);
  t->uflux = alloc_r(P); t->vflux = alloc_r(P);
  t->uflux2 = alloc_r(P); t->vflux2 = alloc_r(P);
  t->util1 = alloc_r(P); t->util2 = alloc_r(P);
  t->fmx = alloc_r(P); t->fmn = alloc_r(P); t->flx = alloc_r(P);
  t->fly = alloc_r(P); t->fldlo = alloc_r(P); t->fmxlo = alloc_r(P);
  t->fmnlo = alloc_r(P); t->fax = alloc_r(P); t->fay = alloc_r(P);
  t->rp = alloc_r(P); t->rm = alloc_r(P); t->flxdiv = alloc_r(P);
  t->tx1 = alloc_r(P); t->ty1 = alloc_r(P); t->fldao = alloc_r(P);
  t->fldan = alloc_r(P);
  t->xmin = alloc_r(K); t->xmax = alloc_r(K);
  /* blkdat defaults of the benchmark configuration */
  t->advtyp = 2; t->advflg = 0; t->btrmas = 0; t->nhybrd = kdm; t->hybrid = 1; t->mxlkta = 0;
  t->isopyc = 0; t->mxlmy = 0; t->nstep = 1; t->diagno = 0;
  t->delt1 = 480.0; t->temdf2 = 0.0; t->temdfc = 1.0; t->thbase = 34.0;
  t->onemm = 9806.0 * 0.001; /* mod_cb_arrays.F90:842-857 */
  t->sigver = 6; /* -DEOS_SIG2 -DEOS_17T, the GLB build */
  return t;
}

void orc_tile_destroy(orc_tile *t) {
  if (!t) return;
  void *ptrs[] = {t->ip, t->iu, t->iv, t->iq, t->ipim1, t->ipip1, t->ipjm1,
                  t->ipjp1, t->ifp, t->ilp, t->isp, t->jfp, t->jlp, t->jsp,
                  t->scp2, t->scp2i, t->scuy, t->scvx, t->aspux, t->aspvy,
                  t->temp, t->saln, t->th3d, t->dp, t->tracer, t->uflx, t->vflx,
                  t->oneta, t->onetamas, t->uflux, t->vflux, t->uflux2,
                  t->vflux2, t->util1, t->util2, t->fmx, t->fmn, t->flx, t->fly,
                  t->fldlo, t->fmxlo, t->fmnlo, t->fax, t->fay, t->rp, t->rm,
                  t->flxdiv, t->tx1, t->ty1, t->fldao, t->fldan, t->xmin,
                  t->xmax, t->theta, t->q2, t->q2l, t->dpo, t->onetao, t->pbavg, t->pbot,
                  t->otemp, t->osaln, t->oth3d, t->otracer, t->oq2, t->oq2l,
                  t->u, t->v, t->dpu, t->dpv, t->ubavg, t->vbavg, t->depthu, t->depthv, t->p, t->utotn, t->vtotn,
                  t->utotm, t->vtotm, t->util3, t->dpmixl, t->dpmold, t->uflxav, t->vflxav, t->dpav, t->dpkmin,
                  t->thkdf4u, t->thkdf4v, t->pold};
  for (size_t q = 0; q < sizeof(ptrs) / sizeof(ptrs[0]); q++) free(ptrs[q]);
  free(t);
}

double *orc_f64(orc_tile *t, const char *name) {
#define F(n) if (!strcmp(name, #n)) return t->n
  F(scp2); F(scp2i); F(scuy); F(scvx); F(aspux); F(aspvy);
  F(temp); F(saln); F(th3d); F(dp); F(tracer); F(uflx); F(vflx);
  F(oneta); F(onetamas); F(uflux); F(vflux); F(uflux2); F(vflux2);
  F(util1); F(util2);
  F(fmx); F(fmn); F(flx); F(fly); F(fldlo); F(fmxlo); F(fmnlo); F(fax); F(fay);
  F(rp); F(rm); F(flxdiv); F(tx1); F(ty1); F(fldao); F(fldan);
  F(xmin); F(xmax); F(theta); F(q2); F(q2l);
  F(dpo); F(onetao); F(pbavg); F(pbot); F(otemp); F(osaln); F(oth3d); F(otracer); F(oq2); F(oq2l);
  F(u); F(v); F(dpu); F(dpv); F(ubavg); F(vbavg); F(depthu); F(depthv); F(p); F(utotn); F(vtotn); F(utotm); F(vtotm);
  F(util3); F(dpmixl); F(dpmold); F(uflxav); F(vflxav); F(dpav); F(dpkmin); F(thkdf4u); F(thkdf4v);
#undef F
  return NULL;
}
int *orc_i32(orc_tile *t, const char *name) {
#define F(n) if (!strcmp(name, #n)) return t->n
  F(ip); F(iu); F(iv); F(iq); F(ipim1); F(ipip1); F(ipjm1); F(ipjp1);
  F(ifp); F(ilp); F(isp); F(jfp); F(jlp); F(jsp);
#undef F
  if (!strcmp(name, "trcflg")) return t->trcflg;
  return NULL;
}

int orc_set_i(orc_tile *t, const char *name, int v) {
#define S(n) if (!strcmp(name, #n)) { t->n = v; return 0; }
  S(advtyp) S(advflg) S(btrmas) S(nhybrd) S(hybrid) S(isopyc) S(mxlmy) S(ntracr) S(mxlkta)
  S(nstep) S(diagno) S(nreg) S(nthreads) S(kk) S(sigver)
#undef S
  return 1;
}
int orc_get_i(const orc_tile *t, const char *name) {
#define G(n) if (!strcmp(name, #n)) return t->n;
  G(advtyp) G(advflg) G(btrmas) G(nhybrd) G(hybrid) G(isopyc) G(mxlmy) G(ntracr) G(mxlkta)
  G(nstep) G(diagno) G(nreg) G(nthreads) G(kk) G(ms) G(xminmax_valid) G(sigver)
  G(idm) G(jdm) G(kdm) G(nbdy) G(ii) G(jj) G(i0) G(j0) G(itdm) G(jtdm)
#undef G
  return -999999;
}
int orc_set_d(orc_tile *t, const char *name, double v) {
#define S(n) if (!strcmp(name, #n)) { t->n = v; return 0; }
  S(delt1) S(temdf2) S(temdfc) S(thbase) S(onemm) S(ra2fac) S(oneta0) S(thkdf2) S(thkdf4)
#undef S
  return 1;
}

/* ---- xctilr, single tile: mod_xc_sm.h:1337-1428 --------------------------- */
void orc_xctilr(const orc_tile *t, double *a, int l1, int ld_, int mh, int nh) {
  orc_xctilr_type(t, a, l1, ld_, mh, nh, 1); /* halo_ps */
}

/* itype as mod_xc.F90:41-44 (halo_ps=1, halo_qs=2, halo_us=3, halo_vs=4, +10: vector field);
 * it only matters on a global grid that includes the arctic (nreg=2, mod_xc_sm.h:1172-1335) */
void orc_xctilr_type(const orc_tile *t, double *a, int l1, int ld_, int mh, int nh, int itype) {
  GEOM(t);
  const size_t P = (size_t)orc_slab(t);
  const double vland = 0.0; /* mod_xc.F90:37, set to 0.0 by xcspmd */
  const int mhl = MAX2(0, MIN2(mh, nb)), nhl = MAX2(0, MIN2(nh, nb));
  if (nhl > 0) {
    if (t->nreg == 2) { /* arctic: mod_xc_sm.h:1215-1320 */
      for (int k = l1; k <= ld_; k++) {
        double *ak = a + P * (size_t)(k - 1);
        for (int j = 1; j <= nhl; j++)
          for (int i = 1; i <= ii; i++) ak[IX(i, 1 - j)] = vland; /* southern boundary is closed */
        const int grid = itype % 10;
        const double sgn = itype < 10 ? 1.0 : -1.0; /* vector field, swap sign */
        for (int j = 1; j <= nhl; j++)
          for (int i = 1; i <= ii; i++) {
            int io, jo;
            if (grid == 1) { io = ii - (i - 1) % ii; jo = jj - 1 - j; }            /* p-grid */
            else if (grid == 2) { io = (ii - (i - 1)) % ii + 1; jo = jj - j; }     /* q-grid */
            else if (grid == 3) { io = (ii - (i - 1)) % ii + 1; jo = jj - 1 - j; } /* u-grid */
            else { io = ii - (i - 1) % ii; jo = jj - j; }                          /* v-grid */
            ak[IX(i, jj + j)] = itype < 10 ? ak[IX(io, jo)] : sgn * ak[IX(io, jo)];
          }
      }
    } else if (t->nreg <= 2) { /* closed in latitude, :1381-1389 */
      for (int k = l1; k <= ld_; k++) {
        double *ak = a + P * (size_t)(k - 1);
        for (int j = 1; j <= nhl; j++)
          for (int i = 1; i <= ii; i++) {
            ak[IX(i, 1 - j)] = vland;
            ak[IX(i, jj + j)] = vland;
          }
      }
    } else { /* periodic (f-plane) in latitude, :1390-1398 */
      for (int k = l1; k <= ld_; k++) {
        double *ak = a + P * (size_t)(k - 1);
        for (int j = 1; j <= nhl; j++)
          for (int i = 1; i <= ii; i++) {
            ak[IX(i, 1 - j)] = ak[IX(i, jj + 1 - j)];
            ak[IX(i, jj + j)] = ak[IX(i, j)];
          }
      }
    }
  }
  if (mhl > 0) {
    if (t->nreg == 0 || t->nreg == 4) { /* closed in longitude, :1403-1411 */
      for (int k = l1; k <= ld_; k++) {
        double *ak = a + P * (size_t)(k - 1);
        for (int j = 1 - nhl; j <= jj + nhl; j++)
          for (int i = 1; i <= mhl; i++) {
            ak[IX(1 - i, j)] = vland;
            ak[IX(ii + i, j)] = vland;
          }
      }
    } else { /* periodic in longitude, :1412-1420 */
      for (int k = l1; k <= ld_; k++) {
        double *ak = a + P * (size_t)(k - 1);
        for (int j = 1 - nhl; j <= jj + nhl; j++)
          for (int i = 1; i <= mhl; i++) {
            ak[IX(1 - i, j)] = ak[IX(ii + 1 - i, j)];
            ak[IX(ii + i, j)] = ak[IX(i, j)];
          }
      }
    }
  }
}

/* ---- xctilr over ipr x jpr tiles: mod_xc_mp.h:4664-4987 ------------------- */
/* Uniform tilings only (every tile row has the same i-splits), which is what
 * the synthetic benchmark uses; N/S first over i=1..ii (:4757-4883), then E/W
 * over j=1-nhl..jj+nhl so that corners propagate (:4885-4978).  A missing
 * neighbour leaves vland (:4786-4787, :4914-4915). */
void orc_world_xctilr(int ipr, int jpr, orc_tile *const *tiles,
                      double *const *a, int l1, int ld_, int mh, int nh) {
  orc_world_xctilr_type(ipr, jpr, tiles, a, l1, ld_, mh, nh, 1); /* halo_ps */
}

/* The ARCTIC version (mod_xc_mp.h:4114-4662) differs for the tiles of the top row (nproc = jpr):
 * their northern neighbour is the twin tile idproc(ipr+1-m,jpr) (:2830), which sends its rows
 * below jj mirrored in i, with the sign flipped for vector fields unless the value is vland
 * (:4263-4372).  On the u and q grids the mirror is shifted by one column (io = ii+2-i), so the
 * first column arrives separately from tile mod(ipr+1-mproc,ipr)+1 (buffer aia, :4319-4331,
 * :4419-4428, :4521-4532). */
void orc_world_xctilr_type(int ipr, int jpr, orc_tile *const *tiles,
                           double *const *a, int l1, int ld_, int mh, int nh, int itype) {
  const double vland = 0.0;
  const int nreg = tiles[0]->nreg;
  const int per_i = !(nreg == 0 || nreg == 4), per_j = nreg > 2;
  const int grid = itype % 10;
  const double sarc = itype < 10 ? 1.0 : -1.0; /* :4211-4215 */
  /* phase 1: north/south */
  for (int n = 0; n < jpr; n++)
    for (int m = 0; m < ipr; m++) {
      const orc_tile *t = tiles[m + ipr * n];
      GEOM(t);
      const size_t P = (size_t)orc_slab(t);
      const int nhl = MAX2(0, MIN2(nh, nb));
      int ns = n - 1, nn = n + 1;
      if (ns < 0) ns = per_j ? jpr - 1 : -1;
      if (nn >= jpr) nn = per_j ? 0 : -1;
      const orc_tile *ts = ns >= 0 ? tiles[m + ipr * ns] : NULL;
      const orc_tile *tn = nn >= 0 ? tiles[m + ipr * nn] : NULL;
      for (int k = l1; k <= ld_; k++) {
        double *ak = a[m + ipr * n] + P * (size_t)(k - 1);
        for (int j = 1; j <= nhl; j++)
          for (int i = 1; i <= ii; i++) {
            double vs = vland, vn = vland;
            if (ts) {
              const size_t lds = (size_t)(ts->idm + 2 * nb);
              const double *as = a[m + ipr * ns] +
                                 (size_t)orc_slab(ts) * (size_t)(k - 1);
              vs = as[(size_t)(i + nb - 1) + lds * (size_t)(ts->jj + 1 - j + nb - 1)];
            }
            if (tn) {
              const size_t ldn = (size_t)(tn->idm + 2 * nb);
              const double *an = a[m + ipr * nn] +
                                 (size_t)orc_slab(tn) * (size_t)(k - 1);
              vn = an[(size_t)(i + nb - 1) + ldn * (size_t)(j + nb - 1)];
            } else if (nreg == 2 && n == jpr - 1) { /* arctic */
              /* what the twin packed for this i (the twin's own loop index is i as well: its
               * ai(l,1) for index i holds a(io,jo) of the twin, :4263-4372) */
              int mt = ipr - 1 - m, io, jo;
              if (grid == 1 || grid == 4) io = ii + 1 - i; /* p,v: ii:1:-1   */
              else io = ii + 2 - i;                        /* u,q: ii+1:2:-1 */
              jo = (grid == 1 || grid == 3) ? jj - 1 - j : jj - j;
              if (io == ii + 1) { /* aia: a(1,jo) of tile mod(ipr+1-mproc,ipr)+1 */
                mt = (ipr - m) % ipr;
                io = 1;
              }
              const orc_tile *tt = tiles[mt + ipr * n];
              const size_t ldt = (size_t)(tt->idm + 2 * nb);
              const double *at = a[mt + ipr * n] + (size_t)orc_slab(tt) * (size_t)(k - 1);
              const double v = at[(size_t)(io + nb - 1) + ldt * (size_t)(jo + nb - 1)];
              vn = (v != vland) ? sarc * v : vland;
            }
            ak[IX(i, 1 - j)] = vs;
            ak[IX(i, jj + j)] = vn;
          }
      }
    }
  /* phase 2: east/west, including the halo rows just filled */
  for (int n = 0; n < jpr; n++)
    for (int m = 0; m < ipr; m++) {
      const orc_tile *t = tiles[m + ipr * n];
      GEOM(t);
      const size_t P = (size_t)orc_slab(t);
      const int mhl = MAX2(0, MIN2(mh, nb)), nhl = MAX2(0, MIN2(nh, nb));
      int mw = m - 1, me = m + 1;
      if (mw < 0) mw = per_i ? ipr - 1 : -1;
      if (me >= ipr) me = per_i ? 0 : -1;
      const orc_tile *tw = mw >= 0 ? tiles[mw + ipr * n] : NULL;
      const orc_tile *te = me >= 0 ? tiles[me + ipr * n] : NULL;
      for (int k = l1; k <= ld_; k++) {
        double *ak = a[m + ipr * n] + P * (size_t)(k - 1);
        for (int j = 1 - nhl; j <= jj + nhl; j++)
          for (int i = 1; i <= mhl; i++) {
            double vw = vland, ve = vland;
            if (tw) {
              const size_t ldw = (size_t)(tw->idm + 2 * nb);
              const double *aw = a[mw + ipr * n] +
                                 (size_t)orc_slab(tw) * (size_t)(k - 1);
              vw = aw[(size_t)(tw->ii + 1 - i + nb - 1) + ldw * (size_t)(j + nb - 1)];
            }
            if (te) {
              const size_t lde = (size_t)(te->idm + 2 * nb);
              const double *ae = a[me + ipr * n] +
                                 (size_t)orc_slab(te) * (size_t)(k - 1);
              ve = ae[(size_t)(i + nb - 1) + lde * (size_t)(j + nb - 1)];
            }
            ak[IX(1 - i, j)] = vw;
            ak[IX(ii + i, j)] = ve;
          }
      }
    }
}

/* ---- bigrid.F90 ----------------------------------------------------------- */
static int indx_alloc(orc_tile *t, int ms) {
  const int nb = t->nbdy;
  free(t->ifp); free(t->ilp); free(t->isp);
  free(t->jfp); free(t->jlp); free(t->jsp);
  t->ms = ms;
  t->ifp = alloc_i((size_t)(t->jdm + 2 * nb) * ms);
  t->ilp = alloc_i((size_t)(t->jdm + 2 * nb) * ms);
  t->isp = alloc_i((size_t)(t->jdm + 2 * nb));
  t->jfp = alloc_i((size_t)(t->idm + 2 * nb) * ms);
  t->jlp = alloc_i((size_t)(t->idm + 2 * nb) * ms);
  t->jsp = alloc_i((size_t)(t->idm + 2 * nb));
  return (t->ifp && t->ilp && t->isp && t->jfp && t->jlp && t->jsp) ? 0 : 1;
}

/* bigrid.F90:418-474 ; returns the number of sections needed in any row */
static int indxi(const orc_tile *t, const int *ipt, int *if_, int *il, int *is,
                 int ms) {
  GEOM(t);
  const size_t lj = (size_t)(t->jdm + 2 * nb);
  int need = 0;
  for (int j = 1 - nb; j <= jj + nb; j++) {
    if (is) {
      is[j + nb - 1] = 0;
      for (int k = 1; k <= ms; k++) {
        if_[(size_t)(j + nb - 1) + lj * (k - 1)] = 0;
        il[(size_t)(j + nb - 1) + lj * (k - 1)] = 0;
      }
    }
    int k = 1;
    int last = ipt[IX(1 - nb, j)];
    if (last == 1 && is) if_[(size_t)(j + nb - 1) + lj * (k - 1)] = 1 - nb;
    for (int i = 2 - nb; i <= ii + nb; i++) {
      if (last == 1 && ipt[IX(i, j)] == 0) {
        if (is) il[(size_t)(j + nb - 1) + lj * (k - 1)] = i - 1;
        k = k + 1;
      } else if (last == 0 && ipt[IX(i, j)] == 1) {
        if (is) {
          if (k > ms) return -1;
          if_[(size_t)(j + nb - 1) + lj * (k - 1)] = i;
        }
      }
      last = ipt[IX(i, j)];
    }
    int nsec;
    if (last == 1) {
      if (is) il[(size_t)(j + nb - 1) + lj * (k - 1)] = ii + nb;
      nsec = k;
    } else {
      nsec = k - 1;
    }
    if (is) is[j + nb - 1] = nsec;
    need = MAX2(need, nsec);
  }
  return need;
}

/* bigrid.F90:476-532 */
static int indxj(const orc_tile *t, const int *jpt, int *jf, int *jl, int *js,
                 int ms) {
  GEOM(t);
  const size_t li = (size_t)(t->idm + 2 * nb);
  int need = 0;
  for (int i = 1 - nb; i <= ii + nb; i++) {
    if (js) {
      js[i + nb - 1] = 0;
      for (int k = 1; k <= ms; k++) {
        jf[(size_t)(i + nb - 1) + li * (k - 1)] = 0;
        jl[(size_t)(i + nb - 1) + li * (k - 1)] = 0;
      }
    }
    int k = 1;
    int last = jpt[IX(i, 1 - nb)];
    if (last == 1 && js) jf[(size_t)(i + nb - 1) + li * (k - 1)] = 1 - nb;
    for (int j = 2 - nb; j <= jj + nb; j++) {
      if (last == 1 && jpt[IX(i, j)] == 0) {
        if (js) jl[(size_t)(i + nb - 1) + li * (k - 1)] = j - 1;
        k = k + 1;
      } else if (last == 0 && jpt[IX(i, j)] == 1) {
        if (js) {
          if (k > ms) return -1;
          jf[(size_t)(i + nb - 1) + li * (k - 1)] = j;
        }
      }
      last = jpt[IX(i, j)];
    }
    int nsec;
    if (last == 1) {
      if (js) jl[(size_t)(i + nb - 1) + li * (k - 1)] = jj + nb;
      nsec = k;
    } else {
      nsec = k - 1;
    }
    if (js) js[i + nb - 1] = nsec;
    need = MAX2(need, nsec);
  }
  return need;
}

int orc_bigrid_stage1(orc_tile *t, double *depth) {
  GEOM(t);
  const int nreg = t->nreg;
  const int lperiod = !(nreg == 0 || nreg == 4);      /* :25-33 via nreg */
  const int lfplane = (nreg == 3 || nreg == 4);       /* :44-45 */
  const int larctic = (nreg == 2);
  if (larctic && !lperiod) { snprintf(g_err, sizeof g_err, "arctic   domain, but non-periodic"); return 2; } /* :48-56 */
  /* :119-154 non-periodic boundaries (part I) */
  if (!lfplane && t->j0 == 0)
    for (int j = 1 - nb; j <= 0; j++)
      for (int i = 1 - nb; i <= ii + nb; i++) depth[IX(i, j)] = 0.0;
  if (!lfplane && !larctic && t->j0 + jj == t->jtdm)
    for (int j = jj + 1; j <= jj + nb; j++)
      for (int i = 1 - nb; i <= ii + nb; i++) depth[IX(i, j)] = 0.0;
  if (!lperiod && t->i0 == 0)
    for (int j = 1 - nb; j <= jj + nb; j++)
      for (int i = 1 - nb; i <= 0; i++) depth[IX(i, j)] = 0.0;
  if (!lperiod && t->i0 + ii == t->itdm)
    for (int j = 1 - nb; j <= jj + nb; j++)
      for (int i = ii + 1; i <= ii + nb; i++) depth[IX(i, j)] = 0.0;
  /* :156-191 single-width inlets / 1-point seas */
  for (int j = 1; j <= jj; j++)
    for (int i = 1; i <= ii; i++)
      if (depth[IX(i, j)] > 0.0) {
        int nzero = 0;
        if (depth[IX(i - 1, j)] <= 0.0) nzero++;
        if (depth[IX(i + 1, j)] <= 0.0) nzero++;
        if (depth[IX(i, j - 1)] <= 0.0) nzero++;
        if (depth[IX(i, j + 1)] <= 0.0) nzero++;
        if (nzero == 4) {
          snprintf(g_err, sizeof g_err,
                   "bigrid: dh(%d,%d) has 4 land neighbours", t->i0 + i,
                   t->j0 + j);
          return 3;
        }
      }
  /* :193-214 */
  const size_t P = (size_t)orc_slab(t);
  for (size_t q = 0; q < P; q++) t->ip[q] = t->iq[q] = t->iu[q] = t->iv[q] = 0;
  for (int j = 1 - nb; j <= jj + nb; j++)
    for (int i = 1 - nb; i <= ii + nb; i++)
      if (depth[IX(i, j)] > 0.) t->ip[IX(i, j)] = 1;
  /* :216-240 */
  for (size_t q = 0; q < P; q++) t->util1[q] = t->util2[q] = t->uflux[q] = 0.0;
  for (int j = 1; j <= jj; j++)
    for (int i = 1; i <= ii; i++) {
      if (t->ip[IX(i - 1, j)] > 0 && t->ip[IX(i, j)] > 0) t->iu[IX(i, j)] = 1;
      if (t->ip[IX(i, j - 1)] > 0 && t->ip[IX(i, j)] > 0) t->iv[IX(i, j)] = 1;
      if (MIN2(MIN2(t->ip[IX(i, j)], t->ip[IX(i - 1, j)]),
               MIN2(t->ip[IX(i, j - 1)], t->ip[IX(i - 1, j - 1)])) > 0)
        t->iq[IX(i, j)] = 1;
      else if ((t->ip[IX(i, j)] > 0 && t->ip[IX(i - 1, j - 1)] > 0) ||
               (t->ip[IX(i - 1, j)] > 0 && t->ip[IX(i, j - 1)] > 0))
        t->iq[IX(i, j)] = 1;
      t->util1[IX(i, j)] = t->iu[IX(i, j)];
      t->util2[IX(i, j)] = t->iv[IX(i, j)];
      t->uflux[IX(i, j)] = t->iq[IX(i, j)]; /* util3 in the reference */
    }
  return 0;
}

int orc_bigrid_stage2(orc_tile *t) {
  GEOM(t);
  const int nreg = t->nreg;
  const int lperiod = !(nreg == 0 || nreg == 4);
  const int lfplane = (nreg == 3 || nreg == 4);
  const int larctic = (nreg == 2);
  /* :244-252 */
  for (int j = 1 - nb; j <= jj + nb; j++)
    for (int i = 1 - nb; i <= ii + nb; i++) {
      t->iu[IX(i, j)] = (int)t->util1[IX(i, j)];
      t->iv[IX(i, j)] = (int)t->util2[IX(i, j)];
      t->iq[IX(i, j)] = (int)t->uflux[IX(i, j)];
    }
  /* :254-297 part II */
  if (!lfplane && t->j0 == 0)
    for (int j = 1 - nb; j <= 0; j++)
      for (int i = 1 - nb; i <= ii + nb; i++)
        t->iq[IX(i, j)] = t->iu[IX(i, j)] = t->iv[IX(i, j)] = 0;
  if (!lfplane && !larctic && t->j0 + jj == t->jtdm)
    for (int j = jj + 1; j <= jj + nb; j++)
      for (int i = 1 - nb; i <= ii + nb; i++)
        t->iq[IX(i, j)] = t->iu[IX(i, j)] = t->iv[IX(i, j)] = 0;
  if (!lperiod && t->i0 == 0)
    for (int j = 1 - nb; j <= jj + nb; j++)
      for (int i = 1 - nb; i <= 0; i++)
        t->iq[IX(i, j)] = t->iu[IX(i, j)] = t->iv[IX(i, j)] = 0;
  if (!lperiod && t->i0 + ii == t->itdm)
    for (int j = 1 - nb; j <= jj + nb; j++)
      for (int i = ii + 1; i <= ii + nb; i++)
        t->iq[IX(i, j)] = t->iu[IX(i, j)] = t->iv[IX(i, j)] = 0;
  /* :316-341 sea-only neighbours (i index for im1/ip1, j index for jm1/jp1) */
  for (int j = 1 - nb + 1; j <= jj + nb - 1; j++)
    for (int i = 1 - nb + 1; i <= ii + nb - 1; i++) {
      t->ipim1[IX(i, j)] = t->ip[IX(i - 1, j)] != 0 ? i - 1 : i;
      t->ipip1[IX(i, j)] = t->ip[IX(i + 1, j)] != 0 ? i + 1 : i;
      t->ipjm1[IX(i, j)] = t->ip[IX(i, j - 1)] != 0 ? j - 1 : j;
      t->ipjp1[IX(i, j)] = t->ip[IX(i, j + 1)] != 0 ? j + 1 : j;
    }
  /* :380-382 segment tables for mass points; ms sized to fit (the reference
   * has a compile-time ms and aborts when it is too small, :454-460) */
  int ms = MAX2(indxi(t, t->ip, NULL, NULL, NULL, 0),
                indxj(t, t->ip, NULL, NULL, NULL, 0));
  if (ms < 1) ms = 1;
  if (indx_alloc(t, ms)) { snprintf(g_err, sizeof g_err, "bigrid: alloc"); return 1; }
  indxi(t, t->ip, t->ifp, t->ilp, t->isp, ms);
  indxj(t, t->ip, t->jfp, t->jlp, t->jsp, ms);
  return 0;
}

int orc_bigrid(orc_tile *t, double *depth) {
  /* :116-117 nreg is defined, so now safe to update halo */
  orc_xctilr(t, depth, 1, 1, t->nbdy, t->nbdy);
  int rc = orc_bigrid_stage1(t, depth);
  if (rc) return rc;
  /* :241-243 */
  orc_xctilr_type(t, t->util1, 1, 1, t->nbdy, t->nbdy, 3); /* halo_us */
  orc_xctilr_type(t, t->util2, 1, 1, t->nbdy, t->nbdy, 4); /* halo_vs */
  orc_xctilr_type(t, t->uflux, 1, 1, t->nbdy, t->nbdy, 2); /* halo_qs */
  return orc_bigrid_stage2(t);
}

/* ---- geopar.F90:311-340 --------------------------------------------------- */
void orc_geopar_metrics(orc_tile *t, const double *scpx, const double *scpy,
                        const double *scux, const double *scuy,
                        const double *scvx, const double *scvy) {
  GEOM(t);
  const double epsil = 1.0e-11; /* mod_cb_arrays.F90:852 */
  const double aspmax = 2.0;    /* geopar.F90: aspmax = 2.0 */
  for (int j = 1 - nb; j <= jj + nb; j++)
    for (int i = 1 - nb; i <= ii + nb; i++) {
      const size_t q = IX(i, j);
      t->scp2[q] = scpx[q] * scpy[q];
      t->scp2i[q] = 1.0 / MAX2(t->scp2[q], epsil);
      t->scuy[q] = scuy[q];
      t->scvx[q] = scvx[q];
      t->aspux[q] = MIN2(MAX2(scux[q], scuy[q]),
                         MIN2(scux[q], scuy[q]) * aspmax) /
                    MAX2(scux[q], epsil);
      t->aspvy[q] = MIN2(MAX2(scvx[q], scvy[q]),
                         MIN2(scvx[q], scvy[q]) * aspmax) /
                    MAX2(scvy[q], epsil);
    }
}

/* ---- advection schemes ---------------------------------------------------- */
#define SEA_P (ip[IX(i, j)] != 0) /* mod_tsadvc.F90:10-12 */
#define SEA_U (iu[IX(i, j)] != 0)
#define SEA_V (iv[IX(i, j)] != 0)

#define MASKS(t)                                         \
  const int *ip = (t)->ip, *iu = (t)->iu, *iv = (t)->iv; \
  (void)ip; (void)iu; (void)iv

#define OMP_J _Pragma("omp parallel for schedule(static, jblk) num_threads(nthr)")

static int jblk_of(const orc_tile *t, int nthr) {
  /* mod_dimensions.F90:143  jblk=(jdm+2*nbdy+mxthrd-1)/mxthrd */
  return (t->jdm + 2 * t->nbdy + nthr - 1) / nthr;
}
static int nthr_of(const orc_tile *t) {
  if (t->nthreads > 0) return t->nthreads;
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* coast zeroing: mod_tsadvc.F90:738-758 (and :301-321, :571-591, :835-855) */
static void coast_zero(const orc_tile *t, double *fx, double *fy, int margin) {
  GEOM(t);
  const size_t lj = (size_t)(t->jdm + 2 * nb), li = (size_t)(t->idm + 2 * nb);
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int l = 1; l <= t->isp[j + nb - 1]; l++) {
      const int f = t->ifp[(size_t)(j + nb - 1) + lj * (l - 1)];
      const int e = t->ilp[(size_t)(j + nb - 1) + lj * (l - 1)];
      if (f >= 1 - margin) fx[IX(f, j)] = 0.0;
      if (e < ii + margin) fx[IX(e + 1, j)] = 0.0;
    }
  for (int i = 1 - margin; i <= ii + margin; i++)
    for (int l = 1; l <= t->jsp[i + nb - 1]; l++) {
      const int f = t->jfp[(size_t)(i + nb - 1) + li * (l - 1)];
      const int e = t->jlp[(size_t)(i + nb - 1) + li * (l - 1)];
      if (f >= 1 - margin) fy[IX(i, f)] = 0.0;
      if (e < jj + margin) fy[IX(i, e + 1)] = 0.0;
    }
}

/* S1 shared by pcm/fct2/fct4 (:686-719, :531-563, :1414-1448): upwind fluxes at
 * iu/iv points, 5-point sea-neighbour extrema at ip points */
static void sweep_upwind_extrema(const orc_tile *t, const double *fld,
                                 const double *u, const double *v, int margin) {
  GEOM(t); MASKS(t);
  double *flx = t->flx, *fly = t->fly, *fmx = t->fmx, *fmn = t->fmn;
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      double q;
      if (SEA_U) {
        if (u[IX(i, j)] >= 0.0) q = fld[IX(i - 1, j)];
        else q = fld[IX(i, j)];
        flx[IX(i, j)] = u[IX(i, j)] * q;
      }
      if (SEA_V) {
        if (v[IX(i, j)] >= 0.0) q = fld[IX(i, j - 1)];
        else q = fld[IX(i, j)];
        fly[IX(i, j)] = v[IX(i, j)] * q;
      }
      if (SEA_P) {
        const int ia = t->ipim1[IX(i, j)], ib = t->ipip1[IX(i, j)];
        const int ja = t->ipjm1[IX(i, j)], jb = t->ipjp1[IX(i, j)];
        fmx[IX(i, j)] = MAX5(fld[IX(i, j)], fld[IX(ia, j)], fld[IX(ib, j)],
                             fld[IX(i, ja)], fld[IX(i, jb)]);
        fmn[IX(i, j)] = MIN5(fld[IX(i, j)], fld[IX(ia, j)], fld[IX(ib, j)],
                             fld[IX(i, ja)], fld[IX(i, jb)]);
      }
    }
}

/* mod_tsadvc.F90:495-643 */
static void advem_pcm(orc_tile *t, double *fld, const double *u,
                      const double *v, const double *fco, const double *fcn,
                      const double *scal, const double *scali, double dt2) {
  GEOM(t); MASKS(t);
  const double onemu = 9806.e-12; /* :519 */
  double *flx = t->flx, *fly = t->fly, *fmx = t->fmx, *fmn = t->fmn,
         *flxdiv = t->flxdiv, *fldao = t->fldao, *fldan = t->fldan;
  const int mbdy_a = 2; /* :527 */
  int margin = mbdy_a - 1;
  sweep_upwind_extrema(t, fld, u, v, margin);
  coast_zero(t, flx, fly, margin);
  tap(t, "ad:22:flx", flx); tap(t, "ad:33:fly", fly);
  margin = mbdy_a - 2; /* :610 */
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        fldao[c] = fld[c] * fco[c] * scal[c];
        flxdiv[c] = ((flx[IX(i + 1, j)] - flx[c]) +
                     (fly[IX(i, j + 1)] - fly[c])) * dt2 * scali[c];
        const double q = fld[c] * (fco[c] + onemu) - flxdiv[c];
        fld[c] = MAX2(fmn[c], MIN2(fmx[c], q / (fcn[c] + onemu)));
        fldan[c] = fld[c] * fcn[c] * scal[c];
      }
  tap(t, "ad:610:flxdv", flxdiv);
}

/* mod_tsadvc.F90:207-493 */
static void advem_mpdata(orc_tile *t, double *fld, const double *u,
                         const double *v, const double *fco, const double *fcn,
                         double posdef, const double *scal,
                         const double *scali, double dt2) {
  GEOM(t); MASKS(t);
  const double onemu = 9806.e-12; /* :236 */
  double *flx = t->flx, *fly = t->fly, *fmx = t->fmx, *fmn = t->fmn,
         *fldlo = t->fldlo, *flxdiv = t->flxdiv, *tx1 = t->tx1, *ty1 = t->ty1,
         *rp = t->rp, *rm = t->rm, *fldao = t->fldao, *fldan = t->fldan;
  const int mbdy_a = 5; /* :241 */
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  int margin = mbdy_a - 1; /* :248 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      double q;
      if (SEA_U) {
        tx1[IX(i, j)] = .5 * fabs(u[IX(i, j)]) * (fld[IX(i, j)] - fld[IX(i - 1, j)]);
        if (u[IX(i, j)] >= 0.0) q = fld[IX(i - 1, j)];
        else q = fld[IX(i, j)];
        flx[IX(i, j)] = u[IX(i, j)] * (q + posdef);
      }
      if (SEA_V) {
        ty1[IX(i, j)] = .5 * fabs(v[IX(i, j)]) * (fld[IX(i, j)] - fld[IX(i, j - 1)]);
        if (v[IX(i, j)] >= 0.0) q = fld[IX(i, j - 1)];
        else q = fld[IX(i, j)];
        fly[IX(i, j)] = v[IX(i, j)] * (q + posdef);
      }
      if (SEA_P) {
        const int ia = t->ipim1[IX(i, j)], ib = t->ipip1[IX(i, j)];
        const int ja = t->ipjm1[IX(i, j)], jb = t->ipjp1[IX(i, j)];
        fmx[IX(i, j)] = MAX5(fld[IX(i, j)], fld[IX(ia, j)], fld[IX(ib, j)],
                             fld[IX(i, ja)], fld[IX(i, jb)]) + posdef;
        fmn[IX(i, j)] = MIN5(fld[IX(i, j)], fld[IX(ia, j)], fld[IX(ib, j)],
                             fld[IX(i, ja)], fld[IX(i, jb)]) + posdef;
      }
    }
  tap(t, "ad:11:tx1", tx1); tap(t, "ad:11:ty1", ty1);
  tap(t, "ad:11:fmx", fmx); tap(t, "ad:11:fmn", fmn);
  coast_zero(t, flx, fly, margin); /* :299-321 */
  tap(t, "ad:22:flx", flx); tap(t, "ad:33:fly", fly);
  margin = mbdy_a - 2; /* :340 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        flxdiv[c] = ((flx[IX(i + 1, j)] - flx[c]) +
                     (fly[IX(i, j + 1)] - fly[c])) * dt2 * scali[c];
        const double q = (fld[c] + posdef) * (fco[c] + onemu) - flxdiv[c];
        fldlo[c] = MAX2(fmn[c], MIN2(fmx[c], q / (fcn[c] + onemu)));
      }
  tap(t, "ad:610:fldlo", fldlo); tap(t, "ad:610:flxdv", flxdiv);
  margin = mbdy_a - 2; /* :369 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      if (SEA_U) {
        const double fco2 = fco[IX(i, j)] + fco[IX(i - 1, j)];
        const double fcn2 = fcn[IX(i, j)] + fcn[IX(i - 1, j)];
        flx[IX(i, j)] = tx1[IX(i, j)] -
                        u[IX(i, j)] * (flxdiv[IX(i, j)] + flxdiv[IX(i - 1, j)]) /
                            ((fco2 + fcn2) + onemu);
      }
      if (SEA_V) {
        const double fco2 = fco[IX(i, j)] + fco[IX(i, j - 1)];
        const double fcn2 = fcn[IX(i, j)] + fcn[IX(i, j - 1)];
        fly[IX(i, j)] = ty1[IX(i, j)] -
                        v[IX(i, j)] * (flxdiv[IX(i, j)] + flxdiv[IX(i, j - 1)]) /
                            ((fco2 + fcn2) + onemu);
      }
    }
  tap(t, "ad: 8:flx", flx); tap(t, "ad: 8:fly", fly);
  margin = mbdy_a - 3; /* :405 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        const double flxdp = MIN2(0.0, flx[IX(i + 1, j)]) - MAX2(0.0, flx[c]);
        const double flxdn = MAX2(0.0, flx[IX(i + 1, j)]) - MIN2(0.0, flx[c]);
        const double flydp = MIN2(0.0, fly[IX(i, j + 1)]) - MAX2(0.0, fly[c]);
        const double flydn = MAX2(0.0, fly[IX(i, j + 1)]) - MIN2(0.0, fly[c]);
        rp[c] = (fmx[c] - fldlo[c]) * (fcn[c] * scal[c]) /
                ((onemu - (flxdp + flydp)) * dt2);
        rm[c] = (fldlo[c] - fmn[c]) * (fcn[c] * scal[c]) /
                ((onemu + (flxdn + flydn)) * dt2);
      }
  tap(t, "ad:16:flp", rp); tap(t, "ad:16:fln", rm);
  margin = mbdy_a - 4; /* :433 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      if (SEA_U) {
        const size_t c = IX(i, j), w = IX(i - 1, j);
        flx[c] = MAX2(0.0, flx[c]) * MIN3(1.0, rp[c], rm[w]) +
                 MIN2(0.0, flx[c]) * MIN3(1.0, rp[w], rm[c]);
      }
      if (SEA_V) {
        const size_t c = IX(i, j), s = IX(i, j - 1);
        fly[c] = MAX2(0.0, fly[c]) * MIN3(1.0, rp[c], rm[s]) +
                 MIN2(0.0, fly[c]) * MIN3(1.0, rp[s], rm[c]);
      }
    }
  tap(t, "ad:18:flx", flx); tap(t, "ad:18:fly", fly);
  margin = mbdy_a - 5; /* :465 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        fldao[c] = fld[c] * fco[c] * scal[c];
        flxdiv[c] = ((flx[IX(i + 1, j)] - flx[c]) +
                     (fly[IX(i, j + 1)] - fly[c])) * dt2 * scali[c];
        fld[c] = MAX2(fmn[c], MIN2(fmx[c], fldlo[c] - flxdiv[c] / (fcn[c] + onemu)));
        fld[c] = fld[c] - posdef;
        fldan[c] = fld[c] * fcn[c] * scal[c];
      }
  tap(t, "ad:620:flxdv", flxdiv);
}

/* mod_tsadvc.F90:645-997 (order=2) and :1370-1706 (order=4); the two differ
 * only in the high-order flux of sweep S3 (:824,:828 vs :1535-1553) */
static void advem_fct(orc_tile *t, int order, double *fld, const double *fldc,
                      const double *u, const double *v, const double *fco,
                      const double *fcn, const double *scal,
                      const double *scali, double dt2) {
  GEOM(t); MASKS(t);
  const double onemu = 9806.e-12;                      /* :671, :1396 */
  const double ft14 = 7.0 / 12.0, ft24 = -1.0 / 12.0;  /* :1398-1399 */
  double *flx = t->flx, *fly = t->fly, *fmx = t->fmx, *fmn = t->fmn,
         *fldlo = t->fldlo, *fmxlo = t->fmxlo, *fmnlo = t->fmnlo,
         *fax = t->fax, *fay = t->fay, *rp = t->rp, *rm = t->rm,
         *flxdiv = t->flxdiv, *fldao = t->fldao, *fldan = t->fldan;
  const int mbdy_a = 5; /* :679 */
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  /* S1 :686-719 */
  int margin = mbdy_a - 1;
  sweep_upwind_extrema(t, fld, u, v, margin);
  tap(t, "ad:11:fmx", fmx); tap(t, "ad:11:fmn", fmn);
  coast_zero(t, flx, fly, margin); /* :738-758 */
  tap(t, "ad:22:flx", flx); tap(t, "ad:33:fly", fly);
  /* S2 :778-797 */
  margin = mbdy_a - 2;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        flxdiv[c] = ((flx[IX(i + 1, j)] - flx[c]) +
                     (fly[IX(i, j + 1)] - fly[c])) * dt2 * scali[c];
        const double q = fld[c] * (fco[c] + onemu) - flxdiv[c];
        fldlo[c] = MAX2(fmn[c], MIN2(fmx[c], q / (fcn[c] + onemu)));
        fmxlo[c] = MAX3(fld[c], fldc[c], fldlo[c]);
        fmnlo[c] = MIN3(fld[c], fldc[c], fldlo[c]);
      }
  tap(t, "ad:610:fldlo", fldlo); tap(t, "ad:610:flxdv", flxdiv);
  tap(t, "ad:610:fmxlo", fmxlo); tap(t, "ad:610:fmnlo", fmnlo);
  /* S3 :817-833 / :1528-1558 */
  margin = mbdy_a - 2;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      double fhx, fhy;
      if (SEA_U) {
        if (order == 2 || iu[IX(i - 1, j)] == 0 || iu[IX(i + 1, j)] == 0)
          fhx = u[IX(i, j)] * 0.5 * (fldc[IX(i, j)] + fldc[IX(i - 1, j)]);
        else
          fhx = u[IX(i, j)] * (ft14 * (fldc[IX(i, j)] + fldc[IX(i - 1, j)]) +
                               ft24 * (fldc[IX(i + 1, j)] + fldc[IX(i - 2, j)]));
        fax[IX(i, j)] = fhx - flx[IX(i, j)];
      }
      if (SEA_V) {
        if (order == 2 || iv[IX(i, j - 1)] == 0 || iv[IX(i, j + 1)] == 0)
          fhy = v[IX(i, j)] * 0.5 * (fldc[IX(i, j)] + fldc[IX(i, j - 1)]);
        else
          fhy = v[IX(i, j)] * (ft14 * (fldc[IX(i, j)] + fldc[IX(i, j - 1)]) +
                               ft24 * (fldc[IX(i, j + 1)] + fldc[IX(i, j - 2)]));
        fay[IX(i, j)] = fhy - fly[IX(i, j)];
      }
    }
  coast_zero(t, fax, fay, margin); /* :835-855 */
  tap(t, "ad:fax0", fax); tap(t, "ad:fay0", fay);
  /* S4 :863-909 */
  margin = mbdy_a - 3;
  const double qdt2 = 1.0 / dt2; /* :865 */
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        const int ia = t->ipim1[c], ib = t->ipip1[c];
        const int ja = t->ipjm1[c], jb = t->ipjp1[c];
        const double fqmax = MAX5(fmxlo[c], fmxlo[IX(ia, j)], fmxlo[IX(ib, j)],
                                  fmxlo[IX(i, ja)], fmxlo[IX(i, jb)]);
        const double fqmin = MIN5(fmnlo[c], fmnlo[IX(ia, j)], fmnlo[IX(ib, j)],
                                  fmnlo[IX(i, ja)], fmnlo[IX(i, jb)]);
        /* note fax(ib,j), fay(i,jb): the sea-only index, not i+1/j+1 (:880-883) */
        const double famax = MAX2(0.0, fax[c]) - MIN2(0.0, fax[IX(ib, j)]) +
                             MAX2(0.0, fay[c]) - MIN2(0.0, fay[IX(i, jb)]);
        const double famin = MAX2(0.0, fax[IX(ib, j)]) - MIN2(0.0, fax[c]) +
                             MAX2(0.0, fay[IX(i, jb)]) - MIN2(0.0, fay[c]);
        if (famax > 0.0) {
          const double qp = (fqmax - fldlo[c]) * fcn[c] * scal[c] * qdt2;
          if (qp < famax) rp[c] = qp / famax;
          else rp[c] = 1.0;
        } else {
          rp[c] = 0.0;
        }
        if (famin > 0.0) {
          const double qm = (fldlo[c] - fqmin) * fcn[c] * scal[c] * qdt2;
          if (qm < famin) rm[c] = qm / famin;
          else rm[c] = 1.0;
        } else {
          rm[c] = 0.0;
        }
        fmx[c] = fqmax;
        fmn[c] = fqmin;
      }
  tap(t, "ad:16:rp", rp); tap(t, "ad:16:rm", rm);
  /* S5 :922-946 */
  margin = mbdy_a - 4;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      double fact;
      if (SEA_U) {
        if (fax[IX(i, j)] < 0.0) fact = MIN2(rp[IX(i - 1, j)], rm[IX(i, j)]);
        else fact = MIN2(rp[IX(i, j)], rm[IX(i - 1, j)]);
        fax[IX(i, j)] = fact * fax[IX(i, j)];
      }
      if (SEA_V) {
        if (fay[IX(i, j)] < 0.0) fact = MIN2(rp[IX(i, j - 1)], rm[IX(i, j)]);
        else fact = MIN2(rp[IX(i, j)], rm[IX(i, j - 1)]);
        fay[IX(i, j)] = fact * fay[IX(i, j)];
      }
    }
  tap(t, "ad:18:fax", fax); tap(t, "ad:18:fay", fay);
  /* S6 :962-982 */
  margin = mbdy_a - 5;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        fldao[c] = fld[c] * fco[c] * scal[c];
        flxdiv[c] = ((fax[IX(i + 1, j)] - fax[c]) +
                     (fay[IX(i, j + 1)] - fay[c])) * dt2 * scali[c];
        fld[c] = MAX2(fmn[c], MIN2(fmx[c], fldlo[c] - flxdiv[c] / (fcn[c] + onemu)));
        fldan[c] = fld[c] * fcn[c] * scal[c];
      }
  tap(t, "ad:620:flxdv", flxdiv);
}


/* advem_fct2c: mod_tsadvc.F90:999-1368 (Baraille 2008: leapfrog FCT2 whose low-order step is
 * sub-cycled, 5 iterations with a local time step so that no cell is emptied; used when
 * btrmas, :96-97).  Scratch uloc..lcalc as :1041-1063; each iteration ends with
 * xctilr(hloc), xctilr(fldlo) (:1186-1187) - single-tile semantics here. */
static void advem_fct2c(orc_tile *t, double *fld, const double *fldc,
                        const double *u, const double *v, const double *fco,
                        const double *fcn, const double *scal,
                        const double *scali, double dt2) {
  GEOM(t); MASKS(t);
  (void)fcn;
  const double epsil = 1.e-10; /* :1027 */
  const size_t P = (size_t)orc_slab(t);
  double *flx = t->flx, *fly = t->fly, *fmx = t->fmx, *fmn = t->fmn,
         *fldlo = t->fldlo, *fax = t->fax, *fay = t->fay, *rp = t->rp,
         *rm = t->rm, *flxdiv = t->flxdiv;
  double *uloc = alloc_r(P), *vloc = alloc_r(P), *hloc = alloc_r(P),
         *dtloc = alloc_r(P), *ucumdt = alloc_r(P), *vcumdt = alloc_r(P),
         *flxcum = alloc_r(P), *flycum = alloc_r(P);
  int *lcalc = alloc_i(P);
  const int mbdy_a = 5; /* :1065 */
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  int margin;
  /* :1072-1086 */
  for (size_t q = 0; q < P; q++) {
    lcalc[q] = 1; flxcum[q] = 0.0; flycum[q] = 0.0; dtloc[q] = 0.0;
    hloc[q] = fco[q]; fldlo[q] = fld[q];
    ucumdt[q] = 0.0; vcumdt[q] = 0.0; uloc[q] = 0.0; vloc[q] = 0.0;
    flx[q] = 0.0; fly[q] = 0.0;
  }
  for (int iter = 1; iter <= 5; iter++) { /* :1088 */
    margin = mbdy_a; /* :1090-1107 */
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_P) {
          const size_t c = IX(i, j);
          const double q = MAX2(u[IX(i + 1, j)], 0.0) - MIN2(u[c], 0.0) +
                           MAX2(v[IX(i, j + 1)], 0.0) - MIN2(v[c], 0.0);
          if (q > 0.0)
            dtloc[c] = MIN2(dt2, hloc[c] / (q * scali[c]));
          else
            dtloc[c] = dt2;
        }
    margin = mbdy_a - 1; /* :1109-1158 */
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int i = 1 - margin; i <= ii + margin; i++) {
        const size_t c = IX(i, j);
        if (SEA_U) {
          if (ucumdt[c] != dt2) {
            if (u[c] >= 0) {
              uloc[c] = MIN2(dt2 - ucumdt[c], dtloc[IX(i - 1, j)]) * u[c];
              flx[c] = fldlo[IX(i - 1, j)] * uloc[c];
              ucumdt[c] = ucumdt[c] + MIN2(dt2 - ucumdt[c], dtloc[IX(i - 1, j)]);
            } else {
              uloc[c] = MIN2(dt2 - ucumdt[c], dtloc[c]) * u[c];
              flx[c] = fldlo[c] * uloc[c];
              ucumdt[c] = ucumdt[c] + MIN2(dt2 - ucumdt[c], dtloc[c]);
            }
            flxcum[c] = flxcum[c] + flx[c];
          } else {
            uloc[c] = 0.0;
            flx[c] = 0.0;
          }
        }
        if (SEA_V) {
          if (vcumdt[c] != dt2) {
            if (v[c] >= 0) {
              vloc[c] = MIN2(dt2 - vcumdt[c], dtloc[IX(i, j - 1)]) * v[c];
              fly[c] = fldlo[IX(i, j - 1)] * vloc[c];
              vcumdt[c] = vcumdt[c] + MIN2(dt2 - vcumdt[c], dtloc[IX(i, j - 1)]);
            } else {
              vloc[c] = MIN2(dt2 - vcumdt[c], dtloc[c]) * v[c];
              fly[c] = fldlo[c] * vloc[c];
              vcumdt[c] = vcumdt[c] + MIN2(dt2 - vcumdt[c], dtloc[c]);
            }
            flycum[c] = flycum[c] + fly[c];
          } else {
            vloc[c] = 0.0;
            fly[c] = 0.0;
          }
        }
      }
    margin = mbdy_a - 2; /* :1160-1184 */
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_P) {
          const size_t c = IX(i, j);
          if (lcalc[c]) {
            const double qp =
                hloc[c] - (uloc[IX(i + 1, j)] - uloc[c] + vloc[IX(i, j + 1)] - vloc[c]) * scali[c];
            if (qp > 0.0)
              fldlo[c] = ((epsil + hloc[c]) * fldlo[c] -
                          (flx[IX(i + 1, j)] - flx[c] + fly[IX(i, j + 1)] - fly[c]) * scali[c]) /
                         (epsil + qp);
            hloc[c] = qp;
            lcalc[c] = ucumdt[IX(i + 1, j)] != dt2 || ucumdt[c] != dt2 ||
                       vcumdt[IX(i, j + 1)] != dt2 || vcumdt[c] != dt2;
          }
        }
    orc_xctilr(t, hloc, 1, 1, mbdy_a, mbdy_a);  /* :1186 */
    orc_xctilr(t, fldlo, 1, 1, mbdy_a, mbdy_a); /* :1187 */
    tap(t, "aditer:fldlo", fldlo); tap(t, "aditer:hloc", hloc);
  }
  /* :1202-1215 high-order minus the cumulated low-order fluxes */
  margin = mbdy_a - 2;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      const size_t c = IX(i, j);
      if (SEA_U) {
        const double fhx = u[c] * 0.5 * (fldc[c] + fldc[IX(i - 1, j)]);
        fax[c] = fhx - flxcum[c] / dt2;
      }
      if (SEA_V) {
        const double fhy = v[c] * 0.5 * (fldc[c] + fldc[IX(i, j - 1)]);
        fay[c] = fhy - flycum[c] / dt2;
      }
    }
  coast_zero(t, fax, fay, margin); /* :1223-1245 */
  tap(t, "ad:ip:0:fax", fax); tap(t, "ad:ip:0:fay", fay);
  /* :1258-1306 */
  margin = mbdy_a - 3;
  const double qdt2 = 1.0 / dt2;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        int ia = i - 1; if (ip[IX(ia, j)] == 0) ia = i;
        int ib = i + 1; if (ip[IX(ib, j)] == 0) ib = i;
        int ja = j - 1; if (ip[IX(i, ja)] == 0) ja = j;
        int jb = j + 1; if (ip[IX(i, jb)] == 0) jb = j;
        const double fqmax = MAX5(fldlo[c], fldlo[IX(ia, j)], fldlo[IX(ib, j)],
                                  fldlo[IX(i, ja)], fldlo[IX(i, jb)]);
        const double fqmin = MIN5(fldlo[c], fldlo[IX(ia, j)], fldlo[IX(ib, j)],
                                  fldlo[IX(i, ja)], fldlo[IX(i, jb)]);
        const double famax = MAX2(0.0, fax[c]) - MIN2(0.0, fax[IX(i + 1, j)]) +
                             MAX2(0.0, fay[c]) - MIN2(0.0, fay[IX(i, j + 1)]);
        const double famin = MAX2(0.0, fax[IX(i + 1, j)]) - MIN2(0.0, fax[c]) +
                             MAX2(0.0, fay[IX(i, j + 1)]) - MIN2(0.0, fay[c]);
        if (famax > epsil) {
          const double qp = (fqmax - fldlo[c]) * hloc[c] * scal[c] * qdt2;
          rp[c] = qp < famax ? qp / famax : 1.0;
        } else {
          rp[c] = 0.0;
        }
        if (famin > epsil) {
          const double qm = (fldlo[c] - fqmin) * hloc[c] * scal[c] * qdt2;
          rm[c] = qm < famin ? qm / famin : 1.0;
        } else {
          rm[c] = 0.0;
        }
        fmx[c] = fqmax;
        fmn[c] = fqmin;
      }
  /* :1311-1335 */
  margin = mbdy_a - 4;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      const size_t c = IX(i, j);
      if (SEA_U) {
        const double fact = fax[c] < 0.0 ? MIN2(rp[IX(i - 1, j)], rm[c])
                                         : MIN2(rp[c], rm[IX(i - 1, j)]);
        fax[c] = fact * fax[c];
      }
      if (SEA_V) {
        const double fact = fay[c] < 0.0 ? MIN2(rp[IX(i, j - 1)], rm[c])
                                         : MIN2(rp[c], rm[IX(i, j - 1)]);
        fay[c] = fact * fay[c];
      }
    }
  /* :1343-1361 */
  margin = mbdy_a - 5;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        flxdiv[c] = ((fax[IX(i + 1, j)] - fax[c]) +
                     (fay[IX(i, j + 1)] - fay[c])) * dt2 * scali[c];
        if (hloc[c] > 0.)
          fld[c] = ((epsil + hloc[c]) * fldlo[c] - flxdiv[c]) / (epsil + hloc[c]);
        else
          fld[c] = fldlo[c];
      }
  tap(t, "ad:fct2c:fld", fld);
  free(uloc); free(vloc); free(hloc); free(dtloc); free(ucumdt); free(vcumdt);
  free(flxcum); free(flycum); free(lcalc);
}

/* mod_tsadvc.F90:69-205 (lconserve is compile-time .false., :32) */
int orc_advem(orc_tile *t, int advtyp, double *fld, const double *fldc,
              const double *u, const double *v, const double *fco,
              const double *fcn, double posdef, const double *scal,
              const double *scali, double dt2, int btrmas) {
  if (advtyp == 0) {
    advem_pcm(t, fld, u, v, fco, fcn, scal, scali, dt2);
  } else if (advtyp == 1) {
    advem_mpdata(t, fld, u, v, fco, fcn, posdef, scal, scali, dt2);
  } else if (advtyp == 2 && btrmas) {
    advem_fct2c(t, fld, fldc, u, v, fco, fcn, scal, scali, dt2);
  } else if (advtyp == 2) {
    advem_fct(t, 2, fld, fldc, u, v, fco, fcn, scal, scali, dt2);
  } else if (advtyp == 4) {
    advem_fct(t, 4, fld, fldc, u, v, fco, fcn, scal, scali, dt2);
  } else {
    snprintf(g_err, sizeof g_err, "error: advem called with advtyp =%4d", advtyp);
    return 5; /* xcstop('advem'), :159-166 */
  }
  return 0;
}


/* ---- equation of state: stmt_fns.h ----------------------------------------
 * The reference picks ONE set of statement functions at compile time (EOS_SIG0/
 * EOS_SIG2 x EOS_7T/9T/12T/17T, stmt_fns.h:2-22 "sigver").  Here sigver is a run
 * time scalar of the tile so that one oracle build covers all eight.  Every
 * expression keeps the Fortran grouping; x**2 = x*x, x**3 = (x*x)*x (gfortran's
 * expansion of small integer powers); parameter expressions (rc6, c101..) are
 * evaluated in double like gfortran's constant folder does. */
typedef struct eos_c79 { double c1, c2, c3, c4, c5, c6, c7, c8, c9; } eos_c79;
static eos_c79 eos_coef79(int sigver) {
  eos_c79 c = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  switch (sigver) {
    case 1: /* stmt_fns.h:53-61 */
      c.c1 = -1.36471E-01; c.c2 = 4.68181E-02; c.c3 = 8.07004E-01; c.c4 = -7.45353E-03;
      c.c5 = -2.94418E-03; c.c6 = 3.43570E-05; c.c7 = 3.48658E-05; break;
    case 2: /* :65-73 */
      c.c1 = 9.77093E+00; c.c2 = -2.26493E-02; c.c3 = 7.89879E-01; c.c4 = -6.43205E-03;
      c.c5 = -2.62983E-03; c.c6 = 2.75835E-05; c.c7 = 3.15235E-05; break;
    case 3: /* :86-96 */
      c.c1 = -4.311829E-02; c.c2 = 5.429948E-02; c.c3 = 8.011774E-01; c.c4 = -7.641336E-03;
      c.c5 = -3.258442E-03; c.c6 = 3.757643E-05; c.c7 = 3.630361E-05; c.c8 = 8.675546E-05;
      c.c9 = 3.995086E-06; break;
    case 4: /* :100-110 */
      c.c1 = 9.903308E+00; c.c2 = -1.618075E-02; c.c3 = 7.819166E-01; c.c4 = -6.593939E-03;
      c.c5 = -2.896464E-03; c.c6 = 3.038697E-05; c.c7 = 3.266933E-05; c.c8 = 1.180109E-04;
      c.c9 = 3.399511E-06; break;
  }
  return c;
}
/* tofsig of the 7- and 9-term fits: root of t**3+a2*t**2+a1*t+a0=0 (:311-323, :348-361) */
static double eos_cubic_root(double a0, double a1, double a2) {
  const double a3rd = 1.0 / 3.0;
  const double x = a3rd * a2;
  const double cubq = a3rd * a1 - x * x;
  const double cubr = a3rd * (0.5 * a1 * a2 - 1.5 * a0) - x * x * x;
  const double cuban =
      a3rd * atan2(sqrt(MAX2(0.0, -(cubq * cubq * cubq + cubr * cubr))), cubr);
  const double cubrl = sqrt(-cubq) * cos(cuban);
  const double cubim = sqrt(-cubq) * sin(cuban);
  return -cubrl + sqrt(3.0) * cubim - a3rd * a2;
}
/* 12-term rational function (:129-172) */
typedef struct eos_c12 {
  double c101, c102, c103, c004, c005, c006, c111, c112, c113, c014, c015, c016;
} eos_c12;
static eos_c12 eos_coef12(int sigver) {
  const double c001 = -1.4627567840659594e-01, c002 = 6.4247392832635697e-02,
               c003 = 8.1213979591704621e-01, c007 = 5.0879498675039621e-03,
               c008 = 1.6333913018305079e-05, c009 = 4.3899924880543972e-06,
               c011 = 1.0000000000000000e+00, c012 = 1.0316374535350838e-02,
               c013 = 8.9521792365142522e-04, c017 = 1.1995545126831476e-05,
               c018 = 5.5234008384648383e-08, c019 = 8.4310335919950873e-09;
  const double prs2pdb = 1.e-4, pref = sigver == 7 ? 0.0 : 2000.e4;
  const double rpdb = pref * prs2pdb;
  eos_c12 c;
  c.c004 = -8.1321489441909698e-03; c.c005 = 4.5199845091090296e-03;
  c.c006 = 4.6347888132781394e-04; c.c014 = -2.8438341552142710e-05;
  c.c015 = -1.1887778959461776e-05; c.c016 = -4.0163964812921489e-06;
  c.c101 = c001 + rpdb * c007; c.c102 = c002 + rpdb * c008; c.c103 = c003 + rpdb * c009;
  c.c111 = c011 + rpdb * c017; c.c112 = c012 + rpdb * c018; c.c113 = c013 + rpdb * c019;
  return c;
}
/* sig(t,s): stmt_fns.h:332 (7T), :368-369 (9T), :419-424 (12T), :503-510 (17T) */
static double eos_sig(int sigver, double t, double s) {
  if (sigver == 1 || sigver == 2) {
    const eos_c79 c = eos_coef79(sigver);
    return (c.c1 + c.c3 * s + t * (c.c2 + c.c5 * s + t * (c.c4 + c.c7 * s + c.c6 * t)));
  }
  if (sigver == 3 || sigver == 4) {
    const eos_c79 c = eos_coef79(sigver);
    return (c.c1 + s * (c.c3 + s * c.c8) +
            t * (c.c2 + s * (c.c5 + s * c.c9) + t * (c.c4 + s * c.c7 + t * c.c6)));
  }
  if (sigver == 7 || sigver == 8) {
    const eos_c12 c = eos_coef12(sigver);
    const double sig_n = c.c101 + (c.c102 + c.c004 * t + c.c005 * s) * t + (c.c103 + c.c006 * s) * s;
    const double sig_d = c.c111 + (c.c112 + c.c014 * t + c.c015 * s) * t + (c.c113 + c.c016 * s) * s;
    return sig_n * (1.0 / sig_d);
  }
  { /* 17-term, :216-290 and :503-510 */
    const double c001 = 9.9984085444849347e+02, c002 = 7.3471625860981584e+00,
                 c003 = -5.3211231792841769e-02, c004 = 3.6492439109814549e-04,
                 c005 = 2.5880571023991390e+00, c006 = 6.7168282786692355e-03,
                 c007 = 1.9203202055760151e-03, c008 = 1.0000000000000000e+00,
                 c009 = 7.2815210113327091e-03, c010 = -4.4787265461983921e-05,
                 c011 = 3.3851002965802430e-07, c012 = 1.3651202389758572e-10,
                 c013 = 1.7632126669040377e-03, c014 = 8.8066583251206474e-06,
                 c015 = 1.8832689434804897e-10, c016 = 5.7463776745432097e-06,
                 c017 = 1.4716275472242334e-09, c018 = 1.1798263740430364e-02,
                 c019 = 9.8920219266399117e-08, c020 = 4.6996642771754730e-06,
                 c021 = 2.5862187075154352e-08, c022 = 3.2921414007960662e-12,
                 c023 = 6.7103246285651894e-06, c024 = 2.4461698007024582e-17,
                 c025 = 9.1534417604289062e-18;
    const double prs2pdb = 1.e-4, pref = sigver == 5 ? 0.0 : 2000.e4;
    const double rpdb = pref * prs2pdb;
    const double c101 = c001 + (c018 - c021 * rpdb) * rpdb, c103 = c003 + (c019 - c022 * rpdb) * rpdb,
                 c105 = c005 + c020 * rpdb, c108 = c008 + c023 * rpdb,
                 c109 = c009 - c025 * (rpdb * rpdb * rpdb), c111 = c011 - c024 * (rpdb * rpdb);
    const double sig_n = c101 + t * (c002 + t * (c103 + t * c004)) + s * (c105 - t * c006 + s * c007);
    const double sig_d = c108 + t * (c109 + t * (c010 + t * (c111 + t * c012))) +
                         s * (c013 - t * (c014 + t * t * c015) +
                              sqrt(MAX2(0.0, s)) * (c016 + t * t * c017));
    return sig_n * (1.0 / sig_d) - 1000.0;
  }
}
/* tofsig(r,s): :323 (7T), :379 (9T), :441-449 (12T), :533 (17T: "NOT AVAILABLE", 99.0) */
static double eos_tofsig(int sigver, double r, double s) {
  if (sigver == 1 || sigver == 2) {
    const eos_c79 c = eos_coef79(sigver);
    const double rc6 = 1.0 / c.c6;
    return eos_cubic_root((c.c1 + c.c3 * s - r) * rc6, (c.c2 + c.c5 * s) * rc6,
                          (c.c4 + c.c7 * s) * rc6);
  }
  if (sigver == 3 || sigver == 4) {
    const eos_c79 c = eos_coef79(sigver);
    const double rc6 = 1.0 / c.c6;
    return eos_cubic_root((c.c1 + s * (c.c3 + s * c.c8) - r) * rc6,
                          (c.c2 + s * (c.c5 + s * c.c9)) * rc6, (c.c4 + s * c.c7) * rc6);
  }
  if (sigver == 7 || sigver == 8) {
    const eos_c12 c = eos_coef12(sigver);
    const double a = (c.c004 - r * c.c014);
    const double b = ((c.c102 + c.c005 * s) - r * (c.c112 + c.c015 * s));
    const double cc = ((c.c101 + (c.c103 + c.c006 * s) * s) - r * (c.c111 + (c.c113 + c.c016 * s) * s));
    return (-b - sqrt(MAX2(0.0, b * b - 4.0 * a * cc))) / (2.0 * a);
  }
  return 99.0;
}
double orc_sig(int sigver, double t, double s) { return eos_sig(sigver, t, s); }
double orc_tofsig(int sigver, double r, double s) { return eos_tofsig(sigver, r, s); }

/* geopar.F90:776-783, :826-843, :851, :862-879: uflux2/vflux2 (and uflux/vflux) are
 * `hugel` everywhere, then 0.0 on every face ifp..ilp+1 of each sea segment, then halo
 * updated: a face that borders a sea cell but is not an iu/iv point is 0.  tsdff_2x reads
 * them there (mod_tsadvc.F90:2316-2321). */
static void geopar_zero_coast_faces(orc_tile *t, double *fx, double *fy) {
  const int nb = t->nbdy;
  const int ncol = t->idm + 2 * nb, nrow = t->jdm + 2 * nb;
  const double hugel = 1.2676506002282294e30; /* 2.0**100, mod_cb_arrays.F90 */
  for (int r = 0; r < nrow; r++)
    for (int c = 0; c < ncol; c++) {
      const size_t q = (size_t)r * ncol + c;
      const int pw = c > 0 ? t->ip[q - 1] : 0, ps = r > 0 ? t->ip[q - ncol] : 0;
      fx[q] = (t->ip[q] != 0 || pw != 0) ? 0.0 : hugel;
      fy[q] = (t->ip[q] != 0 || ps != 0) ? 0.0 : hugel;
    }
}

/* ---- diffusion: mod_tsadvc.F90:2262-2492 ---------------------------------- */
static double harmonc(double aa, double bb) { /* :1770-1771 */
  const double eps_har = 1.0e-20;
  const double a = MAX2(aa, 0.0), b = MAX2(bb, 0.0);
  return 2.0 * a * b / MAX2((a + b), 2.0 * eps_har);
}

static void tsdff(orc_tile *t, int k, int n, double *fld1, double *fld2) {
  GEOM(t); MASKS(t);
  const size_t P = (size_t)orc_slab(t);
  const double eps_har = 1.0e-20;
  const double *dpn = t->dp + P * ((size_t)(k - 1) + (size_t)t->kdm * (n - 1));
  const double *oem = t->onetamas + P * (size_t)(n - 1);
  double *uflux = t->uflux, *vflux = t->vflux, *uflux2 = t->uflux2,
         *vflux2 = t->vflux2, *util1 = t->util1, *util2 = t->util2;
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  int margin = 1;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      if (SEA_U) {
        const double factor =
            t->temdf2 * t->aspux[IX(i, j)] * t->scuy[IX(i, j)] *
            harmonc(dpn[IX(i - 1, j)] * oem[IX(i - 1, j)],
                    dpn[IX(i, j)] * oem[IX(i, j)]);
        uflux[IX(i, j)] = factor * (fld1[IX(i - 1, j)] - fld1[IX(i, j)]);
        if (fld2) uflux2[IX(i, j)] = factor * (fld2[IX(i - 1, j)] - fld2[IX(i, j)]);
      }
      if (SEA_V) {
        const double factor =
            t->temdf2 * t->aspvy[IX(i, j)] * t->scvx[IX(i, j)] *
            harmonc(dpn[IX(i, j - 1)] * oem[IX(i, j - 1)],
                    dpn[IX(i, j)] * oem[IX(i, j)]);
        vflux[IX(i, j)] = factor * (fld1[IX(i, j - 1)] - fld1[IX(i, j)]);
        if (fld2) vflux2[IX(i, j)] = factor * (fld2[IX(i, j - 1)] - fld2[IX(i, j)]);
      }
    }
  margin = 0;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        const double factor =
            -t->delt1 / (t->scp2[c] * MAX2(dpn[c] * oem[c], eps_har));
        util1[c] = ((uflux[IX(i + 1, j)] - uflux[c]) +
                    (vflux[IX(i, j + 1)] - vflux[c])) * factor;
        fld1[c] = fld1[c] + util1[c];
        if (fld2) {
          util2[c] = ((uflux2[IX(i + 1, j)] - uflux2[c]) +
                      (vflux2[IX(i, j + 1)] - vflux2[c])) * factor;
          fld2[c] = fld2[c] + util2[c];
        }
      }
}


/* ---- mod_asselin.F90: Robert-Asselin filter of the scalar fields ---------------------------
 * asselin_save :28-82 (time level t-1 saved, oneta/onetao of both slots, their halos to width 6)
 * asselin_filter :84-286 (margin 0; smooths oneta*dp*scalar, conserving constants) */
void orc_asselin_save(orc_tile *t, int m, int n, int do_halo) {
  GEOM(t); MASKS(t);
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  const int kk = t->kk;
  for (int j = 1; j <= jj; j++) {
    for (int i = 1; i <= ii; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        t->oneta[c + P * (size_t)(n - 1)] = MAX2(t->oneta0, 1.0 + t->pbavg[c + P * (size_t)(n - 1)] / t->pbot[c]);
        t->oneta[c + P * (size_t)(m - 1)] = MAX2(t->oneta0, 1.0 + t->pbavg[c + P * (size_t)(m - 1)] / t->pbot[c]);
        t->onetao[c + P * (size_t)(n - 1)] = t->oneta[c + P * (size_t)(n - 1)];
        t->onetao[c + P * (size_t)(m - 1)] = t->oneta[c + P * (size_t)(m - 1)];
      }
    for (int k = 1; k <= kk; k++)
      for (int i = 1; i <= ii; i++) {
        const size_t c = IX(i, j), ck = c + P * (size_t)(k - 1), cn = ck + P * K * (size_t)(n - 1);
        t->otemp[ck] = t->temp[cn];
        t->osaln[ck] = t->saln[cn];
        t->oth3d[ck] = t->th3d[cn];
        for (int ktr = 1; ktr <= t->ntracr; ktr++)
          t->otracer[ck + P * K * (size_t)(ktr - 1)] = t->tracer[cn + P * K * 2 * (size_t)(ktr - 1)];
      }
    if (t->mxlmy)
      for (int k = 1; k <= kk; k++)
        for (int i = 1; i <= ii; i++) {
          const size_t c = IX(i, j);
          t->oq2[c + P * (size_t)k] = t->q2[c + P * ((size_t)k + (K + 2) * (size_t)(n - 1))];
          t->oq2l[c + P * (size_t)k] = t->q2l[c + P * ((size_t)k + (K + 2) * (size_t)(n - 1))];
        }
  }
  if (do_halo) { /* :77-78 */
    orc_xctilr(t, t->oneta, 1, 2, 6, 6);
    orc_xctilr(t, t->onetao, 1, 2, 6, 6);
  }
}

void orc_asselin_filter(orc_tile *t, int m, int n) {
  GEOM(t); MASKS(t);
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  const int kk = t->kk;
  const double onezm = 9806.e-20; /* :93 */
  const double ra2fac = t->ra2fac;
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  OMP_J
  for (int j = 1; j <= jj; j++) {
    for (int i = 1; i <= ii; i++)
      if (SEA_P) { /* :115-118 */
        const size_t c = IX(i, j);
        t->oneta[c + P * (size_t)(n - 1)] = MAX2(t->oneta0, 1.0 + t->pbavg[c + P * (size_t)(n - 1)] / t->pbot[c]);
        t->oneta[c + P * (size_t)(m - 1)] = MAX2(t->oneta0, 1.0 + t->pbavg[c + P * (size_t)(m - 1)] / t->pbot[c]);
      }
    for (int k = 1; k <= kk; k++) {
      const int latemp = k <= t->nhybrd && t->advflg == 0;
      const int lath3d = (k <= t->nhybrd && t->advflg == 1) || (k == 1 && t->isopyc);
      for (int i = 1; i <= ii; i++)
        if (SEA_P) {
          const size_t c = IX(i, j), ck = c + P * (size_t)(k - 1);
          const size_t cn = ck + P * K * (size_t)(n - 1), cm = ck + P * K * (size_t)(m - 1);
          const double dpold = t->dpo[cn] * t->onetao[c + P * (size_t)(n - 1)];
          const double dpmid = t->dpo[cm] * t->onetao[c + P * (size_t)(m - 1)];
          const double dpnew = t->dp[cn] * t->oneta[c + P * (size_t)(n - 1)];
          double q = 0.5 * ra2fac * (dpold + dpnew - 2.0 * dpmid);
          const double dpmidn = dpmid + q;
          t->dp[cm] = dpmidn / t->oneta[c + P * (size_t)(m - 1)];
          if (dpmidn > onezm) {
            const double qdpmidn = 1.0 / dpmidn;
            double smin, dpsold, dpsmid, dpsnew;
#define RA_FILTER(o, fm, fn)                                   \
  smin = MIN3((o), (fm), (fn));                                \
  dpsold = dpold * ((o) - smin);                               \
  dpsmid = dpmid * ((fm) - smin);                              \
  dpsnew = dpnew * ((fn) - smin);                              \
  q = 0.5 * ra2fac * (dpsold + dpsnew - 2.0 * dpsmid);         \
  (fm) = smin + (dpsmid + q) * qdpmidn
            RA_FILTER(t->osaln[ck], t->saln[cm], t->saln[cn]); /* :144-151 */
            if (latemp) { /* :169-180 */
              RA_FILTER(t->otemp[ck], t->temp[cm], t->temp[cn]);
              t->th3d[cm] = eos_sig(t->sigver, t->temp[cm], t->saln[cm]) - t->thbase;
            } else if (lath3d) { /* :181-192 */
              RA_FILTER(t->oth3d[ck], t->th3d[cm], t->th3d[cn]);
              t->temp[cm] = eos_tofsig(t->sigver, t->th3d[cm] + t->thbase, t->saln[cm]);
            } else { /* :193-198 */
              t->th3d[cm] = t->theta[ck];
              t->temp[cm] = eos_tofsig(t->sigver, t->th3d[cm] + t->thbase, t->saln[cm]);
            }
            for (int ktr = 1; ktr <= t->ntracr; ktr++) { /* :199-225 */
              const size_t o = ck + P * K * (size_t)(ktr - 1);
              const size_t tm = cm + P * K * 2 * (size_t)(ktr - 1), tn = cn + P * K * 2 * (size_t)(ktr - 1);
              RA_FILTER(t->otracer[o], t->tracer[tm], t->tracer[tn]);
            }
#undef RA_FILTER
            if (t->mxlmy) { /* :226-237 */
              const size_t qo = c + P * (size_t)k;
              const size_t qm_ = c + P * ((size_t)k + (K + 2) * (size_t)(m - 1));
              const size_t qn_ = c + P * ((size_t)k + (K + 2) * (size_t)(n - 1));
              dpsold = dpold * t->oq2[qo]; dpsmid = dpmid * t->q2[qm_]; dpsnew = dpnew * t->q2[qn_];
              q = 0.5 * ra2fac * (dpsold + dpsnew - 2.0 * dpsmid);
              t->q2[qm_] = (dpsmid + q) * qdpmidn;
              dpsold = dpold * t->oq2l[qo]; dpsmid = dpmid * t->q2l[qm_]; dpsnew = dpnew * t->q2l[qn_];
              q = 0.5 * ra2fac * (dpsold + dpsnew - 2.0 * dpsmid);
              t->q2l[qm_] = (dpsmid + q) * qdpmidn;
            }
          }
        }
    }
  }
}

/* ---- tsadvc driver: mod_tsadvc.F90:1708-2258 ------------------------------ */
int orc_tsadvc(orc_tile *t, int m, int n, int do_halo) {
  GEOM(t); MASKS(t);
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  const int kk = t->kk;
  static const int mbdy_advtyp[5] = {2, 5, 5, 0, 5}; /* :24-29 */
  const double pdzero = 0.0, pdtemp = 256.0, pdth3d = 32.0; /* :1762 */
  if (m < 1 || m > 2 || n < 1 || n > 2 || m == n) {
    snprintf(g_err, sizeof g_err, "tsadvc: bad leapfrog slots m=%d n=%d", m, n);
    return 6;
  }
  const int aadv = abs(t->advtyp);
  if (aadv > 4 || aadv == 3) {
    snprintf(g_err, sizeof g_err, "error: advem called with advtyp =%4d", t->advtyp);
    return 5;
  }
  /* :1804-1810 */
  for (size_t q = 0; q < P; q++) {
    t->onetamas[q + P * (size_t)(n - 1)] = t->oneta[q + P * (size_t)(n - 1)];
    t->onetamas[q + P * (size_t)(m - 1)] =
        t->btrmas ? t->oneta[q + P * (size_t)(n - 1)] : 1.0;
  }
  /* :1812-1813 */
  for (size_t q = 0; q < P; q++) t->uflux[q] = t->vflux[q] = 0.0;
  /* :1815-1825 */
  const int mbdy = mbdy_advtyp[aadv];
  if (nb < mbdy) {
    snprintf(g_err, sizeof g_err,
             "error: nbdy (dimensions.h) must be at least%3d for the advection "
             "scheme indicated by advtyp", mbdy);
    return 8;
  }
  /* :1827-1836 ("dp halo is up to date") */
  if (do_halo) {
    const int l = mbdy;
    orc_xctilr(t, t->saln, 1, 2 * kk, l, l);
    orc_xctilr(t, t->temp, 1, 2 * kk, l, l);
    orc_xctilr(t, t->th3d, 1, 2 * kk, l, l);
    orc_xctilr_type(t, t->uflx, 1, kk, l, l, 13); /* halo_uv */
    orc_xctilr_type(t, t->vflx, 1, kk, l, l, 14); /* halo_vv */
    for (int ktr = 1; ktr <= t->ntracr; ktr++)
      orc_xctilr(t, t->tracer + P * K * 2 * (size_t)(ktr - 1), 1, 2 * kk, l, l);
    if (t->mxlmy) { /* :1837-1840 */
      orc_xctilr(t, t->q2, 1, 2 * kk + 4, l, l);
      orc_xctilr(t, t->q2l, 1, 2 * kk + 4, l, l);
    }
  }
  const int diag = (t->nstep % 3 == 0) || t->diagno; /* :2065 */
  t->xminmax_valid = diag;
  const double *oem_m = t->onetamas + P * (size_t)(m - 1);
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);

  for (int k = 1; k <= kk; k++) { /* :1842 */
    double *temp_n = t->temp + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
    double *temp_m = t->temp + P * ((size_t)(k - 1) + K * (size_t)(m - 1));
    double *saln_n = t->saln + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
    double *saln_m = t->saln + P * ((size_t)(k - 1) + K * (size_t)(m - 1));
    double *th3d_n = t->th3d + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
    double *th3d_m = t->th3d + P * ((size_t)(k - 1) + K * (size_t)(m - 1));
    const double *dp_n = t->dp + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
    const double *uflx_k = t->uflx + P * (size_t)(k - 1);
    const double *vflx_k = t->vflx + P * (size_t)(k - 1);
    /* :1855-1857 */
    const int latemp = k <= t->nhybrd && t->advflg == 0;
    const int lath3d = (k <= t->nhybrd && t->advflg == 1) || (k == 1 && t->isopyc);
    const int smooth = t->isopyc && k == 1;
    int margin;
    if (smooth) { /* :1860-1897 */
      margin = mbdy - 1;
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++) {
          if (SEA_V) {
            const double vfa = iv[IX(i - 1, j)] != 0 ? vflx_k[IX(i - 1, j)] : vflx_k[IX(i, j)];
            const double vfb = iv[IX(i + 1, j)] != 0 ? vflx_k[IX(i + 1, j)] : vflx_k[IX(i, j)];
            t->vflux[IX(i, j)] = .5 * vflx_k[IX(i, j)] + .25 * (vfa + vfb);
          }
          if (SEA_U) {
            const double ufa = iu[IX(i, j - 1)] != 0 ? uflx_k[IX(i, j - 1)] : uflx_k[IX(i, j)];
            const double ufb = iu[IX(i, j + 1)] != 0 ? uflx_k[IX(i, j + 1)] : uflx_k[IX(i, j)];
            t->uflux[IX(i, j)] = .5 * uflx_k[IX(i, j)] + .25 * (ufa + ufb);
          }
        }
    }
    /* :1905-1942  util1=fco, util2=fcn (told/sold/trold are dead stores) */
    margin = mbdy - 1;
    {
      const double *uu = smooth ? t->uflux : uflx_k;
      const double *vv = smooth ? t->vflux : vflx_k;
      double *util1 = t->util1, *util2 = t->util2;
      const double delt1 = t->delt1;
      OMP_J
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            const double flxdiv = ((uu[IX(i + 1, j)] - uu[c]) +
                                   (vv[IX(i, j + 1)] - vv[c])) * delt1 * t->scp2i[c];
            util1[c] = MAX2(oem_m[c] * dp_n[c] + flxdiv, 0.0);
            util2[c] = MAX2(oem_m[c] * dp_n[c], 0.0);
          }
    }
    /* :1969-2015 */
    int rc = 0;
    const double *ua = uflx_k, *va = vflx_k;
    if (latemp) {
      rc |= orc_advem(t, t->advtyp, temp_n, temp_m, ua, va, t->util1, t->util2,
                      pdtemp, t->scp2, t->scp2i, t->delt1, t->btrmas);
      rc |= orc_advem(t, t->advtyp, saln_n, saln_m, ua, va, t->util1, t->util2,
                      pdzero, t->scp2, t->scp2i, t->delt1, t->btrmas);
    } else if (lath3d && t->hybrid) {
      rc |= orc_advem(t, t->advtyp, th3d_n, th3d_m, ua, va, t->util1, t->util2,
                      pdth3d, t->scp2, t->scp2i, t->delt1, t->btrmas);
      rc |= orc_advem(t, t->advtyp, saln_n, saln_m, ua, va, t->util1, t->util2,
                      pdzero, t->scp2, t->scp2i, t->delt1, t->btrmas);
    } else if (lath3d && t->isopyc) {
      rc |= orc_advem(t, t->advtyp, th3d_n, th3d_m, t->uflux, t->vflux, t->util1,
                      t->util2, pdth3d, t->scp2, t->scp2i, t->delt1, t->btrmas);
      rc |= orc_advem(t, t->advtyp, saln_n, saln_m, t->uflux, t->vflux, t->util1,
                      t->util2, pdzero, t->scp2, t->scp2i, t->delt1, t->btrmas);
    } else {
      rc |= orc_advem(t, t->advtyp, saln_n, saln_m, ua, va, t->util1, t->util2,
                      pdzero, t->scp2, t->scp2i, t->delt1, t->btrmas);
    }
    /* :2016-2034 */
    for (int ktr = 1; ktr <= t->ntracr; ktr++) {
      double *tr = t->tracer + P * K * 2 * (size_t)(ktr - 1);
      double *tr_n = tr + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
      double *tr_m = tr + P * ((size_t)(k - 1) + K * (size_t)(m - 1));
      rc |= orc_advem(t, t->advtyp, tr_n, tr_m, ua, va, t->util1, t->util2,
                      t->trcflg[ktr - 1] == 2 ? pdtemp : pdzero, t->scp2,
                      t->scp2i, t->delt1, t->btrmas);
    }
    if (t->mxlmy) { /* :2035-2048, q2(i,j,k,t): slab k + (kk+2)*(t-1), k = 0..kk+1 */
      const double pdq2 = 1.0; /* :1762 */
      double *q2_n = t->q2 + P * ((size_t)k + (K + 2) * (size_t)(n - 1));
      double *q2_m = t->q2 + P * ((size_t)k + (K + 2) * (size_t)(m - 1));
      double *q2l_n = t->q2l + P * ((size_t)k + (K + 2) * (size_t)(n - 1));
      double *q2l_m = t->q2l + P * ((size_t)k + (K + 2) * (size_t)(m - 1));
      rc |= orc_advem(t, t->advtyp, q2_n, q2_m, ua, va, t->util1, t->util2, pdq2,
                      t->scp2, t->scp2i, t->delt1, t->btrmas);
      rc |= orc_advem(t, t->advtyp, q2l_n, q2l_m, ua, va, t->util1, t->util2, pdq2,
                      t->scp2, t->scp2i, t->delt1, t->btrmas);
    }
    if (rc) return rc;
    /* :2065-2084 */
    if (diag) {
      double smin = 999., smax = -999.;
      for (int j = 1; j <= jj; j++)
        for (int i = 1; i <= ii; i++)
          if (SEA_P && dp_n[IX(i, j)] > t->onemm) {
            smin = MIN2(smin, saln_n[IX(i, j)]);
            smax = MAX2(smax, saln_n[IX(i, j)]);
          }
      t->xmin[k - 1] = smin;
      t->xmax[k - 1] = smax;
    }
  }
  /* :2138-2230 diffusion of the thermodynamic variables and the tracers */
  if (t->temdf2 > 0.0) {
    const int mdf = 2;
    /* uflux/vflux land faces are 0 (:1812-1813); uflux2/vflux2 land faces as geopar left them */
    geopar_zero_coast_faces(t, t->uflux2, t->vflux2);
    if (do_halo) {
      orc_xctilr(t, t->saln + P * K * (size_t)(n - 1), 1, kk, mdf, mdf);
      orc_xctilr(t, t->temp + P * K * (size_t)(n - 1), 1, kk, mdf, mdf);
      orc_xctilr(t, t->th3d + P * K * (size_t)(n - 1), 1, kk, mdf, mdf);
      for (int ktr = 1; ktr <= t->ntracr; ktr++)
        orc_xctilr(t, t->tracer + P * K * (2 * (size_t)(ktr - 1) + (size_t)(n - 1)),
                   1, kk, mdf, mdf);
      if (t->mxlmy) { /* :2143-2146 */
        orc_xctilr(t, t->q2 + P * (K + 2) * (size_t)(n - 1), 1, kk + 2, mdf, mdf);
        orc_xctilr(t, t->q2l + P * (K + 2) * (size_t)(n - 1), 1, kk + 2, mdf, mdf);
      }
    }
    for (int k = 1; k <= kk; k++) {
      double *temp_n = t->temp + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
      double *saln_n = t->saln + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
      double *th3d_n = t->th3d + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
      const int ldtemp = k <= t->nhybrd && t->temdfc > 0.0;
      const int ldth3d = (k <= t->nhybrd && t->temdfc < 1.0) || (k == 1 && t->isopyc);
      if (ldtemp && ldth3d) {
        tsdff(t, k, n, th3d_n, temp_n);
        tsdff(t, k, n, saln_n, NULL);
      } else if (ldtemp) {
        tsdff(t, k, n, temp_n, saln_n);
      } else if (ldth3d) {
        tsdff(t, k, n, th3d_n, saln_n);
      } else {
        tsdff(t, k, n, saln_n, NULL);
      }
      if (t->mxlmy) /* :2180-2183 */
        tsdff(t, k, n, t->q2 + P * ((size_t)k + (K + 2) * (size_t)(n - 1)),
              t->q2l + P * ((size_t)k + (K + 2) * (size_t)(n - 1)));
      for (int ktr = 1; ktr <= t->ntracr; ktr += 2) {
        double *tr1 = t->tracer + P * K * 2 * (size_t)(ktr - 1) +
                      P * ((size_t)(k - 1) + K * (size_t)(n - 1));
        if (ktr + 1 <= t->ntracr)
          tsdff(t, k, n, tr1, tr1 + P * K * 2);
        else
          tsdff(t, k, n, tr1, NULL);
      }
    }
    /* :2199-2229 non-independent thermodynamic variable (margin 0) */
    {
      OMP_J
      for (int j = 1; j <= jj; j++)
        for (int k = 1; k <= kk; k++) {
          double *temp_n = t->temp + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
          double *saln_n = t->saln + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
          double *th3d_n = t->th3d + P * ((size_t)(k - 1) + K * (size_t)(n - 1));
          const double *theta_k = t->theta + P * (size_t)(k - 1);
          const int ldtemp = k <= t->nhybrd && t->temdfc > 0.0;
          const int ldth3d = (k <= t->nhybrd && t->temdfc < 1.0) || (k == 1 && t->isopyc);
          for (int i = 1; i <= ii; i++)
            if (SEA_P) {
              const size_t c = IX(i, j);
              if (ldtemp && ldth3d) {
                const double th3d_t = eos_sig(t->sigver, temp_n[c], saln_n[c]) - t->thbase;
                th3d_n[c] = (1.0 - t->temdfc) * th3d_n[c] + t->temdfc * th3d_t;
                temp_n[c] = eos_tofsig(t->sigver, th3d_n[c] + t->thbase, saln_n[c]);
              } else if (ldtemp) {
                th3d_n[c] = eos_sig(t->sigver, temp_n[c], saln_n[c]) - t->thbase;
              } else if (ldth3d) {
                temp_n[c] = eos_tofsig(t->sigver, th3d_n[c] + t->thbase, saln_n[c]);
              } else {
                th3d_n[c] = theta_k[c];
                temp_n[c] = eos_tofsig(t->sigver, th3d_n[c] + t->thbase, saln_n[c]);
              }
            }
        }
    }
  }
  return 0;
}

static int seterr(const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return 1;
}
#include "cnuity_oracle.inc.c"
