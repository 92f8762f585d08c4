/*
 * oracle/tsadvc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99) of HYCOM's tsadvc(m,n) path:
 *   mod_tsadvc.F90 (tsadvc, advem, advem_pcm, advem_mpdata, advem_fct2,
 *   advem_fct4, tsdff_1x/2x), bigrid.F90 (masks, sea-only neighbours, segment
 *   tables), mod_xc_sm.h / mod_xc_mp.h (xctilr), geopar.F90:311-340 (metrics).
 *
 * PARITY PINNED AGAINST THE REFERENCE'S SOURCE TEXT, not against a reference
 * binary: the reference ships no golden vectors, no tests and no input decks,
 * and no Fortran compiler exists in this image or on the GPU boxes, so there is
 * no gfortran-built oracle/_ref.  Instead oracle/fortran_exec.py translates the Fortran of
 * /root/reference (bigrid.F90, xctilr of mod_xc_sm.h and - on threads - of
 * mod_xc_mp.h incl. ARCTIC, mod_tsadvc.F90
 * with stmt_fns.h, mod_asselin.F90, cnuity.F90) statement by statement into
 * Python - it knows the language, not the algorithms - and
 * tests/test_reference_text.py demands that this restatement reproduce what the
 * reference text computes BIT FOR BIT (unfused IEEE double, i.e. a
 * -ffp-contract=off build; tofsig of the 7/9-term fits to 1e-12 because of
 * libm's atan2/cos); tests/golden/from_reference_text.json keeps digests of
 * those runs for machines without /root/reference.  A second backend,
 * oracle/fortran_to_c.py, compiles the same text to C (oracle/_ref/*.so): equal to
 * the interpreter on every routine, equal to this restatement at the full
 * GLBb0.08 horizontal size, and what the GPU tests compare the device with.  On top: an independent
 * second restatement in numpy (oracle/np_restatement.py) that must agree bit
 * for bit, and the invariants the reference relies on.
 * A compiled-reference pin is one `bash fortran/build_ref.sh` away on any
 * machine with gfortran (tests/golden/from_reference.json).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef TSADVC_ORACLE_H
#define TSADVC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MXTRCR 16

/* One tile's view of mod_dimensions + mod_xc + mod_cb_arrays + the module-level
 * scratch of mod_tsadvc.  2-D slabs are (idm+2*nbdy)*(jdm+2*nbdy), Fortran
 * column-major with lower bounds 1-nbdy (mod_dimensions.F90:62-84). */
typedef struct orc_tile {
  /* mod_dimensions.F90:33,45-49 */
  int idm, jdm, kdm, nbdy, ms;
  int ii, jj, kk, i0, j0, itdm, jtdm;
  /* mod_xc.F90:25-31 */
  int nreg, mnproc;
  /* masks, sea-only neighbour indices, segment tables (bigrid.F90) */
  int *ip, *iu, *iv, *iq;
  int *ipim1, *ipip1, *ipjm1, *ipjp1;
  int *ifp, *ilp, *isp, *jfp, *jlp, *jsp;
  /* metrics (geopar.F90:311-340) */
  double *scp2, *scp2i, *scuy, *scvx, *aspux, *aspvy;
  /* prognostic state (mod_cb_arrays.F90:14-33,118,135-137,147-174) */
  double *temp, *saln, *th3d, *dp;   /* (P,kdm,2) */
  double *tracer;                    /* (P,kdm,2,ntracr) */
  double *uflx, *vflx;               /* (P,kdm) */
  double *q2, *q2l;                  /* (P,0:kdm+1,2) Mellor-Yamada tke fields (mxlmy) */
  /* mod_asselin.F90: dpo (P,kdm,2), onetao (P,2), pbavg (P,3), pbot (P), otemp/osaln/oth3d (P,kdm),
   * otracer (P,kdm,ntracr), oq2/oq2l (P,0:kdm+1) */
  double *dpo, *onetao, *pbavg, *pbot, *otemp, *osaln, *oth3d, *otracer, *oq2, *oq2l;
  double ra2fac, oneta0;
  double *theta;                     /* (P,kdm) isopycnic target densities, mod_cb_arrays.F90 */
  double *oneta, *onetamas;          /* (P,2) */
  double *uflux, *vflux, *uflux2, *vflux2, *util1, *util2; /* (P) */
  /* run-time scalars (blkdat) */
  int mxlkta;   /* Kraus-Turner mixed layer: cnuity advects and diffuses dpmixl (cnuity.F90:1148-1324) */
  int advtyp, advflg, btrmas, nhybrd, hybrid, isopyc, mxlmy, ntracr, nstep,
      diagno;
  int trcflg[ORC_MXTRCR];
  double delt1, temdf2, temdfc, thbase, onemm;
  int sigver;                        /* stmt_fns.h:2-22: 1..8 = {7,9,17,12}-term x sigma-{0,2} */
  /* mod_tsadvc.F90:38-51 scratch */
  double *fmx, *fmn, *flx, *fly, *fldlo, *fmxlo, *fmnlo, *fax, *fay, *rp, *rm,
      *flxdiv, *tx1, *ty1, *fldao, *fldan;
  /* diagnostics (mod_tsadvc.F90:2065-2084) */
  double *xmin, *xmax;               /* (kdm) */
  int xminmax_valid;
  /* OpenMP threads used by the sweeps (0 = runtime default) */
  int nthreads;
  /* cnuity.F90 operands, allocated by orc_cnuity_alloc:
   * u, v, dpu, dpv (P,kdm,2); ubavg, vbavg (P,3); depthu, depthv (P); p (P,kdm+1); utotn, vtotn, utotm, vtotm,
   * util3 (P); dpmixl (P,2); dpmold (P); uflxav, vflxav, dpav (P,kdm); dpkmin (2*kdm) */
  double *u, *v, *dpu, *dpv, *ubavg, *vbavg, *depthu, *depthv, *p, *utotn, *vtotn, *utotm, *vtotm, *util3,
      *dpmixl, *dpmold, *uflxav, *vflxav, *dpav, *dpkmin;
  /* interface-depth diffusion (cnuity.F90:745-1124): coefficients at the u and v points (forfun.F90:2541-2568), scratch */
  double *thkdf4u, *thkdf4v, *pold;
  double thkdf2, thkdf4;
} orc_tile;

orc_tile *orc_tile_create(int idm, int jdm, int kdm, int nbdy, int ii, int jj,
                          int i0, int j0, int itdm, int jtdm, int nreg,
                          int ntracr);
void orc_tile_destroy(orc_tile *t);
/* field access by mod_cb_arrays name ("temp", "ip", "fldlo", ...) */
double *orc_f64(orc_tile *t, const char *name);
int *orc_i32(orc_tile *t, const char *name);
int64_t orc_slab(const orc_tile *t); /* P */
/* run-time scalars by blkdat name ("advtyp", "delt1", ...) */
int orc_set_i(orc_tile *t, const char *name, int v);
int orc_get_i(const orc_tile *t, const char *name);
int orc_set_d(orc_tile *t, const char *name, double v);

/* mod_xc_sm.h:1337-1428  single-tile halo update (closed: vland=0, periodic) */
void orc_xctilr(const orc_tile *t, double *a, int l1, int ld, int mh, int nh);
/* the same with xctilr's itype (mod_xc.F90:41-44); nreg=2 is the single-tile arctic version,
 * mod_xc_sm.h:1172-1335 */
void orc_xctilr_type(const orc_tile *t, double *a, int l1, int ld, int mh, int nh, int itype);
/* mod_xc_mp.h:4664-4987 emulated over an ipr x jpr array of tiles living in one
 * address space: a[m + ipr*n] is the same field on tile (m,n), 0-based. */
void orc_world_xctilr(int ipr, int jpr, orc_tile *const *tiles,
                      double *const *a, int l1, int ld, int mh, int nh);
/* the same for a grid/field type (mod_xc.F90:41-44); nreg=2: the ARCTIC version of mod_xc_mp.h */
void orc_world_xctilr_type(int ipr, int jpr, orc_tile *const *tiles,
                           double *const *a, int l1, int ld, int mh, int nh, int itype);

/* bigrid.F90:116-386; depth is a P-sized slab whose interior 1..ii,1..jj is set.
 * stage1: halo of depth must be current; builds ip and interior iu/iv/iq and
 *         leaves iu/iv/iq as reals in util1/util2/uflux for the halo update.
 * stage2: after the halo update of those three; finishes masks, neighbours,
 *         segment tables.  orc_bigrid = single-tile driver doing both. */
int orc_bigrid_stage1(orc_tile *t, double *depth);
int orc_bigrid_stage2(orc_tile *t);
int orc_bigrid(orc_tile *t, double *depth);

/* geopar.F90:311-340 : scp2, scp2i, aspux, aspvy from scpx,scpy,scux,scuy,
 * scvx,scvy (all P-sized, halos valid). */
void orc_geopar_metrics(orc_tile *t, const double *scpx, const double *scpy,
                        const double *scux, const double *scuy,
                        const double *scvx, const double *scvy);

/* mod_tsadvc.F90:69-205 */
int orc_advem(orc_tile *t, int advtyp, double *fld, const double *fldc,
              const double *u, const double *v, const double *fco,
              const double *fcn, double posdef, const double *scal,
              const double *scali, double dt2, int btrmas);

/* stmt_fns.h: sig(t,s) and tofsig(r,s) of the equation of state `sigver` */
double orc_sig(int sigver, double t, double s);
double orc_tofsig(int sigver, double r, double s);

/* mod_tsadvc.F90:1708-2258.  m,n are the 1-based leapfrog slots.
 * do_halo=1: call the single-tile xctilr exactly where the reference does;
 * do_halo=0: the caller has already refreshed the halos (multi-tile emulation,
 *            orc_tsadvc_halo_list + orc_world_xctilr). Returns 0 or an error. */
int orc_tsadvc(orc_tile *t, int m, int n, int do_halo);

/* cnuity.F90:14-1422 within the scope stated in cnuity_oracle.inc.c; orc_cnuity_alloc adds its operands to
 * the tile (idempotent) */
int orc_cnuity_alloc(orc_tile *t);
int orc_cnuity(orc_tile *t, int m, int n, int do_halo);

/* mod_asselin.F90:28-82 and :84-286 */
void orc_asselin_save(orc_tile *t, int m, int n, int do_halo);
void orc_asselin_filter(orc_tile *t, int m, int n);

/* intermediate taps: the places the reference put pipe_compare_sym* hooks
 * (mod_tsadvc.F90:286-294,358-362,798-802,983-987).  When a tap buffer is set,
 * the named scratch slab is copied into it each time that hook is passed. */
void orc_set_tap(const char *tag, double *buf);
void orc_clear_taps(void);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
