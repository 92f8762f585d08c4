/*
 * oracle/cnuity_oracle.inc.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (included by tsadvc_oracle.c).
 *
 * CPU restatement of HYCOM's continuity equation cnuity(m,n) (cnuity.F90), the producer of the
 * dp(:,:,:,n), uflx, vflx that tsadvc(m,n) consumes, sweep by sweep with
 * the Fortran loop ranges and operation order.  Scope (everything else is refused with an error):
 *   .not.btrmas, no open-boundary faces (iuopn = ivopn = 0, :170-229, :327-358), no STOKES drift,
 *   not (synflt .and. wvelfl) (:1128-1142).
 * PARITY UNPINNED like the rest of the oracle (no Fortran compiler in this image); a second, independent
 * restatement in numpy (oracle/np_restatement.py: cnuity) must agree with it bit for bit.
 */

/* arrays cnuity needs on top of the tsadvc tile, allocated on first use (r_init = NaN) */
int orc_cnuity_alloc(orc_tile *t) {
  if (t->u) return 0;
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  t->u = alloc_r(P * K * 2); t->v = alloc_r(P * K * 2);
  t->dpu = alloc_r(P * K * 2); t->dpv = alloc_r(P * K * 2);
  t->ubavg = alloc_r(P * 3); t->vbavg = alloc_r(P * 3);
  t->depthu = alloc_r(P); t->depthv = alloc_r(P);
  t->p = alloc_r(P * (K + 1));
  t->utotn = alloc_r(P); t->vtotn = alloc_r(P); t->utotm = alloc_r(P); t->vtotm = alloc_r(P);
  t->util3 = alloc_r(P);
  t->dpmixl = alloc_r(P * 2); t->dpmold = alloc_r(P);
  t->uflxav = alloc_r(P * K); t->vflxav = alloc_r(P * K); t->dpav = alloc_r(P * K);
  t->dpkmin = alloc_r(2 * K);
  t->thkdf4u = alloc_r(P); t->thkdf4v = alloc_r(P); t->pold = alloc_r(P);
  if (!t->thkdf4u || !t->thkdf4v || !t->pold) return 1;
  for (size_t q = 0; q < P; q++) { t->thkdf4u[q] = 0.0; t->thkdf4v[q] = 0.0; }
  if (!t->u || !t->v || !t->dpu || !t->dpv || !t->ubavg || !t->vbavg || !t->depthu || !t->depthv || !t->p ||
      !t->utotn || !t->vtotn || !t->utotm || !t->vtotm || !t->util3 || !t->dpmixl || !t->dpmold ||
      !t->uflxav || !t->vflxav || !t->dpav || !t->dpkmin)
    return 1;
  /* geopar.F90:822-871: the flux scratch is zero on the land faces that bound sea segments and is never
   * written there again; p(:,:,1) = 0 (the surface) */
  for (size_t q = 0; q < P; q++) {
    t->uflux[q] = t->vflux[q] = t->uflux2[q] = t->vflux2[q] = 0.0;
    t->p[q] = 0.0;
  }
  return 0;
}

/* cnuity.F90:100-107: the single-tile xctilr calls, width 6, with their grid types */
static void cnuity_halo(orc_tile *t, int m, int n) {
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  const int kk = t->kk;
  orc_xctilr_type(t, t->dpmixl + P * (size_t)(n - 1), 1, 1, 6, 6, 1);
  orc_xctilr_type(t, t->dp, 1, 2 * kk, 6, 6, 1);
  orc_xctilr_type(t, t->dpu + P * K * (size_t)(m - 1), 1, kk, 6, 6, 3);   /* halo_us */
  orc_xctilr_type(t, t->dpv + P * K * (size_t)(m - 1), 1, kk, 6, 6, 4);   /* halo_vs */
  orc_xctilr_type(t, t->u + P * K * (size_t)(m - 1), 1, kk, 6, 6, 13);    /* halo_uv */
  orc_xctilr_type(t, t->v + P * K * (size_t)(m - 1), 1, kk, 6, 6, 14);    /* halo_vv */
  orc_xctilr_type(t, t->ubavg + P * (size_t)(m - 1), 1, 1, 6, 6, 13);
  orc_xctilr_type(t, t->vbavg + P * (size_t)(m - 1), 1, 1, 6, 6, 14);
}

/* cnuity.F90:14-1422.  do_halo as orc_tsadvc.  Returns 0, or an error for an unsupported option. */
int orc_cnuity(orc_tile *t, int m, int n, int do_halo) {
  GEOM(t); MASKS(t);
  const size_t P = (size_t)orc_slab(t), K = (size_t)t->kdm;
  const int kk = t->kk;
  if (orc_cnuity_alloc(t)) return seterr("cnuity: out of memory");
  if (t->btrmas) return seterr("cnuity: btrmas is not restated");
  if (t->thkdf2 != 0.0 && t->thkdf4 != 0.0) return seterr("cnuity: only one of thkdf2 and thkdf4 is non-zero (:758)");
  const double delt1 = t->delt1, epsil = 1.0e-11;   /* mod_cb_arrays.F90:853 */
  const int nthr = nthr_of(t), jblk = jblk_of(t, nthr);
  (void)jblk;
  double *dpn = t->dp + P * K * (size_t)(n - 1), *dpm = t->dp + P * K * (size_t)(m - 1);
  double *dpon = t->dpo + P * K * (size_t)(n - 1), *dpom = t->dpo + P * K * (size_t)(m - 1);
  const double *um = t->u + P * K * (size_t)(m - 1), *vm = t->v + P * K * (size_t)(m - 1);
  const double *dpum = t->dpu + P * K * (size_t)(m - 1), *dpvm = t->dpv + P * K * (size_t)(m - 1);
  const double *ubm = t->ubavg + P * (size_t)(m - 1), *vbm = t->vbavg + P * (size_t)(m - 1);
  double *onmn = t->onetamas + P * (size_t)(n - 1);
  double *uflux = t->uflux, *vflux = t->vflux, *uflux2 = t->uflux2, *vflux2 = t->vflux2;
  double *util1 = t->util1, *util2 = t->util2, *util3 = t->util3;
  double *utotn = t->utotn, *vtotn = t->vtotn, *utotm = t->utotm, *vtotm = t->vtotm;
  int mbdy = 6, margin;
  if (do_halo) cnuity_halo(t, m, n);

  /* :116-156 (use dp'): onetamas = oneta_u = oneta_v = 1.0 */
  margin = mbdy;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++) {
      const size_t c = IX(i, j);
      t->onetamas[c] = 1.0; t->onetamas[c + P] = 1.0;
      utotn[c] = 0.0; vtotn[c] = 0.0; util3[c] = 0.0;
      t->dpmold[c] = t->dpmixl[c + P * (size_t)(n - 1)];
      for (int k = 1; k <= kk; k++) dpon[c + P * (size_t)(k - 1)] = dpn[c + P * (size_t)(k - 1)];
    }

  for (int k = 1; k <= kk; k++) {   /* loop 76 */
    const size_t ko = P * (size_t)(k - 1);
    double *dpk = dpn + ko;
    /* :236-283 low-order fluxes at the old time level and antidiffusive fluxes */
    margin = mbdy - 1;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++) {
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_U) {
          const size_t c = IX(i, j);
          double q;
          utotm[c] = (um[c + ko] + ubm[c]) * t->scuy[c];
          if (utotm[c] >= 0.0)
            q = MIN2(dpk[IX(i - 1, j)], MAX2(0.0, t->depthu[c] - util3[IX(i - 1, j)])) * onmn[IX(i - 1, j)];
          else
            q = MIN2(dpk[c], MAX2(0.0, t->depthu[c] - util3[c])) * onmn[c];
          uflux[c] = utotm[c] * q;
          uflux2[c] = utotm[c] * dpum[c + ko] * 1.0 - uflux[c];   /* oneta_u = 1.0 */
          t->uflx[c + ko] = uflux[c];
        }
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_V) {
          const size_t c = IX(i, j);
          double q;
          vtotm[c] = (vm[c + ko] + vbm[c]) * t->scvx[c];
          if (vtotm[c] >= 0.0)
            q = MIN2(dpk[IX(i, j - 1)], MAX2(0.0, t->depthv[c] - util3[IX(i, j - 1)])) * onmn[IX(i, j - 1)];
          else
            q = MIN2(dpk[c], MAX2(0.0, t->depthv[c] - util3[c])) * onmn[c];
          vflux[c] = vtotm[c] * q;
          vflux2[c] = vtotm[c] * dpvm[c + ko] * 1.0 - vflux[c];
          t->vflx[c + ko] = vflux[c];
        }
    }
    /* :293-311 advance dp with the low-order fluxes (loop 19) */
    margin = mbdy - 2;
    {
      double dpmn[4096 + 64];
      double *mn = (size_t)(t->jdm + 2 * nb) <= sizeof dpmn / sizeof dpmn[0] ? dpmn
                                                                            : (double *)malloc(sizeof(double) * (size_t)(t->jdm + 2 * nb));
      OMP_J
      for (int j = 1 - margin; j <= jj + margin; j++) {
        double dpmin = 999.0;
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            util3[c] = util3[c] + dpk[c];
            dpk[c] = dpk[c] * onmn[c] -
                     ((uflux[IX(i + 1, j)] - uflux[c]) + (vflux[IX(i, j + 1)] - vflux[c])) * delt1 * t->scp2i[c];
            dpom[c + ko] = dpk[c];
            dpmin = MIN2(dpmin, dpk[c]);
          }
        mn[j + nb - 1] = dpmin;
      }
      double dpmin = 999.0;
      for (int j = 1; j <= jj; j++) dpmin = MIN2(dpmin, mn[j + nb - 1]);
      t->dpkmin[k - 1] = dpmin;
      if (mn != dpmn) free(mn);
    }
    /* :378-400 ratios of the largest permissible change to the sum of incoming / outgoing fluxes */
    margin = mbdy - 2;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_P) {
          const size_t c = IX(i, j);
          const int ia = t->ipim1[c], ib = t->ipip1[c], ja = t->ipjm1[c], jb = t->ipjp1[c];
          const double d0 = dpk[c], d1 = dpk[IX(ia, j)], d2 = dpk[IX(ib, j)], d3 = dpk[IX(i, ja)], d4 = dpk[IX(i, jb)];
          double u1 = MAX5(d0, d1, d2, d3, d4);
          double u2 = MAX2(0.0, MIN5(d0, d1, d2, d3, d4));
          u1 = (u1 - d0) /
               (((MAX2(0.0, uflux2[c]) - MIN2(0.0, uflux2[IX(i + 1, j)])) +
                 (MAX2(0.0, vflux2[c]) - MIN2(0.0, vflux2[IX(i, j + 1)])) + epsil) * delt1 * t->scp2i[c]);
          u2 = (u2 - d0) /
               (((MIN2(0.0, uflux2[c]) - MAX2(0.0, uflux2[IX(i + 1, j)])) +
                 (MIN2(0.0, vflux2[c]) - MAX2(0.0, vflux2[IX(i, j + 1)])) - epsil) * delt1 * t->scp2i[c]);
          util1[c] = u1; util2[c] = u2;
        }
    /* :414-441 limit the antidiffusive fluxes; utotn, vtotn keep what was clipped */
    margin = mbdy - 3;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++) {
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_U) {
          const size_t c = IX(i, j);
          double clip;
          if (uflux2[c] >= 0.0) clip = MIN3(1.0, util1[c], util2[IX(i - 1, j)]);
          else clip = MIN3(1.0, util2[c], util1[IX(i - 1, j)]);
          utotn[c] = utotn[c] + uflux2[c] * (1.0 - clip);
          uflux[c] = uflux2[c] * clip;
          t->uflx[c + ko] = t->uflx[c + ko] + uflux[c];
        }
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_V) {
          const size_t c = IX(i, j);
          double clip;
          if (vflux2[c] >= 0.0) clip = MIN3(1.0, util1[c], util2[IX(i, j - 1)]);
          else clip = MIN3(1.0, util2[c], util1[IX(i, j - 1)]);
          vtotn[c] = vtotn[c] + vflux2[c] * (1.0 - clip);
          vflux[c] = vflux2[c] * clip;
          t->vflx[c + ko] = t->vflx[c + ko] + vflux[c];
        }
    }
    /* :449-469 effect of the clipped antidiffusive fluxes on dp (loop 15) */
    margin = mbdy - 4;
    {
      double dpmin = 999.0;
      for (int j = 1 - margin; j <= jj + margin; j++) {
        double dmj = 999.0;
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            dpk[c] = dpk[c] - ((uflux[IX(i + 1, j)] - uflux[c]) + (vflux[IX(i, j + 1)] - vflux[c])) * delt1 * t->scp2i[c];
            t->p[c + P * (size_t)k] = t->p[c + P * (size_t)(k - 1)] + dpk[c];
            dmj = MIN2(dmj, dpk[c]);
          }
        if (j >= 1 && j <= jj) dpmin = MIN2(dpmin, dmj);
      }
      t->dpkmin[kk + k - 1] = dpmin;
    }
  }

  /* :580-683 (not btrmas) restore the nondivergence of the vertically integrated flow */
  for (int k = 1; k <= kk; k++) {   /* loop 77 */
    const size_t ko = P * (size_t)(k - 1);
    double *dpk = dpn + ko;
    const double *pb = t->p + P * (size_t)kk;   /* p(:,:,kk+1) */
    margin = mbdy - 5;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++) {
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_U) {
          const size_t c = IX(i, j);
          double q;
          if (utotn[c] >= 0.0) q = dpk[IX(i - 1, j)] / pb[IX(i - 1, j)];
          else q = dpk[c] / pb[c];
          uflux[c] = utotn[c] * q;
          t->uflx[c + ko] = t->uflx[c + ko] + uflux[c];
        }
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_V) {
          const size_t c = IX(i, j);
          double q;
          if (vtotn[c] >= 0.0) q = dpk[IX(i, j - 1)] / pb[IX(i, j - 1)];
          else q = dpk[c] / pb[c];
          vflux[c] = vtotn[c] * q;
          t->vflx[c + ko] = t->vflx[c + ko] + vflux[c];
        }
    }
    margin = mbdy - 6;
    {
      double dpmin = 999.0;
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            dpk[c] = dpk[c] - ((uflux[IX(i + 1, j)] - uflux[c]) + (vflux[IX(i, j + 1)] - vflux[c])) * delt1 * t->scp2i[c];
            t->p[c + P * (size_t)k] = t->p[c + P * (size_t)(k - 1)] + dpk[c];
            dpmin = MIN2(dpmin, dpk[c]);
          }
      t->dpkmin[k - 1] = dpmin;   /* loop 14 (the loop-19 values of :505-535 are overwritten, like the Fortran) */
    }
  }
  /* NOTE the Fortran updates p(:,:,k+1) inside loop 77 while q of the layers below still reads p(:,:,kk+1):
   * p(kk+1) changes only when k = kk, after its last use (:684-706). */

  /* :716-733 bottom-pressure restoring term */
  margin = mbdy - 6;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        const double q = t->pbot[c] / t->p[c + P * (size_t)kk];
        for (int k = 1; k <= kk; k++) {
          dpn[c + P * (size_t)(k - 1)] = dpn[c + P * (size_t)(k - 1)] * q;
          t->p[c + P * (size_t)k] = t->p[c + P * (size_t)(k - 1)] + dpn[c + P * (size_t)(k - 1)];
        }
        if (t->isopyc) t->dpmixl[c + P * (size_t)(n - 1)] = dpn[c];
      }

  /* :745-963 biharmonic and :973-1124 Laplacian thickness diffusion (literally, interface depth diffusion) */
  if (t->thkdf4 != 0.0 || t->thkdf2 != 0.0) {
    const int bih = t->thkdf4 != 0.0;
    const double onecm = 9806.0 * 0.01;   /* mod_cb_arrays.F90:850 */
    double *pold = t->pold;
    const double *thku = t->thkdf4u, *thkv = t->thkdf4v;   /* (the thkdf2 coefficients live in the same arrays) */
    mbdy = 6;
    if (do_halo) {   /* :761-763, :981-983 */
      orc_xctilr_type(t, t->dpmixl + P * (size_t)(n - 1), 1, 1, 6, 6, 1);
      orc_xctilr_type(t, dpn, 1, kk, 6, 6, 1);
      orc_xctilr_type(t, t->p + P, 1, kk, 6, 6, 1);
    }
    const double dtinv = 1. / delt1;
    const int iflip = t->nstep % 2;   /* :766 (the Laplacian block does not use it beyond pold) */
    margin = mbdy;
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int i = 1 - margin; i <= ii + margin; i++) {
        const size_t c = IX(i, j);
        if (SEA_U) uflux[c] = 0.;
        if (SEA_V) vflux[c] = 0.;
        if (SEA_P) {
          if (!bih) vflux[c] = 0.;   /* :1003 */
          pold[c] = iflip == 1 ? t->p[c + P * (size_t)kk] : 0.;
        }
      }
    /* :790 alternate between upward and downward direction in the k loop (biharmonic); :1013 k = 2,kk */
    const int k0 = bih ? 2 * (1 - iflip) + kk * iflip : 2, k1 = bih ? kk * (1 - iflip) + 2 * iflip : kk;
    const int kstep = bih ? 1 - 2 * iflip : 1;
    for (int k = k0; kstep > 0 ? k <= k1 : k >= k1; k += kstep) {
      double *pk = t->p + P * (size_t)(k - 1);          /* p(:,:,k) */
      const double *pb = t->p + P * (size_t)kk;         /* p(:,:,kk+1) */
      if (bih) {   /* :796-833 */
        const double *dk = dpn + P * (size_t)(k - 1), *dkm = dpn + P * (size_t)(k - 2);
        margin = mbdy - 1;
        OMP_J
        for (int j = 1 - margin; j <= jj + margin; j++)
          for (int i = 1 - margin; i <= ii + margin; i++)
            if (SEA_P) {
              const size_t c = IX(i, j);
              if (MIN2(dk[c], dkm[c]) < onecm) {
                util1[c] = 0.0;
                util2[c] = 0.0;
              } else {
                /* bigrid.F90:343-372: i-1 if sea; else i+1 if sea; otherwise i */
                const int ia = ip[IX(i - 1, j)] != 0 ? i - 1 : (ip[IX(i + 1, j)] != 0 ? i + 1 : i);
                const int ib = ip[IX(i + 1, j)] != 0 ? i + 1 : (ip[IX(i - 1, j)] != 0 ? i - 1 : i);
                const int ja = ip[IX(i, j - 1)] != 0 ? j - 1 : (ip[IX(i, j + 1)] != 0 ? j + 1 : j);
                const int jb = ip[IX(i, j + 1)] != 0 ? j + 1 : (ip[IX(i, j - 1)] != 0 ? j - 1 : j);
                util1[c] = pk[c] - .5 * (pk[IX(ia, j)] + pk[IX(ib, j)]);
                util2[c] = pk[c] - .5 * (pk[IX(i, ja)] + pk[IX(i, jb)]);
                if (util1[c] > 0.0) {
                  if (MIN2(dk[IX(ia, j)], dk[IX(ib, j)]) < onecm) util1[c] = 0.0;
                } else {
                  if (MIN2(dkm[IX(ia, j)], dkm[IX(ib, j)]) < onecm) util1[c] = 0.0;
                }
                if (util2[c] > 0.0) {
                  if (MIN2(dk[IX(i, ja)], dk[IX(i, jb)]) < onecm) util2[c] = 0.0;
                } else {
                  if (MIN2(dkm[IX(i, ja)], dkm[IX(i, jb)]) < onecm) util2[c] = 0.0;
                }
              }
            }
      }
      /* :835-906, :1029-1061 limit fluxes to avoid intertwining interfaces */
      margin = mbdy - 2;
      OMP_J
      for (int j = 1 - margin; j <= jj + margin; j++) {
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_U) {
            const size_t c = IX(i, j), w = IX(i - 1, j);
            double flxhi = .25 * (pb[c] - pk[c]) * t->scp2[c];
            double flxlo = -.25 * (pb[w] - pk[w]) * t->scp2[w];
            double want;
            if (bih) {
              if (iflip == 0) {   /* downward k loop */
                flxhi = MIN2(flxhi, uflux[c] + .25 * (pk[w] - pold[w]) * t->scp2[w]);
                flxlo = MAX2(flxlo, uflux[c] - .25 * (pk[c] - pold[c]) * t->scp2[c]);
              } else {            /* upward k loop */
                flxhi = MIN2(flxhi, uflux[c] + .25 * (pold[c] - pk[c]) * t->scp2[c]);
                flxlo = MAX2(flxlo, uflux[c] - .25 * (pold[w] - pk[w]) * t->scp2[w]);
              }
              want = (delt1 * thku[c]) * (util1[w] - util1[c]);
            } else {
              want = (delt1 * thku[c]) * (pk[w] - pk[c]);
            }
            uflux[c] = MIN2(flxhi, MAX2(flxlo, want));
            t->uflx[c + P * (size_t)(k - 2)] = t->uflx[c + P * (size_t)(k - 2)] + uflux[c] * dtinv;
            t->uflx[c + P * (size_t)(k - 1)] = t->uflx[c + P * (size_t)(k - 1)] - uflux[c] * dtinv;
          }
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_V) {
            const size_t c = IX(i, j), s_ = IX(i, j - 1);
            double flxhi = .25 * (pb[c] - pk[c]) * t->scp2[c];
            double flxlo = -.25 * (pb[s_] - pk[s_]) * t->scp2[s_];
            double want;
            if (bih) {
              if (iflip == 0) {
                flxhi = MIN2(flxhi, vflux[c] + .25 * (pk[s_] - pold[s_]) * t->scp2[s_]);
                flxlo = MAX2(flxlo, vflux[c] - .25 * (pk[c] - pold[c]) * t->scp2[c]);
              } else {
                flxhi = MIN2(flxhi, vflux[c] + .25 * (pold[c] - pk[c]) * t->scp2[c]);
                flxlo = MAX2(flxlo, vflux[c] - .25 * (pold[s_] - pk[s_]) * t->scp2[s_]);
              }
              want = (delt1 * thkv[c]) * (util2[s_] - util2[c]);
            } else {
              want = (delt1 * thkv[c]) * (pk[s_] - pk[c]);
            }
            vflux[c] = MIN2(flxhi, MAX2(flxlo, want));
            t->vflx[c + P * (size_t)(k - 2)] = t->vflx[c + P * (size_t)(k - 2)] + vflux[c] * dtinv;
            t->vflx[c + P * (size_t)(k - 1)] = t->vflx[c + P * (size_t)(k - 1)] - vflux[c] * dtinv;
          }
      }
      /* :908-922, :1063-1078 */
      margin = mbdy - 2;
      OMP_J
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            pold[c] = pk[c];
            pk[c] = pk[c] - ((uflux[IX(i + 1, j)] - uflux[c]) + (vflux[IX(i, j + 1)] - vflux[c])) * t->scp2i[c];
          }
    }
    /* :937-962, :1092-1114 */
    margin = mbdy - 2;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++)
      for (int k = 1; k <= kk; k++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            double *pa = t->p + c + P * (size_t)(k - 1), *pbk = t->p + c + P * (size_t)k;
            if (*pbk < *pa) *pbk = *pa;
            dpn[c + P * (size_t)(k - 1)] = *pbk - *pa;
            if (t->isopyc && k == 1) t->dpmixl[c + P * (size_t)(n - 1)] = dpn[c];
          }
  }

  /* :1144-1324 vertical advection of dpmixl: the excursions of the coordinates immediately above and below the
   * mixed-layer base, interpolated to dpmixl; then thickness diffusion of the mixed layer */
  if (t->hybrid && t->mxlkta) {
    double *dpmx = t->dpmixl + P * (size_t)(n - 1);
    const double *thku = t->thkdf4u, *thkv = t->thkdf4v;
    mbdy = 6;
    margin = mbdy - 2;
    OMP_J
    for (int j = 1 - margin; j <= jj + margin; j++) {
      for (int i = 1 - margin; i <= ii + margin; i++)
        if (SEA_P) { util1[IX(i, j)] = 0.; util2[IX(i, j)] = 0.; }
      for (int k = 1; k <= kk; k++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            const double dpok = dpon[c + P * (size_t)(k - 1)];
            util1[c] = util2[c];
            util2[c] = util2[c] + dpok;
            if (util2[c] >= dpmx[c] && util1[c] < dpmx[c]) {
              const double dpup = t->p[c + P * (size_t)(k - 1)] - util1[c];
              const double dpdn = t->p[c + P * (size_t)k] - util2[c];
              const double q = (util2[c] - dpmx[c]) / MAX2(t->onemm, dpok);
              dpmx[c] = dpmx[c] + (dpdn + q * (dpup - dpdn));
            }
          }
    }
    if (t->thkdf4 != 0. || t->thkdf2 != 0.) {
      const int bih = t->thkdf4 != 0.;
      margin = mbdy - 3;
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++) {
          if (SEA_U) uflux[IX(i, j)] = 0.;
          if (SEA_V) vflux[IX(i, j)] = 0.;
        }
      if (bih) {   /* :1206-1243 */
        margin = mbdy - 4;
        OMP_J
        for (int j = 1 - margin; j <= jj + margin; j++)
          for (int i = 1 - margin; i <= ii + margin; i++)
            if (SEA_P) {
              const size_t c = IX(i, j);
              const int ia = ip[IX(i - 1, j)] != 0 ? i - 1 : (ip[IX(i + 1, j)] != 0 ? i + 1 : i);
              const int ib = ip[IX(i + 1, j)] != 0 ? i + 1 : (ip[IX(i - 1, j)] != 0 ? i - 1 : i);
              const int ja = ip[IX(i, j - 1)] != 0 ? j - 1 : (ip[IX(i, j + 1)] != 0 ? j + 1 : j);
              const int jb = ip[IX(i, j + 1)] != 0 ? j + 1 : (ip[IX(i, j - 1)] != 0 ? j - 1 : j);
              util1[c] = dpmx[c] - 0.5 * (dpmx[IX(ia, j)] + dpmx[IX(ib, j)]);
              util2[c] = dpmx[c] - 0.5 * (dpmx[IX(i, ja)] + dpmx[IX(i, jb)]);
            }
        margin = mbdy - 5;
        OMP_J
        for (int j = 1 - margin; j <= jj + margin; j++)
          for (int i = 1 - margin; i <= ii + margin; i++) {
            const size_t c = IX(i, j);
            if (SEA_U) uflux[c] = (delt1 * thku[c]) * (util1[IX(i - 1, j)] - util1[c]);
            if (SEA_V) vflux[c] = (delt1 * thkv[c]) * (util2[IX(i, j - 1)] - util2[c]);
          }
      } else {     /* :1285-1301 */
        margin = mbdy - 4;
        OMP_J
        for (int j = 1 - margin; j <= jj + margin; j++)
          for (int i = 1 - margin; i <= ii + margin; i++) {
            const size_t c = IX(i, j);
            if (SEA_U) uflux[c] = (delt1 * thku[c]) * (dpmx[IX(i - 1, j)] - dpmx[c]);
            if (SEA_V) vflux[c] = (delt1 * thkv[c]) * (dpmx[IX(i, j - 1)] - dpmx[c]);
          }
      }
      margin = mbdy - 6;
      OMP_J
      for (int j = 1 - margin; j <= jj + margin; j++)
        for (int i = 1 - margin; i <= ii + margin; i++)
          if (SEA_P) {
            const size_t c = IX(i, j);
            dpmx[c] = dpmx[c] - ((uflux[IX(i + 1, j)] - uflux[c]) + (vflux[IX(i, j + 1)] - vflux[c])) * t->scp2i[c];
          }
    }
  }

  /* :1326-1350 cumulative fluxes */
  margin = 0;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int k = 1; k <= kk; k++)
      for (int i = 1 - margin; i <= ii + margin; i++) {
        const size_t c = IX(i, j), ck = c + P * (size_t)(k - 1);
        if (SEA_U) t->uflxav[ck] = t->uflxav[ck] + t->uflx[ck];
        if (SEA_V) t->vflxav[ck] = t->vflxav[ck] + t->vflx[ck];
        if (SEA_P) t->dpav[ck] = t->dpav[ck] + dpn[ck];
      }

  /* :1396-1422 Robert-Asselin time filter of the thickness field */
  if (do_halo) orc_xctilr_type(t, dpn, 1, kk, 6, 6, 1);
  margin = mbdy;
  OMP_J
  for (int j = 1 - margin; j <= jj + margin; j++)
    for (int i = 1 - margin; i <= ii + margin; i++)
      if (SEA_P) {
        const size_t c = IX(i, j);
        for (int k = 1; k <= kk; k++) {
          const size_t ck = c + P * (size_t)(k - 1);
          const double dpold = dpon[ck], dpmid = dpm[ck], dpnew = dpn[ck];
          const double q = 0.5 * t->ra2fac * (dpold + dpnew - 2.0 * dpmid);
          dpom[ck] = dpm[ck];
          dpm[ck] = dpmid + q;
        }
      }
  return 0;
}
