/*
 * hycom_tsadvc_synth.h -- synthetic mod_cb_arrays state for the tsadvc path.
 *
 * The reference repository ships no input decks (regional.grid/depth, restart,
 * forcing live in HYCOM-examples), so benchmarks, smoke
 * test and parity tests use deterministic synthetic fields that are pure
 * functions of the GLOBAL (i,j,k): any tiling sees the same data, which is the
 * reference's own way of testing decompositions (mod_pipe.F90:26-127).
 * The host and the device generators produce identical bits.
 */
#ifndef HYCOM_TSADVC_SYNTH_H
#define HYCOM_TSADVC_SYNTH_H

#include <stdint.h>

#include "hycom_tsadvc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hycom_synth_cfg {
  int32_t itdm, jtdm, kdm, nreg;
  int32_t ntracr, pad;
  uint64_t seed;
  double dx0, dy0; /* grid spacing at the reference latitude (m) */
  double delt1;    /* leapfrog step the mass fluxes are scaled for (s) */
} hycom_synth_cfg;

typedef struct hycom_synth_tile {
  int32_t idm, jdm, nbdy, ii, jj, i0, j0, pad;
} hycom_synth_tile;

/* extra generator-only field ids (metrics); the HYCOM_F_* ids are shared */
enum {
  HYCOM_SYNTH_SCPX = 10, HYCOM_SYNTH_SCPY = 11, HYCOM_SYNTH_SCUX = 12,
  HYCOM_SYNTH_SCUY = 13, HYCOM_SYNTH_SCVX = 14, HYCOM_SYNTH_SCVY = 15,
  HYCOM_SYNTH_ONETA = 16
};

/* global land/sea map, itdm*jtdm bytes, 1 = sea: closed basins have their last
 * row/column land (bigrid.F90:25-45), islands are discs, no cell has 3 or 4 land
 * neighbours (bigrid.F90:156-191) */
int hycom_synth_sea_mask(const hycom_synth_cfg *cfg, uint8_t *sea);

/* fill nk Fortran slabs (layers k0..k0+nk-1) of one tile on the host.
 * lev: 0 = old time level (slot n), 1 = centre (slot m).
 * halo_mode 0: cells outside 1..ii,1..jj receive `fill`; 1: the global value
 * (periodic image or 0.0 beyond a closed edge). */
int hycom_synth_fill_host(const hycom_synth_cfg *cfg, const hycom_synth_tile *tile,
                          const uint8_t *sea, int32_t field, int32_t ktr,
                          int32_t lev, int32_t k0, int32_t nk, int32_t halo_mode,
                          double fill, double *dst);

/* same values written straight into a device mirror of `h` (all kdm layers of
 * time slot tlev) */
int hycom_tsadvc_synth_set_sea(hycom_tsadvc_handle *h, const hycom_synth_cfg *cfg,
                               const uint8_t *sea);
int hycom_tsadvc_synth_fill(hycom_tsadvc_handle *h, const hycom_synth_cfg *cfg,
                            int32_t field, int32_t ktr, int32_t tlev, int32_t lev,
                            int32_t halo_mode, double fill);

/* layers k0 .. k0+nk-1 of ANY mirror `dst_field` filled with generator field `gen` (its layers 1..nk) times
 * `scale`: bench.py fabricates the operands of cnuity (u, v, dpu, dpv, ...) this way for its timing */
int hycom_tsadvc_synth_fill_to(hycom_tsadvc_handle *h, const hycom_synth_cfg *cfg, int32_t gen,
                               int32_t ktr, int32_t lev, int32_t halo_mode, double scale,
                               int32_t dst_field, int32_t dst_tlev, int32_t k0, int32_t nk);

#ifdef __cplusplus
}
#endif
#endif
