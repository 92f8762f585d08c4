/*
 * hycom_tsadvc_b200.h -- C ABI of the B200-native tsadvc(m,n) hot path.
 *
 * The reference (HYCOM-src) has no plugin / FFI interface: the boundary of this
 * path is the Fortran module procedure `mod_tsadvc::tsadvc(m,n)`
 * (mod_tsadvc.F90:21-22,1708-1712), called once per baroclinic step from
 * HYCOM_Run (mod_hycom.F90:2535-2537), which finds all of its operands in
 * mod_cb_arrays / mod_dimensions / mod_xc.  A drop-in therefore replaces the
 * file mod_tsadvc.F90 by a thin ISO_C_BINDING shim (fortran/mod_tsadvc_b200.F90,
 * shown in INTEGRATION.md) that hands `c_loc` of those arrays to the entry
 * points below.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Array arguments are pointers to the FIRST element of the full Fortran array
 * (c_loc(temp) == &temp(1-nbdy,1-nbdy,1,1)), column-major, reals are
 * real(8) (-fdefault-real-8, config/generic-gnu-relo_one:22), integers int32:
 *   2-D  a(1-nbdy:idm+nbdy, 1-nbdy:jdm+nbdy)
 *   3-D  a(.., .., kdm)            uflx, vflx      (mod_cb_arrays.F90:147-174)
 *   4-D  a(.., .., kdm, 2)         temp, saln, th3d, dp  (mod_cb_arrays.F90:14-33)
 *   5-D  tracer(.., .., kdm, 2, ntracr)
 *   (.., .., 2)                    oneta           (mod_cb_arrays.F90:135-137)
 * m and n are the 1-based leapfrog slots of tsadvc(m,n): (:,:,:,n) holds time
 * level t-1 on entry and t+1 on exit, (:,:,:,m) holds t (mod_tsadvc.F90:1717-1728).
 *
 * Every function returns 0 on success or a HYCOM_TSADVC_E* code; the message is
 * available from hycom_tsadvc_last_error().  The reference's error behaviour
 * (print on mnproc==1, then `call xcstop('tsadvc')`, mod_tsadvc.F90:1817-1825
 * and :159-166) is reproduced by the shim from the non-zero return.
 * Like the reference routine, calls are collective over all tiles and not
 * re-entrant for one handle.
 */
#ifndef HYCOM_TSADVC_B200_H
#define HYCOM_TSADVC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYCOM_TSADVC_ABI_VERSION 3
#define HYCOM_TSADVC_MXTRCR 16

enum {
  HYCOM_TSADVC_OK = 0,
  HYCOM_TSADVC_EINVAL = 1,       /* bad argument */
  HYCOM_TSADVC_ECUDA = 2,        /* CUDA runtime error or no usable device */
  HYCOM_TSADVC_EUNSUPPORTED = 3, /* valid in the reference, not built here yet */
  HYCOM_TSADVC_ENBDY = 4,        /* nbdy < mbdy_advtyp: xcstop('tsadvc') :1817-1825 */
  HYCOM_TSADVC_EADVTYP = 5,      /* advem called with bad advtyp: xcstop('advem') :159-166 */
  HYCOM_TSADVC_ENOMEM = 6
};

/* fields of mod_cb_arrays mirrored on the device */
enum {
  HYCOM_F_TEMP = 0,
  HYCOM_F_SALN = 1,
  HYCOM_F_TH3D = 2,
  HYCOM_F_DP = 3,
  HYCOM_F_UFLX = 4,   /* 3-D: tlev ignored */
  HYCOM_F_VFLX = 5,   /* 3-D: tlev ignored */
  HYCOM_F_TRACER = 6, /* with ktr = 1..ntracr */
  HYCOM_F_ONETA = 7,  /* oneta(:,:,tlev): one slab per time slot (k0 = nk = 1) */
  HYCOM_F_THETA = 8,  /* theta(:,:,kdm), 3-D: tlev ignored (mod_cb_arrays.F90:81) */
  HYCOM_F_Q2 = 9,     /* q2(:,:,0:kdm+1,tlev): kdm+2 slabs per slot, layer k0 = 1 is k = 0 */
  HYCOM_F_Q2L = 10,   /* (mod_cb_arrays.F90:513-514; advected and diffused when mxlmy) */
  /* operands of the Robert-Asselin filter (mod_asselin.F90), see hycom_tsadvc_asselin_* below */
  HYCOM_F_DPO = 11,     /* dpo(:,:,kdm,tlev) */
  HYCOM_F_ONETAO = 12,  /* onetao(:,:,tlev): one slab per slot */
  HYCOM_F_PBAVG = 13,   /* pbavg(:,:,3), 3-D: tlev ignored, layer k0 = time slot 1..3 */
  HYCOM_F_PBOT = 14,    /* pbot(:,:), one slab */
  HYCOM_F_OTEMP = 15,   /* otemp, osaln, oth3d (:,:,kdm); otracer(:,:,kdm,ktr): time level t-1 */
  HYCOM_F_OSALN = 16,
  HYCOM_F_OTH3D = 17,
  HYCOM_F_OTRACER = 18,
  HYCOM_F_OQ2 = 19,     /* oq2, oq2l (:,:,0:kdm+1) */
  HYCOM_F_OQ2L = 20,
  /* operands of cnuity (cnuity.F90), see hycom_tsadvc_cnuity_device below */
  HYCOM_F_U = 21,       /* u, v, dpu, dpv (:,:,kdm,tlev) */
  HYCOM_F_V = 22,
  HYCOM_F_DPU = 23,
  HYCOM_F_DPV = 24,
  HYCOM_F_UBAVG = 25,   /* ubavg, vbavg (:,:,3), 3-D: tlev ignored, layer k0 = time slot 1..3 */
  HYCOM_F_VBAVG = 26,
  HYCOM_F_DEPTHU = 27,  /* depthu, depthv (:,:), one slab */
  HYCOM_F_DEPTHV = 28,
  HYCOM_F_P = 29,       /* p(:,:,kdm+1): kdm+1 slabs, 3-D */
  HYCOM_F_DPMIXL = 30,  /* dpmixl(:,:,tlev): one slab per slot */
  HYCOM_F_UFLXAV = 31,  /* uflxav, vflxav, dpav (:,:,kdm), 3-D */
  HYCOM_F_VFLXAV = 32,
  HYCOM_F_DPAV = 33,
  HYCOM_F_UTOTN = 34,   /* utotn, vtotn, dpmold (:,:), one slab */
  HYCOM_F_VTOTN = 35,
  HYCOM_F_DPMOLD = 36,
  HYCOM_F_THKDF4U = 37, /* thkdf4u, thkdf4v (:,:), one slab: the coefficients of the interface-depth diffusion at the */
  HYCOM_F_THKDF4V = 38  /* u and v points (forfun.F90:2541-2568; they hold the thkdf2 ones when thkdf4 = 0)          */
};

/* mod_dimensions.F90:33,45-49 + mod_xc tile geometry (mod_xc_mp.h:2317-3288) */
typedef struct hycom_tsadvc_dims {
  int32_t idm, jdm, kdm, nbdy;  /* array extents of this tile, halo width (6) */
  int32_t ii, jj;               /* tile extents actually used (<= idm, jdm) */
  int32_t i0, j0;               /* offset of the tile in the global grid */
  int32_t itdm, jtdm;           /* global extents */
  int32_t nreg;                 /* mod_xc.F90:25-31: 0 closed, 1 periodic in i, 2 global grid
                                   across the arctic (tripole fold of the top row),
                                   3 periodic in i and j (f-plane), 4 closed f-plane */
  int32_t ipr, jpr;             /* number of tiles in i and j */
  int32_t mproc, nproc;         /* 1-based tile coordinates */
  int32_t ntracr;               /* number of tracers (<= HYCOM_TSADVC_MXTRCR) */
  int32_t device;               /* CUDA device ordinal */
} hycom_tsadvc_dims;

/* run-time scalars tsadvc reads (blkdat.F90; mod_cb_arrays.F90:358-365,813-833) */
typedef struct hycom_tsadvc_params {
  int32_t advtyp;  /* 0 PCM, 1 MPDATA, 2 FCT2, 4 FCT4 */
  int32_t advflg;  /* 0 advect T&S, 1 advect th3d&S */
  int32_t btrmas, nhybrd, hybrid, isopyc, mxlmy;
  int32_t nstep, diagno;
  int32_t trcflg[HYCOM_TSADVC_MXTRCR];
  int32_t sigver;  /* stmt_fns.h:2-22: equation of state the host model was compiled with,
                      1/2 = 7-term sigma-0/2, 3/4 = 9-term, 5/6 = 17-term, 7/8 = 12-term;
                      read only when temdf2 > 0 (mod_tsadvc.F90:2199-2229) */
  double delt1;    /* dt2 of advem */
  double temdf2, temdfc, thbase, onemm;
} hycom_tsadvc_params;

typedef struct hycom_tsadvc_handle hycom_tsadvc_handle;

int hycom_tsadvc_abi_version(void);
const char *hycom_tsadvc_last_error(const hycom_tsadvc_handle *h);

/* device mirrors + scratch are owned by the handle (the analogue of the lazily
 * allocated module scratch, mod_tsadvc.F90:110-147); allocation is lazy */
int hycom_tsadvc_create(const hycom_tsadvc_dims *dims, hycom_tsadvc_handle **out);
int hycom_tsadvc_destroy(hycom_tsadvc_handle *h);
/* run all work of this handle on a caller-owned cudaStream_t (NULL: own stream) */
int hycom_tsadvc_set_stream(hycom_tsadvc_handle *h, void *cuda_stream);
int hycom_tsadvc_synchronize(hycom_tsadvc_handle *h);
/* bytes of device memory currently held (the mem_stat_add ledger,
 * mod_tsadvc.F90:129) */
int64_t hycom_tsadvc_device_bytes(const hycom_tsadvc_handle *h);

/* grid metrics and land/sea masks: scp2, scp2i (geopar.F90:311-340), ip, iu, iv
 * (bigrid.F90:193-297), host pointers, halos valid.  scuy..aspvy (diffusion)
 * may be NULL. */
int hycom_tsadvc_set_static(hycom_tsadvc_handle *h, const double *scp2,
                            const double *scp2i, const double *scuy,
                            const double *scvx, const double *aspux,
                            const double *aspvy, const int32_t *ip,
                            const int32_t *iu, const int32_t *iv);

/* THE drop-in entry: tsadvc(m,n) on host arrays (mod_tsadvc.F90:1708).
 * Copies temp/saln(/th3d/tracer) both time levels, dp(:,:,:,n), uflx, vflx to
 * the device, refreshes the halos, advects, and copies (:,:,:,n) of the advected
 * fields back on 1:ii,1:jj.  xmin/xmax (kdm reals, may be NULL) receive the
 * per-layer salinity range when mod(nstep,3)==0 or diagno (:2065-2094).
 * temdf2 > 0 (:2138-2230): th3d(:,:,:,n) and oneta(:,:,n) are copied in as well, the
 * fields are diffused (tsdff_1x/2x), the non-independent thermodynamic variable is rebuilt
 * with the equation of state `sigver`, and temp, saln, th3d (:,:,:,n) are all copied back.
 * theta (read in exactly-isopycnal layers, k > nhybrd) is constant in time: upload it once
 * with hycom_tsadvc_upload(h, HYCOM_F_THETA, ...).
 * mxlmy: q2 and q2l (both time slots, kdm+2 layers) are not in this argument list; the caller
 * uploads them before the call with hycom_tsadvc_upload(h, HYCOM_F_Q2 | HYCOM_F_Q2L, ...) and
 * downloads slot n afterwards (the shim does, fortran/mod_tsadvc_b200.F90). */
int hycom_tsadvc_step(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                      const hycom_tsadvc_params *prm, double *temp,
                      double *saln, double *th3d, double *tracer,
                      const double *dp, const double *uflx, const double *vflx,
                      const double *oneta, double *xmin, double *xmax);

/* device-resident variant: operands already in the handle's device mirrors */
int hycom_tsadvc_step_device(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                             const hycom_tsadvc_params *prm, double *xmin,
                             double *xmax);

/* host <-> device mirror copies of nk layers starting at layer k0 (1-based) of
 * time slot tlev (1 or 2).  `host` points at nk consecutive Fortran slabs. */
int hycom_tsadvc_upload(hycom_tsadvc_handle *h, int32_t field, int32_t ktr,
                        int32_t tlev, int32_t k0, int32_t nk, const double *host);
int hycom_tsadvc_download(hycom_tsadvc_handle *h, int32_t field, int32_t ktr,
                          int32_t tlev, int32_t k0, int32_t nk, double *host);
/* device address of one slab of a mirror and its row pitch in doubles */
int hycom_tsadvc_device_slab(hycom_tsadvc_handle *h, int32_t field, int32_t ktr,
                             int32_t tlev, int32_t k, void **dev_ptr,
                             int64_t *pitch);

/* xctilr(a,1,ld,mh,nh,halo_ps) for one mirror on a single tile (closed: vland,
 * periodic: wrap; mod_xc_sm.h:1337-1428).  step/step_device call this
 * themselves when ipr*jpr == 1. */
int hycom_tsadvc_halo_local(hycom_tsadvc_handle *h, int32_t field, int32_t ktr,
                            int32_t tlev_or_0_for_both, int32_t mh, int32_t nh);

/* ---- multi-tile runs (ipr*jpr > 1): the communicator ------------------------------------
 * The reference's xctilr / xcminr / xcmaxr live in mod_xc, which owns MPI_COMM_HYCOM
 * (mod_xc_mp.h:2317-3288, :4664-4987, :6091-6389).  Here the HANDLE owns the transport: once a
 * communicator is attached, hycom_tsadvc_step and hycom_tsadvc_step_device are complete on
 * ipr x jpr tiles - first exchange (mod_tsadvc.F90:1829-1836) overlapped with the tile
 * interior, the width-2 diffusion exchange (:2140-2151), the five exchanges inside
 * advem_fct2c (:1186-1187), xcminr/xcmaxr of the salinity range (:2093-2094).  Calls are
 * collective over all tiles, like the reference routine.  Without a communicator those two
 * entries return HYCOM_TSADVC_EUNSUPPORTED on a multi-tile handle (they never skip work).
 *   NCCL: one process per GPU.  Rank 0 obtains the 128-byte id, the host program broadcasts it
 *   (MPI_Bcast(id,128,MPI_BYTE,0,mpi_comm_hycom) in the Fortran shim), every tile calls
 *   hycom_tsadvc_comm_init; rank = mproc-1 + ipr*(nproc-1), ranks = ipr*jpr.  libnccl.so.2 is
 *   loaded at run time (HYCOM_TSADVC_NCCL_LIB overrides the name).
 *   In-process: several handles of ONE process, each driven by its own host thread (tests). */
#define HYCOM_TSADVC_COMM_ID_BYTES 128
typedef struct hycom_tsadvc_local_group hycom_tsadvc_local_group;
int hycom_tsadvc_comm_unique_id(char id[HYCOM_TSADVC_COMM_ID_BYTES]);
int hycom_tsadvc_comm_init(hycom_tsadvc_handle *h, const char id[HYCOM_TSADVC_COMM_ID_BYTES]);
int hycom_tsadvc_comm_version(int32_t *nccl_version_code);
int hycom_tsadvc_local_group_create(int32_t nranks, hycom_tsadvc_local_group **out);
int hycom_tsadvc_local_group_destroy(hycom_tsadvc_local_group *g);
int hycom_tsadvc_comm_attach_local(hycom_tsadvc_handle *h, hycom_tsadvc_local_group *g);
int hycom_tsadvc_comm_detach(hycom_tsadvc_handle *h);
/* 1 (default): the first exchange runs next to the march over the tile interior, the frame
 * follows the unpack; 0: exchange first, then the whole tile (comparison runs) */
int hycom_tsadvc_set_overlap(hycom_tsadvc_handle *h, int32_t enable);
/* xctilr(a(1-nbdy,1-nbdy,1[,tlev]),1,nk, mh,nh, itype) of one mirror through the attached
 * communicator (single tile: hycom_tsadvc_halo_local); tlev 0: both time slots;
 * itype 1 halo_ps, 13 halo_uv, 14 halo_vv (mod_xc.F90:41-44) */
int hycom_tsadvc_xctilr(hycom_tsadvc_handle *h, int32_t field, int32_t ktr,
                        int32_t tlev_or_0_for_both, int32_t mh, int32_t nh, int32_t itype);
/* Deferred diagnostics: with enable != 0 a step never waits for the device; the salinity range
 * of a diagnostic step (mod(nstep,3)==0 or diagno) stays in pinned memory until
 * hycom_tsadvc_saln_range fetches it (waits for that step only).  *nstep receives the step the
 * range belongs to, -1 if none is pending. */
int hycom_tsadvc_set_deferred_range(hycom_tsadvc_handle *h, int32_t enable);
int hycom_tsadvc_saln_range(hycom_tsadvc_handle *h, double *xmin, double *xmax, int32_t *nstep);
/* Tiling-invariant checksum of one mirror, the analogue of the reference's PIPE_CHECK hash
 * (mod_pipe.F90:724-757: MurmurHash3 over the gathered array on the first tile).  Here every
 * interior sea cell (1<=i<=ii, 1<=j<=jj, ip) of every layer contributes
 * mix64(bits(a(i,j,k)) ^ mix64(global cell index)) to a sum modulo 2**64 - independent of the
 * order of summation and of the tiling; global != 0: summed over all tiles (collective). */
int hycom_tsadvc_checksum(hycom_tsadvc_handle *h, int32_t field, int32_t ktr, int32_t tlev,
                          int32_t global, uint64_t *sum);

/* ---- multi-tile runs with a HOST-owned transport: the device side of xctilr --------------
 * mod_xc_mp.h:4664-4987 packs the mh x nh wide edge strips of a (..,..,ld) array, moves them
 * with MPI/SHMEM and unpacks them into the neighbour's halo.  A host program that wants to keep
 * the byte moving to itself (CUDA-aware MPI, torch.distributed: hycom-src_b200/xc.py) uses the
 * pack / unpack entries below and drives the step in parts; the library then performs no
 * exchange of its own.
 * Directions: 0 W, 1 E, 2 S, 3 N, 4 SW, 5 SE, 6 NW, 7 NE.  All eight neighbours are
 * addressed in one round; a corner message comes from the diagonal tile (the reference
 * gets the same values in two hops, N/S then E/W including the fresh N/S lines).
 * A message is [array][k][row][col] over the arrays tsadvc(m,n) exchanges at
 * mod_tsadvc.F90:1829-1836 (advected fields, both time slots; uflx; vflx), halo width
 * mbdy_advtyp (5). */
enum {
  HYCOM_TSADVC_PART_ALL = 0,      /* the whole tile */
  HYCOM_TSADVC_PART_INTERIOR = 1, /* cells that read no halo: may overlap the exchange */
  HYCOM_TSADVC_PART_FRAME = 2     /* the rest, after unpack; completes the step */
};
/* 0-based index (mproc-1 + ipr*(nproc-1)) of the neighbour tile per direction, -1 at a
 * closed edge; periodic edges wrap (possibly onto the tile itself).  nreg=2: the tiles of the
 * top row face their twins across the arctic, idproc(m,jpr+1) = idproc(ipr+1-m,jpr)
 * (mod_xc_mp.h:2830): N is the twin of the tile, NW / NE the twins of its western / eastern
 * neighbour, and a message sent in one of those directions is received from the SAME direction
 * (both twins look north).  Fold messages carry ii, mh+1 and mh columns of nh mirrored rows
 * (mod_xc_mp.h:4263-4372; the u-grid mirror is shifted by one column, :4283-4331). */
int hycom_tsadvc_halo_neighbors(const hycom_tsadvc_handle *h, int32_t nbr[8]);
/* doubles per direction of one message (send and receive sizes are equal for the uniform
 * tilings of hycom-src_b200/geometry.py: tiles of one row share jj, of one column ii) */
int hycom_tsadvc_halo_counts(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                             const hycom_tsadvc_params *prm, int64_t count[8]);
/* pack the edge strips into sendbuf[d] (device pointers; NULL: direction skipped) /
 * unpack recvbuf[d] into the halo (NULL: closed edge, halo := vland = 0.0 as
 * mod_xc_sm.h:1377-1422) on cuda_stream (NULL: the handle's stream) */
int hycom_tsadvc_halo_pack(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                           const hycom_tsadvc_params *prm, double *const sendbuf[8],
                           void *cuda_stream);
int hycom_tsadvc_halo_unpack(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                             const hycom_tsadvc_params *prm, double *const recvbuf[8],
                             void *cuda_stream);
/* tsadvc(m,n) on the device mirrors in two parts so that PART_INTERIOR runs while the halo
 * messages are in flight; PART_FRAME (or PART_ALL) finishes the call (leapfrog slot n is
 * switched to the new time level, diagnostics).  On a multi-tile handle no part touches
 * the halos: pack / transport / unpack come first. */
int hycom_tsadvc_step_device_part(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                  const hycom_tsadvc_params *prm, int32_t part,
                                  double *xmin, double *xmax);
/* PART_FRAME depends on the unpacked halos, not on PART_INTERIOR: given a second stream (the one
 * the unpack ran on; NULL: back to the handle's stream) its launch runs NEXT TO the interior
 * launch instead of behind it and fills the tail of that launch; the rest of the step (time-
 * level switch, diagnostics) waits for both. */
int hycom_tsadvc_set_frame_stream(hycom_tsadvc_handle *h, void *cuda_stream);

/* ---- advem_fct2c on several tiles (advtyp = 2 with btrmas, mod_tsadvc.F90:96-97, 999-1368) ---
 * The scheme calls xctilr(hloc) and xctilr(fldlo) after each of its five iterations
 * (:1186-1187).  A single-tile handle does that inside hycom_tsadvc_step_device; on several
 * tiles the caller drives the scheme per layer batch between the first exchange
 * (hycom_tsadvc_halo_*) and hycom_tsadvc_step_device_part(PART_ALL), which then only finishes
 * the step (time-level switch, diagnostics):
 *     for batch in 0..nbatch-1:  stage 0;  5 x { stage 1; fct2c_halo_pack -> transport ->
 *                                fct2c_halo_unpack };  stage 2
 * A message is [array][k][row][col] over hloc and fldlo of every advected field, the layers
 * of the batch, halo width 5. */
int hycom_tsadvc_fct2c_batches(hycom_tsadvc_handle *h, int32_t *nbatch,
                               int32_t *layers_per_batch);
/* stage 0: :1072-1086 (set-up); 1: one iteration :1090-1184 without its xctilr;
 * 2: :1202-1361 (antidiffusive fluxes, limiter, update) */
int hycom_tsadvc_fct2c_stage(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                             const hycom_tsadvc_params *prm, int32_t batch, int32_t stage);
int hycom_tsadvc_fct2c_halo_counts(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                   const hycom_tsadvc_params *prm, int32_t batch,
                                   int64_t count[8]);
int hycom_tsadvc_fct2c_halo_pack(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                 const hycom_tsadvc_params *prm, int32_t batch,
                                 double *const sendbuf[8], void *cuda_stream);
int hycom_tsadvc_fct2c_halo_unpack(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                   const hycom_tsadvc_params *prm, int32_t batch,
                                   double *const recvbuf[8], void *cuda_stream);

/* ---- diffusion part of tsadvc(m,n) (temdf2 > 0, mod_tsadvc.F90:2138-2230) ---------------
 * On a single tile hycom_tsadvc_step_device / _part(PART_ALL|PART_FRAME) run it themselves.
 * On a multi-tile handle the caller repeats the reference's second exchange
 * (xctilr(saln|temp|th3d|tracer(:,:,:,n), 1,kk, 2,2, halo_ps), :2140-2151) between the
 * advection and the diffusion: diff_halo_pack -> transport -> diff_halo_unpack ->
 * hycom_tsadvc_diffuse_device.  A message is [array][k][row][col] over saln, temp, th3d,
 * tracers of slot n, halo width 2; directions and NULL conventions as above. */
int hycom_tsadvc_diff_halo_counts(hycom_tsadvc_handle *h, int32_t n,
                                  const hycom_tsadvc_params *prm, int64_t count[8]);
int hycom_tsadvc_diff_halo_pack(hycom_tsadvc_handle *h, int32_t n,
                                const hycom_tsadvc_params *prm, double *const sendbuf[8],
                                void *cuda_stream);
int hycom_tsadvc_diff_halo_unpack(hycom_tsadvc_handle *h, int32_t n,
                                  const hycom_tsadvc_params *prm, double *const recvbuf[8],
                                  void *cuda_stream);
/* tsdff_1x/2x of every layer and field + the equation-of-state sweep on the device mirrors
 * (halos of slot n valid to width 1); a no-op returning 0 when temdf2 <= 0 */
int hycom_tsadvc_diffuse_device(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                const hycom_tsadvc_params *prm);

/* ---- next to tsadvc in the time step: the Robert-Asselin filter of
 * the scalar fields, the pointwise consumer of tsadvc's output.  Both calls work on the device
 * mirrors (fill them with hycom_tsadvc_upload, read them back with hycom_tsadvc_download);
 * prm supplies advflg, nhybrd, isopyc, mxlmy, sigver, thbase.
 *   asselin_save(m,n)   mod_asselin.F90:28-82 : oneta/onetao(:,:,n|m) = max(oneta0, 1+pbavg/pbot),
 *                       otemp/osaln/oth3d/otracer(/oq2/oq2l) = slot n; on a single tile also the
 *                       xctilr(oneta|onetao, 1,2, 6,6, halo_ps) of :77-78
 *   asselin_filter(m,n) mod_asselin.F90:84-286: oneta(:,:,n|m), dp(:,:,:,m) and slot m of saln,
 *                       temp, th3d, tracers (, q2, q2l) filtered in place */
int hycom_tsadvc_asselin_save_device(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                     const hycom_tsadvc_params *prm, double oneta0);
int hycom_tsadvc_asselin_filter_device(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                                       const hycom_tsadvc_params *prm, double ra2fac,
                                       double oneta0);

/* ---- upstream of tsadvc in the time step: cnuity(m,n), the continuity
 * equation (cnuity.F90), on the device mirrors.  It produces what tsadvc consumes - dp(:,:,:,n), uflx, vflx -
 * so a device-resident step keeps them out of the host<->device traffic.
 *   on entry : dp both slots, u, v, dpu, dpv (:,:,:,m), ubavg, vbavg (:,:,m), dpmixl(:,:,n), depthu, depthv,
 *              pbot in their mirrors (hycom_tsadvc_upload); uflx, vflx zero on land faces (geopar.F90:822-871)
 *   on exit  : dp(:,:,:,n) = t+1, dp(:,:,:,m) Robert-Asselin filtered, dpo both slots, uflx, vflx, p, utotn,
 *              vtotn, dpmold (, dpmixl(:,:,n) if isopyc); uflxav, vflxav, dpav accumulated when their mirrors
 *              exist (uploaded once)
 * The xctilr calls of :100-107 and :1400 (width 6) are done here (single tile: locally; several tiles: through
 * the communicator).  dpkmin (2*kdm reals, may be NULL) receives the per-layer minima of loops 14 and 15
 * (:471, :679) of this tile when mod(nstep,3) == 0, as the reference evaluates them.
 * thkdf4 != 0 (biharmonic) or thkdf2 != 0 (Laplacian) applies the interface-depth diffusion of :745-1124 with the
 * coefficients in the mirrors HYCOM_F_THKDF4U / _THKDF4V (uploaded once, valid halos) and its three xctilr calls.
 * hybrid .and. mxlkta (Kraus-Turner mixed layer) advects and diffuses dpmixl(:,:,n) as :1144-1324 do.
 * Scope: .not.btrmas, no open-boundary faces, no Stokes drift, not (synflt .and. wvelfl): anything else returns
 * EUNSUPPORTED. */
typedef struct hycom_cnuity_params {
  int32_t btrmas, isopyc, hybrid, mxlkta, nstep, pad;
  double delt1, ra2fac, thkdf2, thkdf4;
} hycom_cnuity_params;
int hycom_tsadvc_cnuity_device(hycom_tsadvc_handle *h, int32_t m, int32_t n,
                               const hycom_cnuity_params *prm, double *dpkmin);

/* number of kernels this library launched on the handle since creation */
int64_t hycom_tsadvc_launch_count(const hycom_tsadvc_handle *h);

/* the analogue of the reference's xctmr0/xctmr1 accumulators (timer 41 = tsadvc,
 * mod_hycom.F90:2535-2537; mod_xc_sm.h:1431-1538), at kernel granularity: when
 * enabled, every launch of the marching kernel is bracketed by CUDA events on the
 * handle's stream.  hycom_tsadvc_get_timing synchronises the stream and returns
 * the accumulated device time (ms) and the number of launches since the last
 * reset; reset != 0 clears the accumulators. */
int hycom_tsadvc_set_timing(hycom_tsadvc_handle *h, int32_t enable);
int hycom_tsadvc_get_timing(hycom_tsadvc_handle *h, double *march_ms,
                            int64_t *march_launches, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif
