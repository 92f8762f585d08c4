"""CPU tests of the oracle (no GPU): the C restatement against the independent numpy
restatement (bit-exact), against the product-side host geometry, and against the
invariants the reference itself relies on (constant-field
preservation, temperature-tracer == temp, bounds, land untouched, tiling invariance,
periodic-shift invariance)."""
import importlib
import sys
import os

import ctypes as C

import numpy as np
import pytest

import util
from util import pkg, syn, cabi

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import np_restatement as npr  # noqa: E402


def _sea_eq(a, b, mask):
    return np.array_equal(np.where(mask, a, 0.0), np.where(mask, b, 0.0))


@pytest.mark.parametrize("nreg", [0, 1, 3, 4])
def test_bigrid_masks_match_host_geometry(oracle, nreg):
    cfg, sea, g, cb = util.make_case(61, 47, 2, nreg=nreg, seed=3)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    for name in ("ip", "iu", "iv"):
        assert np.array_equal(ot.i32(name), getattr(cb, name)), name
    # sea-only neighbour indices == the flag form the kernels use (bigrid.F90:316-341)
    nb = g.nbdy
    ip = cb.ip
    ipim1 = ot.i32("ipim1")
    ii_idx = np.arange(1 - nb, g.idm + nb + 1)[None, :].repeat(g.nrows, 0)
    inner = (slice(1, -1), slice(1, -1))
    expect = np.where(np.roll(ip, 1, axis=1) != 0, ii_idx - 1, ii_idx)
    assert np.array_equal(ipim1[inner], expect[inner])
    ot.close()


@pytest.mark.parametrize("advtyp,nreg,ntracr", [(2, 0, 0), (2, 3, 1), (1, 0, 2), (1, 1, 0),
                                                (0, 0, 0), (4, 0, 1), (4, 3, 0)])
def test_c_oracle_equals_numpy_restatement(oracle, advtyp, nreg, ntracr):
    cfg, sea, g, cb = util.make_case(57, 44, 3, nreg=nreg, ntracr=ntracr, seed=5, advtyp=advtyp,
                                     trcflg=[0, 2][:ntracr])
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ref["temp"][n - 1, k], alt["temp"][n - 1, k], msk), ("temp", k)
        assert _sea_eq(ref["saln"][n - 1, k], alt["saln"][n - 1, k], msk), ("saln", k)
        for q in range(ntracr):
            assert _sea_eq(ref["tracer"][q, n - 1, k], alt["tracer"][q, n - 1, k], msk), ("tracer", q, k)
    # something actually happened
    assert not _sea_eq(ref["saln"][n - 1, 0], cb.saln[n - 1, 0], msk)


@pytest.mark.parametrize("advtyp", [1, 2])
def test_intermediates_match_numpy(oracle, advtyp):
    """stage-by-stage compare at the reference's pipe_compare hook points"""
    cfg, sea, g, cb = util.make_case(40, 36, 1, nreg=0, seed=9, advtyp=advtyp)
    m, n = 2, 1
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    tags = {2: {"ad:610:fldlo": "fldlo", "ad:16:rp": "rp", "ad:16:rm": "rm", "ad:18:fax": "fax",
                "ad:18:fay": "fay"},
            1: {"ad:610:fldlo": "fldlo", "ad:16:flp": "rp", "ad:16:fln": "rm", "ad:18:flx": "flx",
                "ad:18:fly": "fly"}}[advtyp]
    bufs = {t: np.full((g.nrows, g.ncols), np.nan) for t in tags}
    # only saln (the last advem call of layer 1) survives in the taps
    for t, b in bufs.items():
        oracle.lib.orc_set_tap(t.encode(), b.ctypes.data_as(util.cabi.C.c_void_p))
    ot.tsadvc(m, n, 1)
    oracle.lib.orc_clear_taps()
    mb = 5
    saln = npr.halo_single_tile(g, cb.saln, mb, mb)
    uflx = npr.halo_single_tile(g, cb.uflx, mb, mb)
    vflx = npr.halo_single_tile(g, cb.vflx, mb, mb)
    fco, fcn = npr.prolog(g, uflx[0], vflx[0], cb.dp[n - 1, 0], 1.0, cb.delt1, cb.scp2i, cb.ip, 4)
    if advtyp == 2:
        _, inter = npr.advem_fct(g, 2, saln[n - 1, 0], saln[m - 1, 0], uflx[0], vflx[0], fco, fcn,
                                 cb.scp2, cb.scp2i, cb.delt1, cb.ip, cb.iu, cb.iv)
    else:
        _, inter = npr.advem_mpdata(g, saln[n - 1, 0], uflx[0], vflx[0], fco, fcn, 0.0, cb.scp2,
                                    cb.scp2i, cb.delt1, cb.ip, cb.iu, cb.iv)
    nb = g.nbdy
    sea_p = cb.ip != 0
    for t, key in tags.items():
        face = key in ("fax", "fay", "flx", "fly")
        mask = (cb.iu != 0) if key in ("fax", "flx") else (cb.iv != 0) if face else sea_p
        reg = np.zeros_like(mask)
        mg = 1
        reg[nb - mg:nb + g.jj + mg, nb - mg:nb + g.ii + mg] = True
        assert _sea_eq(bufs[t], inter[key], mask & reg), t
    ot.close()


@pytest.mark.parametrize("advtyp", [0, 1, 2, 4])
def test_constant_field_is_preserved_exactly(oracle, advtyp):
    cfg, sea, g, cb = util.make_case(48, 40, 2, nreg=0, seed=2, advtyp=advtyp)
    sea_all = cb.ip != 0
    cb.saln[:] = np.where(sea_all, 35.25, cb.saln)
    ref = util.run_oracle(oracle, cb, sea, 1, 2)
    msk = util.interior_sea(cb)
    if advtyp == 1:
        # MPDATA is not clamped to the constant by construction of fco (flxdiv of a
        # constant equals the thickness change only to rounding): 1e-12 relative
        assert util.rel_err(ref["saln"][1, 0], cb.saln[1, 0], msk) < 1e-12
    else:
        assert _sea_eq(ref["saln"][1, 0], cb.saln[1, 0], msk)


@pytest.mark.parametrize("advtyp", [1, 2])
def test_temperature_tracer_equals_temp(oracle, advtyp):
    """PIPE_TRACER: a trcflg==2 tracer advected with pdtemp must equal temp exactly
    (mod_pipe.F90:1517-1542, mod_tsadvc.F90:2017-2023)"""
    cfg, sea, g, cb = util.make_case(50, 42, 3, nreg=0, ntracr=1, seed=4, advtyp=advtyp, trcflg=[2])
    cb.tracer[0] = cb.temp
    ref = util.run_oracle(oracle, cb, sea, 1, 2)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ref["tracer"][0, 1, k], ref["temp"][1, k], msk)


def test_bounds_and_land_untouched(oracle):
    cfg, sea, g, cb = util.make_case(64, 52, 3, nreg=0, ntracr=1, seed=6, advtyp=2)
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    nb = g.nbdy
    land = cb.ip == 0
    inner = np.zeros_like(land)
    inner[nb:nb + g.jj, nb:nb + g.ii] = True
    for k in range(g.kdm):
        # land cells keep their bits (sentinel 2**100)
        assert np.array_equal(ref["saln"][n - 1, k][land & inner], cb.saln[n - 1, k][land & inner])
        # monotone: within the range of old/centre values of the 9x9 neighbourhood (loose form of
        # mod_tsadvc.F90:876-879,975)
        tr_new = ref["tracer"][0, n - 1, k]
        msk = util.interior_sea(cb)
        assert tr_new[msk].min() >= 0.0
        assert tr_new[msk].max() <= 1.0


def test_tiling_invariance_2x2(oracle):
    """1 tile vs 2x2 tiles with orc_world_xctilr: identical bits (the reference's
    master/slave pipe test, mod_pipe.F90:26-127)"""
    itdm, jtdm, kdm = 62, 50, 2
    cfg = util.make_cfg(itdm, jtdm, kdm, nreg=0, seed=8)
    sea = syn.sea_mask(cfg)
    m, n = 1, 2
    g1 = pkg.partition(itdm, jtdm, kdm, 1, 1, 0)[0]
    cb1 = syn.build_cb_arrays(cfg, g1, sea, m, n, advtyp=2)
    ref = util.run_oracle(oracle, cb1, sea, m, n)
    tiles = pkg.partition(itdm, jtdm, kdm, 2, 2, 0)
    ots, cbs = [], []
    for g in tiles:
        cb = syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=2)
        ot = oracle.tile(g, 0)
        # masks from the global sea map (host geometry == bigrid, tested above)
        ot.i32("ip")[...] = cb.ip
        ot.i32("iu")[...] = cb.iu
        ot.i32("iv")[...] = cb.iv
        nb = g.nbdy
        depth = np.where(cb.ip != 0, 100.0, 0.0)
        # neighbour tables / segment tables need bigrid stage 2 on the final masks
        ot.f64("util1")[...] = cb.iu
        ot.f64("util2")[...] = cb.iv
        ot.f64("uflux")[...] = 0.0
        assert oracle.lib.orc_bigrid_stage2(ot.t) == 0
        ot.i32("ip")[...] = cb.ip
        ot.load_cb(cb)
        ots.append(ot)
        cbs.append(cb)
    for name, ld in (("saln", 2 * kdm), ("temp", 2 * kdm), ("uflx", kdm), ("vflx", kdm)):
        oracle.world_xctilr(2, 2, ots, [o.f64(name) for o in ots], 1, ld, 5, 5)
    for ot in ots:
        ot.tsadvc(m, n, 0)
    nb = g1.nbdy
    for ot, g, cb in zip(ots, tiles, cbs):
        for name in ("temp", "saln"):
            loc = ot.f64(name)[n - 1, :, nb:nb + g.jj, nb:nb + g.ii]
            glb = ref[name][n - 1, :, nb + g.j0:nb + g.j0 + g.jj, nb + g.i0:nb + g.i0 + g.ii]
            sea_t = cb.ip[nb:nb + g.jj, nb:nb + g.ii] != 0
            assert np.array_equal(np.where(sea_t, loc, 0), np.where(sea_t, glb, 0)), (name, g.mproc, g.nproc)
        ot.close()


@pytest.mark.parametrize("ipr,jpr", [(2, 1), (2, 2), (4, 2), (1, 2)])
@pytest.mark.parametrize("itype", [1, 13, 14, 2, 3])
def test_arctic_world_xctilr_matches_single_tile(oracle, ipr, jpr, itype):
    """the ARCTIC xctilr of mod_xc_mp.h over ipr x jpr tiles leaves the halo the single-tile version
    of mod_xc_sm.h leaves on the global array (the two differ only in the sign of a folded zero of a
    vector field: the tiled version keeps vland, :4270-4274)"""
    itdm, jtdm, kk = 48, 36, 2
    rng = np.random.default_rng(4)
    g1 = pkg.partition(itdm, jtdm, kk, 1, 1, 2)[0]
    nb = g1.nbdy
    glob = np.full((kk, g1.nrows, g1.ncols), np.nan)
    core = rng.standard_normal((kk, jtdm, itdm))
    core[rng.random(core.shape) < 0.1] = 0.0            # vland inside the field
    glob[:, nb:nb + jtdm, nb:nb + itdm] = core
    o1 = oracle.tile(g1, 0)
    oracle.lib.orc_xctilr_type(o1.t, glob.ctypes.data_as(C.c_void_p), 1, kk, 5, 5, itype)
    tiles = pkg.partition(itdm, jtdm, kk, ipr, jpr, 2)
    ots = [oracle.tile(g, 0) for g in tiles]
    arrs = []
    for g in tiles:
        a = np.full((kk, g.nrows, g.ncols), np.nan)
        a[:, nb:nb + g.jj, nb:nb + g.ii] = core[:, g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii]
        arrs.append(a)
    oracle.world_xctilr(ipr, jpr, ots, arrs, 1, kk, 5, 5, itype)
    for g, a in zip(tiles, arrs):
        win = (slice(None), slice(g.j0 + nb - 5, g.j0 + nb + g.jj + 5), slice(g.i0 + nb - 5, g.i0 + nb + g.ii + 5))
        loc = a[:, nb - 5:nb + g.jj + 5, nb - 5:nb + g.ii + 5]
        assert not np.isnan(loc).any()
        assert np.array_equal(loc, glob[win]), (g.mproc, g.nproc)     # -0.0 == 0.0
        assert np.isnan(a[:, 0]).all()                                   # sixth halo line untouched
    for o in ots + [o1]:
        o.close()


@pytest.mark.parametrize("ipr,jpr,nreg", [(4, 2, 2), (2, 2, 2), (2, 1, 2), (1, 2, 2), (2, 2, 3), (4, 1, 0)])
def test_exchange_schedule_matches_oracle_world_xctilr(oracle, ipr, jpr, nreg):
    """the product's exchange schedule (xc.neighbors / opp_dir / halo_counts: who sends what to whom,
    which message a tile reads for which direction) with the numpy pack/unpack of tests/np_halo.py,
    run between the tiles of one process, fills every halo exactly like the C oracle's xctilr over
    tiles (mod_xc_mp.h), per grid type: scalar p-grid, u-grid vector, v-grid vector"""
    import np_halo
    xc = importlib.import_module("hycom-src_b200.xc")
    itdm, jtdm, kk = 64, 46, 2
    rng = np.random.default_rng(7)
    tiles = pkg.partition(itdm, jtdm, kk, ipr, jpr, nreg)
    nb = tiles[0].nbdy
    itypes = [1, 13, 14]
    cores = [rng.standard_normal((kk, jtdm, itdm)) for _ in itypes]
    for c in cores:
        c[rng.random(c.shape) < 0.1] = 0.0

    def tile_arrays():
        out = []
        for g in tiles:
            arrs = []
            for c in cores:
                a = np.full((kk, g.nrows, g.ncols), np.nan)
                a[:, nb:nb + g.jj, nb:nb + g.ii] = c[:, g.j0:g.j0 + g.jj, g.i0:g.i0 + g.ii]
                arrs.append(a)
            out.append(arrs)
        return out
    # the oracle
    ref = tile_arrays()
    ots = [oracle.tile(g, 0) for g in tiles]
    for q, it in enumerate(itypes):
        oracle.world_xctilr(ipr, jpr, ots, [ref[t][q] for t in range(len(tiles))], 1, kk, 5, 5, it)
    for o in ots:
        o.close()
    # the product's schedule
    got = tile_arrays()
    bes = [np_halo.NumpyHaloBackend(g, got[t], itypes) for t, g in enumerate(tiles)]
    sends, nbrs = [], []
    for t, g in enumerate(tiles):
        nbr = xc.neighbors(g)
        cnt = bes[t].counts(1, 2)
        send = [bes[t].alloc(c) if nbr[d] >= 0 else None for d, c in enumerate(cnt)]
        bes[t].pack(1, 2, send)
        sends.append(send); nbrs.append(nbr)
    for t, g in enumerate(tiles):
        recv = [sends[nbrs[t][d]][xc.opp_dir(g, d)] if nbrs[t][d] >= 0 else None for d in range(8)]
        bes[t].unpack(1, 2, recv)
    for t, g in enumerate(tiles):
        for q in range(len(itypes)):
            assert np.array_equal(got[t][q], ref[t][q], equal_nan=True), (g.mproc, g.nproc, itypes[q])
    # ... and exactly like the reference's own multi-tile xctilr, executed as written (mod_xc_mp.h in its MPI flavour,
    # one thread per tile: oracle/reference_text_mp.py), where the reference tree is on this machine and the tiling is
    # one the reference admits (a domain periodic in latitude wants jpr = 1)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import reference_text as rt
    if rt.available() and not (nreg == 3 and jpr > 1):
        import reference_text_mp as rmp
        world = rmp.World(tiles, ipr, jpr, nreg, itdm, jtdm, kk)
        txt = tile_arrays()
        for q, it in enumerate(itypes):
            world.xctilr([txt[t][q] for t in range(len(tiles))], 1, kk, 5, 5, it)
        for t, g in enumerate(tiles):
            for q in range(len(itypes)):
                assert np.array_equal(got[t][q], txt[t][q], equal_nan=True), (g.mproc, g.nproc, itypes[q], "reference text")


def test_periodic_shift_invariance(oracle):
    """PIPE_SHIFT (mod_pipe.F90:56-61): on a doubly periodic domain shifting every
    input by (si,sj) shifts the output by the same amount, bit for bit"""
    itdm, jtdm, kdm = 48, 40, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=3, seed=11, advtyp=2)
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    si, sj = 7, 5
    nb = g.nbdy
    cb2 = syn.build_cb_arrays(cfg, g, sea, m, n, advtyp=2)

    def shift_interior(a):
        core = a[..., nb:nb + jtdm, nb:nb + itdm]
        out = np.full_like(a, np.nan)
        out[..., nb:nb + jtdm, nb:nb + itdm] = np.roll(core, (sj, si), axis=(-2, -1))
        return out

    def shift_full(a):  # arrays whose halo must be valid: rebuild the halo by wrapping
        core = np.roll(a[..., nb:nb + jtdm, nb:nb + itdm], (sj, si), axis=(-2, -1))
        return np.pad(core, [(0, 0)] * (a.ndim - 2) + [(nb, nb), (nb, nb)], mode="wrap")

    for name in ("temp", "saln", "th3d", "uflx", "vflx"):
        setattr(cb2, name, shift_interior(getattr(cb, name)))
    for name in ("dp", "scp2", "scp2i"):
        setattr(cb2, name, np.ascontiguousarray(shift_full(getattr(cb, name))))
    sea2 = np.roll(sea, (sj, si), axis=(0, 1))
    cb2.ip, cb2.iu, cb2.iv = pkg.bigrid_masks(sea2, g)
    out2 = util.run_oracle(oracle, cb2, sea2, m, n)
    msk = util.interior_sea(cb2)
    for name in ("temp", "saln"):
        exp = shift_interior(ref[name][n - 1])
        for k in range(kdm):
            assert _sea_eq(out2[name][n - 1, k], exp[k], msk), (name, k)


def test_error_behaviour(oracle):
    cfg, sea, g, cb = util.make_case(30, 30, 1, seed=1, advtyp=3)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    with pytest.raises(RuntimeError, match="advtyp"):
        ot.tsadvc(1, 2, 1)
    ot.close()


def test_xminmax_diagnostic(oracle):
    cfg, sea, g, cb = util.make_case(40, 30, 2, seed=1, advtyp=2, nstep=3)
    ref = util.run_oracle(oracle, cb, sea, 1, 2)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        sel = msk & (cb.dp[1, k] > cb.onemm)
        assert ref["xmin"][k] == ref["saln"][1, k][sel].min()
        assert ref["xmax"][k] == ref["saln"][1, k][sel].max()


# ---- diffusion + equation of state (mod_tsadvc.F90:2138-2230, stmt_fns.h) -------------------

# sigma-0 / sigma-2 of standard seawater: UNESCO-83 gives sigma0(S=35,T=10) = 26.952 and
# sigma2(35,10) = 35.80 (potential density referenced to 2000 dbar, HYCOM's sigma-2 fits agree to
# a few 0.01); every family must reproduce them to the accuracy of its fit.  This pins the
# transcription of the 100-odd coefficients of stmt_fns.h.
@pytest.mark.parametrize("sigver", [1, 2, 3, 4, 5, 6, 7, 8])
def test_eos_known_seawater_density(oracle, sigver):
    want = 26.952 if sigver % 2 == 1 else 35.80
    got = oracle.sig(sigver, 10.0, 35.0)
    assert abs(got - want) < 0.06, (sigver, got)
    # d(sig)/dS ~ 0.78, d(sig)/dT ~ -0.17 at this point for every fit
    ds = oracle.sig(sigver, 10.0, 36.0) - got
    dt = oracle.sig(sigver, 11.0, 35.0) - got
    assert 0.70 < ds < 0.85 and -0.22 < dt < -0.12, (sigver, ds, dt)
    assert oracle.sig(sigver, 10.0, 35.0) == float(npr.sig(sigver, 10.0, 35.0))


@pytest.mark.parametrize("sigver", [1, 2, 3, 4, 7, 8])
def test_eos_tofsig_inverts_sig(oracle, sigver):
    for t, s in [(-1.5, 34.5), (4.0, 34.9), (12.0, 35.5), (28.0, 36.5)]:
        r = oracle.sig(sigver, t, s)
        assert abs(oracle.tofsig(sigver, r, s) - t) < 1e-8, (sigver, t, s)
        assert abs(float(npr.tofsig(sigver, r, s)) - t) < 1e-8
    assert oracle.tofsig(5, 27.0, 35.0) == 99.0 and oracle.tofsig(6, 36.0, 35.0) == 99.0   # stmt_fns.h:533


@pytest.mark.parametrize("sigver,temdfc,nhybrd,ntracr", [
    (6, 1.0, -1, 0), (5, 1.0, -1, 3), (2, 1.0, -1, 0), (4, 1.0, -1, 2), (8, 1.0, -1, 0),
    (8, 0.0, -1, 1), (7, 0.5, -1, 0), (8, 0.5, 2, 2), (2, 0.5, -1, 0), (3, 0.0, 2, 1), (1, 0.0, -1, 0)])
def test_diffusion_c_oracle_equals_numpy_restatement(oracle, sigver, temdfc, nhybrd, ntracr):
    cfg, sea, g, cb = util.make_diffusion_case(53, 41, 3, sigver, temdfc, ntracr=ntracr, nhybrd=nhybrd)
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    # tofsig of the 7/9-term fits goes through atan2/cos/sin: glibc and numpy may differ in the last bit
    libm = sigver <= 4 and (temdfc < 1.0 or (0 <= nhybrd < g.kdm))
    for name in ("temp", "saln", "th3d"):
        for k in range(g.kdm):
            a, b = ref[name][n - 1, k], alt[name][n - 1, k]
            if libm and name != "saln":
                assert util.rel_err(a, b, msk) < 1e-13, (name, k)
            else:
                assert _sea_eq(a, b, msk), (name, k)
    for q in range(ntracr):
        for k in range(g.kdm):
            assert _sea_eq(ref["tracer"][q, n - 1, k], alt["tracer"][q, n - 1, k], msk), ("tracer", q, k)
    # diffusion did something, and only smoothed: the advected-only result differs
    cb0 = util.make_diffusion_case(53, 41, 3, sigver, temdfc, ntracr=ntracr, nhybrd=nhybrd, temdf2=0.0)[3]
    adv = util.run_oracle(oracle, cb0, sea, m, n)
    assert not _sea_eq(ref["saln"][n - 1, 0], adv["saln"][n - 1, 0], msk)


def test_diffusion_preserves_constant_and_conserves(oracle):
    """a constant field stays constant; the dp-weighted integral of a diffused field is conserved
    to rounding (fluxes are antisymmetric: what leaves a cell enters its neighbour)"""
    cfg, sea, g, cb = util.make_diffusion_case(61, 47, 2, 6, 1.0, ntracr=1, temdf2=0.05)
    m, n = 1, 2
    cb.tracer[...] = np.where(np.isfinite(cb.tracer), 0.75, cb.tracer)
    cb0 = util.make_diffusion_case(61, 47, 2, 6, 1.0, ntracr=1, temdf2=0.0)[3]
    cb0.tracer[...] = cb.tracer
    for c in (cb, cb0):   # no massless layers in this test (see the note below)
        c.dp[...] = np.where(c.dp < 9806.0, 9806.0, c.dp)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    adv = util.run_oracle(oracle, cb0, sea, m, n)
    msk = util.interior_sea(cb)
    assert np.all(ref["tracer"][0, n - 1][:, msk] == 0.75)
    for k in range(g.kdm):
        w = np.where(msk, cb.dp[n - 1, k] * cb.oneta[n - 1] * cb.scp2, 0.0)
        # (max(dp,eps_har) at :2314 breaks conservation in massless cells, by design)
        a = float((w * np.where(msk, ref["saln"][n - 1, k], 0.0)).sum())
        b = float((w * np.where(msk, adv["saln"][n - 1, k], 0.0)).sum())
        assert abs(a - b) <= 1e-10 * abs(b), (k, a, b)


# ---- advem_fct2c (btrmas): mod_tsadvc.F90:999-1368 -------------------------------------------
@pytest.mark.parametrize("nreg,ntracr", [(0, 0), (3, 1), (1, 0), (4, 1)])
def test_fct2c_c_oracle_equals_numpy_restatement(oracle, nreg, ntracr):
    cfg, sea, g, cb = util.make_case(57, 44, 3, nreg=nreg, ntracr=ntracr, seed=5, advtyp=2, btrmas=True)
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ref["temp"][n - 1, k], alt["temp"][n - 1, k], msk), ("temp", k)
        assert _sea_eq(ref["saln"][n - 1, k], alt["saln"][n - 1, k], msk), ("saln", k)
        for q in range(ntracr):
            assert _sea_eq(ref["tracer"][q, n - 1, k], alt["tracer"][q, n - 1, k], msk), ("tracer", q, k)
    assert not _sea_eq(ref["saln"][n - 1, 0], cb.saln[n - 1, 0], msk)
    # a different scheme than advem_fct2: results differ from the btrmas=.false. path
    cb2 = util.make_case(57, 44, 3, nreg=nreg, ntracr=ntracr, seed=5, advtyp=2, btrmas=False)[3]
    ref2 = util.run_oracle(oracle, cb2, sea, m, n)
    assert not _sea_eq(ref["saln"][n - 1, 0], ref2["saln"][n - 1, 0], msk)


def test_fct2c_constant_field_and_bounds(oracle):
    cfg, sea, g, cb = util.make_case(70, 50, 2, nreg=0, ntracr=1, seed=9, advtyp=2, btrmas=True)
    m, n = 1, 2
    cb.tracer[...] = np.where(np.isfinite(cb.tracer), 0.5, cb.tracer)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    msk = util.interior_sea(cb)
    err = np.abs(ref["tracer"][0, n - 1][:, msk] - 0.5)
    # fct2c has no [fmn,fmx] clamp: a constant is kept to rounding except next to the ~1 % of
    # cells the generator drives to dp+flxdiv < 0 (":1169 it may happen that the cfl is violated")
    assert np.median(err) < 1e-14 and (err > 1e-12).mean() < 0.08 and err.max() < 1e-2, (err.max(), (err > 1e-12).mean())


# ---- nreg=2: global grid across the arctic, single tile (mod_xc_sm.h:1172-1335) ---------------
@pytest.mark.parametrize("advtyp,ntracr,extra", [(2, 1, {}), (1, 0, {}), (4, 0, {}), (2, 0, {"btrmas": True})])
def test_arctic_c_oracle_equals_numpy_restatement(oracle, advtyp, ntracr, extra):
    cfg, sea, g, cb = util.make_arctic_case(60, 47, 2, ntracr=ntracr, seed=7, advtyp=advtyp, **extra)
    m, n = 1, 2
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    for name in ("ip", "iu", "iv"):          # bigrid with halo_us/halo_vs across the fold
        assert np.array_equal(ot.i32(name), getattr(cb, name)), name
    assert ot.i32("ip")[g.nbdy + g.jj:, g.nbdy:g.nbdy + g.ii].any()   # sea cells in the north halo
    ot.tsadvc(m, n, 1)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ot.f64("temp")[n - 1, k], alt["temp"][n - 1, k], msk), ("temp", k)
        assert _sea_eq(ot.f64("saln")[n - 1, k], alt["saln"][n - 1, k], msk), ("saln", k)
    # the fold is felt: the top rows differ from a run that closes the northern boundary
    cb1 = util.make_arctic_case(60, 47, 2, ntracr=ntracr, seed=7, advtyp=advtyp, **extra)[3]
    g1 = pkg_partition_closed(g)
    cb1.geom = g1
    alt1 = npr.tsadvc(cb1, m, n)
    top = np.zeros_like(msk); top[g.nbdy + g.jj - 3:g.nbdy + g.jj] = True
    assert not _sea_eq(alt["saln"][n - 1, 0], alt1["saln"][n - 1, 0], msk & top)
    ot.close()


def pkg_partition_closed(g):
    import dataclasses
    return dataclasses.replace(g, nreg=1)


# ---- mxlmy: q2, q2l advected and diffused with the thermodynamic fields (:2035-2048, :2180-2183) --
@pytest.mark.parametrize("advtyp,temdf2,nreg", [(2, 0.0, 0), (1, 0.0, 3), (2, 0.02, 0)])
def test_mxlmy_c_oracle_equals_numpy_restatement(oracle, advtyp, temdf2, nreg):
    if temdf2 > 0:
        cfg, sea, g, cb = util.make_diffusion_case(57, 44, 3, 6, 1.0, nreg=nreg, seed=5, advtyp=advtyp, temdf2=temdf2)
    else:
        cfg, sea, g, cb = util.make_case(57, 44, 3, nreg=nreg, seed=5, advtyp=advtyp)
    m, n = 1, 2
    util.add_q2(cfg, sea, g, cb, m, n)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for name in ("q2", "q2l"):
        for k in range(1, g.kdm + 1):
            assert _sea_eq(ref[name][n - 1, k], alt[name][n - 1, k], msk), (name, k)
        assert not _sea_eq(ref[name][n - 1, 1], cb.__dict__[name][n - 1, 1], msk)
        for k in (0, g.kdm + 1):      # the boundary layers are not advected
            assert np.array_equal(ref[name][n - 1, k][msk], cb.__dict__[name][n - 1, k][msk])


# ---- isopycnic coordinates: layer 1 on smoothed mass fluxes (:1859-1897), th3d & saln there, ----
# ---- saln only below (:1848-1850)
@pytest.mark.parametrize("advtyp,nreg", [(2, 0), (1, 3), (0, 0)])
def test_isopyc_c_oracle_equals_numpy_restatement(oracle, advtyp, nreg):
    cfg, sea, g, cb = util.make_case(57, 44, 3, nreg=nreg, seed=5, advtyp=advtyp, isopyc=True, hybrid=False, nhybrd=0)
    m, n = 1, 2
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ref["saln"][n - 1, k], alt["saln"][n - 1, k], msk), ("saln", k)
        assert _sea_eq(ref["th3d"][n - 1, k], alt["th3d"][n - 1, k], msk), ("th3d", k)
        assert _sea_eq(ref["temp"][n - 1, k], cb.temp[n - 1, k], msk)                       # temp is not advected
    assert not _sea_eq(ref["th3d"][n - 1, 0], cb.th3d[n - 1, 0], msk)
    assert _sea_eq(ref["th3d"][n - 1, 1], cb.th3d[n - 1, 1], msk)                          # th3d only in layer 1
    hyb = util.make_case(57, 44, 3, nreg=nreg, seed=5, advtyp=advtyp)[3]
    ref2 = util.run_oracle(oracle, hyb, sea, m, n)
    assert not _sea_eq(ref["saln"][n - 1, 0], ref2["saln"][n - 1, 0], msk)                 # smoothing changes layer 1
    assert _sea_eq(ref["saln"][n - 1, 1], ref2["saln"][n - 1, 1], msk)                     # ... and only layer 1


# isopyc with tracers and mxlmy: the tracers and q2, q2l of layer 1 keep uflx(:,:,1) while the prolog
# (fco, fcn) is built from the smoothed fluxes (:1930-1932 against :2016-2048)
@pytest.mark.parametrize("advtyp,nreg,mxlmy", [(2, 0, True), (1, 3, False), (4, 0, False), (0, 0, True)])
def test_isopyc_with_tracers_c_oracle_equals_numpy_restatement(oracle, advtyp, nreg, mxlmy):
    kw = dict(nreg=nreg, seed=9, advtyp=advtyp, isopyc=True, hybrid=False, nhybrd=0, ntracr=2, trcflg=[2, 0])
    m, n = 2, 1
    cfg, sea, g, cb = util.make_case(57, 44, 3, m=m, n=n, **kw)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb, m, n)
    ref = util.run_oracle(oracle, cb, sea, m, n)
    alt = npr.tsadvc(cb, m, n)
    msk = util.interior_sea(cb)
    for k in range(g.kdm):
        assert _sea_eq(ref["saln"][n - 1, k], alt["saln"][n - 1, k], msk), ("saln", k)
        assert _sea_eq(ref["th3d"][n - 1, k], alt["th3d"][n - 1, k], msk), ("th3d", k)
        for q in range(2):
            assert _sea_eq(ref["tracer"][q, n - 1, k], alt["tracer"][q, n - 1, k], msk), ("tracer", q, k)
        if mxlmy:
            assert _sea_eq(ref["q2"][n - 1, k + 1], alt["q2"][n - 1, k + 1], msk), ("q2", k)
            assert _sea_eq(ref["q2l"][n - 1, k + 1], alt["q2l"][n - 1, k + 1], msk), ("q2l", k)
    assert not _sea_eq(ref["tracer"][1, n - 1, 0], cb.tracer[1, n - 1, 0], msk)


# ---- frozen vectors (tests/golden): the oracle and the numpy restatement reproduce them ---------
import json as _json
import os as _os

_GOLD = _json.load(open(_os.path.join(_os.path.dirname(__file__), "golden", "tsadvc_golden.json")))


@pytest.mark.parametrize("name", sorted(_GOLD))
def test_oracle_reproduces_golden_vectors(oracle, name):
    import sys
    sys.path.insert(0, _os.path.join(_os.path.dirname(__file__), "golden"))
    import make_golden
    got, (cfg, sea, g, cb) = make_golden.run_oracle_case(oracle, name)
    assert got == _GOLD[name], name
    # and so does the independent restatement (the golden bits are not the C code's private opinion)
    alt = npr.tsadvc(cb, 1, 2)
    msk = util.interior_sea(cb)
    flds = dict(temp=alt["temp"], saln=alt["saln"], th3d=alt["th3d"], tracer=alt.get("tracer"))
    assert make_golden.digest(flds, msk, 2) == _GOLD[name], (name, "numpy restatement")


# ---- mod_asselin.F90: asselin_save + asselin_filter ----------------
@pytest.mark.parametrize("sigver,ntracr,extra", [(6, 2, {}), (8, 0, {"advflg": 1}), (2, 1, {"nhybrd": 1}),
                                                  (7, 0, {"isopyc": True, "hybrid": False, "nhybrd": 0})])
def test_asselin_c_oracle_equals_numpy_restatement(oracle, sigver, ntracr, extra):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(57, 44, 3, nreg=0, ntracr=ntracr, seed=5, **extra)
    if sigver == 6:
        util.add_q2(cfg, sea, g, cb, m, n)
    util.add_asselin(cfg, sea, g, cb, m, n, sigver=sigver)
    msk = util.interior_sea(cb)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    # asselin_save
    ot.asselin_save(m, n, 1)
    sv = npr.asselin_save(cb, m, n)
    for name in ("oneta", "onetao"):
        assert np.array_equal(ot.f64(name), sv[name], equal_nan=True), name
    for name in ("otemp", "osaln", "oth3d") + (("otracer",) if ntracr else ()) + (("oq2", "oq2l") if cb.mxlmy else ()):
        a, b = ot.f64(name), sv[name]
        assert np.array_equal(a[..., msk], b[..., msk]), name
    # asselin_filter from the same starting state
    ot2 = util.oracle_tile_from_cb(oracle, cb, sea)
    ot2.asselin_filter(m, n)
    fl = npr.asselin_filter(cb, m, n)
    libm = sigver <= 4 and ("nhybrd" in extra or extra.get("advflg") == 1 or extra.get("isopyc"))
    for name in ("oneta", "dp", "temp", "saln", "th3d") + (("tracer",) if ntracr else ()) + (("q2", "q2l") if cb.mxlmy else ()):
        a, b = ot2.f64(name), fl[name]
        if libm and name in ("temp",):
            assert util.rel_err(a[m - 1], b[m - 1], msk) < 1e-13, name
        else:
            assert np.array_equal(a[..., msk], b[..., msk], equal_nan=True), name
    assert not np.array_equal(ot2.f64("saln")[m - 1, 0][msk], cb.saln[m - 1, 0][msk])      # slot m was filtered
    assert np.array_equal(ot2.f64("saln")[n - 1][:, msk], cb.saln[n - 1][:, msk])           # slot n untouched
    ot.close(); ot2.close()


def test_asselin_conserves_constants(oracle):
    """'version that exactly conserves constant salinity' (mod_asselin.F90:143): a field that is the
    same constant at t-1, t and t+1 comes out as that constant, bit for bit"""
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(40, 30, 2, ntracr=1, seed=3)
    util.add_asselin(cfg, sea, g, cb, m, n)
    cb.tracer[...] = 0.625
    cb.otracer[...] = 0.625
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    ot.asselin_filter(m, n)
    msk = util.interior_sea(cb)
    assert np.all(ot.f64("tracer")[0, m - 1][:, msk] == 0.625)
    ot.close()


# ---------------------------------------------------------------------------------------
# the pin: digests produced by the REFERENCE itself (fortran/build_ref.sh, needs gfortran)
# ---------------------------------------------------------------------------------------
_REF_FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "from_reference.json")


@pytest.mark.skipif(not os.path.exists(_REF_FIXTURE),
                    reason="tests/golden/from_reference.json absent: no Fortran compiler has produced it yet "
                           "(fortran/build_ref.sh); the pin is the executed reference text (tests/test_reference_text.py)")
def test_reference_fixture(oracle):
    """the oracle reproduces the bits of the reference's own tsadvc on the golden cases"""
    import json
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    ref = json.load(open(_REF_FIXTURE))
    for name in sorted(ref):
        got, _ = make_golden.run_oracle_case(oracle, name)
        assert got == ref[name], name


def test_reference_case_io_roundtrip(oracle, tmp_path):
    """fortran/ref_case.py: the files it writes hold the case in the Fortran layout, and its digest of
    out_*.bin files equals the frozen golden digest when those files hold the oracle's result - so the
    only untested link of the pin is the Fortran compiler"""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    sys.path.insert(0, os.path.join(root, "fortran"))
    import make_golden
    import ref_case
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tsadvc_golden.json")))
    for name in ("box_fct2", "periodic_mpdata_tracers"):
        d = tmp_path / ("case_" + name)
        ref_case.write(name, str(d))
        kind, kw = make_golden.CASES[name]
        cfg, sea, g, cb = make_golden.build(kind, kw)
        back = np.fromfile(str(d / "saln.bin"), dtype="<f8").reshape(cb.saln.shape)
        assert np.array_equal(back, cb.saln, equal_nan=True)
        hdr = open(str(d / "case.txt")).read().split()
        assert [int(v) for v in hdr[:3]] == [g.itdm, g.jtdm, g.kdm] and float(hdr[16]) == cb.delt1
        out = util.run_oracle(oracle, cb, sea, 1, 2)
        for nm in ("temp", "saln", "th3d"):
            out[nm].astype("<f8").tofile(str(d / f"out_{nm}.bin"))
        if cb.ntracr:
            out["tracer"].astype("<f8").tofile(str(d / "out_tracer.bin"))
    here = os.getcwd()
    fixture = os.path.join(root, "tests", "golden", "from_reference.json")
    keep = open(fixture).read() if os.path.exists(fixture) else None
    try:
        ref_case.digest(str(tmp_path), ["box_fct2", "periodic_mpdata_tracers"])
        got = json.load(open(fixture))
    finally:
        if keep is None:
            os.remove(fixture)
        else:
            open(fixture, "w").write(keep)
        os.chdir(here)
    for name in got:
        assert got[name] == gold[name], name


# ---------------------------------------------------------------------------------------
# cnuity(m,n) (cnuity.F90): C restatement == numpy restatement
# ---------------------------------------------------------------------------------------
def _np_cnuity(cb, g, st, m, n, ra2fac=0.125, isopyc=False, mxlkta=False):
    """the numpy path with the single-tile xctilr calls of cnuity.F90:100-107 and :1400"""
    thkdf = st.get("_thkdf")
    st = {k: v.copy() for k, v in st.items() if not k.startswith("_")}
    nb = g.nbdy
    H = lambda a, it: npr.halo_single_tile(g, a, 6, 6, it)   # noqa: E731
    st["dpmixl"][n - 1] = H(st["dpmixl"][n - 1], 1)
    st["dp"] = H(st["dp"], 1)
    st["dpu"][m - 1] = H(st["dpu"][m - 1], 3)
    st["dpv"][m - 1] = H(st["dpv"][m - 1], 4)
    st["u"][m - 1] = H(st["u"][m - 1], 13)
    st["v"][m - 1] = H(st["v"][m - 1], 14)
    st["ubavg"][m - 1] = H(st["ubavg"][m - 1], 13)
    st["vbavg"][m - 1] = H(st["vbavg"][m - 1], 14)
    thk = None
    if "thkdf4u" in st:
        thk = dict(thkdf4u=st["thkdf4u"], thkdf4v=st["thkdf4v"], bih=thkdf[1], nstep=cb.nstep, scp2=cb.scp2, halo=H)
    p, utotn, vtotn, dpkmin, dpmold = npr.cnuity(g, st, m, n, cb.ip, cb.iu, cb.iv, cb.scuy, cb.scvx, cb.scp2i,
                                                 st["depthu"], st["depthv"], st["pbot"], cb.delt1, ra2fac, isopyc, thk,
                                                 dict(onemm=cb.onemm) if mxlkta else None)
    st["dp"][n - 1] = H(st["dp"][n - 1], 1)
    npr.cnuity_asselin(g, st, m, n, cb.ip, ra2fac)
    st.update(p=p, utotn=utotn, vtotn=vtotn, dpkmin=dpkmin)
    return st


@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,m,n,isopyc", [
    (90, 70, 4, 0, 1, 2, False),      # closed basin with islands
    (131, 77, 3, 3, 2, 1, False),     # doubly periodic, slots swapped
    (64, 90, 3, 1, 1, 2, True),       # periodic in i, isopyc: dpmixl(n) = dp(1,n)
    (70, 45, 2, 4, 1, 2, False),      # closed f-plane (periodic in j)
])
def test_cnuity_c_oracle_equals_numpy(oracle, itdm, jtdm, kdm, nreg, m, n, isopyc):
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=23, m=m, n=n)
    st = util.add_cnuity(cfg, sea, g, cb, m, n)
    want = _np_cnuity(cb, g, st, m, n, isopyc=isopyc)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    util.oracle_load_cnuity(ot, st)
    ot.set_i("isopyc", int(isopyc))
    ot.cnuity(m, n, 1)
    nb = g.nbdy
    inner = util.interior_sea(cb)
    sea6 = cb.ip != 0
    iu_in = np.zeros_like(inner); iv_in = np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    for k in range(kdm):
        assert np.array_equal(ot.f64("dp")[n - 1, k][inner], want["dp"][n - 1, k][inner]), ("dp.n", k)
        assert np.array_equal(ot.f64("dp")[m - 1, k][sea6], want["dp"][m - 1, k][sea6]), ("dp.m", k)
        assert np.array_equal(ot.f64("dpo")[m - 1, k][sea6], want["dpo"][m - 1, k][sea6]), ("dpo.m", k)
        assert np.array_equal(ot.f64("dpo")[n - 1, k], want["dpo"][n - 1, k], equal_nan=True), ("dpo.n", k)
        assert np.array_equal(ot.f64("uflx")[k][iu_in], want["uflx"][k][iu_in]), ("uflx", k)
        assert np.array_equal(ot.f64("vflx")[k][iv_in], want["vflx"][k][iv_in]), ("vflx", k)
        assert np.array_equal(ot.f64("p")[k + 1][inner], want["p"][k + 1][inner]), ("p", k)
        assert np.array_equal(ot.f64("dpav")[k][inner], want["dpav"][k][inner]), ("dpav", k)
        assert np.array_equal(ot.f64("uflxav")[k][iu_in], want["uflxav"][k][iu_in]), ("uflxav", k)
    assert np.array_equal(ot.f64("utotn")[iu_in], want["utotn"][iu_in])
    assert np.array_equal(ot.f64("vtotn")[iv_in], want["vtotn"][iv_in])
    assert np.array_equal(ot.f64("dpkmin"), want["dpkmin"])
    if isopyc:
        assert np.array_equal(ot.f64("dpmixl")[n - 1][inner], want["dpmixl"][n - 1][inner])
    # the column still sums to pbot after the restoring term, and the thicknesses moved
    colsum = ot.f64("dp")[n - 1].sum(axis=0)
    assert np.allclose(colsum[inner], st["pbot"][inner], rtol=1e-12)
    assert not np.array_equal(ot.f64("dp")[n - 1, 0][inner], st["dp"][n - 1, 0][inner])
    ot.close()


# interface-depth diffusion inside cnuity (cnuity.F90:745-1124): biharmonic in both sweep directions (nstep even:
# downward, odd: upward) and Laplacian
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,bih,nstep,isopyc", [
    (90, 70, 5, 0, True, 4, False),
    (90, 70, 5, 0, True, 7, False),
    (64, 90, 4, 1, True, 3, True),
    (131, 77, 3, 3, False, 2, False),
    (70, 45, 4, 4, False, 5, False),
])
def test_cnuity_thickness_diffusion_c_oracle_equals_numpy(oracle, itdm, jtdm, kdm, nreg, bih, nstep, isopyc):
    m, n = 1, 2
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=29, m=m, n=n, nstep=nstep, **extra)
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=0.01 if bih else 0.02, bih=bih)
    want = _np_cnuity(cb, g, st, m, n, isopyc=isopyc)
    plain = _np_cnuity(cb, g, {k: v for k, v in st.items() if k not in ("thkdf4u", "thkdf4v", "_thkdf")}, m, n,
                       isopyc=isopyc)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    util.oracle_load_cnuity(ot, st)
    ot.set_i("isopyc", int(isopyc))
    ot.cnuity(m, n, 1)
    nb = g.nbdy
    inner = util.interior_sea(cb)
    iu_in = np.zeros_like(inner); iv_in = np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    for k in range(kdm):
        assert np.array_equal(ot.f64("dp")[n - 1, k][inner], want["dp"][n - 1, k][inner]), ("dp.n", k)
        assert np.array_equal(ot.f64("dp")[m - 1, k][inner], want["dp"][m - 1, k][inner]), ("dp.m", k)
        assert np.array_equal(ot.f64("uflx")[k][iu_in], want["uflx"][k][iu_in]), ("uflx", k)
        assert np.array_equal(ot.f64("vflx")[k][iv_in], want["vflx"][k][iv_in]), ("vflx", k)
        assert np.array_equal(ot.f64("p")[k + 1][inner], want["p"][k + 1][inner]), ("p", k)
        assert np.array_equal(ot.f64("dpav")[k][inner], want["dpav"][k][inner]), ("dpav", k)
        assert np.array_equal(ot.f64("uflxav")[k][iu_in], want["uflxav"][k][iu_in]), ("uflxav", k)
    if isopyc:
        assert np.array_equal(ot.f64("dpmixl")[n - 1][inner], want["dpmixl"][n - 1][inner])
    # the diffusion moved interfaces (and only interfaces: the column still sums to pbot), layers stay non-negative
    got = ot.f64("dp")[n - 1]
    assert not np.array_equal(got[1][inner], plain["dp"][n - 1, 1][inner])
    assert np.allclose(got.sum(axis=0)[inner], st["pbot"][inner], rtol=1e-12)
    assert (got[:, inner] >= 0.0).all()
    ot.close()


# hybrid .and. mxlkta (cnuity.F90:1144-1324): dpmixl follows the coordinates around the mixed-layer base and is
# diffused like an interface
@pytest.mark.parametrize("itdm,jtdm,kdm,nreg,mode,nstep", [
    (90, 70, 5, 0, "bih", 4),
    (64, 90, 4, 1, "lap", 3),
    (131, 77, 4, 3, None, 2),
])
def test_cnuity_mxlkta_c_oracle_equals_numpy(oracle, itdm, jtdm, kdm, nreg, mode, nstep):
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=31, m=m, n=n, nstep=nstep)
    st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf={"bih": 0.01, "lap": 0.02, None: 0.0}[mode], bih=mode != "lap")
    util.deepen_dpmixl(st, n)
    want = _np_cnuity(cb, g, st, m, n, mxlkta=True)
    plain = _np_cnuity(cb, g, st, m, n, mxlkta=False)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    util.oracle_load_cnuity(ot, st)
    ot.set_i("mxlkta", 1)
    ot.cnuity(m, n, 1)
    inner = util.interior_sea(cb)
    got = ot.f64("dpmixl")[n - 1]
    assert np.array_equal(got[inner], want["dpmixl"][n - 1][inner])
    assert not np.array_equal(got[inner], plain["dpmixl"][n - 1][inner])
    for k in range(kdm):   # nothing else moves
        assert np.array_equal(ot.f64("dp")[n - 1, k][inner], plain["dp"][n - 1, k][inner])
    # the base of the mixed layer sits below layer 1 somewhere: more than one layer is exercised
    assert (st["dpmixl"][n - 1][inner] > st["dp"][n - 1, 0][inner]).any()
    ot.close()


def test_cnuity_across_the_arctic_c_oracle_equals_numpy(oracle):
    """nreg=2: every xctilr of cnuity goes through the tripole fold with its grid type (halo_ps, halo_us, halo_vs,
    halo_uv, halo_vv: mod_xc.F90:41-44), the interface-depth diffusion included"""
    m, n, kdm = 1, 2, 4
    cfg, sea, g, cb = util.make_arctic_case(96, 70, kdm, seed=17, m=m, n=n, nstep=4)
    st = util.arctic_halos_cnuity(g, util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=0.01, bih=True))
    want = _np_cnuity(cb, g, st, m, n)
    ot = util.oracle_tile_from_cb(oracle, cb, sea)
    util.oracle_load_cnuity(ot, st)
    ot.cnuity(m, n, 1)
    nb = g.nbdy
    inner = util.interior_sea(cb)
    iu_in = np.zeros_like(inner); iv_in = np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    for k in range(kdm):
        assert np.array_equal(ot.f64("dp")[n - 1, k][inner], want["dp"][n - 1, k][inner]), ("dp.n", k)
        assert np.array_equal(ot.f64("dp")[m - 1, k][inner], want["dp"][m - 1, k][inner]), ("dp.m", k)
        assert np.array_equal(ot.f64("uflx")[k][iu_in], want["uflx"][k][iu_in]), ("uflx", k)
        assert np.array_equal(ot.f64("vflx")[k][iv_in], want["vflx"][k][iv_in]), ("vflx", k)
        assert np.array_equal(ot.f64("p")[k + 1][inner], want["p"][k + 1][inner]), ("p", k)
    # the top rows feel the fold: the last interior row of dp(n) differs from a run on the same data as a
    # closed basin would give only through the halo, so just require finite, non-negative thicknesses there
    top = ot.f64("dp")[n - 1][:, nb + g.jj - 1, nb:nb + g.ii]
    assert np.isfinite(top).all() and (top >= 0.0).all()
    ot.close()
