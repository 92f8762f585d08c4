"""Cases and digests shared by the three parties that must agree bit for bit on mod_asselin.F90 and cnuity.F90:
the reference's own source text (tests/golden/make_reference_text_vectors.py writes its digests to
tests/golden/from_reference_text.json), the CPU oracle (tests/test_reference_text.py) and the CUDA path
(tests/test_parity_gpu.py, tests/test_cnuity_gpu.py)."""
import hashlib

import numpy as np

import util

# the configurations are those of the GPU parity tests (test_asselin_device_matches_oracle, test_cnuity_*), so that the
# device is checked against the reference text on cases of the size it is checked against the oracle
ASSELIN = {
    # name: (nreg, sigver, ntracr, extra, mxlmy)
    "asselin:17t_tracers_mxlmy": (0, 6, 2, {}, True),
    "asselin:12t_th3d_periodic": (3, 8, 0, {"advflg": 1}, False),
    "asselin:12t_isopycnal_below_1": (0, 8, 1, {"nhybrd": 1}, False),
    "asselin:12t_sigma0_isopyc": (0, 7, 0, {"isopyc": True, "hybrid": False, "nhybrd": 0}, False),
}

CNUITY = {
    # name: (itdm, jtdm, kdm, nreg, m, n, isopyc, thkdf, bih, nstep, mxlkta, seed)
    "cnuity:periodic_swapped": (131, 77, 3, 3, 2, 1, False, 0.0, True, 3, False, 23),
    "cnuity:isopyc": (64, 90, 3, 1, 1, 2, True, 0.0, True, 3, False, 23),
    "cnuity:fplane": (70, 45, 2, 4, 1, 2, False, 0.0, True, 3, False, 23),
    "cnuity:closed_thkdf4_down": (90, 70, 5, 0, 1, 2, False, 0.01, True, 4, False, 29),
    "cnuity:isopyc_thkdf4_up": (64, 90, 4, 1, 1, 2, True, 0.01, True, 3, False, 29),
    "cnuity:thkdf2": (131, 77, 3, 3, 1, 2, False, 0.02, False, 2, False, 29),
    "cnuity:fplane_thkdf2_odd": (70, 45, 4, 4, 1, 2, False, 0.02, False, 5, False, 29),
    "cnuity:mxlkta_thkdf2": (64, 90, 4, 1, 1, 2, False, 0.02, False, 3, True, 31),
    "cnuity:mxlkta": (131, 77, 4, 3, 1, 2, False, 0.0, True, 2, True, 31),
    "cnuity:arctic_thkdf4": (96, 70, 4, 2, 1, 2, False, 0.01, True, 4, False, 17),
}


def _sha(parts):
    """(`+ 0.0` turns -0.0 into +0.0: which zero `max(0., -0.)` returns is the one thing the three parties may differ
    in - gfortran keeps the first argument on a tie, a C `a > b ? a : b` the second - and a mass flux of -0.0 is 0.0)"""
    h = hashlib.sha256()
    for a in parts:
        h.update((np.ascontiguousarray(a).astype("<f8") + 0.0).tobytes())
    return h.hexdigest()[:32]


def asselin_case(name):
    nreg, sigver, ntracr, extra, mxlmy = ASSELIN[name]
    m, n = 1, 2
    cfg, sea, g, cb = util.make_case(90, 61, 4, nreg=nreg, ntracr=ntracr, seed=41, **extra)
    if mxlmy:
        util.add_q2(cfg, sea, g, cb, m, n)
    util.add_asselin(cfg, sea, g, cb, m, n, sigver=sigver)
    return cfg, sea, g, cb, m, n


def asselin_digest(cb, m, fields):
    """fields: name -> array with the slot axis as in CbArrays (temp/saln/th3d/dp: (2,kdm,..); tracer: (ntracr,2,kdm,..));
    slot m (the filtered time level) on the interior sea cells"""
    msk = util.interior_sea(cb)
    out = {}
    for name in ("temp", "saln", "th3d", "dp"):
        out[name] = _sha([fields[name][m - 1][:, msk]])
    if cb.ntracr:
        out["tracer"] = _sha([fields["tracer"][q, m - 1][:, msk] for q in range(cb.ntracr)])
    return out


def cnuity_case(name):
    itdm, jtdm, kdm, nreg, m, n, isopyc, thkdf, bih, nstep, mxlkta, seed = CNUITY[name]
    extra = dict(isopyc=True, hybrid=False, nhybrd=0) if isopyc else {}
    if nreg == 2:
        cfg, sea, g, cb = util.make_arctic_case(itdm, jtdm, kdm, seed=seed, m=m, n=n, nstep=nstep, **extra)
        st = util.arctic_halos_cnuity(g, util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih))
    else:
        cfg, sea, g, cb = util.make_case(itdm, jtdm, kdm, nreg=nreg, seed=seed, m=m, n=n, nstep=nstep, **extra)
        st = util.add_cnuity(cfg, sea, g, cb, m, n, thkdf=thkdf, bih=bih)
    if mxlkta:
        util.deepen_dpmixl(st, n)
    return cfg, sea, g, cb, st, m, n, isopyc, mxlkta


def cnuity_digest(cb, m, n, dp, uflx, vflx, p, dpmixl_n=None):
    """dp: (2,kdm,..) both time levels; uflx, vflx: (kdm,..); p: (kdm+1,..) interfaces; on the interior sea cells /
    u faces / v faces"""
    g = cb.geom
    nb = g.nbdy
    inner = util.interior_sea(cb)
    iu_in, iv_in = np.zeros_like(inner), np.zeros_like(inner)
    iu_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iu[nb:nb + g.jj, nb:nb + g.ii] != 0
    iv_in[nb:nb + g.jj, nb:nb + g.ii] = cb.iv[nb:nb + g.jj, nb:nb + g.ii] != 0
    out = dict(dp_n=_sha([dp[n - 1][:, inner]]), dp_m=_sha([dp[m - 1][:, inner]]), uflx=_sha([uflx[:, iu_in]]),
               vflx=_sha([vflx[:, iv_in]]), p=_sha([p[1:][:, inner]]))
    if dpmixl_n is not None:
        out["dpmixl_n"] = _sha([dpmixl_n[inner]])
    return out
