"""Synthetic many-cell configurations: the reference's own BaseRun cell replicated with
per-cell surface fractions, location and morphology, mirroring how the reference's multi-grid test
perturbs SUEWS_SiteSelect.txt (Test/code/Test_SUEWS.py:138-173: sfr = normalise(base + 0.5 u)).
Everything is derived from a fixed seed and the global cell index, so any rank can build its slice."""
import numpy as np

# RunControl of the user-facing default set (Release/InputTables/2018c/RunControl.nml):
SCHEMES = {
    "default": dict(snowuse=0, netradiationmethod=3, emissionsmethod=2, storageheatmethod=1, stabilitymethod=3,
                    roughlenheatmethod=2, roughlenmommethod=2, smdmethod=0, waterusemethod=0),
    # Test/BaseRun/2020b/RunControl.nml (the golden run)
    "baserun": dict(snowuse=1, netradiationmethod=13, emissionsmethod=1, storageheatmethod=1, stabilitymethod=3,
                    roughlenheatmethod=2, roughlenmommethod=2, smdmethod=0, waterusemethod=0),
    "snow": dict(snowuse=1, netradiationmethod=3, emissionsmethod=2, storageheatmethod=1, stabilitymethod=3,
                 roughlenheatmethod=2, roughlenmommethod=2, smdmethod=0, waterusemethod=0),
}
# the scheme sweep of BASELINE.json configs[3] / SURVEY 8d(iv): one switch changed against the default set
for _name, _over in (("estm", dict(storageheatmethod=4)), ("anohm", dict(storageheatmethod=3)), ("stab2", dict(stabilitymethod=2)), ("stab4", dict(stabilitymethod=4)),
                     ("netrad13", dict(netradiationmethod=13)), ("netrad300", dict(netradiationmethod=300))):
    SCHEMES[_name] = dict(SCHEMES["default"], **_over)

C2K = float(np.float32(273.15))

# AnOHM columns (AnOHM_Cp, AnOHM_Kk, AnOHM_Ch) of Test/BaseRun/2020b/Input/SUEWS_NonVeg.txt (codes 551, 552, 553),
# SUEWS_Veg.txt and SUEWS_Water.txt; the 2020b argument dump carries zeros because AnOHM is off in that run
ANOHM_EXAMPLE = dict(cpanohm=[2e6, 2e6, 1e6, 1e6, 1e6, 2e6, 4.1e6], kkanohm=[1.2] * 7, chanohm=[4.0] * 7)


def estm_example(surface_area, **over):
    """An ESTM (StorageHeatMethod 4) parameter set as the reference's SUEWS_Translate / ESTM_initials leave it in
    MODULE ESTM_data for the BaseRun site: building class 801 and paved class 807 of
    Test/BaseRun/2020b/Input/SUEWS_ESTMCoefficients.txt, ESTMinput.nml of the same directory, the internal view
    factors and floor temperature hard-coded in ESTM_translate (suews_phys_estm.f95:673-692).  AreaWall is -999 in
    that SiteSelect row (ESTM is off in the BaseRun); 0.4 x SurfaceArea is used here."""
    def pad(v):
        return list(v) + [-999.0] * (5 - len(v))
    f32 = lambda x: float(np.float32(x))   # noqa: E731
    d = dict(
        evolveTibld=[0], TsurfChoice=[0], ibldCHmod=[0], Nibld=[3], Nwall=[3], Nroof=[3], Nground=[4],
        LBC_soil=[8.5], THEAT_ON=[18.0 + C2K], THEAT_OFF=[22.0 + C2K], THEAT_fix=[19.0 + C2K],
        ivf_iw=[f32(0.1)], ivf_ir=[0.0], ivf_ii=[f32(0.9)], ivf_if=[0.0],
        ivf_ww=[f32(0.05)], ivf_wr=[0.0], ivf_wi=[f32(0.95)], ivf_wf=[0.0],
        ivf_rw=[f32(0.05)], ivf_ri=[f32(0.95)], ivf_rf=[0.0],
        ivf_fw=[f32(0.05)], ivf_fr=[0.0], ivf_fi=[f32(0.95)],
        zibld=pad([0.035] * 3), kibld=pad([0.93] * 3), ribld=pad([1500000.0] * 3),
        zwall=pad([0.087, 0.0497, 0.0497]), kwall=pad([0.760645, 0.93, 0.93]), rwall=pad([1046908.363, 1500000.0, 1500000.0]),
        zroof=pad([0.05, 0.15, 0.12]), kroof=pad([1.142, 0.93, 0.05]), rroof=pad([1606620.0, 1500000.0, 290000.0]),
        zground=pad([0.05, 0.10, 0.5, 0.7]), kground=pad([1.16, 1.16, 0.63, 0.63]),
        rground=pad([2285607.0, 2285607.0, 1200000.0, 1200000.0]),
        nroom=[10.0], alb_ibld=[0.5], em_ibld=[1.0], CH_iroof=[0.0015], CH_iwall=[0.0015], CH_ibld=[0.0015],
        AreaWall=[0.4 * float(surface_area)], Tfloor=[f32(20.0) + C2K],
    )
    for k, v in over.items():
        d[k] = list(np.atleast_1d(v))
    return d


def estm_ts_forcing(met, tiair=20.0, offsets=(0.0, 1.0, 2.0, 0.5, -0.5, 0.3, 0.8, 0.1)):
    """Synthetic ESTM surface-temperature forcing [len_sim, 13] (iy id it imin Tiair Tsurf Troof Troad Twall
    Twall_n Twall_e Twall_s Twall_w): the air temperature of the met block plus a fixed offset per column
    ("synthetic Ts forcing = air temperature")."""
    met = np.asarray(met)
    ts = np.zeros((met.shape[0], 13))
    ts[:, :4] = met[:, :4]
    ts[:, 4] = tiair
    for j, off in enumerate(offsets):
        ts[:, 5 + j] = met[:, 11] + off
    return ts


def cell_site_arrays(cell0, ncell, base, seed=2020, offsets=True):
    """Per-cell site values for global cells [cell0, cell0 + ncell)."""
    # one generator per block of 4096 cells keyed by (seed, block): cheap, order-independent
    out = {k: [] for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z", "dTemp_C", "dRH")}
    blk = 4096
    b0, b1 = cell0 // blk, (cell0 + ncell - 1) // blk
    base_sfr = np.asarray(base["sfr"], dtype=np.float64)
    for b in range(b0, b1 + 1):
        rng = np.random.default_rng([seed, b])
        u = rng.random((blk, 7))
        sfr = base_sfr[None, :] + 0.5 * u
        sfr = sfr / sfr.sum(axis=1, keepdims=True)
        lat = 35.0 + 30.0 * rng.random(blk)
        lng = -10.0 + 40.0 * rng.random(blk)
        h = 3.0 + 27.0 * rng.random(blk)
        dT = -2.0 + 4.0 * rng.random(blk)
        dRH = -5.0 + 10.0 * rng.random(blk)
        lo = max(cell0, b * blk) - b * blk
        hi = min(cell0 + ncell, (b + 1) * blk) - b * blk
        out["sfr"].append(sfr[lo:hi])
        out["lat"].append(lat[lo:hi])
        out["lng"].append(lng[lo:hi])
        out["bldgH"].append(h[lo:hi])
        out["EveTreeH"].append(0.8 * h[lo:hi])
        out["DecTreeH"].append(0.8 * h[lo:hi])
        out["Z"].append(np.maximum(50.0, 3.0 * h[lo:hi]))
        out["dTemp_C"].append(dT[lo:hi] if offsets else np.zeros(hi - lo))
        out["dRH"].append(dRH[lo:hi] if offsets else np.zeros(hi - lo))
    return {k: np.concatenate(v) for k, v in out.items()}


def base_with_scheme(base, scheme="default"):
    d = {k.lower(): v for k, v in base.items()}
    for k, v in SCHEMES[scheme].items():
        d[k] = [v]
    if SCHEMES[scheme]["storageheatmethod"] == 3:
        d.update({k: list(v) for k, v in ANOHM_EXAMPLE.items()})
    if SCHEMES[scheme]["snowuse"] == 0:
        # a snow-free start: the BaseRun dump starts with snow on the ground
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    return d


def build_batch(batch, base, cell0=0, scheme="default", offsets=True, ts_row0=None):
    """Fill a SuewsBatch (all its cells) from the base cell + per-cell synthetic site values.  With an ESTM scheme,
    ts_row0 (first row of the surface-temperature forcing) initialises the element temperatures as ESTM_translate
    does on its first call."""
    d = base_with_scheme(base, scheme)
    if SCHEMES[scheme]["storageheatmethod"] in (4, 14):
        from . import estm
        ep = estm_example(d["surfacearea"][0])
        d.update({k.lower(): v for k, v in ep.items()})
        # per-cell surface fractions only move lup_ground, a longwave term no output or later state depends on
        # (radforce = .FALSE., suews_phys_estm.f95:1191); the base cell's value is used for every cell
        d.update({k.lower(): v for k, v in estm.initial_state(ep, d["sfr"], d["emis"], ts_row0).items()})
    missing = batch.load_reference_args(d, cls=0, cells=True)
    site = cell_site_arrays(cell0, batch.ncell, d, offsets=offsets)
    for k, v in site.items():
        batch.set_cell_param(k, v if v.ndim == 2 else v.reshape(-1, 1))
    return missing, site
