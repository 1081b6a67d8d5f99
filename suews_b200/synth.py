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
    if SCHEMES[scheme]["snowuse"] == 0:
        # a snow-free start: the BaseRun dump starts with snow on the ground
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    return d


def build_batch(batch, base, cell0=0, scheme="default", offsets=True):
    """Fill a SuewsBatch (all its cells) from the base cell + per-cell synthetic site values."""
    d = base_with_scheme(base, scheme)
    missing = batch.load_reference_args(d, cls=0, cells=True)
    site = cell_site_arrays(cell0, batch.ncell, d, offsets=offsets)
    for k, v in site.items():
        batch.set_cell_param(k, v if v.ndim == 2 else v.reshape(-1, 1))
    return missing, site
