#!/usr/bin/env python3
"""Differential fuzz: many synthetic cells, several seasons and scheme sets, CUDA library vs the CPU oracle
(test infrastructure; run on the GPU box).  Reports the worst relative differences of the seven parity flux sums
and of the end state.   usage: python tools/diff_fuzz.py [ncell] [ndays]"""
import os, sys, time
import multiprocessing as mp
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import fixtures as fx
import oracle_lib as ol
from suews_b200 import OUT_SUM82, SuewsBatch, synth

PARITY = [5, 6, 7, 8, 9, 16, 71]
STATE = ["state_id", "soilstore_id", "SnowPack", "SnowFrac", "SnowWater", "SnowDens", "alb", "qn1_av", "dqndt", "LAI_id",
         "GDD_id", "SDD_id", "HDD_id", "Tair_av", "porosity_id", "DecidCap_id", "SnowAlb", "IceFrac", "WUDay_id"]


def _oracle(args):
    d0, site, c, met = args
    dd = dict(d0)
    for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
        dd[k.lower()] = np.atleast_1d(site[k][c])
    m = met.copy()
    m[:, 11] += site["dTemp_C"][c]
    m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
    cell = ol.Cell(dd)
    r = cell.run(m, want=("SUEWS",))["SUEWS"][:, 5:]
    st = {k: cell.get_state(k) for k in STATE}
    return r.sum(axis=0), st


def main():
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ndays = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ol.build()
    base = fx.state_init(2005)
    allmet = fx.forcing(2004)
    worst_all = 0.0
    variants = [("default", {}), ("snow", {}), ("baserun", {})]
    if os.environ.get("SU_FUZZ_VARIANTS"):
        bio = dict(theta_bioco2=[0.96] * 3, alpha_enh_bioco2=[0.016] * 3, beta_enh_bioco2=[33.454] * 3)
        variants = [("baserun", dict(stabilitymethod=[2])), ("baserun", dict(stabilitymethod=[4])),
                    ("baserun", dict(netradiationmethod=[1])), ("baserun", dict(netradiationmethod=[2])),
                    ("baserun", dict(netradiationmethod=[11])), ("baserun", dict(netradiationmethod=[12])),
                    ("baserun", dict(netradiationmethod=[100])), ("baserun", dict(netradiationmethod=[200])),
                    ("baserun", dict(netradiationmethod=[300])), ("default", dict(netradiationmethod=[0])),
                    ("baserun", dict(roughlenheatmethod=[1])), ("baserun", dict(roughlenheatmethod=[3])),
                    ("baserun", dict(roughlenheatmethod=[4])), ("baserun", dict(roughlenmommethod=[1])),
                    ("baserun", dict(roughlenmommethod=[3])), ("default", dict(evapmethod=[1])),
                    ("default", dict(emissionsmethod=[4])), ("baserun", dict(emissionsmethod=[22], **bio)),
                    ("default", dict(emissionsmethod=[45], gsmodel=[4])), ("baserun", dict(gsmodel=[1])),
                    ("default", dict(waterusemethod=[1])), ("baserun", dict(ohmincqf=[1])),
                    ("default", dict(storageheatmethod=[0])), ("baserun", dict(aerodynamicresistancemethod=[3])),
                    ("baserun", dict(basetmethod=[2])), ("default", dict(smdmethod=[1]))]
    for scheme, over in variants:
        for day0 in (0, 59, 121, 196, 288, 340):
            met = allmet[day0 * 288:(day0 + ndays) * 288]
            b = SuewsBatch(ncell, 1)
            if over.get("smdmethod", [0])[0] > 0:
                met = met.copy(); met[:, 19] = 30.0
            if over.get("netradiationmethod", [0])[0] in (1, 11, 100):
                # observed-Ldown options need a physical series; the test forcing's column 16 is not one (12-20 W m-2),
                # and with it the iteration-based scheme 11 sits in a violently unstable regime that amplifies 1e-16
                met = met.copy(); met[:, 16] = 0.85 * 5.67e-8 * (met[:, 11] + 273.15) ** 4
            _m, site = synth.build_batch(b, base, cell0=17 * day0, scheme=scheme)
            for k, v in over.items():          # the variant's options win over the scheme set
                b.set_class_param(0, k, np.asarray(v, dtype=np.float64))
            b.upload(); b.set_forcing(met)
            b.advance(0, met.shape[0], OUT_SUM82)
            sums = b.sums()
            fatal, _w, _it = b.diag()
            d0 = synth.base_with_scheme(base, scheme)
            d0.update(over)
            with mp.get_context("fork").Pool(os.cpu_count()) as pool:
                res = pool.map(_oracle, [(d0, site, c, met) for c in range(ncell)])
            worst, wc, wk = 0.0, -1, ""
            for c, (rs, st) in enumerate(res):
                for col in PARITY:
                    e = abs(sums[c, col] - rs[col]) / max(abs(rs[col]), 1e-6 * np.abs(rs[PARITY]).max(), 1e-30)
                    if e > worst: worst, wc, wk = e, c, "sum col %d" % col
            ws, wsc, wsk = 0.0, -1, ""
            for k in STATE:
                g = b.get_state(k)
                for c, (rs, st) in enumerate(res):
                    e = np.max(np.abs(g[c] - st[k]) / np.maximum(np.abs(st[k]), 1e-3 * max(np.abs(st[k]).max(), 1e-6)))
                    if e > ws: ws, wsc, wsk = e, c, k
            print("%-8s %-34s day0 %3d: fatal %d  worst flux-sum rel %.2e (cell %d, %s)   worst state rel %.2e (cell %d, %s)" % (
                scheme, str(over)[:34], day0, int((fatal != 0).sum()), worst, wc, wk, ws, wsc, wsk), flush=True)
            worst_all = max(worst_all, worst)
            b.close()
    print("worst flux-sum relative difference overall: %.3e" % worst_all)


if __name__ == "__main__":
    main()
