"""Met-forcing preparation: the step immediately before the hot path.

Host-side (numpy, float64) restatement of the reference's forcing reader and
hourly -> model-timestep disaggregation:

* ``read_met``           MetRead, SUEWS-SourceCode/suews_ctrl_input.f95:13-135
                         (kPa -> hPa :109, xsmd = -999 when SMDMethod == 0 :96-102)
* ``disaggregate_met``   DisaggregateMet :1380-1640 with
  ``_disagg_datetime``   DisaggregateDateTime :1849-1900
  ``_disagg_lin``        Disagg_Lin :1908-2011 (type 10 period averages, type 20 instantaneous)
  ``_disagg_p_among_n``  DisaggP_amongN :2016-2068 (RainDisaggMethod = 100: even spread)

Output is the [len_sim, 24] ``MetForcingBlock`` consumed by ``suews_cal_multitsteps``
(column order suews_ctrl_driver.f95:3177-3201).  Arithmetic order follows the
reference expression by expression so that the result is bit-identical to the
in-memory array the Fortran binary uses (its text dump only keeps 4-6 decimals).
"""
import numpy as np

NCOL_MET = 24


def read_met(path, smd_method=0):
    """Read an original-resolution SUEWS met file (InputMetFormat = 10)."""
    rows = []
    with open(path) as fh:
        fh.readline()
        for ln in fh:
            t = ln.split()
            if len(t) < NCOL_MET or t[0].startswith("-9"):
                continue
            rows.append([float(x) for x in t[:NCOL_MET]])
    a = np.array(rows, dtype=np.float64)
    a[:, 12] = a[:, 12] * 10.0  # Pres_hPa = Pres_kPa*10.
    if smd_method == 0:
        a[:, 19] = -999.0
    return a


def _days_in_year(y):
    y = np.asarray(y)
    return np.where(((y % 4 == 0) & (y % 100 != 0)) | (y % 400 == 0), 366, 365)


def _nint(x):
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(np.int64)


def _disagg_datetime(dt4, tstep, nper):
    n = dt4.shape[0]
    out = np.zeros((n * nper, 4))
    # 24.0 and 60.0*24.0 are REAL(4) constants (exact); tstep/60.0 is a REAL(4) quotient
    dectime_orig = (dt4[:, 1] - 1) + dt4[:, 2] / 24.0 + dt4[:, 3] / (60.0 * 24.0)
    step = np.float64(np.float32(np.float32(tstep) / np.float32(60.0)) / np.float32(60.0 * 24.0))
    seq = np.arange(nper - 1, -1, -1)
    for i in range(n):
        # (tstep/60.0)/(60.0*24.0)*ii is evaluated in REAL(4) (integer ii promoted), then widened
        off = (np.float32(step) * seq.astype(np.float32)).astype(np.float64)
        td = dectime_orig[i] - off
        # NINT(x*60*60*24)/(60*60*24*1.): integer / REAL(4) -> REAL(4) quotient
        td = (_nint(td * 60 * 60 * 24).astype(np.float32) / np.float32(86400.0)).astype(np.float64)
        iy = np.where(td < 0, int(dt4[i, 0]) - 1, int(dt4[i, 0]))
        nd = _days_in_year(iy)
        td = np.where(td < 0, td + nd, td)
        idd = np.floor(td).astype(np.int64) + 1
        ihm = _nint((td + 1 - idd / 1.0) * 60.0 * 24.0)
        ih = (ihm - ihm % 60) // 60
        ih = np.where(ih < 24, ih, 0)
        im = ihm % 60
        sl = slice(nper * i, nper * (i + 1))
        out[sl, 0], out[sl, 1], out[sl, 2], out[sl, 3] = iy, idd, ih, im
    return out


def _disagg_lin(slow, slow_prev, slow_next, dtype, nper, first_block=True, last_block=True):
    n = slow.shape[0]
    fast = np.full(n * nper, -999.0)
    xn = 2 if nper % 2 == 0 else 1
    seq = np.arange(1, nper + 1)
    half = int(np.floor(nper / 2.0))
    if dtype == 10:
        rows = half + seq  # 1-based
        for i in range(n - 1):
            d = slow[i + 1] - slow[i]
            fast[nper * i + rows - 1] = slow[i] - d / (xn * nper) + d / nper * seq
        nfirst = rows[0] - 1
        if first_block:
            fast[:nfirst] = fast[rows[0] - 1]
        else:
            d = slow[0] - slow_prev
            ii = np.arange(nper - nfirst + 1, nper + 1)
            fast[:nfirst] = slow_prev - d / (xn * nper) + d / nper * ii
        last_start = nper * (n - 2) + rows[-1]  # 0-based index of first unfilled row
        nlast = n * nper - last_start
        if last_block:
            fast[last_start:] = fast[last_start - 1]
        else:
            d = slow_next - slow[n - 1]
            fast[last_start:] = slow[n - 1] - d / (xn * nper) + d / nper * np.arange(1, nlast + 1)
    elif dtype == 20:
        rows = nper + seq
        for i in range(n - 1):
            d = slow[i + 1] - slow[i]
            fast[nper * i + rows - 1] = (slow[i] + d / nper * 2 * (seq - 1) + slow[i]) / 2
        nfirst = rows[0] - 1
        if first_block:
            fast[:nfirst] = fast[rows[0] - 1]
        else:
            d = slow[0] - slow_prev
            ii = np.arange(nper - nfirst + 1, nper + 1)
            fast[:nfirst] = (slow_prev + d / nper * 2 * (ii - 1) + slow_prev) / 2
    else:
        raise ValueError("DisaggType must be 10 or 20")
    if np.any(fast == -999.0):
        raise RuntimeError("Problem in SUEWS_MetDisagg: -999 values in disaggregated data.")  # ErrorHint 13
    return fast


def _random_samples(n, out_of, rng):
    """RandomSamples (suews_ctrl_input.f95:2147-2177): n distinct sub-intervals of 1..out_of by rejection, X = INT(r*OutOf) + 1
    with r uniform on [0, 1).  The reference draws r from an unseeded RANDOM_NUMBER, so its own runs do not repeat; here
    the caller's generator decides."""
    samples = []
    while len(samples) < n:
        x = int(rng.random() * out_of) + 1
        if x not in samples:
            samples.append(x)
    return samples


def _disagg_p_among_n(slow, among_n, nper, rng=None):
    """DisaggP_amongN (:2016-2068): each interval's amount in among_n of its nper sub-intervals, equal shares -- all of
    them for among_n = nper (RainDisaggMethod 100), a random draw otherwise (101)."""
    if among_n > nper:
        raise ValueError("ErrorHint 2: no. of rainy periods cannot exceed number of subintervals")
    if among_n != nper and rng is None:
        raise NotImplementedError("RainDisaggMethod 101/102 draw random sub-intervals: pass rng")
    fast = np.zeros(slow.shape[0] * nper)
    for i in range(slow.shape[0]):
        if slow[i] > 0:
            if among_n == nper:
                fast[nper * i:nper * (i + 1)] = slow[i] / among_n
            else:
                for x in _random_samples(among_n, nper, rng):
                    fast[nper * i + x - 1] = slow[i] / among_n
    return fast


def _disagg_p_among_n_mult(slow, upper_i, among_n5, nper, rng):
    """DisaggP_amongNMult (:2070-2145, RainDisaggMethod 102): the number of wet sub-intervals depends on the interval's
    amount through five intensity bins (MultRainAmongNUpperI / MultRainAmongN of RunControl.nml)."""
    fast = np.zeros(slow.shape[0] * nper)
    for i in range(slow.shape[0]):
        if slow[i] > 0:
            k = next((j for j in range(5) if slow[i] <= upper_i[j]), 4)     # above the last bound: the last bin (ErrorHint 4, a warning)
            n = int(among_n5[k])
            if n > nper:
                raise ValueError("ErrorHint 2: no. of rainy periods cannot exceed number of subintervals")
            xs = range(1, nper + 1) if n == nper else _random_samples(n, nper, rng)
            for x in xs:
                fast[nper * i + x - 1] = slow[i] / n
    return fast


def read_estm_ts(path):
    """An ESTM surface-temperature file (iy id it imin Tiair Tsurf Troof Troad Twall Twall_n/e/s/w), any resolution."""
    rows = []
    with open(path) as fh:
        fh.readline()
        for ln in fh:
            t = ln.split()
            if len(t) >= 13 and not t[0].startswith("-9"):
                rows.append([float(x) for x in t[:13]])
    return np.array(rows, dtype=np.float64)


def disaggregate_estm(ts_orig, tstep=300, resolution_in=3600, disagg_method=1):
    """DisaggregateESTM (suews_ctrl_input.f95:1650-1849), whole-file semantics: the date/time and Disagg_Lin of the met
    block applied to the nine temperature columns; DisaggMethodESTM 1 = period averages, 2 = instantaneous values."""
    ts_orig = np.asarray(ts_orig, dtype=np.float64)
    if disagg_method not in (1, 2):
        raise ValueError("DisaggMethodESTM value should be 1 or 2")  # ErrorHint 2
    nper = resolution_in // tstep
    out = np.full((ts_orig.shape[0] * nper, 13), -999.0)
    out[:, 0:4] = _disagg_datetime(ts_orig[:, 0:4], tstep, nper)
    for c in range(4, 13):
        if not np.all(ts_orig[:, c] == -999):
            out[:, c] = _disagg_lin(ts_orig[:, c], -999.0, -999.0, 10 if disagg_method == 1 else 20, nper)
    return out


def disaggregate_met(met_orig, tstep=300, resolution_in=3600, disagg_method=1, rain_disagg_method=100,
                     first_block=True, last_block=True, prev_row=None, next_row=None, kdown_zen=0, zenith_fn=None,
                     rain_among_n=None, mult_rain_among_n=None, mult_rain_among_n_upper_i=None, rng=None):
    """Disaggregate an original-resolution met block (already passed through ``read_met``).

    kdown_zen: RunControl's KdownZen.  1 = zero the disaggregated K-down while the sun is below the site's horizon and
    rescale each day to the original daily sum (suews_ctrl_input.f95:1568-1600).  That needs the sun position, which this
    package only has on the device (SuewsBatch.set_forcing_orig(kdown_zen=1, site=...) does the whole thing there); here the
    caller supplies zenith_fn(year, dectime) -> zenith [deg] for the forcing's site, e.g. a binding of
    suews_cal_sunposition_c.  Asking for KdownZen = 1 without it raises instead of returning the KdownZen = 0 series."""
    # RainDisaggMethod 101 (rain_among_n = RainAmongN) / 102 (the five bins MultRainAmongN, MultRainAmongNUpperI) put each
    # interval's rain into randomly chosen sub-intervals: rng is a numpy Generator or a seed.  The reference's generator is
    # unseeded, so there is no reference sequence to reproduce -- only the distribution (amounts, number of wet sub-steps).
    if rain_disagg_method in (101, 102):
        if rng is None:
            raise ValueError("RainDisaggMethod %d is random: pass rng (a numpy Generator or a seed)" % rain_disagg_method)
        rng = rng if hasattr(rng, "random") else np.random.default_rng(rng)
        if rain_disagg_method == 101 and rain_among_n in (None, -999):
            raise ValueError("ErrorHint 2: RainDisaggMethod 101 requires RainAmongN")
        if rain_disagg_method == 102 and (mult_rain_among_n is None or mult_rain_among_n_upper_i is None):
            raise ValueError("ErrorHint 2: RainDisaggMethod 102 requires MultRainAmongN and MultRainAmongNUpperI")
    if int(kdown_zen) != 0 and zenith_fn is None:
        raise NotImplementedError("KdownZen = 1 (zenith re-weighting of K-down, suews_ctrl_input.f95:1568-1600) needs the sun position: "
                                  "pass zenith_fn, or let the device do it (SuewsBatch.set_forcing_orig(kdown_zen=1, site=...))")
    met_orig = np.asarray(met_orig, dtype=np.float64)
    nper = resolution_in // tstep
    n = met_orig.shape[0]
    method = np.full(NCOL_MET, 10)
    if disagg_method == 2:
        method[:] = 20
    elif disagg_method == 3:
        method[9:13] = 20
    elif disagg_method != 1:
        raise ValueError("DisaggMethod value should be 1, 2, or 3")  # ErrorHint 2
    if rain_disagg_method not in (100, 101, 102):
        raise ValueError("Disaggregation code for rain not recognised")

    def rain(col):
        if rain_disagg_method == 100:
            return _disagg_p_among_n(col, nper, nper)
        if rain_disagg_method == 101:
            return _disagg_p_among_n(col, int(rain_among_n), nper, rng)
        return _disagg_p_among_n_mult(col, list(mult_rain_among_n_upper_i), list(mult_rain_among_n), nper, rng)
    out = np.full((n * nper, NCOL_MET), -999.0)
    out[:, 0:4] = _disagg_datetime(met_orig[:, 0:4], tstep, nper)
    for c in range(4, NCOL_MET):
        if c == 13:
            out[:, 13] = rain(met_orig[:, 13])
            if np.all(met_orig[:, 15] == -999):
                out[:, 15] = -999.0
            else:
                out[:, 15] = rain(met_orig[:, 15])
        elif c == 23:
            continue  # wind direction: not disaggregated by the reference
        elif np.all(met_orig[:, c] == -999):
            out[:, c] = -999.0
        else:
            sp = prev_row[c] if prev_row is not None else -999.0
            sn = next_row[c] if next_row is not None else -999.0
            out[:, c] = _disagg_lin(met_orig[:, c], sp, sn, method[c], nper, first_block, last_block)
    if int(kdown_zen) == 1:
        out[:, 14] = apply_kdown_zen(out, zenith_fn, tstep)
    if np.all(out[:, 15] == -999):
        out[:, 15] = 0.0
    return out


def apply_kdown_zen(met_fast, zenith_fn, tstep=300):
    """KdownZen = 1 (suews_ctrl_input.f95:1568-1600) on a disaggregated block: K-down (column 15) zero wherever
    zenith_fn(year, dectime) > 90 at dectime - halftimestep -- with the day of year where the year belongs, as the
    reference calls NARP_cal_SunPosition(Met_tt(i, 2), ...) -- then every block of nsd = 24 * nsh rows rescaled to its
    original sum (sequential sums, rows after the last whole block left masked).  Returns the new column."""
    k = np.array(met_fast[:, 14], dtype=np.float64)
    adj = k.copy()
    half = float(np.float32(np.float32(tstep) * np.float32(1.0)) / np.float32(2) / np.float32(24 * 3600))
    dectime = met_fast[:, 1] + met_fast[:, 2] / 24.0 + met_fast[:, 3] / float(np.float32(60.0) * np.float32(24.0))
    for i in range(k.size):
        if zenith_fn(float(met_fast[i, 1]), float(dectime[i] - half)) > 90:
            adj[i] = 0.0
    nsd = 24 * (3600 // tstep)
    for d in range(k.size // nsd):
        s1 = s2 = 0.0
        for j in range(d * nsd, (d + 1) * nsd):
            s1 = s1 + k[j]
            s2 = s2 + adj[j]
        with np.errstate(invalid="ignore", divide="ignore"):
            adj[d * nsd:(d + 1) * nsd] = adj[d * nsd:(d + 1) * nsd] * s1 / s2
    return adj
