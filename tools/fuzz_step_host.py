#!/usr/bin/env python3
"""Random cases for the host build of the kernel's step (tests/host_step.cpp) against the oracle, as long as one likes:

    python tools/fuzz_step_host.py [seed [trials]]

The generator is tests/test_step_host.py's (fuzz_trial): a state from some month of 2004, scheme switches combined at random,
a random cell, shifted forcing; every output value, the end state, the warnings and the fatal codes must be the oracle's
(bit for bit outside the RSL-derived diagnostics).  Needs no GPU."""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import test_step_host as T  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    so = os.path.join(tempfile.mkdtemp(prefix="host_step_"), "host_step.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(ROOT, "tests", "host_step.cpp")])
    H = C.CDLL(so)
    H.host_step_run.argtypes = [T.ip, T.dp, T.dp, T.dp, T.dp, T.dp, C.c_int, T.dp, T.dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(T.dp), C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.c_int, C.c_int, T.dp]
    snaps, met_all = T.year_snapshots(H)
    rng = np.random.default_rng(seed)
    t0, bad = time.time(), 0
    for i in range(trials):
        r = T.fuzz_trial(H, rng, snaps, met_all)
        if r:
            bad += 1
            print("case %d: %s" % (i, r))
    print("seed %d: %d cases, %d differ from the oracle, %.0f s" % (seed, trials, bad, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
