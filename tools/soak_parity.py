#!/usr/bin/env python
"""Randomised parity soak of the CUDA path against the CPU oracle: many seeds of the mixed scene family (every position, flag
combination and clipping class; flat / waves; none / Uniform / Jet + Pipe currents), each through both ways of the surface path
(direct mode and, with SFH_NO_DIRECT=1, the lists). Per-face codes must be equal (drag bits may flip under Jet / Pipe with
probability ~1e-8 per face: counted), sums within 1e-9. Test infrastructure (uses oracle/); run on the GPU box:
    python tools/soak_parity.py [first_seed] [n_seeds]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle
from stonefish_b200 import HydroContext, scenes
from tests.parity import REL_TOL, assert_counts_equal, max_rel_err, run_both

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pyoracle.build()
ctx = HydroContext(0)
worst, flips, faces, runs = 0.0, 0, 0, 0
for seed in range(first, first + count):
    for waves in (False, True):
        for currents in ("none", "uniform", "jetpipe"):
            sc = scenes.mixed_small(seed, waves, currents)
            for mode in ("0", "1"):
                os.environ["SFH_NO_DIRECT"] = mode
                gi, gc, ri, rc = run_both(ctx, pyoracle, sc)
                assert_counts_equal(gi, ri)
                assert np.array_equal(gc & 15, rc & 15), (seed, waves, currents, mode)
                d = int((gc != rc).sum())
                assert d == 0 or currents == "jetpipe", (seed, waves, currents, mode, d)
                flips += d
                err, where = max_rel_err(gi, ri)
                assert err <= REL_TOL, (seed, waves, currents, mode, err, where)
                worst = max(worst, err)
                faces += sc.faces_per_step()
                runs += 1
print(f"soak ok: seeds {first}..{first + count - 1}, {runs} steps, {faces} faces, worst relative error {worst:.3e}, drag-flag flips {flips}")
