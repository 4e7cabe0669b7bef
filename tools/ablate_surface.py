"""Where does the surface path (k_classify + k_surface_lists) spend its time? Same 16 x 1M-face vessels with waves / Jet+Pipe currents switched off in turn."""
import os, sys, numpy as np
os.environ["SFH_PROFILE_ROLES"] = "1"  # time the wet and the clip role of k_surface_lists in launches of their own
sys.path.insert(0, ".")
from stonefish_b200 import HydroContext, scenes
ctx = HydroContext(0)
sc = scenes.config5(16, 1000, 501, unique_meshes=False)
st = sc.bind(ctx)
ctx.set_profiling(True)
def run(tag):
    for _ in range(3): ctx.compute(sc.ocean, st)
    ts = []
    for _ in range(5):
        ctx.compute(sc.ocean, st)
        x = ctx.timings()
        ts.append((x.surface_ms, x.classify_ms, x.wet_ms, x.clip_ms))
    t, c, w, k = np.mean(ts, axis=0)
    fw, fk = x.chunks_wet * 255.1 / 1e6, x.chunks_clip * 255.1 / 1e6
    print(f"{tag:32s} surface {t:.3f} ms ({16e6/t/1e6:.1f} G faces/s) = classify {c:.3f} + wet {w:.3f} ({fw/w:.1f} G/s on {x.chunks_wet} chunks) + clip {k:.3f} ({fk/k:.1f} G/s on {x.chunks_clip})")
run("waves + Jet+Pipe")
sc.ocean.DisableCurrents(); run("waves, no currents")
sc.ocean.EnableCurrents(); sc.ocean.setWaves(None); run("flat, Jet+Pipe")
sc.ocean.DisableCurrents(); run("flat, no currents")
st2 = st.copy(); st2["flags"] = 1; run("flat, no currents (same)")
# localized fields: the same Jet, the Pipe moved 100 m away -> every chunk is outside both (chunk-level pre-test)
from stonefish_b200.hydro import Jet, Pipe
sc.ocean.currents = [Jet((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.3, 5.0), Pipe((-20.0, 100.0, 2.0), (20.0, 100.0, 2.0), 3.0, 5.0, 2.0, 2.0)]
sc.ocean.EnableCurrents(); run("flat, Jet + far Pipe")
