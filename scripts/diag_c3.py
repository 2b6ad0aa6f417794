#!/usr/bin/env python
"""Energy / temperature trace of the full 6mrr system for several engine settings (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import mbhelpers as H
import mollyb200 as mb

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "6mrr.npz")))
for dtype, thermo, nsteps_policy, label in ((np.float64, False, 0, "f64 NVE triggered"), (np.float32, False, 0, "f32 NVE triggered"),
                                            (np.float32, False, 10, "f32 NVE every10"), (np.float32, True, 0, "f32 Andersen triggered"),
                                            (np.float64, True, 0, "f64 Andersen triggered")):
    if os.environ.get("DIAG_NOGRAPH"):
        label += " nograph"
    s = H.sixmrr_system(g, dtype, r_list=1.12, n_steps=nsteps_policy)
    sim = mb.VelocityVerlet(dt=0.0005, coupling=mb.AndersenThermostat(300.0, 1.0) if thermo else None)
    rng = np.random.default_rng(3)
    out = []
    for k in range(6):
        f, pe = mb.forces_energy(s)
        ke = mb.kinetic_energy(s)
        out.append(f"step {100 * k}: T={mb.temperature(s):.1f} PE={pe:.1f} KE={ke:.1f} E={pe + ke:.1f} |F|max={np.abs(f).max():.1f}")
        mb.simulate(s, sim, 100, init_step=100 * k, rng=rng)
    print(label, "graph", s.stats()["graph_mode"], "rebuilds", s.stats()["n_rebuilds"])
    print("   " + "\n   ".join(out), flush=True)
    s.close()
