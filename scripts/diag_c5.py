"""diagnostic: NVE energy of 6mrr (f64) over 2000 steps, PME vs reaction-field cutoff, chunked vs single call"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import mbhelpers as H
import mollyb200 as mb
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "6mrr.npz")))
def etot(s):
    return mb.potential_energy(s), mb.kinetic_energy(s)
for label, mk in (("pme", lambda: H.sixmrr_pme_system(g, np.float64, exact=True, velocities=g["velocities_300K"].copy())),
                  ("crf", lambda: H.sixmrr_system(g, np.float64, r_list=1.2, velocities=g["velocities_300K"].copy()))):
    for cm in (1, 0):
        s = mk()
        pe, ke = etot(s)
        e0 = pe + ke
        out = []
        for k in range(10):
            mb.simulate(s, mb.VelocityVerlet(dt=0.0005, remove_CM_motion=cm), 100, init_step=100 * k)
            pe, ke = etot(s)
            out.append(round(pe + ke - e0, 3))
        print(label, "cm", cm, "chunked dE:", out, "rebuilds", s.stats()["n_rebuilds"], "graph", s.stats()["graph_mode"], flush=True)
        s.close()
    s = mk()
    pe, ke = etot(s); e0 = pe + ke
    mb.simulate(s, mb.VelocityVerlet(dt=0.0005), 1000)
    pe, ke = etot(s)
    print(label, "single call 1000 steps dE:", round(pe + ke - e0, 3), flush=True)
    s.close()
