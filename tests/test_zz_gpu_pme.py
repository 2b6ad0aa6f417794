"""PME reciprocal space on the device against OpenMM's forces_all_pme_exact for 6mrr.

The CUDA side (csrc/pme.cuh, mb_set_pme) was written after the round's GPU budget was spent, so this test has never run;
it is marked xfail (non-strict: an XPASS is the good outcome) and runs in a child process so that a device fault in the
new code cannot take the rest of the suite down with it. The checker it mirrors, oracle/pme.py, is pinned on the CPU
(tests/test_oracle.py::test_6mrr_all_pme_openmm_golden)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(tests)r); sys.path.insert(0, %(root)r)
import mbhelpers as H, mollyb200 as mb
g = dict(np.load(%(golden)r))
sd = H.sixmrr_description(g)
atoms = mb.atoms_from_arrays(sd["mass"], sd["charge"], sd["sigma"], sd["eps"], np.float64)
inters = (mb.LennardJones(cutoff=mb.DistanceCutoff(1.0), use_neighbors=True, weight_special=float(g["lj14scale"])),
          mb.CoulombEwald(dist_cutoff=1.0, error_tol=0.0005, use_neighbors=True, weight_special=float(g["coulomb14scale"])))
nf = mb.GPUNeighborFinder(dist_cutoff=1.2, excluded_pairs=g["excluded"] + 1, special_pairs=g["special"] + 1)
pme = mb.PME(dist_cutoff=1.0, error_tol=0.0005, excluded_pairs=np.concatenate([g["excluded"], g["special"]]) + 1)
s = mb.System(atoms=atoms, coords=sd["coords"], boundary=mb.CubicBoundary(*sd["box"]), velocities=sd["velocities"],
              pairwise_inters=inters, neighbor_finder=nf, dtype=np.float64, specific_inter_lists=H.sixmrr_specific_lists(g),
              general_inters=(pme,))
f, e = mb.forces_energy(s)
from oracle import oracle as o
e += o.lj_dispersion_correction_energy(g["sigma"], g["eps"], g["box"], 1.0)
err = np.linalg.norm(f - g["forces_all_pme_exact"], axis=1).max()
print("max|dF| =", err, "dE =", e - float(g["energy_all_pme_exact"]))
assert err < 1e-6 and abs(e - float(g["energy_all_pme_exact"])) < 1e-3   # reference: 1e-7 / 1e-5 on the CPU in f64
# the reference's small PME case (test/interactions.jl:1683-1697): 3 waters, orthorhombic box, all-pairs path
w = dict(np.load(%(water)r))
for dt_, tol_f, tol_e in ((np.float64, 1e-6, 1e-6), (np.float32, 5e-4, 2e-4)):   # f32: the reference's own tolerances
    atoms = mb.atoms_from_arrays(w["mass"], w["charge"], w["sigma"], w["eps"], dt_)
    s = mb.System(atoms=atoms, coords=w["coords"].astype(dt_), boundary=mb.CubicBoundary(*w["box"]),
                  pairwise_inters=(mb.CoulombEwald(dist_cutoff=0.9, error_tol=0.0005, use_neighbors=True),),
                  neighbor_finder=mb.GPUNeighborFinder(dist_cutoff=0.9, excluded_pairs=w["excluded"] + 1),
                  dtype=dt_, general_inters=(mb.PME(dist_cutoff=0.9, error_tol=0.0005, excluded_pairs=w["excluded"] + 1),))
    f, e = mb.forces_energy(s)
    err = np.linalg.norm(f - w["forces_pme"], axis=1).max()
    print("water3", dt_.__name__, "max|dF| =", err, "dE =", e - float(w["energy_pme"]))
    assert err < tol_f and abs(e - float(w["energy_pme"])) < tol_e
"""


@pytest.mark.xfail(strict=False, reason="first implementation, never run on a GPU (round 1 budget was spent)")
def test_6mrr_all_pme_on_device():
    code = CHILD % dict(tests=os.path.join(ROOT, "tests"), root=ROOT, golden=os.path.join(ROOT, "tests", "golden", "6mrr.npz"),
                        water=os.path.join(ROOT, "tests", "golden", "water3.npz"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0
