"""Generate tests/golden/*.npz from the reference's own data files (build container only).

Reads /root/reference/data (never available on the GPU box) and writes small
fixtures: the 6mrr system (per-atom parameters, exclusions, 1-4 specials derived
by oracle/ffreader.py) and the OpenMM golden forces/energies the reference's
test/protein.jl:206-276 compares against, plus per-pair literals.
Usage: python oracle/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ffreader as fr  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    ff = fr.read_force_field(f"{REF}/force_fields/ff99SBildn.xml", f"{REF}/force_fields/tip3p_standard.xml")
    atoms, box = fr.read_pdb(f"{REF}/6mrr_equil.pdb")
    top = fr.build_topology(atoms, ff)
    coords = np.array([a.xyz for a in atoms], np.float64)
    amber = f"{REF}/openmm_6mrr/amber"
    out = dict(
        box=box, coords=coords, mass=top["mass"], charge=top["charge"], sigma=top["sigma"], eps=top["eps"],
        excluded=top["excluded"], special=top["special"], bonds=top["bonds"], angles=top["angles"],
        torsions=top["torsions"],
        lj14scale=np.float64(ff.lj14scale), coulomb14scale=np.float64(ff.coulomb14scale),
        velocities_300K=np.loadtxt(f"{REF}/openmm_6mrr/velocities_300K.txt"),
    )
    for key in ("bond_idx", "bond_par", "angle_idx", "angle_par", "proper_idx", "proper_par", "improper_idx", "improper_par"):
        out[key] = top[key]
    for name in ("lj_only", "coul_only", "bond_only", "angle_only", "proptor_only", "improptor_only", "all_cut",
                 "all_pme_exact", "all_pme"):
        out[f"forces_{name}"] = np.loadtxt(f"{amber}/forces_{name}.txt")
        out[f"energy_{name}"] = np.float64(open(f"{amber}/energy_{name}.txt").read())
    np.savez_compressed(os.path.join(OUT, "6mrr.npz"), **out)
    print("wrote", os.path.join(OUT, "6mrr.npz"), os.path.getsize(os.path.join(OUT, "6mrr.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
