"""Minimal PDB + OpenMM-XML force-field reader for the oracle (TEST INFRASTRUCTURE).

Restates the small part of the reference's setup layer that the non-bonded
parity tests need: per-atom (mass, charge, sigma, epsilon), the
bond graph, and from it the `eligible` / `special` pair sets

    src/setup.jl:713-855   self, 1-2 (bonds) and 1-3 (angle ends) excluded;
                           proper-torsion ends i-l special (unless excluded)
    src/setup.jl:1851-1892 LJ weight_special = lj14scale, Coulomb weight = coulomb14scale

It matches residues to templates by residue name + atom-name set (the reference
matches by graph isomorphism, src/setup.jl:616-690; for standard PDB names the
two agree) and is validated end-to-end against the OpenMM golden force files in
tests/test_oracle_golden.py. It runs only in the build container (it reads
/root/reference/data); its outputs are committed under tests/golden/.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Template:
    name: str
    atoms: list  # [(name, type, charge)]
    bonds: list  # [(name1, name2)]
    external: list  # [name]


@dataclass
class ForceField:
    type_mass: dict = field(default_factory=dict)
    type_class: dict = field(default_factory=dict)
    type_sigma: dict = field(default_factory=dict)
    type_eps: dict = field(default_factory=dict)
    templates: dict = field(default_factory=dict)
    coulomb14scale: float = 1.0
    lj14scale: float = 1.0
    bonds: list = field(default_factory=list)  # (class1/type1, class2/type2, length, k)
    angles: list = field(default_factory=list)
    propers: list = field(default_factory=list)
    impropers: list = field(default_factory=list)


def read_force_field(*paths) -> ForceField:
    ff = ForceField()
    for p in paths:
        root = ET.parse(p).getroot()
        for t in root.findall("./AtomTypes/Type"):
            ff.type_mass[t.get("name")] = float(t.get("mass"))
            ff.type_class[t.get("name")] = t.get("class")
        for r in root.findall("./Residues/Residue"):
            atoms = [(a.get("name"), a.get("type"), float(a.get("charge", "0"))) for a in r.findall("Atom")]
            bonds = []
            for b in r.findall("Bond"):
                if b.get("atomName1") is not None:
                    bonds.append((b.get("atomName1"), b.get("atomName2")))
                else:  # older from/to index form
                    bonds.append((atoms[int(b.get("from"))][0], atoms[int(b.get("to"))][0]))
            ext = [e.get("atomName") for e in r.findall("ExternalBond")]
            ff.templates[r.get("name")] = Template(r.get("name"), atoms, bonds, ext)
        nb = root.find("./NonbondedForce")
        if nb is not None:
            ff.coulomb14scale = float(nb.get("coulomb14scale"))
            ff.lj14scale = float(nb.get("lj14scale"))
            for a in nb.findall("Atom"):
                key = a.get("type") or a.get("class")
                ff.type_sigma[key] = float(a.get("sigma"))
                ff.type_eps[key] = float(a.get("epsilon"))
    return ff


@dataclass
class PDBAtom:
    name: str
    resname: str
    chain: str
    resnum: int
    xyz: tuple
    hetero: bool


def read_pdb(path):
    atoms = []
    box = None
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec == "CRYST1":
                box = np.array([float(line[6:15]), float(line[15:24]), float(line[24:33])]) / 10.0
            elif rec in ("ATOM  ", "HETATM"):
                atoms.append(PDBAtom(line[12:16].strip(), line[17:20].strip(), line[21], int(line[22:26]),
                                     (float(line[30:38]) / 10.0, float(line[38:46]) / 10.0,
                                      float(line[46:54]) / 10.0), rec == "HETATM"))
    return atoms, box


def build_topology(atoms, ff: ForceField):
    """Returns dict with per-atom params, bonds, excluded pairs, special pairs (0-based, i<j)."""
    # group residues
    residues = []
    cur = None
    for idx, a in enumerate(atoms):
        key = (a.chain, a.resnum, a.resname)
        if cur is None or cur[0] != key:
            cur = (key, [])
            residues.append(cur)
        cur[1].append(idx)
    n = len(atoms)
    types = [None] * n
    charges = np.zeros(n)
    bonds = []
    # protein residues are those with a template that has external bonds
    def is_polymer(resname):
        for cand in (resname, "N" + resname, "C" + resname, "HID", "HIE", "HIP"):
            t = ff.templates.get(cand)
            if t is not None and t.external:
                return True
        return False

    res_template = []
    for ri, (key, idxs) in enumerate(residues):
        names = [atoms[i].name for i in idxs]
        nameset = frozenset(names)
        resname = key[2]
        cands = [resname, "N" + resname, "C" + resname]
        if resname == "HIS":
            cands = ["HID", "HIE", "HIP", "NHID", "NHIE", "NHIP", "CHID", "CHIE", "CHIP"]
        chosen = None
        for c in cands:
            t = ff.templates.get(c)
            if t is not None and frozenset(x[0] for x in t.atoms) == nameset:
                chosen = t
                break
        if chosen is None:
            for t in ff.templates.values():
                if frozenset(x[0] for x in t.atoms) == nameset and len(t.atoms) == len(names):
                    chosen = t
                    break
        if chosen is None:
            raise ValueError(f"could not match residue {key} with atoms {sorted(names)}")
        res_template.append(chosen)
        tmap = {x[0]: x for x in chosen.atoms}
        name_to_idx = {atoms[i].name: i for i in idxs}
        for i in idxs:
            _, ty, ch = tmap[atoms[i].name]
            types[i] = ty
            charges[i] = ch
        for a1, a2 in chosen.bonds:
            bonds.append((name_to_idx[a1], name_to_idx[a2]))
    # peptide bonds between consecutive residues of a chain: C(i) - N(i+1)
    for ri in range(len(residues) - 1):
        (k1, i1), (k2, i2) = residues[ri], residues[ri + 1]
        t1, t2 = res_template[ri], res_template[ri + 1]
        if k1[0] != k2[0]:
            continue
        if "C" in t1.external and "N" in t2.external:
            c = [i for i in i1 if atoms[i].name == "C"]
            nn = [i for i in i2 if atoms[i].name == "N"]
            if c and nn:
                bonds.append((c[0], nn[0]))
    bonds = sorted({(min(a, b), max(a, b)) for a, b in bonds})
    adj = [[] for _ in range(n)]
    for a, b in bonds:
        adj[a].append(b)
        adj[b].append(a)
    excl = set(bonds)
    angles = []
    for j in range(n):
        nb = adj[j]
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                i, k = nb[x], nb[y]
                angles.append((i, j, k))
                excl.add((min(i, k), max(i, k)))
    torsions = []
    special = set()
    for j, k in bonds:
        for i in adj[j]:
            if i == k:
                continue
            for l in adj[k]:
                if l == j or l == i:
                    continue
                torsions.append((i, j, k, l))
                special.add((min(i, l), max(i, l)))
    # a pair that is excluded is skipped entirely; keep special minus excluded
    special -= excl
    mass = np.array([ff.type_mass[t] for t in types])
    sigma = np.array([ff.type_sigma[t] for t in types])
    eps = np.array([ff.type_eps[t] for t in types])
    return dict(types=types, mass=mass, charge=charges, sigma=sigma, eps=eps,
                bonds=np.array(bonds, np.int32).reshape(-1, 2),
                angles=np.array(angles, np.int32).reshape(-1, 3),
                torsions=np.array(torsions, np.int32).reshape(-1, 4),
                excluded=np.array(sorted(excl), np.int32).reshape(-1, 2),
                special=np.array(sorted(special), np.int32).reshape(-1, 2),
                n_residues=len(residues))
