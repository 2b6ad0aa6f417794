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
    type_element: dict = field(default_factory=dict)
    type_sigma: dict = field(default_factory=dict)
    type_eps: dict = field(default_factory=dict)
    templates: dict = field(default_factory=dict)
    coulomb14scale: float = 1.0
    lj14scale: float = 1.0
    bonds: list = field(default_factory=list)  # (class1/type1, class2/type2, length, k)
    angles: list = field(default_factory=list)
    propers: list = field(default_factory=list)
    impropers: list = field(default_factory=list)
    torsion_ordering: str = "default"


def _tc(el, i):
    """(kind, name) of attribute type<i> / class<i>; name '' is the wildcard."""
    if el.get(f"type{i}") is not None:
        return ("type", el.get(f"type{i}"))
    return ("class", el.get(f"class{i}"))


def read_force_field(*paths) -> ForceField:
    ff = ForceField()
    for p in paths:
        root = ET.parse(p).getroot()
        for t in root.findall("./AtomTypes/Type"):
            ff.type_mass[t.get("name")] = float(t.get("mass"))
            ff.type_class[t.get("name")] = t.get("class")
            ff.type_element[t.get("name")] = t.get("element", "")
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
        for b in root.findall("./HarmonicBondForce/Bond"):
            ff.bonds.append((_tc(b, 1), _tc(b, 2), float(b.get("length")), float(b.get("k"))))
        for a in root.findall("./HarmonicAngleForce/Angle"):
            ff.angles.append((_tc(a, 1), _tc(a, 2), _tc(a, 3), float(a.get("angle")), float(a.get("k"))))
        ptf = root.find("./PeriodicTorsionForce")
        if ptf is not None:
            ff.torsion_ordering = ptf.get("ordering", "default")
            for kind, dest in (("Proper", ff.propers), ("Improper", ff.impropers)):
                for t in ptf.findall(kind):
                    terms = []
                    i = 1
                    while t.get(f"periodicity{i}") is not None:
                        terms.append((int(t.get(f"periodicity{i}")), float(t.get(f"phase{i}")), float(t.get(f"k{i}"))))
                        i += 1
                    dest.append(((_tc(t, 1), _tc(t, 2), _tc(t, 3), _tc(t, 4)), terms))
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
    # ---- bonded parameters (OpenMM ForceField semantics: the goldens were generated by OpenMM) -------------
    template_index = [0] * n
    residue_index = [0] * n
    for ri, (key, idxs) in enumerate(residues):
        order = {x[0]: k for k, x in enumerate(res_template[ri].atoms)}
        for i in idxs:
            template_index[i] = order[atoms[i].name]
            residue_index[i] = ri
    bonded = assign_bonded(ff, types, bonds, angles, torsions, adj, template_index, residue_index,
                           [ff.type_element.get(t, "") for t in types])
    mass = np.array([ff.type_mass[t] for t in types])
    sigma = np.array([ff.type_sigma[t] for t in types])
    eps = np.array([ff.type_eps[t] for t in types])
    return dict(types=types, mass=mass, charge=charges, sigma=sigma, eps=eps,
                bonds=np.array(bonds, np.int32).reshape(-1, 2),
                angles=np.array(angles, np.int32).reshape(-1, 3),
                torsions=np.array(torsions, np.int32).reshape(-1, 4),
                excluded=np.array(sorted(excl), np.int32).reshape(-1, 2),
                special=np.array(sorted(special), np.int32).reshape(-1, 2),
                n_residues=len(residues), **bonded)


def _m(spec, atype, ff):
    kind, name = spec
    if name == "":
        return True
    return atype == name if kind == "type" else ff.type_class.get(atype) == name


def _wild(specs):
    return any(name == "" for _, name in specs)


def assign_bonded(ff, types, bonds, angles, torsions, adj, template_index, residue_index, elements):
    """HarmonicBond / HarmonicAngle / PeriodicTorsion parameters per term, following OpenMM's matching rules
    (openmm/app/forcefield.py: first matching bond/angle; propers: first specific match, else first wildcard match;
    impropers: central atom first, any permutation of the others, last specific match else first wildcard match,
    'amber' ordering of the peripheral atoms)."""
    import itertools
    b_idx, b_par = [], []
    for i, j in bonds:
        for s1, s2, r0, k in ff.bonds:
            if (_m(s1, types[i], ff) and _m(s2, types[j], ff)) or (_m(s1, types[j], ff) and _m(s2, types[i], ff)):
                b_idx.append((i, j)); b_par.append((k, r0))
                break
        else:
            raise ValueError(f"no bond parameters for {types[i]}-{types[j]}")
    a_idx, a_par = [], []
    for i, j, k in angles:
        for s1, s2, s3, th0, kk in ff.angles:
            if _m(s2, types[j], ff) and ((_m(s1, types[i], ff) and _m(s3, types[k], ff)) or
                                         (_m(s1, types[k], ff) and _m(s3, types[i], ff))):
                a_idx.append((i, j, k)); a_par.append((kk, th0))
                break
        else:
            raise ValueError(f"no angle parameters for {types[i]}-{types[j]}-{types[k]}")
    t_idx, t_par = [], []
    seen = set()
    for i, j, k, l in torsions:
        key = (i, j, k, l) if i < l else (l, k, j, i)
        if key in seen:
            continue
        seen.add(key)
        match = None
        for specs, terms in ff.propers:
            fwd = all(_m(sp, types[a], ff) for sp, a in zip(specs, (i, j, k, l)))
            rev = all(_m(sp, types[a], ff) for sp, a in zip(specs, (l, k, j, i)))
            if fwd or rev:
                w = _wild(specs)
                if match is None or not w:
                    match = terms
                if not w:
                    break
        if match is None:
            continue
        for per, phase, kk in match:
            if kk != 0.0:
                t_idx.append(key); t_par.append((per, phase, kk))
    i_idx, i_par = [], []
    for c in range(len(types)):
        if len(adj[c]) != 3:
            continue
        tor = (c, adj[c][0], adj[c][1], adj[c][2])
        match = None
        for specs, terms in ff.impropers:
            w = _wild(specs)
            if match is not None and w:
                continue
            if not _m(specs[0], types[c], ff):
                continue
            for perm in itertools.permutations((1, 2, 3)):
                a2, a3, a4 = tor[perm[0]], tor[perm[1]], tor[perm[2]]
                if _m(specs[1], types[a2], ff) and _m(specs[2], types[a3], ff) and _m(specs[3], types[a4], ff):
                    if ff.torsion_ordering == "amber":
                        r2, r3, r4 = residue_index[a2], residue_index[a3], residue_index[a4]
                        ta2, ta3, ta4 = template_index[a2], template_index[a3], template_index[a4]
                        t2, t3, t4 = types[a2], types[a3], types[a4]
                        e2, e3, e4 = elements[a2], elements[a3], elements[a4]
                        if not w:
                            if t2 == t4 and (r2 > r4 or (r2 == r4 and ta2 > ta4)):
                                a2, a4, r2, r4, ta2, ta4 = a4, a2, r4, r2, ta4, ta2
                            if t3 == t4 and (r3 > r4 or (r3 == r4 and ta3 > ta4)):
                                a3, a4, r3, r4, ta3, ta4 = a4, a3, r4, r3, ta4, ta3
                            if t2 == t3 and (r2 > r3 or (r2 == r3 and ta2 > ta3)):
                                a2, a3 = a3, a2
                        else:
                            if e2 == e4 and (r2 > r4 or (r2 == r4 and ta2 > ta4)):
                                a2, a4, r2, r4, ta2, ta4 = a4, a2, r4, r2, ta4, ta2
                            if e3 == e4 and (r3 > r4 or (r3 == r4 and ta3 > ta4)):
                                a3, a4, r3, r4, ta3, ta4 = a4, a3, r4, r3, ta4, ta3
                            if r2 > r3 or (r2 == r3 and ta2 > ta3):
                                a2, a3 = a3, a2
                    match = ((a2, a3, c, a4), terms)
                    break
        if match is not None:
            for per, phase, kk in match[1]:
                if kk != 0.0:
                    i_idx.append(match[0]); i_par.append((per, phase, kk))
    f = lambda a, w: np.array(a, np.float64).reshape(-1, w)
    g = lambda a, w: np.array(a, np.int32).reshape(-1, w)
    return dict(bond_idx=g(b_idx, 2), bond_par=f(b_par, 2), angle_idx=g(a_idx, 3), angle_par=f(a_par, 2),
                proper_idx=g(t_idx, 4), proper_par=f(t_par, 3), improper_idx=g(i_idx, 4), improper_par=f(i_par, 3))
