"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/mollyb200.h declares; without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mollyb200 as mb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mollyb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    L = mb.capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in mollyb200.h but not exported"
    assert sorted(mb.capi.EXPORTED) == declared


def test_struct_layouts():
    assert C.sizeof(mb.capi.MBInter) == 72
    assert mb.atom_dtype(np.float32).itemsize == 32  # src/types.jl:466 "fits into 32 bytes"
    assert mb.atom_dtype(np.float64).itemsize == 56
    from oracle import oracle as o
    assert C.sizeof(o.InterC) == C.sizeof(mb.capi.MBInter)
    assert [f[0] for f in o.InterC._fields_] == [f[0] for f in mb.capi.MBInter._fields_]


def test_no_silent_cpu_fallback():
    if mb.device_count() > 0:
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    L = mb.capi.load()
    rc = L.mb_ctx_create(0, 32, None, C.byref(ctx))
    assert rc == mb.capi.MB_ERR_NOGPU
    assert b"no CPU fallback" in L.mb_last_error()
    sysd = dict(mass=[1.0, 1.0], charge=[0, 0], sigma=[0.3, 0.3], eps=[0.2, 0.2])
    s = mb.System(atoms=mb.atoms_from_arrays(**sysd, dtype=np.float64), coords=np.zeros((2, 3)),
                  boundary=mb.CubicBoundary(2.0), pairwise_inters=(mb.LennardJones(),), dtype=np.float64)
    with pytest.raises(mb.MollyB200Error):
        mb.forces(s)


def test_invalid_arguments_rejected_before_any_compute():
    L = mb.capi.load()
    ctx = C.c_void_p()
    assert L.mb_ctx_create(0, 16, None, C.byref(ctx)) == mb.capi.MB_ERR_INVALID
    assert L.mb_set_box(None, (C.c_double * 3)(1, 1, 1)) == mb.capi.MB_ERR_INVALID


def test_interaction_descriptors():
    d = mb.LennardJones(cutoff=mb.DistanceCutoff(1.2), use_neighbors=True, weight_special=0.5).descriptor()
    assert (d.kind, d.cutoff_kind, d.r_cut, d.weight_special, d.use_neighbors) == (0, 1, 1.2, 0.5, 1)
    d = mb.CoulombReactionField(dist_cutoff=1.0, weight_special=0.8333, use_neighbors=True).descriptor()
    assert (d.kind, d.r_cut, d.solvent_dielectric, d.coulomb_const) == (2, 1.0, 78.3, 138.93545764)
    nf = mb.GPUNeighborFinder(dist_cutoff=1.0, eligible=~np.eye(3, dtype=bool) & ~np.array(
        [[0, 1, 0], [1, 0, 0], [0, 0, 0]], bool), special=np.array([[0, 0, 1], [0, 0, 0], [1, 0, 0]], bool))
    from molly_jl_b200.api import _pairs_from
    assert _pairs_from(nf.eligible, 3, want_true=False).tolist() == [[1, 2]]
    assert _pairs_from(nf.special, 3, want_true=True).tolist() == [[1, 3]]


def test_pme_plan_matches_oracle():
    """Host half of the PME row: alpha, mesh dimensions and B-spline moduli computed by the library
    (mb_pme_plan, no GPU) against the numpy restatement that is pinned on OpenMM's goldens (oracle/pme.py)."""
    import ctypes as C
    import numpy as np
    from oracle import pme
    L = mb.capi.load()
    for box, rc, tol in (((5.676, 5.6627, 6.2963), 1.0, 0.0005), ((3.0, 4.1, 2.2), 0.9, 1e-4), ((1.0, 1.0, 1.0), 0.45, 0.01)):
        b = (C.c_double * 3)(*box)
        alpha = C.c_double()
        mesh = (C.c_int32 * 3)()
        mod = np.zeros(4096)
        assert L.mb_pme_plan(b, rc, tol, 5, C.byref(alpha), mesh, mod.ctypes.data, len(mod)) == 0
        a_ref = pme.pme_alpha(rc, tol)
        k_ref = pme.pme_mesh_dims(box, a_ref, tol)
        assert abs(alpha.value - a_ref) < 1e-15 * a_ref * 4 and tuple(mesh) == k_ref
        m_ref = np.concatenate(pme.bspline_moduli(5, k_ref))
        assert np.allclose(mod[:len(m_ref)], m_ref, rtol=1e-12, atol=1e-15)
    alpha = C.c_double()
    mesh = (C.c_int32 * 3)()
    assert L.mb_pme_plan(None, 1.0, 0.0005, 5, C.byref(alpha), mesh, None, 0) != 0  # bad arguments -> status, no crash
