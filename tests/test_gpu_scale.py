"""GPU parity at production scale: the code paths that only a large mesh (or a small grid)
reaches — every persistent CTA of S_A / F_A streams many tiles through its shared-memory ring, so
the stage index wraps, the phase bit flips, the `empty` barriers exert back-pressure and stages
are reused while bulk copies are in flight (axb_solid_tile.cuh / axb_fluid_tile.cuh).

Two ways to get there:
  * a mesh with > 40 000 solid elements: >= 11 tiles per CTA at the full grid of 3 CTAs per SM
    (also with a two-stage ring: >= 5 wraps per CTA);
  * a small mesh run on a deliberately small grid (AXB_SOLID_GRID / AXB_FLUID_GRID) with a
    two-stage ring: 30+ tiles per CTA, cheap enough to sweep every configuration.

Bar as in test_gpu_parity.py: the -fmad=false build bit-identical to the oracle, the product
build (FMA contraction, lean Newmark formulation, CUDA-graph replay) within 1e-5 relative L2.
"""
import numpy as np
import pytest

from axisem_b200.host import AttenuationModel, SourceParams, build_problem, prem_mesh_spec
from tests.util import apply_state, make_problem, rel_l2, seeded_state

pytestmark = pytest.mark.gpu


def _cmp(name, g, o, strict, tol=1e-5):
    if strict:
        assert np.array_equal(g, o), f"{name}: strict build not bit-identical (rel l2 {rel_l2(g, o):.3e})"
    else:
        assert rel_l2(g, o) <= tol, f"{name}: rel l2 {rel_l2(g, o):.3e} > {tol}"


def _run_pair(prob, strict, n, fields, seed_fields=("disp", "velo", "acc0", "chi", "dchi", "ddchi0")):
    from axisem_b200 import solver
    from oracle import oracle
    G = solver.time_loop(prob, strict=strict)
    O = oracle.make_loop(prob)
    st = seeded_state(G, scale=1e-9, fields=seed_fields)
    for L in (G, O):
        apply_state(L, st)
        L.run(n)
    comps = [0, 2] if prob.src_order == 0 else [0, 1, 2]
    for f in fields:
        g, o = G.get(f), O.get(f)
        if f in ("disp", "velo", "acc0"):
            g, o = g[comps], o[comps]
        assert np.abs(o).max() > 0, f
        _cmp(f, g, o, strict)
    _cmp("seismograms", G.seismograms(), O.seismograms(), strict)
    assert G.gpu_launches > 0
    return G, O


# ---- small grid, shallow ring: every configuration ------------------------------------------
CASES = [("explosion", False, True, "newmark2"), ("explosion", True, True, "newmark2"),
         ("mtr", False, True, "newmark2"), ("mtr", True, True, "newmark2"),
         ("mtp", False, True, "newmark2"), ("mtp", True, True, "newmark2"),
         ("mtp", True, False, "newmark2"), ("mtr", True, True, "symplec4"),
         ("explosion", False, True, "symplec4")]


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("src,anel,cg,scheme", CASES)
def test_ring_wrap_on_a_small_grid(src, anel, cg, scheme, strict, monkeypatch):
    """40 x 36 mesh (~1100 solid elements = 140 tiles, ~20 fluid tiles) on 4 solid / 2 fluid CTAs
    with two-stage rings: 35 tiles per solid CTA, 17 ring wraps each."""
    monkeypatch.setenv("AXB_SOLID_GRID", "4")
    monkeypatch.setenv("AXB_SOLID_STAGES", "2")
    monkeypatch.setenv("AXB_FLUID_GRID", "2")
    monkeypatch.setenv("AXB_FLUID_STAGES", "2")
    n = 30 if scheme == "newmark2" else 8
    prob = make_problem(src, anel=anel, ntheta=40, nr=36, niter=n, scheme=scheme, coarse_grained=cg,
                        anisotropic=True)
    assert prob.mesh.nel_solid >= 1000 and prob.mesh.nel_fluid >= 250
    fields = ["disp", "velo", "chi", "dchi"] + (["memvar", "src_dev_tm1", "src_tr_tm1"] if anel else [])
    seeds = ("disp", "velo", "acc0", "chi", "dchi", "ddchi0") if scheme == "newmark2" else ("disp", "velo", "chi", "dchi")
    _run_pair(prob, strict, n, fields, seeds)


# ---- production-size meshes at the full grid --------------------------------------------------
BIG = [("mtr", True, True, "newmark2", 25), ("explosion", False, True, "newmark2", 25),
       ("mtp", True, False, "newmark2", 12), ("mtr", True, True, "symplec4", 5)]


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("stages", [0, 2])
@pytest.mark.parametrize("src,anel,cg,scheme,n", BIG)
def test_production_scale_mesh(src, anel, cg, scheme, n, stages, monkeypatch):
    """256 x 224 mesh: > 40 000 solid elements (>= 5000 solid tiles on 444 CTAs, >= 11 per CTA),
    > 10 000 fluid elements; dipole cg4, monopole elastic, quadrupole with memory variables at
    all 25 points, Newmark and symplec4.  Strict build bit-identical to the oracle in every state
    array and the seismograms, product build within 1e-5; both also with a two-stage ring."""
    if stages:
        if src != "mtr":
            pytest.skip("two-stage ring: dipole cases only (the oracle run dominates the cost)")
        monkeypatch.setenv("AXB_SOLID_STAGES", str(stages))
        monkeypatch.setenv("AXB_FLUID_STAGES", str(stages))
    spec = prem_mesh_spec(ntheta=256, nr_target=224)
    att = AttenuationModel(coarse_grained=cg) if anel else None
    prob = build_problem(spec, SourceParams(src_type2=src, t_0=40.0), anel=anel, att=att, niter=n,
                         time_scheme=scheme)
    assert prob.mesh.nel_solid > 40000 and prob.mesh.nel_fluid > 10000, (prob.mesh.nel_solid, prob.mesh.nel_fluid)
    from axisem_b200 import solver
    from oracle import oracle
    O = oracle.make_loop(prob)
    seeds = ("disp", "velo", "acc0", "chi", "dchi", "ddchi0") if scheme == "newmark2" else ("disp", "velo", "chi", "dchi")
    st = seeded_state(O, scale=1e-9, fields=seeds)
    apply_state(O, st)
    O.run(n)
    fields = ["disp", "velo", "chi", "dchi"] + (["memvar", "src_dev_tm1", "src_tr_tm1"] if anel else [])
    comps = [0, 2] if prob.src_order == 0 else [0, 1, 2]
    ref = {f: O.get(f) for f in fields}
    ref_seis = O.seismograms()
    del O
    for strict in (True, False):
        G = solver.time_loop(prob, strict=strict)
        apply_state(G, st)
        G.run(n)
        for f in fields:
            g, o = G.get(f), ref[f]
            if f in ("disp", "velo"):
                g, o = g[comps], o[comps]
            _cmp(f, g, o, strict)
        _cmp("seismograms", G.seismograms(), ref_seis, strict)
        G.close()


# ---- the formulation / launch variants of the product build -------------------------------------
@pytest.mark.parametrize("lean,graph", [("0", "0"), ("0", "1"), ("1", "0"), ("1", "1")])
@pytest.mark.parametrize("src", ["explosion", "mtr"])
def test_product_variants_agree_with_the_oracle(src, lean, graph, monkeypatch):
    """The product library in its four combinations of {reference statement order, lean Newmark}
    x {direct launches, CUDA-graph replay}: each within 1e-5 of the oracle after 120 steps, the
    state handed back in the reference's variables (velo, acc0, dchi, ddchi0) in the middle of
    the run and at its end."""
    monkeypatch.setenv("AXB_LEAN", lean)
    monkeypatch.setenv("AXB_GRAPH", graph)
    from axisem_b200 import solver
    from oracle import oracle
    n = 120
    prob = make_problem(src, anel=True, niter=n, dump=True, strain_it=30, seis_it=2)
    G, O = solver.time_loop(prob), oracle.make_loop(prob)
    st = seeded_state(G, scale=1e-9)
    for L in (G, O):
        apply_state(L, st)
        L.run(50)
    comps = [0, 2] if src == "explosion" else [0, 1, 2]
    for f in ("velo", "acc0", "dchi", "ddchi0"):
        g, o = G.get(f), O.get(f)
        if f in ("velo", "acc0"):
            g, o = g[comps], o[comps]
        _cmp(f + " (mid-run)", g, o, False)
    for L in (G, O):
        L.run(1)
        L.run(n - 51)
    assert G.iter == O.iter == n and G.nseismo == O.nseismo and G.nstrain == O.nstrain
    for f in ("disp", "velo", "acc0", "chi", "dchi", "ddchi0", "memvar"):
        g, o = G.get(f), O.get(f)
        if f in ("disp", "velo", "acc0"):
            g, o = g[comps], o[comps]
        _cmp(f, g, o, False)
    _cmp("seismograms", G.seismograms(), O.seismograms(), False)
    _cmp("snapshots", G.snapshots(), O.snapshots(), False)


@pytest.mark.parametrize("graph", ["0", "1"])
def test_strict_build_is_bit_identical_with_and_without_graph_replay(graph, monkeypatch):
    monkeypatch.setenv("AXB_GRAPH", graph)
    prob = make_problem("mtr", anel=True, niter=40, seis_it=3)
    _run_pair(prob, True, 40, ["disp", "velo", "acc0", "chi", "dchi", "ddchi0", "memvar"])


def test_source_time_function_fed_step_by_step(monkeypatch):
    """The e2e pattern of bench.py: per step axb_set_stf_values (values travel as kernel
    arguments), axb_run(1) from the step graph, axb_fetch_seismograms — bit-identical (strict) to
    the oracle that had the whole table from the start."""
    from axisem_b200 import solver
    from oracle import oracle
    n = 60
    prob = make_problem("mtr", anel=True, niter=n, t_0=5.0)
    assert np.abs(prob.stf[:n]).max() > 0
    O = oracle.make_loop(prob)
    O.run(n)
    stf = prob.stf.copy()
    prob.stf = np.zeros_like(stf)
    G = solver.time_loop(prob, strict=True)
    buf = np.zeros(1, np.float32)
    for k in range(n):
        buf[0] = stf[k]
        G.set_stf_values(k, buf)
        buf[0] = np.nan                      # the call must not have kept a reference to the buffer
        G.run(1, sync=False)
        G.seismograms(G.nseismo - 1, 1)
    assert np.array_equal(G.seismograms(), O.seismograms())
    assert np.array_equal(G.get("disp"), O.get("disp"))


def test_halo_wait_is_bounded(monkeypatch):
    """A neighbour that never delivers: the wait inside the corrector gives up after
    AXB_HALO_TIMEOUT_MS, the abort flag lets every later wait through, and axb_synchronize
    reports it (the reference's pcheck stops the run, commpi.F90:64-111)."""
    from axisem_b200 import solver
    from axisem_b200.capi import AxbError
    monkeypatch.setenv("AXB_HALO_TIMEOUT_MS", "200")
    probs = [make_problem("mtr", niter=20, rank=r, nranks=2) for r in range(2)]
    lib, loops = solver.time_loop_group(probs)
    with pytest.raises(AxbError, match="HALO EXCHANGE TIMED OUT"):
        loops[0].run(3)                      # rank 1 is never stepped
