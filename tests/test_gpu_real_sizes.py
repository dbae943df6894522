"""GPU: element-wise parity against the oracle at the REAL sizes of BASELINE.json's configs (bench.py WORKLOADS
c1 / c2 / c3 / sweep-min, a c4-shaped GSO build) and at the exact tile shape and k-point schedule of the target
workload (4x4x4, nao 200, neo 150: whole transfer-momentum units on a 64-row auxiliary slice, so the oracle
finishes in seconds while stage 1 runs the same 64x80 3M tiles, M/N tails and block chaining as the bench).

Thresholds are the reference's own: 1e-10 max-abs between two routes to the same integrals
(test_eri_transform_gdf.py:54,68,83,94; test_eri_transform_uhf.py:68; test_eri_transform_gso.py:155)."""
import numpy as np
import pytest

from helpers import problem, gso_basis

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _both(gdf, **kw):
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    return et.get_emb_eri(gdf.cell, gdf, **kw), oe.get_emb_eri(gdf.cell, gdf, **kw)


def test_c1_hchain_shape(dev):
    """configs[0]: 1x1x3, nao 4, naux 30, neo 6 -- s4, s1, s8, with and without time reversal"""
    gdf, C, basis = problem([1, 1, 3], 4, 30, 6, seed=11)
    for kw in (dict(), dict(symmetry=1), dict(symmetry=8), dict(t_reversal_symm=False)):
        got, ref = _both(gdf, C_ao_lo=C, basis=basis, **kw)
        assert got.shape == ref.shape and np.abs(got - ref).max() < TOL, kw


def test_c2_graphene_shape(dev):
    """configs[1]: 3x3x1, nao 26, naux 150, neo 40 (25 blocks with time reversal, 81 without)"""
    gdf, C, basis = problem([3, 3, 1], 26, 150, 40, seed=12)
    got, ref = _both(gdf, C_ao_lo=C, basis=basis)
    assert got.shape == (1, 820, 820) and np.abs(got - ref).max() < TOL
    got2, ref2 = _both(gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False)
    assert np.abs(got2 - ref2).max() < TOL
    assert np.abs(got2 - got).max() < TOL               # time reversal vs plain (test_eri_transform_gdf.py:94)


def test_c3_nio_unrestricted_shape(dev):
    """configs[2]: 2x2x2, nao 78, naux 400, neo 106, two spins -> aa, ab, bb (test_eri_transform_uhf.py:36-40, 68)"""
    gdf, C, basis = problem([2, 2, 2], 78, 400, 106, spin=2, seed=13)
    got, ref = _both(gdf, C_ao_lo=C, basis=basis)
    npair = 106 * 107 // 2
    assert got.shape == ref.shape == (3, npair, npair)
    err = [float(np.abs(got[s] - ref[s]).max()) for s in range(3)]
    assert max(err) < TOL, err
    assert np.array_equal(got[0], got[0].T) and np.array_equal(got[2], got[2].T)
    assert not np.allclose(got[1], got[1].T)             # ab is a full product, not symmetric


def test_c4_shaped_gso_large_neo(dev):
    """configs[3] shaped: 2x2x1, generalised spin orbitals, neo large against nao (test_eri_transform_gso.py)"""
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    kmesh, nao, naux, neo = [2, 2, 1], 60, 300, 120
    gdf, C, _ = problem(kmesh, nao, naux, 8, spin=2, seed=14)
    basis = gso_basis(kmesh, nao, neo, seed=14)
    got = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    ref = oe.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    assert got.shape == ref.shape == (1, neo * (neo + 1) // 2, neo * (neo + 1) // 2)
    assert np.abs(got - ref).max() < TOL
    got2 = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False)
    assert np.abs(got2 - ref).max() < TOL                # real R-space basis: both schedules agree (l.155)


def test_sweep_minimum_shape(dev):
    """configs[4] lower corner: 2x2x2, nao 100, naux 500, neo 50"""
    gdf, C, basis = problem([2, 2, 2], 100, 500, 50, seed=15)
    got, ref = _both(gdf, C_ao_lo=C, basis=basis)
    assert got.shape == (1, 1275, 1275) and np.abs(got - ref).max() < TOL
    from libdmet_preview_b200 import eri_transform as et
    got_h = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, source="host")
    assert np.array_equal(got_h, got)                   # blocks staged from host memory: bit-identical


@pytest.mark.parametrize("unit_pick", ["weight1", "weight2"])
def test_target_schedule_units_on_aux_slice(dev, unit_pick):
    """Whole transfer-momentum units of the 4x4x4 target schedule (nao 200, neo 150) on a 64-row auxiliary slice:
    the exact tile configuration, N tail (150 in 2 x 80), block chaining and symmetrise flags of the bench, checked
    element-wise.  The oracle runs the same unit through its `kL_subset` hook (the reference's MPI variant shards
    the kL loop the same way, eri_transform_mpi.py:151-157)."""
    from libdmet_preview_b200 import eri_transform as et
    from libdmet_preview_b200.schedule import build_schedule
    from oracle import eri_transform as oe
    kmesh, nao, naux, neo = [4, 4, 4], 200, 64, 150
    gdf, C, basis = problem(kmesh, nao, naux, neo, seed=16)
    sch = build_schedule(gdf.kpts_scaled, True)
    assert sch.nblocks == 1184 and sch.ngram == 64
    want = 1 if unit_pick == "weight1" else 2
    # the weight-1 unit with the most symmetrised blocks / the first weight-2 unit
    cands = [u for u in range(len(sch.units)) if sch.units[u][1] == want]
    u = cands[1] if len(cands) > 1 else cands[0]
    kL, w, blocks = sch.units[u]
    assert len(blocks) >= 32
    CT = et.build_CT(gdf, C, basis)
    eri = et.emb_eri_device(gdf, CT, schedule=sch, items=[(u, 0, naux)])
    got = dev.to_host(et.finalize_eri(eri, neo, 4, 1))
    ref = oe.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_lo=C, basis=basis, kL_subset={kL}, restore=False)
    assert got.shape == ref.shape == (1, 11325, 11325)
    assert np.abs(ref).max() > 1e-3
    assert np.abs(got - ref).max() < TOL
    # same unit with the blocks staged from host memory (the e2e route of bench.py): bit-identical
    eri_h = et.emb_eri_device(gdf, CT, schedule=sch, items=[(u, 0, naux)], source="host")
    assert np.array_equal(dev.to_host(et.finalize_eri(eri_h, neo, 4, 1)), got)


def test_target_shape_full_aux_one_block_pair(dev):
    """nao 200, naux 1000, neo 150 at full auxiliary length on the smallest mesh that has a weight-2 unit (1x1x3):
    4 blocks of the bench's exact (naux*nao) x neo x nao GEMM shape against the oracle."""
    gdf, C, basis = problem([1, 1, 3], 200, 1000, 150, seed=17)
    got, ref = _both(gdf, C_ao_lo=C, basis=basis)
    assert got.shape == (1, 11325, 11325)
    assert np.abs(got - ref).max() < TOL
