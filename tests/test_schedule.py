"""CPU: the host-side schedule replay equals the reference loop structure (eri_transform.py:338-382) as restated by
the oracle, block counts match BASELINE.md section 3, and the rank assignment is a deterministic partition."""
import numpy as np
import pytest

from libdmet_preview_b200 import schedule as sch
from libdmet_preview_b200.synthetic import trs_block_count
from oracle import eri_transform as o_eri, fourier as o_f, pyscf_lib as olib


def oracle_schedule(kscaled, trs, tol=1e-6, center=None):
    """instrumented copy of the oracle's loop nest: returns [(kL, weight, [(i, j, sym)])]"""
    ks0 = np.array(kscaled, dtype=float)
    ks = ks0 - center if center is not None else ks0
    nk = len(ks)
    w = o_eri.get_weights_t_reversal(ks0) if trs else np.ones(nk, dtype=int)
    units = []
    for kL in range(nk):
        if w[kL] <= 0:
            continue
        vis = np.zeros(nk, dtype=bool)
        blocks = []
        for i in range(nk):
            if vis[i]:
                continue
            vis[i] = True
            for j in range(nk):
                kc = -ks[i] + ks[j] + ks[kL]
                if o_f.max_abs(np.round(kc) - kc) > tol:
                    continue
                if trs:
                    jm = o_f.kpt_member(-ks[j], ks)
                    assert len(jm) == 1
                    jm = jm[0]
                    blocks.append((i, j, int(not vis[jm])))
                    vis[jm] = True
                else:
                    blocks.append((i, j, 0))
        units.append((kL, int(w[kL]) if trs else 0, blocks))
    return units


@pytest.mark.parametrize("kmesh", [[1, 1, 1], [1, 1, 3], [3, 3, 1], [2, 2, 2], [2, 2, 1], [1, 4, 3], [4, 4, 4]])
@pytest.mark.parametrize("trs", [True, False])
def test_schedule_matches_reference_loop(kmesh, trs):
    ks = sch.make_kpts_scaled(kmesh)
    assert np.array_equal(ks, o_f.make_kpts_scaled(kmesh))
    s = sch.build_schedule(ks, trs)
    ref = oracle_schedule(ks, trs)
    assert [(u[0], u[1], [tuple(map(int, b)) for b in u[2]]) for u in s.units] == ref


def test_schedule_with_center_shift():
    kmesh = [2, 1, 3]
    ks = sch.make_kpts_scaled(kmesh)
    c = np.array([0.0, 0.0, 0.0])
    assert [(u[0], u[1], u[2]) for u in sch.build_schedule(ks, True, kscaled_center=c).units] == oracle_schedule(
        ks, True, center=c)


def test_block_counts_of_baseline_md():
    expect = {(1, 1, 3): 4, (3, 3, 1): 25, (2, 2, 2): 36, (2, 2, 1): 10, (2, 2, 4): 104, (3, 3, 3): 196,
              (4, 4, 2): 336, (4, 4, 4): 1184}
    for km, b in expect.items():
        assert trs_block_count(list(km))[0] == b
        assert trs_block_count(list(km), False)[0] == int(np.prod(km)) ** 2
    assert trs_block_count([4, 4, 4])[1] == 64


def test_weights_and_helpers():
    ks = sch.make_kpts_scaled([2, 2, 2])
    assert list(sch.time_reversal_weights(ks)) == list(o_eri.get_weights_t_reversal(ks)) == [1] * 8
    ks = sch.make_kpts_scaled([1, 1, 3])
    assert list(sch.time_reversal_weights(ks)) == [1, 2, 0]
    assert np.array_equal(sch.cell_vectors([2, 3, 1]), olib.cartesian_prod([np.arange(2), np.arange(3), np.arange(1)]))


def test_assign_units_partition():
    s = sch.build_schedule(sch.make_kpts_scaled([4, 4, 4]), True)
    costs = s.unit_cost(8.4e10, 1.28e11)
    for n in (1, 2, 4, 8, 40):
        parts = sch.assign_units(costs, n)
        flat = sorted(u for p in parts for u in p)
        assert flat == list(range(len(costs)))
        assert parts == sch.assign_units(costs, n)
        loads = [sum(costs[u] for u in p) for p in parts]
        if n <= 8:
            assert max(loads) / (sum(loads) / n) < 1.2      # LPT balance at the target mesh
