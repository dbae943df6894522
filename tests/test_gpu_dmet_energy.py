"""GPU: converged HF-in-DMET energy (north_star: <= 1e-8 Ha on the converged DMET energy).

A self-consistent k-point Hartree-Fock on a seeded synthetic GDF tensor (tests/hf_in_dmet.py) plays the role of the
PySCF KRHF object of libdmet/test/test_mfd.py.  The CUDA path then runs the whole loop body -- `Lattice.set_Ham`
(AO -> LO transforms, seven k2R transforms), lattice HF (`oracle/mfd.py` <- libdmet/routine/mfd.py, host LAPACK as in
the reference), `get_emb_basis`, `embHam` (embedding ERI + one-body part + J/K), impurity HF to convergence,
`get_H_dmet` -- and must satisfy what the reference test asserts, with its thresholds:
    lattice HF density matrix = mean-field density matrix             < 1e-8     (test_mfd.py:113)
    folded density matrix is a fixed point of the impurity HF         < 1e-8     (test_mfd.py:138)
    fragment energy = k-point HF energy per cell                      < 1e-8 Ha  (test_mfd.py:153)
and agree with the same loop run on the oracle: |E_gpu - E_oracle| < 1e-8 Ha, restricted and unrestricted."""
import numpy as np
import pytest

from helpers import OracleLattice
import hf_in_dmet as hd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kmesh,nao,naux,nocc,sym", [
    ([1, 1, 3], 4, 30, 2, 4),          # configs[0] shape (H chain 1x1x3), restricted, s4
    ([1, 1, 3], 4, 30, 2, 1),          # same, s1 integrals (the reference test uses eri_symmetry = 1)
    ([2, 2, 1], 6, 20, 3, 4),          # 2-D mesh
    ([1, 1, 3], 5, 24, (3, 2), 4),     # unrestricted, spin-polarised
    ([2, 1, 2], 6, 20, (3, 2), 1),     # unrestricted, s1
])
def test_converged_dmet_energy(dev, kmesh, nao, naux, nocc, sym):
    from libdmet_preview_b200 import lattice as lat, synthetic
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=61)
    spin = 2 if np.ndim(nocc) else 1
    restricted = spin == 1
    hcore = hd.gapped_hcore(kmesh, nao, nocc, seed=3)
    mf = hd.lattice_scf(gdf, hcore, nocc)
    C = synthetic.make_C_ao_lo(kmesh, nao, seed=62, spin=(2 if spin == 2 else None))
    ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * len(gdf.kpts_scaled))
    H0 = 0.125
    E_ref = mf["e_cell"] + H0

    L = lat.Lattice(gdf.cell, kmesh)
    L.set_val_virt_core(nao, 0, 0)                     # impurity = the whole cell, full bath (see test_hf_in_dmet.py)
    L.set_Ham(None, gdf, C, eri_symmetry=sym, ovlp=ovlp, hcore=hcore, rdm1=mf["rdm1"], vhf=mf["vhf"], H0=H0)
    O = OracleLattice(gdf, C, hcore, ovlp, mf["rdm1"], mf["vhf"], eri_symmetry=sym, H0=H0)
    O.val_idx, O.virt_idx = list(range(nao)), []

    filling = nocc / float(nao) if restricted else [nocc[0] / float(nao), nocc[1] / float(nao)]
    nelec_emb = nao if restricted else [nao, nao]
    gpu = hd.dmet_cycle(L, hd.ProductMods(), filling, restricted, nelec_emb)
    ora = hd.dmet_cycle(O, hd.OracleMods, filling, restricted, nelec_emb)

    assert gpu["basis"].shape[-1] == 2 * nao
    assert abs(gpu["E_lattice_HF"] - E_ref) < 1e-10                     # test_mfd.py:111-112
    assert gpu["rdm_diff"] < 1e-8                                       # :113
    assert gpu["fixed_point_diff"] < 1e-8                               # :138
    assert abs(gpu["E_frag"] - E_ref) < 1e-8, (gpu["E_frag"], E_ref)    # :153
    # CUDA path against the oracle, converged numbers
    assert abs(gpu["E_frag"] - ora["E_frag"]) < 1e-8
    assert abs(gpu["E_imp"] - ora["E_imp"]) < 1e-8
    assert np.abs(np.asarray(gpu["ImpHam"].H2["ccdd"]) - np.asarray(ora["ImpHam"].H2["ccdd"])).max() < 1e-10
    assert np.abs(np.asarray(gpu["ImpHam"].H1["cd"]) - np.asarray(ora["ImpHam"].H1["cd"])).max() < 1e-10


def test_two_dmet_iterations_with_update_Ham(dev):
    """second iteration after the charge-self-consistency step `Lattice.update_Ham` (test_mfd.py:115): the density
    matrix of the first lattice HF replaces the mean-field one and the loop body is repeated -- on a converged mean
    field the energy must not move"""
    from libdmet_preview_b200 import lattice as lat, synthetic
    kmesh, nao, naux, nocc = [1, 1, 3], 4, 30, 2
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=61)
    hcore = hd.gapped_hcore(kmesh, nao, nocc, seed=3)
    mf = hd.lattice_scf(gdf, hcore, nocc)
    C = synthetic.make_C_ao_lo(kmesh, nao, seed=62)
    ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * len(gdf.kpts_scaled))
    L = lat.Lattice(gdf.cell, kmesh)
    L.set_val_virt_core(nao, 0, 0)
    L.set_Ham(None, gdf, C, eri_symmetry=4, ovlp=ovlp, hcore=hcore, rdm1=mf["rdm1"], vhf=mf["vhf"])
    first = hd.dmet_cycle(L, hd.ProductMods(), nocc / float(nao), True, nao)
    L.update_Ham(first["rhoT"] * 2.0, vhf=mf["vhf"])
    second = hd.dmet_cycle(L, hd.ProductMods(), nocc / float(nao), True, nao)
    assert abs(second["E_frag"] - first["E_frag"]) < 1e-8
    assert abs(second["E_frag"] - mf["e_cell"]) < 1e-8
    full = L.expand(second["rhoT"])[0]
    assert np.abs(full - full.dot(full)).max() < 1e-10                  # idempotent (test_mfd.py:117-120)
