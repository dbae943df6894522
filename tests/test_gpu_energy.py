"""GPU: energy parity.  The same Hartree-Fock impurity solver (oracle/scf.py) run on the embedding Hamiltonian built
by the CUDA path and on the one built by the oracle must converge to the same energy within 1e-8 Ha
(the reference pins HF-in-DMET energies to 1e-8, libdmet/test/test_mfd.py:153)."""
import numpy as np
import pytest

from helpers import problem, mean_field, OracleLattice

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spin,sym,kmesh", [(1, 4, [1, 1, 3]), (1, 1, [2, 2, 1]), (2, 4, [1, 1, 3])])
def test_impurity_hf_energy(dev, spin, sym, kmesh):
    from libdmet_preview_b200 import lattice as lat, slater
    from oracle import slater as osl, scf as oscf
    nao, naux, nval = 6, 16, 4
    gdf, C, _ = problem(kmesh, nao, naux, 2, spin=spin)
    if spin == 1:
        C = C if C.ndim == 3 else C[0]
    hcore, ovlp, vhf, rdm1 = mean_field(kmesh, nao, nval // 2 + 1, spin=spin)
    L = lat.Lattice(gdf.cell, kmesh)
    L.set_val_virt_core(nval, nao - nval, 0)
    L.set_Ham(None, gdf, C, eri_symmetry=sym, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf, H0=0.25)
    O = OracleLattice(gdf, C, hcore, ovlp, rdm1, vhf, eri_symmetry=sym, H0=0.25)
    O.val_idx, O.virt_idx = list(range(nval)), list(range(nval, nao))
    rho = L.rdm1_lo_R * (0.5 if spin == 1 else 1.0)
    basis = slater.get_emb_basis(L, rho)
    Ham, _ = slater.embHam(L, basis, None)
    Ref, _ = osl.embHam(O, basis, None)
    nocc = [basis.shape[-1] // 2] * spin if spin == 1 else [basis.shape[-1] // 2, basis.shape[-1] // 2 - 1]
    E1, dm1 = oscf.hf_energy(Ham.H0, Ham.H1["cd"], Ham.H2["ccdd"], nocc)
    E2, dm2 = oscf.hf_energy(Ref.H0, Ref.H1["cd"], Ref.H2["ccdd"], nocc)
    assert abs(E1 - E2) < 1e-8, (E1, E2)
    assert np.abs(dm1 - dm2).max() < 1e-6
    assert abs(E1) > 1e-3                      # a non-trivial number
