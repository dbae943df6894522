"""Oracle: a small Hartree-Fock impurity solver on an embedding Hamiltonian, used to turn integral parity into
energy parity (north_star: <= 1e-8 Ha on the energy).  Mirrors what the reference's tests do with
libdmet.solver.scf.SCF on `ImpHam` (libdmet/test/test_mfd.py:128-138): RHF/UHF iterations with the J/K convention of
libdmet/solver/scf.py:255-352, orthonormal embedding orbitals, fixed electron number.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np
import scipy.linalg as la

from .slater import _get_veff


def hf_energy(H0, H1, H2, nelec_per_spin, max_iter=200, tol=1e-12, dm0=None):
    """H1 (spin, n, n); H2 as returned by embHam (spin_pair, ...) in the order aa, bb, ab for spin 2.
    Returns (E_total, rdm1 (spin, n, n) per spin)."""
    H1 = np.asarray(H1)
    spin, n = H1.shape[0], H1.shape[-1]
    nocc = list(nelec_per_spin) if np.ndim(nelec_per_spin) else [int(nelec_per_spin)] * spin

    def fock_and_energy(dm):
        if spin == 1:
            veff = _get_veff(dm * 2.0, H2)                      # restricted: spin-traced density (slater.py:478-486)
            F = H1 + veff
            E = H0 + np.sum((H1[0] + 0.5 * veff[0]) * (dm[0] * 2.0))
        else:
            veff = _get_veff(dm, H2)
            F = H1 + veff
            E = H0 + sum(np.sum((H1[s] + 0.5 * veff[s]) * dm[s]) for s in range(2))
        return F, E

    dm = np.zeros_like(H1)
    if dm0 is None:
        for s in range(spin):
            e, c = la.eigh(H1[s])
            dm[s] = c[:, :nocc[s]].dot(c[:, :nocc[s]].T)
    else:
        dm[...] = dm0
    E_old, errs, focks = None, [], []
    for it in range(max_iter):
        F, E = fock_and_energy(dm)
        err = np.concatenate([(F[s].dot(dm[s]) - dm[s].dot(F[s])).ravel() for s in range(spin)])
        if E_old is not None and abs(E - E_old) < tol and np.abs(err).max() < 1e-10:
            break
        E_old = E
        focks.append(F.copy())
        errs.append(err)
        focks, errs = focks[-8:], errs[-8:]
        if len(errs) > 1 and np.abs(err).max() > 1e-9:          # DIIS (plain iterations once nearly converged)
            m = len(errs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    B[i, j] = errs[i].dot(errs[j])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            c = np.linalg.lstsq(B, rhs, rcond=1e-14)[0][:m]
            F = sum(ci * Fi for ci, Fi in zip(c, focks))
        for s in range(spin):
            e, c = la.eigh(F[s])
            dm[s] = c[:, :nocc[s]].dot(c[:, :nocc[s]].T)
    _, E = fock_and_energy(dm)
    return E, dm
