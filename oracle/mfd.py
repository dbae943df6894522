"""Oracle: lattice mean field without self-consistency -- the host-side piece of a DMET iteration that turns the
lattice Fock matrix (+ correlation potential) into the density matrix the bath is built from.

Numpy restatement (own code, same algorithm) of libdmet/routine/mfd.py:33-108 (`DiagRHF`, `DiagUHF` and their k / -k
symmetric variants), 235-427 (`HF`: diagonalise per k-point, assign occupations, rho_k = C f C^dagger, FFT to the
stripe, energy per cell) and 862-957 (`check_nelec`, `assignocc`, zero-temperature branch with fractional filling of a
degenerate HOMO).  Finite temperature (`ftsystem`), the pyscf-driven `scf=True` branch and non-local correlation
potentials are outside the hot-path scope.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by the HF-in-DMET loop of the energy-parity tests, which
assert what libdmet/test/test_mfd.py:138,153,161 assert.

`lattice` is duck-typed: getFock(kspace), getH1(kspace), getH0(), FFTtoT(A), nkpts, use_hcore_as_emb_ham (optional),
cell_idx2pos / cell_pos2idx (symm=True only).  `vcor` is None (no correlation potential), or an object with
.get(i, kspace) -> (spin, n, n) and .islocal() like libdmet.dmet.Hubbard.VcorLocal.
"""
import numpy as np
import scipy.linalg as la

IMAG_DISCARD_TOL = 1e-7     # libdmet/settings.py


def _vcor_k(vcor, i, s):
    return 0.0 if vcor is None else vcor.get(i, True)[s]


def DiagRHF(Fock, vcor, **kwargs):
    """eigenpairs of Fock[0, k] (+ vcor) for every k (mfd.py:33-46)"""
    Fock = Fock[None] if Fock.ndim == 3 else Fock
    nk, n = Fock.shape[-3], Fock.shape[-1]
    ew = np.empty((nk, n))
    ev = np.empty((nk, n, n), dtype=np.complex128)
    for k in range(nk):
        ew[k], ev[k] = la.eigh(Fock[0, k] + _vcor_k(vcor, k, 0))
    return ew, ev


def DiagRHF_symm(Fock, vcor, lattice, **kwargs):
    """as DiagRHF, with the eigenpairs at -k taken as the conjugates of those at k (mfd.py:48-67)"""
    Fock = Fock[None] if Fock.ndim == 3 else Fock
    nk, n = Fock.shape[-3], Fock.shape[-1]
    ew = np.empty((nk, n))
    ev = np.empty((nk, n, n), dtype=np.complex128)
    done = set()
    for k in range(nk):
        mk = lattice.cell_pos2idx(-lattice.cell_idx2pos(k))
        if mk in done:
            ew[k], ev[k] = ew[mk], ev[mk].conj()
        else:
            ew[k], ev[k] = la.eigh(Fock[0, k] + _vcor_k(vcor, k, 0))
            done.add(k)
    return ew, ev


def DiagUHF(Fock, vcor, **kwargs):
    """both spin channels (mfd.py:69-84); a spin-less Fock stack is used for both"""
    Fock = np.asarray((Fock, Fock)) if Fock.ndim == 3 else Fock
    nk, n = Fock.shape[-3], Fock.shape[-1]
    ew = np.empty((2, nk, n))
    ev = np.empty((2, nk, n, n), dtype=np.complex128)
    for s in range(2):
        for k in range(nk):
            ew[s, k], ev[s, k] = la.eigh(Fock[s, k] + _vcor_k(vcor, k, s))
    return ew, ev


def DiagUHF_symm(Fock, vcor, lattice, **kwargs):
    """mfd.py:86-108"""
    Fock = np.asarray((Fock, Fock)) if Fock.ndim == 3 else Fock
    nk, n = Fock.shape[-3], Fock.shape[-1]
    ew = np.empty((2, nk, n))
    ev = np.empty((2, nk, n, n), dtype=np.complex128)
    done = set()
    for k in range(nk):
        mk = lattice.cell_pos2idx(-lattice.cell_idx2pos(k))
        if mk in done:
            for s in range(2):
                ew[s, k], ev[s, k] = ew[s, mk], ev[s, mk].conj()
        else:
            for s in range(2):
                ew[s, k], ev[s, k] = la.eigh(Fock[s, k] + _vcor_k(vcor, k, s))
            done.add(k)
    return ew, ev


def check_nelec(nelec, ncells=None, tol=1e-5):
    """electron number rounded to the nearest integer, with a warning when it was not one (mfd.py:860-885);
    returns (int nelec, nelec per cell or None)"""
    if abs(nelec - np.round(nelec)) > tol:
        import warnings
        warnings.warn("HF: nelec is rounded to integer nelec = %d (original %.2f)" % (np.round(nelec), nelec))
    nelec = int(np.round(nelec))
    per_cell = None
    if ncells is not None:
        per_cell = nelec / float(ncells)
        if abs(per_cell - np.round(per_cell)) <= tol:
            per_cell = int(np.round(per_cell))
    return nelec, per_cell


def assignocc(ew, nelec, beta=np.inf, mu0=0.0, fix_mu=False, thr_deg=1e-6, Sz=None, ncore=0, nvirt=0):
    """zero-temperature occupations (mfd.py:887-957): levels below mu - thr_deg are filled; electrons left over are
    spread evenly over the levels within thr_deg of mu.  `nelec` per spin for RHF, total for UHF; a pair
    (nelec_a, nelec_b) -- or Sz -- fixes the two spin channels separately."""
    if beta < np.inf:
        raise NotImplementedError("finite temperature is outside the oracle's scope")
    ew = np.asarray(ew)
    if Sz is None and not np.ndim(nelec):
        srt = np.sort(ew, axis=None, kind="mergesort")
        nelec = check_nelec(nelec, None)[0]
        if np.sum(ew < mu0 - thr_deg) <= nelec <= np.sum(ew <= mu0 + thr_deg):
            mu = mu0                                        # "we prefer not to change mu"
        else:
            mu = 0.5 * (srt[nelec - 1] + srt[nelec])
        occ = 1.0 * (ew < mu - thr_deg)
        left = nelec - np.sum(occ)
        if left > 0:
            near = np.logical_and(ew <= mu + thr_deg, ew >= mu - thr_deg)
            occ += (float(left) / np.sum(near)) * near
        return occ, mu, 0.0
    assert ew.shape[0] == 2
    if not np.ndim(nelec):
        nelec = [(nelec + Sz) * 0.5, (nelec - Sz) * 0.5]
    mu0 = list(mu0) if np.ndim(mu0) else [mu0, mu0]
    occ = np.empty_like(ew)
    mu, nerr = np.zeros(2), np.zeros(2)
    for s in range(2):
        occ[s], mu[s], nerr[s] = assignocc(ew[s], nelec[s], beta, mu0[s], fix_mu=fix_mu, thr_deg=thr_deg)
    return occ, mu, nerr


def _mid_gap(srt, nelec):
    if nelec <= 0:
        return srt[0]
    if nelec >= len(srt):
        return srt[-1]
    return 0.5 * (srt[nelec - 1] + srt[nelec])


def HF(lattice, vcor, filling, restricted, mu0=None, beta=np.inf, ires=False, scf=False, use_hcore=None, **kwargs):
    """mfd.py:235-427 (scf=False): rho (spin, ncells, n, n) per spin channel, mu, energy per cell incl. the
    correlation-potential term; with `ires` also the dict of mfd.py:421-424."""
    assert not scf, "the pyscf-driven self-consistent branch is not restated"
    if use_hcore is None:
        use_hcore = getattr(lattice, "use_hcore_as_emb_ham", False)
    if use_hcore:
        Fock = lattice.getH1(kspace=True)
        FockT = H1T = lattice.getH1(kspace=False)
    else:
        Fock, FockT, H1T = lattice.getFock(kspace=True), lattice.getFock(kspace=False), lattice.getH1(kspace=False)
    Fock = np.asarray(Fock)
    symm = kwargs.get("symm", False)
    if restricted:
        ew, ev = (DiagRHF_symm if symm else DiagRHF)(Fock, vcor, lattice=lattice)
        ew, ev = ew[None], ev[None]
    else:
        ew, ev = (DiagUHF_symm if symm else DiagUHF)(Fock, vcor, lattice=lattice)

    if np.ndim(filling):                                   # a filling per spin channel (mfd.py:303-321)
        nelec = [check_nelec(ew.size * filling[s] * 0.5, None)[0] for s in range(2)]
        srt = [np.sort(ew[s], axis=None, kind="mergesort") for s in range(2)]
        if mu0 is None:
            mu0 = [_mid_gap(srt[s], nelec[s]) for s in range(2)]
    else:
        nelec = check_nelec(ew.size * filling, None)[0]    # RHF: per spin, UHF: total
        srt = np.sort(ew, axis=None, kind="mergesort")
        if mu0 is None:
            mu0 = _mid_gap(srt, nelec)
    occ, mu, nerr = assignocc(ew, nelec, beta, mu0, fix_mu=kwargs.get("fix_mu", False),
                              thr_deg=kwargs.get("tol_deg", 1e-6))

    spin, nk = ev.shape[0], ev.shape[1]
    rho = np.empty_like(ev)
    rhoT = np.empty_like(ev)
    for s in range(spin):
        for k in range(nk):
            rho[s, k] = (ev[s, k] * occ[s, k]).dot(ev[s, k].conj().T)
        rhoT[s] = lattice.FFTtoT(rho[s])
    if np.abs(np.asarray(rhoT).imag).max() < IMAG_DISCARD_TOL:
        rhoT = rhoT.real

    def with_spin(A):
        A = np.asarray(A)
        A = A[None] if A.ndim == 3 else A
        return A if A.shape[0] == spin else np.asarray([A[0]] * spin)
    FockT, H1T = with_spin(FockT), with_spin(H1T)
    vT = None if vcor is None else np.asarray(vcor.get(0, False))
    if spin == 1:
        E0 = np.sum((FockT + H1T) * rhoT) + lattice.getH0()
        E = E0 + (0.0 if vT is None else np.sum(vT[0] * rhoT[0, 0]))
    else:
        E0 = 0.5 * np.sum((FockT + H1T) * rhoT) + lattice.getH0()
        E = E0 + (0.0 if vT is None else 0.5 * np.sum(vT[0] * rhoT[0, 0] + vT[1] * rhoT[1, 0]))
    E = float(np.real(E))
    if not ires:
        return rhoT, mu, E
    if np.ndim(mu):
        homo = tuple(srt[s][max(np.searchsorted(srt[s], mu[s], side="right") - 1, 0)] for s in range(2))
        lumo = tuple(srt[s][min(np.searchsorted(srt[s], mu[s], side="left"), len(srt[s]) - 1)] for s in range(2))
        gap = np.array((lumo[0] - homo[0], lumo[1] - homo[1]))
    else:
        homo = srt[max(np.searchsorted(srt, mu, side="right") - 1, 0)]
        lumo = srt[min(np.searchsorted(srt, mu, side="left"), len(srt) - 1)]
        gap = lumo - homo
    res = {"gap": gap, "e": ew, "coef": ev, "nerr": nerr, "rho_k": rho, "E0": float(np.real(E0)), "E": E,
           "mo_occ": occ, "homo": homo, "lumo": lumo}
    return rhoT, mu, E, res
