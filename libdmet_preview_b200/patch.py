"""Drop-in installation into an existing libDMET (when `libdmet` and PySCF are importable):

    import libdmet_preview_b200.patch as p; p.install()

rebinds the names the reference's callers actually resolve:
  * `libdmet.basis_transform.eri_transform.get_emb_eri / get_unit_eri / get_emb_eri_fast_gdf / transform_gdf_to_lo`
  * `libdmet.routine.slater.get_emb_eri / get_unit_eri` -- imported BY NAME at module load (slater.py:32-33), so the
    attribute on `slater` must be replaced too
  * `libdmet.system.fourier.k2R / R2k / FFTtoK / FFTtoT` and the copies `libdmet.system.lattice` pulled in with
    `from libdmet.system.fourier import *` (lattice.py:23)
  * `libdmet.basis_transform.make_basis.transform_h1_to_lo / multiply_basis` (looked up through the module at call
    time, lattice.py:609)
  * `libdmet.routine.slater.get_emb_basis / embBasis / get_emb_Ham / embHam`
  * `libdmet.routine.slater.get_H_dmet / transformResults / get_rho_glob_R / get_rho_glob_k` (the last two also in
    `libdmet.routine.slater_helper`, from where `slater` star-imports them)
  * `libdmet.routine.spinless.get_emb_basis / embBasis / get_emb_Ham / embHam / get_H_dmet / transformResults` and
    `libdmet.basis_transform.eri_transform.get_emb_eri_gso` (resolved inside `spinless.__embHam2e` at call time)
Non-GDF density-fitting objects keep going to the reference's own drivers, and every branch this package does not
mirror falls through to the reference's function: those branches raise `UnsupportedBranch` (a NotImplementedError
subclass) from argument checks at the top of the call, before the embedding ERI or anything else of consequence has
been computed.  Only that exception triggers the fallback -- a NotImplementedError from torch or CUDA is a real
failure and propagates.  `uninstall()` restores the originals.
"""
import functools

from ._lib import UnsupportedBranch

_saved = {}


def _with_fallback(ours, ref):
    """`ours`, except that a branch we do not mirror (UnsupportedBranch) is served by the reference's function"""
    @functools.wraps(ours)
    def call(*args, **kwargs):
        try:
            return ours(*args, **kwargs)
        except UnsupportedBranch:
            return ref(*args, **kwargs)
    return call


def _swap(mod, name, new):
    _saved.setdefault((mod, name), getattr(mod, name))
    setattr(mod, name, new)


def install():
    import libdmet.basis_transform.eri_transform as r_eri
    import libdmet.basis_transform.make_basis as r_mb
    import libdmet.routine.slater as r_sl
    import libdmet.system.fourier as r_f
    import libdmet.system.lattice as r_lat
    from pyscf.pbc import df as pdf
    from . import eri_transform as eri, fourier, make_basis, slater

    ref_get_emb_eri = r_eri.get_emb_eri

    def get_emb_eri(cell, mydf, *args, **kwargs):
        gdf_like = hasattr(mydf, "load") or (isinstance(mydf, pdf.GDF) and not isinstance(mydf, pdf.MDF))
        if gdf_like and kwargs.get("incore", True):
            return eri.get_emb_eri(cell, mydf, *args, **kwargs)
        return ref_get_emb_eri(cell, mydf, *args, **kwargs)       # MDF / FFTDF / AFTDF / outcore: reference route

    def get_unit_eri(cell, mydf, *args, **kwargs):
        gdf_like = hasattr(mydf, "load") or (isinstance(mydf, pdf.GDF) and not isinstance(mydf, pdf.MDF))
        if gdf_like and kwargs.get("incore", True):
            return eri.get_unit_eri(cell, mydf, *args, **kwargs)
        return _saved[(r_eri, "get_unit_eri")](cell, mydf, *args, **kwargs)

    _swap(r_eri, "get_unit_eri", get_unit_eri)
    _swap(r_eri, "get_emb_eri", get_emb_eri)
    _swap(r_eri, "get_emb_eri_fast_gdf", eri.get_emb_eri_fast_gdf)
    _swap(r_eri, "transform_gdf_to_lo", eri.transform_gdf_to_lo)
    _swap(r_sl, "get_emb_eri", get_emb_eri)
    _swap(r_sl, "get_unit_eri", get_unit_eri)
    for m in (r_f, r_lat):
        for n in ("k2R", "R2k", "FFTtoK", "FFTtoT"):
            _swap(m, n, getattr(fourier, n))
    _swap(r_mb, "transform_h1_to_lo", make_basis.transform_h1_to_lo)
    _swap(r_mb, "multiply_basis", make_basis.multiply_basis)
    ref_get_emb_basis = r_sl.get_emb_basis

    def get_emb_basis(lattice, rho=None, local=True, kind='svd', **kwargs):
        if not local or kwargs.get("localize_bath") is not None:      # model-Hamiltonian baths: reference route
            return ref_get_emb_basis(lattice, rho, local=local, kind=kind, **kwargs)
        return slater.get_emb_basis(lattice, rho, local=local, kind=kind, **kwargs)

    for n in ("get_emb_basis", "embBasis"):
        _swap(r_sl, n, get_emb_basis)
    ham = _with_fallback(slater.get_emb_Ham, r_sl.get_emb_Ham)
    for n in ("get_emb_Ham", "embHam"):
        _swap(r_sl, n, ham)
    # energy side and global density matrix
    import libdmet.routine.slater_helper as r_slh
    import libdmet.routine.spinless as r_sp
    from . import spinless
    for n in ("get_H_dmet", "transformResults"):
        _swap(r_sl, n, _with_fallback(getattr(slater, n), getattr(r_sl, n)))
    for n in ("get_rho_glob_R", "get_rho_glob_k"):
        f = _with_fallback(getattr(slater, n), getattr(r_slh, n))
        _swap(r_slh, n, f)
        _swap(r_sl, n, f)
    # generalised spin orbitals
    basis_gso = _with_fallback(spinless.get_emb_basis, r_sp.get_emb_basis)
    for n in ("get_emb_basis", "embBasis"):
        _swap(r_sp, n, basis_gso)
    ham_gso = _with_fallback(spinless.get_emb_Ham, r_sp.get_emb_Ham)
    for n in ("get_emb_Ham", "embHam"):
        _swap(r_sp, n, ham_gso)
    for n in ("get_H_dmet", "transformResults"):
        _swap(r_sp, n, _with_fallback(getattr(spinless, n), getattr(r_sp, n)))
    ref_gso = r_eri.get_emb_eri_gso

    def get_emb_eri_gso(cell, mydf, *args, **kwargs):
        gdf_like = hasattr(mydf, "load") or (isinstance(mydf, pdf.GDF) and not isinstance(mydf, pdf.MDF))
        if gdf_like and kwargs.get("incore", True):
            return eri.get_emb_eri_gso(cell, mydf, *args, **kwargs)
        return ref_gso(cell, mydf, *args, **kwargs)

    _swap(r_eri, "get_emb_eri_gso", get_emb_eri_gso)
    return sorted("%s.%s" % (m.__name__, n) for (m, n) in _saved)


def uninstall():
    for (mod, name), old in _saved.items():
        setattr(mod, name, old)
    _saved.clear()
