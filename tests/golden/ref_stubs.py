"""Stub environment that lets the UNMODIFIED reference modules under /root/reference be imported and executed in a
container without PySCF / h5py (used only by tests/golden/make_golden.py to generate fixtures; never by the product).

* every `pyscf.*`, `h5py`, `matplotlib.*`, `mpi4py*`, `block2`, ... module that the reference imports is fabricated on
  demand; unknown attributes resolve to inert placeholder classes so that `class X(pyscf.scf.hf.RHF)` style
  definitions import cleanly;
* the handful of PySCF helpers the embedding-Hamiltonian path actually CALLS are bound to the numpy restatements of
  `oracle/pyscf_lib.py` (pack_tril, unpack_tril, hermi_sum, dot, r_e2, _conc_mos, ao2mo.restore, hf.dot_eri_dm,
  cartesian_prod, KPT_DIFF_TOL) -- so the fixtures pin the reference's own control flow (schedule, symmetrisation,
  weights, spin ordering, chunking, Fourier conventions, SVD bath, embHam assembly), not PySCF's C arithmetic;
* the GDF tensor comes from an in-memory provider through stubs of `pyscf.pbc.df.df._load3c`, `pyscf.df.addons.load`
  and `h5py.File`.
"""
import contextlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
_STUB_ROOTS = ("pyscf", "h5py", "matplotlib", "mpi4py", "mpi4pyscf", "block2", "pyblock2", "libdmet_solid",
               "seaborn", "ase", "spglib", "numba_stub")

GDF_REGISTRY = {}     # cderi key -> provider
H5_WRITTEN = {}       # file name -> {dataset name: array} written through the h5py stub
FILE_NAO = {}         # path of a real cderi file -> nao (PySCF's loader knows it from the cell)


class _PlaceholderMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _Placeholder(object, metaclass=_PlaceholderMeta):
    """inert base for fabricated attributes: usable as base class, callable, attribute chains allowed"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        if name[:1].isupper():      # CamelCase -> a class (usable as a base class)
            cls = type(name, (_Placeholder,), {"__module__": self.__name__})
            setattr(self, name, cls)
            return cls
        import importlib            # lower case -> a (callable) sub-module
        return importlib.import_module(full)

    def __call__(self, *a, **k):
        return _Placeholder()


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        parent, _, child = module.__name__.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, module)
        _populate(module)


def _populate(m):
    from oracle import pyscf_lib as olib
    name = m.__name__
    if name == "pyscf.lib":
        m.pack_tril = olib.pack_tril
        m.unpack_tril = olib.unpack_tril
        m.hermi_sum = olib.hermi_sum
        m.dot = olib.dot
        m.einsum = np.einsum
        m.cartesian_prod = olib.cartesian_prod
        m.HERMITIAN, m.ANTIHERMI, m.SYMMETRIC = olib.HERMITIAN, olib.ANTIHERMI, olib.SYMMETRIC
        m.current_memory = lambda: (0.0, 0.0)

        def prange(start, end, step):
            for i in range(start, end, step):
                yield i, min(i + step, end)
        m.prange = prange
        m.map_with_prefetch = lambda func, *iterables: map(func, *iterables)
        m.logger = importlib.import_module("pyscf.lib.logger")
        m.param = importlib.import_module("pyscf.lib.param")
    elif name == "pyscf.lib.param":
        m.BOHR = 0.52917721092
    elif name == "pyscf.lib.logger":
        m.debug3 = m.debug2 = m.debug1 = m.debug = m.info = m.note = m.warn = lambda *a, **k: None
        m.DEBUG = 5
    elif name == "pyscf.pbc.lib.kpts_helper":
        m.KPT_DIFF_TOL = olib.KPT_DIFF_TOL
        m.is_zero = lambda k: bool(np.abs(np.asarray(k)).max() < olib.KPT_DIFF_TOL) if np.size(k) else True
        m.gamma_point = m.is_zero
        m.member = lambda kpt, kpts: np.where(np.abs(np.asarray(kpts) - np.asarray(kpt)).max(axis=1) < 1e-6)[0]
        m.unique = lambda kpts: np.unique(np.asarray(kpts).round(6), axis=0, return_index=True, return_inverse=True)
    elif name == "pyscf.ao2mo":
        m.restore = olib.restore
        m._ao2mo = importlib.import_module("pyscf.ao2mo._ao2mo")
        m.incore = importlib.import_module("pyscf.ao2mo.incore")
    elif name == "pyscf.ao2mo._ao2mo":
        m.r_e2 = lambda Lpq, mo, sl, tao, ao_loc, out=None: olib.r_e2(Lpq, mo, sl, out=out)
    elif name == "pyscf.ao2mo.incore":
        m._conc_mos = lambda moi, moj, compact=False: (False, False) + olib.conc_mos(moi, moj)
    elif name == "pyscf.scf.hf":
        m.dot_eri_dm = olib.dot_eri_dm
    elif name == "pyscf.pbc.df":
        class AFTDF(object):
            pass

        class FFTDF(object):
            pass

        class GDF(AFTDF):
            """in-memory GDF: `_cderi` is a registry key"""
            blockdim = 240
            max_memory = 4000

            def __init__(self, cell=None, kpts=None):
                self.cell, self.kpts, self._cderi = cell, kpts, None

            def build(self):
                return self

        class MDF(GDF):
            pass
        m.AFTDF, m.FFTDF, m.GDF, m.MDF = AFTDF, FFTDF, GDF, MDF
        m.df = importlib.import_module("pyscf.pbc.df.df")
    elif name == "pyscf.pbc.df.df":
        @contextlib.contextmanager
        def _load3c(cderi, label, kpti_kptj, kptij_label=None):
            if cderi not in GDF_REGISTRY:        # a real file on disk: PySCF's lookup restated over h5lite
                from libdmet_preview_b200 import h5lite
                with h5lite.File(cderi) as feri:
                    yield olib.load3c(feri, label, kpti_kptj, kptij_label, FILE_NAO[cderi])
                return
            prov = GDF_REGISTRY[cderi]
            ks = prov.cell.get_scaled_kpts(np.asarray(kpti_kptj))
            from oracle.fourier import kpt_member
            ki = int(kpt_member(ks[0], prov.kpts_scaled)[0])
            kj = int(kpt_member(ks[1], prov.kpts_scaled)[0])
            blk = prov.load(ki, kj)
            if ki == kj:      # PySCF stores k_i == k_j blocks s2-packed; sr_loop unpacks them (eri_transform.py:216-217)
                yield olib.pack_tril(blk)
            else:
                yield blk.reshape(prov.naux, -1)
        m._load3c = _load3c
    elif name == "pyscf.df.addons":
        @contextlib.contextmanager
        def load(cderi, dataname):
            if cderi not in GDF_REGISTRY:
                from libdmet_preview_b200 import h5lite
                with h5lite.File(cderi) as feri:
                    yield feri[dataname]
                return
            prov = GDF_REGISTRY[cderi]
            yield np.empty((prov.naux, 1))
        m.load = load
    elif name == "h5py":
        from libdmet_preview_b200 import h5lite
        Group = h5lite.Group       # the reference asks isinstance(entry, h5py.Group) (eri_transform.py:172)

        class _RealOut(object):
            """mode "w" on a name ending in .h5: a real file through h5lite.Writer"""

            def __init__(self, fname):
                self.w = h5lite.Writer(fname)

            def __setitem__(self, key, value):
                self.w[key] = np.asarray(value)

            def close(self):
                self.w.close()

        class File(object):
            """mode "r": a view of a registered in-memory GDF (only `j3c-kptij` is read through h5py by the reference,
            the blocks go through `_load3c`), or -- for a name that is not registered -- the real file, opened with
            h5lite; mode "w": datasets land in H5_WRITTEN[fname] (real file when the name ends in .h5)"""

            def __new__(cls, fname, mode="r"):
                if mode == "w" and fname.endswith(".h5"):
                    return _RealOut(fname)
                if mode != "w" and fname not in GDF_REGISTRY:
                    return h5lite.File(fname)
                return object.__new__(cls)

            def __init__(self, fname, mode="r"):
                self.mode = mode
                if mode == "w":
                    self.out = H5_WRITTEN[fname] = {}
                else:
                    self.prov = GDF_REGISTRY[fname]

            def __enter__(self):
                return self

            def __exit__(self, *a):
                return False

            def close(self):
                pass

            def __getitem__(self, key):
                if key == "j3c-kptij":      # absolute k-point pairs, j <= i (PySCF's file order)
                    k = np.asarray(self.prov.kpts)
                    n = len(k)
                    return np.asarray([(k[i], k[j]) for i in range(n) for j in range(i + 1)])
                raise KeyError(key)

            def __setitem__(self, key, value):
                self.out[key] = np.array(value)

            def __contains__(self, key):
                return self.mode != "w" and key == "j3c-kptij"
        m.File, m.Group = File, Group
    elif name == "pyscf.pbc.tools":
        def super_cell(cell, kmesh):
            return cell.super_cell(kmesh)
        m.super_cell = super_cell


# sub-modules whose public names the reference's package `__init__` files re-export (looked up lazily)
_REEXPORT = {
    "libdmet.utils": ["misc", "logger"],
    "libdmet.lo": ["lowdin", "iao"],
    "libdmet.basis_transform": ["make_basis", "eri_transform"],
}


class _RefNamespace(types.ModuleType):
    """a reference package without its eager `__init__`: attributes resolve to sub-modules or to names the real
    `__init__` would have re-exported"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        import importlib
        try:
            return importlib.import_module(self.__name__ + "." + name)
        except ModuleNotFoundError as e:
            if e.name != self.__name__ + "." + name:
                raise
        for sub in _REEXPORT.get(self.__name__, []):
            mod = importlib.import_module(self.__name__ + "." + sub)
            if hasattr(mod, name):
                return getattr(mod, name)
        raise AttributeError("%s has no attribute %s" % (self.__name__, name))


def install():
    """Put the stubs and the reference on the import path.  `libdmet` and its sub-packages are registered as bare
    namespace packages so that their eager `__init__` imports (solvers, plotting, ...) do not run."""
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("the reference tree %s is needed to generate golden fixtures" % REF_ROOT)
    for pkg in ("libdmet", "libdmet.basis_transform", "libdmet.system", "libdmet.routine", "libdmet.utils",
                "libdmet.lo", "libdmet.solver", "libdmet.dmet", "libdmet.integral"):
        if pkg not in sys.modules:
            mod = _RefNamespace(pkg)
            mod.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
            sys.modules[pkg] = mod
            parent, _, child = pkg.rpartition(".")
            if parent:
                setattr(sys.modules[parent], child, mod)
    sys.modules["libdmet"].__version__ = "0.5"
    import importlib as il
    sys.modules["libdmet"].settings = il.import_module("libdmet.settings")
