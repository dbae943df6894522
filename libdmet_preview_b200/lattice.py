"""The part of `libdmet.system.lattice.Lattice` the embedding-Hamiltonian path reads (lattice.py:31-56, 120-231,
304-351, 399-411, 591-673), usable without PySCF: k-mesh bookkeeping, k2R / R2k method wrappers, expand /
extract_stripe, orbital partition, and `set_Ham` from explicit mean-field matrices with the AO->LO transforms and
the seven k2R transforms running on the GPU.  A real libdmet `Lattice` works with `slater.embHam` as well -- only
attributes are read.
"""
import numpy as np

from . import fourier, make_basis
from .fourier import IMAG_DISCARD_TOL
from .make_basis import add_spin_dim
from .schedule import make_kpts_scaled, cell_vectors


class Lattice(object):
    def __init__(self, cell, kmesh):
        self.mol = self.cell = cell
        self.kmesh = [int(x) for x in kmesh]
        self.nscsites = self.nao = int(cell.nao_nr())
        self.dim = cell.dimension
        self.csize = np.asarray(self.kmesh)
        self.ncells = int(np.prod(self.csize))
        self.cells = cell_vectors(self.kmesh)
        self.celldict = dict(zip(map(tuple, self.cells), range(self.ncells)))
        self.kpts_scaled = make_kpts_scaled(self.kmesh)
        self.kpts = self.kpts_abs = cell.get_abs_kpts(self.kpts_scaled)
        self.nkpts = len(self.kpts)
        self.phase_R2k = fourier.get_phase_R2k_scaled(self.kmesh)            # exp(-iRk)      (lattice.py:56)
        self.phase_k2R = self.phase_R2k.conj().T / self.nkpts                # exp(+ikR) / N  (lattice.py:55)
        self.val_idx, self.virt_idx, self.core_idx = [], [], []
        self.df = None
        self.C_ao_lo = None
        self.kmf = None
        self.JK_imp = self.JK_emb = self.JK_core = None
        self.has_Ham = False
        self.restricted = None
        self.is_model = False
        self.H0 = 0.0
        self.use_hcore_as_emb_ham = False
        self.vxc_ao_k = self.vxc_lo_k = self.vxc_lo_R = None

    # ---- orbital partition (lattice.py:99-156) ----
    @property
    def ncore(self):
        return len(self.core_idx)

    @property
    def nval(self):
        return len(self.val_idx)

    @property
    def nvirt(self):
        return len(self.virt_idx)

    @property
    def nimp(self):
        return self.nval + self.nvirt

    @property
    def imp_idx(self):
        return list(self.val_idx) + list(self.virt_idx)

    def set_val_virt_core(self, val, virt, core):
        self.core_idx = list(core) if hasattr(core, "__iter__") else list(range(0, core))
        self.val_idx = list(val) if hasattr(val, "__iter__") else list(range(self.ncore, self.ncore + val))
        self.virt_idx = list(virt) if hasattr(virt, "__iter__") else \
            list(range(self.ncore + self.nval, self.ncore + self.nval + virt))

    # ---- cell index arithmetic (lattice.py:192-207) ----
    def cell_idx2pos(self, idx):
        return self.cells[idx % self.ncells]

    def cell_pos2idx(self, pos):
        return self.celldict[tuple(pos % self.csize)]

    def add(self, i, j):
        return self.cell_pos2idx(self.cell_idx2pos(i) + self.cell_idx2pos(j))

    def subtract(self, i, j):
        return self.cell_pos2idx(self.cell_idx2pos(i) - self.cell_idx2pos(j))

    # ---- Fourier wrappers (lattice.py:209-219, 399-411) ----
    def FFTtoK(self, A):
        return fourier.FFTtoK(A, self.kmesh)

    def FFTtoT(self, B, tol=IMAG_DISCARD_TOL):
        return fourier.FFTtoT(B, self.kmesh, tol=tol)

    def k2R(self, A, tol=IMAG_DISCARD_TOL):
        return fourier.k2R(A, self.kmesh, tol=tol)

    def R2k(self, B):
        return fourier.R2k(B, self.kmesh)

    def k2R_basis(self, basis_k):
        return self.k2R(basis_k)

    def R2k_basis(self, basis_R):
        return self.R2k(basis_R)

    # ---- stripe <-> full (lattice.py:304-351) ----
    def expand(self, A, dense=False):
        """stripe ((spin,) ncells, n, n) -> full translation-invariant matrix: block (i + j, j) = A[i]; all-zero
        stripes are skipped unless `dense` (same result, the reference's shortcut)"""
        assert A.shape[-3] == self.ncells
        n = A.shape[-1]
        stripes = A.reshape((-1, self.ncells, n, n))
        full = np.zeros((stripes.shape[0], self.ncells * n, self.ncells * n), dtype=A.dtype)
        target = np.asarray([[self.add(i, j) for j in range(self.ncells)] for i in range(self.ncells)])
        for i in range(self.ncells):
            if not dense and np.allclose(stripes[:, i], 0.0):
                continue
            for j in range(self.ncells):
                r = target[i, j]
                full[:, r * n:(r + 1) * n, j * n:(j + 1) * n] = stripes[:, i]
        return full.reshape(A.shape[:-3] + full.shape[1:])

    def extract_stripe(self, A):
        """first block column of a full matrix"""
        if A.ndim not in (2, 3):
            raise ValueError("unknown shape of A, %s" % (A.shape,))
        n = A.shape[-1] // self.ncells
        return A.reshape(A.shape[:-2] + (self.ncells, n, self.ncells, n))[..., :, :, 0, :]

    # ---- Hamiltonian (lattice.py:416-515, 591-673) ----
    _H1_NAMES = ("hcore", "fock", "fock_hf", "veff", "vhf")      # transformed with C^dagger . C and given a spin axis

    def set_Ham(self, kmf, df, C_ao_lo, eri_symmetry=4, ovlp=None, hcore=None, rdm1=None, fock=None, veff=None,
                vhf=None, vj=None, vk=None, vxc=None, use_hcore_as_emb_ham=False, H0=0.0, hcore_hf_add=None):
        """Same argument list as the reference.  Without PySCF pass kmf=None together with ovlp, hcore, rdm1 and
        vhf (or vj and vk: vhf = vj - vk/2 restricted, vj[0] + vj[1] - vk unrestricted) as
        ((spin,) nkpts, nao, nao) arrays."""
        self.kmf, self.df = kmf, df
        self.C_ao_lo = np.asarray(C_ao_lo)
        if kmf is not None:                               # pull what was not supplied from the mean-field object
            ovlp = kmf.get_ovlp() if ovlp is None else ovlp
            hcore = kmf.get_hcore() if hcore is None else hcore
            rdm1 = kmf.make_rdm1() if rdm1 is None else rdm1
            if vhf is None and (vj is None or vk is None):
                vj, vk = kmf.get_jk(dm_kpts=rdm1)
        missing = [n for n, v in (("ovlp", ovlp), ("hcore", hcore), ("rdm1", rdm1)) if v is None]
        if missing:
            raise ValueError("set_Ham: %s required when no mean-field object is given" % ", ".join(missing))
        if vhf is None:
            if vj is None or vk is None:
                raise ValueError("set_Ham: vhf (or vj and vk) is required when no mean-field object is given")
            vj, vk = np.asarray(vj), np.asarray(vk)
            vhf = vj - 0.5 * vk if vk.ndim == 3 else vj[0] + vj[1] - vk
        self.ovlp_ao_k, self.hcore_ao_k, self.rdm1_ao_k = (np.asarray(x) for x in (ovlp, hcore, rdm1))
        self.vhf_ao_k = np.asarray(vhf)
        self.veff_ao_k = self.vhf_ao_k if veff is None else np.asarray(veff)
        self.fock_ao_k = self.hcore_ao_k + self.veff_ao_k if fock is None else np.asarray(fock)
        self.fock_hf_ao_k = self.hcore_ao_k + self.vhf_ao_k
        self.hcore_hf_add = hcore_hf_add
        if hcore_hf_add is not None:
            self.fock_hf_ao_k = self.fock_hf_ao_k + hcore_hf_add
        if vxc is not None:
            self.vxc_ao_k = np.asarray(vxc)
        self.spin = 1 if self.C_ao_lo.ndim == 3 else self.C_ao_lo.shape[0]
        self.restricted = (self.spin == 1)
        if eri_symmetry not in (1, 4, 8) or (eri_symmetry == 8 and not self.restricted):
            raise AssertionError("eri_symmetry must be 1, 4 or (restricted only) 8")
        self.eri_symmetry = eri_symmetry
        self.transform_obj_to_lo()
        self.H0 = H0
        self.has_Ham = True
        self.use_hcore_as_emb_ham = use_hcore_as_emb_ham

    def transform_obj_to_lo(self):
        """AO -> LO for hcore, ovlp, fock, fock_hf, veff, vhf (C^dagger h C) and rdm1 (C^-1 D C^-dagger), then the
        seven k -> R transforms; all on the device (lattice.py:591-673, non-GHF)."""
        C = self.C_ao_lo
        if C.shape[-2] != self.hcore_ao_k.shape[-1]:
            raise NotImplementedError("generalised (GHF) coefficient layout is outside the hot path")
        for name in self._H1_NAMES:
            lo_k = make_basis.transform_h1_to_lo(getattr(self, name + "_ao_k"), C)
            setattr(self, name + "_lo_k", add_spin_dim(lo_k, self.spin))
        self.ovlp_lo_k = make_basis.transform_h1_to_lo(self.ovlp_ao_k, C)
        self.rdm1_lo_k = add_spin_dim(make_basis.transform_rdm1_to_lo(self.rdm1_ao_k, C, self.ovlp_ao_k), self.spin)
        for name in self._H1_NAMES + ("ovlp", "rdm1"):
            setattr(self, name + "_lo_R", self.k2R(getattr(self, name + "_lo_k")))
        if self.vxc_ao_k is not None:
            self.vxc_lo_k = add_spin_dim(make_basis.transform_h1_to_lo(self.vxc_ao_k, C), self.spin)
            self.vxc_lo_R = self.k2R(self.vxc_lo_k)

    def update_Ham(self, rdm1_lo_R, veff=None, vhf=None, **kwargs):
        """lattice.py:565-589 (Knizia's charge self-consistency step): the DMET density matrix replaces the
        mean-field one, rdm1_lo_R -> rdm1_lo_k -> rdm1_ao_k on the device, and every LO-basis quantity is rebuilt
        by `set_Ham`.  With `veff` / `vhf` supplied the stored J and K are kept, exactly as in the reference (there
        `vhf` is then re-formed from the OLD vj, vk -- i.e. it is the stored `vhf_ao_k`); without them J and K have
        to be rebuilt from the new density by the mean-field object."""
        self.rdm1_lo_R = rdm1_lo_R
        self.rdm1_lo_k = self.R2k(self.rdm1_lo_R)
        self.rdm1_ao_k = make_basis.transform_rdm1_to_ao(self.rdm1_lo_k, self.C_ao_lo)
        if veff is None and vhf is None:
            if self.kmf is None:
                raise ValueError("update_Ham without veff / vhf rebuilds J and K from the new density and needs the "
                                 "mean-field object (kmf)")
            vhf_use = None
        else:
            vhf_use = self.vhf_ao_k if vhf is None else vhf
        self.set_Ham(self.kmf, self.df, self.C_ao_lo, self.eri_symmetry, ovlp=self.ovlp_ao_k, hcore=self.hcore_ao_k,
                     rdm1=self.rdm1_ao_k, fock=None, veff=veff, vhf=vhf_use, vxc=None, H0=self.H0,
                     use_hcore_as_emb_ham=self.use_hcore_as_emb_ham, hcore_hf_add=self.hcore_hf_add)

    def expand_orb(self, C):
        """translation-invariant orbitals C[T] ((spin,) ncells, nao, nmo) -> the full supercell coefficient matrix,
        block (T + R, R) = C[T] (lattice.py:353-377)"""
        C = np.asarray(C)
        if C.ndim not in (3, 4):
            raise ValueError("unknown shape of C, %s" % (C.shape,))
        assert C.shape[-3] == self.ncells
        nao, nmo = C.shape[-2:]
        big = np.zeros(C.shape[:-3] + (self.ncells * nao, self.ncells * nmo), dtype=C.dtype)
        for i in range(self.ncells):
            for j in range(self.ncells):
                r = self.add(i, j)
                big[..., r * nao:(r + 1) * nao, j * nmo:(j + 1) * nmo] = C[..., i, :, :]
        return big

    def transpose(self, A):
        """transpose of a translation-invariant matrix in stripe form: A^T[n] = A[-n]^T (lattice.py:379-397)"""
        A = np.asarray(A)
        if A.ndim not in (3, 4):
            raise ValueError("unknown shape of A, %s" % (A.shape,))
        minus = [self.cell_pos2idx(-self.cell_idx2pos(n)) for n in range(self.ncells)]
        return np.ascontiguousarray(np.swapaxes(A[..., minus, :, :], -1, -2))

    def update_lo(self, C_ao_lo):
        self.C_ao_lo = np.asarray(C_ao_lo)
        self.transform_obj_to_lo()

    def getH0(self):
        return self.H0

    def getH1(self, kspace=True):
        return self.hcore_lo_k if kspace else self.hcore_lo_R

    def getFock(self, kspace=True):
        return self.fock_lo_k if kspace else self.fock_lo_R

    def get_ovlp(self, kspace=True):
        return self.ovlp_lo_k if kspace else self.ovlp_lo_R

    def get_JK_imp(self):
        return self.JK_imp

    getImpJK = get_JK_imp

    def get_JK_core(self):
        return self.JK_core
