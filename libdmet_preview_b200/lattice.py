"""The part of `libdmet.system.lattice.Lattice` the embedding-Hamiltonian path reads (lattice.py:31-56, 120-231,
304-351, 399-411, 591-673), usable without PySCF: k-mesh bookkeeping, k2R / R2k method wrappers, expand /
extract_stripe, orbital partition, and `set_Ham` from explicit mean-field matrices with the AO->LO transforms and
the seven k2R transforms running on the GPU.  A real libdmet `Lattice` works with `slater.embHam` as well -- only
attributes are read.
"""
import itertools as it

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
        assert A.shape[-3] == self.ncells
        nsc = A.shape[-1]
        nsites = nsc * self.ncells
        lead = A.shape[:-3]
        A4 = A.reshape((-1, self.ncells, nsc, nsc))
        big = np.zeros((A4.shape[0], nsites, nsites), dtype=A.dtype)
        if dense:
            rows = range(self.ncells)
        else:
            rows = [i for i in range(self.ncells) if not np.allclose(A4[:, i], 0.0)]
        for i, j in it.product(rows, range(self.ncells)):
            idx = self.add(i, j)
            big[:, idx * nsc:(idx + 1) * nsc, j * nsc:(j + 1) * nsc] = A4[:, i]
        return big.reshape(lead + (nsites, nsites))

    def extract_stripe(self, A):
        ncells = self.ncells
        nsc = A.shape[-1] // ncells
        if A.ndim == 2:
            return A.reshape((ncells, nsc, ncells, nsc))[:, :, 0]
        elif A.ndim == 3:
            return A.reshape((A.shape[0], ncells, nsc, ncells, nsc))[:, :, :, 0]
        raise ValueError("unknown shape of A, %s" % (A.shape,))

    # ---- Hamiltonian (lattice.py:416-515, 591-673) ----
    def set_Ham(self, kmf, df, C_ao_lo, eri_symmetry=4, ovlp=None, hcore=None, rdm1=None, fock=None, veff=None,
                vhf=None, vj=None, vk=None, vxc=None, use_hcore_as_emb_ham=False, H0=0.0, hcore_hf_add=None):
        """Same argument list as the reference.  Without PySCF `kmf` is None and ovlp, hcore, rdm1 and vhf (or
        vj and vk: vhf = vj - vk/2 restricted, vj[0]+vj[1]-vk unrestricted) must be given as
        ((spin,) nkpts, nao, nao) arrays."""
        self.kmf = kmf
        self.df = df
        self.C_ao_lo = np.asarray(C_ao_lo)
        if kmf is not None:
            ovlp = kmf.get_ovlp() if ovlp is None else ovlp
            hcore = kmf.get_hcore() if hcore is None else hcore
            rdm1 = kmf.make_rdm1() if rdm1 is None else rdm1
            if (vj is None or vk is None) and vhf is None:
                vj, vk = kmf.get_jk(dm_kpts=rdm1)
        if ovlp is None or hcore is None or rdm1 is None:
            raise ValueError("set_Ham: ovlp, hcore and rdm1 are required when no mean-field object is given")
        ovlp, hcore, rdm1 = np.asarray(ovlp), np.asarray(hcore), np.asarray(rdm1)
        if vhf is None:
            if vj is None or vk is None:
                raise ValueError("set_Ham: vhf (or vj and vk) is required when no mean-field object is given")
            vj, vk = np.asarray(vj), np.asarray(vk)
            vhf = vj - vk * 0.5 if vk.ndim == 3 else vj[0] + vj[1] - vk      # pbc_helper.get_veff, HF
        vhf = np.asarray(vhf)
        if veff is None:
            veff = vhf
        if fock is None:
            fock = hcore + veff
        fock_hf = hcore + vhf
        if hcore_hf_add is not None:
            fock_hf = fock_hf + hcore_hf_add
        self.ovlp_ao_k, self.hcore_ao_k, self.rdm1_ao_k = ovlp, hcore, rdm1
        self.fock_ao_k, self.fock_hf_ao_k = np.asarray(fock), np.asarray(fock_hf)
        self.hcore_hf_add = hcore_hf_add
        self.veff_ao_k, self.vhf_ao_k = np.asarray(veff), vhf
        if vxc is not None:
            self.vxc_ao_k = np.asarray(vxc)
        if self.C_ao_lo.ndim == 3:
            self.spin, self.restricted = 1, True
        else:
            self.spin = self.C_ao_lo.shape[0]
            self.restricted = (self.spin == 1)
        self.eri_symmetry = eri_symmetry
        assert self.eri_symmetry in [1, 4, 8]
        if not self.restricted:
            assert self.eri_symmetry != 8
        self.transform_obj_to_lo()
        self.H0 = H0
        self.has_Ham = True
        self.use_hcore_as_emb_ham = use_hcore_as_emb_ham

    def transform_obj_to_lo(self):
        """lattice.py:591-673 (non-GHF): six transform_h1_to_lo + transform_rdm1_to_lo, then seven k2R."""
        C = self.C_ao_lo
        if C.shape[-2] != self.hcore_ao_k.shape[-1]:
            raise NotImplementedError("generalised (GHF) coefficient layout is outside the hot path")
        t = make_basis.transform_h1_to_lo
        self.hcore_lo_k = add_spin_dim(t(self.hcore_ao_k, C), self.spin)
        self.ovlp_lo_k = t(self.ovlp_ao_k, C)
        self.fock_lo_k = add_spin_dim(t(self.fock_ao_k, C), self.spin)
        self.fock_hf_lo_k = add_spin_dim(t(self.fock_hf_ao_k, C), self.spin)
        self.veff_lo_k = add_spin_dim(t(self.veff_ao_k, C), self.spin)
        self.vhf_lo_k = add_spin_dim(t(self.vhf_ao_k, C), self.spin)
        self.rdm1_lo_k = add_spin_dim(make_basis.transform_rdm1_to_lo(self.rdm1_ao_k, C, self.ovlp_ao_k), self.spin)
        self.hcore_lo_R = self.k2R(self.hcore_lo_k)
        self.ovlp_lo_R = self.k2R(self.ovlp_lo_k)
        self.fock_lo_R = self.k2R(self.fock_lo_k)
        self.fock_hf_lo_R = self.k2R(self.fock_hf_lo_k)
        self.veff_lo_R = self.k2R(self.veff_lo_k)
        self.vhf_lo_R = self.k2R(self.vhf_lo_k)
        self.rdm1_lo_R = self.k2R(self.rdm1_lo_k)
        if self.vxc_ao_k is not None:
            self.vxc_lo_k = add_spin_dim(t(self.vxc_ao_k, C), self.spin)
            self.vxc_lo_R = self.k2R(self.vxc_lo_k)

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
