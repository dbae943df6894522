"""Shared builders of seeded synthetic problems for the tests."""
import numpy as np

from libdmet_preview_b200 import synthetic
from oracle import fourier as o_fourier, make_basis as o_mb


def problem(kmesh, nao, naux, neo, spin=1, seed=0, nlo=None):
    nlo = nao if nlo is None else nlo
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=100 + seed)
    if spin == 1:
        C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=200 + seed)
    else:
        C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=200 + seed, spin=spin)
    basis = synthetic.make_emb_basis(kmesh, nlo, neo, seed=300 + seed, spin=spin)
    return gdf, C, basis


class OracleLattice(o_fourier.StripeLattice):
    """duck-typed lattice for the oracle's embHam: same inputs as libdmet_preview_b200.lattice.Lattice.set_Ham,
    transformed with the oracle's own (numpy/scipy) routines (lattice.py:591-673)."""

    def __init__(self, gdf, C_ao_lo, hcore, ovlp, rdm1, vhf, eri_symmetry=4, H0=0.0):
        super().__init__(gdf.kmesh, gdf.nao)
        self.cell, self.df, self.C_ao_lo = gdf.cell, gdf, np.asarray(C_ao_lo)
        self.is_model = False
        self.eri_symmetry = eri_symmetry
        self.H0 = H0
        self.spin = 1 if self.C_ao_lo.ndim == 3 else self.C_ao_lo.shape[0]
        t = o_mb.transform_h1_to_lo
        add = o_mb.add_spin_dim
        C = self.C_ao_lo
        self.hcore_lo_k = add(t(hcore, C), self.spin)
        self.ovlp_lo_k = t(ovlp, C)
        self.vhf_lo_k = add(t(vhf, C), self.spin)
        self.fock_lo_k = self.hcore_lo_k + self.vhf_lo_k
        Cs = add(C, self.spin)
        d = add(np.asarray(rdm1), self.spin)
        out = np.zeros((self.spin, self.nkpts, C.shape[-1], C.shape[-1]), dtype=np.complex128)
        for s in range(self.spin):
            for k in range(self.nkpts):
                Cinv = Cs[s, k].conj().T.dot(ovlp[k])
                out[s, k] = Cinv.dot(d[s, k]).dot(Cinv.conj().T)       # make_basis.py:591-620
        self.rdm1_lo_k = out
        self.rdm1_lo_R = self.k2R(self.rdm1_lo_k)
        self.val_idx, self.virt_idx, self.core_idx = [], [], []
        self.JK_core = None

    @property
    def imp_idx(self):
        return list(self.val_idx) + list(self.virt_idx)

    def getH0(self):
        return self.H0

    # what mfd.HF reads (lattice.py:209-231)
    def getH1(self, kspace=True):
        return self.hcore_lo_k if kspace else self.k2R(self.hcore_lo_k)

    def getFock(self, kspace=True):
        return self.fock_lo_k if kspace else self.k2R(self.fock_lo_k)

    def FFTtoT(self, A):
        return o_fourier.FFTtoT(A, self.kmesh)

    def FFTtoK(self, A):
        return o_fourier.FFTtoK(A, self.kmesh)


def mean_field(kmesh, nao, nocc, spin=1, seed=0):
    """hcore, ovlp (identity: orthonormal AOs), vhf, rdm1 in the AO basis with time-reversal structure."""
    nk = int(np.prod(kmesh))
    ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * nk)
    if spin == 1:
        hcore = synthetic.make_hermitian_k(kmesh, nao, seed=400 + seed)
        vhf = synthetic.make_hermitian_k(kmesh, nao, seed=500 + seed, scale=0.3)
        rdm1 = synthetic.make_rdm1_k(hcore + vhf, nocc) * 2.0
    else:
        hcore = synthetic.make_hermitian_k(kmesh, nao, seed=400 + seed)
        vhf = synthetic.make_hermitian_k(kmesh, nao, seed=500 + seed, scale=0.3, spin=spin)
        rdm1 = np.asarray([synthetic.make_rdm1_k(hcore + vhf[s], nocc + (1 - s)) for s in range(spin)])
    return hcore, ovlp, vhf, rdm1


class GSOLattice(object):
    """duck-typed lattice for the generalised-spin-orbital embedding Hamiltonian: (3, nkpts, nao, nao) stacks
    (aa, bb, ab) for hcore / ovlp / fock_hf and one generalised (nkpts, 2 nao, 2 nao) density matrix, all seeded.
    `fourier_mod` supplies R2k (oracle.fourier or libdmet_preview_b200.fourier)."""

    def __init__(self, gdf, C_ao_lo, fourier_mod, eri_symmetry=4, seed=0, H0=0.5):
        self.kmesh = list(gdf.kmesh)
        self.cell, self.df, self.C_ao_lo = gdf.cell, gdf, np.asarray(C_ao_lo)
        self.nscsites = self.nao = gdf.nao
        self.ncells = self.nkpts = len(gdf.kpts_scaled)
        self.is_model = False
        self.eri_symmetry = eri_symmetry
        self.use_hcore_as_emb_ham = False
        self.JK_core = None
        self._H0 = H0
        self._f = fourier_mod
        nao = gdf.nao

        def stack3(sd, scale):
            aa = synthetic.make_hermitian_k(self.kmesh, nao, seed=sd, scale=scale)
            bb = -synthetic.make_hermitian_k(self.kmesh, nao, seed=sd + 1, scale=scale)
            big = synthetic.make_hermitian_k(self.kmesh, 2 * nao, seed=sd + 2, scale=0.2 * scale)
            return np.asarray((aa, bb, big[:, :nao, nao:]))
        self.hcore_lo_k = stack3(600 + seed, 1.0)
        self.fock_hf_lo_k = self.hcore_lo_k + stack3(610 + seed, 0.3)
        eye = np.asarray([np.eye(nao, dtype=np.complex128)] * self.nkpts)
        self.ovlp_lo_k = np.asarray((eye, eye, 0 * eye))
        self.rdm1_lo_k = synthetic.make_rdm1_k(synthetic.make_hermitian_k(self.kmesh, 2 * nao, seed=620 + seed), nao)
        # energy side (get_H_dmet / transformResults): all but the last orbital of cell 0 form the impurity, so that
        # the embedding space has impurity AND environment spin orbitals and every scaling branch is exercised
        self.imp_idx = list(range(nao - 1))
        self.nimp = nao - 1

    def R2k_basis(self, basis):
        return self._f.R2k(basis, self.kmesh)

    def getH1(self, kspace=True):
        return self.hcore_lo_k

    def getFock(self, kspace=True):
        return self.fock_hf_lo_k

    def get_ovlp(self, kspace=True):
        return self.ovlp_lo_k

    def get_JK_imp(self):
        return None

    getImpJK = get_JK_imp

    def getH0(self):
        return self._H0


def gso_basis(kmesh, nao, nemb, seed=0):
    """real orthonormal GSO basis (ncells, 2 nao, nemb)"""
    nk = int(np.prod(kmesh))
    q, _ = np.linalg.qr(np.random.default_rng(700 + seed).standard_normal((nk * 2 * nao, nemb)))
    return q.reshape(nk, 2 * nao, nemb)


# cderi-file cases of tests/golden/make_golden.py (FILE_CASES): layout options used when the file was written
FILE_CASES = {
    "eri_file_122": dict(nsegments=2, pack_diagonal=True, drop={(2, 1): 8, (3, 3): 7}),
    "eri_file_113": dict(nsegments=1, pack_diagonal=True, drop={}),
}


def write_case_file(path, gdf, name, version="v1"):
    """the cderi file of a golden case, rewritten from the seeded provider (the fixtures store seeds, not tensors)"""
    from libdmet_preview_b200.gdf_file import write_gdf_file
    c = FILE_CASES[name]
    return write_gdf_file(str(path), gdf, version=version, nsegments=c["nsegments"], pack_diagonal=c["pack_diagonal"],
                          naux_of=c["drop"])
