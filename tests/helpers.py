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
