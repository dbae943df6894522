"""libdmet_preview_b200: sm_100a (B200) implementation of libDMET's embedding-Hamiltonian hot path behind the
reference's own Python signatures (eri_transform.get_emb_eri, fourier.k2R/R2k, make_basis.transform_h1_to_lo,
slater.get_emb_basis / embHam).  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
