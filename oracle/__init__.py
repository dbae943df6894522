"""CPU oracle for the embedding-Hamiltonian hot path of gkclab/libdmet_preview.

TEST INFRASTRUCTURE ONLY.  This package is a plain-numpy restatement of the reference's algorithm for
`get_emb_eri` (GDF) / `k2R` / `R2k` / `transform_h1_to_lo` / `get_emb_basis` / `embHam`.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and
only as the checker or the timed CPU baseline -- never as part of the product path in `libdmet_preview_b200/`.

Every function cites the reference file:line it follows (paths relative to /root/reference).

Pinning status (see DESIGN.md "Oracle"):
  * The arithmetic of the reference lives in PySCF (`pyscf>=2.0`, un-pinned, not vendored, not installed in
    this image) -- `oracle/pyscf_lib.py` restates the published semantics of the handful of PySCF helpers
    the path calls.
  * The reference holds no stored golden arrays for this path.  The oracle is pinned instead against the
    reference's OWN Python control flow: `tests/golden/make_golden.py` imports the unmodified modules from
    /root/reference over a stub of `pyscf`/`h5py` (whose numerical helpers are the same restatements) and
    stores the results of `libdmet.basis_transform.eri_transform.get_emb_eri`, `libdmet.system.fourier.k2R/R2k`,
    `make_basis.transform_h1_to_lo`, `slater.get_emb_basis` / `get_emb_Ham` as fixtures in `tests/golden/`.
    The one input fixture the reference ships for this path (`libdmet/routine/test/rdm1_lo`) is used as is.
"""
