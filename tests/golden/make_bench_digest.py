"""Oracle digest of the bench workloads (test infrastructure; run once on the CPU, output committed).

    python tests/golden/make_bench_digest.py target [--procs 8]        ->  tests/golden/bench_digest.json

bench.py defines the GDF tensor of a workload by a pool of synthetic blocks (`synthetic.PooledGDF`) that is the same
for every number of GPUs.  The full embedding ERI of the target workload costs the oracle hours, but its restriction
to a few embedding orbitals is the same algorithm with a narrower C_ao_emb: here the ORACLE's `get_emb_eri_fast_gdf`
(same schedule, symmetrisation flags, weights, Gram products) is run with the `C_ao_eo` entry on the sampled
orbitals O, which gives eri[tri(a,b), tri(c,d)] for a, b, c, d in O exactly as the full run would.  bench.py compares
the same entries of the GPU result at every N with these numbers (`parity.max_abs_vs_oracle_digest`).
The transfer momenta are dealt out to worker processes through the oracle's `kL_subset` hook.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "bench_digest.json")


def setup(name, npool):
    import bench
    from libdmet_preview_b200 import synthetic
    from oracle import eri_transform as oe
    kmesh, nao, naux, neo, nspin = bench.workload(name)
    gdf = synthetic.PooledGDF(synthetic.SyntheticGDF(kmesh, nao, naux, seed=bench.GDF_SEED), npool)
    C_ao_lo, basis = bench.make_inputs(kmesh, nao, neo, nspin)
    orbs = bench.sample_orbitals(neo)
    C_ao_emb = oe.build_C_ao_emb(gdf, C_ao_lo, basis)                 # (spin, nk, nao, neo) / nk^0.75
    C_ao_eo = C_ao_emb[..., orbs] * (len(gdf.kpts_scaled) ** 0.75)
    return gdf, C_ao_eo, orbs


class Cached(object):
    """keeps the distinct pool blocks a worker meets (the small e2e pools fit in memory)"""

    def __init__(self, gdf, limit):
        self.g, self.limit, self.c = gdf, limit, {}
        for a in ("kmesh", "nao", "naux", "cell", "kpts", "kpts_scaled", "blockdim"):
            setattr(self, a, getattr(gdf, a))

    def load(self, ki, kj):
        s = self.g.pool_index(ki, kj)
        if s in self.c:
            return self.c[s]
        L = self.g.load(ki, kj)
        if len(self.c) < self.limit:
            self.c[s] = L
        return L


def work(args):
    name, npool, kLs = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import eri_transform as oe
    gdf, C_ao_eo, orbs = setup(name, npool)
    prov = Cached(gdf, 6)
    return oe.get_emb_eri_fast_gdf(gdf.cell, prov, C_ao_eo=C_ao_eo, kL_subset=set(kLs), restore=False)


def digest(name, npool, procs):
    from libdmet_preview_b200.schedule import build_schedule
    gdf, C_ao_eo, orbs = setup(name, npool)
    sch = build_schedule(gdf.kpts_scaled, True)
    kLs = [u[0] for u in sch.units]
    chunks = [kLs[i::procs] for i in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        parts = pool.map(work, [(name, npool, c) for c in chunks if c])
    eri = sum(parts)
    return {"npool": npool, "orbitals": [int(o) for o in orbs], "eri_s4_lower": eri.tolist(),
            "nblocks": sch.nblocks, "ngram": sch.ngram}


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--procs", type=int, default=8)
    a = ap.parse_args()
    have = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in a.workloads:
        kmesh, nao, naux, neo, nspin = bench.workload(name)
        have[name] = {"resident": digest(name, bench.pool_blocks(kmesh, nao, naux), a.procs),
                      "e2e": digest(name, bench.E2E_POOL, a.procs)}
        with open(OUT, "w") as f:
            json.dump(have, f, indent=0)
        print(name, "done", flush=True)


if __name__ == "__main__":
    main()
