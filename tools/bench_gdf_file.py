"""Disk-backed get_emb_eri: write a synthetic GDF tensor as a PySCF-layout cderi file, then time the build with the
blocks (a) in host memory, (b) read from the file through gdf_file.GDFFile (page cache warm), (c) resident in HBM
after the first pass (ResidentGDF over the file).  Prints one JSON line.

    python tools/bench_gdf_file.py --kmesh 2 2 2 --nao 64 --naux 300 --neo 48 [--dir /tmp]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmesh", type=int, nargs=3, default=[2, 2, 2])
    ap.add_argument("--nao", type=int, default=64)
    ap.add_argument("--naux", type=int, default=300)
    ap.add_argument("--neo", type=int, default=48)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from libdmet_preview_b200 import synthetic, eri_transform as et
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file

    gdf = synthetic.SyntheticGDF(args.kmesh, args.nao, args.naux, seed=1)
    C = synthetic.make_C_ao_lo(args.kmesh, args.nao, seed=2)
    basis = synthetic.make_emb_basis(args.kmesh, args.nao, args.neo, seed=3)
    path = os.path.join(args.dir, "ldm_bench_cderi.h5")
    t0 = time.perf_counter()
    write_gdf_file(path, gdf)
    t_write = time.perf_counter() - t0
    nk = len(gdf.kpts_scaled)

    class HostGDF(object):            # every block already in host memory (what the e2e leg of bench.py streams)
        def __init__(self):
            for a in ("kpts_scaled", "kmesh", "nao", "naux", "cell", "kpts"):
                setattr(self, a, getattr(gdf, a))
            self.blocks = {(i, j): gdf.load(i, j) for i in range(nk) for j in range(nk)}

        def load(self, ki, kj):
            return self.blocks[(ki, kj)]

    def timed(provider, **kw):
        best, out, st = 1e30, None, {}
        for _ in range(args.reps):
            st = {}
            torch.cuda.synchronize()
            t = time.perf_counter()
            out = et.get_emb_eri(gdf.cell, provider, C_ao_lo=C, basis=basis, stats=st, **kw)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        return best, out, st

    host = HostGDF()
    t_host, e_host, st_host = timed(host)
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    t_file, e_file, st_file = timed(f)                 # stored entries shipped as they are, unpacked on the device
    et.DEVICE_UNPACK = False
    t_asm, e_asm, st_asm = timed(f)                    # blocks assembled with numpy on the host (the reference's way)
    et.DEVICE_UNPACK = True
    res = et.ResidentGDF(f)
    et.get_emb_eri(gdf.cell, res, C_ao_lo=C, basis=basis)
    t_res, e_res, st_res = timed(res)
    print(json.dumps({
        "tool": "bench_gdf_file", "kmesh": args.kmesh, "nao": args.nao, "naux": args.naux, "neo": args.neo,
        "file_bytes": os.path.getsize(path), "write_s": round(t_write, 3),
        "h2d_bytes_per_build": st_host.get("h2d_bytes"),
        "host_memory_s": round(t_host, 4), "file_s": round(t_file, 4), "file_host_assembly_s": round(t_asm, 4),
        "resident_s": round(t_res, 4), "file_h2d_bytes": st_file.get("h2d_bytes"),
        "file_h2d_gbs": round(st_file.get("h2d_bytes", 0) / t_file / 1e9, 2),
        "file_block_gbs": round(st_host.get("h2d_bytes", 0) / t_file / 1e9, 2),
        "host_assembly_equals_bitwise": bool(np.array_equal(e_asm, e_host)),
        "host_h2d_gbs": round(st_host.get("h2d_bytes", 0) / t_host / 1e9, 2),
        "file_equals_host_bitwise": bool(np.array_equal(e_file, e_host)),
        "resident_max_abs_diff": float(np.abs(e_res - e_host).max()),
        "note": "file read through h5lite into the pinned staging ring (page cache warm after the write); file_s: "
                "stored entries cross PCIe as they are and are unpacked on the device (ldm_eri_block_stored); "
                "file_host_assembly_s: packed / swapped entries expanded with numpy first; file_block_gbs = expanded "
                "block bytes per second of the file path",
    }))
    os.remove(path)


if __name__ == "__main__":
    main()
