"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): `get_emb_eri(..., use_mpi=True)` -- the reference's own
keyword for its multi-process path (eri_transform.py:71) -- equals the oracle, as t_eri_transform_gdf_mpi.py:37-41
asserts for the MPI variant (< 1e-10)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from helpers import problem
from libdmet_preview_b200 import eri_transform as et
from oracle import eri_transform as oe
gdf, C, basis = problem([2, 2, 1], 9, 26, 8, spin=2)
for kw in (dict(), dict(nsplit=3), dict(t_reversal_symm=False, symmetry=1), dict(source="host")):
    got = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, use_mpi=True, **kw)
    if rank == 0:
        okw = {k: v for k, v in kw.items() if k not in ("nsplit", "source")}
        ref = oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, **okw)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-10, (kw, np.abs(got - ref).max())
    else:
        assert got is None
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("DIST_OK")
'''


def test_two_rank_nccl_equals_oracle(dev, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
