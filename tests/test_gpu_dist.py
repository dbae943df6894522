"""GPU, 2..8 ranks over NCCL (skipped on a single-GPU box): `get_emb_eri(..., use_mpi=True)` -- the reference's own
keyword for its multi-process path (eri_transform.py:71) -- equals the oracle, as t_eri_transform_gdf_mpi.py:37-41
asserts for the MPI variant (< 1e-10): unrestricted, auxiliary-split items, without time reversal, blocks staged
from host memory, the GSO build (eri_transform_mpi.py:226-388) and a restricted build with more units than ranks.
Run on hardware with `gpurun --gpus 2` / `--gpus 8`; logs under profiles/dist_nccl_r02_*.txt."""
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
from helpers import problem, gso_basis
from libdmet_preview_b200 import eri_transform as et
from oracle import eri_transform as oe
world = dist.get_world_size()
def check(tag, got, ref_fn):
    if rank == 0:
        ref = ref_fn()
        err = float(np.abs(got - ref).max())
        print("[dist x%%d] %%-44s max|gpu - oracle| = %%.2e" %% (world, tag, err), flush=True)
        assert got.shape == ref.shape and err < 1e-10, (tag, err)
    else:
        assert got is None
# unrestricted (aa, ab, bb), four ways of feeding / splitting the work
gdf, C, basis = problem([2, 2, 1], 9, 26, 8, spin=2)
for kw in (dict(), dict(nsplit=3), dict(t_reversal_symm=False, symmetry=1), dict(source="host")):
    got = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, use_mpi=True, **kw)
    okw = {k: v for k, v in kw.items() if k not in ("nsplit", "source")}
    check("unrestricted %%s" %% kw, got, lambda: oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, **okw))
# restricted, more transfer momenta than ranks, blocks per launch given with the serial keyword
gdf, C, basis = problem([2, 2, 4], 12, 40, 14)
got = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, use_mpi=True, group=2, kl_group=2)
check("restricted 2x2x4", got, lambda: oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis))
# every rank gets the sum (all_ranks) and it is the same bits everywhere
both = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, use_mpi=True, all_ranks=True, return_device=True)
digest = torch.stack([both.sum(), both.abs().max()])
lo, hi = digest.clone(), digest.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi)
if rank == 0:
    assert np.abs(both.cpu().numpy() - got).max() < 1e-12
# generalised spin orbitals (eri_transform_mpi.py:226-388)
gdf, C, _ = problem([2, 1, 2], 7, 20, 4, spin=2)
gb = gso_basis([2, 1, 2], 7, 9, seed=3)
for kw in (dict(), dict(t_reversal_symm=False), dict(nsplit=2, symmetry=1)):
    got = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=gb, use_mpi=True, **kw)
    okw = {k: v for k, v in kw.items() if k != "nsplit"}
    check("GSO %%s" %% kw, got, lambda: oe.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=gb, **okw))
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("DIST_OK")
'''


def test_multi_rank_nccl_equals_oracle(dev, tmp_path):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    nproc = min(ngpu, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
