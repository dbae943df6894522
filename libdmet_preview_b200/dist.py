"""Multi-GPU embedding-ERI build: one process per GPU, transfer momenta k_L sharded over the ranks, one sum-reduce
of the partial s4 ERI at the end.

This extends the reference's own multi-process design (libdmet/basis_transform/eri_transform_mpi.py:35-55 `assign_workload`,
:151-157 per-rank k_L loop, :203-210 `mpi.reduce_inplace`, :212-223 rank 0 finishes with `eri_restore`) with NCCL over
NVLink in place of MPI: every k_L contributes an additive term to `eri`, so there is no data-path exchange during
the computation and a single collective at the end.  The units are dealt out by longest-processing-time on the
algorithmic flop count of each k_L (block count differs between k_L and weight-2 momenta need two Gram products).

`compute_partial` is injectable so that the sharding / reduction logic can be exercised on CPU (gloo, world size 2)
in tests with a stand-in for the CUDA pipeline; the product path always uses `eri_transform.emb_eri_device`.
"""
import numpy as np
import torch
import torch.distributed as dist

from .schedule import build_schedule, assign_units, work_items, choose_split, KPT_DIFF_TOL


def unit_costs(schedule, nao, naux, nemb, nspin):
    npair = nemb * (nemb + 1) // 2
    f_block = 8.0 * naux * nao * nemb * (nao + nemb) * nspin
    f_gram = float(naux) * npair * (npair + 1) * (1 if nspin == 1 else 4)
    return schedule.unit_cost(f_block, f_gram)


def rank_items(schedule, nao, naux, nemb, nspin, world_size, nsplit=None, speeds=None):
    """work items (unit, l0, l1) per rank -- deterministic, identical on every rank.  When the transfer momenta do
    not divide evenly over the ranks (36 units on 8 GPUs at 4x4x4 would cap the speed-up at 7.2x) each unit is
    split along the auxiliary index into `nsplit` independent, additive pieces.  `speeds`: relative throughput of
    the ranks (identical list on every rank), used when the build is bound by each rank's own host link."""
    costs = unit_costs(schedule, nao, naux, nemb, nspin)
    if nsplit is None:
        nsplit = choose_split(costs, world_size, speeds=speeds, max_split=4 if speeds is None else 8)
    items = work_items(schedule, naux, nsplit)
    icost = [costs[u] * (l1 - l0) / float(naux) for (u, l0, l1) in items]
    parts = assign_units(icost, world_size, speeds)
    return [[items[i] for i in p] for p in parts]


_H2D_SPEEDS = {}


def measure_h2d_speeds(group=None, nbytes=256 << 20, reps=3):
    """Pinned host -> device bandwidth of every rank with ALL ranks copying at the same time, in GB/s (identical
    list on every rank).  On a multi-GPU box the ranks' links are not equal (the 8-GPU boxes of this pool deliver
    23-35 GB/s per rank when all copy at once, profiles/probe_h2d_r02_8gpu.json) and a build that streams the GDF
    tensor from host memory finishes when the slowest link has moved its share -- so the shares are made
    proportional to these rates."""
    world = dist.get_world_size(group)
    key = (id(group), world)
    if key in _H2D_SPEEDS:                       # the links do not change during the life of the process group
        return _H2D_SPEEDS[key]
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    src.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier(group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    e1.synchronize()
    mine = torch.tensor([reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine, group=group)
    _H2D_SPEEDS[key] = [float(x.item()) for x in allr]
    return _H2D_SPEEDS[key]


def sharded_partial(schedule, shape, compute_partial, group=None, all_ranks=False, nsplit=None, speeds=None):
    """Each rank computes its items with `compute_partial(items) -> tensor`, then the partials are summed onto
    rank 0 (or all ranks).  shape = (nao, naux, nemb, nspin).  Returns the reduced tensor on rank 0 (all ranks if
    all_ranks) and the local partial elsewhere."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nao, naux, nemb, nspin = shape
    mine = rank_items(schedule, nao, naux, nemb, nspin, world, nsplit, speeds)[rank]
    part = compute_partial(mine)
    if world > 1:
        if all_ranks:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(part, dst=0, op=dist.ReduceOp.SUM, group=group)
    return part


def _blocks_per_launch(kwargs, et):
    """one name for the blocks-per-launch option on both routes: `group` as in the serial call when it is a number
    (`group_blocks` is kept as an alias); a ProcessGroup passed as `group` still selects the communicator"""
    g = kwargs.get("group", None)
    blocks = kwargs.get("group_blocks", None)
    pg = None
    if g is not None and not isinstance(g, (int, np.integer)):
        pg = g
    elif g is not None and blocks is None:
        blocks = int(g)
    if "process_group" in kwargs:
        pg = kwargs["process_group"]
    return (et.DEFAULT_GROUP if blocks is None else blocks), pg


def get_emb_eri_sharded(cell, mydf, C_ao_lo=None, basis=None, kscaled_center=None, symmetry=4,
                        kconserv_tol=KPT_DIFF_TOL, unit_eri=False, t_reversal_symm=True, C_ao_eo=None,
                        return_device=False, all_ranks=False, gso=False, basis_k=None, incore=True, fout="H2.h5",
                        feri=None, max_memory=None, swap_idx=None, **kwargs):
    """`get_emb_eri(..., use_mpi=True)` / `get_emb_eri_gso(..., use_mpi=True)`: same result layout as the serial
    call on rank 0 (None elsewhere unless all_ranks).  Restricted / unrestricted (eri_transform_mpi.py:57-223) and
    GSO (:226-388: same k_L sharding, one ERI block from Lambda_a - Lambda_b).  Without an initialised
    torch.distributed group the call is the serial build (one rank owns every k_L)."""
    from . import eri_transform as et
    from .device import get_device
    if not incore and not t_reversal_symm:
        raise NotImplementedError                                   # eri_transform_mpi.py:146
    if gso and not incore:
        raise NotImplementedError("GSO outcore ERI is not built")
    blocks, pg = _blocks_per_launch(kwargs, et)
    if not dist.is_initialized():
        kw = {k: v for k, v in kwargs.items() if k not in ("group", "group_blocks", "process_group")}
        if blocks is not None:
            kw["group"] = blocks
        common = dict(C_ao_lo=C_ao_lo, basis=basis, feri=feri, kscaled_center=kscaled_center, symmetry=symmetry,
                      max_memory=max_memory, kconserv_tol=kconserv_tol, unit_eri=unit_eri, swap_idx=swap_idx,
                      t_reversal_symm=t_reversal_symm, incore=incore, fout=fout, return_device=return_device, **kw)
        if gso:
            return et.get_emb_eri_gso(cell, mydf, basis_k=basis_k, **common)
        return et.get_emb_eri_fast_gdf(cell, mydf, C_ao_eo=C_ao_eo, **common)
    provider = et.as_provider(cell, mydf, feri=feri)
    if gso:
        CT = et.build_CT_gso(provider, C_ao_lo, basis, basis_k, unit_eri)
    else:
        CT = et.build_CT(provider, C_ao_lo, basis, C_ao_eo, unit_eri)
    nspin, nkpts, nemb, nao = CT.shape
    schedule = build_schedule(provider.kpts_scaled, t_reversal_symm, kconserv_tol, kscaled_center)
    imag = et._imag_buffer(t_reversal_symm, kwargs, 1 if gso else nspin * (nspin + 1) // 2, nemb)

    def compute(items):
        return et.emb_eri_device(provider, CT, schedule=schedule, items=items,
                                 source=kwargs.get("source", "auto"), group=blocks,
                                 kl_group=kwargs.get("kl_group", et.DEFAULT_KL_GROUP), stats=kwargs.get("stats", None),
                                 gso=gso, imag=imag)

    # a build that streams its blocks from host memory is bound by each rank's own host link: shares proportional to
    # the measured rates (device-generated and resident tensors: equal shares)
    speeds = kwargs.get("rank_speeds", None)
    host_fed = kwargs.get("source", "auto") == "host" or not (
        isinstance(provider, et.ResidentGDF) or (hasattr(provider, "keys") and hasattr(provider, "scale")))
    if speeds is None and host_fed and kwargs.get("balance_h2d", True) and dist.get_world_size(pg) > 1 \
            and torch.cuda.is_available() and dist.get_backend(pg) == "nccl":
        speeds = measure_h2d_speeds(pg)
    if isinstance(kwargs.get("stats", None), dict):
        kwargs["stats"]["rank_speeds"] = speeds
    eri = sharded_partial(schedule, (nao, provider.naux, nemb, nspin), compute, group=pg, all_ranks=all_ranks,
                          nsplit=kwargs.get("nsplit", None), speeds=speeds)
    if imag is not None and dist.get_world_size(pg) > 1:
        # rank 0 reports max|Im| of the complete Lambda^dagger Lambda (eri_transform_mpi.py:205-209)
        if all_ranks:
            dist.all_reduce(imag, op=dist.ReduceOp.SUM, group=pg)
        else:
            dist.reduce(imag, dst=0, op=dist.ReduceOp.SUM, group=pg)
    if dist.get_rank(pg) != 0 and not all_ranks:
        return None
    et._report_imag(imag, kwargs)
    nsp = 1 if gso else nspin
    if not incore:
        return et.write_outcore(eri, nemb, nsp, fout)
    eri = et.finalize_eri(eri, nemb, symmetry, nsp)
    return eri if return_device else get_device().to_host(eri)
