"""Host-side replay of the k-point schedule of `get_emb_eri_fast_gdf`.

The kernels never re-derive k-point logic: which (k_i, k_j) blocks belong to which transfer momentum k_L, which of
them the reference symmetrises (hermi_sum) and the time-reversal weights are decided here, following
libdmet/basis_transform/eri_transform.py:142-157 (weights) and :338-382 (visit order, k-conservation test,
partner lookup) exactly -- the schedule decides which blocks are symmetrised, so it must be identical.
"""
import numpy as np

KPT_DIFF_TOL = 1e-6   # pyscf.pbc.lib.kpts_helper.KPT_DIFF_TOL, the default of `kconserv_tol` (eri_transform.py:46)


def make_kpts_scaled(kmesh):
    """Scaled k-points in numpy.fft order, C-ordered product over the mesh axes (libdmet/system/fourier.py:46-53)."""
    axes = [np.fft.fftfreq(int(n), 1.0) for n in kmesh]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.ravel() for g in grids], axis=-1)


def cell_vectors(kmesh):
    """Integer cell positions, C-ordered (fourier.py:39-44, lattice.py:44-47)."""
    grids = np.meshgrid(*[np.arange(int(n)) for n in kmesh], indexing="ij")
    return np.stack([g.ravel() for g in grids], axis=-1)


def round_to_FBZ(kpts, tol=1e-10, wrap_around=True):
    """fourier.py:55-65."""
    k = np.asarray(kpts, dtype=float)
    k = k - np.floor(k)
    if wrap_around:
        k[k > (0.5 - tol)] -= 1.0
    else:
        k[k > (1.0 - tol)] = 0.0
    return k


def kpt_member(kpt, kpts, tol=KPT_DIFF_TOL):
    """Indices of `kpt` in `kpts` modulo reciprocal lattice vectors (fourier.py:73-81)."""
    kpts = np.reshape(kpts, (len(kpts), np.size(kpt)))
    dk = kpts - np.ravel(kpt)
    dk = np.linalg.norm(dk - np.round(dk), axis=-1)
    return np.where(dk < tol)[0]


def time_reversal_weights(kpts_scaled, tol=KPT_DIFF_TOL):
    """weights in {0, 1, 2}: a k-point paired with a later -k gets 2 and the partner 0 (eri_transform.py:142-157)."""
    kr = round_to_FBZ(kpts_scaled, tol=tol)
    nk = len(kr)
    w = np.ones(nk, dtype=int)
    for i in range(nk):
        if w[i] != 1:
            continue
        s = kr[i][None, :] + kr[i + 1:]
        s = s - np.round(s)
        hit = np.where(np.max(np.abs(s), axis=1) < tol)[0] if s.size else []
        if len(hit):
            w[i] = 2
            w[i + 1 + hit[0]] = 0
    assert w.sum() == nk
    return w


class EriSchedule(object):
    """units: list of (kL, weight, [(ki, kj, sym), ...]); weight is 1 or 2 with time reversal, 0 (meaning complex
    Lambda^dagger Lambda) without.  `sym` marks the blocks that get Lij + Lij^T."""

    def __init__(self, nkpts, t_reversal_symm, weights, units):
        self.nkpts = nkpts
        self.t_reversal_symm = t_reversal_symm
        self.weights = weights
        self.units = units

    @property
    def nblocks(self):
        return sum(len(u[2]) for u in self.units)

    @property
    def ngram(self):
        """number of real Gram products of stage 3 (1 per weight-1 unit, 2 otherwise)"""
        return sum(1 if u[1] == 1 else 2 for u in self.units)

    def unit_cost(self, flop_block, flop_gram):
        return [flop_block * len(u[2]) + flop_gram * (1 if u[1] == 1 else 2) for u in self.units]


_SCHEDULES = {}


def build_schedule(kpts_scaled, t_reversal_symm=True, kconserv_tol=KPT_DIFF_TOL, kscaled_center=None):
    """Replay of eri_transform.py:308-382.  `kpts_scaled` are the unshifted scaled k-points (what
    cell.get_scaled_kpts(mydf.kpts) returns); the optional centre shift only enters the conservation test and the
    partner lookup, as in the reference (l.266-268 vs l.309).  The schedule is a pure function of its arguments and
    a DMET loop asks for the same one every iteration, so the last few are kept (an EriSchedule is read-only)."""
    k0 = np.array(kpts_scaled, dtype=float)
    key = (k0.tobytes(), k0.shape, bool(t_reversal_symm), float(kconserv_tol),
           None if kscaled_center is None else np.asarray(kscaled_center, dtype=float).tobytes())
    hit = _SCHEDULES.get(key)
    if hit is not None:
        return hit
    sch = _build_schedule(k0, t_reversal_symm, kconserv_tol, kscaled_center)
    if len(_SCHEDULES) >= 8:
        _SCHEDULES.pop(next(iter(_SCHEDULES)))
    _SCHEDULES[key] = sch
    return sch


def _build_schedule(k0, t_reversal_symm, kconserv_tol, kscaled_center):
    nk = len(k0)
    ks = k0 - kscaled_center if kscaled_center is not None else k0
    if t_reversal_symm:
        weights = time_reversal_weights(k0)
    else:
        weights = np.ones(nk, dtype=int)
    # j(i, kL): the k_j with -k_i + k_j + k_L integer (l.349-351); a uniform mesh has exactly one per (i, kL)
    units = []
    minus = None
    if t_reversal_symm:
        minus = []
        for j in range(nk):
            jm = kpt_member(-ks[j], ks)
            assert len(jm) == 1
            minus.append(int(jm[0]))
    for kL in range(nk):
        if weights[kL] <= 0:
            continue
        visited = np.zeros(nk, dtype=bool)
        blocks = []
        for i in range(nk):
            if visited[i]:
                continue
            visited[i] = True
            kc = -ks[i][None, :] + ks + ks[kL][None, :]
            ok = np.max(np.abs(np.round(kc) - kc), axis=1) <= kconserv_tol
            for j in np.where(ok)[0]:
                j = int(j)
                if t_reversal_symm:
                    jm = minus[j]
                    blocks.append((i, j, 0 if visited[jm] else 1))
                    visited[jm] = True
                else:
                    blocks.append((i, j, 0))
        units.append((kL, int(weights[kL]) if t_reversal_symm else 0, blocks))
    return EriSchedule(nk, t_reversal_symm, weights, units)


def work_items(schedule, naux, nsplit=1, units=None):
    """Independent work items (unit index, l0, l1): a transfer momentum restricted to the auxiliary rows [l0, l1).
    Rows of the packed 3-index tensor are independent through stage 1 and enter `eri` additively in stage 3
    (eri += sum_L X[L,P] X[L,Q]), so items can run on different GPUs with no exchange but the final sum."""
    idx = range(len(schedule.units)) if units is None else units
    nsplit = max(1, min(int(nsplit), naux))
    cuts = [(naux * s) // nsplit for s in range(nsplit + 1)]
    return [(u, cuts[s], cuts[s + 1]) for u in idx for s in range(nsplit) if cuts[s + 1] > cuts[s]]


def choose_split(costs, nranks, max_split=4, tol=0.03, speeds=None):
    """smallest aux split whose LPT assignment is balanced to `tol` (1 when the units already divide evenly).
    `speeds`: relative throughput of the ranks (see assign_units)."""
    sp = [1.0] * nranks if speeds is None else [float(x) for x in speeds]
    best = 1
    for ns in range(1, max_split + 1):
        c = [x / ns for x in costs for _ in range(ns)]
        parts = assign_units(c, nranks, speeds)
        times = [sum(c[u] for u in p) / sp[r] for r, p in enumerate(parts)]
        if max(times) <= (1.0 + tol) * sum(c) / sum(sp):
            return ns
        best = ns
    return best


def assign_units(costs, nranks, speeds=None):
    """Longest-processing-time assignment of schedule units to ranks (the reference's MPI variant deals the units
    out round-robin, eri_transform_mpi.py:35-55; LPT balances unequal block counts better).  With `speeds` (relative
    throughput per rank, e.g. the host->device bandwidth each rank measured for a host-streamed build) a unit goes to
    the rank that would finish it first.  Returns a list of unit-index lists, deterministic."""
    sp = [1.0] * nranks if speeds is None else [float(x) for x in speeds]
    assert len(sp) == nranks and min(sp) > 0.0
    order = sorted(range(len(costs)), key=lambda u: (-costs[u], u))
    load = [0.0] * nranks
    out = [[] for _ in range(nranks)]
    for u in order:
        r = min(range(nranks), key=lambda x: ((load[x] + costs[u]) / sp[x], x))
        out[r].append(u)
        load[r] += costs[u]
    for r in range(nranks):
        out[r].sort()
    return out
