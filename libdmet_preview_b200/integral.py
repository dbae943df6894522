"""`Integral` container returned by `embHam` (libdmet/system/integral.py:61-105) and `get_eri_format` (:883-928)."""
import numpy as np


class Integral(object):
    """norb, restricted, bogoliubov, H0, H1 = {"cd": (spin, norb, norb)}, H2 = {"ccdd": (spin_pair,) + s1 | s4 | s8
    layout}, ovlp -- the attribute set impurity solvers read from the reference's Integral."""

    def __init__(self, norb, restricted, bogoliubov, H0, H1, H2, ovlp=None):
        self.norb, self.restricted, self.bogoliubov, self.H0 = norb, restricted, bogoliubov, H0
        self.H1 = {"cd": H1} if isinstance(H1, np.ndarray) else H1
        self.H2 = {"ccdd": H2} if isinstance(H2, np.ndarray) else H2
        for key, v in self.H1.items():
            if v is not None and not (v.ndim == 3 and v.shape[-1] == norb):
                raise Exception("invalid shape %s, should have shape (spin, %s, %s)" % (str(v.shape), norb, norb))
        for key, v in self.H2.items():
            if v is not None and v.ndim not in (5, 3, 2):
                raise Exception("invalid H2 shape: %s" % str(v.shape))
        self.ovlp = np.eye(norb) if ovlp is None else ovlp


def get_eri_format(eri, nao):
    """('s1' | 's4' | 's8', spin_dim in {0, 1, 3}) deduced from rank and size"""
    eri = np.asarray(eri)
    npair = nao * (nao + 1) // 2
    size = {"s1": nao ** 4, "s4": npair * npair, "s8": npair * (npair + 1) // 2}
    by_rank = {5: "s1", 3: "s4"}
    if eri.ndim in by_rank:
        fmt = by_rank[eri.ndim]
        spin_dim, rem = divmod(eri.size, size[fmt])
        if rem:
            raise Exception("%s: eri.shape %s not consistent with nao %s" % (fmt, str(eri.shape), nao))
    elif eri.ndim == 4 and eri.size == size["s1"]:
        fmt, spin_dim = "s1", 0
    elif eri.ndim == 2 and eri.size == size["s4"]:
        fmt, spin_dim = "s4", 0
    elif eri.ndim == 2 and eri.size == size["s8"]:
        fmt, spin_dim = "s8", 1
    elif eri.ndim == 1 and eri.size == size["s8"]:
        fmt, spin_dim = "s8", 0
    else:
        raise ValueError("Unknown ERI shape %s, nao %s" % (str(eri.shape), nao))
    if spin_dim not in (0, 1, 3):
        raise Exception("spin_dim(%s) incorrect" % spin_dim)
    return fmt, spin_dim
