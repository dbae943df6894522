"""`Integral` container returned by `embHam` (libdmet/system/integral.py:61-105) and `get_eri_format` (:883-928)."""
import numpy as np


class Integral(object):
    def __init__(self, norb, restricted, bogoliubov, H0, H1, H2, ovlp=None):
        """H1: {"cd": (spin, norb, norb)};  H2: {"ccdd": (spin_pair,) + s1 | s4 | s8 layout}."""
        self.norb = norb
        self.restricted = restricted
        self.bogoliubov = bogoliubov
        self.H0 = H0
        if isinstance(H1, np.ndarray):
            H1 = {"cd": H1}
        if isinstance(H2, np.ndarray):
            H2 = {"ccdd": H2}
        for key in H1:
            if not (H1[key] is None or (H1[key].ndim == 3 and H1[key].shape[-1] == self.norb)):
                raise Exception("invalid shape %s, should have shape (spin, %s, %s)"
                                % (str(H1[key].shape), self.norb, self.norb))
        self.H1 = H1
        for key in H2:
            if H2[key] is not None and H2[key].ndim not in (5, 3, 2):
                raise Exception("invalid H2 shape: %s" % str(H2[key].shape))
        self.H2 = H2
        self.ovlp = np.eye(self.norb) if ovlp is None else ovlp


def get_eri_format(eri, nao):
    """('s1' | 's4' | 's8', spin_dim in {0, 1, 3})."""
    eri = np.asarray(eri)
    nao_pair = nao * (nao + 1) // 2
    s1_size = nao ** 4
    s4_size = nao_pair * nao_pair
    s8_size = nao_pair * (nao_pair + 1) // 2
    if eri.ndim == 5:
        fmt, spin_dim = 's1', eri.size // s1_size
        ok = spin_dim * s1_size == eri.size
    elif eri.ndim == 4 and eri.size == s1_size:
        fmt, spin_dim, ok = 's1', 0, True
    elif eri.ndim == 3:
        fmt, spin_dim = 's4', eri.size // s4_size
        ok = spin_dim * s4_size == eri.size
    elif eri.ndim == 2 and eri.size == s4_size:
        fmt, spin_dim, ok = 's4', 0, True
    elif eri.ndim == 2 and eri.size == s8_size:
        fmt, spin_dim, ok = 's8', 1, True
    elif eri.ndim == 1 and eri.size == s8_size:
        fmt, spin_dim, ok = 's8', 0, True
    else:
        raise ValueError("Unknown ERI shape %s, nao %s" % (str(eri.shape), nao))
    if not ok or spin_dim not in (0, 1, 3):
        raise Exception("spin_dim(%s) incorrect for ERI shape %s, nao %s" % (spin_dim, str(eri.shape), nao))
    return fmt, spin_dim
