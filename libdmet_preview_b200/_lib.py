"""ctypes binding of libldm_b200.so (C ABI declared in include/ldm_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a handle without an
sm_100 GPU, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldm_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/ldm_b200.h (checked by tests/test_abi.py)
SIGNATURES = {
    "ldm_version": (C.c_int, []),
    "ldm_last_error": (C.c_char_p, []),
    "ldm_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "ldm_destroy": (C.c_int, [vp]),
    "ldm_set_option": (C.c_int, [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "ldm_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "ldm_host_free": (C.c_int, [vp]),
    "ldm_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "ldm_dev_free": (C.c_int, [vp, vp]),
    "ldm_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "ldm_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "ldm_memset": (C.c_int, [vp, vp, C.c_int, C.c_size_t, vp]),
    "ldm_stream_sync": (C.c_int, [vp, vp]),
    "ldm_zgemm_tn": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               c_i32p, vp, c_i64p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int]),
    "ldm_dgemm_tn": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int64,
                               C.c_double, C.c_int, C.c_int]),
    "ldm_mirror_lower": (C.c_int, [vp, vp, vp, C.c_int, C.c_int64]),
    "ldm_phase_transform": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_double,
                                      C.c_int, C.c_int, c_f64p]),
    "ldm_lattice_dft": (C.c_int, [vp, vp, vp, vp, c_i32p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                  c_f64p]),
    "ldm_ztranspose": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "ldm_d2z": (C.c_int, [vp, vp, vp, vp, C.c_int64]),
    "ldm_ksum_real": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int64, C.c_double, c_f64p]),
    "ldm_pack_tril": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, c_f64p]),
    "ldm_restore_s1": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "ldm_restore_s8": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "ldm_jk_s4": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int]),
    "ldm_jk_s4_symm": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int]),
    "ldm_scale_eri": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp]),
    "ldm_synth_block": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_double]),
    "ldm_eri_begin": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]),
    "ldm_eri_set_store": (C.c_int, [vp, vp, C.c_int]),
    "ldm_eri_set_mode": (C.c_int, [vp, C.c_int]),
    "ldm_eri_set_imag": (C.c_int, [vp, vp]),
    "ldm_max_abs": (C.c_int, [vp, vp, vp, C.c_int64, c_f64p]),
    "ldm_eri_block_host": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    "ldm_eri_block_stored": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int64, C.c_int]),
    "ldm_unpack_stored": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "ldm_eri_block_store": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldm_eri_block_synth": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_double]),
    "ldm_eri_end_kl": (C.c_int, [vp, C.c_int]),
    "ldm_eri_finish": (C.c_int, [vp]),
    "ldm_eri_end": (C.c_int, [vp]),
    "ldm_release_workspaces": (C.c_int, [vp]),
    "ldm_eri_stats": (C.c_int, [vp, c_i64p, c_i64p]),
    "ldm_launch_count": (C.c_int64, [vp]),
    "ldm_eri_kernel_time": (C.c_int, [vp, C.c_int, c_f64p, c_i64p]),
}

_lib = None


def load():
    """Load libldm_b200.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libldm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` -- "
            "there is no CPU fallback for this package." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class UnsupportedBranch(NotImplementedError):
    """a branch of the reference this package does not mirror (model Hamiltonians, DFT / QSGW one-body terms, the
    non-interacting bath, ...).  `patch.install()` routes exactly these calls to the reference's own function; any
    other error -- including a NotImplementedError raised inside torch or CUDA -- propagates."""


class LdmError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().ldm_last_error()
        raise LdmError("libldm_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
