"""CPU: the shared library builds with nvcc (no GPU needed), loads, and exports every symbol include/ldm_b200.h
declares; the ctypes table binds exactly that set.  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ldm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ldm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header(built_lib):
    from libdmet_preview_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.ldm_version() >= 100
    assert isinstance(lib.ldm_last_error(), bytes)


def test_no_gpu_means_loud_failure(built_lib):
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libdmet_preview_b200 import device, eri_transform, synthetic
    with pytest.raises(RuntimeError):
        device.Device()
    gdf = synthetic.SyntheticGDF([1, 1, 2], 3, 4)
    with pytest.raises(RuntimeError):
        eri_transform.get_emb_eri(gdf.cell, gdf, C_ao_lo=synthetic.make_C_ao_lo([1, 1, 2], 3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "libdmet_preview_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_header_is_plain_c_and_links_from_c(built_lib, tmp_path):
    """the boundary is a C ABI: include/ldm_b200.h compiles as C99 (no C++ or torch types in the signatures) and a C
    client linked against the library can call it (version / error string only -- no GPU here)"""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text('#include <stdio.h>\n#include "ldm_b200.h"\n'
                   'int main(void) { printf("%d %s\\n", ldm_version(), ldm_last_error() ? "ok" : "null"); return 0; }\n')
    exe = str(tmp_path / "client")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", exe, "-L", libdir, "-l:libldm_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) >= 100 and out[1] == "ok"


def test_unsupported_branches_are_refused_before_any_work():
    """`patch.install()` hands branches this package does not mirror to the reference; they must be refused by an
    argument check (no device, no ERI build) with the dedicated exception -- a plain NotImplementedError from torch
    or CUDA must NOT trigger that fallback"""
    import types
    import numpy as np
    from libdmet_preview_b200 import slater, spinless, patch
    from libdmet_preview_b200._lib import UnsupportedBranch
    assert issubclass(UnsupportedBranch, NotImplementedError)
    lat = types.SimpleNamespace(is_model=False)
    basis = np.zeros((1, 2, 3, 4))
    for kw in (dict(int_bath=False), dict(dft=True), dict(qsgw=True)):
        try:
            slater.get_emb_Ham(lat, basis, None, **kw)
            raise AssertionError("not refused: %s" % kw)
        except UnsupportedBranch:
            pass
    try:
        slater.get_emb_Ham(types.SimpleNamespace(is_model=True), basis, None)
        raise AssertionError("model Hamiltonian not refused")
    except UnsupportedBranch:
        pass
    try:
        spinless.get_emb_Ham(lat, basis[0], None, 0.0, int_bath=False)
        raise AssertionError("GSO non-interacting bath not refused")
    except UnsupportedBranch:
        pass
    calls = []

    def ours(x):
        if x == 0:
            raise UnsupportedBranch("not mirrored")
        raise NotImplementedError("a real failure")

    f = patch._with_fallback(ours, lambda x: calls.append(x) or "reference")
    assert f(0) == "reference" and calls == [0]
    try:
        f(1)
        raise AssertionError("a plain NotImplementedError must propagate")
    except UnsupportedBranch:
        raise AssertionError("wrong exception type")
    except NotImplementedError:
        pass
    assert calls == [0]
