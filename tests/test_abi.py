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
