"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

from helpers import ROOT


def declared_symbols():
    names = []
    for h in ("lvt_c.h", "lvt_kernels.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        names += re.findall(r"LVT_API\s+[\w\s\*]+?\b(lvtk?_\w+)\s*\(", txt)
    return sorted(set(names))


def test_headers_declare_the_reference_abi():
    names = declared_symbols()
    for ref in ("lvt_create", "lvt_destroy", "lvt_track", "lvt_track_with_external_corners", "lvt_get_status"):
        assert ref in names  # lvt/src/lvt_c.h:55-62
    assert len(names) >= 25


def test_headers_compile_as_c():
    src = "#include \"lvt_kernels.h\"\nint main(void){ lvt_params_c p; (void)p; return 0; }\n"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-fsyntax-only"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("which", ["oracle", "cuda"])
def test_library_exports_every_symbol(which, oracle):
    if which == "oracle":
        path = oracle.path
    else:
        import lvt_b200
        path = lvt_b200.LIB_PATH
        if not os.path.exists(path):
            import __graft_entry__
            __graft_entry__.build()
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_product_does_not_reference_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    for d, _, files in os.walk(os.path.join(ROOT, "lvt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "liblvt_oracle" not in txt and "lvto.h" not in txt and "oracle/_build" not in txt, f


def test_cuda_library_reports_gpu_and_fails_without_device():
    import lvt_b200
    import torch
    lib = lvt_b200.load()
    assert lib.is_gpu
    if not torch.cuda.is_available():
        from lvt_b200 import configs
        with pytest.raises(lvt_b200.LvtError):
            lib.create(configs.make_params("kitti_synth"))  # no silent CPU path


def test_cxx_lvt_system_mirror_over_the_c_abi(oracle, tmp_path):
    """include/lvt_system.hpp (the reference's lvt_system method names over the C ABI) compiles as C++11 and
    tracks through the oracle library, which exports the same symbols as the CUDA build"""
    import subprocess
    exe = str(tmp_path / "cxx_smoke")
    so_dir = os.path.dirname(oracle.path)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx_system_smoke.cpp"), "-o", exe, "-L", so_dir, "-llvt_oracle",
                    "-Wl,-rpath," + so_dir], check=True)
    r = subprocess.run([exe, "320", "200"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
