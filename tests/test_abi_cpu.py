"""The C-ABI library loads without a GPU and exports every symbol include/theatergen.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "theatergen.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__  # noqa: F401  (build() makes sure the .so exists)
    __graft_entry__.build()
    from theatergen_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in theatergen.h but not exported"
    assert set(_lib.exported_symbols()) == set(names), set(_lib.exported_symbols()) ^ set(names)


def test_no_compute_without_gpu_is_an_error_not_a_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from theatergen_b200 import _lib
    l = _lib.lib()
    assert l.tg_version() >= 100
    rc = l.tg_gemm(256, 64, 64, None, 0, 256, 128, 64, 64, None, None, 1, 0, None, 0, 0, 1.0, 256, 64, None)
    assert rc != 0 and len(l.tg_last_error()) > 0   # fails loudly: no device, no CPU path
    from theatergen_b200.unet import TGUNet  # noqa: F401
    assert "oracle" not in open(os.path.join(ROOT, "theatergen_b200", "unet.py")).read().replace("oracle/unet.py", "")


def test_plain_c_host_loads_the_library(tmp_path):
    """A C99 host (tests/c_host/abi_check.c, no CUDA headers, no Python): include/theatergen.h compiles as C with -pedantic, the
    library loads with dlopen, every declared entry point resolves and a call with null operands is refused with a message."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    import __graft_entry__
    __graft_entry__.build()
    from theatergen_b200 import _lib
    exe = str(tmp_path / "abi_check")
    src = os.path.join(ROOT, "tests", "c_host", "abi_check.c")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                        "-ldl", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, _lib.LIB_PATH, *_declared()], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout, r.stderr)
