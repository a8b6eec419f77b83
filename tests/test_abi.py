"""CPU tier: the C-ABI library builds/loads and exports every symbol include/crs.h declares.
No compute call is made (there is no GPU here and no CPU implementation to call)."""
import os
import re

import pytest

from compressed_rag_suite_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "crs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = N.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in crs.h but not exported by libcrs.so"
    assert set(names) == set(N.SYMBOLS), "ctypes table and crs.h disagree"


def test_version_and_error_string():
    lib = N.lib()
    assert lib.crs_version() >= 100
    assert isinstance(lib.crs_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import ctypes as C
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = N.lib().crs_index_create(C.byref(h), 384, N.CRS_F16, N.CRS_COSINE, 0, 0, 0)
    assert rc == N.CRS_ECUDA
    with pytest.raises(RuntimeError):
        N.check(rc)
    from compressed_rag_suite_b200.rag import VectorStore, Chunk
    import numpy as np
    vs = VectorStore({"collection_name": "nogpu"})
    with pytest.raises(RuntimeError):          # the product fails loudly, it does not fall back
        vs.create_index([Chunk("t", "chunk_0", 0, 1)], np.ones((1, 8), dtype=np.float32))


def test_bad_arguments_are_valueerrors():
    import ctypes as C
    h = C.c_void_p()
    lib = N.lib()
    assert lib.crs_index_create(C.byref(h), 0, N.CRS_F16, N.CRS_COSINE, 0, 0, 0) == N.CRS_EINVAL
    assert lib.crs_index_create(C.byref(h), 384, N.CRS_F32, N.CRS_COSINE, 0, 0, 0) == N.CRS_EINVAL
    assert lib.crs_index_create(C.byref(h), 5000, N.CRS_F16, N.CRS_COSINE, 0, 0, 0) == N.CRS_EINVAL
    with pytest.raises(ValueError):
        N.check(N.CRS_EINVAL)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "compressed_rag_suite_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def _build_c_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(str(tmp_path), "crs_example")
    libdir = os.path.join(root, "compressed_rag_suite_b200")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                          os.path.join(root, "examples", "crs_example.c"), "-o", exe, "-L" + libdir, "-lcrs",
                          "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_is_plain_c_and_a_c_host_links_and_fails_loudly_without_gpu(tmp_path):
    """include/crs.h compiles as strict C99 from a C caller; with no GPU the program reports
    CRS_ECUDA from crs_index_create instead of computing anything on the CPU."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    run = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout + run.stderr
    else:
        assert run.returncode == 2 and "no CPU implementation" in run.stderr
