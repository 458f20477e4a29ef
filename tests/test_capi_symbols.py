"""The C-ABI libraries load on a CPU-only box and export every symbol include/*.h declares.
No compute entry point is called here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DECL = re.compile(r"^\s*(?:const\s+)?(?:int|int64_t|void|double|const char\*|char\*|qg_index\*)\s*\*?\s*(q[gh]_[a-z0-9_]+)\s*\(", re.M)


def declared(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(DECL.findall(src)))


def test_gpu_library_exports_every_declared_symbol(capi):
    names = declared("quiver_gpu.h")
    assert len(names) >= 28
    lib = capi.load()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTED_SYMBOLS) == names, set(capi.EXPORTED_SYMBOLS) ^ set(names)
    assert lib.qg_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    header = os.path.join(ROOT, "include", "quiver_host.h")
    if not os.path.exists(header):
        pytest.skip("host layer not built in this tree")
    names = declared("quiver_host.h")
    assert names
    lib = ctypes.CDLL(os.path.join(ROOT, "quiver_b200", "lib", "libquiverhost.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_a_device(capi):
    """On a box without a GPU every compute entry point fails loudly with QG_ERR_CUDA."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.QuiverGpuError) as e:
        capi.Index(8, capi.L2)
    assert e.value.code == capi.QG_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under quiver_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quiver_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn), errors="replace") as f:
                    txt = f.read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/", txt, re.M):
                    # comments that cite oracle/synth.h as the generator's twin are fine
                    hits = [ln for ln in txt.splitlines()
                            if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|#include.*oracle", ln)]
                    if hits:
                        bad.append((fn, hits[:2]))
    assert not bad, bad
