"""The C-ABI library loads on a machine without a GPU and exports every symbol include/dsgd.h declares; the
ctypes binding covers the same set; without a GPU the library fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dsgd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsgd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from distributed_sgd_b200 import native
    lib = C.CDLL(native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dsgd.h but not exported by libdsgd.so"


def test_ctypes_binding_covers_the_header():
    from distributed_sgd_b200 import native
    assert sorted(native.ABI) == declared_symbols()
    native.lib()                                   # resolves every bound symbol with its argtypes


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from distributed_sgd_b200 import native
    with pytest.raises(native.DsgdError) as e:
        native.NativeCtx(0, 16, 0.1)
    assert e.value.code == native.ERR_CUDA and "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distributed_sgd_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("# oracle", "") or f == "README", \
                    f"{f} mentions the oracle"
