"""The C-ABI library loads on a CPU-only machine, exports every symbol include/tealeaf_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from tl_testutil import has_gpu


def declared_symbols():
    from exploringsycl_b200._lib import HEADER_PATH
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    from exploringsycl_b200._lib import SIGNATURES, lib
    names = declared_symbols()
    assert len(names) >= 55
    L = lib()
    for n in names:
        assert hasattr(L, n), "library does not export %s" % n
        assert n in SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(SIGNATURES) == names


def test_version_and_device_count():
    from exploringsycl_b200 import lib
    assert b"sm_100a" in lib().tl_version()
    assert lib().tl_device_count() >= 0


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from exploringsycl_b200 import Chunk, TeaLeafError
    with pytest.raises(TeaLeafError) as e:
        Chunk(10, 10)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sub in ("exploringsycl_b200", "include", "c_kernels"):
        for dp, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".c")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "tealeaf_oracle" not in txt and "import oracle" not in txt \
                        and "from oracle" not in txt, os.path.join(dp, f)
