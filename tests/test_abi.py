"""CPU: the C-ABI library loads, exports every symbol include/sfq_b200.h declares, and refuses to
work without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sfq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfq_[a-z_0-9]+)\s*\(", src)))


def test_header_and_python_mirror_agree():
    from slimfastq_b200 import api

    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    from slimfastq_b200 import api

    L = api.load_library()
    for name in declared_symbols():
        assert getattr(L, name) is not None
    assert L.sfq_version() == b"2.04/6 b200"


def test_no_cpu_fallback():
    import torch

    import slimfastq_b200 as S

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.SfqError):
        S.Codec()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "slimfastq_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) :
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "sfq_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_container_struct_sizes_match_python():
    from slimfastq_b200 import container as K

    assert K.FILE_HDR.size == 80 and K.BLOB_HDR.size == 132
