"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/neutts_b200.h
declares; size queries and argument validation work; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "neutts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from neutts_air_b200 import _lib, build

    build.build()
    L = _lib.lib()
    names = _declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/neutts_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names            # the ctypes binding covers exactly the header
    assert L.nt_abi_version() == 3


def test_workspace_queries_and_validation():
    from neutts_air_b200 import _lib

    L = _lib.lib()
    cfg = _lib.LMConfig(217472, 896, 4864, 24, 14, 2, 64, 1e-6, 1e6, 1, 2048, 64, 32, 2048)
    n = L.nt_lm_workspace_bytes(C.byref(cfg))
    assert 10e6 < n < 200e6
    bad = _lib.LMConfig(217472, 896, 4864, 24, 14, 2, 128, 1e-6, 1e6, 1, 2048, 64, 32, 2048)
    assert L.nt_lm_workspace_bytes(C.byref(bad)) == 0 and b"head_dim" in L.nt_last_error()
    cc = _lib.CodecConfig(1024, 12, 16, 64, 4096, 32, 7, 1920, 480, 4, 8, 1e-6, 1e4, 1e2, 1, 1, 256)
    assert L.nt_codec_workspace_bytes(C.byref(cc)) > 1e6
    a = _lib.GemmArgs()
    assert L.nt_gemm(C.byref(a), None) == -1                         # NT_ERR_INVALID: empty problem
    with pytest.raises(ValueError):
        _lib.check(-1)
    with pytest.raises(RuntimeError):
        _lib.check(-2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_engines_fail_loudly_without_gpu():
    from neutts_air_b200.codec import CodecDecoder, CodecShape
    from neutts_air_b200.lm import LMShape, SpeechLM

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SpeechLM(LMShape(), {}, device="cuda")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CodecDecoder(CodecShape(), {}, device="cuda")


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the shipped packages may import it."""
    for pkg in ("neutts_air_b200", "neutts", "neuttsair"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
