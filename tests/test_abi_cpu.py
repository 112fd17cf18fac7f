"""The C-ABI boundary without a GPU: libvidchap.so loads, exports every entry point include/vidchap.h declares, the
ctypes binding covers each of them, and the product refuses to run (loudly, no CPU fallback) when no B200 is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vidchap.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def so_path():
    p = os.path.join(ROOT, "vidchapters_b200", "libvidchap.so")
    if not os.path.exists(p):
        import __graft_entry__ as g
        g.build()
    return p


def test_header_symbols_are_exported_and_bound(so_path):
    from vidchapters_b200 import lib as L
    syms = declared_symbols()
    assert len(syms) >= 29 and "vc_gemm_bf16" in syms and "vc_attn_bwd" in syms and "vc_beam_topk" in syms
    dll = ctypes.CDLL(so_path)
    for s in syms:
        assert hasattr(dll, s), f"{s} is declared in include/vidchap.h but not exported by libvidchap.so"
    bound = set(L.SIGNATURES) | {"vc_version", "vc_last_error", "vc_device_check"}
    missing = [s for s in syms if s not in bound]
    assert not missing, f"declared but not bound in vidchapters_b200/lib.py: {missing}"
    extra = [s for s in L.SIGNATURES if s not in syms]
    assert not extra, f"bound but not declared in the header: {extra}"
    loaded = L.load()
    assert loaded.vc_version() == L.ABI_VERSION
    # argument structs of the binding have the sizes the C side uses (a layout slip would corrupt every call)
    dll.vc_last_error.restype = ctypes.c_char_p
    assert ctypes.sizeof(L.AttnBwdArgs) > ctypes.sizeof(L.AttnArgs) > 100 and ctypes.sizeof(L.GemmArgs) > 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_gpu(so_path):
    from vidchapters_b200 import lib as L
    from vidchapters_b200.ops import CudaOps
    loaded = L.load()
    assert loaded.vc_device_check() != 0                      # status code + message, never abort()
    assert len(loaded.vc_last_error()) > 0
    with pytest.raises(RuntimeError):
        CudaOps()
    from vidchapters_b200 import TINY, Vid2Seq

    class Tok:
        pad_token_id, eos_token_id = 0, 1

        def __len__(self):
            return TINY["base_vocab"] + TINY["num_bins"]

    m = Vid2Seq("t5-base", num_features=10, tokenizer=Tok(), t5_config=dict(TINY, num_features=10))
    v = torch.randn(1, 10, 768)
    ids = torch.randint(2, 100, (1, 5))
    with pytest.raises(RuntimeError):                          # no CPU / PyTorch fallback of the forward
        m(v, {"input_ids": ids, "attention_mask": ids != 0}, {"input_ids": ids, "attention_mask": ids != 0})
