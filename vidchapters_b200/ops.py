"""CudaOps — the op table of the Vid2Seq hot path, each op one call into libvidchap.so (include/vidchap.h).

Tensors are torch CUDA tensors used only as device-memory handles (`data_ptr()`); the arithmetic happens in the
hand-written sm_100a kernels.  All ops launch on torch's current CUDA stream.  Nothing here has a CPU path: a
missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_RELU_BWD, ACT_GELU_BWD = 0, 1, 2, 3, 4


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vidchapters_b200 ops need CUDA tensors (no CPU fallback exists)")


class CudaOps:
    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("vidchapters_b200 needs a CUDA device (B200, sm_100a); none is available")
        _lib.check(self.lib.vc_device_check())
        self.launches = 0  # kernels launched through this table (bench.py reports it as gpu_launches)

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, out, *, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, pre_out=None,
             aux=None, alpha=1.0, splits=1, atomic=False, tile_n=0):
        """out[M,N] = epi(alpha * op(A) @ op(B)^T).  A: [M,K] (or [K,M] if a_mn); B: [N,K] (or [K,N] if b_mn)."""
        _chk_cuda(A, B, out, bias, residual, pre_out, aux)
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
        assert A.dim() == 2 and B.dim() == 2 and out.dim() == 2
        assert A.stride(1) == 1 and B.stride(1) == 1 and out.stride(1) == 1
        M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
        N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
        assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
        assert out.shape[0] == M and out.shape[1] == N, (out.shape, M, N)
        a = _lib.GemmArgs()
        a.A, a.B = A.data_ptr(), B.data_ptr()
        a.lda, a.ldb = A.stride(0), B.stride(0)
        a.M, a.N, a.K = M, N, K
        a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
        a.out, a.ldo = out.data_ptr(), out.stride(0)
        a.out_fp32 = int(out.dtype == torch.float32)
        assert out.dtype in (torch.float32, torch.bfloat16)
        a.atomic = int(atomic)
        a.bias = None if bias is None else bias.data_ptr()
        if residual is not None:
            assert residual.dtype == torch.float32 and residual.stride(1) == 1
            a.residual, a.ldr = residual.data_ptr(), residual.stride(0)
        a.act = act
        if pre_out is not None:
            assert pre_out.dtype == torch.bfloat16 and pre_out.stride(0) == out.stride(0)
            a.pre_out = pre_out.data_ptr()
        if aux is not None:
            assert aux.dtype == torch.bfloat16 and aux.stride(1) == 1
            a.aux, a.ld_aux = aux.data_ptr(), aux.stride(0)
        a.alpha = float(alpha)
        a.splits = int(splits)
        a.tile_n = int(tile_n)
        _lib.check(self.lib.vc_gemm_bf16(C.byref(a), self._stream()))
        self.launches += 1
        return out
