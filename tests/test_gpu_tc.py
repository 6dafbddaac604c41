"""GPU: tcgen05 building blocks and the bf16 tensor-core field kernels against fp32 references."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (256, 48), (48, 256), (256, 128), (16, 16), (64, 112)])
def test_tcgen05_gemm_building_block(N, K):
    """smem descriptors / instruction descriptor / 128B swizzle / bulk-copy weight image / TMEM epilogue."""
    from spurfies_b200 import _lib
    from spurfies_b200.packing import pack_sw128
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).cuda().to(torch.bfloat16).contiguous()
    W = torch.randn(N, K, generator=g).cuda()
    Wp = pack_sw128(W)
    out = torch.zeros(128, N, device="cuda")
    _lib.call("spf_tc_gemm_test", _lib.ptr(A), _lib.ptr(Wp), N, K, _lib.ptr(out), _lib.stream())
    torch.cuda.synchronize()
    ref = A.float() @ W.to(torch.bfloat16).float().t()
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err   # bf16 products are exact in fp32; only the accumulation order differs
