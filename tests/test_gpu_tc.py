"""tcgen05 building blocks on real hardware: the UMMA self-test kernel against a float64 matmul of
the rounded operands (all four operand-staging variants, both 16-bit formats, ragged K)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from mipnerf_pl_b200 import _cabi  # noqa: E402

DEV = "cuda:0"


def run_selftest(n, k, precision, variant, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(128, k, generator=g).to(DEV)
    b = (torch.randn(n, k, generator=g) * 0.2).to(DEV)
    d = torch.full((128, n), float("nan"), device=DEV)
    scratch = torch.zeros(((k + 63) // 64) * n * 128, dtype=torch.uint8, device=DEV)
    rc = _cabi.lib().mipnerf_b200_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k,
                                               _cabi.PRECISIONS[precision], variant, scratch.data_ptr(),
                                               scratch.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    dt = torch.bfloat16 if precision == "bf16" else torch.float16
    ref = (a.to(dt).double() @ b.to(dt).double().T)
    return d, ref


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n,k", [(256, 256), (256, 96), (128, 64), (128, 256)])
def test_umma_selftest(n, k, precision, variant):
    d, ref = run_selftest(n, k, precision, variant)
    err = float((d.double() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 2e-5 * max(scale, 1.0) * (k ** 0.5), f"max err {err} (scale {scale})"


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n", [128, 256])
def test_umma_selftest_sw64_tail(n, precision):
    """32-wide K slab in the 64-byte-swizzle layout (the tail of the 96-d IPE features)."""
    d, ref = run_selftest(n, 32, precision, 4)
    err = float((d.double() - ref).abs().max())
    assert err <= 2e-5 * max(float(ref.abs().max()), 1.0) * (32 ** 0.5), err


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n,k", [(256, 256), (128, 64), (256, 96 + 32), (128, 32)])
def test_umma_selftest_a_sw32_blocks(n, k, precision):
    """A as dense K = 16 blocks in the 32-byte-swizzle layout against B as 64-byte-swizzle K = 32 slabs: the operand
    layouts of the level kernels' activation tile and weight stages."""
    d, ref = run_selftest(n, k, precision, 8)
    err = float((d.double() - ref).abs().max())
    assert err <= 2e-5 * max(float(ref.abs().max()), 1.0) * (k ** 0.5), err
