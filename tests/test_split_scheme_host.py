"""CPU check of the host-side operand preparation of the mixed fp16 + 2x e4m3 split: the float64
emulation of what `tc_gemm_bf16x3<.., MIX=1>` accumulates stays inside the parity budget, the fp16 pass alone does not,
and the byte tensor has the layout the kernel's 8-bit tensor map expects."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_tensor import mixed_operands

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("scale_b", [0.3, 0.01, 20.0])
def test_mixed_split_emulation_is_inside_the_budget(scale_b):
    rng = np.random.default_rng(7)
    A = rng.random((96, 256)).astype(np.float32)                     # U[0,1) like the data operand
    B = (rng.standard_normal((64, 256)) * scale_b).astype(np.float32)
    a16, a8h, a8l, a8, sa = mixed_operands(A)
    b16, b8h, b8l, b8, sb = mixed_operands(B)
    for x16, h8, l8 in ((a16, a8h, a8l), (b16, b8h, b8l)):
        assert 128 < np.abs(x16.astype(np.float64)).max() <= 256      # one power-of-two scale puts the maximum here
        assert np.abs(h8).max() <= 256 and np.abs(l8).max() <= 256    # inside e4m3's range (max 448)
    a = a16.astype(np.float64) * 32.0
    b = b16.astype(np.float64) * 64.0
    assert np.abs(a).max() < 65504 and np.abs(b).max() < 65504        # the fp16 operands with their share of 2^11
    c = 2.0 ** -11 / (sa * sb)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    full = c * (a @ b.T + a8l @ b8h.T + a8h @ b8l.T)
    single = c * (a @ b.T)
    assert rel_err(full, want) < 4e-5
    assert rel_err(single, want) > 1e-4


def test_mixed_split_byte_layout():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((5, 64)).astype(np.float32)
    x16, h8, l8, packed, s = mixed_operands(X)
    assert packed.shape == (5, 128) and packed.dtype == np.uint8
    dec = lambda b: torch.from_numpy(np.ascontiguousarray(b)).view(torch.float8_e4m3fn).to(torch.float64).numpy()
    for chunk in range(2):                                           # per 32 K-elements: 32 bytes hi, then 32 bytes lo
        np.testing.assert_array_equal(dec(packed[:, chunk * 64:chunk * 64 + 32]), h8[:, chunk * 32:chunk * 32 + 32])
        np.testing.assert_array_equal(dec(packed[:, chunk * 64 + 32:chunk * 64 + 64]), l8[:, chunk * 32:chunk * 32 + 32])
    np.testing.assert_allclose(h8 / s, X, rtol=2.0 ** -4)             # e4m3: 4 significant bits
