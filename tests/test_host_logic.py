"""Host-side logic that runs without a GPU: the stacked AdaLN modulation table."""
import torch

from fastdm_b200.blocks import AdaLNTable, ModChunk, _f32, _mod


class _Lin:
    """Stand-in for an unquantised QLinear: weight is the [K, N] view the blocks use, bias [N]."""

    def __init__(self, n, k, g):
        self.weight = (torch.randn(n, k, generator=g) * 0.1).to(torch.bfloat16).t()
        self.bias = torch.randn(n, generator=g).to(torch.bfloat16)

    def forward(self, x):
        return torch.addmm(self.bias, x, self.weight)


def test_adaln_table_equals_separate_linears():
    g = torch.Generator().manual_seed(0)
    k = 64
    lins = [_Lin(6 * k, k, g), _Lin(6 * k, k, g), _Lin(3 * k, k, g), _Lin(9 * k, k, g)]
    want = None
    cond = torch.randn(2, k, generator=g).to(torch.bfloat16)
    want = [lin.forward(cond) for lin in lins]          # before the table re-points the weights
    table = AdaLNTable(lins)
    assert table.weight_store.shape == (24 * k, k)
    for lin, w in zip(lins, want):                        # the re-pointed views still compute the same thing
        assert torch.equal(lin.forward(cond), w)
    tabs = table.compute(cond)
    for i, (lin, w, n_chunks) in enumerate(zip(lins, want, (6, 6, 3, 9))):
        chunks = table.chunks(tabs, i, n_chunks)
        assert len(chunks) == n_chunks and all(isinstance(c, ModChunk) for c in chunks)
        ref = w.chunk(n_chunks, dim=1)
        for c, r in zip(chunks, ref):
            assert c.f32.is_contiguous() and c.one_plus.is_contiguous() and c.same.is_contiguous()
            assert c.one_plus.dtype == torch.bfloat16 and c.same.dtype == torch.bfloat16 and c.f32.dtype == torch.float32
            # same dot products; the stacked GEMM may round a few last bits differently than the small one
            assert torch.allclose(c.f32, r.float(), rtol=2e-2, atol=2e-2)
            a, cc = _mod(c, c)
            # the fused LayerNorm-modulate kernel takes (1 + scale) and shift in the model dtype; gates stay fp32
            assert torch.equal(a, c.one_plus) and torch.equal(cc, c.same) and torch.equal(_f32(c), c.f32)
            assert torch.equal(c.same.float(), c.f32)
            assert torch.equal(c.one_plus, 1 + c.same)                               # (1 + x) evaluated in bf16


def test_mod_helpers_on_plain_tensors():
    x = torch.randn(2, 8).to(torch.bfloat16)
    a, c = _mod(x, x)
    assert a.dtype == torch.bfloat16 and torch.equal(a, 1 + x) and torch.equal(c, x)     # bf16 stays bf16
    assert torch.equal(_f32(x), x.float())
    a, c = _mod(x.half(), x.half())                                                       # other dtypes widen to fp32
    assert a.dtype == torch.float32 and torch.equal(a, (1 + x.half()).float()) and torch.equal(c, x.half().float())
