"""CPU, world_size 2, gloo: the host-side logic of the Ulysses sequence-parallel attention
(fastdm_b200/ulysses.py) -- pack / all-to-all / head-sharded attention / all-to-all / unpack --
against the single-process attention on the gathered tensors. The attention arithmetic itself is
injected from the oracle (there is no GPU here); the CUDA path is covered by tests/test_gpu_ulysses.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ops_ref as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_attn(q, k, v, heads, hd, scale):
    return R.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous(), heads, heads, hd, scale=scale)


def _worker(rank, world, port, S, H, hd, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastdm_b200.ulysses import UlyssesAttention

        torch.manual_seed(0)
        d = H * hd
        qkv = torch.randn(1, S, 3 * d).to(torch.bfloat16)       # same on every rank
        ul = UlyssesAttention(H, hd)
        local = ul.shard_tokens(qkv, dim=1)
        assert local.shape == (1, S // world, 3 * d)
        out_local = ul.attention(local, hd ** -0.5, attention_fn=_oracle_attn)
        full = ul.gather_tokens(out_local, dim=1)
        want = R.scaled_dot_product_attention(qkv[:, :, :d].contiguous(), qkv[:, :, d:2 * d].contiguous(),
                                              qkv[:, :, 2 * d:].contiguous(), H, H, hd, scale=hd ** -0.5)
        ok = torch.equal(full, want)   # per head the arithmetic is identical -> bit-equal
        # pack/unpack are inverses
        from fastdm_b200.ulysses import _pack, _unpack
        x = torch.randn(S // world, 3 * d).to(torch.bfloat16)
        ok = ok and torch.equal(_unpack(_pack(x, H, hd, world, 3), H, hd, 3), x)
        if rank == 0:
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ulysses_attention_world2_gloo():
    world, S, H, hd = 2, 96, 4, 64
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, H, hd, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_ulysses_rejects_indivisible_heads():
    from fastdm_b200.ulysses import UlyssesAttention

    ul = UlyssesAttention(5, 64)   # no process group: P = 1, anything divides
    assert ul.P == 1
    x = torch.randn(1, 10, 3 * 5 * 64).to(torch.bfloat16)
    y = ul.attention(x, 0.125, attention_fn=_oracle_attn)
    want = _oracle_attn(x[:, :, :320], x[:, :, 320:640], x[:, :, 640:], 5, 64, 0.125)
    assert torch.equal(y, want)
