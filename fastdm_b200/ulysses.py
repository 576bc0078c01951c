"""Ulysses sequence-parallel attention for the long-sequence DiT models (Wan2.2: 40 heads,
Qwen-Image: 24 heads) -- new functionality, the reference has no multi-GPU path.

Layout: every rank holds a contiguous token shard [S/P, d] through the whole block (all GEMMs,
norms, RoPE, FFN and the cross-attention are token-local, weights replicated). Around the
self-attention two all-to-alls re-shard between tokens and heads:

    local qkv [S/P, 3, H, hd] --pack--> [P, S/P, 3, H/P, hd] --all_to_all--> [S, 3, H/P, hd]
        attention over the FULL sequence for this rank's H/P heads  (same kernel, strided views)
    out [S, H/P, hd] = [P, S/P, H/P, hd] --all_to_all--> [P(head group), S/P, H/P, hd] --unpack--> [S/P, H, hd]

The exchange is NCCL all_to_all_single (NVLink 5 / NVSwitch: uniform bandwidth to every peer, so a
flat all-to-all is the right schedule). `UlyssesAttention.qkv_projection_overlapped` splits the
fused QKV projection into its Q | K | V column groups and ships each group on a side stream while
the next group's GEMM runs, so only the last group's exchange is exposed.

The head-sharded attention is mathematically the single-GPU attention head by head, so parity =
gather(P-GPU output) vs 1-GPU output to the attention kernel's own tolerance.

On CPU tensors (the world_size-2 gloo tests of the host logic) the pack/unpack layout steps run as
torch view ops and the attention itself is injected by the test (the oracle); the CUDA path always
uses the fdm_ulysses_* kernels and fdm_attn_fwd.
"""
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import ops


def _pack(x: torch.Tensor, H: int, hd: int, P: int, n_seg: int) -> torch.Tensor:
    """[S, n_seg*H*hd] -> [P, S, n_seg*(H/P)*hd]"""
    if x.is_cuda:
        return ops.ulysses_pack_heads(x, H, hd, P, n_seg)
    S = x.shape[0]
    v = x[:, : n_seg * H * hd].reshape(S, n_seg, P, H // P, hd)
    return v.permute(2, 0, 1, 3, 4).reshape(P, S, n_seg * (H // P) * hd).contiguous()


def _unpack(x: torch.Tensor, H: int, hd: int, n_seg: int) -> torch.Tensor:
    """[P, S, n_seg*(H/P)*hd] -> [S, n_seg*H*hd]"""
    if x.is_cuda:
        return ops.ulysses_unpack_heads(x, H, hd, n_seg)
    P, S, _ = x.shape
    v = x.reshape(P, S, n_seg, H // P, hd).permute(1, 2, 0, 3, 4)
    return v.reshape(S, n_seg * H * hd).contiguous()


class UlyssesAttention:
    def __init__(self, num_heads: int, head_dim: int, group: Optional[dist.ProcessGroup] = None,
                 stub_comm: bool = False):
        self.H, self.hd = num_heads, head_dim
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.H % self.P:
            raise ValueError(f"Ulysses needs heads ({self.H}) divisible by ranks ({self.P})")
        # stub_comm replaces the exchange by a local copy: used ONLY to measure exposed all-to-all time
        self.stub_comm = stub_comm
        self.comm_stream = None

    # ---- the exchange ----------------------------------------------------------------------------
    def _a2a(self, send: torch.Tensor, async_op: bool = False):
        recv = torch.empty_like(send)
        if self.P == 1 or self.stub_comm:
            recv.copy_(send)
            return recv, None
        work = dist.all_to_all_single(recv, send, group=self.group, async_op=async_op)
        return recv, work

    def shard_tokens(self, x: torch.Tensor, dim: int = 1) -> torch.Tensor:
        """This rank's contiguous token chunk of a replicated tensor."""
        n = x.shape[dim]
        if n % self.P:
            raise ValueError(f"sequence length {n} not divisible by {self.P} ranks")
        c = n // self.P
        return x.narrow(dim, self.rank * c, c).contiguous()

    def gather_tokens(self, x: torch.Tensor, dim: int = 1) -> torch.Tensor:
        if self.P == 1:
            return x
        parts = [torch.empty_like(x) for _ in range(self.P)]
        dist.all_gather(parts, x.contiguous(), group=self.group)
        return torch.cat(parts, dim=dim)

    def local_mask(self, sparse_mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """A block mask given for all H heads ([B, H, nbq, nbk]) -> the [B, H/P, nbq, nbk] slice of this rank's
        heads (a mask that already has H/P heads is taken as is)."""
        if sparse_mask is None or self.P == 1 or sparse_mask.shape[1] == self.H // self.P:
            return sparse_mask
        if sparse_mask.shape[1] != self.H:
            raise ValueError(f"sparse mask has {sparse_mask.shape[1]} heads, expected {self.H} or {self.H // self.P}")
        hp = self.H // self.P
        return sparse_mask[:, self.rank * hp:(self.rank + 1) * hp].contiguous()

    # ---- attention on a local fused qkv shard ------------------------------------------------------
    def attention(self, qkv_local: torch.Tensor, scale: Optional[float] = None, sparse_mask=None,
                  block_q: int = 128, block_k: int = 64,
                  attention_fn: Optional[Callable] = None) -> torch.Tensor:
        """qkv_local [1, S/P, 3*H*hd] (q|k|v, q and k already normalised + rotated with THIS shard's
        positions) -> attention output for the local tokens [1, S/P, H*hd].
        sparse_mask, if given, is the [1, H, nbq, nbk] mask of all heads (sliced here) or the [1, H/P, nbq, nbk]
        mask of this rank's heads."""
        H, hd, P = self.H, self.hd, self.P
        d = H * hd
        if qkv_local.shape[0] != 1:
            raise NotImplementedError("Ulysses path is written for batch 1 (CFG halves are separate forwards)")
        x = qkv_local[0]
        S_loc = x.shape[0]
        send = _pack(x, H, hd, P, 3)                             # [P, S/P, 3*(H/P)*hd]
        recv, _ = self._a2a(send)
        full = recv.view(1, P * S_loc, 3 * d // P)               # [1, S, 3, H/P, hd]
        dp = d // P
        q, k, v = full[:, :, :dp], full[:, :, dp:2 * dp], full[:, :, 2 * dp:]
        if attention_fn is not None:
            o = attention_fn(q, k, v, H // P, hd, scale)
        else:
            o = ops.attention(q, k, v, H // P, hd, scale, self.local_mask(sparse_mask), block_q, block_k)
        back, _ = self._a2a(o.reshape(P, S_loc, dp).contiguous())  # chunk p: head group p of my tokens
        return _unpack(back, H, hd, 1).view(1, S_loc, d)

    # ---- QKV projection with the exchange overlapped ------------------------------------------------
    def qkv_projection_overlapped(self, project: List[Callable[[], torch.Tensor]], scale=None,
                                  sparse_mask=None, block_q=128, block_k=64) -> torch.Tensor:
        """`project` = three thunks producing this shard's q, k, v ([S/P, H*hd] each, q/k already
        normalised + rotated). Group g's all-to-all runs on a side stream while group g+1 is being
        projected on the compute stream; only V's exchange is exposed."""
        H, hd, P = self.H, self.hd, self.P
        d = H * hd
        dp = d // P
        if len(project) != 3:
            raise ValueError("qkv_projection_overlapped takes the three thunks of q, k and v")
        cur = torch.cuda.current_stream()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        recvs, events = [], []
        for thunk in project:
            x = thunk()                                           # [S/P, d] on the compute stream
            if x.ndim != 2 or x.shape[1] != d:
                # [B*S/P, d] with B > 1 would be exchanged (and attended to) as one long sequence
                raise NotImplementedError(f"Ulysses: projections must be one sequence's [S/P, {d}] shard (batch 1), "
                                          f"got {tuple(x.shape)}")
            send = _pack(x, H, hd, P, 1)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                recv, _ = self._a2a(send)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            send.record_stream(self.comm_stream)
            recv.record_stream(cur)
            recvs.append(recv)
            events.append(done)
        for e in events:
            cur.wait_event(e)
        S_loc = recvs[0].shape[1]
        q, k, v = (r.view(1, P * S_loc, dp) for r in recvs)
        o = ops.attention(q, k, v, H // P, hd, scale, self.local_mask(sparse_mask), block_q, block_k)
        back, _ = self._a2a(o.reshape(P, S_loc, dp).contiguous())
        return _unpack(back, H, hd, 1).view(1, S_loc, d)
