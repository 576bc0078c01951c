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

The post-attention exchange of O is hidden the same way, from the other side: the rank's H/P heads are attended to in a
few head groups -- launched on two alternating compute streams, so that one group's tail wave and the next group's
first CTAs share the machine --, and each group's all-to-all runs on the communication stream under the remaining
groups' attention; only the last (smallest) group's exchange is exposed (`UlyssesAttention.attend_and_return`).

The head-sharded attention is mathematically the single-GPU attention head by head, so parity =
gather(P-GPU output) vs 1-GPU output to the attention kernel's own tolerance.

On CPU tensors (the world_size-2 gloo tests of the host logic) the pack/unpack layout steps run as
torch view ops and the attention itself is injected by the test (the oracle); the CUDA path always
uses the fdm_ulysses_* kernels and fdm_attn_fwd.
"""
import os
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
        self.attn_stream = None
        # head groups of the post-attention overlap (1 = one attention call, then one all-to-all). Measured at N = 2 on the
        # Wan step (profiles/r02_ulysses.txt): 1 group 2644 ms / 40.9 ms exposed, 2 groups 2655 / 45.5, 4 groups 2667 / 58.5
        # -- NCCL's copy kernels take SMs (and whole SM pairs) away from the attention CTAs they run under, which costs more
        # than the exchange they hide. Default 1; the peer-memory epilogue (`scatter`) is the way to hide this exchange.
        self.o_groups = int(os.environ.get("FDM_ULYSSES_O_GROUPS", "1"))
        # peer-memory epilogue: the attention kernel stores O straight into the owners' buffers (torch symmetric memory
        # maps every rank's buffer into every process); FDM_ULYSSES_SCATTER=0 keeps the NCCL all-to-all
        self.scatter = os.environ.get("FDM_ULYSSES_SCATTER", "1") == "1"
        self._o_sym = None      # [(buffer, handle)] x 2, allocated for one (rows, width, dtype) on first use
        self._o_turn = 0
        self._o_stub = None

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

    # ---- attention over the full sequence for this rank's heads + the exchange of O back to token shards -------------
    def _head_groups(self, hp: int) -> List[range]:
        g = max(1, min(self.o_groups, hp))
        if self.P == 1:
            g = 1
        sizes = [hp // g + (1 if i < hp % g else 0) for i in range(g)]      # larger groups first: the exposed one is small
        out, h = [], 0
        for n in sizes:
            out.append(range(h, h + n))
            h += n
        return out

    def _symmetric_out(self, rows: int, width: int, dtype, device):
        """Two [rows, width] output buffers in symmetric memory (every rank's copy is mapped into every process), used in
        turn: layer L's buffer is read by its out-projection while a faster peer may already be writing layer L+1's."""
        key = (rows, width, dtype)
        if self._o_sym is None or self._o_sym[0] != key:
            import torch.distributed._symmetric_memory as symm_mem

            group = self.group if self.group is not None else dist.group.WORLD
            pairs = []
            for _ in range(2):
                buf = symm_mem.empty((rows, width), dtype=dtype, device=device)
                pairs.append((buf, symm_mem.rendezvous(buf, group.group_name)))
            self._o_sym = (key, pairs)
        self._o_turn ^= 1
        return self._o_sym[1][self._o_turn]

    def _attend_scatter(self, q, k, v, scale, mask, block_q, block_k) -> torch.Tensor:
        """One kernel for attention + the exchange of O: the epilogue stores every query row into the buffer of the rank
        that owns the row (NVLink peer stores, overlapped with the other CTAs' math), then a device-side barrier across
        the ranks (~7 us) orders the owners' reads after everybody's stores. No all-to-all, no unpack pass."""
        H, hd, P = self.H, self.hd, self.P
        hp, d = H // P, H * hd
        S_loc = q.shape[1] // P
        col0 = self.rank * hp * hd * q.element_size()     # this rank's head columns in every owner's [S/P, H*hd] buffer
        if self.stub_comm:
            # timing aid only (see stub_comm): the same stores, into local memory
            if self._o_stub is None or self._o_stub.shape != (P, S_loc, d):
                self._o_stub = torch.empty((P, S_loc, d), device=q.device, dtype=q.dtype)
            ptrs = [self._o_stub[p_].data_ptr() + col0 for p_ in range(P)]
            ops.attention_scatter(q, k, v, hp, hd, ptrs, S_loc, d, scale, mask, block_q, block_k)
            return self._o_stub[self.rank].view(1, S_loc, d)
        buf, hdl = self._symmetric_out(S_loc, d, q.dtype, q.device)
        ptrs = [int(hdl.buffer_ptrs[p_]) + col0 for p_ in range(P)]
        ops.attention_scatter(q, k, v, hp, hd, ptrs, S_loc, d, scale, mask, block_q, block_k)
        hdl.barrier(channel=0)
        return buf.view(1, S_loc, d)

    def attend_and_return(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale, sparse_mask=None,
                          block_q: int = 128, block_k: int = 64) -> torch.Tensor:
        """q, k, v [1, S, (H/P)*hd] (views with any token stride): attention for this rank's heads over the full sequence,
        then the all-to-all that hands every rank its token shard of all heads -> [1, S/P, H*hd].
        The heads go group by group; a group's exchange runs on the communication stream while the next groups are being
        attended to (on two alternating compute streams -- consecutive groups overlap at their wave boundaries)."""
        H, hd, P = self.H, self.hd, self.P
        hp = H // P
        S = q.shape[1]
        S_loc = S // P
        mask = self.local_mask(sparse_mask)
        if self.scatter and P > 1 and q.is_cuda and hd == 128 and q.dtype == torch.bfloat16 and q.shape[0] == 1 \
                and (S_loc * P == S):
            return self._attend_scatter(q, k, v, scale, mask, block_q, block_k)
        groups = self._head_groups(hp)
        if len(groups) == 1 or not q.is_cuda:
            o = ops.attention(q, k, v, hp, hd, scale, mask, block_q, block_k)
            back, _ = self._a2a(o.reshape(P, S_loc, hp * hd).contiguous())
            return _unpack(back, H, hd, 1).view(1, S_loc, H * hd)
        cur = torch.cuda.current_stream()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        if self.attn_stream is None:
            self.attn_stream = torch.cuda.Stream()
        start = torch.cuda.Event()
        start.record(cur)
        self.attn_stream.wait_event(start)
        # staging [group][P, S/P, heads_g * hd]; assembled at the end into [S/P, P, H/P, hd] = [S/P, H*hd]
        recvs, dones = [], []
        for gi, hs in enumerate(groups):
            stream = cur if gi % 2 == 0 else self.attn_stream
            cols = slice(hs.start * hd, hs.stop * hd)
            with torch.cuda.stream(stream):
                m = None if mask is None else mask[:, hs.start:hs.stop].contiguous()
                o = ops.attention(q[:, :, cols], k[:, :, cols], v[:, :, cols], len(hs), hd, scale, m, block_q, block_k)
                ready = torch.cuda.Event()
                ready.record(stream)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                recv, _ = self._a2a(o.view(P, S_loc, len(hs) * hd))
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            for t in (o, recv):
                t.record_stream(self.comm_stream)
                t.record_stream(cur)
            recvs.append(recv)
            dones.append(done)
        for e in dones:
            cur.wait_event(e)
        for t in (q, k, v):
            t.record_stream(self.attn_stream)
        out = torch.empty((S_loc, P, hp, hd), device=q.device, dtype=recvs[0].dtype)
        for hs, recv in zip(groups, recvs):
            out[:, :, hs.start:hs.stop] = recv.view(P, S_loc, len(hs), hd).transpose(0, 1)
        return out.view(1, S_loc, H * hd)

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
            back, _ = self._a2a(o.reshape(P, S_loc, dp).contiguous())  # chunk p: head group p of my tokens
            return _unpack(back, H, hd, 1).view(1, S_loc, d)
        return self.attend_and_return(q, k, v, scale, sparse_mask, block_q, block_k)

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
        return self.attend_and_return(q, k, v, scale, sparse_mask, block_q, block_k)
