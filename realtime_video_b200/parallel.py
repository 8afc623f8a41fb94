"""Single-stream multi-GPU mode: Ulysses-style sequence parallelism for the DiT block stack.

The hot path is ONE strictly sequential stream (B=1; every pass needs the previous one), so it does
not shard into independent units; ``bench.py --gpus N`` therefore defaults to
independent replicas.  This module is the latency-oriented alternative: all N GPUs work on the
same block.

  * token rows are sharded contiguously (L/N rows per rank) for everything token-wise:
    LayerNorm/modulation, QKV / o / cross-attention / FFN GEMMs, RMSNorm+RoPE — weights replicated;
  * self-attention is sharded by HEAD (heads/N per rank, full sequence): one all-to-all turns
    row-sharded q/k/v [L/N, heads*128] into head-sharded [L, heads/N*128] (K and V land directly
    in the rank's head-sharded KV-cache slot), a second one brings the attention output back;
  * per layer and rank that is 4 x (L/N x 5120) bf16 out and in (21 MB at N=8) over NVLink.
The collectives are torch.distributed all_to_all_single / all_gather (NCCL on GPUs; gloo in the
CPU tests).  Results are bit-identical to the single-GPU path: every kernel computes each output
row / head exactly as before, only the placement changes (tests/test_parallel.py).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class SequenceParallel:
    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before SequenceParallel()")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # -- sharding arithmetic -------------------------------------------------------------------
    def rows(self, L: int):
        """(first global row, number of local rows) of this rank for a sequence of L tokens."""
        if L % self.world != 0:
            raise ValueError(f"sequence length {L} is not divisible by {self.world} ranks")
        n = L // self.world
        return self.rank * n, n

    def local_heads(self, heads: int) -> int:
        if heads % self.world != 0:
            raise ValueError(f"{heads} heads are not divisible by {self.world} ranks")
        return heads // self.world

    # -- collectives ---------------------------------------------------------------------------
    def rows_to_heads(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [L/N, H*128] (my rows, all heads) -> [L, H/N*128] (all rows in global order, my heads).
        ``out`` may be a row-contiguous view (e.g. a KV-cache slot) that receives the result."""
        n, width = x.shape
        W = self.world
        blk = width // W
        send = x.view(n, W, blk).transpose(0, 1).contiguous()          # [W, n, blk]: chunk d -> rank d
        if out is None:
            out = torch.empty(W * n, blk, dtype=x.dtype, device=x.device)
        if not out.is_contiguous():
            raise ValueError("rows_to_heads: output must be contiguous")
        dist.all_to_all_single(out.view(W, n, blk), send, group=self.group)   # chunk s <- rank s's rows
        return out

    def heads_to_rows(self, x: torch.Tensor) -> torch.Tensor:
        """x [L, H/N*128] (all rows, my heads) -> [L/N, H*128] (my rows, all heads)."""
        L, blk = x.shape
        W = self.world
        n = L // W
        recv = torch.empty(W, n, blk, dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recv, x.contiguous().view(W, n, blk), group=self.group)
        return recv.transpose(0, 1).reshape(n, W * blk)                # heads of rank s at columns s*blk

    def gather_rows(self, x: torch.Tensor) -> torch.Tensor:
        """x [L/N, C] -> [L, C] on every rank."""
        out = torch.empty(self.world * x.shape[0], *x.shape[1:], dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out
