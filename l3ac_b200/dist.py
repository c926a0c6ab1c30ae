"""Utterance-level data parallelism: one process per GPU, weights replicated, no collective inside the forward.

Every operator of the path is per-utterance (GRN and InstanceNorm reduce over time/channel of one sample; there
are no batch statistics), so a batch shards by contiguous utterance ranges and the only exchange step is the final
gather of token indices (int32, B*T_tok*4 bytes) and/or waveforms.  ``torch.distributed``
is plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_items`` utterances for ``rank`` (first ranks get the remainder)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world_size: int) -> List[int]:
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0] for r in range(world_size)]


def shard_batch(batch: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """This rank's utterances of a batch every rank holds (or a view of an identically-shaped placeholder)."""
    lo, hi = shard_bounds(batch.shape[0], dist.get_world_size(group), dist.get_rank(group))
    return batch[lo:hi]


def gather_batch(local: torch.Tensor, n_items: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gathers per-rank results (indices ``(b, T_tok)`` or waveforms ``(b, T)``) back into utterance order.

    Shards may be ragged by one utterance; they are padded to the largest shard for the collective and trimmed after.
    """
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"rank holds {local.shape[0]} utterances, expected {sizes[dist.get_rank(group)]}")
    biggest = max(sizes)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(s == biggest for s in sizes):
        return out
    return torch.cat([out[r * biggest:r * biggest + sizes[r]] for r in range(world)], dim=0)


class ShardedCodec:
    """``encode_audio`` / ``decode_audio`` over a process group: shard by utterance, run locally, gather."""

    def __init__(self, codec, group: Optional[dist.ProcessGroup] = None):
        self.codec, self.group = codec, group

    def encode_audio(self, audio: torch.Tensor):
        n = audio.shape[0]
        q, idx = self.codec.encode_audio(shard_batch(audio, self.group))
        return gather_batch(q, n, self.group), {k: gather_batch(v, n, self.group) for k, v in idx.items()}

    def decode_audio(self, audio_feature: torch.Tensor = None, indices: torch.Tensor = None):
        src = audio_feature if audio_feature is not None else indices
        n = src.shape[0]
        local = shard_batch(src, self.group)
        wav = self.codec.decode_audio(local) if audio_feature is not None else self.codec.decode_audio(indices=local)
        return gather_batch(wav, n, self.group)
