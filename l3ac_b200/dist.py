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


class PendingGather:
    """An all-gather in flight (issued with ``async_op=True``: the collective waits for the producer on the issuing stream,
    the issuing stream does not wait for the collective).  ``wait()`` joins it into the current stream and returns the
    gathered tensor in utterance order."""

    def __init__(self, work, out: torch.Tensor, sizes: List[int], biggest: int):
        self._work, self._out, self._sizes, self._biggest = work, out, sizes, biggest

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if all(s == self._biggest for s in self._sizes):
            return self._out
        b = self._biggest
        return torch.cat([self._out[r * b:r * b + n] for r, n in enumerate(self._sizes)], dim=0)


def gather_batch_async(local: torch.Tensor, n_items: int, group: Optional[dist.ProcessGroup] = None) -> PendingGather:
    """``gather_batch`` without blocking the issuing stream: decode of the local shard can be queued behind it at once."""
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
    work = dist.all_gather_into_tensor(out, padded.contiguous(), group=group, async_op=True)
    return PendingGather(work, out, sizes, biggest)


class ShardedCodec:
    """``encode_audio`` / ``decode_audio`` over a process group: shard by utterance, run locally, gather.

    Two call styles.  *Global*: every rank passes the same whole batch (``encode_audio`` / ``decode_audio``), the rank's
    contiguous utterance range is processed and the results are all-gathered back into utterance order.  *Local shard*
    (``encode_shard`` / ``decode_shard``): every rank already holds its own utterances (a data-parallel serving job); only
    the token indices (and, on request, waveforms) are exchanged, asynchronously, so that the rank's decode never waits for
    the slowest rank's encode."""

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

    # ---- local-shard style ------------------------------------------------------------------------------------------
    def encode_shard(self, local_audio: torch.Tensor, n_items: Optional[int] = None):
        """Encodes this rank's utterances; returns (q_local, indices_dict_local, PendingGather of the int32 indices of the
        whole job).  ``n_items`` = utterances over all ranks (default: world_size * local batch, i.e. equal shards)."""
        world = dist.get_world_size(self.group)
        n = world * local_audio.shape[0] if n_items is None else n_items
        q, idx = self.codec.encode_audio(local_audio)
        return q, idx, gather_batch_async(idx["indices"], n, self.group)

    def decode_shard(self, audio_feature: torch.Tensor = None, indices: torch.Tensor = None, gather: bool = False,
                     n_items: Optional[int] = None, **kw):
        """Decodes this rank's utterances; with ``gather=True`` additionally returns a PendingGather of all waveforms."""
        wav = self.codec.decode_audio(audio_feature, **kw) if audio_feature is not None else self.codec.decode_audio(indices=indices, **kw)
        if not gather:
            return wav
        world = dist.get_world_size(self.group)
        n = world * wav.shape[0] if n_items is None else n_items
        return wav, gather_batch_async(wav, n, self.group)
