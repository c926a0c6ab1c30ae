"""CPU: the N>1 host logic (utterance sharding + gather) on the gloo backend, world_size 2 and 3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from l3ac_b200.dist import ShardedCodec, gather_batch, shard_batch, shard_bounds, shard_sizes


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _FakeCodec:
    """Stands in for the CUDA codec: a per-utterance deterministic map, so sharding must be transparent."""

    def encode_audio(self, audio):
        idx = (audio.abs().sum(dim=1, keepdim=True) * 1000).to(torch.int32).repeat(1, 7)
        return audio[:, :7, None].repeat(1, 1, 4), {"indices": idx, "level_indices": idx.float()[..., None].repeat(1, 1, 6)}

    def decode_audio(self, audio_feature=None, indices=None):
        src = audio_feature[..., 0] if audio_feature is not None else indices.float()
        return src.repeat(1, 3)


def _worker(rank, world, port, n_items):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        audio = torch.randn(n_items, 50, generator=g)
        local = shard_batch(audio)
        lo, hi = shard_bounds(n_items, world, rank)
        assert torch.equal(local, audio[lo:hi])
        assert torch.equal(gather_batch(local.clone(), n_items), audio)
        ref, sharded = _FakeCodec(), ShardedCodec(_FakeCodec())
        q, idx = sharded.encode_audio(audio)
        rq, ridx = ref.encode_audio(audio)
        assert torch.equal(q, rq) and torch.equal(idx["indices"], ridx["indices"])
        assert torch.equal(sharded.decode_audio(indices=idx["indices"]), ref.decode_audio(indices=ridx["indices"]))
        assert torch.equal(sharded.decode_audio(q), ref.decode_audio(rq))
        # local-shard style: each rank brings its own utterances; async index / waveform gathers
        ql, idxl, pending = sharded.encode_shard(local, n_items=n_items)
        assert torch.equal(ql, rq[lo:hi]) and torch.equal(pending.wait(), ridx["indices"])
        wl, pw = sharded.decode_shard(indices=idxl["indices"], gather=True, n_items=n_items)
        assert torch.equal(wl, ref.decode_audio(indices=ridx["indices"])[lo:hi])
        assert torch.equal(pw.wait(), ref.decode_audio(indices=ridx["indices"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 8), (2, 5), (3, 7)])
def test_sharded_codec_gloo(world, n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n_items), nprocs=world, join=True)
