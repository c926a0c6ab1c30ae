"""GPU, world size 2 over NCCL: the real ``ShardedCodec`` (ragged utterance shards, index AND waveform gathers) must give
exactly what one GPU gives for the same batch.  Needs two visible GPUs (``gpurun --gpus 2``); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_items, tmp):
    import l3ac_b200
    from helpers import make_audio, model_config
    from l3ac_b200.dist import ShardedCodec, shard_bounds
    from l3ac_b200.spec import init_state_dicts
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        codec = l3ac_b200.get_model("1k5bps", pretrained=False)
        codec.network.load_state_dicts(init_state_dicts(model_config("1k5bps"), seed=23, jitter=True))
        codec.network.to(dev).eval()
        audio = make_audio(n_items, 2.3, seed=8).to(dev)            # every rank holds the whole batch (global style)
        sharded = ShardedCodec(codec)
        with torch.inference_mode():
            q, idx = sharded.encode_audio(audio)
            wav = sharded.decode_audio(indices=idx["indices"])
            wav_q = sharded.decode_audio(q)
            # local-shard style with asynchronous gathers
            lo, hi = shard_bounds(n_items, world, rank)
            ql, idxl, pending = sharded.encode_shard(audio[lo:hi], n_items=n_items)
            wl, pw = sharded.decode_shard(indices=idxl["indices"], gather=True, n_items=n_items)
            all_idx, all_wav = pending.wait(), pw.wait()
            # single-GPU truth on this rank's own device
            q1, idx1 = codec.encode_audio(audio)
            wav1 = codec.decode_audio(indices=idx1["indices"])
        torch.cuda.synchronize()
        assert q.shape[0] == n_items and wav.shape[0] == n_items
        assert torch.equal(idx["indices"], idx1["indices"]) and torch.equal(idx["level_indices"], idx1["level_indices"])
        assert torch.equal(q, q1) and torch.equal(wav, wav1) and torch.equal(wav_q, wav1)
        assert torch.equal(all_idx, idx1["indices"]) and torch.equal(all_wav, wav1)
        assert torch.equal(ql, q1[lo:hi]) and torch.equal(wl, wav1[lo:hi])
        with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 4])
def test_sharded_codec_nccl_equals_single_gpu(cuda_lib, tmp_path, n_items):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ["PYTHONPATH"] = os.pathsep.join([os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)),
                                                os.environ.get("PYTHONPATH", "")])
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
