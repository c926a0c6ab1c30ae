"""GPU: the step-level C ABI (l3ac_create / l3ac_encode / l3ac_decode, csrc/codec.cu) against the oracle, against the
operator-level Python sequence (same kernels, Python-side packing) and through a plain-C host with no Python in the loop."""
import ctypes
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import l3ac_b200
from l3ac_b200 import _lib, ops
from l3ac_b200.engine import Engine
from helpers import config_path, golden_case, make_audio, model_config, snr_db
from l3ac_b200.config import L3ACConfig
from l3ac_b200.spec import init_state_dicts
from oracle import l3ac_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent


def _engines(name, weights, monkeypatch, precision="bf16"):
    mc = model_config(name)
    native = Engine(mc, weights, DEV, precision=precision)
    assert native.native is not None, "the tensor-core precisions must run through the step-level ABI"
    monkeypatch.setenv("L3AC_ENGINE", "python")
    python = Engine(mc, weights, DEV, precision=precision)
    assert python.native is None
    return mc, native, python


def _prefolded(weights):
    """The same checkpoint with every weight-norm pair replaced by its folded ``.weight`` (both packers take either form) and
    a constant DynamicPositionBias table (last MLP layer zeroed): the two packers then see bit-identical fp32 inputs and no
    packer-side arithmetic is left, so any output difference would be a difference in the launch sequence."""
    out = {}
    for mod, sd in weights.items():
        new = {}
        for k, v in sd.items():
            if k.endswith("parametrizations.weight.original0"):
                p = k[:-len(".parametrizations.weight.original0")]
                vv = sd[p + ".parametrizations.weight.original1"].float()
                norm = vv.flatten(1).norm(dim=1).reshape(v.shape)
                new[p + ".weight"] = (vv * (v.float() / norm)).contiguous()
            elif k.endswith("parametrizations.weight.original1"):
                continue
            elif k.endswith("dynamic_pos_bias.mlp.4.weight"):
                new[k] = torch.zeros_like(v)
            else:
                new[k] = v
        out[mod] = new
    return out


@pytest.mark.parametrize("precision", ["bf16", "split"])
@pytest.mark.parametrize("name", ["1kbps", "3kbps", "0k75bps"])
def test_native_engine_is_the_python_sequence(cuda_lib, name, precision, monkeypatch):
    """l3ac_encode / l3ac_decode launch the same kernels with the same arguments as the operator-level Python sequence:
    with packer-side arithmetic taken out (pre-folded weights) the indices, features and waveforms are BIT-identical."""
    weights = _prefolded(init_state_dicts(model_config(name), seed=11, jitter=True))
    mc, native, python = _engines(name, weights, monkeypatch, precision)
    audio = make_audio(3, 2.0 + 0.37, seed=5).to(DEV)          # not a multiple of the hop: the library pads
    with torch.inference_mode():
        qn, dn = native.encode(audio)
        qp, dp = python.encode(audio)
        wn = native.decode(indices=dp["indices"])
        wp = python.decode(indices=dp["indices"])
        wq = native.decode(qp)
        w64 = native.decode(indices=dp["indices"].long())
    torch.cuda.synchronize()
    assert dn["indices"].dtype == torch.int32 and torch.equal(dn["indices"], dp["indices"])
    assert torch.equal(dn["level_indices"], dp["level_indices"]) and torch.equal(qn, qp)
    assert torch.equal(wn, wp)
    assert torch.equal(wq, wn) and torch.equal(w64, wn)         # q_feature / int64 index inputs: the same launches after dequantize


def test_native_engine_matches_oracle(cuda_lib, monkeypatch):
    """Through the public API (native engine underneath) against the CPU oracle: indices and waveform."""
    mc, weights, audio, g = golden_case("1kbps")
    codec = l3ac_b200.L3AC(L3ACConfig(config_file=config_path("1kbps")))
    codec.network.load_state_dicts(weights)
    codec.network.cuda()
    assert codec.network.engine.native is not None
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio.to(DEV))
        wav = codec.decode_audio(indices=torch.from_numpy(g["indices"]).to(DEV))
    agree = float((idx["indices"].cpu().numpy() == g["indices"]).mean())
    stride = int(g["wav_stride"])
    snr = snr_db(torch.from_numpy(g["wav"]), wav.cpu()[:, ::stride])
    # the operator-level Python sequence (torch-side packing) on the same input: the decoder amplifies ulp-level differences
    # between the two packers' folded weights, so the two are compared through their distance to the reference
    monkeypatch.setenv("L3AC_ENGINE", "python")
    py = Engine(mc, weights, DEV)
    with torch.inference_mode():
        wav_py = py.decode(indices=torch.from_numpy(g["indices"]).to(DEV))
    snr_py = snr_db(torch.from_numpy(g["wav"]), wav_py.cpu()[:, ::stride])
    print(f"native engine vs golden: index agreement {agree:.5f}, decode SNR {snr:.1f} dB (python sequence: {snr_py:.1f} dB; bf16 decode side)")
    assert agree >= 0.999
    assert snr > 18.0 and abs(snr - snr_py) < 3.0


def test_micro_batches_of_odd_length_clips(cuda_lib, monkeypatch):
    """Clips whose length is neither a multiple of the hop nor of four samples, processed in many micro-batches (row slices of
    the batch are then only 4-byte aligned): every micro-batching gives the result of the one-clip-at-a-time run."""
    mc = model_config("1kbps")
    weights = init_state_dicts(mc, seed=13, jitter=True)
    monkeypatch.setenv("L3AC_CHUNK_SECONDS", "3")
    eng = Engine(mc, weights, DEV)
    assert eng.native is not None
    audio = make_audio(45, 16005 / 16000.0, seed=21).to(DEV)         # 45 x 16005 samples: 23 micro-batches of 2 (and 1) clips
    assert audio.shape[1] == 16005 and len(eng._chunks(*audio.shape)) > 20
    with torch.inference_mode():
        q, d = eng.encode(audio)
        wav = eng.decode(indices=d["indices"])
        for i in (0, 1, 2, 44):
            qi, di = eng.encode(audio[i:i + 1])
            assert torch.equal(di["indices"], d["indices"][i:i + 1]) and torch.equal(qi, q[i:i + 1])
            assert torch.equal(eng.decode(indices=di["indices"]), wav[i:i + 1])
    assert wav.shape == (45, d["indices"].shape[1] * mc.hop_length)


def test_workspace_contract(cuda_lib):
    """l3ac_workspace_bytes is the exact high-water mark: the call succeeds with it and is refused one block below it;
    bad arguments are rejected with L3AC_EINVAL, not a crash."""
    mc = model_config("1kbps")
    weights = init_state_dicts(mc, seed=3, jitter=True)
    nc = ops.NativeCodec(mc, weights, DEV)
    lib = _lib.load()
    B, T = 2, 16000
    need = lib.l3ac_workspace_bytes(nc.handle, B, T)
    assert need > 0 and need % 512 == 0
    assert lib.l3ac_workspace_bytes(nc.handle, 2 * B, T) > need
    audio = make_audio(B, 1.0, seed=2).to(DEV)
    t_tok = -(-T // nc.hop)
    idx = torch.empty((B, t_tok), device=DEV, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(need, device=DEV, dtype=torch.uint8)
    assert lib.l3ac_encode(nc.handle, audio.data_ptr(), B, T, ws.data_ptr(), need, None, idx.data_ptr(), None, st) == 0
    torch.cuda.synchronize()
    ref = nc.encode(audio)[1]
    assert torch.equal(idx, ref)
    wav = torch.empty((B, t_tok * nc.hop), device=DEV, dtype=torch.float32)
    assert lib.l3ac_decode(nc.handle, idx.data_ptr(), 0, None, B, t_tok, ws.data_ptr(), need, wav.data_ptr(), st) == 0
    # the figure is the larger of the two calls' high-water marks: one block less and (at least) that call is refused
    rc = [lib.l3ac_encode(nc.handle, audio.data_ptr(), B, T, ws.data_ptr(), need - 512, None, idx.data_ptr(), None, st),
          lib.l3ac_decode(nc.handle, idx.data_ptr(), 0, None, B, t_tok, ws.data_ptr(), need - 512, wav.data_ptr(), st)]
    assert -1 in rc and set(rc) <= {0, -1}
    assert lib.l3ac_encode(nc.handle, audio.data_ptr(), B, T, ws.data_ptr(), 4096, None, idx.data_ptr(), None, st) == -1
    assert b"workspace" in lib.l3ac_last_error()
    assert lib.l3ac_encode(nc.handle, None, B, T, ws.data_ptr(), need, None, idx.data_ptr(), None, st) == -1
    assert lib.l3ac_decode(nc.handle, None, 0, None, B, t_tok, ws.data_ptr(), need, audio.data_ptr(), st) == -1
    torch.cuda.synchronize()


def test_quantizer_entry_points(cuda_lib):
    """l3ac_quantize / l3ac_dequantize (the handle's VQEmbed.forward / to_features) against the operator-level calls."""
    mc = model_config("3kbps")
    weights = init_state_dicts(mc, seed=5, jitter=True)
    nc = ops.NativeCodec(mc, weights, DEV)
    eng = Engine(mc, weights, DEV)
    lib = _lib.load()
    feat = torch.randn(3, 211, 128, device=DEV)
    q0, i0, l0, _ = eng.quantize(feat)
    q = torch.empty_like(q0); idx = torch.empty_like(i0); lvl = torch.empty_like(l0)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.l3ac_quantize(nc.handle, feat.data_ptr(), 3, 211, q.data_ptr(), idx.data_ptr(), lvl.data_ptr(), st) == 0
    assert torch.equal(q, q0) and torch.equal(idx, i0) and torch.equal(lvl, l0)
    for ind in (i0, i0.long()):
        d = torch.empty_like(q0)
        assert lib.l3ac_dequantize(nc.handle, ind.data_ptr(), int(ind.dtype == torch.int64), 3, 211, d.data_ptr(), st) == 0
        assert torch.equal(d, eng.dequantize(ind)) and torch.equal(d, q0)
    assert lib.l3ac_quantize(nc.handle, None, 3, 211, q.data_ptr(), idx.data_ptr(), None, st) == -1


def test_create_rejects_bad_checkpoints(cuda_lib):
    mc = model_config("1kbps")
    weights = init_state_dicts(mc, seed=3)
    del weights["decoder"]["blocks.0.bias"]
    with pytest.raises(ValueError, match="decoder.blocks.0.bias"):
        ops.NativeCodec(mc, weights, DEV)
    weights = init_state_dicts(mc, seed=3)
    weights["encoder"]["blocks.0.conv_1.bias"] = torch.zeros(79)
    with pytest.raises(ValueError, match="expected 80"):
        ops.NativeCodec(mc, weights, DEV)
    with pytest.raises(ValueError, match="rotary"):
        ops.NativeCodec(model_config("rotary"), init_state_dicts(model_config("rotary"), seed=3), DEV)


def _dump_weights(path, mc, weights):
    cfg = _lib.CodecConfig()
    cfg.feature_dim, cfg.n_encoder_stages, cfg.n_decoder_stages = mc.feature_dim, len(mc.encoder_dims), len(mc.decoder_dims)
    for name in ("encoder_dims", "encoder_depths", "compress_rates", "decoder_dims", "decoder_depths", "decode_rates"):
        for i, v in enumerate(getattr(mc, name)):
            getattr(cfg, name)[i] = int(v)
    cfg.en_coder_depth, cfg.en_coder_window_size = mc.en_coder_depth, mc.en_coder_window_size
    cfg.en_coder_compress_rate, cfg.en_coder_dynamic_pos = mc.en_coder_compress_rate, int(mc.en_coder_dynamic_pos)
    cfg.n_levels = len(mc.levels)
    for i, v in enumerate(mc.levels):
        cfg.levels[i] = int(v)
    cfg.precision = 0
    items = [(f"{m}.{k}", t) for m, sd in weights.items() for k, t in sd.items()]
    with open(path, "wb") as f:
        f.write(b"L3ACW1\0\0")
        f.write(bytes(cfg))
        f.write(struct.pack("<i", len(items)))
        for name, t in items:
            a = t.detach().cpu().float().contiguous().numpy()
            f.write(struct.pack("<i", len(name)) + name.encode() + struct.pack("<q", a.size))
            f.write(a.tobytes())


def test_plain_c_host_runs_the_path(cuda_lib, tmp_path):
    """tests/c/codec_driver.c (gcc, no Python, no CUDA headers): l3ac_create -> l3ac_encode_host -> l3ac_decode_host on a
    batch that spans several micro-batches; indices and waveform equal the Python host's (same library, same packing)."""
    exe = tmp_path / "codec_driver"
    lib_dir = _lib.LIB_PATH.parent
    subprocess.run(["gcc", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "codec_driver.c"), "-L", str(lib_dir),
                    "-ll3ac_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    mc = model_config("1kbps")
    weights = init_state_dicts(mc, seed=7, jitter=True)
    _dump_weights(tmp_path / "w.bin", mc, weights)
    B, secs = 70, 10.0                                            # 700 s of audio: three micro-batches on three streams
    audio = make_audio(B, secs, seed=9)
    with open(tmp_path / "a.bin", "wb") as f:
        f.write(struct.pack("<ii", B, audio.shape[1]))
        f.write(audio.numpy().tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "w.bin"), str(tmp_path / "a.bin"), str(tmp_path / "o.bin"), "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    print("C host:", r.stdout.strip())
    raw = (tmp_path / "o.bin").read_bytes()
    b, t_tok, hop = struct.unpack("<iii", raw[:12])
    assert (b, hop) == (B, mc.hop_length)
    idx = np.frombuffer(raw, dtype=np.int32, count=b * t_tok, offset=12).reshape(b, t_tok)
    wav = np.frombuffer(raw, dtype=np.float32, count=b * t_tok * hop, offset=12 + 4 * b * t_tok).reshape(b, t_tok * hop)
    nc = ops.NativeCodec(mc, weights, DEV)
    with torch.inference_mode():
        ref_idx, ref_wav = [], []
        for lo in range(0, B, 24):                                # any batching: every reduction on the path is per clip
            q, i, _ = nc.encode(audio[lo:lo + 24].to(DEV))
            ref_idx.append(i)
            ref_wav.append(nc.decode(indices=i))
        ref_idx, ref_wav = torch.cat(ref_idx).cpu().numpy(), torch.cat(ref_wav).cpu().numpy()
    assert np.array_equal(idx, ref_idx)
    assert np.array_equal(wav, ref_wav)
