"""CPU: host-side logic -- config names and fields, checkpoint key inventory, C-ABI exports, error behaviour."""
import ctypes
import json
import re
from pathlib import Path

import pytest
import torch

import l3ac_b200
from helpers import CONFIGS, GOLDEN, model_config
from l3ac_b200 import _lib
from l3ac_b200.config import CONFIG_DIR, L3ACConfig, ModelConfig
from l3ac_b200.spec import init_state_dicts, network_spec

ROOT = Path(__file__).resolve().parent.parent


def test_list_models():
    """l3ac/__init__.py:17-18 lists every TOML stem, ``debug`` included (which cannot be loaded, there as here)."""
    assert set(l3ac_b200.list_models()) == set(CONFIGS) | {"debug"}
    with pytest.raises(ValueError):
        l3ac_b200.get_model("debug", pretrained=False)


def test_pretrained_missing_raises(tmp_path, monkeypatch):
    """get_model(pretrained=True) must not hand back a random-weight codec: no files and no network -> an exception."""
    cfg = L3ACConfig(config_file=CONFIG_DIR / "1kbps.toml", model_dir=tmp_path)
    codec = l3ac_b200.L3AC(cfg)
    import requests

    def no_network(url, **kw):
        raise requests.ConnectionError(f"offline: {url}")
    monkeypatch.setattr(requests, "get", no_network)
    with pytest.raises(requests.ConnectionError):
        codec.load_pretrained()
    assert codec.config.weight_url.format("encoder").endswith("weights/1kbps.v1/encoder.pt")
    # files present -> loaded without touching the network
    codec.network.save_model(model_path=cfg.model_path)
    codec.load_pretrained()


def test_engine_invalidated_by_standard_weight_loading():
    """load_state_dict on the network or one stage, and in-place parameter edits, must drop the packed engine."""
    net = l3ac_b200.EnCodec(model_config("3kbps"), seed=1)
    sentinel = object()
    net._engine, net._engine_key = sentinel, None
    net.decoder.load_state_dict(net.decoder.state_dict())
    assert net._engine is None
    net._engine = sentinel
    net.load_state_dict(net.state_dict())
    assert net._engine is None
    v0 = sum(p._version for p in net.parameters())
    with torch.no_grad():
        next(net.quantizer.parameters()).mul_(1.0)
    assert sum(p._version for p in net.parameters()) == v0 + 1      # the engine cache key changes with it


def test_chunk_data_round_trip():
    """ChunkData, l3ac/codec.py:164-195: chunk i > 0 carries a prefix_len overlap that ``data`` drops again."""
    from l3ac_b200.network import ChunkData
    x = torch.arange(1000)
    c = ChunkData(chunk_len=300, prefix_len=50, original_data=x)
    chunks = c.chunk_data
    assert [len(t) for t in chunks] == [300, 350, 350, 150]
    assert torch.equal(chunks[1], x[250:600])
    assert torch.equal(ChunkData(chunk_len=300, prefix_len=50, chunk_data=chunks).data, x)
    with pytest.raises(AssertionError):
        ChunkData(chunk_len=10, prefix_len=10, original_data=x)


@pytest.mark.parametrize("name", CONFIGS + ("rotary",))
def test_spec_matches_reference_checkpoint_keys(name):
    """Key names and shapes equal the reference state_dicts (dumped from the reference by oracle/make_golden.py)."""
    ref = json.loads((GOLDEN / "state_dict_keys.json").read_text())[name]
    spec = network_spec(model_config(name))
    assert set(spec) == set(ref)
    for mod in spec:
        assert {k: list(v[0]) for k, v in spec[mod].items()} == ref[mod], mod


def test_config_fields_and_validation(tmp_path):
    cfg = L3ACConfig(config_file=CONFIG_DIR / "1kbps.toml")
    assert cfg.sample_rate == 16000 and cfg.model_name == "1kbps" and cfg.model_tag == "1kbps.v1"
    mc = cfg.network_config
    assert mc.feature_dim == 128 and mc.compress_rates == (6, 5, 3) and mc.hop_length == 270
    assert mc.levels == (7, 7, 7, 7, 7, 7)
    bad = tmp_path / "bad.toml"
    bad.write_text('model_tag = "x"\n[network_config]\nfeature_dim = 128\n')
    with pytest.raises(ValueError):
        L3ACConfig(config_file=bad)
    with pytest.raises(ValueError):
        ModelConfig(compress_rates=(2, 2), encoder_dims=(8, 16), encoder_depths=(1, 1))


def test_network_state_dict_round_trip(tmp_path):
    mc = model_config("3kbps")
    net = l3ac_b200.EnCodec(mc, seed=1)
    assert list(net.trainable_modules) == ["encoder", "quantizer", "decoder", "en_encoder", "en_decoder"]
    w = init_state_dicts(mc, seed=2, jitter=True)
    net.load_state_dicts(w)
    for mod, sd in w.items():
        got = getattr(net, mod).state_dict()
        assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    net.save_model(model_path=tmp_path / "m")
    net2 = l3ac_b200.EnCodec(mc, seed=5)
    net2.load_model(model_path=tmp_path / "m")
    assert torch.equal(net2.decoder.state_dict()["blocks.0.bias"], w["decoder"]["blocks.0.bias"])
    padded, n = net.preprocess(torch.zeros(2, 1000))
    assert n == 1000 and padded.shape == (2, 1056)


def test_header_symbols_are_bound_and_exported():
    """Every function the header declares has a ctypes prototype, and the built library exports it."""
    header = (ROOT / "include" / "l3ac_b200.h").read_text()
    declared = set(re.findall(r"\b(l3ac_[a-z0-9_]+)\s*\(", header)) - {"l3ac_gemm_desc"}
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    if not _lib.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    # dlopen needs libcuda only for kernel launches; resolving symbols works without a GPU
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.l3ac_abi_version() == 1


def test_gemm_desc_layout_matches_header():
    """The ctypes mirror of l3ac_gemm_desc has the C layout (11 pointers, 3 int64, 9 int32 -> 152 bytes)."""
    assert ctypes.sizeof(_lib.GemmDesc) == 11 * 8 + 3 * 8 + 9 * 4 + 4
    assert _lib.GemmDesc.lda.offset == 88 and _lib.GemmDesc.B.offset == 112 and _lib.GemmDesc.out_dtype.offset == 144


def test_no_cpu_fallback():
    """The product refuses to run without CUDA instead of silently computing on the host."""
    net = l3ac_b200.EnCodec(model_config("1kbps"))
    codec = l3ac_b200.L3AC(L3ACConfig(config_file=CONFIG_DIR / "1kbps.toml"))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            codec.encode_audio(torch.zeros(1, 16000))
        with pytest.raises(RuntimeError):
            net.engine
    src = "".join(p.read_text() for p in (ROOT / "l3ac_b200").glob("*.py"))
    assert "oracle" not in src.replace("# oracle", ""), "product code must not reference the oracle"


@pytest.mark.parametrize("name,tokens,codes,bps", [("0k75bps", 44.44, 117649, 748.6), ("1kbps", 59.26, 117649, 998.2),
                                                  ("1k5bps", 88.89, 117649, 1497.3), ("3kbps", 166.67, 250047, 2988.6)])
def test_get_model_info_matches_readme_table(name, tokens, codes, bps):
    """The model table of the reference README.md:71-76 (tokens/s, codebook size, bitrate) and SURVEY's MAC counts."""
    net = l3ac_b200.EnCodec(model_config(name))
    info = l3ac_b200.get_model_info(net)
    assert info["codebook_size"] == codes
    assert abs(info["frame_rate"] - tokens) < 0.01 and abs(info["bps"] - bps) < 0.1
    gmac = float(info["macs"].split()[0])
    want = {"0k75bps": 67.29, "1kbps": 83.39, "1k5bps": 84.30, "3kbps": 72.82}[name] / 2      # SURVEY section 8d, GFLOP = 2 * GMAC
    assert abs(gmac - want) / want < 0.03, (gmac, want)


def test_step_level_structs_match_the_header(tmp_path):
    """ctypes mirrors of l3ac_codec_config / l3ac_tensor have the layout a C compiler gives the header's structs, and the
    plain-C host (tests/c/codec_driver.c) compiles and links against the built library with gcc alone."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "l3ac_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(l3ac_codec_config), offsetof(l3ac_codec_config, en_coder_depth),\n'
                   '  offsetof(l3ac_codec_config, levels), offsetof(l3ac_codec_config, decode_rates), offsetof(l3ac_codec_config, precision),\n'
                   '  sizeof(l3ac_tensor), offsetof(l3ac_tensor, numel)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    cc, t = _lib.CodecConfig, _lib.Tensor
    assert got == [ctypes.sizeof(cc), cc.en_coder_depth.offset, cc.levels.offset, cc.decode_rates.offset, cc.precision.offset,
                   ctypes.sizeof(t), t.numel.offset]
    if not _lib.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "codec_driver.c"),
                    "-L", str(_lib.LIB_PATH.parent), "-ll3ac_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}", "-o", str(tmp_path / "driver")], check=True)
