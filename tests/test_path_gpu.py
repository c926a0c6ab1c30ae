"""GPU: the full encode -> quantize -> decode path through the public API against the oracle and golden vectors."""
import numpy as np
import pytest
import torch

import l3ac_b200
from l3ac_b200 import ops
from helpers import CONFIGS, bf16_operand_emulation, config_path, golden_case, make_audio, max_abs, model_config, snr_db
from l3ac_b200.config import CONFIG_DIR, L3ACConfig
from l3ac_b200.spec import init_state_dicts
from oracle import l3ac_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(name, weights, precision):
    codec = l3ac_b200.L3AC(L3ACConfig(config_file=config_path(name)), precision=precision)
    codec.network.load_state_dicts(weights)
    codec.network.cuda()
    codec.network.eval()
    return codec


@pytest.mark.parametrize("name", CONFIGS)
def test_fp32_mode_matches_golden(cuda_lib, name):
    """fp32 mode against the reference's golden vectors.

    Bounds (tests/golden/fp32_floor.json + fp64_wave.npz, made by tools/fp32_floor.py): the reference's own fp32 decode
    is 6.5e-5 ... 9.4e-4 max-abs (85-99 dB) away from the same forward in exact (fp64) arithmetic, and moves by up to
    2.1e-5 when torch merely uses 1 instead of 8 threads -- a flat 1e-5 is below the reference's self-noise.  So: latents
    to 2e-6, indices equal, waveform (a) within the reference's own distance to exact arithmetic and < 6e-5 absolute,
    and (b) as close to the exact waveform as the reference is (max-abs within a factor 1.25, SNR within 3 dB)."""
    import json
    from helpers import GOLDEN
    mc, weights, audio, g = golden_case(name)
    floor = json.loads((GOLDEN / "fp32_floor.json").read_text())[name]
    exact = torch.from_numpy(np.load(GOLDEN / "fp64_wave.npz")[name])
    codec = build(name, weights, "fp32")
    taps = {}
    with torch.inference_mode():
        q, idx = codec.network.engine.encode(audio.to(DEV), taps)
        wav_from_ref_idx = codec.decode_audio(indices=torch.from_numpy(g["indices"]).to(DEV))
        wav = codec.decode_audio(q)
    z_err = max_abs(taps["z"].cpu(), torch.from_numpy(g["z"]))
    agree = float((idx["indices"].cpu().numpy() == g["indices"]).mean())
    stride = int(g["wav_stride"])
    ref_wav = torch.from_numpy(g["wav"])
    ours = wav_from_ref_idx.cpu()[:, ::stride]
    wav_err = max_abs(ours, ref_wav)
    ours_vs_exact, ref_vs_exact = max_abs(ours, exact), max_abs(ref_wav, exact)
    print(f"[{name}] z max-abs {z_err:.2e}  index agreement {agree:.5f}  wav max-abs {wav_err:.2e} "
          f"snr {snr_db(ref_wav, ours):.1f} dB | vs exact arithmetic: ours {ours_vs_exact:.2e} "
          f"({snr_db(exact, ours):.1f} dB), reference {ref_vs_exact:.2e} ({snr_db(exact, ref_wav):.1f} dB)")
    assert z_err < 2e-6
    assert agree == 1.0
    assert wav_err < min(6e-5, floor["wave_max_abs_fp32_vs_fp64"])
    assert snr_db(ref_wav, ours) > 85.0
    assert ours_vs_exact < 1.25 * ref_vs_exact and snr_db(exact, ours) > snr_db(exact, ref_wav) - 3.0
    assert idx["indices"].dtype == torch.int32 and idx["level_indices"].dtype == torch.float32
    assert q.shape == (audio.shape[0], g["indices"].shape[1], 128) and wav.shape[1] == g["indices"].shape[1] * mc.hop_length
    assert max_abs(wav.cpu()[:, ::stride], ref_wav) < 6e-5


@pytest.mark.parametrize("name", CONFIGS)
def test_bf16_mode_tolerance(cuda_lib, name):
    """Default mode (bf16 tensor-core decode side, fp32 encode side): index agreement >= 99.9 %, waveform SNR stated."""
    mc, weights, audio, g = golden_case(name)
    codec = build(name, weights, "bf16")
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio.to(DEV))
        wav = codec.decode_audio(indices=torch.from_numpy(g["indices"]).to(DEV))
    agree = float((idx["indices"].cpu().numpy() == g["indices"]).mean())
    stride = int(g["wav_stride"])
    ref_wav = torch.from_numpy(g["wav"])
    snr = snr_db(ref_wav, wav.cpu()[:, ::stride])
    err = max_abs(wav.cpu()[:, ::stride], ref_wav)
    print(f"[{name}] bf16: index agreement {agree:.5f}  wav snr {snr:.1f} dB  max-abs {err:.3e}")
    assert agree >= 0.999
    assert snr > 17.0 and err < 0.3        # random-init nets are not contractive: SURVEY.md section 8d (iii)


@pytest.mark.parametrize("name", ["1kbps", "3kbps"])
def test_split_mode_is_fp32_class(cuda_lib, name):
    """precision="split": encode AND decode on the tensor cores with 3-term split-bf16 operands -- indices equal, waveform at
    the fp32-class level (the bf16 decode of the same weights sits at 19-24 dB)."""
    mc, weights, audio, g = golden_case(name)
    codec = build(name, weights, "split")
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio.to(DEV))
        wav = codec.decode_audio(indices=torch.from_numpy(g["indices"]).to(DEV))
    agree = float((idx["indices"].cpu().numpy() == g["indices"]).mean())
    stride = int(g["wav_stride"])
    ref_wav = torch.from_numpy(g["wav"])
    snr = snr_db(ref_wav, wav.cpu()[:, ::stride])
    err = max_abs(wav.cpu()[:, ::stride], ref_wav)
    print(f"[{name}] split: index agreement {agree:.5f}  wav snr {snr:.1f} dB  max-abs {err:.3e}")
    assert agree >= 0.999
    assert snr > 65.0 and err < 1e-3


def test_api_surface_and_invariants(cuda_lib):
    name = "1k5bps"
    mc = model_config(name)
    weights = init_state_dicts(mc, seed=11, jitter=True)
    codec = build(name, weights, "fp32")
    audio = make_audio(3, 1.7, seed=9).to(DEV)
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio)
        a = codec.decode_audio(q)                       # positional q_feature, README.md:61
        b = codec.decode_audio(indices=idx["indices"])
        c = codec.decode_audio(indices=idx["indices"].long())
        assert torch.equal(a, b) and torch.equal(b, c)  # decode(q_feature) == decode(indices=) bit-exact
        q1, idx1 = codec.encode_audio(audio[1:2])
        assert torch.equal(idx1["indices"], idx["indices"][1:2])    # batch independence
        out = codec.network(audio)
        assert torch.equal(out["indices"], idx["indices"]) and out["generated_audio"].shape == audio.shape
    with pytest.raises(RuntimeError):
        codec.encode_audio(audio[0])                    # 1-D input, like the reference's conv shape error
    with pytest.raises(AttributeError):
        codec.decode_audio()
    # oracle agreement on a ragged length (not a multiple of hop) and an empty-ish clip
    orc = O.Oracle(mc.as_dict(), weights)
    oq, oidx = orc.encode_audio(audio.cpu())
    assert (oidx["indices"] == idx["indices"].cpu()).float().mean() >= 0.999
    assert max_abs(orc.decode_audio(indices=idx["indices"].cpu()), b.cpu()) < 1e-4
    with torch.inference_mode():                        # empty batch: empty outputs with the right trailing shapes
        qe, ide = codec.encode_audio(audio[:0])
        we = codec.decode_audio(indices=ide["indices"])
    assert qe.shape == (0, q.shape[1], 128) and ide["indices"].shape == (0, q.shape[1]) and we.shape == (0, b.shape[1])
    short = make_audio(1, 0.01, seed=2).to(DEV)       # 160 samples -> one hop
    with torch.inference_mode():
        qs, ids = codec.encode_audio(short)
    assert ids["indices"].shape == (1, 1)
    assert torch.equal(ids["indices"].cpu(), orc.encode_audio(short.cpu())[1]["indices"])


def test_round_trip_at_bench_size(cuda_lib):
    """BASELINE config #2 shape (1kbps, 10 s clips, a micro-batch of it): properties that need no oracle run."""
    codec = l3ac_b200.get_model("1kbps", pretrained=False)
    codec.network.cuda()
    audio = make_audio(8, 10.0, seed=3).to(DEV)
    audio[4:] = audio[:4]                                   # duplicate clips must give identical results
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio)
        wav = codec.decode_audio(indices=idx["indices"])
    assert idx["indices"].shape == (8, 593) and wav.shape == (8, 160110)
    assert int(idx["indices"].min()) >= 0 and int(idx["indices"].max()) < 7 ** 6
    assert torch.equal(idx["indices"][4:], idx["indices"][:4]) and torch.equal(wav[4:], wav[:4])
    assert torch.isfinite(wav).all() and float(wav.abs().max()) <= 1.0
    lv = idx["level_indices"].long()
    basis = torch.tensor([1, 7, 49, 343, 2401, 16807], device=DEV)
    assert torch.equal((lv * basis).sum(-1).int(), idx["indices"])   # mixed-radix checksum of the level digits


@pytest.mark.parametrize("name,seconds,batch,t_tok,t_pad", [("0k75bps", 10.0, 6, 445, 160200), ("1k5bps", 10.0, 6, 889, 160020),
                                                             ("3kbps", 30.0, 3, 5000, 480000), ("1kbps", 60.0, 2, 3556, 960120)])
def test_other_configs_at_baseline_sizes(cuda_lib, name, seconds, batch, t_tok, t_pad):
    """BASELINE configs #3-#5 (other hop / token rates, 30 s and 60 s clips, decode from indices) at their full clip lengths:
    size-independent properties, plus index agreement and decode SNR of one clip against the oracle on the same weights."""
    mc = model_config(name)
    weights = init_state_dicts(mc, seed=11, jitter=True)
    codec = build(name, weights, "bf16")
    audio = make_audio(batch, seconds, seed=77).to(DEV)
    audio[-1] = audio[0]                                       # duplicate clips must give identical results at any batch position
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio)
        wav_i = codec.decode_audio(indices=idx["indices"])     # config #5: decode from indices only
        wav_q = codec.decode_audio(q)
        q1, idx1 = codec.encode_audio(audio[:1])               # batch independence (a different micro-batch / graph path)
    levels = torch.tensor(mc.levels, device=DEV)
    basis = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.long, device=DEV), levels[:-1]]), 0)
    assert idx["indices"].shape == (batch, t_tok) and wav_i.shape == (batch, t_pad)
    assert int(idx["indices"].min()) >= 0 and int(idx["indices"].max()) < int(torch.prod(levels))
    assert torch.equal((idx["level_indices"].long() * basis).sum(-1).int(), idx["indices"])     # mixed-radix checksum
    assert torch.equal(wav_i, wav_q)                           # decode(q_feature) == decode(indices=) bit for bit
    assert torch.equal(idx["indices"][-1], idx["indices"][0]) and torch.equal(wav_i[-1], wav_i[0])
    assert torch.equal(idx1["indices"][0], idx["indices"][0])
    assert torch.isfinite(wav_i).all() and float(wav_i.abs().max()) <= 1.0
    if seconds <= 30.0:                                        # the oracle needs ~1.5 s of CPU per 10 s clip
        orc = O.Oracle(mc.as_dict(), weights)
        _, oidx = orc.encode_audio(audio[:1].cpu())
        agree = float((oidx["indices"] == idx["indices"][:1].cpu()).float().mean())
        owav = orc.decode_audio(indices=idx["indices"][:1].cpu())
        snr = snr_db(owav, wav_i[:1].cpu())
        codec32 = build(name, weights, "fp32")
        with torch.inference_mode():
            wav32 = codec32.decode_audio(indices=idx["indices"][:1])
        snr32 = snr_db(owav, wav32.cpu())
        with bf16_operand_emulation():
            ewav = orc.decode_audio(indices=idx["indices"][:1].cpu())
        snr_emu = snr_db(owav, ewav)
        print(f"[{name} {seconds:g} s] index agreement {agree:.5f}, decode SNR vs oracle: fp32 mode {snr32:.1f} dB, bf16 mode {snr:.1f} dB, "
              f"bf16-operand emulation of the oracle {snr_emu:.1f} dB")
        assert agree >= 0.999           # north-star bound for the default (split-bf16 encode) mode
        assert snr32 > 60.0             # fp32 mode: the 1e-5 class (measured 74-91 dB)
        # bf16 decode side: these jittered random-init networks amplify operand rounding by 6-10 dB per decoder stage
        # (tools/snr_diag.py: dec_feature 47 dB -> dec_up3 15 dB -> waveform 6-19 dB depending on config and seed).  The bound is
        # therefore relative: the kernels must be as accurate as the "bf16-in / fp32-accumulate" model of the same network
        # (the oracle with its dense operands rounded to bf16), within 4 dB; the absolute tolerance on the golden weights is
        # stated in test_bf16_mode_tolerance.
        assert snr > snr_emu - 4.0 and snr > 3.0


def test_many_short_clips(cuda_lib):
    """BASELINE config #5 corner: hundreds of 1 s clips (two micro-batches, sample count not a multiple of any tile size)."""
    codec = l3ac_b200.get_model("1kbps", pretrained=False)
    codec.network.cuda()
    audio = make_audio(300, 1.0, seed=9).to(DEV)
    audio[299] = audio[0]
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio)
        wav = codec.decode_audio(indices=idx["indices"])
        q1, idx1 = codec.encode_audio(audio[:1])
        wav1 = codec.decode_audio(indices=idx1["indices"])
    assert idx["indices"].shape == (300, 60) and wav.shape == (300, 16200)
    assert torch.equal(idx["indices"][299], idx["indices"][0]) and torch.equal(wav[299], wav[0])
    assert torch.equal(idx1["indices"][0], idx["indices"][0]) and torch.equal(wav1[0], wav[0])
    assert torch.isfinite(wav).all()


def test_micro_batching_is_transparent(cuda_lib):
    """Batches larger than one 160 s micro-batch are processed in chunks; results must equal per-clip processing."""
    codec = l3ac_b200.get_model("3kbps", pretrained=False)
    codec.network.cuda()
    codec.network.engine.max_chunk_samples = 16000 * 12          # force 3 chunks for 7 clips of 5 s (issued on side streams)
    codec.network.engine.graph_max_samples = 0
    audio = make_audio(7, 5.0, seed=21).to(DEV)
    with torch.inference_mode():
        q, idx = codec.encode_audio(audio)
        wav = codec.decode_audio(indices=idx["indices"])
        q1, idx1 = codec.encode_audio(audio[5:6])
        wav1 = codec.decode_audio(indices=idx1["indices"])
    assert idx["indices"].shape[0] == 7 and wav.shape[0] == 7
    assert torch.equal(idx["indices"][5:6], idx1["indices"]) and torch.equal(q[5:6], q1)
    assert torch.equal(wav[5:6], wav1)


def test_pinned_host_input_is_uploaded_per_micro_batch(cuda_lib):
    """A pinned host batch spanning several micro-batches is uploaded chunk by chunk inside encode_audio; same results."""
    codec = l3ac_b200.get_model("1kbps", pretrained=False)
    codec.network.cuda()
    codec.network.engine.max_chunk_samples = 16000 * 12
    codec.network.engine.graph_max_samples = 0
    host = make_audio(5, 5.0, seed=41).pin_memory()
    with torch.inference_mode():
        q0, idx0 = codec.encode_audio(host.to(DEV))
        q1, idx1 = codec.encode_audio(host)
    assert q1.device.type == "cuda" and torch.equal(idx0["indices"], idx1["indices"]) and torch.equal(q0, q1)
    # decode_audio(out=pinned host tensor): the waveform is downloaded per micro-batch inside the call
    with torch.inference_mode():
        wav = codec.decode_audio(indices=idx0["indices"])
        host = torch.empty(tuple(wav.shape), dtype=torch.float32).pin_memory()
        got = codec.decode_audio(indices=idx0["indices"], out=host)
        small = torch.empty((1, wav.shape[1]), dtype=torch.float32).pin_memory()
        codec.network.engine.graph_max_samples = 16000 * 40
        got1 = codec.decode_audio(indices=idx0["indices"][:1], out=small)            # CUDA-graph path
    assert got is host and torch.equal(host, wav.cpu()) and got1 is small and torch.equal(small, wav[:1].cpu())
    with pytest.raises(ValueError):
        codec.decode_audio(indices=idx0["indices"], out=torch.empty(tuple(wav.shape)))   # not pinned


@pytest.mark.parametrize("engine_kind", ["python", "native"])
def test_micro_batch_graphs_match_eager(cuda_lib, engine_kind, monkeypatch):
    """Large batches on the operator-level (Python) path: each micro-batch is captured into its own CUDA graph on its second
    appearance and replayed on its stream (pinned host input uploaded straight into the graph's input); results must be
    bit-identical to the eager launch sequence.  With the step-level C ABI underneath (native) the micro-batches are launched
    directly -- no graphs -- and the same invariants hold."""
    monkeypatch.setenv("L3AC_ENGINE", engine_kind)
    codec = l3ac_b200.get_model("1kbps", pretrained=False)
    codec.network.cuda()
    eng = codec.network.engine
    assert (eng.native is not None) == (engine_kind == "native")
    eng.max_chunk_samples = 16000 * 12                            # 4 micro-batches for 7 clips of 5 s
    eng.graph_max_samples = 0
    a, b = make_audio(7, 5.0, seed=51), make_audio(7, 5.0, seed=52).pin_memory()
    with torch.inference_mode():
        eng.graph_chunks = False
        ref = []
        launches0 = ops.LAUNCHES
        for x in (a.to(DEV), b):
            q, idx = codec.encode_audio(x)
            ref.append((q, idx["indices"], codec.decode_audio(indices=idx["indices"])))
        per_call = (ops.LAUNCHES - launches0) // 2                # kernels of one encode + decode of this batch
        eng.graph_chunks = True
        launches0 = ops.LAUNCHES
        for rep in range(3):                                      # (eager, capture + replay), replay, replay
            for x, (q0, i0, w0) in zip((a.to(DEV), b), ref):
                q, idx = codec.encode_audio(x)
                w = codec.decode_audio(indices=idx["indices"])
                assert torch.equal(q, q0) and torch.equal(idx["indices"], i0) and torch.equal(w, w0), rep
        if engine_kind == "python":
            assert len(eng._graphs) == 2 * len(eng._chunks(7, 80000)) >= 6      # micro-batch slots x (encode, decode)
            # replayed kernels are counted like launched ones: 1 eager call + the capture's warm-up run + 5 replays (the few
            # kernels outside the graphs -- dequantize -- are not repeated by the warm-up run)
            assert abs((ops.LAUNCHES - launches0) - 7 * per_call) <= 8
        else:
            assert len(eng._graphs) == 0 and ops.LAUNCHES - launches0 == 6 * per_call


def test_cuda_graph_path_matches_eager(cuda_lib):
    """Small batches replay a captured CUDA graph; results must be bit-identical to the eager launch sequence."""
    codec = l3ac_b200.get_model("1kbps", pretrained=False)
    codec.network.cuda()
    eng = codec.network.engine
    audio = make_audio(2, 3.0, seed=31).to(DEV)
    with torch.inference_mode():
        eng.graph_max_samples = 0                       # eager
        q0, idx0 = codec.encode_audio(audio)
        w0 = codec.decode_audio(indices=idx0["indices"])
        eng.graph_max_samples = 16000 * 40              # graphed: first call captures, second replays
        for _ in range(2):
            q1, idx1 = codec.encode_audio(audio)
            w1 = codec.decode_audio(indices=idx1["indices"])
        other = make_audio(2, 3.0, seed=32).to(DEV)     # same shape, different data -> same graph, new results
        q2, idx2 = codec.encode_audio(other)
        eng.graph_max_samples = 0
        q3, idx3 = codec.encode_audio(other)
    assert torch.equal(idx0["indices"], idx1["indices"]) and torch.equal(q0, q1) and torch.equal(w0, w1)
    assert torch.equal(idx2["indices"], idx3["indices"]) and torch.equal(q2, q3)
    assert len(eng._graphs) == 2


@pytest.mark.parametrize("precision", ["fp32", "bf16", "split"])
def test_rotary_config_matches_golden(cuda_lib, precision):
    """en_coder_dynamic_pos = false (rotary positions inside LocalMHA, l3ac/local_trans.py:29,36) on a
    user-defined config, against golden vectors produced by the unmodified reference (tests/golden/rotary.toml / .npz)."""
    mc, weights, audio, g = golden_case("rotary")
    codec = build("rotary", weights, precision)
    taps = {}
    with torch.inference_mode():
        q, idx = codec.network.engine.encode(audio.to(DEV), taps)
        wav = codec.decode_audio(indices=torch.from_numpy(g["indices"]).to(DEV))
    agree = float((idx["indices"].cpu().numpy() == g["indices"]).mean())
    z_err = max_abs(taps["z"].cpu(), torch.from_numpy(g["z"]))
    stride = int(g["wav_stride"])
    ref_wav = torch.from_numpy(g["wav"])
    snr = snr_db(ref_wav, wav.cpu()[:, ::stride])
    print(f"[rotary {precision}] z max-abs {z_err:.2e}  index agreement {agree:.5f}  wav snr {snr:.1f} dB")
    assert agree >= 0.999
    if precision == "fp32":
        assert z_err < 5e-6 and agree == 1.0 and snr > 80.0
    elif precision == "split":
        assert snr > 55.0
    else:
        assert snr > 10.0


def test_forward_dict_and_stage_modules(cuda_lib):
    """EnCodec.forward (l3ac/en_codec.py:53-72) returns the reference's full dict, and the five trainable modules are
    callable with the reference's channels-first boundaries; all against the oracle on the same weights (fp32 mode)."""
    name = "1k5bps"
    mc = model_config(name)
    weights = init_state_dicts(mc, seed=13, jitter=True)
    codec = build(name, weights, "fp32")
    net = codec.network
    audio = make_audio(2, 1.3, seed=4)
    cfg = mc.as_dict()
    w = {m: {k: v.float() for k, v in sd.items()} for m, sd in weights.items()}
    padded, n = O.preprocess(cfg, audio)
    feature = O.encoder(w["encoder"], cfg, padded.unsqueeze(1))
    trans = O.en_encoder(w["en_encoder"], cfg, feature)
    oq, oidx, _ = O.quantizer_forward(w["quantizer"], cfg, trans)
    q_feature = O.en_decoder(w["en_decoder"], cfg, oq)
    y = O.decoder(w["decoder"], cfg, q_feature)
    with torch.inference_mode():
        out = net(audio.to(DEV))
        hid = out["hidden_feature"]
        assert set(out) == {"generated_audio", "embedded_audio", "indices", "commit_loss", "hidden_feature"}
        assert set(hid) == {"encoded_feature", "encoded_trans_feature", "quantized_trans_feature", "quantized_feature"}
        assert out["generated_audio"].shape == audio.shape and float(out["commit_loss"].sum()) == 0.0
        assert torch.equal(out["indices"].cpu(), oidx["indices"])
        assert hid["encoded_feature"].shape == feature.shape and out["embedded_audio"].shape == q_feature.shape
        assert max_abs(hid["encoded_feature"].cpu(), feature) < 2e-5
        assert max_abs(hid["encoded_trans_feature"].cpu(), trans) < 2e-5
        assert max_abs(hid["quantized_trans_feature"].cpu(), oq) < 1e-6
        assert max_abs(hid["quantized_feature"].cpu(), q_feature) < 2e-5 and torch.equal(out["embedded_audio"], hid["quantized_feature"])
        assert max_abs(out["generated_audio"].cpu(), y[:, 0, :n]) < 6e-5
        # callable sub-modules, reference layouts
        f = net.encoder(padded.unsqueeze(1).to(DEV))
        assert f.shape == feature.shape and max_abs(f.cpu(), feature) < 2e-5
        t = net.en_encoder(feature.to(DEV))
        assert max_abs(t.cpu(), trans) < 2e-5
        qf, qi, loss = net.quantizer(trans.to(DEV))
        assert torch.equal(qi["indices"].cpu(), oidx["indices"]) and max_abs(qf.cpu(), oq) < 1e-6
        d = net.en_decoder(oq.to(DEV))
        assert d.shape == q_feature.shape and max_abs(d.cpu(), q_feature) < 2e-5
        wv = net.decoder(q_feature.to(DEV))
        assert wv.shape == y.shape and max_abs(wv.cpu(), y) < 6e-5
        # a conv-encoder input that is not a multiple of the strides is floored like the reference's strided convs
        odd = padded[:, :padded.shape[1] - 7].unsqueeze(1)
        fo = net.encoder(odd.to(DEV))
        fo_ref = O.encoder(w["encoder"], cfg, odd)
        assert fo.shape == fo_ref.shape and max_abs(fo.cpu(), fo_ref) < 2e-5


def test_chunked_unit_api(cuda_lib):
    """extract_unit / decode_unit / ChunkData (l3ac/codec.py:122-195): one long clip through the base Codec's chunked
    compress (encoder -> quantizer) / decompress (decoder) path, against the same algorithm run on the oracle's stages."""
    from l3ac_b200.network import ChunkData
    name = "3kbps"
    mc = model_config(name)
    weights = init_state_dicts(mc, seed=17, jitter=True)
    codec = build(name, weights, "fp32")
    net = codec.network
    audio = make_audio(1, 2.6, seed=6)
    window = 16000
    cfg = mc.as_dict()
    w = {m: {k: v.float() for k, v in sd.items()} for m, sd in weights.items()}
    hop = mc.hop_length
    padded, n = O.preprocess(cfg, audio)
    win = window // hop * hop
    ref_chunks = ChunkData(chunk_len=win, prefix_len=hop, original_data=padded[0]).chunk_data
    ref_idx, ref_wav = [], []
    for x in ref_chunks:
        feat = O.encoder(w["encoder"], cfg, x[None, None, :]).permute(0, 2, 1)
        qf, idx, _ = O.quantizer_forward(w["quantizer"], cfg, feat)
        ref_idx.append(idx["indices"][0])
        ref_wav.append(O.decoder(w["decoder"], cfg, qf.permute(0, 2, 1))[0, 0])
    ref_audio = ChunkData(chunk_len=len(ref_wav[0]), prefix_len=hop, chunk_data=ref_wav).data[None, :]
    with torch.inference_mode():
        ci, cq = net.extract_unit(audio.to(DEV), process_window=window)
        wav_i = net.decode_unit(chunk_indices=ci)
        wav_q = net.decode_unit(chunk_q_feature=cq)
    assert len(ci.chunk_data) == len(ref_chunks) == 3
    for a, b in zip(ci.chunk_data, ref_idx):
        assert torch.equal(a.cpu(), b)
    assert ci.chunk_len == win // hop and ci.prefix_len == 1
    assert wav_i.shape == ref_audio.shape == (1, padded.shape[1]) and torch.equal(wav_i, wav_q)
    assert max_abs(wav_i.cpu(), ref_audio) < 6e-5
    assert torch.equal(ci.data.cpu(), torch.cat([ref_idx[0]] + [r[1:] for r in ref_idx[1:]]))
