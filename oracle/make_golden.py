"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference, with
oracle/shim/local_attention standing in for the un-vendored PyPI dependency) on seeded weights and inputs, and
asserts that oracle/l3ac_oracle.py reproduces it bit for bit.  Run in the build container only:

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The weights come from l3ac_b200.spec.init_state_dicts (seeded, jittered so that GRN,
biases and affine terms are non-trivial) and are loaded into the reference with strict=True, which also proves
that the product's parameter inventory is the reference's checkpoint format.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "shim"))
sys.path.insert(0, "/root/reference")

import l3ac as ref_pkg                                          # noqa: E402  (the reference)
from l3ac_b200.config import CONFIG_DIR, L3ACConfig            # noqa: E402
from l3ac_b200.spec import init_state_dicts, network_spec      # noqa: E402
from oracle.l3ac_oracle import Oracle                           # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
# (config, seconds, batch): 10 s clips span several attention windows (1kbps: 750 frames = 4.2 s)
CASES = [("1kbps", 10.0, 1), ("3kbps", 6.0, 1), ("0k75bps", 3.1, 2), ("1k5bps", 3.1, 2),
         ("rotary", 3.1, 2)]          # tests/golden/rotary.toml: the rotary-position path (no named config uses it)


def config_file(name: str, config_dir) -> Path:
    """Named configs live in the package's configs/; test-only configs (rotary) next to the golden vectors."""
    p = Path(config_dir) / f"{name}.toml"
    return p if p.exists() else GOLDEN / f"{name}.toml"

WEIGHT_SEED, AUDIO_SEED = 7, 1234


def make_audio(batch: int, seconds: float, seed: int = AUDIO_SEED) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    n = int(round(seconds * 16000))
    return (0.1 * torch.randn(batch, n, generator=g)).clamp(-1, 1)


def weights_digest(weights) -> str:
    h = hashlib.sha256()
    for mod in weights:
        for k, v in weights[mod].items():
            h.update(k.encode())
            h.update(v.contiguous().numpy().tobytes())
    return h.hexdigest()


def stats(t: torch.Tensor):
    t = t.double()
    return [float(t.sum()), float((t * t).sum()), float(t.abs().max())]


def main(only=None):
    torch.set_num_threads(8)
    keys_path = GOLDEN / "state_dict_keys.json"
    keys = json.loads(keys_path.read_text()) if (only and keys_path.exists()) else {}
    for name, seconds, batch in CASES:
        if only and name not in only:
            continue
        rcfg = ref_pkg.L3ACConfig(config_file=config_file(name, ref_pkg.CONFIG_DIR))
        ref = ref_pkg.L3AC(rcfg)
        ref.network.eval()
        mc = L3ACConfig(config_file=config_file(name, CONFIG_DIR)).network_config
        spec = network_spec(mc)
        for mod, net in ref.network.trainable_modules.items():
            sd = net.state_dict()
            assert list(sd) == list(spec[mod]), f"{name}.{mod}: key mismatch"
            assert all(tuple(v.shape) == spec[mod][k][0] for k, v in sd.items())
        keys[name] = {mod: {k: list(s[0]) for k, s in spec[mod].items()} for mod in spec}
        weights = init_state_dicts(mc, seed=WEIGHT_SEED, jitter=True)
        for mod, net in ref.network.trainable_modules.items():
            net.load_state_dict(weights[mod], strict=True)
        audio = make_audio(batch, seconds)
        hooks, taps = [], {}
        hooks.append(ref.network.encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("enc_feature", o)))
        hooks.append(ref.network.en_encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("trans_feature", o)))
        hooks.append(ref.network.quantizer.project_in.register_forward_hook(lambda m, i, o: taps.__setitem__("z", o)))
        hooks.append(ref.network.en_decoder.register_forward_hook(lambda m, i, o: taps.__setitem__("dec_feature", o)))
        with torch.inference_mode():
            q, idx = ref.encode_audio(audio)
            wav = ref.decode_audio(indices=idx["indices"])
            wav_q = ref.decode_audio(q)
        for h in hooks:
            h.remove()
        assert torch.equal(wav, wav_q)
        # the oracle must reproduce the reference exactly
        orc = Oracle(mc.as_dict(), weights)
        otaps = {}
        oq, oidx = orc.encode_audio(audio, otaps)
        owav = orc.decode_audio(indices=oidx["indices"], taps=otaps)
        assert torch.equal(oidx["indices"], idx["indices"]) and torch.equal(oidx["level_indices"], idx["level_indices"])
        assert torch.equal(oq, q) and torch.equal(owav, wav)
        for k in ("enc_feature", "trans_feature", "z", "dec_feature"):
            assert torch.equal(otaps[k], taps[k]), k
        np.savez_compressed(
            GOLDEN / f"{name}.npz",
            seconds=np.float64(seconds), batch=np.int64(batch), weight_seed=np.int64(WEIGHT_SEED),
            audio_seed=np.int64(AUDIO_SEED), weights_sha256=np.array(weights_digest(weights)),
            audio_stats=np.array(stats(audio)),
            indices=idx["indices"].numpy(), level_indices=idx["level_indices"].numpy().astype(np.int8),
            z=taps["z"].numpy(), q_feature_stats=np.array(stats(q)),
            enc_feature_stats=np.array(stats(taps["enc_feature"])), trans_feature_stats=np.array(stats(taps["trans_feature"])),
            dec_feature_stats=np.array(stats(taps["dec_feature"])),
            wav=(wav.numpy() if name == "1kbps" else wav.numpy()[:, ::8]), wav_stride=np.int64(1 if name == "1kbps" else 8),
            wav_stats=np.array(stats(wav)))
        print(name, "T_tok", tuple(idx["indices"].shape), "wav", tuple(wav.shape), "ok")
    (GOLDEN / "state_dict_keys.json").write_text(json.dumps(keys, indent=0, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
