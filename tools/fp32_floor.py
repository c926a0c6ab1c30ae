"""Rounding floor of the reference's own fp32 decode: oracle (== reference, bit for bit) in fp32 against the SAME forward
evaluated in fp64 on the same weights and indices.  Any independent fp32 implementation (other summation order, other
libm) sits at this distance from the reference; it is the bound a 'pure-fp32 parity mode' can be held to.
Test infrastructure (imports oracle/).  Usage: python tools/fp32_floor.py [config ...]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from l3ac_b200.config import CONFIG_DIR, L3ACConfig          # noqa: E402
from l3ac_b200.spec import init_state_dicts                   # noqa: E402
from oracle import l3ac_oracle as O                           # noqa: E402


def main(names):
    out, waves = {}, {}
    for name in names:
        g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
        mc = L3ACConfig(config_file=CONFIG_DIR / f"{name}.toml").network_config
        w = init_state_dicts(mc, seed=int(g["weight_seed"]), jitter=True)
        cfg = mc.as_dict()
        idx = torch.from_numpy(g["indices"]).to(torch.int32)
        orc = O.Oracle(cfg, w)
        t32 = {}
        wav32 = orc.decode_audio(indices=idx, taps=t32)
        w64 = {m: {k: v.double() for k, v in sd.items()} for m, sd in w.items()}
        codes = O.fsq_indices_to_codes(idx, cfg["vq_config"]["levels"]).double()
        feat = torch.nn.functional.linear(codes, w64["quantizer"]["project_out.weight"], w64["quantizer"]["project_out.bias"])
        dec64 = O.en_decoder(w64["en_decoder"], cfg, feat)
        t64 = {}
        wav64 = O.decoder(w64["decoder"], cfg, dec64, t64).squeeze(1)
        d = (wav32.double() - wav64).abs()
        snr = 10 * torch.log10(wav64.pow(2).sum() / (wav32.double() - wav64).pow(2).sum())
        stage = {k: float((t32[k].double() - t64[k]).abs().max() / t64[k].abs().max()) for k in t64}
        torch.set_num_threads(1)
        wav32_1t = orc.decode_audio(indices=idx)
        torch.set_num_threads(8)
        out[name] = dict(wave_max_abs_fp32_vs_fp64=float(d.max()), wave_snr_db=float(snr),
                         rel_max_err_per_stage=stage,
                         reorder_max_abs_1_vs_8_threads=float((wav32 - wav32_1t).abs().max()))
        print(name, json.dumps(out[name]))
        stride = int(g["wav_stride"])
        waves[name] = wav64[:, ::stride].to(torch.float32).numpy()     # exact-arithmetic waveform (fp32 storage: 6e-8 rounding)
    np.savez_compressed(ROOT / "tests" / "golden" / "fp64_wave.npz", **waves)
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    res = main(sys.argv[1:] or ["1kbps", "3kbps", "0k75bps", "1k5bps"])
    (ROOT / "profiles" / "r02_fp32_floor.json").write_text(json.dumps(res, indent=1))
    (ROOT / "tests" / "golden" / "fp32_floor.json").write_text(json.dumps(res, indent=1))
