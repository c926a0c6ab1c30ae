"""Decode-side error diagnosis: per-tap SNR of the bf16 and fp32 engines against the oracle for a config / weight seed."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import l3ac_b200  # noqa: E402
from helpers import make_audio, model_config, snr_db  # noqa: E402
from l3ac_b200.config import CONFIG_DIR, L3ACConfig  # noqa: E402
from l3ac_b200.spec import init_state_dicts  # noqa: E402
from oracle import l3ac_oracle as O  # noqa: E402

name, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
mc = model_config(name)
weights = init_state_dicts(mc, seed=seed, jitter=True)
audio = make_audio(1, seconds, seed=77)
orc = O.Oracle(mc.as_dict(), weights)
_, oidx = orc.encode_audio(audio)
otaps = {}
owav = orc.decode_audio(indices=oidx["indices"], taps=otaps)
for prec in ("fp32", "bf16"):
    codec = l3ac_b200.L3AC(L3ACConfig(config_file=CONFIG_DIR / f"{name}.toml"), precision=prec)
    codec.network.load_state_dicts(weights)
    codec.network.cuda()
    taps = {}
    with torch.inference_mode():
        wav = codec.network.engine.decode(indices=oidx["indices"].cuda(), taps=taps)
    msg = [f"{name} seed {seed} {prec}: wav {snr_db(owav, wav.cpu()):.1f} dB"]
    for k in ("dec_feature", "dec_up0", "dec_up1", "dec_up2", "dec_up3"):
        if k in taps and k in otaps:
            a, b = otaps[k], taps[k].cpu()
            if a.shape != b.shape:
                b = b.transpose(1, 2)
            msg.append(f"{k} {snr_db(a, b):.1f}")
    print(" | ".join(msg), flush=True)
