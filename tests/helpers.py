"""Shared test utilities (seeded inputs, golden cases, error metrics)."""
import hashlib
from pathlib import Path

import numpy as np
import torch

from l3ac_b200.config import CONFIG_DIR, L3ACConfig
from l3ac_b200.spec import init_state_dicts

GOLDEN = Path(__file__).resolve().parent / "golden"
CONFIGS = ("0k75bps", "1kbps", "1k5bps", "3kbps")


def make_audio(batch: int, seconds: float, seed: int = 1234) -> torch.Tensor:
    """0.1 * randn clipped to [-1, 1], 16 kHz mono (same generator as oracle/make_golden.py)."""
    g = torch.Generator().manual_seed(seed)
    n = int(round(seconds * 16000))
    return (0.1 * torch.randn(batch, n, generator=g)).clamp(-1, 1)


def config_path(name: str) -> Path:
    """Named configs live in the package; test-only configs (``rotary``) next to the golden vectors."""
    p = CONFIG_DIR / f"{name}.toml"
    return p if p.exists() else GOLDEN / f"{name}.toml"


def model_config(name: str):
    return L3ACConfig(config_file=config_path(name)).network_config


def weights_digest(weights) -> str:
    h = hashlib.sha256()
    for mod in weights:
        for k, v in weights[mod].items():
            h.update(k.encode())
            h.update(v.contiguous().numpy().tobytes())
    return h.hexdigest()


def golden_case(name: str):
    """Returns (model_config, weights, audio, golden npz) and checks the regenerated weights are the committed ones."""
    g = np.load(GOLDEN / f"{name}.npz")
    mc = model_config(name)
    weights = init_state_dicts(mc, seed=int(g["weight_seed"]), jitter=True)
    assert weights_digest(weights) == str(g["weights_sha256"]), "seeded weights differ from the golden run (torch RNG drift?)"
    audio = make_audio(int(g["batch"]), float(g["seconds"]), int(g["audio_seed"]))
    return mc, weights, audio, g


def snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    ref, test = ref.double().flatten(), test.double().flatten()
    noise = (ref - test).pow(2).sum().clamp_min(1e-300)
    return float(10 * torch.log10(ref.pow(2).sum() / noise))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())


class bf16_operand_emulation:
    """Context manager: the oracle's dense decode-side ops (conv1d with groups == 1 and >= 24 input channels, linear with
    >= 24 input features) see their input and weight rounded to bf16, products accumulated in fp32 -- the arithmetic model
    of the tensor-core decode path ("bf16-in / fp32-accumulate").  Depthwise, 1-channel and quantiser layers stay fp32, as
    they do in the kernels.  Attention products are not rounded, so the emulation is, if anything, slightly MORE accurate
    than the real path."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.conv1d, self.linear = F, F.conv1d, F.linear
        bf = lambda t: t.to(torch.bfloat16).to(torch.float32)

        def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if groups == 1 and w.shape[1] >= 24 and w.shape[0] > 1:
                x, w = bf(x), bf(w)
            return self.conv1d(x, w, b, stride, padding, dilation, groups)

        def linear(x, w, b=None):
            if w.shape[1] >= 24:
                x, w = bf(x), bf(w)
            return self.linear(x, w, b)

        F.conv1d, F.linear = conv1d, linear
        return self

    def __exit__(self, *exc):
        self.F.conv1d, self.F.linear = self.conv1d, self.linear
        return False
