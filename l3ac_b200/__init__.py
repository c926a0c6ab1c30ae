"""l3ac_b200 -- B200-native drop-in for the L3AC encode / quantize / decode hot path.

Public surface mirrors the reference package (l3ac/__init__.py):
``list_models()``, ``get_model(config_name)``, ``L3ACConfig``, ``L3AC`` with ``config``, ``network``,
``encode_audio(audio)`` and ``decode_audio(audio_feature=None, indices=None)``.
"""
from __future__ import annotations

import logging
from pathlib import Path

import torch

from .config import CONFIG_DIR, L3ACConfig, ModelConfig
from .network import EnCodec

__version__ = "0.1"
log = logging.getLogger("L3AC")


def list_models() -> list[str]:
    """l3ac/__init__.py:17-18: the TOML stems under configs/ (``debug`` included, like the reference)."""
    return sorted(p.relative_to(CONFIG_DIR).stem for p in CONFIG_DIR.rglob("*.toml"))


def get_model(config_name, pretrained: bool = True, precision: str = "bf16") -> "L3AC":
    """l3ac/__init__.py:21-25.  ``pretrained=True`` (the reference's only behaviour) downloads the per-module ``.pt`` files
    that are missing under ``<model_dir>/<name>.<version>/`` and loads them; it RAISES when they cannot be had (no silent
    random-weight codec).  ``pretrained=False`` (extension) keeps the seeded random initialisation."""
    config = L3ACConfig(config_file=CONFIG_DIR / f"{config_name}.toml")
    codec = L3AC(config, precision=precision)
    if pretrained:
        codec.load_pretrained()
    return codec


def get_model_info(model: "EnCodec", eval_flops_seconds=10, sample_rate: int = 16000) -> dict:
    """l3ac/__init__.py:28-51 without the ``ptflops`` dependency: the same dictionary keys, with MACs counted
    analytically (conv / linear layers and the useful, unmasked attention products) for ``eval_flops_seconds`` of audio."""
    import math
    from .spec import HEADS, is_compressed, network_spec
    mc = model.mc
    T = math.ceil(eval_flops_seconds * sample_rate / mc.hop_length) * mc.hop_length
    macs = 0
    # encoder
    t = T
    macs += t * (5 * 4 * 7 + 20 * 80 + 81 * mc.encoder_dims[0])
    unit = lambda c, n: n * (7 * c + 8 * c * c)
    for i, s in enumerate(mc.compress_rates):
        macs += mc.encoder_depths[i] * unit(mc.encoder_dims[i], t)
        t //= s
        macs += t * mc.encoder_dims[i] * s * mc.encoder_dims[i + 1]
    macs += mc.encoder_depths[-1] * unit(mc.encoder_dims[-1], t) + t * 3 * mc.encoder_dims[-1] * mc.feature_dim
    t_f, d, inner, ff = t, mc.feature_dim, HEADS * (mc.feature_dim // 4), int(mc.feature_dim * 4 * 2 / 3)

    def trans(n, w, depth):
        keys = sum((w if p >= w else 0) + (p % w) + 1 for p in range(n))
        return depth * (n * (3 * inner * d + inner * d + 3 * ff * d) + 2 * HEADS * keys * (d // 4))

    w, r = mc.en_coder_window_size, mc.en_coder_compress_rate
    t_tok = t_f // r
    if is_compressed(mc):
        macs += trans(t_f, w * r, 1) + t_tok * r * d * d + trans(t_tok, w, 2)
        macs += trans(t_tok, w, mc.en_coder_depth - 2) + trans(t_f, w * r, 2)
    else:
        macs += trans(t_f, w, 1) + trans(t_f, w, mc.en_coder_depth)
    macs += 2 * t_tok * d * len(mc.levels)
    # decoder
    t = t_f
    macs += t * 3 * d * mc.decoder_dims[0]
    for i, s in enumerate(mc.decode_rates):
        c = mc.decoder_dims[i]
        macs += mc.decoder_depths[i] * unit(c, t) + t * (4 * 7 + 4 * c) + t * c * mc.decoder_dims[i + 1]
        t *= s
    c = mc.decoder_dims[-1]
    macs += t * (3 * (7 * c * c + c * c) + 7 * c)
    params = sum(int(math.prod(shape)) for spec in network_spec(mc).values() for shape, _, _ in spec.values())
    codebook_size = math.prod(mc.levels)
    frame_rate = sample_rate / mc.hop_length
    return {"macs": f"{macs / 1e9:.2f} GMac", "params": f"{params / 1e6:.2f} M", "codebook_size": codebook_size,
            "frame_rate": frame_rate, "bps": frame_rate * math.log2(codebook_size),
            "receptive_field": mc.en_coder_window_size / frame_rate}


class L3AC:
    """l3ac/__init__.py:84-121."""

    def __init__(self, config: L3ACConfig, precision: str = "bf16"):
        self.config = config
        self.network = EnCodec(config.network_config, precision=precision)

    def download_weights(self):
        """l3ac/__init__.py:90-102: fetch every missing ``<module>.pt`` from ``config.weight_url`` (HTTP errors raise)."""
        import requests
        self.config.model_path.mkdir(parents=True, exist_ok=True)
        for module_name in self.network.trainable_modules:
            weight_url = self.config.weight_url.format(module_name)
            weight_path = self.config.model_path / f"{module_name}.pt"
            if weight_path.exists():
                log.info(f"{module_name}({weight_path}) already exists, skip download")
                continue
            log.warning(f"Downloading {module_name}({weight_url}) to {weight_path}")
            response = requests.get(weight_url, timeout=60)
            response.raise_for_status()
            weight_path.write_bytes(response.content)

    def load_pretrained(self):
        """l3ac/__init__.py:104-106.  Unlike the reference's ``load_model`` (which only logs a missing file), a module file
        that is still absent after the download step is an error: the caller asked for pretrained weights."""
        self.download_weights()
        missing = [n for n in self.network.trainable_modules if not (self.config.model_path / f"{n}.pt").exists()]
        if missing:
            raise FileNotFoundError(f"pretrained weights missing under {self.config.model_path}: {missing}")
        self.network.load_model(model_path=self.config.model_path)

    def encode_audio(self, audio_data: torch.Tensor):
        """(B, T) fp32 -> (q_feature (B, T_tok, F), {"indices": int32 (B, T_tok), "level_indices": fp32 (B, T_tok, D)})."""
        return self.network.engine.encode(audio_data)

    def decode_audio(self, audio_feature: torch.Tensor = None, indices: torch.Tensor = None, out: torch.Tensor = None) -> torch.Tensor:
        """q_feature (B, T_tok, F) or indices (B, T_tok) -> audio (B, T_tok * hop_length) fp32.  ``out`` (optional extension):
        a pinned host tensor that receives the waveform, downloaded micro-batch by micro-batch while the rest decodes."""
        return self.network.engine.decode(audio_feature, indices, out=out)
