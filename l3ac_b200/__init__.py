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
    """l3ac/__init__.py:17-18."""
    return sorted(p.relative_to(CONFIG_DIR).stem for p in CONFIG_DIR.rglob("*.toml"))


def get_model(config_name, pretrained: bool = True, precision: str = "bf16") -> "L3AC":
    """l3ac/__init__.py:21-25.  ``pretrained`` loads ``<model_dir>/<name>.<version>/*.pt`` when present (there is
    no download step here: the box has no network); otherwise the seeded random initialisation is kept."""
    config = L3ACConfig(config_file=CONFIG_DIR / f"{config_name}.toml")
    codec = L3AC(config, precision=precision)
    if pretrained:
        codec.load_pretrained()
    return codec


class L3AC:
    """l3ac/__init__.py:84-121."""

    def __init__(self, config: L3ACConfig, precision: str = "bf16"):
        self.config = config
        self.network = EnCodec(config.network_config, precision=precision)

    def load_pretrained(self):
        """l3ac/__init__.py:104-106 without the HTTP download (l3ac/__init__.py:90-102)."""
        self.network.load_model(model_path=self.config.model_path)

    def encode_audio(self, audio_data: torch.Tensor):
        """(B, T) fp32 -> (q_feature (B, T_tok, F), {"indices": int32 (B, T_tok), "level_indices": fp32 (B, T_tok, D)})."""
        return self.network.engine.encode(audio_data)

    def decode_audio(self, audio_feature: torch.Tensor = None, indices: torch.Tensor = None) -> torch.Tensor:
        """q_feature (B, T_tok, F) or indices (B, T_tok) -> audio (B, T_tok * hop_length) fp32."""
        return self.network.engine.decode(audio_feature, indices)
