"""Config objects with the reference's field names.

Mirrors ``ModelConfig`` (l3ac/codec.py:13-36 + l3ac/en_codec.py:9-19) and ``L3ACConfig``
(l3ac/__init__.py:54-81), read from the same TOML layout (l3ac/configs/*.toml).  Unknown keys are
rejected (the reference uses pydantic ``extra='forbid'``, l3ac/xtract/config.py:8); errors are
``ValueError`` (pydantic's ``ValidationError`` is a ``ValueError`` subclass).  The pydantic-settings
environment-variable override (l3ac/xtract/config.py:25-31) is not reproduced: it is outside the hot path.
"""
from __future__ import annotations

import dataclasses
import math
import tomllib
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Tuple

CONFIG_DIR = Path(__file__).parent / "configs"


def _default_vq():
    return dict(name="super_fsq", levels=[7, 7, 7, 7, 7, 7], noise_rate=0.5)


@dataclass
class ModelConfig:
    feature_dim: int = 256
    compress_rates: Tuple[int, ...] = (9, 5)
    encoder_dims: Tuple[int, ...] = (24, 96, 192)
    encoder_depths: Tuple[int, ...] = (1, 1, 2)
    decode_rates: Tuple[int, ...] = (5, 3, 3)
    decoder_dims: Tuple[int, ...] = (256, 128, 64, 32)
    decoder_depths: Tuple[int, ...] = (3, 2, 1, 1)
    base_unit: str = "normal"
    use_norm: bool = True
    use_snake_act: bool = True
    decoder_last_layer: Optional[str] = None
    vq_config: dict = field(default_factory=_default_vq)
    en_coder_depth: int = 2
    en_coder_window_size: int = 500
    en_coder_dynamic_pos: bool = False
    en_coder_compress_rate: int = 1
    en_coder_cache_size: int = 0

    def __post_init__(self):
        for name in ("compress_rates", "encoder_dims", "encoder_depths", "decode_rates", "decoder_dims",
                     "decoder_depths"):
            setattr(self, name, tuple(int(v) for v in getattr(self, name)))
        # l3ac/codec.py:32-36
        if not (len(self.compress_rates) + 1 == len(self.encoder_dims) == len(self.encoder_depths)):
            raise ValueError("compress_rates / encoder_dims / encoder_depths lengths are inconsistent")
        if not (len(self.decode_rates) + 1 == len(self.decoder_dims) == len(self.decoder_depths)):
            raise ValueError("decode_rates / decoder_dims / decoder_depths lengths are inconsistent")

    @property
    def hop_length(self) -> int:
        """prod(compress_rates) * en_coder_compress_rate (l3ac/codec.py:27-30, l3ac/en_codec.py:16-19)."""
        return math.prod(self.compress_rates) * self.en_coder_compress_rate

    @property
    def levels(self) -> Tuple[int, ...]:
        return tuple(int(v) for v in self.vq_config["levels"])

    def as_dict(self) -> dict:
        d = dataclasses.asdict(self)
        d["hop_length"] = self.hop_length
        return d


def _from_table(cls, table: dict, what: str):
    names = {f.name for f in dataclasses.fields(cls)}
    extra = set(table) - names
    if extra:
        raise ValueError(f"{what}: extra keys are not permitted: {sorted(extra)}")
    return cls(**table)


@dataclass
class L3ACConfig:
    config_file: Optional[Path] = None
    model_name: str = "debug"
    sample_rate: int = 16000
    model_version: str = "v0.0"
    model_dir: Path = field(default_factory=lambda: Path.home() / ".cache" / "l3ac")
    weight_url: Optional[str] = None
    network_config: Optional[ModelConfig] = None

    def __post_init__(self):
        if self.config_file is not None:
            self.config_file = Path(self.config_file)
            with open(self.config_file, "rb") as fh:
                table = tomllib.load(fh)
            names = {f.name for f in dataclasses.fields(self)} - {"config_file"}
            extra = set(table) - names
            if extra:
                raise ValueError(f"{self.config_file.name}: extra keys are not permitted: {sorted(extra)}")
            defaults = L3ACConfig.__dataclass_fields__
            for key, value in table.items():          # TOML is the lowest-priority source
                current = getattr(self, key)
                f = defaults[key]
                dflt = f.default_factory() if f.default_factory is not dataclasses.MISSING else f.default
                if current == dflt or current is None:
                    setattr(self, key, value)
        if isinstance(self.network_config, dict):
            self.network_config = _from_table(ModelConfig, self.network_config, "network_config")
        if self.network_config is None:
            self.network_config = ModelConfig()
        self.model_dir = Path(self.model_dir)
        if self.weight_url is None:     # l3ac/__init__.py:74-81
            self.weight_url = ("https://huggingface.co/zhai-lw/L3AC/resolve/main/weights/"
                               f"{self.model_name}.{self.model_version}/" "{}.pt")

    @property
    def model_tag(self) -> str:
        return f"{self.model_name}.{self.model_version}"

    @property
    def model_path(self) -> Path:
        return self.model_dir / self.model_tag
