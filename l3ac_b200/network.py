"""``codec.network``: an ``nn.Module`` with the reference's module names, checkpoint keys and helpers.

Mirrors ``EnCodec`` / ``Codec`` / ``xnn.Module`` (l3ac/en_codec.py:22-72, l3ac/codec.py:39-84,
l3ac/xtract/nn/module.py:11-54) as far as the encode/decode path needs: the five trainable modules hold the
parameters under the reference's ``state_dict`` keys (so ``.pt`` files interchange), ``.cuda()/.to()/.eval()``
work as usual, and the forward is executed by ``Engine`` on the CUDA kernels.  The submodules are parameter
containers only -- calling them (``network.encoder(x)``) goes through the engine as well.
"""
from __future__ import annotations

import logging
import math
import pathlib
from typing import Dict, Optional

import torch
from torch import nn

from .config import ModelConfig
from .engine import Engine
from .spec import MODULE_NAMES, init_state_dicts, network_spec

log = logging.getLogger("L3AC")


class ParamTree(nn.Module):
    """Holds parameters under dotted reference names by nesting anonymous child modules."""

    def __init__(self, tensors=None):
        super().__init__()
        for key, t in (tensors or {}).items():
            self._insert(key.split("."), t)

    def _insert(self, path, t):
        if len(path) == 1:
            self.register_parameter(path[0], nn.Parameter(t.clone(), requires_grad=False))
            return
        child = self._modules.get(path[0])
        if child is None:
            child = ParamTree()
            self.add_module(path[0], child)
        child._insert(path[1:], t)


class _Stage(ParamTree):
    """A trainable module of the reference (encoder, quantizer, ...) bound to its engine entry point."""

    def __init__(self, tensors, owner, fn_name):
        super().__init__(tensors)
        object.__setattr__(self, "_owner", owner)
        self._fn_name = fn_name

    def forward(self, *args, **kwargs):
        return getattr(self._owner, self._fn_name)(*args, **kwargs)


class EnCodec(nn.Module):
    def __init__(self, mc: ModelConfig, seed: Optional[int] = None, precision: str = "bf16",
                 encoder_precision: Optional[str] = None):
        super().__init__()
        self.mc = mc
        self.precision = precision
        self.encoder_precision = encoder_precision
        network_spec(mc)                                   # validates the layer options early
        sds = init_state_dicts(mc, seed=0 if seed is None else seed)
        self.encoder = _Stage(sds["encoder"], self, "_call_encoder")
        self.quantizer = _Stage(sds["quantizer"], self, "_call_quantizer")
        self.decoder = _Stage(sds["decoder"], self, "_call_decoder")
        self.en_encoder = _Stage(sds["en_encoder"], self, "_call_en_encoder")
        self.en_decoder = _Stage(sds["en_decoder"], self, "_call_en_decoder")
        self.quantizer.to_features = self._to_features      # VQEmbed.to_features, l3ac/vq/__init__.py:20-23
        self._engine: Optional[Engine] = None
        self._engine_key = None

    # ---- reference helpers ----------------------------------------------------------------------
    @property
    def trainable_modules(self) -> Dict[str, nn.Module]:
        """l3ac/codec.py:67-73 + l3ac/en_codec.py:46-51 (same order)."""
        return {name: getattr(self, name) for name in MODULE_NAMES}

    @property
    def fill_length(self) -> int:
        return self.mc.hop_length

    def preprocess(self, audio_data: torch.Tensor):
        """Codec.preprocess, l3ac/codec.py:79-84: right zero-pad to a multiple of hop_length."""
        length = audio_data.shape[-1]
        pad_len = math.ceil(length / self.fill_length) * self.fill_length - length
        return nn.functional.pad(audio_data, (0, pad_len)), length

    def save_model(self, model_dir=None, model_path=None):
        """l3ac/xtract/nn/module.py:36-41 (one ``<name>.pt`` state_dict per trainable module)."""
        model_path = pathlib.Path(model_path or pathlib.Path(model_dir) / "l3ac_b200.EnCodec")
        model_path.mkdir(exist_ok=True, parents=True)
        for name, module in self.trainable_modules.items():
            torch.save(module.state_dict(), model_path / f"{name}.pt")

    def load_model(self, model_dir=None, model_path=None):
        """l3ac/xtract/nn/module.py:43-54: strict per-module load; missing files are logged, not fatal."""
        model_path = pathlib.Path(model_path or pathlib.Path(model_dir) / "l3ac_b200.EnCodec")
        if not model_path.exists():
            log.warning(f"Model path ({model_path}) does not exist.")
            return
        for name, module in self.trainable_modules.items():
            module_path = model_path / f"{name}.pt"
            if module_path.exists():
                log.info(f"Loading module({name}) from ({module_path})")
                module.load_state_dict(torch.load(module_path, map_location="cpu", weights_only=True))
            else:
                log.info(f"Module({name})'s path: ({module_path}) does not exist.")
        self.invalidate()

    def load_state_dicts(self, weights: Dict[str, Dict[str, torch.Tensor]]):
        for name, sd in weights.items():
            getattr(self, name).load_state_dict(sd, strict=True)
        self.invalidate()

    # ---- engine management ------------------------------------------------------------------------
    def invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):          # .cuda() / .to(): weights moved -> repack lazily
        self.invalidate()
        return super()._apply(fn, *a, **k)

    @property
    def engine(self) -> Engine:
        dev = next(self.parameters()).device
        key = (str(dev), self.precision, self.encoder_precision)
        if self._engine is None or self._engine_key != key:
            weights = {n: m.state_dict() for n, m in self.trainable_modules.items()}
            self._engine = Engine(self.mc, weights, dev, precision=self.precision,
                                  encoder_precision=self.encoder_precision)
            self._engine_key = key
        return self._engine

    # ---- stage entry points (channels-first at this boundary, like the reference modules) ----------
    def _call_encoder(self, audio_b1t: torch.Tensor):
        raise NotImplementedError("call codec.encode_audio(); the conv encoder is fused with en_encoder in this build")

    def _call_en_encoder(self, feature):
        raise NotImplementedError("call codec.encode_audio(); the conv encoder is fused with en_encoder in this build")

    def _call_quantizer(self, trans_feature: torch.Tensor):
        q, idx, lvl, _ = self.engine.quantize(trans_feature.to(torch.float32))
        return q, {"indices": idx, "level_indices": lvl}, torch.zeros(1, device=q.device, dtype=q.dtype)

    def _to_features(self, indices: torch.Tensor):
        return self.engine.dequantize(indices)

    def _call_en_decoder(self, q_feature):
        raise NotImplementedError("call codec.decode_audio(); en_decoder is fused with the conv decoder in this build")

    def _call_decoder(self, feature):
        raise NotImplementedError("call codec.decode_audio(); en_decoder is fused with the conv decoder in this build")

    def forward(self, audio_data: torch.Tensor):
        """EnCodec.forward, l3ac/en_codec.py:53-72 (inference subset of the returned dict)."""
        length = audio_data.shape[-1]
        q, idx = self.engine.encode(audio_data)
        y = self.engine.decode(q)
        return {"generated_audio": y[..., :length], "indices": idx["indices"],
                "commit_loss": torch.zeros(1, device=y.device),
                "hidden_feature": dict(quantized_trans_feature=q)}
