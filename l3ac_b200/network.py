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
        # module.load_state_dict(sd) on one stage (the reference's load_model does exactly this) must drop the packed engine
        self.register_load_state_dict_post_hook(lambda module, incompatible: owner.invalidate())

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
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

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
        # the packed engine (folded weight-norm, bf16 / split copies, bias tables, CUDA graphs) is keyed on the parameters'
        # version counters too, so in-place edits (p.copy_, p.mul_, optimiser steps) repack instead of serving stale weights
        key = (str(dev), self.precision, self.encoder_precision, sum(p._version for p in self.parameters()))
        if self._engine is None or self._engine_key != key:
            weights = {n: m.state_dict() for n, m in self.trainable_modules.items()}
            self._engine = Engine(self.mc, weights, dev, precision=self.precision,
                                  encoder_precision=self.encoder_precision)
            self._engine_key = key
        return self._engine

    # ---- stage entry points (channels-first at this boundary, like the reference modules) ----------
    def _call_encoder(self, audio_b1t: torch.Tensor):
        """Encoder.forward, l3ac/modules.py:114-116: (B, 1, T) -> (B, F, T_f)."""
        if audio_b1t.dim() != 3 or audio_b1t.shape[1] != 1:
            raise RuntimeError(f"encoder expects (batch, 1, samples), got {tuple(audio_b1t.shape)}")
        return self.engine.run_stage("conv_encoder", audio_b1t[:, 0]).permute(0, 2, 1)

    def _call_en_encoder(self, feature: torch.Tensor):
        """LocalEncoder / CompressedLocalEncoderWithCache.forward: (B, F, T_f) -> (B, T_tok, F)."""
        return self.engine.run_stage("en_encoder", feature.permute(0, 2, 1))

    def _call_quantizer(self, trans_feature: torch.Tensor):
        """VQEmbed.forward, l3ac/vq/__init__.py:25-30: (q_features, indices dict, vq_loss = [0.])."""
        q, idx, lvl, _ = self.engine.quantize(trans_feature.to(torch.float32))
        return q, {"indices": idx, "level_indices": lvl}, torch.zeros(1, device=q.device, dtype=q.dtype)

    def _to_features(self, indices: torch.Tensor):
        return self.engine.dequantize(indices)

    def _call_en_decoder(self, q_trans_feature: torch.Tensor):
        """LocalDecoder / CompressedLocalDecoderWithCache.forward: (B, T_tok, F) -> (B, F, T_f)."""
        return self.engine.run_stage("en_decoder", q_trans_feature).permute(0, 2, 1)

    def _call_decoder(self, q_feature: torch.Tensor):
        """Decoder.forward, l3ac/modules.py:200-201: (B, F, T_f) -> (B, 1, T)."""
        return self.engine.run_stage("conv_decoder", q_feature.permute(0, 2, 1)).unsqueeze(1)

    def forward(self, audio_data: torch.Tensor):
        """EnCodec.forward, l3ac/en_codec.py:53-72: the same keys, shapes and layouts as the reference's dict."""
        length = audio_data.shape[-1]
        r = self.engine.forward_all(audio_data)
        q_feature = r["quantized_feature"].permute(0, 2, 1)
        return {"generated_audio": r["audio"][..., :length],
                "embedded_audio": q_feature,
                "indices": r["indices"],
                "commit_loss": torch.zeros(1, device=q_feature.device, dtype=q_feature.dtype),
                "hidden_feature": dict(encoded_feature=r["encoded_feature"].permute(0, 2, 1),
                                       encoded_trans_feature=r["encoded_trans_feature"],
                                       quantized_trans_feature=r["quantized_trans_feature"],
                                       quantized_feature=q_feature)}

    # ---- legacy chunked API of the base Codec (l3ac/codec.py:111-161) ---------------------------------
    def compress(self, audio_data: torch.Tensor):
        """Codec.compress, l3ac/codec.py:111-114: encoder -> quantizer (the base class's path: no en_encoder).
        audio_data (B, 1, T) -> (indices dict, q_feature (B, T_f, F))."""
        feature = self._call_encoder(audio_data).permute(0, 2, 1)
        q_feature, indices, _ = self._call_quantizer(feature.contiguous())
        return indices, q_feature

    def decompress(self, indices: torch.Tensor = None, q_feature: torch.Tensor = None):
        """Codec.decompress, l3ac/codec.py:116-120: (to_features ->) decoder; returns (B, 1, T)."""
        if q_feature is None:
            q_feature = self._to_features(indices)
        return self._call_decoder(q_feature.permute(0, 2, 1))

    @torch.no_grad()
    def extract_unit(self, audio_data: torch.Tensor, process_window: int = 5 * 16000):
        """Codec.extract_unit, l3ac/codec.py:122-147: chunked compress of one long clip (1, T) with one-hop overlap.
        Returns (ChunkData of index chunks, ChunkData of q_feature chunks).  The reference feeds ``compress`` an index DICT
        per chunk and takes ``indices[0]``, which raises KeyError there; here the chunk's ``indices`` tensor is used."""
        assert len(audio_data) == 1, "Only support batch size 1"
        audio_data, _ = self.preprocess(audio_data)
        process_window = process_window // self.fill_length * self.fill_length
        chunk_audio = ChunkData(chunk_len=process_window, prefix_len=self.fill_length, original_data=audio_data[0])
        chunk_indices, chunk_q_feature = [], []
        for x in chunk_audio.chunk_data:
            indices, q_feature = self.compress(x[None, None, :])
            chunk_indices.append(indices["indices"][0])
            chunk_q_feature.append(q_feature[0])
        codec_chunk_len, codec_prefix_len = process_window // self.mc.hop_length, self.fill_length // self.mc.hop_length
        return (ChunkData(chunk_len=codec_chunk_len, prefix_len=codec_prefix_len, chunk_data=chunk_indices),
                ChunkData(chunk_len=codec_chunk_len, prefix_len=codec_prefix_len, chunk_data=chunk_q_feature))

    @torch.no_grad()
    def decode_unit(self, chunk_indices=None, chunk_q_feature=None, audio_length: int = None):
        """Codec.decode_unit, l3ac/codec.py:149-156: decode every chunk, drop each later chunk's one-hop prefix, concatenate."""
        if chunk_q_feature is None:
            chunk_audio = [self.decompress(indices=x[None, :])[0, 0] for x in chunk_indices.chunk_data]
        else:
            chunk_audio = [self.decompress(q_feature=x[None, :, :])[0, 0] for x in chunk_q_feature.chunk_data]
        chunk_audio = ChunkData(chunk_len=len(chunk_audio[0]), prefix_len=self.fill_length, chunk_data=chunk_audio)
        return chunk_audio.data[None, :]


class ChunkData:
    """l3ac/codec.py:164-195: a sequence either whole (``original_data``) or as overlapping chunks (``chunk_data``): chunk 0 is
    ``[0, chunk_len)``, chunk i > 0 is ``[i*chunk_len - prefix_len, (i+1)*chunk_len)``; ``data`` drops the prefixes again."""

    def __init__(self, chunk_len: int, prefix_len: int, original_data=None, chunk_data=None):
        assert chunk_len > prefix_len
        self.chunk_len = chunk_len
        self.prefix_len = prefix_len
        self._original_data = original_data
        self._chunk_data = chunk_data

    @property
    def data(self):
        if self._original_data is not None:
            return self._original_data
        parts = [self._chunk_data[0]] + [x[self.prefix_len:] for x in self._chunk_data[1:]]
        return torch.cat(parts, dim=0)

    @property
    def chunk_data(self):
        if self._chunk_data is not None:
            return self._chunk_data
        n = len(self._original_data)
        return [self._original_data[:self.chunk_len] if i == 0 else self._original_data[i - self.prefix_len:i + self.chunk_len]
                for i in range(0, n, self.chunk_len)]
