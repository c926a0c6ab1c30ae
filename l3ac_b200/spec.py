"""Parameter inventory of the codec network in the reference's checkpoint format.

One ``state_dict`` per trainable module (``encoder``, ``quantizer``, ``decoder``, ``en_encoder``,
``en_decoder``; l3ac/codec.py:67-73, l3ac/en_codec.py:46-51), with the reference's key names,
including the weight-norm parametrisation pairs ``parametrizations.weight.original0`` (g) /
``original1`` (v) (l3ac/layers.py:11-25).  The same inventory drives random initialisation,
checkpoint loading and weight packing, so a reference ``.pt`` file loads unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import ModelConfig

HEADS = 6                      # l3ac/local_trans.py:52
POS_BIAS_DIM_DIV = 2           # DynamicPositionBias(dim=dim // 2), l3ac/local_trans.py:30
MODULE_NAMES = ("encoder", "quantizer", "decoder", "en_encoder", "en_decoder")

# init kinds
TRUNC, ZEROS, ONES, UNIFORM, WN_G = "trunc_normal_0.02", "zeros", "ones", "uniform_fan_in", "weight_norm_g"
INV_FREQ = "rotary_inv_freq"        # SinusoidalEmbeddings buffer: 1 / theta ** (arange(0, d, 2) / d), theta = 10000
Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str, object]]"


def _wn(spec, p, v_shape):
    """weight-normed Conv1d / Linear with bias (l3ac/layers.py:24-25)."""
    out = v_shape[0]
    spec[f"{p}.bias"] = ((out,), ZEROS, None)
    spec[f"{p}.parametrizations.weight.original0"] = ((out,) + (1,) * (len(v_shape) - 1), WN_G, None)
    spec[f"{p}.parametrizations.weight.original1"] = (tuple(v_shape), TRUNC, None)


def _plain(spec, p, w_shape, bias=True):
    """torch-default initialised Conv1d / Linear (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in)))."""
    fan_in = int(math.prod(w_shape[1:]))
    spec[f"{p}.weight"] = (tuple(w_shape), UNIFORM, fan_in)
    if bias:
        spec[f"{p}.bias"] = ((w_shape[0],), UNIFORM, fan_in)


def _norm(spec, p, dim):
    spec[f"{p}.weight"] = ((dim,), ONES, None)
    spec[f"{p}.bias"] = ((dim,), ZEROS, None)


def _conv_unit(spec, p, dim):
    """ConvUnit (l3ac/modules.py:16-30)."""
    _wn(spec, f"{p}.dw_conv", (dim, 1, 7))
    _norm(spec, f"{p}.norm", dim)
    _wn(spec, f"{p}.pw_conv1", (4 * dim, dim))
    spec[f"{p}.act.alpha"] = ((1, 1, 4 * dim), ONES, None)
    spec[f"{p}.grn.gamma"] = ((1, 4 * dim), ZEROS, None)
    spec[f"{p}.grn.beta"] = ((1, 4 * dim), ZEROS, None)
    _wn(spec, f"{p}.pw_conv2", (dim, 4 * dim))


def encoder_spec(mc: ModelConfig) -> Spec:
    """Encoder (l3ac/modules.py:71-113) with V3FirstBlock stem (l3ac/tconv/__init__.py:8-27)."""
    s = OrderedDict()
    dims = mc.encoder_dims
    for i in range(5):
        _wn(s, f"blocks.0.blocks.{i}.1", (4, 1, 7))
    _wn(s, "blocks.0.conv_1", (80, 20, 1))
    _wn(s, "blocks.0.conv_2", (dims[0], 81, 1))
    blk = 1
    for i, stride in enumerate(mc.compress_rates):
        for j in range(mc.encoder_depths[i]):
            _conv_unit(s, f"blocks.{blk}.{j}.module", dims[i])
        blk += 1
        _wn(s, f"blocks.{blk}.0", (dims[i + 1], dims[i], stride))
        if mc.use_norm:
            _norm(s, f"blocks.{blk}.1", dims[i + 1])
        blk += 1
    for j in range(mc.encoder_depths[-1]):
        _conv_unit(s, f"blocks.{blk}.{j}.module", dims[-1])
    _wn(s, f"blocks.{blk + 1}", (mc.feature_dim, dims[-1], 3))
    return s


def decoder_spec(mc: ModelConfig) -> Spec:
    """Decoder (l3ac/modules.py:135-198) with EnhanceBlock (l3ac/tconv/__init__.py:30-38)."""
    s = OrderedDict()
    dims = mc.decoder_dims
    _wn(s, "blocks.0", (dims[0], mc.feature_dim, 3))
    blk = 1
    for i in range(len(mc.decode_rates)):
        for j in range(mc.decoder_depths[i]):
            _conv_unit(s, f"blocks.{blk}.{j}.module", dims[i])
        blk += 1
        for k in range(4):
            _wn(s, f"blocks.{blk}.blocks.{k}.1", (1, 1, 7))
        _norm(s, f"blocks.{blk}.merge_layer.0", 4)
        _plain(s, f"blocks.{blk}.merge_layer.1", (dims[i], 4, 1))
        blk += 1
        _wn(s, f"blocks.{blk}.0", (dims[i + 1], dims[i], 1))
        if mc.use_norm:
            _norm(s, f"blocks.{blk}.2", dims[i + 1])
        blk += 1
    c = dims[-1]
    p = f"blocks.{blk}.block"
    for j in range(3):
        q = f"{p}.0.{j}.module.block"
        s[f"{q}.0.alpha"] = ((1, c, 1), ONES, None)
        _wn(s, f"{q}.1", (c, c, 7))
        s[f"{q}.2.alpha"] = ((1, c, 1), ONES, None)
        _wn(s, f"{q}.3", (c, c, 1))
    s[f"{p}.1.alpha"] = ((1, c, 1), ONES, None)
    _wn(s, f"{p}.2", (1, c, 7))
    return s


def _local_trans(spec, p, dim, depth, dynamic_pos):
    """LocalTrans (l3ac/local_trans.py:7-53): LocalMHA + FeedForward of local-attention==1.11.2."""
    dim_head = dim // 4
    inner = HEADS * dim_head
    ff_inner = int(dim * 4 * 2 / 3)
    for l in range(depth):
        _norm(spec, f"{p}.layers.{l}.0.norm", dim)
        _plain(spec, f"{p}.layers.{l}.0.to_qkv", (3 * inner, dim), bias=False)
        if not dynamic_pos:     # use_rotary_pos_emb (l3ac/local_trans.py:29,36): the angle table's persistent buffer
            spec[f"{p}.layers.{l}.0.attn_fn.rel_pos.inv_freq"] = ((dim_head // 2,), INV_FREQ, dim_head)
        _plain(spec, f"{p}.layers.{l}.0.to_out", (dim, inner), bias=False)
        _norm(spec, f"{p}.layers.{l}.1.0", dim)
        _plain(spec, f"{p}.layers.{l}.1.1", (2 * ff_inner, dim), bias=False)
        _plain(spec, f"{p}.layers.{l}.1.4", (dim, ff_inner), bias=False)
    if dynamic_pos:
        h = dim // POS_BIAS_DIM_DIV
        _plain(spec, f"{p}.dynamic_pos_bias.mlp.0", (h, 1))
        _plain(spec, f"{p}.dynamic_pos_bias.mlp.2", (h, h))
        _plain(spec, f"{p}.dynamic_pos_bias.mlp.4", (HEADS, h))


def is_compressed(mc: ModelConfig) -> bool:
    """l3ac/en_codec.py:25 -- plain LocalEncoder/Decoder vs the Compressed...WithCache pair."""
    return not (mc.en_coder_compress_rate == 1 and mc.en_coder_cache_size == 0)


def en_encoder_spec(mc: ModelConfig) -> Spec:
    s = OrderedDict()
    dim, dyn = mc.feature_dim, mc.en_coder_dynamic_pos
    if is_compressed(mc):       # l3ac/local_trans.py:145-165 (depth=3 -> 1 + 2)
        _local_trans(s, "down_trans.trans", dim, 3 // 2, dyn)
        _wn(s, "down_trans.down_layer", (dim, dim, mc.en_coder_compress_rate))
        _local_trans(s, "local_trans", dim, 3 - 3 // 2, dyn)
    else:                       # l3ac/en_codec.py:27-29 (depth=1)
        _local_trans(s, "local_trans", dim, 1, dyn)
    return s


def en_decoder_spec(mc: ModelConfig) -> Spec:
    s = OrderedDict()
    dim, dyn = mc.feature_dim, mc.en_coder_dynamic_pos
    if is_compressed(mc):       # l3ac/local_trans.py:168-186
        _local_trans(s, "up_trans.trans", dim, 2, dyn)
        _local_trans(s, "local_trans", dim, mc.en_coder_depth - 2, dyn)
    else:
        _local_trans(s, "local_trans", dim, mc.en_coder_depth, dyn)
    return s


def quantizer_spec(mc: ModelConfig) -> Spec:
    """VQEmbed (l3ac/vq/__init__.py:6-15): plain nn.Linear project_in / project_out."""
    s = OrderedDict()
    d = len(mc.levels)
    _plain(s, "project_in", (d, mc.feature_dim))
    _plain(s, "project_out", (mc.feature_dim, d))
    return s


def network_spec(mc: ModelConfig) -> Dict[str, Spec]:
    return OrderedDict(
        encoder=encoder_spec(mc), quantizer=quantizer_spec(mc), decoder=decoder_spec(mc),
        en_encoder=en_encoder_spec(mc), en_decoder=en_decoder_spec(mc))


def init_state_dicts(mc: ModelConfig, seed: int = 0, jitter: bool = False) -> Dict[str, "OrderedDict[str, torch.Tensor]"]:
    """Seeded random weights with the reference's initialisation distributions.

    ``jitter=True`` additionally perturbs every parameter that is an identity at construction
    (GRN gamma/beta, Snake alpha, norm affines, zero biases, weight-norm gains) so that every
    term of the forward is exercised by parity tests (weight set "W1").
    Deterministic for a given torch version: CPU generator, fixed draw order.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = OrderedDict()
    for mod, spec in network_spec(mc).items():
        sd = OrderedDict()
        for key, (shape, kind, arg) in spec.items():
            if kind == TRUNC:
                t = torch.empty(shape)
                torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
            elif kind == UNIFORM:
                bound = 1.0 / math.sqrt(arg)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            elif kind == ONES:
                t = torch.ones(shape)
                if jitter:
                    if key.endswith("alpha"):
                        t = 0.5 + 1.5 * torch.rand(shape, generator=g)
                    else:
                        t = t + 0.1 * torch.randn(shape, generator=g)
            elif kind == ZEROS:
                t = torch.zeros(shape)
                if jitter:
                    std = 0.1 if (".grn." in key or ".norm." in key or ".merge_layer.0." in key) else 0.02
                    t = std * torch.randn(shape, generator=g)
            elif kind == INV_FREQ:
                t = 1.0 / (10000 ** (torch.arange(0, arg, 2).float() / arg))
            elif kind == WN_G:
                t = None    # filled below from v
            else:
                raise AssertionError(kind)
            sd[key] = t
        for key in list(sd):
            if key.endswith("original0"):
                v = sd[key[:-1] + "1"]
                gain = v.flatten(1).norm(dim=1).reshape(spec[key][0])      # weight_norm init: g = ||v||
                if jitter:
                    gain = gain * (1.0 + 0.1 * torch.randn(gain.shape, generator=g))
                sd[key] = gain
        out[mod] = sd
    return out
