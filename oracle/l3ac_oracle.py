"""CPU oracle for the L3AC encode -> quantize -> decode hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file; the product package
(``l3ac_b200``) never does, and fails loudly when its CUDA library is missing.

What it is: a plain, functional, fp32 torch-CPU restatement of the reference forward, driven by the
reference's own checkpoint format (one ``state_dict`` per trainable module, weight-norm
parametrisation keys included).  Every function cites the reference ``file:line`` it follows
(paths relative to /root/reference).  It does not import the reference, so it travels to the GPU box.

How it is pinned: the reference ships no tests, fixtures or golden vectors, so
parity is pinned by running the *unmodified* reference (imported from /root/reference in the build
container, with ``oracle/shim/local_attention`` standing in for the absent PyPI dependency
``local-attention==1.11.2``) on the same weights/inputs: ``oracle/make_golden.py`` does that, asserts
this file agrees with it, and commits the outputs under ``tests/golden``.
PARITY UNPINNED for the attention arithmetic only: ``local-attention`` itself could not be installed,
its algorithm is restated from its published source (see ``oracle/shim/local_attention``).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence

import torch
import torch.nn.functional as F

EPS = 1e-8          # l3ac/xtract/nn/utils.py:33
SD = Mapping[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# parameter helpers
# ----------------------------------------------------------------------------------------------
def wn_weight(sd: SD, prefix: str) -> torch.Tensor:
    """Effective weight of a weight-normed layer: w = g * v / ||v|| (norm over all dims but 0).

    Follows torch.nn.utils.parametrizations.weight_norm as applied by l3ac/layers.py:17-18;
    ``original0`` is g, ``original1`` is v.  A plain ``<prefix>.weight`` key is also accepted.
    """
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    g = sd[prefix + ".parametrizations.weight.original0"]
    v = sd[prefix + ".parametrizations.weight.original1"]
    return torch._weight_norm(v, g, 0)


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """l3ac/layers.py:29-33."""
    return x + (alpha + EPS).reciprocal() * torch.sin(alpha * x).pow(2)


def channel_norm_cf(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """channels-first ChannelNorm, l3ac/layers.py:50-56,76-78 (x: (B, C, T))."""
    eps = torch.tensor(EPS)
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return weight.view(1, -1, 1) * x + bias.view(1, -1, 1)


def grn(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """l3ac/layers.py:112-115 (channels_last; the norm is over time AND channel)."""
    g_x = torch.norm(x, p=2, dim=[1, 2], keepdim=True)
    n_x = g_x / (g_x.mean(dim=-1, keepdim=True) + torch.tensor(EPS))
    return gamma * (x * n_x) + beta + x


def trend_pool(x: torch.Tensor, k: int) -> torch.Tensor:
    """l3ac/tconv/base.py:8-14."""
    if k > 1:
        args = dict(kernel_size=k, stride=1, padding=k // 2)
        return F.avg_pool1d(F.max_pool1d(x.abs(), **args), **args)
    return x


# ----------------------------------------------------------------------------------------------
# conv stack
# ----------------------------------------------------------------------------------------------
def base_block(sd: SD, p: str, x: torch.Tensor, pool_kernels: Sequence[int], dilation_rate: int) -> torch.Tensor:
    """l3ac/tconv/base.py:27-45: parallel TrendPool -> Conv1d(1 -> each, k7) branches, concatenated."""
    outs = []
    for i, pk in enumerate(pool_kernels):
        dil = pk // dilation_rate + 1
        pad = (7 - 1) * dil // 2
        w = wn_weight(sd, f"{p}.blocks.{i}.1")
        outs.append(F.conv1d(trend_pool(x, pk), w, sd[f"{p}.blocks.{i}.1.bias"], dilation=dil, padding=pad))
    return torch.cat(outs, dim=1)


def first_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """V3FirstBlock, l3ac/tconv/__init__.py:8-27 (pool kernels 1,5,11,21,45; dilation_rate 99)."""
    h = base_block(sd, p, x, (1, 5, 11, 21, 45), 99)
    h = F.conv1d(h, wn_weight(sd, f"{p}.conv_1"), sd[f"{p}.conv_1.bias"])
    h = F.gelu(h)
    y = torch.cat([h, x], dim=1)
    return F.conv1d(y, wn_weight(sd, f"{p}.conv_2"), sd[f"{p}.conv_2.bias"])


def enhance_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """EnhanceBlock, l3ac/tconv/__init__.py:30-44 (pool kernels 1,3,5,9; dilation_rate 2)."""
    yi = base_block(sd, p, x[:, :1, :], (1, 3, 5, 9), 2)
    yi = F.instance_norm(yi, weight=sd[f"{p}.merge_layer.0.weight"], bias=sd[f"{p}.merge_layer.0.bias"], eps=1e-5)
    y = F.conv1d(yi, sd[f"{p}.merge_layer.1.weight"], sd[f"{p}.merge_layer.1.bias"])
    return x + y * x


def conv_unit(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """Residual(ConvUnit), l3ac/modules.py:10-44 + l3ac/xtract/nn/layers.py:59-62.  x: (B, C, T)."""
    c = x.shape[1]
    h = F.conv1d(x, wn_weight(sd, f"{p}.dw_conv"), sd[f"{p}.dw_conv.bias"], padding=3, groups=c)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (c,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], torch.tensor(EPS).item())
    h = F.linear(h, wn_weight(sd, f"{p}.pw_conv1"), sd[f"{p}.pw_conv1.bias"])
    h = snake(h, sd[f"{p}.act.alpha"])
    h = grn(h, sd[f"{p}.grn.gamma"], sd[f"{p}.grn.beta"])
    h = F.linear(h, wn_weight(sd, f"{p}.pw_conv2"), sd[f"{p}.pw_conv2.bias"])
    return x + h.permute(0, 2, 1)


def legacy_unit(sd: SD, p: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    """Residual(LegacyUnit), l3ac/modules.py:47-64."""
    h = snake(x, sd[f"{p}.block.0.alpha"])
    h = F.conv1d(h, wn_weight(sd, f"{p}.block.1"), sd[f"{p}.block.1.bias"], dilation=dilation, padding=3 * dilation)
    h = snake(h, sd[f"{p}.block.2.alpha"])
    h = F.conv1d(h, wn_weight(sd, f"{p}.block.3"), sd[f"{p}.block.3.bias"])
    return x + h


def encoder(sd: SD, cfg: dict, x: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """Encoder.forward, l3ac/modules.py:71-116.  x: (B, 1, T) -> (B, feature_dim, T_f)."""
    h = first_block(sd, "blocks.0", x)
    if taps is not None:
        taps["enc_stem"] = h
    blk = 1
    for stride, depth in zip(cfg["compress_rates"], cfg["encoder_depths"][:-1]):
        for j in range(depth):
            h = conv_unit(sd, f"blocks.{blk}.{j}.module", h)
        blk += 1
        h = F.conv1d(h, wn_weight(sd, f"blocks.{blk}.0"), sd[f"blocks.{blk}.0.bias"], stride=stride)
        h = channel_norm_cf(h, sd[f"blocks.{blk}.1.weight"], sd[f"blocks.{blk}.1.bias"])
        if taps is not None:
            taps[f"enc_down{(blk // 2) - 1}"] = h
        blk += 1
    for j in range(cfg["encoder_depths"][-1]):
        h = conv_unit(sd, f"blocks.{blk}.{j}.module", h)
    blk += 1
    return F.conv1d(h, wn_weight(sd, f"blocks.{blk}"), sd[f"blocks.{blk}.bias"], padding=1)


def decoder(sd: SD, cfg: dict, x: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """Decoder.forward, l3ac/modules.py:135-201 (decoder_last_layer='legacy').  (B,F,T_f) -> (B,1,T)."""
    assert cfg.get("decoder_last_layer") == "legacy"
    h = F.conv1d(x, wn_weight(sd, "blocks.0"), sd["blocks.0.bias"], padding=1)
    blk = 1
    for si, (stride, depth) in enumerate(zip(cfg["decode_rates"], cfg["decoder_depths"][:-1])):
        for j in range(depth):
            h = conv_unit(sd, f"blocks.{blk}.{j}.module", h)
        blk += 1
        h = enhance_block(sd, f"blocks.{blk}", h)
        blk += 1
        h = F.conv1d(h, wn_weight(sd, f"blocks.{blk}.0"), sd[f"blocks.{blk}.0.bias"])
        h = F.interpolate(h, scale_factor=stride, mode="linear", align_corners=False)
        h = channel_norm_cf(h, sd[f"blocks.{blk}.2.weight"], sd[f"blocks.{blk}.2.bias"])
        if taps is not None:
            taps[f"dec_up{si}"] = h
        blk += 1
    p = f"blocks.{blk}.block"
    for j, dil in enumerate((1, 3, 9)):
        h = legacy_unit(sd, f"{p}.0.{j}.module", h, dil)
    h = snake(h, sd[f"{p}.1.alpha"])
    h = F.conv1d(h, wn_weight(sd, f"{p}.2"), sd[f"{p}.2.bias"], padding=3)
    return torch.tanh(h)


# ----------------------------------------------------------------------------------------------
# windowed transformer (local-attention==1.11.2 restated; call sites l3ac/local_trans.py:23-53)
# ----------------------------------------------------------------------------------------------
def dynamic_position_bias(sd: SD, p: str, w: int) -> torch.Tensor:
    """DynamicPositionBias(dim=64, heads)(w, 2w) -> (heads, w, 2w); l3ac/local_trans.py:30,43."""
    j = 2 * w
    w0 = sd[f"{p}.mlp.0.weight"]
    d = torch.arange(j, dtype=w0.dtype, device=w0.device)[:, None]         # fp32 in the reference (fp64 only in tools/fp32_floor.py)
    h = F.silu(F.linear(d, sd[f"{p}.mlp.0.weight"], sd[f"{p}.mlp.0.bias"]))
    h = F.silu(F.linear(h, sd[f"{p}.mlp.2.weight"], sd[f"{p}.mlp.2.bias"]))
    table = F.linear(h, sd[f"{p}.mlp.4.weight"], sd[f"{p}.mlp.4.bias"])          # (2w, heads)
    idx = (torch.arange(w, j, device=w0.device)[:, None] - torch.arange(j, device=w0.device)[None, :]).abs()
    return table[idx].permute(2, 0, 1)


def rotary_tables(inv_freq: torch.Tensor, n: int):
    """SinusoidalEmbeddings.forward of local-attention: angles t * inv_freq (fp32 product) for t = 0..n-1, duplicated to the
    full head dim; returns (cos, sin) of shape (n, d)."""
    t = torch.arange(n, device=inv_freq.device).type_as(inv_freq)
    freqs = torch.einsum("i , j -> i j", t, inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1)
    return freqs.cos(), freqs.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x.reshape(*x.shape[:-1], 2, x.shape[-1] // 2).unbind(dim=-2)
    return torch.cat((-x2, x1), dim=-1)


def local_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, bias, w: int, inv_freq=None) -> torch.Tensor:
    """LocalAttention(window_size=w, causal, look_backward=1, look_forward=0, autopad, exact_windowsize=False).

    q,k,v: (B*heads, n, d).  Keys of window i are [window i-1 ; window i]; window -1 is padding
    (value -1, masked out).  bias: (heads, w, 2w) or None; ``inv_freq`` (rotary mode, l3ac/local_trans.py:29,36): keys are
    rotated by their position 0..2w-1 inside [previous ; own], queries by w..2w-1, after q has been scaled.
    """
    bh, n0, d = q.shape
    rem = (-n0) % w
    if rem:
        q, k, v = (F.pad(t, (0, 0, 0, rem)) for t in (q, k, v))
    n = q.shape[1]
    nw = n // w
    pos = torch.arange(n, device=q.device).reshape(1, nw, w)

    def look_around(t, pad_value):
        prev = F.pad(t, (0,) * (2 * (t.dim() - 2)) + (1, 0), value=pad_value)[:, :nw]
        return torch.cat([prev, t], dim=2)

    bq = q.reshape(bh, nw, w, d) * (d ** -0.5)
    bk = look_around(k.reshape(bh, nw, w, d), -1.)
    bv = look_around(v.reshape(bh, nw, w, d), -1.)
    if inv_freq is not None:
        cos, sin = rotary_tables(inv_freq, 2 * w)
        scale = torch.ones(1, device=q.device)
        bq = (bq * cos[-w:] * scale) + (rotate_half(bq) * sin[-w:] * scale)
        bk = (bk * cos * scale ** -1) + (rotate_half(bk) * sin * scale ** -1)
    q_pos = pos[..., :, None]
    k_pos = look_around(pos, -1)[..., None, :]
    sim = torch.einsum("bhie,bhje->bhij", bq, bk)
    if bias is not None:
        heads = bias.shape[0]
        sim = sim + bias.repeat(bh // heads, 1, 1)[:, None]
    neg = -torch.finfo(sim.dtype).max
    sim = sim.masked_fill(q_pos < k_pos, neg)
    sim = sim.masked_fill(k_pos == -1, neg)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhje->bhie", attn, bv)
    return out.reshape(bh, n, d)[:, :n0]


def local_trans(sd: SD, p: str, x: torch.Tensor, depth: int, w: int, heads: int = 6) -> torch.Tensor:
    """LocalTrans.forward, l3ac/local_trans.py:42-48; LocalMHA(prenorm) + FeedForward(GEGLU)."""
    rotary = f"{p}.dynamic_pos_bias.mlp.0.weight" not in sd           # use_rotary_pos_emb = not use_dynamic_pos_bias
    bias = None if rotary else dynamic_position_bias(sd, f"{p}.dynamic_pos_bias", w)
    b, n, dim = x.shape
    for l in range(depth):
        a = f"{p}.layers.{l}.0"
        h = F.layer_norm(x, (dim,), sd[f"{a}.norm.weight"], sd[f"{a}.norm.bias"], 1e-5)
        q, k, v = F.linear(h, sd[f"{a}.to_qkv.weight"]).chunk(3, dim=-1)
        q, k, v = (t.reshape(b, n, heads, -1).transpose(1, 2).reshape(b * heads, n, -1) for t in (q, k, v))
        o = local_attention(q, k, v, bias, w, sd[f"{a}.attn_fn.rel_pos.inv_freq"] if rotary else None)
        o = o.reshape(b, heads, n, -1).transpose(1, 2).reshape(b, n, -1)
        x = F.linear(o, sd[f"{a}.to_out.weight"]) + x
        f = f"{p}.layers.{l}.1"
        h = F.layer_norm(x, (dim,), sd[f"{f}.0.weight"], sd[f"{f}.0.bias"], 1e-5)
        val, gate = F.linear(h, sd[f"{f}.1.weight"]).chunk(2, dim=-1)
        x = F.linear(val * F.gelu(gate), sd[f"{f}.4.weight"]) + x
    return x


def en_encoder(sd: SD, cfg: dict, feature: torch.Tensor) -> torch.Tensor:
    """LocalEncoder / CompressedLocalEncoderWithCache, l3ac/local_trans.py:56-74,145-165; en_codec.py:25-44."""
    x = feature.permute(0, 2, 1)
    w, r = cfg["en_coder_window_size"], cfg["en_coder_compress_rate"]
    if r == 1 and cfg.get("en_coder_cache_size", 0) == 0:
        return local_trans(sd, "local_trans", x, 1, w)
    x = local_trans(sd, "down_trans.trans", x, 3 // 2, w * r)
    x = F.conv1d(x.permute(0, 2, 1), wn_weight(sd, "down_trans.down_layer"), sd["down_trans.down_layer.bias"],
                 stride=r).permute(0, 2, 1)
    return local_trans(sd, "local_trans", x, 3 - 3 // 2, w)


def en_decoder(sd: SD, cfg: dict, q_feature: torch.Tensor) -> torch.Tensor:
    """LocalDecoder / CompressedLocalDecoderWithCache, l3ac/local_trans.py:77-94,114-126,168-186."""
    w, r, depth = cfg["en_coder_window_size"], cfg["en_coder_compress_rate"], cfg["en_coder_depth"]
    if r == 1 and cfg.get("en_coder_cache_size", 0) == 0:
        return local_trans(sd, "local_trans", q_feature, depth, w).permute(0, 2, 1)
    x = local_trans(sd, "local_trans", q_feature, depth - 2, w)
    x = F.interpolate(x.permute(0, 2, 1), scale_factor=r, mode="linear", align_corners=False).permute(0, 2, 1)
    x = local_trans(sd, "up_trans.trans", x, 2, w * r)
    return x.permute(0, 2, 1)


# ----------------------------------------------------------------------------------------------
# FSQ bottleneck
# ----------------------------------------------------------------------------------------------
def fsq_basis(levels: Sequence[int], device=None) -> torch.Tensor:
    """l3ac/vq/fsq.py:15 -- cumprod([1, L0, ..., L4]); dim 0 is least significant."""
    return torch.cumprod(torch.tensor([1] + list(levels[:-1]), device=device), dim=0, dtype=torch.int32)


def fsq_quantize(z: torch.Tensor, levels: Sequence[int]):
    """SuperFSQ.forward in eval mode, l3ac/vq/fsq.py:30-68 + fsq_act.py:38-39 + fsq.py:21."""
    lv = torch.tensor(list(levels), dtype=torch.int32, device=z.device)
    shape = z.shape
    z2 = z.reshape(-1, len(levels))
    act = (torch.tanh(z2) + 1) / 2
    level_idx = (act * (lv - 1)).round()
    q_act = level_idx / (lv - 1)
    indices = (level_idx * fsq_basis(levels, z.device)).sum(dim=-1).to(torch.int32)
    q_z = q_act * 2 - 1
    return q_z.reshape(shape), indices.reshape(shape[:-1]), level_idx.reshape(shape)


def fsq_indices_to_codes(indices: torch.Tensor, levels: Sequence[int]) -> torch.Tensor:
    """l3ac/vq/fsq.py:70-81."""
    lv = torch.tensor(list(levels), dtype=torch.int32, device=indices.device)
    level_idx = (indices.unsqueeze(-1) // fsq_basis(levels, indices.device)) % lv
    return (level_idx / (lv - 1)) * 2 - 1


def quantizer_forward(sd: SD, cfg: dict, x: torch.Tensor):
    """VQEmbed.forward, l3ac/vq/__init__.py:25-30."""
    levels = cfg["vq_config"]["levels"]
    z = F.linear(x, sd["project_in.weight"], sd["project_in.bias"])
    q_z, indices, level_idx = fsq_quantize(z, levels)
    q_feature = F.linear(q_z, sd["project_out.weight"], sd["project_out.bias"])
    return q_feature, {"indices": indices, "level_indices": level_idx}, z


def quantizer_to_features(sd: SD, cfg: dict, indices: torch.Tensor) -> torch.Tensor:
    """VQEmbed.to_features, l3ac/vq/__init__.py:20-23."""
    codes = fsq_indices_to_codes(indices, cfg["vq_config"]["levels"])
    return F.linear(codes, sd["project_out.weight"], sd["project_out.bias"])


# ----------------------------------------------------------------------------------------------
# facade
# ----------------------------------------------------------------------------------------------
def hop_length(cfg: dict) -> int:
    """l3ac/codec.py:27-30 + l3ac/en_codec.py:16-19."""
    return math.prod(cfg["compress_rates"]) * cfg["en_coder_compress_rate"]


def preprocess(cfg: dict, audio: torch.Tensor):
    """Codec.preprocess, l3ac/codec.py:79-84 (right zero-pad to a multiple of hop_length)."""
    length = audio.shape[-1]
    hop = hop_length(cfg)
    pad = math.ceil(length / hop) * hop - length
    return F.pad(audio, (0, pad)), length


class Oracle:
    """Mirror of the reference ``L3AC`` facade (l3ac/__init__.py:84-121) over plain state dicts.

    ``weights``: {"encoder": sd, "en_encoder": sd, "quantizer": sd, "en_decoder": sd, "decoder": sd}.
    ``cfg``: the ``[network_config]`` table of the TOML as a dict.
    """

    def __init__(self, cfg: dict, weights: Dict[str, SD], device="cpu"):
        """``device``: "cpu" for the oracle proper; a CUDA device only for ``bench.py --impl reference-gpu`` (the same torch
        forward through ATen/cuBLAS/cuDNN kernels, TF32 off -- the reference's own GPU path, never a parity checker)."""
        self.cfg = dict(cfg)
        self.device = torch.device(device)
        self.w = {m: {k: v.detach().to(torch.float32).to(self.device) for k, v in sd.items()} for m, sd in weights.items()}

    @torch.no_grad()
    def encode_audio(self, audio: torch.Tensor, taps: dict | None = None):
        """l3ac/__init__.py:108-114."""
        audio, _ = preprocess(self.cfg, audio.to(device=self.device, dtype=torch.float32))
        feature = encoder(self.w["encoder"], self.cfg, audio.unsqueeze(1), taps)
        trans = en_encoder(self.w["en_encoder"], self.cfg, feature)
        q_feature, idx, z = quantizer_forward(self.w["quantizer"], self.cfg, trans)
        if taps is not None:
            taps.update(enc_feature=feature, trans_feature=trans, z=z)
        return q_feature, idx

    @torch.no_grad()
    def decode_audio(self, audio_feature: torch.Tensor = None, indices: torch.Tensor = None, taps: dict | None = None):
        """l3ac/__init__.py:116-121."""
        if audio_feature is None:
            audio_feature = quantizer_to_features(self.w["quantizer"], self.cfg, indices.to(self.device))
        audio_feature = audio_feature.to(self.device)
        q_feature = en_decoder(self.w["en_decoder"], self.cfg, audio_feature)
        if taps is not None:
            taps["dec_feature"] = q_feature
        return decoder(self.w["decoder"], self.cfg, q_feature, taps).squeeze(1)
