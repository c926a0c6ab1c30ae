"""Restatement of ``local_attention.transformer`` / ``local_attention.local_attention`` v1.11.x
(public algorithm, written from its documented behaviour; see package docstring).

Only the constructor arguments the reference passes (``l3ac/local_trans.py:30,34-39``) are
supported; anything else raises so that a silent semantic drift is impossible.
"""
import torch
import torch.nn.functional as F
from torch import nn


class DynamicPositionBias(nn.Module):
    """MLP 1 -> dim -> dim -> heads on integer distances; ``forward(i, j)`` gives (heads, i, j)."""

    def __init__(self, dim, heads):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(1, dim), nn.SiLU(),
            nn.Linear(dim, dim), nn.SiLU(),
            nn.Linear(dim, heads),
        )

    def forward(self, i, j):
        assert j >= i
        device = next(self.parameters()).device
        rel_dist = torch.arange(j, dtype=torch.float, device=device)
        bias = self.mlp(rel_dist[:, None])                      # (j, heads)
        i_seq = torch.arange(j - i, j, device=device)
        j_seq = torch.arange(j, device=device)
        idx = (i_seq[:, None] - j_seq[None, :]).abs()           # (i, j)
        return bias[idx].permute(2, 0, 1)                       # (heads, i, j)


def _look_around(x, backward, forward, pad_value, dim=2):
    """x: (b, windows, n, ...) -> concat of [w-backward .. w+forward] windows along ``dim``."""
    t = x.shape[1]
    dims = (len(x.shape) - dim) * (0, 0)
    padded = F.pad(x, (*dims, backward, forward), value=pad_value)
    parts = [padded[:, ind:(ind + t), ...] for ind in range(forward + backward + 1)]
    return torch.cat(parts, dim=dim)


class SinusoidalEmbeddings(nn.Module):
    """Rotary angle table: ``forward(x)`` -> (freqs (n, dim), scale = ones(1)) for n = x.shape[-2]; ``inv_freq`` is a
    persistent buffer (it appears in the state dict as ``...attn_fn.rel_pos.inv_freq``)."""

    def __init__(self, dim, scale_base=None, use_xpos=False, theta=10000):
        super().__init__()
        assert not use_xpos
        inv_freq = 1. / (theta ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer('inv_freq', inv_freq)

    def forward(self, x):
        t = torch.arange(x.shape[-2], device=x.device).type_as(self.inv_freq)
        freqs = torch.einsum('i , j -> i j', t, self.inv_freq)
        return torch.cat((freqs, freqs), dim=-1), torch.ones(1, device=x.device)


def rotate_half(x):
    x1, x2 = x.reshape(*x.shape[:-1], 2, x.shape[-1] // 2).unbind(dim=-2)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, freqs, scale=1):
    """Window-relative rotation: keys at positions 0..2w-1 of [previous ; own], queries at the last q_len positions."""
    q_len = q.shape[-2]
    q_freqs = freqs[..., -q_len:, :]
    inv_scale = scale ** -1
    q = (q * q_freqs.cos() * scale) + (rotate_half(q) * q_freqs.sin() * scale)
    k = (k * freqs.cos() * inv_scale) + (rotate_half(k) * freqs.sin() * inv_scale)
    return q, k


class LocalAttention(nn.Module):
    def __init__(self, window_size, causal=False, look_backward=1, look_forward=None, dropout=0.,
                 autopad=False, exact_windowsize=False, scale=None, dim=None,
                 use_rotary_pos_emb=True, use_xpos=False, xpos_scale_base=None, shared_qk=False):
        super().__init__()
        look_forward = (0 if causal else 1) if look_forward is None else look_forward
        assert not (causal and look_forward > 0)
        assert not use_xpos and not shared_qk and dropout == 0.
        self.rel_pos = SinusoidalEmbeddings(dim) if (use_rotary_pos_emb and dim is not None) else None
        self.window_size, self.causal = window_size, causal
        self.look_backward, self.look_forward = look_backward, look_forward
        self.autopad, self.exact_windowsize, self.scale = autopad, exact_windowsize, scale

    def forward(self, q, k, v, mask=None, attn_bias=None):
        assert mask is None
        w, pad_value = self.window_size, -1
        lead = q.shape[:-2]
        q, k, v = (t.reshape(-1, *t.shape[-2:]) for t in (q, k, v))        # pack '* n d'
        orig_n = q.shape[1]
        if self.autopad:
            rem = (-orig_n) % w
            if rem:
                q, k, v = (F.pad(t, (0, 0, 0, rem), value=0.) for t in (q, k, v))
        b, n, d = q.shape
        scale = d ** -0.5 if self.scale is None else self.scale
        assert n % w == 0
        windows = n // w
        b_t = torch.arange(n, device=q.device).reshape(1, windows, w)
        bq, bk, bv = (t.reshape(b, windows, w, d) for t in (q, k, v))
        bq = bq * scale
        la = dict(backward=self.look_backward, forward=self.look_forward, pad_value=pad_value)
        bk, bv = _look_around(bk, **la), _look_around(bv, **la)
        if self.rel_pos is not None:
            pos_emb, xpos_scale = self.rel_pos(bk)
            bq, bk = apply_rotary_pos_emb(bq, bk, pos_emb, scale=xpos_scale)
        bq_t = b_t[..., :, None]
        bq_k = _look_around(b_t, **la)[..., None, :]
        pad_mask = bq_k == pad_value
        sim = torch.einsum('bhie,bhje->bhij', bq, bk)
        if attn_bias is not None:
            heads = attn_bias.shape[0]
            assert b % heads == 0
            sim = sim + attn_bias.repeat(b // heads, 1, 1)[:, None]       # 'h i j -> (b h) 1 i j'
        mask_value = -torch.finfo(sim.dtype).max
        if self.causal:
            causal_mask = bq_t < bq_k
            if self.exact_windowsize:
                causal_mask = causal_mask | (bq_t > (bq_k + w * self.look_backward))
            sim = sim.masked_fill(causal_mask, mask_value)
        sim = sim.masked_fill(pad_mask, mask_value)
        attn = sim.softmax(dim=-1)
        out = torch.einsum('bhij,bhje->bhie', attn, bv)
        out = out.reshape(b, n, d)[:, :orig_n]
        return out.reshape(*lead, orig_n, d)


class LocalMHA(nn.Module):
    def __init__(self, *, dim, window_size, dim_head=64, heads=8, dropout=0., causal=False, prenorm=False,
                 qk_rmsnorm=False, qk_scale=8, use_xpos=False, xpos_scale_base=None, exact_windowsize=None,
                 gate_values_per_head=False, **kwargs):
        super().__init__()
        assert not qk_rmsnorm and not gate_values_per_head and not use_xpos
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim) if prenorm else None
        self.heads = heads
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.attn_fn = LocalAttention(dim=dim_head, window_size=window_size, causal=causal, autopad=True,
                                      scale=None, exact_windowsize=True if exact_windowsize is None else exact_windowsize,
                                      dropout=dropout, **kwargs)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, mask=None, attn_bias=None):
        if self.norm is not None:
            x = self.norm(x)
        b, n, _ = x.shape
        q, k, v = self.to_qkv(x).chunk(3, dim=-1)
        q, k, v = (t.reshape(b, n, self.heads, -1).transpose(1, 2) for t in (q, k, v))   # b h n d
        out = self.attn_fn(q, k, v, mask=mask, attn_bias=attn_bias)
        out = out.transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class GEGLU(nn.Module):
    def forward(self, x):
        x, gate = x.chunk(2, dim=-1)
        return x * F.gelu(gate)


def FeedForward(dim, mult=4, dropout=0.):
    inner = int(dim * mult * 2 / 3)
    return nn.Sequential(
        nn.LayerNorm(dim),
        nn.Linear(dim, inner * 2, bias=False),
        GEGLU(),
        nn.Dropout(dropout),
        nn.Linear(inner, dim, bias=False),
    )
