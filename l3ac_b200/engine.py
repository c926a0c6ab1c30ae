"""Forward engine: packs reference-format weights once, then runs encode / decode as kernel sequences.

The composition follows the reference forward op by op (citations per method); the arithmetic lives in
``libl3ac_b200.so``.  Weight packing (weight-norm folding ``w = g v / ||v||``, tap-major conv layouts,
GEGLU column interleave, the input-independent attention bias table, GRN folded to a per-channel affine)
is one-time, input-independent preprocessing done with torch on the parameters' device.

Precision modes
  ``fp32``  every GEMM on the fp32 SIMT path (parity mode, 1e-5 class).
  ``split`` both sides on the tensor cores with 3-term split-bf16 operands (x = hi + lo; hi*hi + lo*hi + hi*lo, fp32
            accumulate): fp32-class results (decode SNR vs the reference > 70 dB where ``bf16`` gives 6-27 dB at random
            init) at roughly three times the tensor work of ``bf16`` on the decode side.
  ``bf16``  decode side (en_decoder + decoder) GEMM operands in bf16 on the tcgen05 path with fp32
            accumulation and fp32 residual stream.  The encode side also runs on the tcgen05 path but with
            split-bf16 operands (x = hi + lo, three MMAs per k-block: hi*hi + lo*hi + hi*lo, fp32 accumulate),
            because token indices flip under plain bf16 rounding (93.6 % index agreement
            at bf16, 100 % / 99.985 % with the 3-term split).  ``encoder_precision="fp32"`` keeps the encode
            side on the SIMT path instead.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import ops
from .config import ModelConfig
from .spec import HEADS, is_compressed

EPS = 1e-8          # ChannelNorm eps, l3ac/xtract/nn/utils.py:33
LN_EPS = 1e-5       # nn.LayerNorm default inside local_attention
FF_PAD = 352        # FeedForward inner 341 padded to a multiple of 16 bf16 elements / 32 B


def fold_weight_norm(sd, prefix: str) -> torch.Tensor:
    """Effective weight of a weight-normed layer (l3ac/layers.py:17-18): g * v / ||v||, norm over dims != 0."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    g = sd[prefix + ".parametrizations.weight.original0"].float()
    v = sd[prefix + ".parametrizations.weight.original1"].float()
    norm = v.flatten(1).norm(dim=1).reshape(g.shape)
    return v * (g / norm)


def _taps_major(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight (Co, Ci, k) -> GEMM weight (Co, k*Ci) with the tap index outermost."""
    return w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()


class _Linear:
    """A packed GEMM weight: fp32 master + the copy the chosen path needs (bf16, or a split (hi, lo) bf16 pair)."""

    def __init__(self, w: torch.Tensor, bias: Optional[torch.Tensor], kind):
        self.w32 = w.contiguous().float()
        self.w16 = self.w32.to(torch.bfloat16).contiguous() if kind == torch.bfloat16 else None
        self.wsp = None
        if kind == ops.SPLIT:
            hi = self.w32.to(torch.bfloat16)
            self.wsp = ops.Split(hi.contiguous(), (self.w32 - hi.float()).to(torch.bfloat16).contiguous())
        self.bias = None if bias is None else bias.contiguous().float()

    def weight_for(self, a):
        if isinstance(a, ops.Split):
            return self.wsp
        return self.w32 if a.dtype == torch.float32 else self.w16


class Engine:
    def __init__(self, mc: ModelConfig, weights: Dict[str, Dict[str, torch.Tensor]], device, precision: str = "bf16",
                 max_chunk_seconds: float = 330.0, encoder_precision: Optional[str] = None):
        if precision not in ("fp32", "bf16", "split"):
            raise ValueError(f"precision must be 'fp32', 'bf16' or 'split', got {precision!r}")
        encoder_precision = encoder_precision or ("fp32" if precision == "fp32" else "split")
        if encoder_precision not in ("fp32", "split"):
            raise ValueError(f"encoder_precision must be 'fp32' or 'split', got {encoder_precision!r}")
        if mc.decoder_last_layer != "legacy" or mc.base_unit != "normal" or not mc.use_norm or not mc.use_snake_act:
            raise NotImplementedError("only the layer options used by the four named configs are built")
        if mc.en_coder_cache_size != 0:
            raise NotImplementedError("en_coder_cache_size must be 0 (the reference asserts the same)")
        self.mc = mc
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("l3ac_b200 runs on CUDA devices only; move the network with .cuda() first")
        self.precision = precision
        import os
        # micro-batch size (tools/sweep_chunks.sh, 64 x 10 s): 110 s -> 21.4 ms, 160 -> 20.9, 240 -> 20.1, 330 -> 19.6 (two micro-batches of
        # 32 clips), one micro-batch of 64 -> 19.5 but no overlap of the host copies (e2e 29.2 k vs 30.6 k audio-s/s)
        max_chunk_seconds = float(os.environ.get("L3AC_CHUNK_SECONDS", max_chunk_seconds))     # tuning knobs (bench sweeps)
        self.max_chunk_samples = int(max_chunk_seconds * 16000)
        # Small batches are launch-bound (~220 kernels per encode+decode step, ~1 ms of GPU work for one 10 s clip):
        # the kernel sequence of a given (batch, length) is captured once into a CUDA graph and replayed.
        self.graph_max_samples = int(40.0 * 16000)
        self.graph_cache_size = 32
        # Large batches: every micro-batch (encode and decode separately, one graph instance per micro-batch slot so that the
        # instances can run concurrently on their streams) is captured into a CUDA graph on its second appearance and
        # replayed from then on -- ~360 ctypes launches (10 ms of host time per 64 x 10 s step) become six graph launches.
        self.graph_chunks = os.environ.get("L3AC_GRAPH_CHUNKS", "1") != "0"
        self.graph_capture_after = int(os.environ.get("L3AC_GRAPH_AFTER", 2))   # capture a shape on its n-th appearance
        self._graphs = {}
        self._graph_seen = {}
        # Micro-batches are independent, so they are issued round-robin on a few streams: the HBM-bound stencil kernels of
        # one micro-batch then overlap the tensor-core GEMMs (one persistent, shared-memory-heavy CTA per SM) of another.
        self.num_streams = int(os.environ.get("L3AC_STREAMS", 4))    # measured: 2 -> +11 %, 4 -> +14 % at 64 x 10 s
        self._streams = None
        # Fused tcgen05 MLP kernel (mlp_fused.cu) for the bf16 (decode-side) ConvUnits: the 4C hidden activation stays in
        # TMEM / shared memory.  Bit-identical to the two-GEMM path.  Per 24-clip chunk on B200: C=48 288 vs 429 us, C=96
        # 222 vs 298 us, C=256 206 vs 278 us (it re-streams 1 MB of weights per 128-row tile out of L2);
        # C=512 needs more TMEM columns than an SM has and stays on two GEMM launches.
        self.fused_mlp = True
        self.thin_tc = os.environ.get("L3AC_THIN_TC", "1") != "0"     # fused tensor-core ConvUnit for the C = 24 / 48 encoder stages
        # The same kernel with plain bf16 operands for the decode-side C = 48 unit: measured SLOWER than dwconv7_ln + the tcgen05
        # fused MLP (664 vs 190 + 256 us per 24 clips: with one m-tile per warp and 16 warps per SM the kernel is latency-bound,
        # 21 % issue efficiency), so it is off; kept as a tested operand mode of the entry point.
        self.thin_tc_decode = os.environ.get("L3AC_THIN_TC_DECODE", "0") != "0"
        self.thin_impl = os.environ.get("L3AC_THIN_IMPL", "tcgen05")  # "mma_sync": the register-level cross-check kernel
        self.fused_mlp_max_c = 256
        # row-kernel fusions of the thin decode stages (bf16, C = 48 / 96): thread-per-row dwconv7 + LN, Upsample + ChannelNorm fused
        # with the next unit's dwconv7 + LN, EnhanceBlock gate + 1x1 up conv in one kernel ("0": the separate kernels)
        self.dwconv_rows = os.environ.get("L3AC_DWCONV_ROWS", "1") != "0"
        self.hidden_block_bytes = 0              # >0: L2-blocked ConvUnit MLP (measured slower, see _run_conv_unit)
        self.dec_dtype = {"bf16": torch.bfloat16, "split": ops.SPLIT, "fp32": torch.float32}[precision]
        self.enc_dtype = ops.SPLIT if encoder_precision == "split" else torch.float32
        # Step-level C ABI (csrc/codec.cu): for the product precision the library itself packs the checkpoint and owns the
        # launch sequence; encode / decode below are then thin callers of l3ac_encode / l3ac_decode per micro-batch.  The
        # Python sequence further down stays for the other precisions, the rotary path, debug taps, the per-stage entry
        # points and the per-launch instrumentation of bench.py (ops.OP_HOOK) -- the same kernels in the same order.
        self.native = None
        knobs_default = all(os.environ.get(k, d) == d for k, d in (("L3AC_THIN_TC", "1"), ("L3AC_THIN_TC_DECODE", "0"), ("L3AC_THIN_IMPL", "tcgen05"),
                                                                    ("L3AC_STEM_IMPL", "tcgen05"), ("L3AC_TAIL_IMPL", "tcgen05"), ("L3AC_ATT_IMPL", "tcgen05"),
                                                                    ("L3AC_DWCONV_ROWS", "1")))
        if (os.environ.get("L3AC_ENGINE", "native") == "native" and precision in ("bf16", "split") and encoder_precision == "split" and knobs_default
                and mc.en_coder_dynamic_pos and mc.feature_dim == 128 and mc.encoder_dims[0] == 24 and mc.decoder_dims[-1] == 24):
            self.native = ops.NativeCodec(mc, weights, self.device, precision=precision)
        w = {m: {k: v.detach().to(self.device) for k, v in sd.items()} for m, sd in weights.items()}
        with torch.no_grad():
            self._pack_encoder(w["encoder"])
            self._pack_en_encoder(w["en_encoder"])
            self._pack_quantizer(w["quantizer"])
            self._pack_en_decoder(w["en_decoder"])
            self._pack_decoder(w["decoder"])

    # ------------------------------------------------------------------ packing
    def _conv_unit(self, sd, p, kind):
        dw = fold_weight_norm(sd, f"{p}.dw_conv")                      # (C, 1, 7)
        gamma, beta = sd[f"{p}.grn.gamma"].float().flatten(), sd[f"{p}.grn.beta"].float().flatten()
        return dict(
            dw_w=dw[:, 0, :].t().contiguous(), dw_b=sd[f"{p}.dw_conv.bias"].float().contiguous(),
            ln_w=sd[f"{p}.norm.weight"].float().contiguous(), ln_b=sd[f"{p}.norm.bias"].float().contiguous(),
            pw1=_Linear(fold_weight_norm(sd, f"{p}.pw_conv1"), sd[f"{p}.pw_conv1.bias"], kind),
            alpha=sd[f"{p}.act.alpha"].float().flatten().contiguous(),
            ialpha=(1.0 / (sd[f"{p}.act.alpha"].float().flatten() + 1e-8)).contiguous(),
            # GRN (l3ac/layers.py:112-115): n_x = g/(g+1e-8) == 1 to within 1e-8/g, so gamma*(x*n_x)+beta+x is the
            # per-channel affine (1+gamma) x + beta (absolute deviation <= 1e-8*|gamma|).
            scale=(1.0 + gamma).contiguous(), shift=beta.contiguous(),
            pw2=_Linear(fold_weight_norm(sd, f"{p}.pw_conv2"), sd[f"{p}.pw_conv2.bias"], kind),
        )

    def _pack_encoder(self, sd):
        mc = self.mc
        ek = self.enc_dtype
        bw = torch.stack([fold_weight_norm(sd, f"blocks.0.blocks.{i}.1")[:, 0, :] for i in range(5)])   # (5,4,7)
        bb = torch.cat([sd[f"blocks.0.blocks.{i}.1.bias"].float() for i in range(5)])
        self.stem = dict(
            branch_w=bw.contiguous(), branch_b=bb.contiguous(),
            w1=fold_weight_norm(sd, "blocks.0.conv_1")[:, :, 0].contiguous(), b1=sd["blocks.0.conv_1.bias"].float(),
            w2=fold_weight_norm(sd, "blocks.0.conv_2")[:, :, 0].contiguous(), b2=sd["blocks.0.conv_2.bias"].float())
        self.stem_plan = None                                                 # tcgen05 stem: packed lazily on first use
        self.stem_impl = os.environ.get("L3AC_STEM_IMPL", "tcgen05")          # "mma_sync": the register-level cross-check kernel
        self.enc_stages = []
        blk = 1
        for i, stride in enumerate(mc.compress_rates):
            units = [self._conv_unit(sd, f"blocks.{blk}.{j}.module", ek) for j in range(mc.encoder_depths[i])]
            blk += 1
            down = _Linear(_taps_major(fold_weight_norm(sd, f"blocks.{blk}.0")), sd[f"blocks.{blk}.0.bias"], ek)
            self.enc_stages.append(dict(units=units, stride=stride, down=down,
                                        cn_w=sd[f"blocks.{blk}.1.weight"].float().contiguous(),
                                        cn_b=sd[f"blocks.{blk}.1.bias"].float().contiguous()))
            blk += 1
        self.enc_last = [self._conv_unit(sd, f"blocks.{blk}.{j}.module", ek) for j in range(mc.encoder_depths[-1])]
        self.enc_out = _Linear(_taps_major(fold_weight_norm(sd, f"blocks.{blk + 1}")), sd[f"blocks.{blk + 1}.bias"], ek)

    def _local_trans(self, sd, p, depth, window, kind):
        """LocalTrans weights + the position term.  Dynamic mode (l3ac/local_trans.py:30,43): the DynamicPositionBias table
        f[h][d], d = q_pos - k_pos in [0, 2w) (the MLP input is the integer distance, so the table is input-independent).
        Rotary mode (l3ac/local_trans.py:29,36): cos / sin of the bucket angles t * inv_freq, t in [0, 2w), computed on the
        CPU in fp32 exactly like SinusoidalEmbeddings (the attention kernels then see a zero bias table)."""
        q = f"{p}.dynamic_pos_bias.mlp"
        rotary = None
        if self.mc.en_coder_dynamic_pos:
            d = torch.arange(2 * window, dtype=torch.float32, device=self.device)[:, None]
            h = torch.nn.functional.silu(torch.nn.functional.linear(d, sd[f"{q}.0.weight"].float(), sd[f"{q}.0.bias"].float()))
            h = torch.nn.functional.silu(torch.nn.functional.linear(h, sd[f"{q}.2.weight"].float(), sd[f"{q}.2.bias"].float()))
            table = torch.nn.functional.linear(h, sd[f"{q}.4.weight"].float(), sd[f"{q}.4.bias"].float()).t().contiguous()
        else:
            table = torch.zeros((HEADS, 2 * window), device=self.device)
            rotary = []
            for l in range(depth):
                inv_freq = sd[f"{p}.layers.{l}.0.attn_fn.rel_pos.inv_freq"].detach().float().cpu()
                t = torch.arange(2 * window).type_as(inv_freq)
                freqs = torch.einsum("i , j -> i j", t, inv_freq)
                freqs = torch.cat((freqs, freqs), dim=-1)
                rotary.append((freqs.cos().contiguous().to(self.device), freqs.sin().contiguous().to(self.device)))
        layers = []
        for l in range(depth):
            a, f = f"{p}.layers.{l}.0", f"{p}.layers.{l}.1"
            w1 = sd[f"{f}.1.weight"].float()                       # (2*inner, dim): [value rows ; gate rows]
            inner = w1.shape[0] // 2
            dim = w1.shape[1]
            w1i = torch.zeros((2 * FF_PAD, dim), device=self.device)
            w1i[0:2 * inner:2] = w1[:inner]                        # interleave (value_i, gate_i) column pairs
            w1i[1:2 * inner:2] = w1[inner:]
            w2 = torch.zeros((dim, FF_PAD), device=self.device)
            w2[:, :inner] = sd[f"{f}.4.weight"].float()
            layers.append(dict(
                ln1_w=sd[f"{a}.norm.weight"].float().contiguous(), ln1_b=sd[f"{a}.norm.bias"].float().contiguous(),
                qkv=_Linear(sd[f"{a}.to_qkv.weight"], None, kind), out=_Linear(sd[f"{a}.to_out.weight"], None, kind),
                ln2_w=sd[f"{f}.0.weight"].float().contiguous(), ln2_b=sd[f"{f}.0.bias"].float().contiguous(),
                ff1=_Linear(w1i, None, kind), ff2=_Linear(w2, None, kind)))
        return dict(layers=layers, window=window, table=table, rotary=rotary)

    def _pack_en_encoder(self, sd):
        mc = self.mc
        w, r = mc.en_coder_window_size, mc.en_coder_compress_rate
        ek = self.enc_dtype
        if is_compressed(mc):
            self.enc_trans_frame = self._local_trans(sd, "down_trans.trans", 3 // 2, w * r, ek)
            self.enc_trans_down = _Linear(_taps_major(fold_weight_norm(sd, "down_trans.down_layer")),
                                          sd["down_trans.down_layer.bias"], ek)
            self.enc_trans_token = self._local_trans(sd, "local_trans", 3 - 3 // 2, w, ek)
        else:
            self.enc_trans_frame = None
            self.enc_trans_down = None
            self.enc_trans_token = self._local_trans(sd, "local_trans", 1, w, ek)

    def _pack_quantizer(self, sd):
        self.vq = dict(w_in=sd["project_in.weight"].float().contiguous(), b_in=sd["project_in.bias"].float().contiguous(),
                       w_out=sd["project_out.weight"].float().contiguous(),
                       b_out=sd["project_out.bias"].float().contiguous())

    def _pack_en_decoder(self, sd):
        mc = self.mc
        bf16 = self.dec_dtype
        w, r = mc.en_coder_window_size, mc.en_coder_compress_rate
        if is_compressed(mc):
            self.dec_trans_token = self._local_trans(sd, "local_trans", mc.en_coder_depth - 2, w, bf16)
            self.dec_trans_frame = self._local_trans(sd, "up_trans.trans", 2, w * r, bf16)
        else:
            self.dec_trans_token = self._local_trans(sd, "local_trans", mc.en_coder_depth, w, bf16)
            self.dec_trans_frame = None

    def _pack_decoder(self, sd):
        mc = self.mc
        bf16 = self.dec_dtype
        self.dec_in = _Linear(_taps_major(fold_weight_norm(sd, "blocks.0")), sd["blocks.0.bias"], bf16)
        self.dec_stages = []
        blk = 1
        for i, stride in enumerate(mc.decode_rates):
            units = [self._conv_unit(sd, f"blocks.{blk}.{j}.module", bf16) for j in range(mc.decoder_depths[i])]
            blk += 1
            e = f"blocks.{blk}"
            enh = dict(
                conv_w=torch.stack([fold_weight_norm(sd, f"{e}.blocks.{k}.1")[0, 0] for k in range(4)]).contiguous(),
                conv_b=torch.cat([sd[f"{e}.blocks.{k}.1.bias"].float() for k in range(4)]).contiguous(),
                in_w=sd[f"{e}.merge_layer.0.weight"].float().contiguous(),
                in_b=sd[f"{e}.merge_layer.0.bias"].float().contiguous(),
                merge_w=sd[f"{e}.merge_layer.1.weight"].float()[:, :, 0].contiguous(),
                merge_b=sd[f"{e}.merge_layer.1.bias"].float().contiguous())
            blk += 1
            up = _Linear(fold_weight_norm(sd, f"blocks.{blk}.0")[:, :, 0], sd[f"blocks.{blk}.0.bias"], bf16)
            self.dec_stages.append(dict(units=units, enh=enh, up=up, stride=stride,
                                        cn_w=sd[f"blocks.{blk}.2.weight"].float().contiguous(),
                                        cn_b=sd[f"blocks.{blk}.2.bias"].float().contiguous()))
            blk += 1
        p = f"blocks.{blk}.block"
        self.dec_legacy = []
        for j, dil in enumerate((1, 3, 9)):
            q = f"{p}.0.{j}.module.block"
            self.dec_legacy.append(dict(
                dil=dil, alpha0=sd[f"{q}.0.alpha"].float().flatten().contiguous(),
                conv=_Linear(_taps_major(fold_weight_norm(sd, f"{q}.1")), sd[f"{q}.1.bias"], bf16),
                alpha1=sd[f"{q}.2.alpha"].float().flatten().contiguous(),
                pw=_Linear(fold_weight_norm(sd, f"{q}.3")[:, :, 0], sd[f"{q}.3.bias"], bf16)))
        self.dec_tail = dict(alpha=sd[f"{p}.1.alpha"].float().flatten().contiguous(),
                             w=fold_weight_norm(sd, f"{p}.2")[0].t().contiguous(),       # (7, C)
                             bias=float(sd[f"{p}.2.bias"].float().item()))
        # fused tail (bf16 decode mode, 24 channels, dilations 1/3/9): weights in mma B-fragment order
        self.dec_tail_fused = None
        self.dec_tail_plan = None
        if bf16 == ops.SPLIT and mc.decoder_dims[-1] == 24 and os.environ.get("L3AC_TAIL_IMPL", "tcgen05") == "tcgen05":
            # precision="split": the same fused tcgen05 tail with 3-term split operands (l3ac_decoder_tail_tc_split)
            self.dec_tail_plan = ops.TailPlan(
                torch.stack([fold_weight_norm(sd, f"{p}.0.{j}.module.block.1") for j in range(3)]),
                torch.stack([u["conv"].bias for u in self.dec_legacy]),
                torch.stack([fold_weight_norm(sd, f"{p}.0.{j}.module.block.3")[:, :, 0] for j in range(3)]),
                torch.stack([u["pw"].bias for u in self.dec_legacy]),
                torch.stack([u["alpha0"] for u in self.dec_legacy]), torch.stack([u["alpha1"] for u in self.dec_legacy]),
                [u["dil"] for u in self.dec_legacy], self.dec_tail["alpha"], self.dec_tail["w"], self.dec_tail["bias"], self.device)
        if bf16 == torch.bfloat16 and mc.decoder_dims[-1] == 24:
            convs, pws = [], []
            for j in range(3):
                q = f"{p}.0.{j}.module.block"
                wc = fold_weight_norm(sd, f"{q}.1")                                         # (24, 24, 7)
                convs.append(ops.pack_mma_b_fragments(_taps_major(wc), k_pad=176))          # K = tap * 24 + channel
                pws.append(ops.pack_mma_b_fragments(fold_weight_norm(sd, f"{q}.3")[:, :, 0]))
            # product path: the tcgen05 kernel (weights packed into a plan by the library); the register-level mma.sync
            # kernel below stays as a cross-check (L3AC_TAIL_IMPL=mma_sync)
            if os.environ.get("L3AC_TAIL_IMPL", "tcgen05") == "tcgen05":
                self.dec_tail_plan = ops.TailPlan(
                    torch.stack([fold_weight_norm(sd, f"{p}.0.{j}.module.block.1") for j in range(3)]),
                    torch.stack([u["conv"].bias for u in self.dec_legacy]),
                    torch.stack([fold_weight_norm(sd, f"{p}.0.{j}.module.block.3")[:, :, 0] for j in range(3)]),
                    torch.stack([u["pw"].bias for u in self.dec_legacy]),
                    torch.stack([u["alpha0"] for u in self.dec_legacy]), torch.stack([u["alpha1"] for u in self.dec_legacy]),
                    [u["dil"] for u in self.dec_legacy], self.dec_tail["alpha"], self.dec_tail["w"], self.dec_tail["bias"], self.device)
            self.dec_tail_fused = dict(
                conv_frags=torch.stack(convs).contiguous(), pw_frags=torch.stack(pws).contiguous(),
                conv_bias=torch.stack([u["conv"].bias for u in self.dec_legacy]).contiguous(),
                pw_bias=torch.stack([u["pw"].bias for u in self.dec_legacy]).contiguous(),
                alpha0=torch.stack([u["alpha0"] for u in self.dec_legacy]).contiguous(),
                alpha1=torch.stack([u["alpha1"] for u in self.dec_legacy]).contiguous(),
                dilations=[u["dil"] for u in self.dec_legacy], alpha_f=self.dec_tail["alpha"], w_f=self.dec_tail["w"],
                bias_f=self.dec_tail["bias"])

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _lin(a, lin: _Linear, B, T, K, **kw):
        return ops.gemm(a, lin.weight_for(a), B=B, T=T, K=K, bias=lin.bias, **kw)

    @staticmethod
    def _as_operand(x: torch.Tensor, kind):
        """fp32 activation -> GEMM A operand of the requested kind."""
        if kind == torch.float32:
            return x
        if isinstance(x, ops.Split):          # the producer already wrote the split pair
            return x
        if kind == ops.SPLIT:
            return ops.split_bf16(x)
        return x.to(kind)

    def _run_conv_unit(self, x, u, act_dtype, out_kind=torch.float32, ch0_out: Optional[list] = None, a_pre=None):
        """Residual(ConvUnit) -- l3ac/modules.py:32-44.

        The 4C-wide hidden tensor is the largest activation of the path.  Optionally (``hidden_block_bytes`` > 0) the
        two point-wise GEMMs run back to back over row blocks sized so that one block of the hidden tensor stays in the
        126 MB L2.  Measured on B200 (profiles/r01_bench_history.md) this LOSES: blocks of ~20 k rows are only ~4 tiles
        per CTA, and the fill/drain of the persistent GEMM costs more than the saved HBM traffic (C=256 pw_conv2: 636
        -> 331 TFLOP/s), so it is off by default; keeping the hidden tensor on chip needs the fused MLP kernel.

        ``out_kind=ops.SPLIT`` (encode side, last unit before a GEMM consumer): the result is written as the split-bf16
        pair directly by the producing kernel, which removes a separate fp32 -> split pass over the tensor."""
        B, T, C = x.shape
        if act_dtype == ops.SPLIT and C in (24, 48) and self.thin_tc and self.thin_impl == "tcgen05":
            # thin encode-side stages: the whole unit in one tcgen05 kernel (3-term split operands, fp32-class)
            if u.get("plan") is None:
                u["plan"] = ops.ConvUnitPlan(u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS, u["pw1"].w32, u["pw1"].bias, u["alpha"],
                                             u["scale"], u["shift"], u["pw2"].w32, u["pw2"].bias, self.device)
            return ops.convunit_umma(x, u["plan"], out_dtype=out_kind)
        if act_dtype == ops.SPLIT and C in (24, 48) and self.thin_tc:
            # (cross-check path, L3AC_THIN_IMPL=mma_sync: the register-level kernel)
            return ops.convunit_thin_tc(x, u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS, u["pw1"].w32, u["pw1"].bias,
                                        u["alpha"], u["scale"], u["shift"], u["pw2"].w32, u["pw2"].bias, out_dtype=out_kind)
        if act_dtype == torch.bfloat16 and C == 48 and self.thin_tc_decode and out_kind == torch.float32:
            # decode-side C = 48 unit (1.9 M rows per 24 clips): dwconv + LN + MLP in one register-resident mma.sync kernel
            # with the decode side's bf16 operands, instead of dwconv7_ln + the (HBM-bound at this width) tcgen05 fused MLP
            return ops.convunit_thin_tc(x, u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS, u["pw1"].w32, u["pw1"].bias,
                                        u["alpha"], u["scale"], u["shift"], u["pw2"].w32, u["pw2"].bias, operands=torch.bfloat16)
        if C == 24 and act_dtype != torch.bfloat16:      # thin full-rate encoder stage: one fused fp32 kernel
            return ops.convunit_thin(x, u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS, u["pw1"].w32, u["pw1"].bias,
                                     u["alpha"], u["scale"], u["shift"], u["pw2"].w32, u["pw2"].bias, out_dtype=out_kind)
        if a_pre is not None:
            a = a_pre                        # (dwconv7 + LayerNorm already done by the fused up-layer kernel)
        elif act_dtype == torch.bfloat16 and C in (48, 96) and self.dwconv_rows and B <= 65535:
            # decode side's thin stages: thread-per-row kernel, parameters in a plan (constant bank)
            if u.get("dw_plan") is None:
                u["dw_plan"] = ops.DwconvPlan(u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS)
            a = ops.dwconv7_ln_plan(x, u["dw_plan"])
        else:
            a = ops.dwconv7_ln(x, u["dw_w"], u["dw_b"], u["ln_w"], u["ln_b"], EPS, out_dtype=act_dtype)
        M = B * T
        esz = {torch.float32: 4, torch.bfloat16: 2, ops.SPLIT: 4}[act_dtype]
        rows_blk = max(128 * 148, (self.hidden_block_bytes // (4 * C * esz)) // 128 * 128)
        if act_dtype == torch.bfloat16 and self.fused_mlp and 16 <= C <= self.fused_mlp_max_c and C % 16 == 0:
            r = ops.convunit_mlp(a, u["pw1"].w16, u["pw1"].bias, u["alpha"], u["scale"], u["shift"], u["pw2"].w16,
                                 u["pw2"].bias, x, ialpha=u["ialpha"], want_ch0=ch0_out is not None)
            if ch0_out is not None:          # (the unit in front of an EnhanceBlock also emits channel 0 as a compact plane)
                ch0_out.append(r[1])
                return r[0]
            return r
        if self.hidden_block_bytes <= 0 or act_dtype == torch.float32 or M <= rows_blk + rows_blk // 2:
            h = self._lin(a, u["pw1"], B, T, C, act=ops.ACT_SNAKE, alpha=u["alpha"], scale=u["scale"], shift=u["shift"],
                          out_dtype=act_dtype)
            return self._lin(h, u["pw2"], B, T, 4 * C, residual=x, out_dtype=out_kind)
        out = torch.empty_like(x)
        a2, x2, o2 = a.view(M, C), x.view(M, C), out.view(M, C)
        split = act_dtype == ops.SPLIT
        hbuf = ops._empty_act((rows_blk, 4 * C), x.device, act_dtype)[0]          # reused by every block
        for m0 in range(0, M, rows_blk):
            m1 = min(M, m0 + rows_blk)
            n = m1 - m0
            ab = ops.Split(a2.hi[m0:m1], a2.lo[m0:m1]) if split else a2[m0:m1]
            hb = ops.Split(hbuf.hi[:n], hbuf.lo[:n]) if split else hbuf[:n]
            self._lin(ab, u["pw1"], 1, n, C, act=ops.ACT_SNAKE, alpha=u["alpha"], scale=u["scale"], shift=u["shift"],
                      out_dtype=act_dtype, out=hb)
            self._lin(hb, u["pw2"], 1, n, 4 * C, residual=x2[m0:m1], out=o2[m0:m1])
        return out

    def _run_local_trans(self, x, lt, act_dtype):
        """LocalTrans.forward -- l3ac/local_trans.py:42-48 (LocalMHA prenorm + GEGLU FeedForward)."""
        B, T, D = x.shape
        for li, L in enumerate(lt["layers"]):
            a = ops.layernorm(x, L["ln1_w"], L["ln1_b"], LN_EPS, out_dtype=act_dtype)
            w = lt["window"]
            if lt["rotary"] is not None:
                # rotary mode: fp32 q/k/v -> rotated per-window segments in the attention operand kind (rotary.cu)
                cos_t, sin_t = lt["rotary"][li]
                seg = ops.rotary_pack(self._lin(a, L["qkv"], B, T, D), HEADS, w, cos_t, sin_t, out_dtype=act_dtype)
                if act_dtype == torch.float32:
                    o = ops.local_attention(seg, lt["table"], HEADS, w)
                else:
                    o = ops.local_attention_tc(seg, lt["table"], HEADS, w, out_dtype=act_dtype)
                o = ops.rotary_unpack(o, B, T, w)
            elif act_dtype == torch.float32:
                qkv = self._lin(a, L["qkv"], B, T, D)                               # fp32 (B,T,576)
                o = ops.local_attention(qkv, lt["table"], HEADS, w)
            else:   # q/k/v stay in the GEMM operand format (bf16 or split pair) and feed the tensor-core attention
                qkv = self._lin(a, L["qkv"], B, T, D, out_dtype=act_dtype)
                o = ops.local_attention_tc(qkv, lt["table"], HEADS, w, out_dtype=act_dtype)
            x = self._lin(o, L["out"], B, T, o.shape[-1], residual=x)
            a = ops.layernorm(x, L["ln2_w"], L["ln2_b"], LN_EPS, out_dtype=act_dtype)
            g = self._lin(a, L["ff1"], B, T, D, act=ops.ACT_GEGLU, out_dtype=act_dtype)   # (B,T,352)
            x = self._lin(g, L["ff2"], B, T, FF_PAD, residual=x)
        return x

    # ------------------------------------------------------------------ encode
    def conv_encoder(self, audio: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """Encoder.forward -- l3ac/modules.py:71-116.  audio (B, T) fp32 -> feature (B, T_f, F) channels-last fp32.
        A length that is not a multiple of a stage's stride is floored like the reference's strided Conv1d does."""
        f32 = self.enc_dtype            # operand kind of the encode-side GEMMs (fp32 SIMT or split-bf16 tcgen05)
        if f32 == ops.SPLIT and self.thin_tc and self.stem_impl == "tcgen05":
            if self.stem_plan is None:
                self.stem_plan = ops.StemPlan(**self.stem, device=self.device)
            x = ops.stem_umma(audio.contiguous(), self.stem_plan)               # (B, T, 24)
        else:
            stem = ops.stem_tc if (f32 == ops.SPLIT and self.thin_tc) else ops.stem
            x = stem(audio.contiguous(), **self.stem)
        if taps is not None:
            taps["enc_stem"] = x
        # The last ConvUnit before a GEMM consumer writes that GEMM's operand kind directly (split pair on the tensor-core
        # path); only when debug taps want the fp32 tensor, or the operand is fp32 anyway, is the stream kept in fp32.
        direct = f32 if (f32 == ops.SPLIT and self.hidden_block_bytes <= 0) else torch.float32
        for si, st in enumerate(self.enc_stages):
            B_, T_, C_ = x.shape
            s = st["stride"]
            keep = (T_ // s) * s
            last_kind = direct if keep == T_ else torch.float32
            for ui, u in enumerate(st["units"]):
                x = self._run_conv_unit(x, u, f32, out_kind=last_kind if ui == len(st["units"]) - 1 else torch.float32)
            if keep != T_:
                x = x[:, :keep].contiguous()
            x = self._lin(self._as_operand(x, f32), st["down"], B_, keep // s, s * C_)   # Conv1d(k=s, stride=s) as a GEMM
            x = ops.layernorm(x, st["cn_w"], st["cn_b"], EPS)                       # channels-first ChannelNorm
            if taps is not None:
                taps[f"enc_down{si}"] = x
        B_, T_, C_ = x.shape
        for ui, u in enumerate(self.enc_last):
            x = self._run_conv_unit(x, u, f32, out_kind=direct if ui == len(self.enc_last) - 1 else torch.float32)
        x = self._lin(self._as_operand(x, f32), self.enc_out, B_, T_, C_, taps=3, tap_shift0=-1)   # Conv1d(k3, pad 1)
        if taps is not None:
            taps["enc_feature"] = x
        return x

    def en_encoder(self, x: torch.Tensor) -> torch.Tensor:
        """LocalEncoder / CompressedLocalEncoderWithCache.forward -- l3ac/local_trans.py:63-74,138-142,161-165.
        feature (B, T_f, F) channels-last -> trans_feature (B, T_tok, F)."""
        f32 = self.enc_dtype
        if self.enc_trans_frame is not None:
            B_, T_, _ = x.shape
            x = self._run_local_trans(x, self.enc_trans_frame, f32)
            r = self.mc.en_coder_compress_rate
            keep = (T_ // r) * r
            if keep != T_:
                x = x[:, :keep].contiguous()
            x = self._lin(self._as_operand(x, f32), self.enc_trans_down, B_, keep // r, r * x.shape[-1])
        return self._run_local_trans(x, self.enc_trans_token, f32)

    def encode_features(self, audio: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """preprocess + encoder + en_encoder (l3ac/__init__.py:109-111) -> trans_feature (B, T_tok, F)."""
        B, T0 = audio.shape
        hop = self.mc.hop_length
        T = math.ceil(T0 / hop) * hop                      # Codec.preprocess, l3ac/codec.py:79-84
        if T != T0:
            audio = torch.nn.functional.pad(audio, (0, T - T0))
        return self.en_encoder(self.conv_encoder(audio, taps))

    def quantize(self, trans_feature: torch.Tensor, want_z: bool = False):
        """VQEmbed.forward -- l3ac/vq/__init__.py:25-30."""
        return ops.fsq_quantize(trans_feature.contiguous(), self.vq["w_in"], self.vq["b_in"], self.vq["w_out"],
                                self.vq["b_out"], self.mc.levels, want_z=want_z)

    def dequantize(self, indices: torch.Tensor) -> torch.Tensor:
        """VQEmbed.to_features -- l3ac/vq/__init__.py:20-23."""
        return ops.fsq_dequantize(indices.contiguous(), self.vq["w_out"], self.vq["b_out"], self.mc.levels)

    def _chunks(self, B: int, T: int):
        per = max(1, self.max_chunk_samples // max(T, 1))
        n = -(-B // per)                                     # balanced micro-batches (64 clips -> 22 + 21 + 21, not 24 + 24 + 16):
        base, extra = divmod(B, n) if n else (0, 0)          # the streams then finish together
        out, lo = [], 0
        for i in range(n):
            hi = lo + base + (1 if i < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out

    # ------------------------------------------------------------------ CUDA graphs (small, launch-bound batches)
    def _graphed(self, key, fn, inp: torch.Tensor):
        """Runs ``fn(inp)`` through a cached CUDA graph once the shape has been seen before; returns fresh outputs.

        A shape is captured on its SECOND appearance (variable-length workloads would otherwise pay warm-up + capture +
        instantiation, ~3x the eager cost, on almost every call and evict useful graphs); the first call runs eagerly.
        Capture uses the thread-local error mode (a DataLoader pin-memory thread issuing CUDA calls must not abort it) and
        any capture failure falls back to eager execution for that shape."""
        slot = self._graphs.get(key)
        if slot is None:
            seen = self._graph_seen.get(key, 0)
            if seen < 0 or seen + 1 < self.graph_capture_after:
                if seen >= 0:
                    if len(self._graph_seen) > 4096:
                        self._graph_seen.clear()
                    self._graph_seen[key] = seen + 1
                return fn(inp.to(self.device, non_blocking=True))
            if len(self._graphs) >= self.graph_cache_size:
                self._graphs.pop(next(iter(self._graphs)))          # drop the oldest capture (and its memory pool)
            static_in = inp.to(self.device, copy=True)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                            # warm-up outside capture (lazy init, attribute calls)
                fn(static_in)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            launches0 = ops.LAUNCHES
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    outs = fn(static_in)
            except RuntimeError:
                self._graph_seen[key] = -1                           # never try this shape again
                torch.cuda.synchronize(self.device)
                return fn(inp.to(self.device, non_blocking=True))
            slot = (graph, static_in, outs, ops.LAUNCHES - launches0)
            ops.LAUNCHES = launches0                                 # (captured, not launched)
            self._graphs[key] = slot
        graph, static_in, outs, n_kernels = slot
        static_in.copy_(inp, non_blocking=True)                      # (a pinned host slice is uploaded straight into the graph's input)
        graph.replay()
        ops.LAUNCHES += n_kernels                                    # the kernels of this library inside the replayed graph
        return tuple(o.clone() for o in outs)

    def _chunk_graphs(self, n_chunks: int, taps) -> bool:
        """Per-micro-batch CUDA graphs pay when PYTHON issues the ~120 launches of a micro-batch (10.9 -> 6.2 ms of host time per
        64 x 10 s step).  The step-level C ABI issues them in ~1 ms, so with it the micro-batches are launched directly (a
        256 x 30 s batch = 24 + 24 micro-batches would otherwise thrash the graph cache: every miss is a warm-up run, a capture
        and an instantiation).  On the Python path graphs are used only while every slot of the call fits the cache."""
        return (self.graph_chunks and not self._use_native() and n_chunks > 1 and 2 * n_chunks + 8 <= self.graph_cache_size
                and taps is None and ops.OP_HOOK is None and not torch.cuda.is_current_stream_capturing())

    def _run_chunks(self, fn, chunks):
        """Runs ``fn(lo, hi)`` for every micro-batch, round-robin over ``num_streams`` side streams; returns the results
        in order.  All side streams are joined back into the caller's stream before returning."""
        if len(chunks) <= 1 or self.num_streams <= 1 or torch.cuda.is_current_stream_capturing():
            return [fn(lo, hi) for lo, hi in chunks]
        if self._streams is None or len(self._streams) != self.num_streams:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.num_streams)]
        cur = torch.cuda.current_stream(self.device)
        for st in self._streams:
            st.wait_stream(cur)
        results = []
        for i, (lo, hi) in enumerate(chunks):
            with torch.cuda.stream(self._streams[i % self.num_streams]):
                results.append(fn(lo, hi))
        for st in self._streams:
            cur.wait_stream(st)
        for r in results:                                    # outputs were allocated on a side stream, are consumed on `cur`
            for t in (r if isinstance(r, (tuple, list)) else (r,)):
                if isinstance(t, torch.Tensor):
                    t.record_stream(cur)
        return results

    def _use_native(self) -> bool:
        return self.native is not None and ops.OP_HOOK is None

    def _encode_one_chunk(self, audio: torch.Tensor):
        if self._use_native():
            return self.native.encode(audio.contiguous())
        q, idx, lvl, _ = self.quantize(self.encode_features(audio))
        return q, idx, lvl

    def encode(self, audio: torch.Tensor, taps: Optional[dict] = None):
        """L3AC.encode_audio -- l3ac/__init__.py:108-114."""
        if audio.dim() != 2:
            raise RuntimeError(f"encode_audio expects a (batch, samples) tensor, got shape {tuple(audio.shape)}")
        # A pinned host batch that will be processed in several micro-batches is uploaded micro-batch by micro-batch on the
        # streams that consume it, so that the copies overlap the kernels of the other micro-batches.
        staged_on_host = (audio.device.type == "cpu" and audio.dtype == torch.float32 and audio.is_pinned()
                          and len(self._chunks(*audio.shape)) > 1 and taps is None)
        if not staged_on_host:
            audio = audio.to(device=self.device, dtype=torch.float32)
        if audio.shape[0] == 0 or audio.shape[1] == 0:            # empty batch: empty results of the right shapes
            if audio.shape[1] == 0:
                raise RuntimeError("encode_audio got clips of length 0")
            t_tok = math.ceil(audio.shape[1] / self.mc.hop_length)
            z = lambda *sh, dt=torch.float32: torch.zeros(sh, device=self.device, dtype=dt)
            return z(0, t_tok, self.mc.feature_dim), {"indices": z(0, t_tok, dt=torch.int32),
                                                      "level_indices": z(0, t_tok, len(self.mc.levels))}
        if taps is None and 0 < audio.numel() <= self.graph_max_samples and not torch.cuda.is_current_stream_capturing():
            with torch.cuda.device(self.device):
                q, idx, lvl = self._graphed(("enc",) + tuple(audio.shape), self._encode_one_chunk, audio.contiguous())
            return q, {"indices": idx, "level_indices": lvl}
        n_all = audio.shape[0]

        chunks = self._chunks(*audio.shape)
        use_graphs = self._chunk_graphs(len(chunks), taps)

        def run(lo, hi):
            if use_graphs:
                return self._graphed(("encc", chunks.index((lo, hi)), hi - lo, audio.shape[1]), self._encode_one_chunk, audio[lo:hi])
            a = audio[lo:hi].to(self.device, non_blocking=True) if staged_on_host else audio[lo:hi]
            if taps is None and self._use_native():
                return self._encode_one_chunk(a)
            t = self.encode_features(a, taps if (lo == 0 and hi == n_all) else None)
            if taps is not None:
                taps["trans_feature"] = t
            q, idx, lvl, z = self.quantize(t, want_z=taps is not None)
            if taps is not None:
                taps["z"] = z
            return q, idx, lvl

        outs = self._run_chunks(run, chunks)
        if len(outs) == 1:
            q, idx, lvl = outs[0]
        else:
            q, idx, lvl = (torch.cat([o[i] for o in outs], dim=0) for i in range(3))
        return q, {"indices": idx, "level_indices": lvl}

    # ------------------------------------------------------------------ decode
    def en_decoder(self, feat: torch.Tensor) -> torch.Tensor:
        """LocalDecoder / CompressedLocalDecoderWithCache.forward -- l3ac/local_trans.py:86-94,123-126,182-186.
        q_trans_feature (B, T_tok, F) -> q_feature (B, T_f, F), channels-last."""
        adt = self.dec_dtype
        x = self._run_local_trans(feat, self.dec_trans_token, adt)
        if self.dec_trans_frame is not None:                                        # UpTransV2, l3ac/local_trans.py:123-126
            x = ops.upsample_linear_cn(x, self.mc.en_coder_compress_rate)
            x = self._run_local_trans(x, self.dec_trans_frame, adt)
        return x

    def conv_decoder(self, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """Decoder.forward -- l3ac/modules.py:135-201.  q_feature (B, T_f, F) channels-last -> audio (B, T)."""
        adt = self.dec_dtype
        B, T, F = x.shape
        x = self._lin(self._as_operand(x, adt), self.dec_in, B, T, F, taps=3, tap_shift0=-1)   # Conv1d(k3, pad 1)
        a_pre = None
        for si, st in enumerate(self.dec_stages):
            ch0 = []
            for ui, u in enumerate(st["units"]):
                x = self._run_conv_unit(x, u, adt, ch0_out=ch0 if ui == len(st["units"]) - 1 else None, a_pre=a_pre if ui == 0 else None)
            a_pre = None
            B, T, C = x.shape
            c_out = st["up"].w32.shape[0]
            if adt == torch.bfloat16 and self.dwconv_rows and (C, c_out) in ((48, 24), (96, 48)) and B <= 65535:
                # thin stages: EnhanceBlock gate + the 1x1 up conv in one kernel (the gated bf16 activation stays in registers)
                if st.get("enhup_plan") is None:
                    e = st["enh"]
                    st["enhup_plan"] = ops.EnhUpPlan(e["in_w"], e["in_b"], e["merge_w"], e["merge_b"], st["up"].w32, st["up"].bias, self.device)
                y = ops.enhance_up(x, st["enh"]["conv_w"], st["enh"]["conv_b"], st["enhup_plan"], ch0=ch0[0] if ch0 else None)
            else:
                a = ops.enhance(x, out_dtype=torch.float32 if adt == ops.SPLIT else adt, ch0=ch0[0] if ch0 else None, **st["enh"])    # EnhanceBlock
                y = self._lin(self._as_operand(a, adt), st["up"], B, T, C)              # Conv1d 1x1
            nxt = self.dec_stages[si + 1]["units"] if si + 1 < len(self.dec_stages) else []
            if (adt == torch.bfloat16 and self.dwconv_rows and nxt and y.shape[-1] in (48, 96) and st["stride"] in (2, 3) and B <= 65535
                    and taps is None):
                # Upsample + ChannelNorm fused with the next unit's dwconv7 + LayerNorm: x_up is written once, never re-read
                if st.get("updw_plan") is None:
                    u0 = nxt[0]
                    st["updw_plan"] = ops.UpDwPlan(st["stride"], st["cn_w"], st["cn_b"], EPS, u0["dw_w"], u0["dw_b"], u0["ln_w"], u0["ln_b"], EPS)
                x, a_pre = ops.upsample_cn_dwconv7_ln(y, st["updw_plan"])
            else:
                x = ops.upsample_linear_cn(y, st["stride"], st["cn_w"], st["cn_b"], EPS)   # Upsample + ChannelNorm
            if taps is not None:
                taps[f"dec_up{si}"] = x
        B, T, C = x.shape
        if adt == ops.SPLIT and self.dec_tail_plan is not None:                     # fp32-class fused tail (3-term split operands)
            return ops.decoder_tail_tc(x, self.dec_tail_plan, split=True)
        if self.dec_tail_fused is not None:                                         # 3 LegacyUnits + tail conv in one kernel
            if self.dec_tail_plan is not None:
                return ops.decoder_tail_tc(x, self.dec_tail_plan)
            return ops.decoder_tail(x, **self.dec_tail_fused)
        for u in self.dec_legacy:                                                   # Residual(LegacyUnit), modules.py:47-64
            d = u["dil"]
            a = self._as_operand(ops.snake(x, u["alpha0"]), adt) if adt == ops.SPLIT else ops.snake(x, u["alpha0"], out_dtype=adt)
            h = self._lin(a, u["conv"], B, T, C, taps=7, tap_shift0=-3 * d, tap_step=d, act=ops.ACT_SNAKE,
                          alpha=u["alpha1"], out_dtype=adt)
            x = self._lin(h, u["pw"], B, T, C, residual=x)
        return ops.tail_conv_tanh(x, self.dec_tail["alpha"], self.dec_tail["w"], self.dec_tail["bias"])

    def decode_features(self, feat: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """en_decoder + decoder (l3ac/__init__.py:119-120).  feat (B, T_tok, F) fp32 -> audio (B, T)."""
        if taps is None and self._use_native():
            return self.native.decode(feat=feat.contiguous())
        x = self.en_decoder(feat)
        if taps is not None:
            taps["dec_feature"] = x
        return self.conv_decoder(x, taps)

    # ------------------------------------------------------------------ stage-level entry points (EnCodec sub-modules)
    def run_stage(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """One trainable module of the reference on a whole batch (micro-batched like encode / decode):
        ``conv_encoder`` (B, T) -> (B, T_f, F); ``en_encoder`` (B, T_f, F) -> (B, T_tok, F); ``en_decoder`` (B, T_tok, F) ->
        (B, T_f, F); ``conv_decoder`` (B, T_f, F) -> (B, T).  All channels-last."""
        fn = getattr(self, name)
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        frame = math.prod(self.mc.compress_rates)
        per_item = {"conv_encoder": 1, "en_encoder": frame, "en_decoder": self.mc.hop_length, "conv_decoder": frame}[name]
        if x.shape[0] == 0:
            raise RuntimeError(f"{name} got an empty batch")
        outs = self._run_chunks(lambda lo, hi: fn(x[lo:hi].contiguous()), self._chunks(x.shape[0], x.shape[1] * per_item))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def forward_all(self, audio: torch.Tensor) -> dict:
        """EnCodec.forward -- l3ac/en_codec.py:53-72: every intermediate the reference returns, channels-last."""
        if audio.dim() != 2:
            raise RuntimeError(f"forward expects a (batch, samples) tensor, got shape {tuple(audio.shape)}")
        audio = audio.to(device=self.device, dtype=torch.float32)
        hop = self.mc.hop_length
        T0 = audio.shape[1]
        T = math.ceil(T0 / hop) * hop
        if T != T0:
            audio = torch.nn.functional.pad(audio, (0, T - T0))

        def run(lo, hi):
            feature = self.conv_encoder(audio[lo:hi])
            trans = self.en_encoder(feature)
            q, idx, lvl, _ = self.quantize(trans)
            q_feature = self.en_decoder(q)
            return feature, trans, q, idx, lvl, q_feature, self.conv_decoder(q_feature)

        outs = self._run_chunks(run, self._chunks(*audio.shape))
        keys = ("encoded_feature", "encoded_trans_feature", "quantized_trans_feature", "indices", "level_indices",
                "quantized_feature", "audio")
        return {k: (outs[0][i] if len(outs) == 1 else torch.cat([o[i] for o in outs], dim=0)) for i, k in enumerate(keys)}

    def decode(self, audio_feature: Optional[torch.Tensor] = None, indices: Optional[torch.Tensor] = None,
               taps: Optional[dict] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """L3AC.decode_audio -- l3ac/__init__.py:116-121.

        ``out`` (an extension): a pinned host tensor (B, T_tok * hop) fp32 that receives the waveform; every micro-batch is
        downloaded on its own stream as soon as it is decoded, overlapping the kernels of the others.  The returned tensor
        is ``out`` and the copies are complete (the call synchronises the device)."""
        if out is not None:
            if out.device.type != "cpu" or not out.is_pinned() or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("out must be a contiguous pinned host tensor of dtype float32")
        if audio_feature is None:
            if indices is None:
                # the reference fails inside quantizer.to_features(None) with AttributeError
                raise AttributeError("decode_audio needs audio_feature or indices ('NoneType' has no attribute 'unsqueeze')")
            if indices.shape[0] == 0:
                return torch.zeros((0, indices.shape[1] * self.mc.hop_length), device=self.device, dtype=torch.float32)
            audio_feature = self.dequantize(indices.to(self.device))
        feat = audio_feature.to(device=self.device, dtype=torch.float32).contiguous()
        B, T_tok, _ = feat.shape
        if B == 0:
            return torch.zeros((0, T_tok * self.mc.hop_length), device=self.device, dtype=torch.float32)
        if out is not None and tuple(out.shape) != (B, T_tok * self.mc.hop_length):
            raise ValueError(f"out must have shape {(B, T_tok * self.mc.hop_length)}, got {tuple(out.shape)}")

        def finish(wav):
            if out is None:
                return wav
            out.copy_(wav, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return out

        if taps is None and 0 < B * T_tok * self.mc.hop_length <= self.graph_max_samples and \
                not torch.cuda.is_current_stream_capturing():
            with torch.cuda.device(self.device):
                return finish(self._graphed(("dec", B, T_tok), lambda f: (self.decode_features(f),), feat)[0])
        chunks = self._chunks(B, T_tok * self.mc.hop_length)

        use_graphs = self._chunk_graphs(len(chunks), taps)

        def run(lo, hi):
            if use_graphs:
                wav = self._graphed(("decc", chunks.index((lo, hi)), hi - lo, T_tok), lambda f: (self.decode_features(f),), feat[lo:hi])[0]
            else:
                wav = self.decode_features(feat[lo:hi].contiguous(), taps if (lo == 0 and hi == B) else None)
            if out is not None and len(chunks) > 1:
                out[lo:hi].copy_(wav, non_blocking=True)          # on this micro-batch's stream
            return wav

        outs = self._run_chunks(run, chunks)
        if out is not None and len(chunks) > 1:
            torch.cuda.current_stream(self.device).synchronize()   # the side streams were joined into the current one
            return out
        return finish(outs[0] if len(outs) == 1 else torch.cat(outs, dim=0))
