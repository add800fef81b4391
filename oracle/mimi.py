"""Oracle (test infrastructure): Mimi streaming encode / decode on the CPU, fp32.

Restates the reference's ``MimiModel`` data path with explicit per-layer state:

* causal conv with carried left context      – ``modules/conv.py:245-274``
* transposed conv with overlap-add carry     – ``modules/conv.py:340-362``
* SEANet encoder / decoder / residual block  – ``modules/seanet.py:90-93, 170-239, 323-392``
* learnt 2x down / (depth-wise) up-sampling  – ``modules/resample.py:14-119``
* transformer bottlenecks                    – see ``oracle/transformer.py``
* split residual VQ                          – ``quantization/core_vq.py:178-186, 270-297, 507-528``,
                                               ``quantization/vq.py:126-151, 269-287``
* orchestration                              – ``models/compression.py:338-429``

State-dict keys are the reference's.
"""
from __future__ import annotations

import typing as tp
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import transformer as tr


@dataclass
class ConvSpec:
    key: str          # state-dict prefix of the nn.Conv1d / nn.ConvTranspose1d
    kind: str         # conv | convtr
    stride: int = 1
    dilation: int = 1
    groups: int = 1
    replicate: bool = False   # pad_mode == "replicate" (down-sampling conv only)
    elu_before: bool = False  # the nn.ELU that precedes this conv in the nn.Sequential


class MimiOracle:
    def __init__(self, sd: tp.Dict[str, torch.Tensor], cfg):
        self.sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        self.cfg = cfg
        self.num_codebooks = cfg.num_codebooks
        self.tr_spec = tr.TransformerSpec(
            d_model=cfg.tr_d_model, num_heads=cfg.tr_num_heads, num_layers=cfg.tr_num_layers,
            context=cfg.tr_context, norm="layer_norm", gating="none", positional_embedding="rope",
            max_period=cfg.tr_max_period, layer_scale=True)
        self._centroid_cache: dict[str, torch.Tensor] = {}
        self._plan()
        self.batch: int | None = None
        self.trace: dict[str, torch.Tensor] | None = None   # set to {} to record per-module outputs

    # ------------------------------------------------------------------ structure
    def _plan(self) -> None:
        """Ordered op list of the SEANet encoder and decoder (seanet.py:170-236 / 323-388)."""
        cfg = self.cfg
        enc: list = []
        idx = 0
        enc.append(("conv", ConvSpec(f"encoder.model.{idx}.conv.conv", "conv")))
        idx += 1
        for ratio in reversed(cfg.ratios):
            for j in range(cfg.n_residual_layers):
                enc.append(("res", f"encoder.model.{idx}", cfg.dilation_base ** j))
                idx += 1
            idx += 1
            enc.append(("conv", ConvSpec(f"encoder.model.{idx}.conv.conv", "conv", stride=ratio,
                                         elu_before=True)))
            idx += 1
        idx += 1
        enc.append(("conv", ConvSpec(f"encoder.model.{idx}.conv.conv", "conv", elu_before=True)))
        dec: list = []
        idx = 0
        dec.append(("conv", ConvSpec(f"decoder.model.{idx}.conv.conv", "conv")))
        idx += 1
        for ratio in cfg.ratios:
            idx += 1
            dec.append(("conv", ConvSpec(f"decoder.model.{idx}.convtr.convtr", "convtr", stride=ratio,
                                         elu_before=True)))
            idx += 1
            for j in range(cfg.n_residual_layers):
                dec.append(("res", f"decoder.model.{idx}", cfg.dilation_base ** j))
                idx += 1
        idx += 1
        dec.append(("conv", ConvSpec(f"decoder.model.{idx}.conv.conv", "conv", elu_before=True)))
        self.enc_plan, self.dec_plan = enc, dec
        s = cfg.resample_stride
        self.down_spec = ConvSpec("downsample.conv.conv.conv", "conv", stride=s, replicate=True)
        self.up_spec = ConvSpec("upsample.convtr.convtr.convtr", "convtr", stride=s, groups=cfg.dimension)

    # ------------------------------------------------------------------ state
    def streaming(self, batch: int) -> None:
        """``MimiModel.streaming(B)`` – fresh state for every streaming child (streaming.py:131-137)."""
        self.batch = batch
        self.exec_mask = torch.ones(batch, dtype=torch.bool)
        self.conv_state: dict[str, dict] = {}
        self.enc_tr = tr.init_state(self.tr_spec, batch, torch.float32)
        self.dec_tr = tr.init_state(self.tr_spec, batch, torch.float32)

    def stop_streaming(self) -> None:
        self.batch = None

    def set_exec_mask(self, mask: torch.Tensor) -> None:
        """streaming.py:183-211."""
        self.exec_mask[:] = mask
        self.enc_tr.exec_mask[:] = mask
        self.dec_tr.exec_mask[:] = mask

    def reset_streaming(self, reset_mask: torch.Tensor | None = None) -> None:
        """streaming.py:139-156 with the per-module ``reset`` of conv.py:166-169, 281-286."""
        assert self.batch is not None
        if reset_mask is None:
            reset_mask = torch.ones(self.batch, dtype=torch.bool)
        self.exec_mask |= reset_mask
        for st in self.conv_state.values():
            if "previous" in st:
                st["previous"][reset_mask] = 0
                st["first"][reset_mask] = True
            else:
                st["partial"][reset_mask] = 0
        tr.reset_state(self.enc_tr, reset_mask)
        tr.reset_state(self.dec_tr, reset_mask)

    # ------------------------------------------------------------------ convolutions
    def _conv(self, spec: ConvSpec, x: torch.Tensor) -> torch.Tensor:
        """``StreamingConv1d.forward`` (conv.py:245-274)."""
        w = self.sd[spec.key + ".weight"]
        b = self.sd.get(spec.key + ".bias")
        k_eff = (w.shape[-1] - 1) * spec.dilation + 1
        tp_ = k_eff - spec.stride
        B, C, T = x.shape
        assert T > 0 and T % spec.stride == 0
        st = self.conv_state.get(spec.key)
        if st is None:
            st = {"previous": torch.zeros(B, C, tp_), "first": torch.ones(B, dtype=torch.bool)}
            self.conv_state[spec.key] = st
        m = self.exec_mask.view(-1, 1, 1)
        if tp_ and spec.replicate:
            assert T >= tp_
            st["previous"] = torch.where(st["first"].view(-1, 1, 1) & m, x[..., :1], st["previous"])
        if tp_:
            x = torch.cat([st["previous"], x], dim=-1)
        y = F.conv1d(x, w, b, stride=spec.stride, dilation=spec.dilation, groups=spec.groups)
        if tp_:
            st["previous"] = torch.where(m, x[..., -tp_:], st["previous"])
            if spec.replicate:
                st["first"] = torch.where(self.exec_mask, torch.zeros_like(st["first"]), st["first"])
        return y

    def _convtr(self, spec: ConvSpec, x: torch.Tensor) -> torch.Tensor:
        """``StreamingConvTranspose1d.forward`` (conv.py:340-362)."""
        w = self.sd[spec.key + ".weight"]
        b = self.sd.get(spec.key + ".bias")
        K, S = w.shape[-1], spec.stride
        B = x.shape[0]
        cout = w.shape[1] * spec.groups
        st = self.conv_state.get(spec.key)
        if st is None:
            st = {"partial": torch.zeros(B, cout, K - S)}
            self.conv_state[spec.key] = st
        y = F.conv_transpose1d(x, w, b, stride=S, groups=spec.groups)
        pt = K - S
        if pt > 0:
            y[..., :pt] += st["partial"]
            tail = y[..., -pt:]
            if b is not None:
                tail = tail - b[:, None]
            st["partial"] = torch.where(self.exec_mask.view(-1, 1, 1), tail, st["partial"])
            y = y[..., :-pt]
        return y

    def _run(self, plan: list, x: torch.Tensor) -> torch.Tensor:
        for item in plan:
            if item[0] == "res":
                # SEANetResnetBlock with true_skip: x + conv1(elu(conv3(elu(x)))) (seanet.py:56-93)
                _, base, dil = item
                h = self._conv(ConvSpec(base + ".block.1.conv.conv", "conv", dilation=dil), F.elu(x))
                h = self._conv(ConvSpec(base + ".block.3.conv.conv", "conv"), F.elu(h))
                x = x + h
                name = base
            else:
                spec = item[1]
                if spec.elu_before:
                    x = F.elu(x)
                x = self._conv(spec, x) if spec.kind == "conv" else self._convtr(spec, x)
                name = spec.key
            if self.trace is not None:   # "encoder.model.3.conv.conv" -> "enc.3"
                side, _, idx = name.split(".")[:3]
                self.trace[f"{side[:3]}.{idx}"] = x
        return x

    # ------------------------------------------------------------------ quantizer
    def _centroids(self, prefix: str) -> torch.Tensor:
        """``EuclideanCodebook.embedding`` (core_vq.py:178-186)."""
        c = self._centroid_cache.get(prefix)
        if c is None:
            c = self.sd[prefix + ".embedding_sum"] / self.sd[prefix + ".cluster_usage"].clamp(min=1e-5)[:, None]
            self._centroid_cache[prefix] = c
        return c

    def _levels(self) -> list[tuple[str, int]]:
        n_sem = self.cfg.q_n_semantic
        return [("rvq_first", n_sem), ("rvq_rest", self.num_codebooks - n_sem)]

    def quantize(self, latent: torch.Tensor, return_margins: bool = False):
        """``SplitResidualVectorQuantizer.encode`` (vq.py:269-279) -> int64 [B, K, T].

        With ``return_margins`` also returns, per code, the relative gap between the best and the
        second-best squared distance (used by the margin-aware comparator in the tests).
        """
        codes, margins = [], []
        for name, n_levels in self._levels():
            p = f"quantizer.{name}"
            res = F.conv1d(latent, self.sd[p + ".input_proj.weight"])          # vq.py:135
            for level in range(n_levels):
                cb = self._centroids(f"{p}.vq.layers.{level}._codebook")
                flat = res.transpose(1, 2).reshape(-1, res.shape[1])
                dists = torch.cdist(flat[None], cb[None], p=2)[0]                # core_vq.py:274
                idx = dists.argmin(dim=-1)
                if return_margins:
                    two = torch.topk(dists.double() ** 2, 2, dim=-1, largest=False).values
                    margins.append(((two[:, 1] - two[:, 0]) / two[:, 1].clamp(min=1e-30))
                                   .view(res.shape[0], res.shape[2]))
                idx = idx.view(res.shape[0], res.shape[2])
                res = res - F.embedding(idx, cb).transpose(1, 2)                 # core_vq.py:514-516
                codes.append(idx)
        out = torch.stack(codes, dim=1)
        if return_margins:
            return out, torch.stack(margins, dim=1)
        return out

    def dequantize(self, codes: torch.Tensor) -> torch.Tensor:
        """``SplitResidualVectorQuantizer.decode`` (vq.py:281-287) -> fp32 [B, 512, T]."""
        n_sem = self.cfg.q_n_semantic
        out = None
        for name, lo, hi in (("rvq_first", 0, n_sem), ("rvq_rest", n_sem, codes.shape[1])):
            if hi <= lo:
                continue
            p = f"quantizer.{name}"
            q = None
            for level in range(hi - lo):
                cb = self._centroids(f"{p}.vq.layers.{level}._codebook")
                e = F.embedding(codes[:, lo + level], cb).transpose(1, 2)
                q = e if q is None else q + e
            q = F.conv1d(q, self.sd[p + ".output_proj.weight"])
            out = q if out is None else out + q
        return out

    # ------------------------------------------------------------------ model
    def _transformer(self, which: str, st: tr.TransformerState, x: torch.Tensor) -> torch.Tensor:
        """``ProjectedTransformer.forward`` with conv_layout (transformer.py:971-983)."""
        y = tr.forward(self.sd, f"{which}.transformer", self.tr_spec, x.transpose(1, 2), st)
        return y.transpose(1, 2)

    def encode_to_latent(self, pcm: torch.Tensor) -> torch.Tensor:
        """``_encode_to_unquantized_latent`` in streaming mode (compression.py:338-374)."""
        assert self.batch is not None, "call streaming(B) first"
        fs = self.cfg.frame_size
        if pcm.shape[-1] % fs != 0 or pcm.shape[-1] == 0:
            raise RuntimeError(f"Invalid input x of length {pcm.shape[-1]}.")
        emb = self._run(self.enc_plan, pcm)
        emb = self._transformer("encoder_transformer", self.enc_tr, emb)
        lat = self._conv(self.down_spec, emb)
        if self.trace is not None:
            self.trace["enc.tr"] = emb
            self.trace["enc.latent"] = lat
        return lat

    def encode(self, pcm: torch.Tensor) -> torch.Tensor:
        return self.quantize(self.encode_to_latent(pcm))

    def decode_latent(self, latent: torch.Tensor) -> torch.Tensor:
        up = self._convtr(self.up_spec, latent)
        emb = self._transformer("decoder_transformer", self.dec_tr, up)
        if self.trace is not None:
            self.trace["dec.latent"] = latent
            self.trace["dec.up"] = up
            self.trace["dec.tr"] = emb
        return self._run(self.dec_plan, emb)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """``MimiModel.decode`` (compression.py:406-429)."""
        return self.decode_latent(self.dequantize(codes))
