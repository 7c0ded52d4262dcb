"""DetectorEngine: the candidate-sourcing forward (GeneralizedRCNN.inference, lvc/modeling/meta_arch/rcnn.py:177-322)
as a static-shape pipeline of liblvcb200 launches over zero-bordered channels-last bf16 planes.

    preprocess+stem gather -> 7x7 stem GEMM -> maxpool -> 33/16 bottlenecks (1x1 / 3x3(9 shifted taps) / 1x1+residual GEMMs)
    -> FPN (lateral GEMM, upsample-add, 3x3 GEMM, p6 subsample) -> RPN head (3x3 GEMM, fused 1x1 logits|deltas GEMM)
    -> top-k select + decode + NMS + merge -> fused multi-level RoIAlign -> fc1 / fc2 / fused predictor GEMMs
    -> softmax + decode + per-class NMS + top-k + postprocess

Every arithmetic step is a kernel of liblvcb200.so; torch provides device buffers, the stream and (optionally) CUDA-graph
capture of the whole sequence.  Weights come from a state dict with the reference's names; FrozenBN
(batch_norm.py:45-65) is folded into the conv weights / bias in fp32 before the bf16 cast.

Two precision modes (DESIGN.md "Precision policy"):
  * ``precision="bf16"`` (default, the throughput mode): bf16 operands and bf16 activation planes, fp32 accumulation.
  * ``precision="strict"``: the reference is fp32 end to end (wrappers.py:94-98), so every activation and weight is carried as a
    bf16 hi/lo PAIR (16 significant bits) and every dense layer is the three-term product A_hi W_hi + A_lo W_hi + A_hi W_lo on the
    same tcgen05 pipe (gemm_tc.cu, SPLIT): features, logits and scores then agree with the fp32 reference to ~1e-5 relative,
    inside BASELINE.json's 1e-3 contract, at 3x the tensor work and 2x the activation bytes.
"""
import ctypes
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from .. import _lib, ops
from ..config import DetectorConfig

STRIDES = (4, 8, 16, 32, 64)


def _fold_bn(sd, prefix, eps=1e-5):
    w = sd[prefix + ".weight"].float()
    if prefix + ".norm.weight" in sd:
        scale = sd[prefix + ".norm.weight"].float() * (sd[prefix + ".norm.running_var"].float() + eps).rsqrt()
        bias = sd[prefix + ".norm.bias"].float() - sd[prefix + ".norm.running_mean"].float() * scale
        w = w * scale.view(-1, 1, 1, 1)
    else:
        b = sd.get(prefix + ".bias")
        bias = b.float() if b is not None else torch.zeros(w.shape[0])
    return w, bias


def _to_gemm_weight(w):
    """OIHW -> [O, (kh, kw, I)]: tap-major K, matching the row-shift order of the shift-GEMM."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _dense_weight(w, device, strict, taps=1):
    """fp32 [N, taps*K] -> the GEMM's W operand: bf16, or the [hi | lo] pair layout of strict mode."""
    if strict:
        return ops.split_weight(w, taps).to(device)
    return w.contiguous().to(device=device, dtype=torch.bfloat16)


class _Conv:
    def __init__(self, sd, prefix, device, relu, strict=False):
        w, b = _fold_bn(sd, prefix)
        self.cout, self.cin, self.k, _ = w.shape
        self.w = _dense_weight(_to_gemm_weight(w), device, strict, self.k * self.k)
        self.b = b.to(device)
        self.relu = relu


class DetectorEngine:
    def __init__(self, cfg: DetectorConfig, state_dict: Dict[str, torch.Tensor], device="cuda", use_cuda_graph=False,
                 precision="bf16", max_shapes=4):
        _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LvcB200Error("DetectorEngine needs a CUDA device (no CPU fallback)")
        if precision not in ("bf16", "strict"):
            raise ValueError("precision must be 'bf16' or 'strict'")
        self.cfg, self.device = cfg, torch.device(device)
        self.use_cuda_graph = use_cuda_graph
        self.precision, self.strict = precision, precision == "strict"
        strict = self.strict
        self.max_shapes = max_shapes    # activation sets / graphs kept alive at once (LRU over (n, Hpad, Wpad, dtype))
        sd = state_dict
        dev = self.device
        bu = "backbone.bottom_up."
        # stem 7x7/2 conv as a 3x3 shift-GEMM over the 4x4 space-to-depth image (see misc.cu: stem_s2d4_kernel):
        # W4[(py*2+px)*64 + o, (ty*3+tx)*64 + (iy*4+ix)*3 + c] = w[o, c, kh, kw], kh = 4ty + iy - 2py - 1, kw = 4tx + ix - 2px - 1
        w, b = _fold_bn(sd, bu + "stem.conv1")
        w4 = torch.zeros(2, 2, 64, 3, 3, 64)
        for py in range(2):
            for px in range(2):
                for ty in range(3):
                    for iy in range(4):
                        kh = 4 * ty + iy - 2 * py - 1
                        if not 0 <= kh < 7:
                            continue
                        for tx in range(3):
                            for ix in range(4):
                                kw = 4 * tx + ix - 2 * px - 1
                                if 0 <= kw < 7:
                                    ch = (iy * 4 + ix) * 3
                                    w4[py, px, :, ty, tx, ch:ch + 3] = w[:, :, kh, kw]
        self.stem_w = _dense_weight(w4.reshape(256, 576), dev, strict, 9)
        self.stem_b = b.repeat(4).to(dev)
        self.blocks = []
        for si, nblocks in enumerate(cfg.blocks_per_stage):
            stage = si + 2
            for bi in range(nblocks):
                p = f"{bu}res{stage}.{bi}."
                blk = dict(stage=stage, stride=2 if (bi == 0 and stage > 2) else 1,
                           conv1=_Conv(sd, p + "conv1", dev, True, strict), conv2=_Conv(sd, p + "conv2", dev, True, strict),
                           conv3=_Conv(sd, p + "conv3", dev, True, strict),
                           shortcut=_Conv(sd, p + "shortcut", dev, False, strict) if (p + "shortcut.weight") in sd else None,
                           last=bi == nblocks - 1)
                self.blocks.append(blk)
        self.lateral = {l: _Conv(sd, f"backbone.fpn_lateral{l}", dev, False, strict) for l in (2, 3, 4, 5)}
        self.fpn_out = {l: _Conv(sd, f"backbone.fpn_output{l}", dev, False, strict) for l in (2, 3, 4, 5)}
        rp = "proposal_generator.rpn_head."
        self.rpn_conv = _Conv(sd, rp + "conv", dev, True, strict)
        A = len(cfg.anchor_ratios)
        self.A = A
        wh = torch.zeros(16, 256)
        bh = torch.zeros(16)
        wh[:A] = sd[rp + "objectness_logits.weight"].float().view(A, 256)
        wh[A:A + 4 * A] = sd[rp + "anchor_deltas.weight"].float().view(4 * A, 256)
        bh[:A] = sd[rp + "objectness_logits.bias"].float()
        bh[A:A + 4 * A] = sd[rp + "anchor_deltas.bias"].float()
        self.rpn_head_w, self.rpn_head_b = _dense_weight(wh, dev, strict), bh.to(dev)
        # box head: fc1 input index c*49+h*7+w (box_head.py:86-87) -> permuted to the pooler's (h, w, c) order
        res = cfg.pooler_resolution
        self.fcs = []
        # models without a detection box head (ProposalNetwork, the box corrector's GeneralizedRCNNRegOnly) stop after the RPN / FPN
        self.has_box_head = "roi_heads.box_predictor.cls_score.weight" in sd
        for i in range(cfg.num_fc if self.has_box_head else 0):
            w = sd[f"roi_heads.box_head.fc{i + 1}.weight"].float()
            if i == 0:
                w = w.view(-1, 256, res, res).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
            self.fcs.append((_dense_weight(w, dev, strict), sd[f"roi_heads.box_head.fc{i + 1}.bias"].float().to(dev), w.shape[0]))
        K = cfg.num_classes
        self.cls_cols = (K + 1 + 15) // 16 * 16
        self.cosine = cfg.output_layer == "CosineSimOutputLayers"
        if self.has_box_head:
            wc = sd["roi_heads.box_predictor.cls_score.weight"].float()
            if self.cosine:  # fast_rcnn.py:830-837 (first forward of a freshly loaded model)
                wc = wc / (wc.norm(p=2, dim=1, keepdim=True) + 1e-5)
            wb = sd["roi_heads.box_predictor.bbox_pred.weight"].float()
            wp = torch.zeros(self.cls_cols + 4 * K, wc.shape[1])
            bp = torch.zeros(self.cls_cols + 4 * K)
            wp[:K + 1] = wc
            wp[self.cls_cols:] = wb
            if not self.cosine:
                bp[:K + 1] = sd["roi_heads.box_predictor.cls_score.bias"].float()
            bp[self.cls_cols:] = sd["roi_heads.box_predictor.bbox_pred.bias"].float()
            self.pred_w, self.pred_b, self.pred_cols = _dense_weight(wp, dev, strict), bp.to(dev), wp.shape[0]
        self.mean = torch.tensor(cfg.pixel_mean, dtype=torch.float32, device=dev)
        self.inv_std = (1.0 / torch.tensor(cfg.pixel_std, dtype=torch.float32)).to(dev)
        # res stages run as layer-chain launches (gemm_chain.cu).  res2 stays on per-layer launches: its N = 64 layers issue one tiny
        # MMA group per 24 KB operand block, which the leaner single-layer producer loop feeds faster (profiles/r01_gemm_layers_*.md)
        # strict mode runs every layer as its own SPLIT launch (the chain / 2-CTA / fused-upsample variants are bf16-mode kernels)
        self.use_chain = os.environ.get("LVCB200_CHAIN", "1") != "0" and not strict
        self.fuse_upsample = os.environ.get("LVCB200_FUSE_UPSAMPLE", "1") != "0" and not strict   # same-box A/B: equal in burst, -1.0 % sustained (0.8 GB less HBM traffic per step)
        self.chain_stages = tuple(int(x) for x in os.environ.get("LVCB200_CHAIN_STAGES", "3,4,5").split(",") if x)
        # per input-shape state (activation buffers, staging, CUDA graph, layer-chain plans), LRU-bounded: real datasets yield many
        # (Hpad, Wpad) and an activation set is ~0.8 GB per R101 image
        self._states = OrderedDict()
        self._cur = None
        self.debug = None  # set to a dict to capture intermediates (tests)

    # ------------------------------------------------------------------ buffers
    def _buf(self, name, shape, dtype=torch.bfloat16, zero=False):
        bufs = self._cur["bufs"]
        key = (name, tuple(shape), dtype)
        t = bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            bufs[key] = t
        return t

    def _plane(self, name, n, H, W, C, dtype=torch.bfloat16):
        if self.strict and dtype == torch.bfloat16:
            S = ops.PairPlane.rows_for(n, H, W)
            return ops.PairPlane(self._buf(name, (2 * S, C), dtype, zero=True), n, H, W, C)
        return ops.Plane(self._buf(name, (n, H + 2, W + 2, C), dtype), H, W, C)

    # ------------------------------------------------------------------ layers
    def _conv(self, name, x: ops.Plane, conv: _Conv, residual: Optional[ops.Plane] = None, out_dtype=torch.bfloat16,
              upsample_add: Optional[ops.Plane] = None):
        n = x.n
        out = self._plane(name, n, x.H, x.W, conv.cout, out_dtype)
        PW = x.PW
        if conv.k == 3:
            shifts = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
            taps = 9
        else:
            shifts, taps = (0,), 1
        if self.strict:
            ops.gemm(x.full, conv.w, bias=conv.b, residual=residual.full if residual is not None else None, out=out.full,
                     relu=conv.relu, taps=taps, shifts=shifts, K=conv.cin, M=x.M, plane_hw=(x.PH, x.PW), split_rows=x.split_rows)
            return out
        ops.gemm(x.t.view(-1, x.C), conv.w, bias=conv.b, residual=residual.t.view(-1, conv.cout) if residual is not None else None,
                 out=out.t.view(-1, conv.cout), relu=conv.relu, taps=taps, shifts=shifts, K=conv.cin, plane_hw=(x.PH, x.PW),
                 upsample_add=upsample_add)
        return out

    def _subsample(self, name, x: ops.Plane):
        Ho, Wo = (x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1
        out = self._plane(name, x.n, Ho, Wo, x.C)
        lib = _lib.load()
        _lib.check(lib.lvcb200_subsample2(_lib.ptr(x.t), x.n, x.H, x.W, x.C, _lib.ptr(out.t), _lib.stream_ptr()), "subsample2")
        if self.strict:   # a pure copy: the lo half moves the same way
            _lib.check(lib.lvcb200_subsample2(_lib.ptr(x.lo), x.n, x.H, x.W, x.C, _lib.ptr(out.lo), _lib.stream_ptr()), "subsample2")
        return out

    # ------------------------------------------------------------------ forward pieces
    def backbone(self, img_ptrs, img_dtype, sizes_dev, n, Hpad, Wpad):
        lib = _lib.load()
        H4, W4 = Hpad // 4, Wpad // 4
        x4 = self._plane("stem_x4", n, H4, W4, 64)
        s2 = self._plane("stem_s2", n, H4, W4, 256)     # stem output, 2x2 pixels per cell: channel ((Y&1)*2 + (X&1))*64 + o
        x = self._plane("pool", n, H4, W4, 64)
        PW = x4.PW
        shifts = [(ty - 1) * PW + (tx - 1) for ty in range(3) for tx in range(3)]
        if self.strict:
            _lib.check(lib.lvcb200_stem_s2d4_pair(_lib.ptr(img_ptrs), img_dtype, _lib.ptr(sizes_dev), n, Hpad, Wpad, _lib.ptr(self.mean),
                                                  _lib.ptr(self.inv_std), _lib.ptr(x4.full), x4.split_rows, _lib.stream_ptr()), "stem_s2d4_pair")
            ops.gemm(x4.full, self.stem_w, bias=self.stem_b, out=s2.full, relu=True, taps=9, shifts=shifts, K=64, M=x4.M,
                     plane_hw=(x4.PH, x4.PW), split_rows=x4.split_rows)
            _lib.check(lib.lvcb200_maxpool_s2d_pair(_lib.ptr(s2.full), s2.split_rows, n, H4, W4, 64, _lib.ptr(x.full), x.split_rows,
                                                    _lib.stream_ptr()), "maxpool_s2d_pair")
        else:
            _lib.check(lib.lvcb200_stem_s2d4(_lib.ptr(img_ptrs), img_dtype, _lib.ptr(sizes_dev), n, Hpad, Wpad, _lib.ptr(self.mean),
                                             _lib.ptr(self.inv_std), _lib.ptr(x4.t), _lib.stream_ptr()), "stem_s2d4")
            ops.gemm(x4.t.view(-1, 64), self.stem_w, bias=self.stem_b, out=s2.t.view(-1, 256), relu=True, taps=9, shifts=shifts, K=64,
                     plane_hw=(x4.PH, x4.PW))
            _lib.check(lib.lvcb200_maxpool_s2d(_lib.ptr(s2.t), n, H4, W4, 64, _lib.ptr(x.t), _lib.stream_ptr()), "maxpool_s2d")
        if self.debug is not None:
            self.debug["stem_pool"] = x
        feats = {}
        i = 0
        while i < len(self.blocks):
            # one res stage = ONE layer-chain launch (tile-granular dependencies between its 10-70 GEMMs, gemm_chain.cu); the
            # stride-2 subsample feeding the stage's first block runs before it as its own kernel
            stage = self.blocks[i]["stage"]
            with ops.gemm_chain(self.use_chain and stage in self.chain_stages):
                while i < len(self.blocks) and self.blocks[i]["stage"] == stage:
                    blk, tag = self.blocks[i], f"b{i}"
                    xin = self._subsample(tag + "_sub", x) if blk["stride"] == 2 else x
                    o1 = self._conv(tag + "_c1", xin, blk["conv1"])
                    o2 = self._conv(tag + "_c2", o1, blk["conv2"])
                    sc = self._conv(tag + "_sc", xin, blk["shortcut"]) if blk["shortcut"] is not None else x
                    x = self._conv(tag + "_c3", o2, blk["conv3"], residual=sc)
                    i += 1
            feats[stage] = x
        return feats

    def fpn(self, feats):
        lib = _lib.load()
        out = {}
        prev = None
        for l in (5, 4, 3, 2):
            # top-down path (fpn.py:128-134): the nearest-2x upsampling of the coarser level is added in the lateral conv's epilogue
            # (one rounding to bf16); LVCB200_FUSE_UPSAMPLE=0 keeps the separate read-modify-write kernel
            fuse = prev is not None and self.fuse_upsample
            lat = self._conv(f"lat{l}", feats[l], self.lateral[l], upsample_add=prev if fuse else None)
            if prev is not None and self.strict:
                _lib.check(lib.lvcb200_upsample2_add_pair(_lib.ptr(prev.full), prev.split_rows, prev.n, prev.H, prev.W, 256, _lib.ptr(lat.full),
                                                          lat.split_rows, lat.H, lat.W, _lib.stream_ptr()), "upsample2_add_pair")
            elif prev is not None and not fuse:
                _lib.check(lib.lvcb200_upsample2_add(_lib.ptr(prev.t), prev.n, prev.H, prev.W, 256, _lib.ptr(lat.t), lat.H, lat.W,
                                                     _lib.stream_ptr()), "upsample2_add")
            prev = lat
            out[l] = self._conv(f"p{l}", lat, self.fpn_out[l])
        out[6] = self._subsample("p6", out[5])
        return out

    def rpn(self, pyramid, sizes_dev):
        cfg = self.cfg
        levels = []
        for l in (2, 3, 4, 5, 6):
            p = pyramid[l]
            t = self._conv(f"rpn_t{l}", p, self.rpn_conv)
            head = self._buf(f"rpn_h{l}", (p.n * p.PH * p.PW, 16), torch.float32)
            if self.strict:
                ops.gemm(t.full, self.rpn_head_w, bias=self.rpn_head_b, out=head, K=256, M=t.M, split_rows=t.split_rows)
            else:
                ops.gemm(t.t.view(-1, 256), self.rpn_head_w, bias=self.rpn_head_b, out=head, K=256)
            off = (p.PW + 1) * 16
            levels.append(dict(logits=head, deltas=head, H=p.H, W=p.W, A=self.A, offset_l=off, offset_d=off + self.A,
                               strides_l=(p.PH * p.PW * 16, p.PW * 16, 16), strides_d=(p.PH * p.PW * 16, p.PW * 16, 16)))
        if self.debug is not None:
            self.debug["rpn_levels"] = levels
        return ops.rpn_proposals(levels, sizes_dev, cfg.anchor_sizes, cfg.anchor_ratios, STRIDES, cfg.rpn_pre_nms_topk,
                                 cfg.rpn_post_nms_topk, cfg.rpn_nms_thresh, cfg.rpn_min_box_size, cfg.rpn_bbox_weights)

    def roi_heads(self, pyramid, props, counts, sizes_dev, out_sizes_dev):
        cfg = self.cfg
        n, P = props.shape[0], props.shape[1]
        R = n * P
        rois, roi_image = ops.make_rois(props, counts, self._buf("rois", (R, 5), torch.float32), self._buf("roi_image", (R,), torch.int32))
        scales = [1.0 / s for s in STRIDES[:4]]
        row_scale = None
        if self.strict:
            # pooler in fp32 on the merged (hi + lo) planes, then back to a pair for the FC layers
            planes = []
            for l in (2, 3, 4, 5):
                pl = pyramid[l]
                planes.append(ops.Plane(pl.merged(self._buf(f"p{l}_f32", tuple(pl.t.shape), torch.float32)), pl.H, pl.W, pl.C))
            pooled = ops.roi_pool_fpn(planes, scales, rois, cfg.pooler_resolution, cfg.pooler_sampling_ratio,
                                      out_dtype=torch.float32, out_layout=ops.OUT_NHWC)
            S = (R + 127) // 128 * 128
            x, _ = ops.pair_split(pooled.view(R, -1), self._buf("pooled_pair", (2 * S, pooled[0].numel()), zero=True))
            for i, (w, b, nout) in enumerate(self.fcs):
                x = ops.gemm(x, w, bias=b, relu=True, out=self._buf(f"fc{i}", (2 * S, nout), zero=True), M=R, split_rows=S)
            pred = ops.gemm(x, self.pred_w, bias=self.pred_b, out=self._buf("pred", (R, self.pred_cols), torch.float32), M=R, split_rows=S)
            if self.cosine:  # scores = scale * (x / (|x| + 1e-5)) . w_hat   (fast_rcnn.py:826-840)
                row_scale = ops.row_inv_norm(x, cfg.cosine_scale, 1e-5, lo_off=S * x.shape[1], rows=R, out=self._buf("row_scale", (R,), torch.float32))
            head_dbg = (x[:R].float() + x[S:S + R].float()) if self.debug is not None else None
        else:
            planes = [pyramid[l] for l in (2, 3, 4, 5)]
            pooled = ops.roi_pool_fpn(planes, scales, rois, cfg.pooler_resolution, cfg.pooler_sampling_ratio,
                                      out_dtype=torch.bfloat16, out_layout=ops.OUT_NHWC)
            x = pooled.view(R, -1)
            for i, (w, b, nout) in enumerate(self.fcs):
                x = ops.gemm(x, w, bias=b, relu=True, out=self._buf(f"fc{i}", (R, nout)))
            pred = ops.gemm(x, self.pred_w, bias=self.pred_b, out=self._buf("pred", (R, self.pred_cols), torch.float32))
            if self.cosine:
                row_scale = ops.row_inv_norm(x, cfg.cosine_scale, 1e-5, out=self._buf("row_scale", (R,), torch.float32))
            head_dbg = x
        K = cfg.num_classes
        if self.debug is not None:
            self.debug.update(pooled=pooled, head=head_dbg, pred=pred, rois=rois, roi_image=roi_image, row_scale=row_scale)
        return ops.detections(pred[:, : K + 1], pred[:, self.cls_cols:], props.view(-1, 4), roi_image, sizes_dev, out_sizes_dev, K,
                              max_rois_per_image=P, weights=cfg.roi_bbox_weights, score_thresh=cfg.score_thresh_test,
                              nms_thresh=cfg.nms_thresh_test, topk=cfg.detections_per_image, row_scale=row_scale)

    # ------------------------------------------------------------------ whole forward on device-resident inputs
    def forward_device(self, img_ptrs, img_dtype, sizes_dev, out_sizes_dev, n, Hpad, Wpad):
        feats = self.backbone(img_ptrs, img_dtype, sizes_dev, n, Hpad, Wpad)
        pyramid = self.fpn(feats)
        props, plogits, counts = self.rpn(pyramid, sizes_dev)
        if self.debug is not None:
            self.debug.update(feats=feats, pyramid=pyramid, props=props, prop_logits=plogits, prop_counts=counts)
        return self.roi_heads(pyramid, props, counts, sizes_dev, out_sizes_dev)

    def run_features(self, images: List[torch.Tensor]):
        """Backbone + FPN only (eager): returns (pyramid {2..6: Plane}, image sizes on the device).  Used by the box corrector
        (GeneralizedRCNNRegOnly.inference, rcnn.py:340-377) and ProposalNetwork."""
        st, args = self._prepare(images, None)
        ptrs, img_dtype, sizes_dev, _, n, Hpad, Wpad = args
        ops.PLAN_LOG = st["plan_keys"]
        try:
            return self.fpn(self.backbone(ptrs, img_dtype, sizes_dev, n, Hpad, Wpad)), sizes_dev
        finally:
            ops.PLAN_LOG = None

    def run_proposals(self, images: List[torch.Tensor]):
        """Backbone + FPN + RPN (ProposalNetwork.forward, rcnn.py:413-488): proposals [n,P,4], logits [n,P], counts [n]."""
        pyramid, sizes_dev = self.run_features(images)
        return self.rpn(pyramid, sizes_dev)

    def run(self, images: List[torch.Tensor], out_sizes=None):
        """images: list of [3,H,W] CUDA tensors (BGR, 0..255), all fp32 or all uint8 (what the reference's DatasetMapper yields).
        Returns (boxes [n,100,4], scores, classes, rows, counts)."""
        if not self.has_box_head:
            raise _lib.LvcB200Error("this engine was built without a detection box head (use run_features / run_proposals)")
        st, args = self._prepare(images, out_sizes)
        if not self.use_cuda_graph or self.debug is not None:
            return self._forward_logged(st, *args)
        if st["graph"] is None:
            if st["warm"] < 1:   # first call eager: allocates every buffer, sets kernel attributes, builds the layer-chain plans
                st["warm"] += 1
                return self._forward_logged(st, *args)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                st["result"] = self.forward_device(*args)
            st["graph"] = g
        st["graph"].replay()
        return st["result"]

    def _prepare(self, images, out_sizes):
        _lib.require_cuda(*images)
        cfg = self.cfg
        n = len(images)
        if all(im.dtype == torch.uint8 for im in images):
            images, img_dtype = [im.contiguous() for im in images], _lib.U8
        else:
            images, img_dtype = [im.contiguous().float() for im in images], _lib.F32
        sizes = [tuple(im.shape[-2:]) for im in images]
        d = cfg.size_divisibility
        Hpad = (max(s[0] for s in sizes) + d - 1) // d * d
        Wpad = (max(s[1] for s in sizes) + d - 1) // d * d
        out_sizes = out_sizes or sizes
        key = (n, Hpad, Wpad, img_dtype)
        st = self._state(key)
        # per-call metadata (image pointers, sizes, output sizes): one small async H2D from a ring of pinned staging buffers (a slot is
        # rewritten only after the copy that read it has completed: back-to-back run() calls never see each other's metadata), skipped
        # when nothing changed; the int32 views the kernels read are refreshed on the stream
        meta = [[im.data_ptr(), s[0], s[1], o[0], o[1]] for im, s, o in zip(images, sizes, out_sizes)]
        if meta != st["last"]:
            slot = st["ring"][st["ring_pos"] % len(st["ring"])]
            st["ring_pos"] += 1
            if slot[1] is not None:
                slot[1].synchronize()
            slot[0].copy_(torch.tensor(meta, dtype=torch.int64))
            st["meta"].copy_(slot[0], non_blocking=True)
            slot[1] = torch.cuda.current_stream(self.device).record_event()
            st["sizes"].copy_(st["meta"][:, 1:3])
            st["outs"].copy_(st["meta"][:, 3:5])
            st["ptrs_c"].copy_(st["meta"][:, 0])
            st["last"] = meta
        self._keepalive = images
        return st, (st["ptrs_c"], img_dtype, st["sizes"], st["outs"], n, Hpad, Wpad)

    def _forward_logged(self, st, *args):
        ops.PLAN_LOG = st["plan_keys"]
        try:
            return self.forward_device(*args)
        finally:
            ops.PLAN_LOG = None

    def _state(self, key):
        st = self._states.get(key)
        if st is None:
            n = key[0]
            while len(self._states) >= self.max_shapes:      # evict the least recently used shape: buffers, graph, chain plans
                _, old = self._states.popitem(last=False)
                torch.cuda.synchronize(self.device)
                ops.drop_plans(old["plan_keys"])
                old.clear()
            st = dict(bufs={}, plan_keys=[], last=None, graph=None, result=None, warm=0, ring_pos=0,
                      ring=[[torch.zeros((n, 5), dtype=torch.int64).pin_memory(), None] for _ in range(4)],
                      meta=torch.zeros((n, 5), dtype=torch.int64, device=self.device),
                      sizes=torch.zeros((n, 2), dtype=torch.int32, device=self.device),
                      outs=torch.zeros((n, 2), dtype=torch.int32, device=self.device),
                      ptrs_c=torch.zeros(n, dtype=torch.int64, device=self.device))   # persistent: captured by the CUDA graph
            self._states[key] = st
        else:
            self._states.move_to_end(key)
        self._cur = st
        return st
