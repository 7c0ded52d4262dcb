"""Torch-tensor front end of the fused device ops in liblvcb200.so (device memory + streams come from torch;
the arithmetic is all in the library).  Everything here requires CUDA tensors and raises otherwise."""
import ctypes
import math

import torch

from . import _lib
from ._lib import BF16, F32, OUT_NCHW, OUT_NHWC, ChainPlan, DetParams, FMap, GemmDesc, RpnLevel, RpnParams

_ws = {}
_ws_retired = []     # outgrown scratch buffers stay alive: a captured CUDA graph may still hold their raw pointers
PLAN_LOG = None      # an engine sets this to a list to learn which layer-chain plans its buffers own (evicted with them)
GEMM_EVENTS = None   # bench.py sets this to a list to time every GEMM launch with CUDA events on the launch stream
GEMM_RECORD = None   # bench.py sets this to a list to record every GEMM descriptor of a step (replayed alone inside a CUDA graph)


GEMM_CHAIN = None    # inside `with gemm_chain():` gemm() calls are collected here and issued as ONE layer-chain launch


class gemm_chain:
    """Context manager: every `gemm()` issued inside is recorded instead of launched; on exit the recorded layers run as one
    persistent launch with tile-granular dependencies (lvcb200_gemm_chain_*).  Plans (tensor maps + layer table in a device
    workspace) are cached per exact descriptor list, so steady-state calls -- and CUDA-graph capture -- only enqueue the run.
    Layers the chain kernel does not take (fp32 heads, N or K not a multiple of 64) make it fall back to per-layer launches."""
    _plans = {}

    def __init__(self, enabled=True):
        self.enabled = enabled

    def __enter__(self):
        global GEMM_CHAIN
        if self.enabled:
            assert GEMM_CHAIN is None, "gemm_chain contexts do not nest"
            GEMM_CHAIN = []
        return self

    def __exit__(self, et, ev, tb):
        global GEMM_CHAIN
        if not self.enabled:
            return False
        rec, GEMM_CHAIN = GEMM_CHAIN, None
        if et is None and rec:
            run_chain(rec)
        return False


def _desc_key(d):
    return (d.a_dtype, d.A, d.lda, d.M_rows, d.W, d.ldw, d.bias, d.residual, d.ldr, d.D, d.ldd, d.d_dtype, d.M, d.N, d.K, d.taps,
            tuple(d.shift), d.relu, d.plane_h, d.plane_w, d.upsample_add, d.split_rows)


def chain_eligible(d):
    return d.a_dtype == BF16 and d.d_dtype == BF16 and d.N % 64 == 0 and d.K % 64 == 0 and not d.upsample_add and not d.split_rows and d.relu != 2


def run_chain(rec, record=True):
    """rec: list of (GemmDesc, keepalive tensors).  One lvcb200_gemm_chain_run (plan cached), or per-layer launches."""
    lib = _lib.load()
    if len(rec) < 2 or not all(chain_eligible(d) for d, _ in rec):
        if record and GEMM_RECORD is not None:
            GEMM_RECORD.extend(rec)
        replay_gemms(rec)
        return
    if record and GEMM_RECORD is not None:
        GEMM_RECORD.append((rec, "chain"))
    key = tuple(_desc_key(d) for d, _ in rec)
    ent = gemm_chain._plans.get(key)
    if ent is None:
        import torch
        if torch.cuda.is_current_stream_capturing():
            raise _lib.LvcB200Error("gemm_chain: plan must be built before CUDA-graph capture (run the step once eagerly)")
        n = len(rec)
        arr = (GemmDesc * n)(*[d for d, _ in rec])
        nbytes = lib.lvcb200_gemm_chain_workspace(arr, n)
        if nbytes == 0:
            _lib.check(-1, "lvcb200_gemm_chain_workspace")
        dev = rec[0][1][0].device
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        plan = ChainPlan()
        torch.cuda.current_stream().synchronize()
        _lib.check(lib.lvcb200_gemm_chain_plan(arr, n, _lib.ptr(ws), ws.numel(), ctypes.byref(plan)), "lvcb200_gemm_chain_plan")
        ent = (plan, ws, [t for _, t in rec])       # keep the workspace and every operand alive with the plan
        gemm_chain._plans[key] = ent
        if PLAN_LOG is not None:
            PLAN_LOG.append(key)
    _lib.check(lib.lvcb200_gemm_chain_run(ctypes.byref(ent[0]), _lib.stream_ptr()), "lvcb200_gemm_chain_run")


def replay_gemms(descs):
    """Re-issue recorded dense launches (same buffers) on the current stream; used to time the dense layers back to back.
    Entries are (GemmDesc, tensors) for a single launch or (list of those, "chain") for a layer-chain launch."""
    lib = _lib.load()
    sp = _lib.stream_ptr()
    for d, tag in descs:
        if isinstance(d, list):
            run_chain(d, record=False)
        else:
            _lib.check(lib.lvcb200_gemm_bf16(ctypes.byref(d), sp), "lvcb200_gemm_bf16")


def flatten_recorded(descs):
    """All per-layer descriptors of a recorded launch list (chains expanded)."""
    out = []
    for d, tag in descs:
        out += [x for x, _ in d] if isinstance(d, list) else [d]
    return out


def drop_plans(keys):
    """Forget layer-chain plans (and the operand references they hold) whose buffers are being released."""
    for k in keys:
        gemm_chain._plans.pop(k, None)


def _workspace(tag, nbytes, device):
    """Growable scratch per (op, device).  A buffer that is outgrown is RETIRED, not freed: kernels captured into a CUDA
    graph hold its raw pointer, and returning it to the caching allocator would let a later replay scribble over whatever
    tensor reuses the block."""
    key = (tag, device.index)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def _dt(t):
    if t.dtype == torch.float16:
        return _lib.F16
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise _lib.LvcB200Error(f"unsupported dtype {t.dtype}")


# ---------------------------------------------------------------------------------------------- pooler
def assign_boxes_to_levels(boxes, min_level=2, max_level=5, canonical_box_size=224, canonical_level=4):
    """detectron2/modeling/poolers.py:23-59 on device.  boxes [R,4] fp32 -> int64 [R]."""
    _lib.require_cuda(boxes)
    b = boxes.detach().to(torch.float32).contiguous()
    out = torch.empty(b.shape[0], dtype=torch.int64, device=b.device)
    if b.shape[0]:
        rc = _lib.load().lvcb200_assign_boxes_to_levels(_lib.ptr(b), b.shape[0], min_level, max_level, canonical_box_size,
                                                        canonical_level, _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_assign_boxes_to_levels")
    return out


class Plane:
    """A channels-last feature plane [n, H+2b, W+2b, C] (b = border, 0 or 1) viewed through lvcb200_fmap."""

    def __init__(self, tensor, H, W, C, border=1, c_stride=None):
        self.t, self.H, self.W, self.C, self.border = tensor, H, W, C, border
        self.c_stride = c_stride or C
        self.PH, self.PW = H + 2 * border, W + 2 * border

    @property
    def n(self):
        return self.t.shape[0]

    def fmap(self, scale):
        es = self.t.element_size()
        base = self.t.data_ptr() + (self.border * self.PW + self.border) * self.c_stride * es
        return FMap(base, self.H, self.W, self.PH * self.PW, self.PW, self.c_stride, scale)

    def valid(self):
        b = self.border
        return self.t[:, b:b + self.H, b:b + self.W, : self.C]

    @staticmethod
    def from_nchw(x, dtype=torch.bfloat16, border=1):
        n, c, h, w = x.shape
        t = torch.zeros((n, h + 2 * border, w + 2 * border, c), dtype=dtype, device=x.device)
        t[:, border:border + h, border:border + w] = x.permute(0, 2, 3, 1).to(dtype)
        return Plane(t, h, w, c, border)

    def to_nchw(self):
        return self.valid().permute(0, 3, 1, 2).float().contiguous()


class PairPlane(Plane):
    """Strict-mode plane: x = hi + lo, two bf16 planes in one [2 * split_rows, C] matrix (hi half at row 0, lo half at row
    split_rows, a multiple of 128 >= n * PH * PW) -- the operand format of the SPLIT GEMM (lvcb200_gemm_desc.split_rows)."""

    def __init__(self, full, n, H, W, C, border=1):
        PH, PW = H + 2 * border, W + 2 * border
        M, S = n * PH * PW, full.shape[0] // 2
        assert full.dim() == 2 and full.shape[1] == C and S >= M and S % 128 == 0
        super().__init__(full[:M].view(n, PH, PW, C), H, W, C, border)
        self.full, self.split_rows, self.M = full, S, M
        self.lo = full[S:S + M].view(n, PH, PW, C)

    @staticmethod
    def rows_for(n, H, W, border=1):
        return ((n * (H + 2 * border) * (W + 2 * border) + 127) // 128) * 128

    def to_nchw(self):
        b = self.border
        v = self.t[:, b:b + self.H, b:b + self.W].float() + self.lo[:, b:b + self.H, b:b + self.W].float()
        return v.permute(0, 3, 1, 2).contiguous()

    def merged(self, out=None):
        """fp32 plane [n, PH, PW, C] = hi + lo (lvcb200_pair_merge)."""
        if out is None:
            out = torch.empty(self.t.shape, dtype=torch.float32, device=self.full.device)
        _lib.check(_lib.load().lvcb200_pair_merge(_lib.ptr(self.full), self.split_rows, self.M, self.C, _lib.ptr(out), _lib.stream_ptr()),
                   "lvcb200_pair_merge")
        return out

    @staticmethod
    def from_nchw(x, border=1):
        n, c, h, w = x.shape
        S = PairPlane.rows_for(n, h, w, border)
        full = torch.zeros((2 * S, c), dtype=torch.bfloat16, device=x.device)
        p = PairPlane(full, n, h, w, c, border)
        v = x.permute(0, 2, 3, 1).float()
        hi = v.bfloat16()
        p.t[:, border:border + h, border:border + w] = hi
        p.lo[:, border:border + h, border:border + w] = (v - hi.float()).bfloat16()
        return p


def pair_split(x, out=None):
    """fp32 [rows, cols] -> bf16 pair matrix [2 * split_rows, cols] (lvcb200_pair_split); returns (full, split_rows)."""
    _lib.require_cuda(x)
    x = x.contiguous()
    rows, cols = x.shape
    S = (rows + 127) // 128 * 128
    if out is None:
        out = torch.zeros((2 * S, cols), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (2 * S, cols)
    _lib.check(_lib.load().lvcb200_pair_split(_lib.ptr(x), rows, cols, _lib.ptr(out), S, _lib.stream_ptr()), "lvcb200_pair_split")
    return out, S


def split_weight(w, taps=1):
    """fp32 [N, taps*K] -> bf16 [N, taps*2K] with [hi | lo] per tap (W operand of the SPLIT GEMM)."""
    N = w.shape[0]
    K = w.shape[1] // taps
    w = w.float()
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    return torch.stack([hi.view(N, taps, K), lo.view(N, taps, K)], dim=2).reshape(N, 2 * taps * K).contiguous()


def row_inv_norm(x, scale, eps=1e-5, lo_off=0, rows=None, out=None):
    """out[r] = scale / (||x_r|| + eps) (CosineSimOutputLayers.forward, fast_rcnn.py:826-829); x bf16 [R, C] (a pair when lo_off != 0) or fp32."""
    _lib.require_cuda(x)
    R = rows if rows is not None else x.shape[0]
    if out is None:
        out = torch.empty(R, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().lvcb200_row_inv_norm(_lib.ptr(x), _dt(x), lo_off, R, x.shape[1], x.stride(0), scale, eps, _lib.ptr(out),
                                                _lib.stream_ptr()), "lvcb200_row_inv_norm")
    return out


def make_rois(props, counts, rois=None, roi_image=None):
    """Padded proposals [n, P, 4] + counts [n] int32 -> pooler-format rois [n*P, 5] and roi_image [n*P] int32 (-1 = padding)."""
    _lib.require_cuda(props, counts)
    n, P = props.shape[0], props.shape[1]
    props = props.contiguous()
    if rois is None:
        rois = torch.empty((n * P, 5), dtype=torch.float32, device=props.device)
    if roi_image is None:
        roi_image = torch.empty(n * P, dtype=torch.int32, device=props.device)
    _lib.check(_lib.load().lvcb200_make_rois(_lib.ptr(props), _lib.ptr(counts), n, P, _lib.ptr(rois), _lib.ptr(roi_image), _lib.stream_ptr()),
               "lvcb200_make_rois")
    return rois, roi_image


def roi_pool_fpn(planes, scales, rois, pooled=7, sampling_ratio=0, out_dtype=torch.float32, out_layout=OUT_NCHW,
                 canonical_box_size=224, canonical_level=4, return_levels=False):
    """ROIPooler.forward (poolers.py:191-246) in one launch.  planes: list[Plane]; rois [R,5] fp32."""
    _lib.require_cuda(rois, *[p.t for p in planes])
    r = rois.detach().to(torch.float32).contiguous()
    R = r.shape[0]
    C = planes[0].C
    arr = (FMap * len(planes))(*[p.fmap(s) for p, s in zip(planes, scales)])
    out = torch.empty((R, C * pooled * pooled), dtype=out_dtype, device=r.device)
    lv = torch.empty(R, dtype=torch.int64, device=r.device) if return_levels else None
    min_level = int(round(-math.log2(scales[0])))
    if R:
        rc = _lib.load().lvcb200_roi_pool_fpn(arr, len(planes), _dt(planes[0].t), C, _lib.ptr(r), R, pooled, sampling_ratio,
                                              canonical_box_size, canonical_level, min_level, _lib.ptr(out), _dt(out), out_layout,
                                              out.shape[1], _lib.ptr(lv), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_roi_pool_fpn")
    if out_layout == OUT_NCHW:
        out = out.view(R, C, pooled, pooled)
    else:
        out = out.view(R, pooled, pooled, C)
    return (out, lv) if return_levels else out


# ---------------------------------------------------------------------------------------------- RPN
def cell_anchors(size, ratios):
    """generate_cell_anchors (anchor_generator.py:173-208) for one size; computed in double then cast to fp32."""
    out = []
    area = float(size) ** 2
    for r in ratios:
        w = math.sqrt(area / r)
        h = r * w
        out += [-w / 2.0, -h / 2.0, w / 2.0, h / 2.0]
    return out


def rpn_proposals(level_inputs, image_sizes, anchor_sizes, anchor_ratios, strides=(4, 8, 16, 32, 64), pre_nms_topk=1000,
                  post_nms_topk=1000, nms_thresh=0.7, min_box_size=0.0, weights=(1.0, 1.0, 1.0, 1.0), nms_mode=-1):
    """RPN.predict_proposals on device.

    level_inputs: list of dicts(logits=fp32 tensor, deltas=fp32 tensor, H, W, A, strides_l=(img,row,pix), strides_d=(img,row,pix),
    offset_l / offset_d = element offset of (n=0,y=0,x=0,a=0)); helper `rpn_level_dense` builds one for [N,HWA] / [N,HWA,4].
    image_sizes: int32 [N,2] device tensor (h, w).  Returns proposals [N,post,4], logits [N,post], counts [N] int32."""
    lib = _lib.load()
    N = image_sizes.shape[0]
    L = len(level_inputs)
    levels = (RpnLevel * L)()
    keep_alive = []
    for i, li in enumerate(level_inputs):
        lg, dl = li["logits"], li["deltas"]
        _lib.require_cuda(lg, dl)
        assert lg.dtype == torch.float32 and dl.dtype == torch.float32
        keep_alive += [lg, dl]
        lv = levels[i]
        lv.logits = lg.data_ptr() + 4 * li.get("offset_l", 0)
        lv.deltas = dl.data_ptr() + 4 * li.get("offset_d", 0)
        lv.H, lv.W, lv.A, lv.stride = li["H"], li["W"], li["A"], strides[i]
        lv.img_stride_l, lv.row_stride_l, lv.pix_stride_l = li["strides_l"]
        lv.img_stride_d, lv.row_stride_d, lv.pix_stride_d = li["strides_d"]
        ca = cell_anchors(anchor_sizes[i], anchor_ratios)
        for j, v in enumerate(ca):
            lv.cell_anchors[j] = v
    p = RpnParams(N, L, pre_nms_topk, post_nms_topk, nms_thresh, min_box_size, (ctypes.c_float * 4)(*weights), nms_mode)
    dev = image_sizes.device
    ws = _workspace("rpn", lib.lvcb200_rpn_proposals_workspace(ctypes.byref(p)), dev)
    props = torch.empty((N, post_nms_topk, 4), dtype=torch.float32, device=dev)
    logits = torch.empty((N, post_nms_topk), dtype=torch.float32, device=dev)
    counts = torch.empty(N, dtype=torch.int32, device=dev)
    rc = lib.lvcb200_rpn_proposals(levels, ctypes.byref(p), _lib.ptr(image_sizes), _lib.ptr(props), _lib.ptr(logits), _lib.ptr(counts),
                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_rpn_proposals")
    return props, logits, counts


def rpn_level_dense(logits, deltas, H, W, A):
    """Reference layout: logits [N, H*W*A], deltas [N, H*W*A, 4] contiguous fp32."""
    return dict(logits=logits.contiguous(), deltas=deltas.contiguous(), H=H, W=W, A=A,
                strides_l=(H * W * A, 0, A), strides_d=(H * W * A * 4, 0, A * 4))


# ---------------------------------------------------------------------------------------------- detections
def detections(cls_logits, box_deltas, proposals, roi_image, image_sizes, out_sizes, num_classes, max_rois_per_image=1000,
               weights=(10.0, 10.0, 5.0, 5.0), score_thresh=0.05, nms_thresh=0.5, topk=100, nms_mode=-1, row_scale=None,
               class_agnostic=False):
    """softmax + decode + threshold + per-class NMS + top-k + detector_postprocess (fast_rcnn.py:95-137,440-493;
    postprocessing.py:10-79) on device.  Returns boxes [N,topk,4], scores, classes (int64), rows (int64), counts (int32)."""
    lib = _lib.load()
    _lib.require_cuda(cls_logits, box_deltas, proposals, roi_image, image_sizes, out_sizes)
    assert cls_logits.dtype == torch.float32 and box_deltas.dtype == torch.float32 and proposals.dtype == torch.float32
    assert cls_logits.stride(1) == 1 and box_deltas.stride(1) == 1
    proposals = proposals.contiguous()
    N = image_sizes.shape[0]
    R = cls_logits.shape[0]
    dev = cls_logits.device
    p = DetParams(N, num_classes, max_rois_per_image, int(class_agnostic), (ctypes.c_float * 4)(*weights), score_thresh, nms_thresh,
                  topk, nms_mode)
    ws = _workspace("det", lib.lvcb200_detections_workspace(ctypes.byref(p)), dev)
    boxes = torch.empty((N, topk, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((N, topk), dtype=torch.float32, device=dev)
    classes = torch.empty((N, topk), dtype=torch.int64, device=dev)
    rows = torch.empty((N, topk), dtype=torch.int64, device=dev)
    counts = torch.empty(N, dtype=torch.int32, device=dev)
    rc = lib.lvcb200_detections(_lib.ptr(cls_logits), cls_logits.stride(0), _lib.ptr(row_scale), _lib.ptr(box_deltas),
                                box_deltas.stride(0), _lib.ptr(proposals), _lib.ptr(roi_image), R, ctypes.byref(p),
                                _lib.ptr(image_sizes), _lib.ptr(out_sizes), _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(classes),
                                _lib.ptr(rows), _lib.ptr(counts), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_detections")
    return boxes, scores, classes, rows, counts


def apply_deltas_clip(boxes, deltas, weights, roi_image=None, image_sizes=None):
    """Class-agnostic Box2BoxTransform.apply_deltas (+ Boxes.clip when roi_image / image_sizes are given)."""
    _lib.require_cuda(boxes, deltas, roi_image, image_sizes)
    assert boxes.dtype == torch.float32 and deltas.dtype == torch.float32 and deltas.stride(1) == 1
    boxes = boxes.contiguous()
    out = torch.empty_like(boxes)
    w = (ctypes.c_float * 4)(*weights)
    rc = _lib.load().lvcb200_apply_deltas_clip(_lib.ptr(boxes), _lib.ptr(deltas), deltas.stride(0), _lib.ptr(roi_image), _lib.ptr(image_sizes),
                                               boxes.shape[0], w, int(roi_image is not None), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_apply_deltas_clip")
    return out


# ---------------------------------------------------------------------------------------------- kNN
class KnnBank:
    """Support bank after the all-gather: centred + normalised once (lvcb200_knn_prepare)."""

    def __init__(self, bank, bank_cls, cosine=True):
        """cosine=False: QUERY_EXPAND.COSINE_SIM False, neighbours by -cdist (run_nearest_neighbours.py:154-159)."""
        _lib.require_cuda(bank, bank_cls)
        lib = _lib.load()
        self.bank = bank.detach().to(torch.float32).contiguous()
        self.cls = bank_cls.detach().to(torch.int64).contiguous()
        self.S, self.D = self.bank.shape
        self.cosine = cosine
        self.prepared = torch.empty(lib.lvcb200_knn_prepared_bytes(self.S, self.D), dtype=torch.uint8, device=bank.device)
        prep = lib.lvcb200_knn_prepare if cosine else lib.lvcb200_knn_prepare_euclid
        rc = prep(_lib.ptr(self.bank), self.S, self.D, _lib.ptr(self.prepared), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_knn_prepare")

    def tc_eligible(self):
        return self.cosine and 64 <= self.S <= 4096 and self.D % 8 == 0 and 32 <= self.D <= 4096

    def verify(self, queries, query_cls, topk=10, knn=10, return_sim=False, path="auto"):
        """path: "auto" (tensor-core path when the bank shape allows it), "tc" / "tc3" (bf16-pair scores, top-k in the GEMM epilogue,
        exact re-scoring of uncertain queries), "tc1" (round-1 path: TF32 scores + exact re-rank), or "simt" (exact fp32 FMA)."""
        _lib.require_cuda(queries, query_cls)
        q = queries.detach().to(torch.float32).contiguous()
        qc = query_cls.detach().to(torch.int64).contiguous()
        Q = q.shape[0]
        dev = q.device
        top_idx = torch.empty((Q, topk), dtype=torch.int64, device=dev)
        votes = torch.empty((Q, topk), dtype=torch.int64, device=dev)
        keep = torch.empty(Q, dtype=torch.uint8, device=dev)
        sim = torch.empty((Q, topk), dtype=torch.float32, device=dev) if return_sim else None
        lib = _lib.load()
        use_tc = path in ("tc", "tc1", "tc3") or (path == "auto" and self.tc_eligible() and Q > 0)
        if use_tc:
            if path in ("tc1", "tc3"):     # explicit version (A/B, tests): 1 = TF32 scores + exact re-rank, 3 = bf16 pairs + epilogue top-k
                _lib.check(lib.lvcb200_knn_tc_select(int(path[2])), "lvcb200_knn_tc_select")
            elif path == "tc":
                _lib.check(lib.lvcb200_knn_tc_select(3), "lvcb200_knn_tc_select")
            ws = _workspace("knn", lib.lvcb200_knn_tc_workspace(Q, self.S, self.D), dev)
            rc = lib.lvcb200_knn_verify_tc(_lib.ptr(self.prepared), _lib.ptr(self.cls), self.S, self.D, _lib.ptr(q), _lib.ptr(qc), Q,
                                           topk, knn, _lib.ptr(top_idx), _lib.ptr(sim), _lib.ptr(votes), _lib.ptr(keep), _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "lvcb200_knn_verify_tc")
        else:
            fn = lib.lvcb200_knn_verify if self.cosine else lib.lvcb200_knn_verify_euclid
            rc = fn(_lib.ptr(self.prepared), _lib.ptr(self.cls), self.S, self.D, _lib.ptr(q), _lib.ptr(qc), Q,
                    topk, knn, _lib.ptr(top_idx), _lib.ptr(sim), _lib.ptr(votes), _lib.ptr(keep), _lib.stream_ptr())
            _lib.check(rc, "lvcb200_knn_verify")
        return dict(top_idx=top_idx, votes=votes, keep=keep, top_sim=sim)


# ---------------------------------------------------------------------------------------------- dense
def gemm(A, W, bias=None, residual=None, out=None, out_dtype=torch.bfloat16, relu=False, taps=1, shifts=(0,), K=None,
         M=None, plane_hw=None, upsample_add=None, split_rows=0):
    """D = act(sum_t A[m+shift_t, :K] @ W[:, t*K:(t+1)*K]^T + bias + residual).  A [rows, >=K] bf16 (row pitch = stride(0)),
    W [N, taps*K] bf16.  plane_hw=(PH, PW) zeroes border rows of a zero-bordered plane.  upsample_add: a coarser Plane (half the
    interior size) whose nearest-2x upsampling is added in the epilogue (FPN top-down path, fpn.py:128-134).
    split_rows != 0 (strict mode): A, residual and a bf16 `out` are hi/lo pair matrices [2 * split_rows, .] (M = rows of one half),
    W is `split_weight(w, taps)`; three-term bf16 product with fp32-grade accuracy (lvcb200_gemm_desc.split_rows)."""
    _lib.require_cuda(A, W, bias, residual, out)
    assert A.dtype == W.dtype and A.dtype in (torch.bfloat16, torch.float32) and A.stride(1) == 1 and W.stride(1) == 1
    N = W.shape[0]
    K = K or (W.shape[1] // (taps * (2 if split_rows else 1)))
    if split_rows:
        assert M is not None and split_rows >= M and A.shape[0] >= split_rows + M
    M = M if M is not None else A.shape[0]
    if out is None:
        if split_rows and out_dtype != torch.float32:
            out = torch.zeros((2 * split_rows, N), dtype=out_dtype, device=A.device)
        else:
            out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    assert out.stride(1) == 1
    d = GemmDesc()
    d.a_dtype = _dt(A)
    d.A, d.lda, d.M_rows = A.data_ptr(), A.stride(0), A.shape[0]
    d.W, d.ldw = W.data_ptr(), W.stride(0)
    d.bias = bias.data_ptr() if bias is not None else None
    d.residual, d.ldr = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
    d.D, d.ldd, d.d_dtype = out.data_ptr(), out.stride(0), _dt(out)
    d.M, d.N, d.K, d.taps = M, N, K, taps
    for i, s in enumerate(shifts):
        d.shift[i] = int(s)
    d.relu = 2 if relu == "gelu" else int(relu)
    d.split_rows = int(split_rows)
    d.plane_h, d.plane_w = plane_hw if plane_hw else (0, 0)
    if upsample_add is not None:
        _lib.require_cuda(upsample_add.t)
        d.upsample_add, d.ldu = upsample_add.t.data_ptr(), upsample_add.c_stride
        d.up_plane_h, d.up_plane_w = upsample_add.PH, upsample_add.PW
    if GEMM_CHAIN is not None:
        assert upsample_add is None, "upsample_add layers are not chainable"
        GEMM_CHAIN.append((d, (A, W, bias, residual, out)))
        return out
    if GEMM_RECORD is not None:
        GEMM_RECORD.append((d, (A, W, bias, residual, out, upsample_add)))   # keep the tensors alive with the descriptor
    if GEMM_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.load().lvcb200_gemm_bf16(ctypes.byref(d), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_gemm_bf16")
    if GEMM_EVENTS is not None:
        e1.record()
        GEMM_EVENTS.append((e0, e1, (M, N, K * taps)))
    return out
