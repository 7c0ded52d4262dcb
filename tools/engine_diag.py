"""GPU diagnostic: DetectorEngine vs the oracle (bf16-emulating and fp32) stage by stage.  Prints statistics only."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lvc_b200.config import DetectorConfig  # noqa: E402
from lvc_b200.modeling import DetectorEngine  # noqa: E402
from lvc_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import model as OM  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)), float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def iou_matrix(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None] - inter + 1e-12)


def main(depth=50, layer="FastRCNNOutputLayers", sizes=((320, 416), (300, 400))):
    cfg = DetectorConfig(depth=depth, output_layer=layer)
    sd = synthetic_state_dict(cfg, 0)
    ims = [torch.rand(3, h, w, generator=torch.Generator().manual_seed(100 + i)) * 255 for i, (h, w) in enumerate(sizes)]
    eng = DetectorEngine(cfg, sd)
    eng.debug = {}
    out = eng.run([im.cuda() for im in ims])
    torch.cuda.synchronize()
    dbg = eng.debug
    for mode in ("bf16-emulating oracle", "fp32 oracle"):
        col = {}
        t = time.time()
        ref = OM.detector_forward(cfg, sd, ims, device="cuda", collect=col, emulate_bf16=mode.startswith("bf16"))
        print(f"== engine vs {mode}  (oracle {time.time() - t:.1f}s)")
        for l in (2, 3, 4, 5):
            print(f" res{l} rel/maxrel", rel(dbg["feats"][l].to_nchw().cpu().numpy(), col["features_res"][f"res{l}"].numpy()) if "features_res" in col else "n/a")
        for l in (2, 3, 4, 5, 6):
            print(f" p{l} rel/maxrel", rel(dbg["pyramid"][l].to_nchw().cpu().numpy(), col["features"][f"p{l}"].numpy()))
        for n in range(len(ims)):
            c = int(dbg["prop_counts"][n])
            pb = dbg["props"][n, :c].cpu().numpy(); pl = dbg["prop_logits"][n, :c].cpu().numpy()
            rb, rl = col["proposals"][n]
            m = iou_matrix(rb, pb).max(1) if c else np.zeros(len(rb))
            same_order = float((np.abs(pl[: min(c, len(rl))] - rl[: min(c, len(rl))]) < 1e-3).mean())
            print(f" img{n}: proposals engine {c} oracle {len(rb)}; oracle props matched IoU>0.9: {(m > 0.9).mean():.3f}, >0.99: {(m > 0.99).mean():.3f}; logits agree(1e-3) {same_order:.3f}")
        R = dbg["pooled"].shape[0]
        print(" head rel", rel(dbg["head"].float().cpu().numpy()[: len(col["head"])], col["head"].numpy()) if len(col["head"]) == R else f"skip (R {R} vs {len(col['head'])})")
        boxes, scores, classes, rows, counts = out
        for n in range(len(ims)):
            c = int(counts[n])
            eb, es, ec = boxes[n, :c].cpu().numpy(), scores[n, :c].cpu().numpy(), classes[n, :c].cpu().numpy()
            r = ref[n]
            if len(r["scores"]) and c:
                m = iou_matrix(r["pred_boxes"], eb)
                j = m.argmax(1)
                ok = (m.max(1) > 0.9) & (ec[j] == r["pred_classes"]) & (np.abs(es[j] - r["scores"]) < 0.02)
                print(f" img{n}: dets engine {c} oracle {len(r['scores'])}; oracle dets matched (IoU>.9, class, |ds|<.02): {ok.mean():.3f}; top score {es[0]:.4f} vs {r['scores'][0]:.4f}")
            else:
                print(f" img{n}: dets engine {c} oracle {len(r['scores'])}")


if __name__ == "__main__":
    main()
    main(101, "CosineSimOutputLayers", ((256, 320),))
