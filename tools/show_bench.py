"""Print the key numbers of a bench.py JSON line."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)   # `| head` closes the pipe early

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ({d['ms_per_step']:.3f} ms/step, n_gpus {d['n_gpus']});  e2e {d['e2e']['value']:.1f} ({d['e2e']['ms_per_step']:.3f} ms)")
r = d.get("roofline") or {}
print("roofline:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("achieved", "frac", "gemm_ms_per_step", "gemm_share_of_step", "traffic")})
if "knn" in d:
    print("knn tc:", d["knn"]["tensor_core_path"], " simt:", d["knn"]["simt_exact_path"])
    if "cpu_baseline" in d["knn"]:
        print("knn cpu:", d["knn"]["cpu_baseline"])
if "box_corrector" in d:
    print("corrector:", d["box_corrector"])
if "ops" in d:
    for k, v in d["ops"].items():
        print("op", k, v)
if "cpu_baseline" in d:
    print("cpu:", d["cpu_baseline"])
print("clocks:", d.get("clocks"), " launches:", d.get("gpu_launches"))
