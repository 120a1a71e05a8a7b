"""Golden vectors of the statistics leg of `calculate_metrics.py gen` from the UNMODIFIED reference
(calculate_stats_for_iterable_nvs :133-240, calculate_metrics_from_stats_nvs :289-322), CPU, one gloo rank.
The detector networks need downloads, so a deterministic fake (cases.FakeDetector) is placed in the reference's own
`_detector_cache`; everything after the detector call is the reference's code.  Build container only.
    python tests/golden/make_golden_metrics.py  -> tests/golden/metrics.pt"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

for m in ("kornia", "litdata"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path[:0] = ["/root/reference"]
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29611")
torch.distributed.init_process_group("gloo", rank=0, world_size=1)
import calculate_metrics as CM  # noqa: E402
import scipy.linalg  # noqa: E402

# scipy >= 1.16 (this image) dropped sqrtm's `disp` argument that the reference passes (:315); adapt the LIBRARY call
# signature so the reference's own expression runs unmodified
_sqrtm = scipy.linalg.sqrtm
scipy.linalg.sqrtm = lambda a, disp=True, **k: (_sqrtm(a, **k), None) if not disp else _sqrtm(a, **k)

det = cases.FakeDetector()
CM._detector_cache["fid"] = det
metrics = ["fid", "joint_fid", "psnr"]
batches = cases.synth_metric_batches()
it = [dict(src=s, tgt=t, images=i) for s, t, i in batches]
for r, ref in CM.calculate_stats_for_iterable_nvs(it, metrics=metrics, verbose=False, device=torch.device("cpu")):
    pass
results = CM.calculate_metrics_from_stats_nvs(stats=r.stats, ref=ref.stats, metrics=metrics, verbose=False)
out = dict(stats=r.stats, ref=ref.stats, results=results, metrics=metrics)
print({k: (v if not isinstance(v, dict) else {kk: getattr(vv, "shape", vv) for kk, vv in v.items()}) for k, v in r.stats.items()})
print(results)
torch.save(out, os.path.join(HERE, "metrics.pt"))
