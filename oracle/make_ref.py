"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference under oracle/_ref/ (git-ignored).

The reference (danielcodelavin/vivid) is pure Python: there is nothing to compile, so "building"
the reference checker means placing the handful of modules the hot path imports next to the oracle, byte for byte,
where they travel to the GPU box like a built .so does (oracle/_ref/ is git-ignored, not gpurun-ignored).  Nothing
under oracle/_ref/ is ever committed, imported by the product package, or edited.

    python oracle/make_ref.py            # needs /root/reference (the build container); no-op on the GPU box

Layout:  oracle/_ref/current/   {training/, torch_utils/, dnnlib/, generate_images.py, calculate_metrics.py}
         oracle/_ref/snapshot/  {training/, dnnlib/, generate_images.py}      (experiments/code; torch_utils shared)
         oracle/_ref/MANIFEST.json   file -> sha256 of the source it was staged from
Used by: oracle/ref_loader.py (bench.py --impl reference / cpu_baseline with kind "reference"; tests).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")

CURRENT = ["training/__init__.py", "training/models.py", "training/utils.py", "training/encoders.py",
           "training/custom_litdata_loader.py", "torch_utils", "dnnlib", "generate_images.py", "calculate_metrics.py"]
SNAPSHOT = ["training/__init__.py", "training/models.py", "training/utils.py", "training/encoders.py",
            "training/custom_litdata_loader.py", "dnnlib", "generate_images.py"]


def _stage(src_root, rel, dst_root, manifest):
    src = os.path.join(src_root, rel)
    if os.path.isdir(src):
        for name in sorted(os.listdir(src)):
            if name.endswith(".py"):
                _stage(src_root, os.path.join(rel, name), dst_root, manifest)
        return
    dst = os.path.join(dst_root, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    with open(src, "rb") as f:
        manifest[os.path.relpath(dst, DST)] = hashlib.sha256(f.read()).hexdigest()


def make(verbose=True):
    """Stage the reference modules; returns the destination or None when /root/reference is absent."""
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/make_ref: /root/reference not present; keeping whatever oracle/_ref already holds")
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in CURRENT:
        _stage(REF, rel, os.path.join(DST, "current"), manifest)
    for rel in SNAPSHOT:
        _stage(os.path.join(REF, "experiments", "code"), rel, os.path.join(DST, "snapshot"), manifest)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/make_ref: staged {len(manifest)} reference modules under {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
