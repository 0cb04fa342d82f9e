"""TEST / BASELINE INFRASTRUCTURE -- recipe that makes the reference runnable where /root/reference does not exist (the GPU box).

The reference (wcq99681-svg/YOLO-AD-Refine) is pure Python on this path, so "building" it is copying the `ultralytics` package (sources and
cfg yamls only) and the 701 model yaml from /root/reference, where they lie, into oracle/_ref/ -- an OUTPUT directory that is git-ignored (so the
reference's sources never enter this repository's history) but not gpurun-ignored (so it travels to the GPU box like the built libyad.so).
oracle/ref_shims.py puts it on sys.path when /root/reference is absent.

Users: `bench.py --impl reference` (the reference's own CPU implementation as the baseline arm, cpu_baseline.kind = "reference") and the plugin
test that drives the reference's own DetectionModel._predict_once / ops.non_max_suppression with libyad-backed modules.  Product code never
imports it.       python -m oracle.build_ref
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("YAD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yaml", ".yml", ".json", ".txt")


def build(verbose=True):
    """copy the runnable part of the reference into oracle/_ref (idempotent); returns the path, or None when the source tree is absent"""
    pkg = os.path.join(SRC, "ultralytics")
    if not os.path.isdir(pkg):
        return DST if os.path.isdir(os.path.join(DST, "ultralytics")) else None
    n = 0
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in files:
            if f.endswith(KEEP_EXT):
                s, d = os.path.join(root, f), os.path.join(DST, rel, f)
                if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                    shutil.copyfile(s, d)
                n += 1
    os.makedirs(os.path.join(DST, "z-yaml"), exist_ok=True)
    for f in os.listdir(os.path.join(SRC, "z-yaml")):
        if f.endswith((".yaml", ".yml")):
            shutil.copyfile(os.path.join(SRC, "z-yaml", f), os.path.join(DST, "z-yaml", f))
    if verbose:
        print(f"oracle/_ref: {n} files of the reference's ultralytics package + z-yaml/*.yaml")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
