"""
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product package.

Stages the UNMODIFIED reference (read-only at /root/reference, build container only) under the
git-ignored ``baseline/_ref/`` so that it travels to the GPU box with the snapshot (``baseline/_ref`` is
listed in .gitignore, not in .gpurunignore; no reference source enters the history):

    baseline/_ref/src, eval, conf, expconf.conf   byte-identical copies (checked by sha256 below)
    baseline/_ref/MANIFEST.json                   file -> sha256 of what was staged

Users: ``bench.py --impl reference`` / ``cpu_baseline`` (times the reference's own renderer on the host
cores, ``cpu_baseline.kind = "reference"``) and tests/test_gpu_reference_callers.py (runs the reference's
``eval/gen_video.py`` itself against ``dropin/src``).  Called by ``__graft_entry__.build()``.
"""
import hashlib
import json
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PIXELNERF_REFERENCE", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
PARTS = ("src", "eval", "conf", "expconf.conf")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged_root():
    """Path of the staged reference tree, or None."""
    ok = os.path.isfile(os.path.join(DST, "src", "model", "models.py.backup2"))
    return DST if ok else None


def stage(verbose=True):
    """Copies the reference tree when it is present; returns the staged root or None."""
    if not os.path.isdir(os.path.join(SRC, "src")):
        return staged_root()
    manifest = {}
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            for root, _dirs, files in os.walk(d):
                for fn in files:
                    p = os.path.join(root, fn)
                    manifest[os.path.relpath(p, DST)] = _sha(p)
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[part] = _sha(d)
    for rel, digest in manifest.items():  # byte-identical to the source tree
        assert _sha(os.path.join(SRC, rel)) == digest, rel
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return DST


if __name__ == "__main__":
    stage()
