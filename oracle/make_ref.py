"""Recipe that stages the UNMODIFIED reference package under ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

    python oracle/make_ref.py            # /root/reference/rajni/**/*.py -> oracle/_ref/rajni/

The reference (dRaniwal/RAJNI-ViT) is pure Python with no packaging metadata, so there is nothing to
compile or pip-install: "building" it means copying its ``rajni`` package byte for byte.  ``oracle/_ref/``
is listed in ``.gitignore`` (reference sources never enter this repo's history) but not in
``.gpurunignore``, so the copy travels to the GPU box, where /root/reference does not exist.
``__graft_entry__.build()`` runs this whenever /root/reference is present.

Who may use ``oracle/_ref``: ``bench.py --impl reference`` and bench.py's ``cpu_baseline`` /
``gpu_eager_reference`` legs (the reference's own ``evaluate_model`` + ``RAJNIViTWrapper`` on the stand-in
ViT, ``kind: "reference"``), and ``tests/``.  Nothing in ``rajni_vit_b200/`` imports it.

``MANIFEST.json`` records the sha256 of every staged file next to the sha256 of its source, so a stale or
edited copy is detectable (``verify()``).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/rajni"
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "rajni")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    """True when a staged copy exists (on the GPU box: the one that travelled with the snapshot)."""
    return os.path.isfile(os.path.join(DST, "__init__.py")) and os.path.isfile(os.path.join(DST_ROOT, "MANIFEST.json"))


def verify() -> bool:
    """Every staged file still has the hash recorded when it was copied."""
    try:
        man = json.load(open(os.path.join(DST_ROOT, "MANIFEST.json")))
        return all(_sha(os.path.join(DST_ROOT, rel)) == h for rel, h in man["files"].items())
    except Exception:
        return False


def make(src: str = SRC) -> str:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found (the reference only exists in the build container)")
    if os.path.isdir(DST_ROOT):
        shutil.rmtree(DST_ROOT)
    files = {}
    for root, dirs, names in os.walk(src):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for name in sorted(names):
            if not name.endswith(".py"):
                continue
            s = os.path.join(root, name)
            rel = os.path.join("rajni", os.path.relpath(s, src))
            d = os.path.join(DST_ROOT, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            files[rel] = _sha(d)
            assert files[rel] == _sha(s)
    json.dump({"source": src, "files": files}, open(os.path.join(DST_ROOT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return DST_ROOT


def import_reference():
    """Import the staged package as ``rajni`` and return the module (raises if it is missing or was edited)."""
    if not available() or not verify():
        raise RuntimeError("oracle/_ref is missing or does not match its manifest: run `python oracle/make_ref.py` in the build container")
    if DST_ROOT not in sys.path:
        sys.path.insert(0, DST_ROOT)
    import rajni                                        # noqa: E402  (the reference's own package name)
    if not os.path.abspath(rajni.__file__).startswith(DST_ROOT):
        raise RuntimeError(f"`rajni` resolved to {rajni.__file__}, not to oracle/_ref")
    return rajni


if __name__ == "__main__":
    print("staged", make())
