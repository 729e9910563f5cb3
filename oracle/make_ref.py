#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- stages the UNMODIFIED reference for use as the checker / the timed baseline.

The reference (jtang25/Beta-VAE) is pure Python; "building" it is staging its own source files,
byte for byte, from where they lie under /root/reference into the untracked directory
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the tree the same
way the built ``.so`` does; /root/reference itself does not exist there).  Nothing under
``oracle/_ref`` is ever imported by the product (``beta-vae_b200/``); its users are

  * ``bench.py --impl reference`` / ``cpu_baseline`` / ``gpu_eager_baseline``: time the reference's own
    ``models.beta_vae.BetaVAE`` class (``kind: "reference"``) instead of the oracle port;
  * ``tests/test_gpu_integration.py``: run the reference's own ``src/training/train.py`` on the drop-in model,
    exactly as INTEGRATION.md section 3 describes;
  * ``tests/test_oracle_golden.py``: (CPU, when present) cross-check the oracle port against the live class.

    python oracle/make_ref.py            # stage (idempotent); prints the manifest
    python oracle/make_ref.py --check    # exit 1 when the staged copy differs from /root/reference

What is staged: ``src/**/*.py`` (2.7 kLoC), ``scripts/generate_demo_data.py`` (the synthetic-slice generator the
debug config trains on) and ``configs/*.yaml``.  A manifest with the sha256 of every file is written to
``oracle/_ref/MANIFEST.json`` so a stale or edited copy is detectable.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BVAE_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def _wanted(root):
    out = []
    for base, dirs, files in os.walk(os.path.join(root, "src")):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for f in files:
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(base, f), root))
    out.append(os.path.join("scripts", "generate_demo_data.py"))
    cdir = os.path.join(root, "configs")
    if os.path.isdir(cdir):
        out += [os.path.join("configs", f) for f in sorted(os.listdir(cdir)) if f.endswith((".yaml", ".yml"))]
    return sorted(out)


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def available():
    """True when a staged copy is present (what bench.py / the tests ask)."""
    return os.path.exists(os.path.join(DST, "MANIFEST.json")) and \
        os.path.exists(os.path.join(DST, "src", "models", "beta_vae.py"))


def src_dir():
    return os.path.join(DST, "src")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print(f"make_ref: {REF} not present (GPU box?) -- keeping whatever is staged: available={available()}")
        return available()
    files = _wanted(REF)
    manifest = {}
    for rel in files:
        s, d = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and _sha(d) == _sha(s)):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"make_ref: staged {len(files)} files from {REF} into {DST}")
    return True


def check():
    if not available():
        print("make_ref: nothing staged")
        return 1
    with open(os.path.join(DST, "MANIFEST.json")) as f:
        man = json.load(f)["files"]
    bad = [rel for rel, h in man.items() if not os.path.exists(os.path.join(DST, rel)) or _sha(os.path.join(DST, rel)) != h]
    if os.path.isdir(os.path.join(REF, "src")):
        bad += [rel for rel in man if _sha(os.path.join(REF, rel)) != man[rel]]
    print("make_ref: OK" if not bad else f"make_ref: MISMATCH {sorted(set(bad))}")
    return 1 if bad else 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    sys.exit(check() if a.check else (0 if stage() else 1))
