"""Stage the UNMODIFIED reference sources the oracle needs into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

    python -m oracle.build_ref            # in the build container, where /root/reference exists

The reference is a pure-Python tree without setup.py / pyproject.toml: there is nothing to compile or pip-install.  What the
checker and the timed CPU arm need is the reference's own code for the hot path and its callers:

    model/{model_factory,autoregressive_transformer,base_transformer,trajectory_model}.py     the path itself
    trainer.py, class_weights.json                                                            the caller that drives it
    data_loader/{data_loader,image_loader,sequence_retriver}.py, utils.py                     the data formats either side (8(f))

They are copied byte for byte into oracle/_ref/ -- git-ignored (never part of the history), not gpurun-ignored (travels to
the GPU box exactly like a built .so), so that `bench.py --impl reference`, the `cpu_baseline` leg and the tests can run the
REAL reference there ("kind": "reference").  Nothing under oracle/_ref/ is imported by the product package.  A manifest with
the sha256 of every staged file is written next to them; `verify()` re-checks it before use.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("VIDEOCAD_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = [
    "model/model_factory.py", "model/autoregressive_transformer.py", "model/base_transformer.py", "model/trajectory_model.py",
    "trainer.py", "class_weights.json",
    "data_loader/data_loader.py", "data_loader/image_loader.py", "data_loader/sequence_retriver.py", "utils.py",
]
MANIFEST = "MANIFEST.json"


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def source_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model", "autoregressive_transformer.py"))


def staged() -> bool:
    return os.path.isfile(os.path.join(REF_DST, MANIFEST)) and os.path.isfile(os.path.join(REF_DST, "trainer.py"))


def build(force: bool = False) -> str:
    """Copy the files (only when the source tree is present); returns the staging directory."""
    if not source_available():
        if staged():
            return REF_DST
        raise RuntimeError(f"reference sources not found under {REF_SRC} and nothing staged under {REF_DST}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        digest = _sha(src)
        if force or not os.path.exists(dst) or _sha(dst) != digest:
            if os.path.exists(dst):
                os.chmod(dst, 0o644)
            shutil.copyfile(src, dst)
        manifest[rel] = digest
    with open(os.path.join(REF_DST, MANIFEST), "w") as f:
        json.dump(dict(source=REF_SRC, files=manifest), f, indent=1, sort_keys=True)
    return REF_DST


def verify() -> bool:
    """True when every staged file still has the digest recorded at staging time (i.e. is the unmodified reference)."""
    if not staged():
        return False
    with open(os.path.join(REF_DST, MANIFEST)) as f:
        manifest = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(REF_DST, rel)) and _sha(os.path.join(REF_DST, rel)) == d for rel, d in manifest.items())


if __name__ == "__main__":
    print(build(), "verified" if verify() else "NOT verified")
