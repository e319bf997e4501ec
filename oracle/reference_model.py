"""Import the UNMODIFIED reference model from /root/reference (TEST INFRASTRUCTURE ONLY).

Source: /root/reference where it exists (the build container), else the byte-for-byte copies staged under oracle/_ref/ by
oracle/build_ref.py (git-ignored, shipped to the GPU box like a built .so).  `vit_pytorch` and `timm` are not installed in this
image, so the restatement/stub under oracle/shims/ is put on sys.path first.  Nothing is written to
/root/reference and no reference source enters the repository's history.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
_STAGED = os.path.join(_HERE, "_ref")  # byte-for-byte copies made by oracle/build_ref.py (git-ignored; travels with gpurun)


def _pick_root() -> str:
    env = os.environ.get("VIDEOCAD_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "model", "autoregressive_transformer.py")):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "autoregressive_transformer.py")) and \
        os.path.isfile(os.path.join(REFERENCE_ROOT, "trainer.py"))


def _prepare_path():
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)


def build_reference_model(cfg: dict, device="cpu"):
    """ModelFactory.create_model(...) of the real reference (model/model_factory.py:15-36)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _prepare_path()
    from model.model_factory import ModelFactory  # type: ignore

    cfg = dict(cfg)
    cfg.setdefault("state_dim", 1644)
    cfg.setdefault("act_dim", 7)
    model, model_type = ModelFactory().create_model(cfg.get("model_name", "autoregressive"), cfg, device)
    return model, model_type


def import_trainer():
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _prepare_path()
    import trainer  # type: ignore

    return trainer
