"""Import the UNMODIFIED reference model from /root/reference (TEST INFRASTRUCTURE ONLY).

Works only where /root/reference exists (the build container).  `vit_pytorch` and `timm` are not
installed in this image, so the restatement/stub under oracle/shims/ is put on sys.path first
(SURVEY.md 8(c)).  Nothing is copied from, or written to, /root/reference.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("VIDEOCAD_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "autoregressive_transformer.py"))


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
