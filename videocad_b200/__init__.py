"""videocad_b200 -- B200-native (sm_100a) hot path of ghadinehme/VideoCAD's behaviour-cloning model.

Public surface (mirrors the reference's model package for this path):
    AutoRegressiveTransformer   drop-in nn.Module            (reference: model/autoregressive_transformer.py)
    ModelFactory, ModelType     drop-in factory               (reference: model/model_factory.py)
    build_library()             compile libvideocad_b200.so   (nvcc, -gencode arch=compute_100a,code=sm_100a)
"""
from .model import AutoRegressiveTransformer, ViTParams  # noqa: F401
from .model_factory import ModelFactory, ModelType  # noqa: F401


def build_library(force: bool = False, verbose: bool = False) -> str:
    from .build import build as _build

    return _build(force=force, verbose=verbose)
