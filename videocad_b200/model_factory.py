"""Drop-in for the reference's model/model_factory.py:1-36: same `create_model(model_name, model_config, device,
state_dict=None) -> (nn.Module, ModelType)` contract, same prefix stripping and strict=False load, but the module
returned is the B200-native AutoRegressiveTransformer."""
from enum import Enum

from .model import AutoRegressiveTransformer


class ModelType(Enum):
    MULTI_CLASSES = "multi_classes"


class ModelFactory:
    def create_model(self, model_name, model_config, device, state_dict=None):
        # model_name is ignored, exactly as in the reference (model_factory.py:15-23)
        model = AutoRegressiveTransformer(**model_config).to(device)
        model_type = ModelType.MULTI_CLASSES
        if state_dict:
            print("Loading state dict")
            new_state_dict = {}
            for k, v in state_dict.items():
                if k.startswith("module._orig_mod."):
                    new_state_dict[k.replace("module._orig_mod.", "")] = v
                elif k.startswith("module."):
                    new_state_dict[k.replace("module.", "")] = v
                else:
                    new_state_dict[k] = v
            model.load_state_dict(new_state_dict, strict=False)
        return model, model_type
