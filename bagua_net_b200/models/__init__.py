"""Benchmark model families: the reference's published numbers are VGG16 (README.md:52-84);
BASELINE.json adds ResNet-50."""
from .resnet import ResNet, resnet18, resnet50, resnet101  # noqa: F401
from .vgg import VGG, vgg16, vgg19  # noqa: F401

_REGISTRY = {"vgg16": vgg16, "vgg19": vgg19, "resnet18": resnet18, "resnet50": resnet50, "resnet101": resnet101}


def build_model(name: str, **kw):
    if name not in _REGISTRY:
        raise KeyError(f"unknown model {name!r}; have {sorted(_REGISTRY)}")
    return _REGISTRY[name](**kw)
