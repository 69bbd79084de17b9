"""Benchmark model families: the reference's published numbers are VGG16 (README.md:52-84);
BASELINE.json adds ResNet-50."""
from .resnet import ResNet, resnet18, resnet34, resnet50, resnet101, resnet152  # noqa: F401
from .vgg import VGG, vgg11, vgg13, vgg16, vgg19  # noqa: F401

_REGISTRY = {"vgg11": vgg11, "vgg13": vgg13, "vgg16": vgg16, "vgg19": vgg19, "resnet18": resnet18, "resnet34": resnet34,
             "resnet50": resnet50, "resnet101": resnet101, "resnet152": resnet152}


def build_model(name: str, **kw):
    if name not in _REGISTRY:
        raise KeyError(f"unknown model {name!r}; have {sorted(_REGISTRY)}")
    return _REGISTRY[name](**kw)
