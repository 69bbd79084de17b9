"""Model zoo of the benchmarks (bagua_net_b200/models): architecture parity by parameter count with the canonical
definitions (the reference's benchmark is torchvision's VGG16: reference README.md:52-84), shape checks on the CPU,
and state_dict compatibility between the eager and the fused layouts where it is promised."""
import pytest
import torch

from bagua_net_b200.models import build_model

# million parameters of the canonical ImageNet models (torchvision)
CANONICAL = {"vgg11": 132.86, "vgg13": 133.05, "vgg16": 138.36, "vgg19": 143.67, "resnet18": 11.69, "resnet34": 21.80,
             "resnet50": 25.56, "resnet101": 44.55, "resnet152": 60.19}


@pytest.mark.parametrize("name", sorted(CANONICAL))
def test_parameter_counts_match_the_canonical_architectures(name):
    m = build_model(name)
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - CANONICAL[name]) < 0.02


@pytest.mark.parametrize("name", ["vgg11", "vgg16", "resnet18", "resnet34", "resnet50"])
def test_forward_shapes_on_cpu(name):
    kw = dict(num_classes=7)
    if name.startswith("vgg"):
        kw.update(width_div=8, fc_dim=64, image_size=64)
    m = build_model(name, **kw).eval()
    with torch.no_grad():
        assert m(torch.randn(2, 3, 64, 64)).shape == (2, 7)


def test_unknown_model_is_a_clear_error():
    with pytest.raises(KeyError, match="unknown model"):
        build_model("alexnet")


def test_resnet_fused_layout_keeps_the_state_dict_keys():
    a, b = build_model("resnet18", num_classes=5), build_model("resnet18", num_classes=5, fused=True)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
