"""ResNet family (BASELINE.json config: "ResNet-50 bf16 torch DDP img/sec")."""
from __future__ import annotations

import torch
from torch import nn

from ..ops.fused_nn import conv_bn_act


class Bottleneck(nn.Module):
    expansion = 4

    fused = False     # set by ResNet(fused=True): conv + BatchNorm + ReLU (+ residual) through bagua_net_b200.ops.fused_nn

    def __init__(self, cin, width, stride=1, down=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(width * 4)
        self.relu = nn.ReLU(inplace=True)
        self.down = down

    def forward(self, x):
        if self.fused:
            idt = x if self.down is None else conv_bn_act(x, self.down[0], self.down[1], relu=False)
            out = conv_bn_act(x, self.conv1, self.bn1)
            out = conv_bn_act(out, self.conv2, self.bn2)
            return conv_bn_act(out, self.conv3, self.bn3, relu=True, res=idt)
        idt = x if self.down is None else self.down(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class BasicBlock(nn.Module):
    expansion = 1

    fused = False

    def __init__(self, cin, width, stride=1, down=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.down = down

    def forward(self, x):
        if self.fused:
            idt = x if self.down is None else conv_bn_act(x, self.down[0], self.down[1], relu=False)
            out = conv_bn_act(x, self.conv1, self.bn1)
            return conv_bn_act(out, self.conv2, self.bn2, relu=True, res=idt)
        idt = x if self.down is None else self.down(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes: int = 1000, base: int = 64, fused: bool = False):
        super().__init__()
        self.fused = fused
        self.cin = base
        self.conv1 = nn.Conv2d(3, base, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(base)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make(block, base, layers[0], 1)
        self.layer2 = self._make(block, base * 2, layers[1], 2)
        self.layer3 = self._make(block, base * 4, layers[2], 2)
        self.layer4 = self._make(block, base * 8, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(base * 8 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, (Bottleneck, BasicBlock)):
                m.fused = fused

    def _make(self, block, width, n, stride):
        down = None
        if stride != 1 or self.cin != width * block.expansion:
            down = nn.Sequential(nn.Conv2d(self.cin, width * block.expansion, 1, stride, bias=False),
                                 nn.BatchNorm2d(width * block.expansion))
        blocks = [block(self.cin, width, stride, down)]
        self.cin = width * block.expansion
        blocks += [block(self.cin, width) for _ in range(1, n)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        if self.fused:
            x = self.maxpool(conv_bn_act(x, self.conv1, self.bn1))
            x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
            return self.fc(torch.flatten(self.avgpool(x), 1))
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(**kw) -> ResNet:
    return ResNet(BasicBlock, [2, 2, 2, 2], **kw)


def resnet34(**kw) -> ResNet:
    return ResNet(BasicBlock, [3, 4, 6, 3], **kw)


def resnet50(**kw) -> ResNet:
    return ResNet(Bottleneck, [3, 4, 6, 3], **kw)


def resnet101(**kw) -> ResNet:
    return ResNet(Bottleneck, [3, 4, 23, 3], **kw)


def resnet152(**kw) -> ResNet:
    return ResNet(Bottleneck, [3, 8, 36, 3], **kw)
