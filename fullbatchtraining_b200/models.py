"""Model construction: the drop-in for ``fullbatch.models.construct_model`` (reference fullbatch/models/models.py:14-52)
for the ResNet family named by the hot path (CIFAR stem, BatchNorm2d, ReLU, downsample 'C').

The returned ``torch.nn.Module`` has the reference's module tree, hence the same ``state_dict`` keys
(``stem.0.weight``, ``layers.{s}.{b}.conv1.weight``, ``...downsample.{1,2}...``, ``fc.weight`` -- the checkpoint format of
fullbatch/training/utils.py:43-51 and hubconf.py:37-40) and the same ``parameters()`` order (the flat-buffer order of
fullbatch/training/utils.py:34).  Modules are created in the same order and initialised with the same calls as
fullbatch/models/resnets.py:45-126, so ``torch.manual_seed(s); construct_model(...)`` gives bit-identical initial
weights (checked in tests/test_models.py against the golden fingerprints of the reference).

The module's own ``forward`` is plain PyTorch and is only used for evaluation / debugging; training goes through
``engine.FullBatchEngine`` which runs the sm_100a kernels on the same parameter storage.
"""
import torch

_DEPTHS = {
    18: ("basic", [2, 2, 2, 2]),
    34: ("basic", [3, 4, 6, 3]),
    50: ("bottleneck", [3, 4, 6, 3]),
    101: ("bottleneck", [3, 4, 23, 3]),
    152: ("bottleneck", [3, 8, 36, 3]),
}


def _conv(cin, cout, k, stride=1):
    return torch.nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=False)


class ResidualBlock(torch.nn.Module):
    """BasicBlock (expansion 1) or Bottleneck (expansion 4, stride on the 3x3): resnets.py:195-230 / :271-316."""

    def __init__(self, kind, inplanes, planes, stride, downsample):
        super().__init__()
        self.kind = kind
        if kind == "basic":
            self.conv1 = _conv(inplanes, planes, 3, stride)
            self.bn1 = torch.nn.BatchNorm2d(planes)
            self.nonlin = torch.nn.ReLU(inplace=True)
            self.conv2 = _conv(planes, planes, 3)
            self.bn2 = torch.nn.BatchNorm2d(planes)
        else:
            self.conv1 = _conv(inplanes, planes, 1)
            self.bn1 = torch.nn.BatchNorm2d(planes)
            self.conv2 = _conv(planes, planes, 3, stride)
            self.bn2 = torch.nn.BatchNorm2d(planes)
            self.conv3 = _conv(planes, planes * 4, 1)
            self.bn3 = torch.nn.BatchNorm2d(planes * 4)
            self.nonlin = torch.nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def conv_bn_pairs(self):
        pairs = [("conv1", "bn1"), ("conv2", "bn2")]
        if self.kind != "basic":
            pairs.append(("conv3", "bn3"))
        return pairs

    def forward(self, x):
        out = x
        pairs = self.conv_bn_pairs()
        for i, (c, b) in enumerate(pairs):
            out = getattr(self, b)(getattr(self, c)(out))
            if i + 1 < len(pairs):
                out = self.nonlin(out)
        identity = x if self.downsample is None else self.downsample(x)
        return self.nonlin(out + identity)


class ResNet(torch.nn.Module):
    def __init__(self, depth, channels=3, classes=10, zero_init_residual=False):
        super().__init__()
        if depth not in _DEPTHS:
            raise ValueError(f"ResNet depth {depth} is not supported by the B200 path")
        kind, layers = _DEPTHS[depth]
        self.depth, self.kind = depth, kind
        expansion = 1 if kind == "basic" else 4
        inplanes = 64
        self.stem = torch.nn.Sequential(_conv(channels, inplanes, 3), torch.nn.BatchNorm2d(inplanes),
                                        torch.nn.ReLU(inplace=True))
        stages = []
        width = inplanes
        for s, nblocks in enumerate(layers):
            stride = 1 if s == 0 else 2
            planes = width
            downsample = None
            if stride != 1 or inplanes != planes * expansion:
                # downsample 'C' (resnets.py:147-152); built before the blocks of the stage like the reference does
                downsample = torch.nn.Sequential(torch.nn.AvgPool2d(kernel_size=stride, stride=stride),
                                                 _conv(inplanes, planes * expansion, 1),
                                                 torch.nn.BatchNorm2d(planes * expansion))
            blocks = [ResidualBlock(kind, inplanes, planes, stride, downsample)]
            inplanes = planes * expansion
            for _ in range(1, nblocks):
                blocks.append(ResidualBlock(kind, inplanes, planes, 1, None))
            stages.append(torch.nn.Sequential(*blocks))
            width *= 2
        self.layers = torch.nn.Sequential(*stages)
        self.avgpool = torch.nn.AdaptiveAvgPool2d((1, 1))
        self.fc = torch.nn.Linear(inplanes, classes)
        for m in self.modules():  # resnets.py:109-114
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, torch.nn.BatchNorm2d):
                torch.nn.init.constant_(m.weight, 1)
                torch.nn.init.constant_(m.bias, 0)
        if zero_init_residual:  # resnets.py:119-126
            for m in self.modules():
                if isinstance(m, ResidualBlock):
                    torch.nn.init.constant_(getattr(m, m.conv_bn_pairs()[-1][1]).weight, 0)

    def forward(self, x):
        x = self.layers(self.stem(x))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def construct_model(cfg_model, channels=3, classes=10):
    """Same call as the reference factory (models.py:14).  ``cfg_model`` needs ``name`` and ``depth``; the reference keys
    ``stem / convolution / nonlin_fn / normalization / downsample`` are validated against the only combination the
    accelerated path implements.  ``zero_init_residual`` reproduces the reference test for the substring
    'skip_residual' in ``initialization`` (models.py:22), which is False for the shipped configs ('skip-residual')."""
    get = (lambda k, d=None: cfg_model.get(k, d)) if isinstance(cfg_model, dict) else \
        (lambda k, d=None: getattr(cfg_model, k, d))
    name = str(get("name", "")).lower()
    if "resnet" not in name:
        raise ValueError(f"model {get('name')} is not on the accelerated path (ResNet family only, no fallback)")
    expected = dict(stem="CIFAR", convolution="Standard", nonlin_fn="ReLU", normalization="BatchNorm2d", downsample="C")
    for key, val in expected.items():
        got = get(key, val)
        if str(got).lower() != val.lower():
            raise ValueError(f"model.{key}={got!r} is not supported by the B200 path (only {val!r})")
    zero_init = "skip_residual" in str(get("initialization", ""))
    return ResNet(int(get("depth")), channels, classes, zero_init_residual=zero_init)
