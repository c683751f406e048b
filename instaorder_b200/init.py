"""Fresh-model initialisation identical to the reference constructor.

``models.__dict__[algo](params)`` of the reference builds ``resnet50_cls(in_channels, num_classes)`` and then calls
``utils.init_weights(model, 'xavier')`` (reference models/single_stage_model.py:24-25).  What a fresh model holds is
therefore the result of three passes over the network, all drawing from torch's global CPU generator:

1. the ``nn.Conv2d`` / ``nn.Linear`` constructors' own default ``reset_parameters`` (kaiming-uniform with
   ``a = sqrt(5)``, the linear bias uniform in +-1/sqrt(fan_in)) in *construction* order -- a layer's downsample
   convolution is built before the blocks of that layer (reference models/backbone/resnet_cls.py:180-199);
2. ``kaiming_normal_(fan_out, relu)`` on every convolution in ``modules()`` order (resnet_cls.py:162-167);
3. ``init_weights``: ``xavier_normal_(gain=0.02)`` on every convolution / linear weight, zero linear bias,
   ``normal_(1, 0.02)`` on every BatchNorm weight, zero BatchNorm bias, in ``Module.apply`` order
   (utils/common_utils.py:35-65).

Only pass 3 decides the values, but passes 1 and 2 advance the generator, so all three are replayed here on scratch
tensors of the reference's shapes.  After ``torch.manual_seed(s)`` the returned ``state_dict`` equals the reference
constructor's bit for bit (tests/test_init.py compares it with the unmodified reference).  Host-side, once per model.
"""
import math

import torch
from torch.nn import init as tinit

_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))      # (planes, blocks, stride): ResNet-50


def _modules(in_channels, num_classes):
    """(name, kind, shape) of every parameterised module in ``modules()`` / ``apply`` order, and the same list in
    constructor order."""
    reg, ctor = [], []

    def conv(name, cout, cin, k, into):
        into.append((name, "conv", (cout, cin, k, k)))

    def bn(name, c, into):
        into.append((name, "bn", (c,)))

    conv("conv1", 64, in_channels, 7, reg)
    bn("bn1", 64, reg)
    ctor.extend(reg)
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(_LAYERS, 1):
        for b in range(blocks):
            p = "layer%d.%d." % (li, b)
            block, ds = [], []
            conv(p + "conv1", planes, inplanes, 1, block)
            bn(p + "bn1", planes, block)
            conv(p + "conv2", planes, planes, 3, block)
            bn(p + "bn2", planes, block)
            conv(p + "conv3", planes * 4, planes, 1, block)
            bn(p + "bn3", planes * 4, block)
            if b == 0:                                   # stride != 1 or inplanes != planes * 4: always for block 0
                conv(p + "downsample.0", planes * 4, inplanes, 1, ds)
                bn(p + "downsample.1", planes * 4, ds)
            reg.extend(block + ds)                       # attribute order inside Bottleneck: ... bn3, relu, downsample
            ctor.extend(ds + block)                      # _make_layer builds the downsample branch first
            inplanes = planes * 4
    heads = []
    if isinstance(num_classes, (list, tuple)):           # resnet_cls.py:153-160
        heads.append(("fc_occ", "linear", (int(num_classes[0]), 2048)))
        heads.append(("fc_depth", "linear", (int(num_classes[1]), 2048)))
    else:
        heads.append(("fc", "linear", (int(num_classes), 2048)))
    reg.extend(heads)
    ctor.extend(heads)
    return reg, ctor


def reference_init_state_dict(num_classes, in_channels=5, prefix="module."):
    """The ``state_dict`` (reference key names, fp32 CPU tensors, BN buffers included) of a freshly constructed
    reference model, drawn from torch's global CPU generator exactly as the reference constructor draws it."""
    reg, ctor = _modules(in_channels, num_classes)
    # pass 1: constructor defaults (values discarded, generator advanced)
    for _, kind, shape in ctor:
        if kind == "conv":
            tinit.kaiming_uniform_(torch.empty(shape), a=math.sqrt(5))
        elif kind == "linear":
            tinit.kaiming_uniform_(torch.empty(shape), a=math.sqrt(5))
            bound = 1.0 / math.sqrt(shape[1])
            tinit.uniform_(torch.empty(shape[0]), -bound, bound)
    # pass 2: resnet_cls.py:162-167
    for _, kind, shape in reg:
        if kind == "conv":
            tinit.kaiming_normal_(torch.empty(shape), mode="fan_out", nonlinearity="relu")
    # pass 3: init_weights(model, 'xavier'), gain 0.02
    sd = {}
    for name, kind, shape in reg:
        if kind == "bn":
            w = torch.empty(shape)
            tinit.normal_(w, 1.0, 0.02)
            sd[prefix + name + ".weight"] = w
            sd[prefix + name + ".bias"] = torch.zeros(shape)
            sd[prefix + name + ".running_mean"] = torch.zeros(shape)
            sd[prefix + name + ".running_var"] = torch.ones(shape)
            sd[prefix + name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
        else:
            w = torch.empty(shape)
            tinit.xavier_normal_(w, gain=0.02)
            sd[prefix + name + ".weight"] = w
            if kind == "linear":
                sd[prefix + name + ".bias"] = torch.zeros(shape[0])
    return sd
