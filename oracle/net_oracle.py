"""TEST INFRASTRUCTURE ONLY -- CPU (or any-device) fp32 torch restatement of the two
neural stages, driven purely by a reference-format ``state_dict``.

Pinned by tests/test_oracle_pins.py against the reference modules executed in the
build container on the seeded synthetic checkpoint (tests/golden/ref_hrnet.npz,
ref_head.npz).  Real-checkpoint parity is UNPINNED (SHAPY_A weights are a licensed
download).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.

  hrnet_forward   regressor/human_shape/models/backbone/hrnet.py:426-498 (+175-193 fuse,
                  torchvision BasicBlock / Bottleneck semantics, hrnet.py:13)
  head_forward    regressor/human_shape/models/common/networks.py:536-592 (+MLP 387-396)
"""
import torch
import torch.nn.functional as F

STAGES = {  # config/network_defaults.py:92-132
    2: dict(num_modules=1, channels=(48, 96)),
    3: dict(num_modules=4, channels=(48, 96, 192)),
    4: dict(num_modules=3, channels=(48, 96, 192, 384)),
}
EPS = 1e-5


def _conv(sd, name, x, stride=1):
    w = sd[name + '.weight']
    b = sd.get(name + '.bias')
    return F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2)


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'],
                        sd[name + '.bias'], False, 0.0, EPS)


def _basic(sd, p, x):
    out = F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x)))
    out = _bn(sd, p + '.bn2', _conv(sd, p + '.conv2', out))
    return F.relu(out + x)


def _bottleneck(sd, p, x):
    out = F.relu(_bn(sd, p + '.bn1', _conv(sd, p + '.conv1', x)))
    out = F.relu(_bn(sd, p + '.bn2', _conv(sd, p + '.conv2', out)))
    out = _bn(sd, p + '.bn3', _conv(sd, p + '.conv3', out))
    if (p + '.downsample.0.weight') in sd:        # layer1.0: conv + BN
        idt = _bn(sd, p + '.downsample.1', _conv(sd, p + '.downsample.0', x))
    elif (p + '.downsample.weight') in sd:        # conv_layers.N: bare 1x1 conv, hrnet.py:361-373
        idt = _conv(sd, p + '.downsample', x)
    else:
        idt = x
    return F.relu(out + idt)


def _hr_module(sd, p, xs):
    n = len(xs)
    xs = list(xs)
    for i in range(n):
        for k in range(4):
            xs[i] = _basic(sd, f'{p}.branches.{i}.{k}', xs[i])
    outs = []
    for i in range(n):
        y = None
        for j in range(n):
            if j == i:
                t = xs[j]
            elif j > i:
                q = f'{p}.fuse_layers.{i}.{j}'
                t = _bn(sd, q + '.1', _conv(sd, q + '.0', xs[j]))
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
            else:
                t = xs[j]
                for k in range(i - j):
                    q = f'{p}.fuse_layers.{i}.{j}.{k}'
                    t = _bn(sd, q + '.1', _conv(sd, q + '.0', t, stride=2))
                    if k != i - j - 1:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hrnet_forward(sd: dict, x: torch.Tensor) -> dict:
    x = F.relu(_bn(sd, 'bn1', _conv(sd, 'conv1', x, stride=2)))
    x = F.relu(_bn(sd, 'bn2', _conv(sd, 'conv2', x, stride=2)))
    for k in range(4):
        x = _bottleneck(sd, f'layer1.{k}', x)
    # transition1: [3x3 256->48 ; 3x3 s2 256->96]
    xs = [F.relu(_bn(sd, 'transition1.0.1', _conv(sd, 'transition1.0.0', x))),
          F.relu(_bn(sd, 'transition1.1.0.1', _conv(sd, 'transition1.1.0.0', x, stride=2)))]
    for m in range(STAGES[2]['num_modules']):
        xs = _hr_module(sd, f'stage2.{m}', xs)
    xs = xs + [F.relu(_bn(sd, 'transition2.2.0.1', _conv(sd, 'transition2.2.0.0', xs[-1], stride=2)))]
    for m in range(STAGES[3]['num_modules']):
        xs = _hr_module(sd, f'stage3.{m}', xs)
    xs = xs + [F.relu(_bn(sd, 'transition3.3.0.1', _conv(sd, 'transition3.3.0.0', xs[-1], stride=2)))]
    for m in range(STAGES[4]['num_modules']):
        xs = _hr_module(sd, f'stage4.{m}', xs)
    out = {f'layer{i + 1}': t for i, t in enumerate(xs)}

    def subsample(name, t, n):
        for i in range(n):
            t = F.relu(_bn(sd, f'{name}.{3 * i + 1}', _conv(sd, f'{name}.{3 * i}', t, stride=2)))
        return t
    feat = torch.cat([subsample('subsample_4', xs[0], 3), subsample('subsample_3', xs[1], 2),
                      subsample('subsample_2', xs[2], 1), xs[3]], dim=1)
    for k in range(5):
        feat = _bottleneck(sd, f'conv_layers.{k}', feat)
    out['concat'] = feat.mean(dim=(2, 3))
    return out


def head_forward(sd: dict, feats: torch.Tensor, num_stages: int = 3, prefix: str = 'regressor.'):
    """p_0 = mean; p_{k+1} = p_k + MLP(cat[f, p_k]); MLP = 3 Linear, no activation (eval)."""
    def mlp(x):
        x = F.linear(x, sd[prefix + 'module.layer_000.0.weight'], sd[prefix + 'module.layer_000.0.bias'])
        x = F.linear(x, sd[prefix + 'module.layer_001.0.weight'], sd[prefix + 'module.layer_001.0.bias'])
        return F.linear(x, sd[prefix + 'module.output_layer.weight'], sd[prefix + 'module.output_layer.bias'])
    p = sd[prefix + 'mean_param'].expand(feats.shape[0], -1)
    outs = []
    for _ in range(num_stages):
        p = p + mlp(torch.cat([feats, p], dim=1))
        outs.append(p)
    return outs
