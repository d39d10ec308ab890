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


def _cb(sd, conv, bn, x, stride=1, rnd=None):
    """conv (+ eval-mode BN).  rnd=None: the reference arithmetic (conv, then BN).  rnd given (tests of the
    plain-fp16 mode, BASELINE config 2): BN scale folded into the weights as the product does, both operands
    rounded by `rnd` (fp16), fp32 accumulation, shift added in fp32."""
    if rnd is None:
        y = _conv(sd, conv, x, stride)
        return _bn(sd, bn, y) if bn is not None else y
    w = sd[conv + '.weight']
    b = sd.get(conv + '.bias')
    b = torch.zeros(w.shape[0], dtype=w.dtype, device=w.device) if b is None else b
    if bn is not None:
        s = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + EPS)
        w = w * s[:, None, None, None]
        b = (b - sd[bn + '.running_mean']) * s + sd[bn + '.bias']
    y = F.conv2d(rnd(x), rnd(w), None, stride=stride, padding=w.shape[-1] // 2)
    return y + b[None, :, None, None]


def _id(t):
    return t


def _basic(sd, p, x, rnd=None):
    out = F.relu(_cb(sd, p + '.conv1', p + '.bn1', x, 1, rnd))
    out = _cb(sd, p + '.conv2', p + '.bn2', out, 1, rnd)
    return F.relu(out + (rnd or _id)(x))


def _bottleneck(sd, p, x, rnd=None):
    out = F.relu(_cb(sd, p + '.conv1', p + '.bn1', x, 1, rnd))
    out = F.relu(_cb(sd, p + '.conv2', p + '.bn2', out, 1, rnd))
    out = _cb(sd, p + '.conv3', p + '.bn3', out, 1, rnd)
    if (p + '.downsample.0.weight') in sd:        # layer1.0: conv + BN
        idt = _cb(sd, p + '.downsample.0', p + '.downsample.1', x, 1, rnd)
    elif (p + '.downsample.weight') in sd:        # conv_layers.N: bare 1x1 conv, hrnet.py:361-373
        idt = _cb(sd, p + '.downsample', None, x, 1, rnd)
    else:
        idt = x
    return F.relu(out + (rnd or _id)(idt))


def _hr_module(sd, p, xs, rnd=None):
    r = rnd or _id
    n = len(xs)
    xs = list(xs)
    for i in range(n):
        for k in range(4):
            xs[i] = _basic(sd, f'{p}.branches.{i}.{k}', xs[i], rnd)
    outs = []
    for i in range(n):
        y = None
        for j in range(n):
            if j == i:
                t = r(xs[j])
            elif j > i:
                q = f'{p}.fuse_layers.{i}.{j}'
                t = _cb(sd, q + '.0', q + '.1', xs[j], 1, rnd)
                t = F.interpolate(r(t), scale_factor=2 ** (j - i), mode='nearest')
            else:
                t = xs[j]
                for k in range(i - j):
                    q = f'{p}.fuse_layers.{i}.{j}.{k}'
                    t = _cb(sd, q + '.0', q + '.1', t, 2, rnd)
                    if k != i - j - 1:
                        t = F.relu(t)
                t = r(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hrnet_forward(sd: dict, x: torch.Tensor, rnd=None) -> dict:
    """rnd=None: reference arithmetic.  rnd = fp16 round trip: emulates the product's plain-fp16 mode (every stored
    activation and every conv operand rounded to fp16, fp32 accumulation; the stem conv reads fp32 images)."""
    r = rnd or _id
    x = F.relu(_cb(sd, 'conv1', 'bn1', x, 2, None))
    x = F.relu(_cb(sd, 'conv2', 'bn2', x, 2, rnd))
    for k in range(4):
        x = _bottleneck(sd, f'layer1.{k}', x, rnd)
    # transition1: [3x3 256->48 ; 3x3 s2 256->96]
    xs = [F.relu(_cb(sd, 'transition1.0.0', 'transition1.0.1', x, 1, rnd)),
          F.relu(_cb(sd, 'transition1.1.0.0', 'transition1.1.0.1', x, 2, rnd))]
    for m in range(STAGES[2]['num_modules']):
        xs = _hr_module(sd, f'stage2.{m}', xs, rnd)
    xs = xs + [F.relu(_cb(sd, 'transition2.2.0.0', 'transition2.2.0.1', xs[-1], 2, rnd))]
    for m in range(STAGES[3]['num_modules']):
        xs = _hr_module(sd, f'stage3.{m}', xs, rnd)
    xs = xs + [F.relu(_cb(sd, 'transition3.3.0.0', 'transition3.3.0.1', xs[-1], 2, rnd))]
    for m in range(STAGES[4]['num_modules']):
        xs = _hr_module(sd, f'stage4.{m}', xs, rnd)
    out = {f'layer{i + 1}': r(t) for i, t in enumerate(xs)}

    def subsample(name, t, n):
        for i in range(n):
            t = F.relu(_cb(sd, f'{name}.{3 * i}', f'{name}.{3 * i + 1}', t, 2, rnd))
        return t
    feat = torch.cat([subsample('subsample_4', xs[0], 3), subsample('subsample_3', xs[1], 2),
                      subsample('subsample_2', xs[2], 1), r(xs[3])], dim=1)
    for k in range(5):
        feat = _bottleneck(sd, f'conv_layers.{k}', feat, rnd)
    out['concat'] = r(feat).mean(dim=(2, 3))
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
