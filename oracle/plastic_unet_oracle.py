"""ORACLE — test infrastructure only.  Never imported by the product path (plastic-unet_b200/).

A CPU restatement (plain PyTorch fp32/fp64 functional ops + a numpy restatement of the plastic head) of
the reference's Plastic U-Net hot path, driven by a ``state_dict`` so that it is independent of both
the reference's classes and of the B200 modules.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it, and only as the checker /
CPU baseline.

Pinning: ``oracle/make_golden.py`` imports the real reference (``/root/reference/src/unet``) in the
build container, checks that every function here reproduces it **bit-exactly** on CPU, and writes the
golden vectors under ``tests/golden/`` that the test-suite re-checks on every run (the reference has
no tests / golden vectors of its own).  The batched and coord-conv extensions
have no reference: they are pinned by composition (per-sample loops over pinned functions, analytic
coordinate values) — see tests/test_oracle.py.

Each function cites the reference file:line it follows (paths relative to /root/reference/).
"""
import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# plastic head + trace — src/unet/unet_p.py:69-88 (identical copy at src/unet/unet_p_res.py:115-134)
# ------------------------------------------------------------------------------------------------
def plastic_head(x_map, w, alpha, hebb, alfa_type='free'):
    """x_map: [nbf, nbf] (the 1-channel output viewed 2-D, unet_p.py:70) -> (activ, activout)."""
    if alfa_type == 'free':
        activ = x_map.mm(w + torch.mul(alpha, hebb))  # unet_p.py:73
    elif alfa_type == 'yoked':
        activ = x_map.mm(w + alpha * hebb)  # unet_p.py:75
    else:
        raise ValueError("Must select one plasticity coefficient type ('free' or 'yoked')")  # unet_p.py:77
    return activ, torch.sigmoid(activ)  # unet_p.py:79


def trace_update(hebb, activin, activout, eta, rule):
    """unet_p.py:81-86.  Only row 0 of the bmm survives the [0]."""
    if rule == 'hebb':
        return (1 - eta) * hebb + eta * torch.bmm(activin.unsqueeze(2), activout.unsqueeze(1))[0]  # unet_p.py:82
    elif rule == 'oja':
        return hebb + eta * torch.mul((activin[0].unsqueeze(1) - torch.mul(hebb, activout[0].unsqueeze(0))),
                                      activout[0].unsqueeze(0))  # unet_p.py:84
    raise ValueError("Must select one learning rule ('hebb' or 'oja')")  # unet_p.py:86


def trace_update_batched(hebb, x_maps, s_maps, eta, rule):
    """Batched extension (no reference): mean over samples of the per-sample reference update from the
    shared trace — both rules are affine in the per-sample term for a fixed hebb (SURVEY.md §8c recipe 2)."""
    outs = [trace_update(hebb, x_maps[b], s_maps[b], eta, rule) for b in range(x_maps.shape[0])]
    return torch.stack(outs, 0).mean(0)


def head_numpy(X, w, alpha, hebb, eta, rule):
    """numpy float64 restatement of head + row-0 trace for one map (closed forms of SURVEY.md §8a rows 8-10)."""
    X, w, alpha, hebb = (np.asarray(a, dtype=np.float64) for a in (X, w, alpha, hebb))
    eta = float(eta)
    A = X @ (w + alpha * hebb)
    S = 1.0 / (1.0 + np.exp(-A))
    x0, s0 = X[0], S[0]
    if rule == 'hebb':
        hn = (1.0 - eta) * hebb + eta * np.outer(x0, s0)
    elif rule == 'oja':
        hn = hebb + eta * (x0[:, None] - hebb * s0[None, :]) * s0[None, :]
    else:
        raise ValueError(rule)
    return A, S, hn


# ------------------------------------------------------------------------------------------------
# UNetp body — src/unet/unet_p.py:54-67, blocks :96-177
# ------------------------------------------------------------------------------------------------
def _bn(sd, key, x, training):
    """nn.BatchNorm2d (unet_p.py:106,109).  In training mode updates the running stats in `sd` in place."""
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'], sd[key + '.bias'],
                        training, 0.1, 1e-5)


def _double_conv(sd, p, x, batch_norm, training):
    """unet_p.py:96-122: (conv3x3 pad 1 => [BN] => ReLU) * 2; Sequential indices 0,2 (no BN) or 0,1,3,4 (BN)."""
    if batch_norm:
        x = F.conv2d(x, sd[p + '.conv.0.weight'], sd[p + '.conv.0.bias'], padding=1)
        x = F.relu(_bn(sd, p + '.conv.1', x, training))
        x = F.conv2d(x, sd[p + '.conv.3.weight'], sd[p + '.conv.3.bias'], padding=1)
        return F.relu(_bn(sd, p + '.conv.4', x, training))
    x = F.relu(F.conv2d(x, sd[p + '.conv.0.weight'], sd[p + '.conv.0.bias'], padding=1))
    return F.relu(F.conv2d(x, sd[p + '.conv.2.weight'], sd[p + '.conv.2.bias'], padding=1))


def _up(sd, p, x1, x2, batch_norm, bilinear, training):
    """unet_p.py:159-167."""
    if bilinear:
        x1 = F.interpolate(x1, scale_factor=2, mode='bilinear', align_corners=True)  # unet_p.py:153
    else:
        x1 = F.conv_transpose2d(x1, sd[p + '.up.weight'], sd[p + '.up.bias'], stride=2)  # unet_p.py:155
    diffX = x1.size()[2] - x2.size()[2]
    diffY = x1.size()[3] - x2.size()[3]
    x2 = F.pad(x2, (diffX // 2, int(diffX / 2), diffY // 2, int(diffY / 2)))  # unet_p.py:163-164
    x = torch.cat([x2, x1], dim=1)  # unet_p.py:165
    return _double_conv(sd, p + '.conv', x, batch_norm, training)


def unetp_body(sd, x, batch_norm=False, bilinear=False, training=True, depth=4):
    """x [B,C,H,W] -> 1-channel map [B,n_classes,H',W'] (unet_p.py:58-67).  depth=4 is the reference."""
    feats = [_double_conv(sd, 'inc.conv', x, batch_norm, training)]  # unet_p.py:58
    for k in range(1, depth + 1):
        feats.append(_double_conv(sd, 'down%d.mpconv.1' % k, F.max_pool2d(feats[-1], 2), batch_norm, training))  # :59-62
    y = feats[-1]
    for j in range(1, depth + 1):
        y = _up(sd, 'up%d' % j, y, feats[depth - j], batch_norm, bilinear, training)  # :63-66
    return F.conv2d(y, sd['outc.conv.weight'], sd['outc.conv.bias'])  # :67


# ------------------------------------------------------------------------------------------------
# UNetpRes body — src/unet/unet_p_res.py:71-113, blocks :142-272
# ------------------------------------------------------------------------------------------------
def _residual_block(sd, p, x, batch_norm, training):
    """unet_p_res.py:166-189.  Leading ReLU is in place => the skip adds relu(input)."""
    r = F.relu(x)
    if batch_norm:
        b = _bn(sd, p + '.conv.1', r, training)
        i1, i2 = 2, 3
    else:
        b = r
        i1, i2 = 1, 2
    a = F.relu(F.conv2d(b, sd['%s.conv.%d.conv.weight' % (p, i1)], sd['%s.conv.%d.conv.bias' % (p, i1)], padding=1))  # conv_module :142-164
    y = F.conv2d(a, sd['%s.conv.%d.conv.weight' % (p, i2)], sd['%s.conv.%d.conv.bias' % (p, i2)], padding=1)
    return y.add(r)  # unet_p_res.py:188


def _res_stack(sd, p, x, batch_norm, training):
    """`down` (unet_p_res.py:256-272) and `middle` (:223-238): conv3x3 -> residual_block x2 -> ReLU."""
    x = F.conv2d(x, sd[p + '.0.weight'], sd[p + '.0.bias'], padding=1)
    x = _residual_block(sd, p + '.1', x, batch_norm, training)
    x = _residual_block(sd, p + '.2', x, batch_norm, training)
    return F.relu(x)


def _drop2d(x, p, training, masks):
    """nn.Dropout2d (unet_p_res.py:209,248).  `masks`: None -> draw with torch's RNG like the reference;
    else a list that is consumed front to back ([B,C] scale tensors) so that tests can inject the noise."""
    if not training or p == 0:
        return x
    if masks is None:
        return F.dropout2d(x, p, True)
    return x * masks.pop(0).view(x.shape[0], x.shape[1], 1, 1)


def _res_up(sd, p, x1, x2, dropout_ratio, training, masks):
    """unet_p_res.py:213-220."""
    x = F.conv_transpose2d(x1, sd[p + '.dconv.weight'], sd[p + '.dconv.bias'], stride=2)  # :207, :214
    diffX = x2.size()[2] - x.size()[2]
    diffY = x2.size()[3] - x.size()[3]
    x = F.pad(x, (diffX // 2, int(diffX / 2), diffY // 2, int(diffY / 2)))  # :217
    x = torch.cat([x, x2], dim=1)  # :218
    x = _drop2d(x, dropout_ratio, training, masks)  # :209
    return _res_stack(sd, p + '.uconv.1.mconv', x, False, training)  # middle(batch_norm=False) :210


def unetpres_body(sd, x, dropout_ratio=0.5, batch_norm=False, training=True, masks=None, depth=4):
    """unet_p_res.py:71-113.  depth=4 is the reference."""
    skips = []
    for k in range(1, depth + 1):
        xc = _res_stack(sd, 'conv%d.dconv' % k, x, batch_norm, training)  # :73,78,83,88
        skips.append(xc)
        x = F.max_pool2d(xc, 2)  # pool_drop :240-253
        x = _drop2d(x, dropout_ratio / 2 if k == 1 else dropout_ratio, training, masks)  # :39,42,45,48
    y = _res_stack(sd, 'mid.mconv', x, batch_norm, training)  # :93
    for k in range(depth, 0, -1):
        y = _res_up(sd, 'uconv%d' % k, y, skips[k - 1], dropout_ratio, training, masks)  # :97-110
    return F.conv2d(y, sd['outc.conv.weight'], sd['outc.conv.bias'])  # :112


# ------------------------------------------------------------------------------------------------
# UNetpCoord body — src/coord_conv_script.py:69-96 (AddCoords), :104-126 (CoordConv), :153-194 (topology)
# (Keras/TF, NHWC, not runnable here: parity unpinned by the reference; pinned analytically in tests)
# ------------------------------------------------------------------------------------------------
def add_coords(x, with_r=False):
    """coord_conv_script.py:69-96 restated for NCHW: xx varies along width (j), yy along height (i);
    value = 2*index/(dim-1) - 1; rr = sqrt((xx-0.5)^2 + (yy-0.5)^2) (:93-95)."""
    B, _, H, W = x.shape
    j = torch.arange(W, dtype=torch.float32, device=x.device)
    i = torch.arange(H, dtype=torch.float32, device=x.device)
    xx = (j / (W - 1) * 2 - 1).view(1, 1, 1, W).expand(B, 1, H, W)
    yy = (i / (H - 1) * 2 - 1).view(1, 1, H, 1).expand(B, 1, H, W)
    chans = [x, xx.to(x.dtype), yy.to(x.dtype)]
    if with_r:
        chans.append(torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2).to(x.dtype))
    return torch.cat(chans, dim=1)


def unetpcoord_body(sd, x, with_r=False, depth=4):
    """coord_conv_script.py:153-194 with the sigmoid output conv replaced by a linear 1x1 conv feeding the plastic head."""
    y = F.relu(F.conv2d(add_coords(x, with_r), sd['stem.conv.weight'], sd['stem.conv.bias']))  # :153
    feats = [_double_conv(sd, 'enc0', y, False, False)]  # :155-156
    for k in range(1, depth + 1):
        feats.append(_double_conv(sd, 'enc%d.mpconv.1' % k, F.max_pool2d(feats[-1], 2), False, False))  # :157-172
    y = feats[-1]
    for k in range(depth, 0, -1):
        u = F.conv_transpose2d(y, sd['dec%d.up.weight' % k], sd['dec%d.up.bias' % k], stride=2)  # :174,179,184,189
        y = _double_conv(sd, 'dec%d.conv' % k, torch.cat([u, feats[k - 1]], dim=1), False, False)  # :175-192
    return F.conv2d(y, sd['outc.conv.weight'], sd['outc.conv.bias'])  # :194 (without the sigmoid)


# ------------------------------------------------------------------------------------------------
# whole-model forward: body + head + trace
# ------------------------------------------------------------------------------------------------
BODIES = {'unetp': unetp_body, 'unetpres': unetpres_body, 'unetpcoord': unetpcoord_body}


def forward(kind, sd, x, hebb, rule='hebb', alfa_type='free', **body_kw):
    """-> (activ logits, activout, hebb').  B == 1: 2-D [nbf,nbf] outputs exactly as the reference
    (unet_p.py:70,79,88).  B > 1: batched extension — outputs [B,nbf,nbf], shared hebb, mean trace update."""
    o = BODIES[kind](sd, x, **body_kw)
    B = x.shape[0]
    nbf = sd['w'].shape[0]
    if B == 1:
        activin = o.view(nbf, nbf)  # unet_p.py:70
        activ, activout = plastic_head(activin, sd['w'], sd['alpha'], hebb, alfa_type)
        return activ, activout, trace_update(hebb, activin, activout, sd['eta'], rule)
    maps = o.view(B, nbf, nbf)
    pairs = [plastic_head(maps[b], sd['w'], sd['alpha'], hebb, alfa_type) for b in range(B)]
    activ = torch.stack([p[0] for p in pairs], 0)
    activout = torch.stack([p[1] for p in pairs], 0)
    return activ, activout, trace_update_batched(hebb, maps, activout, sd['eta'], rule)


def init_state_unetp(nbf=128, seed=0, base=8, depth=4):
    """A freshly initialised UNetp state_dict with the reference's shapes and init distributions (unet_p.py:30-32 for the
    plastic parameters, nn.Conv2d / nn.ConvTranspose2d defaults for the body, unet_p.py:34-47) built from stock torch.nn
    layers — for the CPU-baseline timing leg when neither the reference nor its bytecode (oracle/_ref) is present."""
    torch.manual_seed(seed)
    sd = {'w': .01 * torch.randn(nbf, nbf), 'alpha': .01 * torch.rand(nbf, nbf), 'eta': .01 * torch.ones(1)}

    def conv(prefix, cin, cout, k=3):
        m = torch.nn.Conv2d(cin, cout, k, padding=k // 2)
        sd[prefix + '.weight'], sd[prefix + '.bias'] = m.weight.detach(), m.bias.detach()

    def dconv(prefix, cin, cout):
        conv(prefix + '.conv.0', cin, cout)
        conv(prefix + '.conv.2', cout, cout)

    c = [base * 2 ** i for i in range(depth)] + [base * 2 ** (depth - 1)]
    dconv('inc.conv', 1, c[0])
    for k in range(1, depth + 1):
        dconv('down%d.mpconv.1' % k, c[k - 1], c[k])
    for j in range(1, depth + 1):
        skip = c[depth - j]
        out = c[depth - j - 1] if depth - j - 1 >= 0 else c[0]
        m = torch.nn.ConvTranspose2d(skip, skip, 2, stride=2)
        sd['up%d.up.weight' % j], sd['up%d.up.bias' % j] = m.weight.detach(), m.bias.detach()
        dconv('up%d.conv' % j, 2 * skip, out)
    m = torch.nn.Conv2d(c[0], 1, 1)
    sd['outc.conv.weight'], sd['outc.conv.bias'] = m.weight.detach(), m.bias.detach()
    return sd


def leaf_state(sd, dtype=torch.float32, requires_grad=True):
    """Clone a state_dict into leaf tensors (float params get requires_grad) for autograd through `forward`."""
    out = {}
    for k, v in sd.items():
        if v.is_floating_point():
            t = v.detach().clone().to(dtype)
            if requires_grad and not (k.endswith('running_mean') or k.endswith('running_var')):
                t.requires_grad_(True)
            out[k] = t
        else:
            out[k] = v.detach().clone()
    return out


# ------------------------------------------------------------------------------------------------
# the training loop the metric times — src/train.py:66-70, :88-112
# ------------------------------------------------------------------------------------------------
def train_steps(kind, sd, images, masks, rule, lr=1e-4, gamma=1.0, steplr=1e9, hebb=None, **body_kw):
    """B=1 sequential steps exactly as train.py:91-112 (zero_grad, fwd with detached hebb, BCELoss, .item(),
    backward, Adam.step, StepLR.step).  `sd` must be a leaf_state(); it is updated in place.
    -> (losses, hebb)"""
    params = [v for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1.0 * lr)  # train.py:66
    sched = torch.optim.lr_scheduler.StepLR(opt, gamma=gamma, step_size=steplr)  # :67-68
    crit = torch.nn.BCELoss()  # :70
    nbf = sd['w'].shape[0]
    if hebb is None:
        hebb = torch.zeros(nbf, nbf, dtype=sd['w'].dtype)  # train.py:88
    losses = []
    for img, mask in zip(images, masks):  # train.py:91
        opt.zero_grad()
        _, y_pred, hebb = forward(kind, sd, img[None], hebb.detach(), rule=rule, **body_kw)  # :99
        loss = crit(y_pred.view(-1), mask.view(-1))  # :101-105
        losses.append(loss.item())  # :106
        loss.backward()  # :110
        opt.step()  # :111
        sched.step()  # :112
    return losses, hebb.detach()


def train_steps_batched(kind, sd, batches, rule, lr=1e-4, hebb=None, **body_kw):
    """Batched extension of train.py:91-112 (no reference: the reference is B=1 only, unet_p.py:55-56): per step one
    batch [B,C,H,W] through `forward` (shared trace, mean of the per-sample trace updates), BCELoss over all B*nbf^2
    outputs, backward, Adam.step; the trace is carried detached (train.py:99).  At B == 1 this is train_steps().
    `batches`: iterable of (x [B,C,H,W], target [B,nbf,nbf]).  -> (losses, hebb)"""
    params = [v for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(params, lr=1.0 * lr)  # train.py:66
    crit = torch.nn.BCELoss()  # :70
    nbf = sd['w'].shape[0]
    if hebb is None:
        hebb = torch.zeros(nbf, nbf, dtype=sd['w'].dtype)  # train.py:88
    losses = []
    for x, target in batches:
        opt.zero_grad()
        _, y_pred, hebb = forward(kind, sd, x, hebb.detach(), rule=rule, **body_kw)  # :99
        loss = crit(y_pred.reshape(-1), target.reshape(-1))  # :101-105
        losses.append(loss.item())  # :106
        loss.backward()  # :110
        opt.step()  # :111
    return losses, hebb.detach()


def bce_mean(pred, target):
    """nn.BCELoss (train.py:70): mean over elements of -(t*max(log p,-100) + (1-t)*max(log(1-p),-100))."""
    return F.binary_cross_entropy(pred, target)


def pad_101_to_128(x):
    """Input construction named by BASELINE.json configs: zero-pad 101x101 to 128x128, 13 px top/left,
    14 px bottom/right (the reference has no padding step; SURVEY.md §8d defines it)."""
    return F.pad(x, (13, 14, 13, 14))


def tol_report(a, b):
    """max-abs error relative to max|b| and L2-relative error (the two figures the parity tests bound)."""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom, float((a - b).norm()) / max(float(b.norm()), 1e-30)

