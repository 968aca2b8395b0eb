"""Generate tests/golden/*.npz from the REAL reference and pin the oracle to it.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

For every case it (1) builds the unmodified reference module (``from unet import UNetp, UNetpRes``),
(2) runs forward + backward on CPU fp32 with seeded inputs, (3) asserts that
``oracle/plastic_unet_oracle.py`` driven by the same state_dict reproduces the reference **bit-exactly**
(outputs, trace, every parameter gradient), and (4) stores state_dict + inputs + outputs + gradients.
The test-suite then checks oracle-vs-golden (CPU) and CUDA-vs-golden / CUDA-vs-oracle (GPU).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference/src')

import plastic_unet_oracle as orc  # noqa: E402
from unet import UNetp, UNetpRes  # noqa: E402  (the reference, read-only)

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
FULL_GRAD_KEYS_MAX = 6  # store these many full parameter gradients per case (+ norms/sums of all)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def replay_dropout_masks(seed, shapes_p):
    """The Dropout2d noise the reference draws after torch.manual_seed(seed): ATen feature_dropout does
    input.new_empty([B,C,1,1]).bernoulli_(1-p).div_(1-p) once per active Dropout2d, in call order."""
    torch.manual_seed(seed)
    return [torch.empty(B, C, 1, 1).bernoulli_(1 - p).div_(1 - p).view(B, C) for (B, C, p) in shapes_p]


def res_dropout_shapes(B, neurons, p, depth=4):
    shapes = [(B, neurons * 2 ** (k - 1), p / 2 if k == 1 else p) for k in range(1, depth + 1)]
    shapes += [(B, neurons * 2 ** k, p) for k in range(depth, 0, -1)]
    return [(b, c, pp) for (b, c, pp) in shapes if pp > 0]


_WEIGHT_FILES = {}


def save_weights(wname, sd0):
    """Weights are stored once per (architecture, seed) and shared by the cases that use them."""
    if wname in _WEIGHT_FILES:
        for k, v in sd0.items():
            assert torch.equal(v, _WEIGHT_FILES[wname][k]), (wname, k)
        return
    _WEIGHT_FILES[wname] = {k: v.clone() for k, v in sd0.items()}
    np.savez_compressed(os.path.join(OUT, wname + '.npz'), **{k: v.numpy() for k, v in sd0.items()})


def run_case(name, wname, kind, ctor, ctor_kw, body_kw, n_in, seed_w, seed_x, train_mode=True, dropout_seed=None):
    torch.manual_seed(seed_w)
    net = quiet(ctor, 1, 1, torch.device('cpu'), **ctor_kw)
    net.train(train_mode)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    save_weights(wname, sd0)
    nbf = net.nbf
    g = torch.Generator().manual_seed(seed_x)
    x = torch.rand(1, 1, n_in, n_in, generator=g)
    hebb = 0.05 * torch.randn(nbf, nbf, generator=g)
    target = (torch.rand(nbf * nbf, generator=g) > 0.6).float()
    R = torch.randn(nbf, nbf, generator=g)

    # ---- reference
    xr = x.clone().requires_grad_(True)
    hr = hebb.clone().requires_grad_(True)
    if dropout_seed is not None:
        torch.manual_seed(dropout_seed)
    out_r, hebb_r = net(xr, hr)
    loss_r = torch.nn.BCELoss()(out_r.view(-1), target) + (hebb_r * R).sum()
    loss_r.backward()
    grads_r = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    sd_after = {k: v.detach().clone() for k, v in net.state_dict().items()}  # BN running stats after the step

    # ---- oracle on the same state_dict (must be bit-exact)
    masks = None
    kw = dict(body_kw)
    mask_list = []
    if kind == 'unetpres':
        kw['training'] = train_mode
        if train_mode and ctor_kw.get('dropout_ratio', 0.5) > 0:
            mask_list = replay_dropout_masks(dropout_seed, res_dropout_shapes(1, ctor_kw.get('neurons', 16), ctor_kw.get('dropout_ratio', 0.5)))
            masks = [m.clone() for m in mask_list]
            kw['masks'] = masks
    elif kind == 'unetp':
        kw['training'] = train_mode
    sdo = orc.leaf_state(sd0)
    xo = x.clone().requires_grad_(True)
    ho = hebb.clone().requires_grad_(True)
    activ_o, out_o, hebb_o = orc.forward(kind, sdo, xo, ho, rule=ctor_kw.get('rule', 'hebb'), **kw)
    loss_o = orc.bce_mean(out_o.view(-1), target) + (hebb_o * R).sum()
    loss_o.backward()

    def same(a, b, what):
        assert a.shape == b.shape, (name, what, a.shape, b.shape)
        assert torch.equal(a, b), "%s: oracle != reference for %s (max abs diff %g)" % (name, what, float((a - b).abs().max()))

    same(out_o, out_r, 'activout')
    same(hebb_o, hebb_r, 'hebb_new')
    same(xo.grad, xr.grad, 'x.grad')
    same(ho.grad, hr.grad, 'hebb.grad')
    for k, gr in grads_r.items():
        same(sdo[k].grad, gr, 'grad ' + k)
    for k in sd_after:
        if k.endswith('running_mean') or k.endswith('running_var'):
            same(sdo[k], sd_after[k], 'buffer ' + k)

    # ---- store
    keys = list(grads_r.keys())
    conv_keys = [k for k in keys if k.endswith('weight')]
    full = ['w', 'alpha', 'eta'] + [conv_keys[i] for i in sorted(set([0, len(conv_keys) // 2, len(conv_keys) - 1]))]
    blob = {
        'meta_kind': np.array(kind), 'meta_rule': np.array(ctor_kw.get('rule', 'hebb')), 'meta_train': np.array(int(train_mode)),
        'x': x.numpy(), 'hebb': hebb.numpy(), 'target': target.numpy(), 'R': R.numpy(),
        'activ': activ_o.detach().numpy(), 'activout': out_r.detach().numpy(), 'hebb_new': hebb_r.detach().numpy(),
        'loss': np.array(float(loss_r)), 'grad_x': xr.grad.numpy(), 'grad_hebb': hr.grad.numpy(),
        'grad_keys': np.array(keys),
        'grad_l2': np.array([float(grads_r[k].double().norm()) for k in keys]),
        'grad_sum': np.array([float(grads_r[k].double().sum()) for k in keys]),
    }
    for k in full[:FULL_GRAD_KEYS_MAX]:
        blob['grad::' + k] = grads_r[k].numpy()
    blob['weights_file'] = np.array(wname)
    blob['ctor_kw'] = np.array(repr(ctor_kw))
    for k, v in sd_after.items():
        if k.endswith('running_mean') or k.endswith('running_var'):
            blob['sd_after::' + k] = v.numpy()
    for i, m in enumerate(mask_list):
        blob['mask::%02d' % i] = m.numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print('%-28s ok  loss=%.6f  min|logit|=%.3e  files: %s.npz' % (name, float(loss_r), float(activ_o.abs().min()), name))


def run_train_case(name, wname, ctor, ctor_kw, kind, n_in, steps, lr, seed_w):
    """train.py:91-112 on the reference vs oracle.train_steps — losses, final trace and weights bit-exact."""
    torch.manual_seed(seed_w)
    net = quiet(ctor, 1, 1, torch.device('cpu'), **ctor_kw)
    net.train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    save_weights(wname, sd0)
    g = torch.Generator().manual_seed(11)
    imgs = torch.rand(steps, 1, n_in, n_in, generator=g)
    masks = (torch.rand(steps, net.nbf, net.nbf, generator=g) > 0.7).float()
    opt = torch.optim.Adam(net.parameters(), lr=1.0 * lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, gamma=0.5, step_size=2)
    crit = torch.nn.BCELoss()
    hebb = net.initialZeroHebb()
    losses = []
    from torch.autograd import Variable
    for img, mask in zip(imgs, masks):
        opt.zero_grad()
        y_pred, hebb = net(Variable(img[None], requires_grad=False), Variable(hebb, requires_grad=False))
        loss = crit(y_pred.view(-1), Variable(mask.view(-1), requires_grad=False))
        losses.append(loss.item())
        loss.backward()
        opt.step()
        sched.step()
    sdo = orc.leaf_state(sd0)
    losses_o, hebb_o = orc.train_steps(kind, sdo, imgs, masks, ctor_kw.get('rule', 'hebb'), lr=lr, gamma=0.5, steplr=2,
                                       **({'dropout_ratio': 0.0} if kind == 'unetpres' else {}))
    assert losses == losses_o, (losses, losses_o)
    assert torch.equal(hebb.detach(), hebb_o)
    for k, v in net.state_dict().items():
        assert torch.equal(v, sdo[k].detach()), k
    blob = {'imgs': imgs.numpy(), 'masks': masks.numpy(), 'losses': np.array(losses), 'hebb_final': hebb.detach().numpy(),
            'lr': np.array(lr), 'meta_rule': np.array(ctor_kw.get('rule', 'hebb')), 'meta_kind': np.array(kind),
            'final_keys': np.array(list(net.state_dict().keys())),
            'final_l2': np.array([float(v.double().norm()) for v in net.state_dict().values()])}
    blob['weights_file'] = np.array(wname)
    blob['ctor_kw'] = np.array(repr(ctor_kw))
    blob['final::w'] = net.w.detach().numpy()
    blob['final::alpha'] = net.alpha.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print('%-28s ok  losses=%s' % (name, ['%.5f' % l for l in losses]))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # bit-reproducible reductions
    run_case('unetp_hebb_n32', 'w_unetp_n32_s0', 'unetp', UNetp, dict(rule='hebb', nbf=32), {}, 32, 0, 1)
    run_case('unetp_oja_n32', 'w_unetp_n32_s0', 'unetp', UNetp, dict(rule='oja', nbf=32, alfa_type='yoked'), {}, 32, 0, 2)
    run_case('unetp_crop_n32_in37', 'w_unetp_n32_s0', 'unetp', UNetp, dict(rule='hebb', nbf=32), {}, 37, 0, 6)  # 37 -> 32 (S12): crops the skips
    run_case('unetp_bn_bilinear_n32', 'w_unetp_bn_bil_n32_s4', 'unetp', UNetp,
             dict(rule='oja', nbf=32, batch_norm=True, bilinear_upsample=True), dict(batch_norm=True, bilinear=True), 32, 4, 5)
    run_case('unetpres_hebb_n21', 'w_unetpres4_n21_s7', 'unetpres', UNetpRes, dict(neurons=4, dropout_ratio=0.0, rule='hebb', nbf=21),
             dict(dropout_ratio=0.0), 21, 7, 8)
    run_case('unetpres_oja_n21_dropout', 'w_unetpres4_n21_s7', 'unetpres', UNetpRes, dict(neurons=4, dropout_ratio=0.5, rule='oja', nbf=21),
             dict(dropout_ratio=0.5), 21, 7, 9, dropout_seed=123)
    run_case('unetpres_bn_n21_eval', 'w_unetpres2_bn_n21_s10', 'unetpres', UNetpRes,
             dict(neurons=2, dropout_ratio=0.5, rule='oja', nbf=21, batch_norm=True), dict(dropout_ratio=0.5, batch_norm=True),
             21, 10, 11, train_mode=False)
    run_case('unetpres_bn_n32_train', 'w_unetpres2_bn_n32_s12', 'unetpres', UNetpRes,
             dict(neurons=2, dropout_ratio=0.0, rule='hebb', nbf=32, batch_norm=True), dict(dropout_ratio=0.0, batch_norm=True),
             32, 12, 13)
    run_train_case('train_unetp_hebb_n32', 'w_unetp_n32_s0', UNetp, dict(rule='hebb', nbf=32), 'unetp', 32, 4, 1e-3, 0)
    run_train_case('train_unetpres_oja_n21', 'w_unetpres4_n21_s7', UNetpRes, dict(neurons=4, dropout_ratio=0.0, rule='oja', nbf=21),
                   'unetpres', 21, 4, 1e-3, 7)


if __name__ == '__main__':
    main()
