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


def weight_fingerprint(sd0):
    """Per-key float64 (sum, L2) of a state_dict: lets a test that re-creates seeded weights prove they are the reference's."""
    keys = list(sd0.keys())
    return {'w_keys': np.array(keys), 'w_sum': np.array([float(sd0[k].double().sum()) for k in keys]),
            'w_l2': np.array([float(sd0[k].double().norm()) for k in keys])}


def run_case(name, wname, kind, ctor, ctor_kw, body_kw, n_in, seed_w, seed_x, train_mode=True, dropout_seed=None,
             weights_by_seed=False):
    """weights_by_seed: do not store the (multi-MB) initial weights; store the seed + a per-key fingerprint instead.  The
    tests re-create them with the drop-in constructor (same RNG consumption as the reference, tests/test_boundary.py)."""
    torch.manual_seed(seed_w)
    net = quiet(ctor, 1, 1, torch.device('cpu'), **ctor_kw)
    net.train(train_mode)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    if not weights_by_seed:
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
    blob['weights_file'] = np.array('' if weights_by_seed else wname)
    if weights_by_seed:
        blob['weights_seed'] = np.array(seed_w)
        blob.update(weight_fingerprint(sd0))
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


def run_margin_case(name, ctor, ctor_kw, kind, n_in, steps=300, lr=3e-3):
    """A weight state with decision margin: the reference trained for `steps` steps (train.py:91-112) on separable
    synthetic data (mask = a smooth function of the image), so that the logits of a held-out image are far from
    every threshold of eval.py:48-50 — the mask tests then hold on EVERY pixel (no exempt pixels).  The data seed is
    searched until min_pixel,threshold |logit - logit(thr)| > MARGIN."""
    from torch.autograd import Variable
    MARGIN = 2e-4
    thr = np.concatenate([[0.5], np.linspace(0.3, 0.7, 31)])
    thr_logit = torch.from_numpy(np.log(thr / (1 - thr))).float()
    for seed in range(100, 140):
        torch.manual_seed(30)
        net = quiet(ctor, 1, 1, torch.device('cpu'), **ctor_kw)
        net.train()
        g = torch.Generator().manual_seed(seed)
        nbf = net.nbf

        def sample():
            # a bright disc on a dim noisy background; the mask is the disc
            cy, cx = (torch.rand(2, generator=g) * (n_in - 12) + 6).tolist()
            r = float(torch.rand(1, generator=g)) * 5 + 7
            yy, xx = torch.meshgrid(torch.arange(n_in).float(), torch.arange(n_in).float(), indexing='ij')
            m = (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).float()
            img = 0.15 * torch.rand(n_in, n_in, generator=g) + 0.8 * m
            return img[None], m

        opt = torch.optim.Adam(net.parameters(), lr=lr)
        crit = torch.nn.BCELoss()
        hebb = net.initialZeroHebb()
        for _ in range(steps):
            img, m = sample()
            opt.zero_grad()
            y_pred, hebb = net(Variable(img[None], requires_grad=False), Variable(hebb, requires_grad=False))
            loss = crit(y_pred.view(-1), m.view(-1))
            loss.backward()
            opt.step()
        net.eval()
        img, m = sample()
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        with torch.no_grad():
            out_r, hebb_r = net(img[None], hebb.detach())
            activ_o, out_o, hebb_o = orc.forward(kind, sd, img[None], hebb.detach(), rule=ctor_kw.get('rule', 'hebb'), training=False)
        assert torch.equal(out_o, out_r) and torch.equal(hebb_o, hebb_r)
        margin = float((activ_o.view(-1, 1) - thr_logit.view(1, -1)).abs().min())
        if margin > MARGIN:
            break
    else:
        raise SystemExit('no seed with margin > %g found' % MARGIN)
    wname = 'w_' + name
    save_weights(wname, sd)
    blob = {'meta_kind': np.array(kind), 'meta_rule': np.array(ctor_kw.get('rule', 'hebb')), 'meta_train': np.array(0),
            'x': img[None].numpy(), 'hebb': hebb.detach().numpy(), 'target': m.numpy(), 'activ': activ_o.numpy(),
            'activout': out_r.numpy(), 'hebb_new': hebb_r.numpy(), 'margin': np.array(margin), 'final_loss': np.array(float(loss)),
            'weights_file': np.array(wname), 'ctor_kw': np.array(repr(ctor_kw))}
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print('%-28s ok  seed=%d  train loss=%.4f  min margin to any of 32 thresholds=%.3e  logit range [%.2f, %.2f]  mask px %d / IoU@0.5 %.3f'
          % (name, seed, float(loss), margin, float(activ_o.min()), float(activ_o.max()), int((out_r > 0.5).sum()),
             float(((out_r > 0.5) & (m > 0)).sum()) / max(1.0, float(((out_r > 0.5) | (m > 0)).sum()))))


def run_infer_tail_golden():
    """Inference tail: run the REFERENCE's own eval.score_model_best_iou (through a stand-in net that
    replays fixed predictions), eval.eval_net's fast_iou_metric, infer.py's threshold and utils.rle_encode.encode; assert
    that oracle/infer_tail_oracle.py returns identical values; store inputs + outputs."""
    import infer_tail_oracle as ito
    import ref_loader
    ref = ref_loader.load('/root/reference/src')
    g = torch.Generator().manual_seed(77)

    def blobs(n, R, C, smooth=5):
        z = torch.randn(n, 1, R, C, generator=g)
        k = torch.ones(1, 1, smooth, smooth) / smooth ** 2
        return torch.nn.functional.conv2d(z, k, padding=smooth // 2)[:, 0]

    class ReplayNet:  # just enough of the nn.Module surface for eval.py:33-45
        def __init__(self, preds):
            self.preds, self.i = preds, 0

        def eval(self):
            return self

        def initialZeroHebb(self):
            return torch.zeros(1)

        def __call__(self, x, hebb):
            y = self.preds[self.i]
            self.i += 1
            return y, hebb

    blob = {}
    for tag, B, R, C in (('a', 10, 37, 37), ('b', 5, 101, 101)):
        base = blobs(B, R, C)
        labels = (base > 0.1).float()
        logits = 9.0 * (base - 0.1) + 0.6 * blobs(B, R, C, 3)  # correlated with the labels: IoU varies over the sweep
        preds = torch.sigmoid(logits)
        # edge cases: empty / full predictions, empty label, last column-major pixel set, checkerboard
        preds[0].fill_(0.01)
        preds[1].fill_(0.99)
        labels[2].zero_()
        if B > 5:
            preds[3].fill_(0.2)
            preds[3][-1, -1] = 0.9
            preds[3][0, 0] = 0.9
            ii, jj = torch.meshgrid(torch.arange(R), torch.arange(C), indexing='ij')
            preds[4] = ((ii + jj) % 2).float() * 0.8 + 0.1
        X = np.zeros((B, 1, R, C), dtype=np.float64)
        y_valid = labels.numpy().astype(np.float64)
        # eval.py:20-64 on the replayed predictions
        thr_best, iou_best = ref.eval_.score_model_best_iou(ReplayNet([p for p in preds]), X, y_valid, torch.device('cpu'))
        preds_np = preds.numpy()
        o_thr, o_iou, o_ious = ito.sweep_best_iou(y_valid, [p for p in preds_np])
        assert o_thr == thr_best and o_iou == iou_best and o_iou.dtype == iou_best.dtype, (o_thr, thr_best, o_iou, iou_best)
        ref_ious = np.array([ref.iou_metric.iou_metric_batch(y_valid, preds_np > t) for t in ito.sweep_thresholds()])
        assert np.array_equal(ref_ious, o_ious) and ref_ious.dtype == o_ious.dtype
        # eval.py:100: fast_iou_metric on the flattened prediction / target of each image
        fast = np.array([ref.iou_metric.fast_iou_metric(y_true_in=y_valid[b].astype(np.float32).reshape(-1), y_pred_in=preds_np[b].reshape(-1))
                         for b in range(B)])
        fast_o = np.array([ito.fast_iou_metric(y_valid[b].astype(np.float32).reshape(-1), preds_np[b].reshape(-1)) for b in range(B)])
        assert np.array_equal(fast, fast_o)
        # infer.py:81,99 at a few mask thresholds
        rles = {}
        for mt in (0.5, 0.35, 0.62):
            enc = [ref.rle_encode.encode(np.round(preds_np[b] > mt)) for b in range(B)]
            enc_o = [ito.rle_encode(np.round(preds_np[b] > mt)) for b in range(B)]
            assert enc == enc_o
            alt = [ref.rle_encode.rle_encode((preds_np[b] > mt).astype(np.uint8)) for b in range(B)]  # the second implementation agrees
            assert alt == enc
            rles[mt] = enc
            blob['%s_rle_%g' % (tag, mt)] = np.array(enc)
            blob['%s_mask_%g' % (tag, mt)] = np.stack([(preds_np[b] > mt).astype(np.uint8) for b in range(B)])
        blob.update({tag + '_preds': preds_np, tag + '_labels': labels.numpy(), tag + '_ious': ref_ious,
                     tag + '_thr_best': np.array(thr_best), tag + '_iou_best': np.array(iou_best), tag + '_fast_iou': fast})
        print('infer_tail[%s]  B=%d %dx%d  best thr %.4f iou %.4f  fast_iou %s  rle lens %s' %
              (tag, B, R, C, thr_best, iou_best, np.round(fast, 3), [len(s) for s in rles[0.5]]))
    np.savez_compressed(os.path.join(OUT, 'infer_tail.npz'), **blob)


def main():
    os.makedirs(OUT, exist_ok=True)
    run_infer_tail_golden()
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
    # BASELINE sizes (configs[1] and configs[3]): UNetp @128 and the script variant UNetpRes(neurons=8) @101
    # (unet_p_res_script.py:30,781-788: default dropout 0.5, hebb); weights re-created from the seed by the tests
    run_case('unetp_oja_n128', '', 'unetp', UNetp, dict(rule='oja', nbf=128), {}, 128, 20, 21, weights_by_seed=True)
    run_case('unetpres8_hebb_n101_dropout', '', 'unetpres', UNetpRes, dict(neurons=8, rule='hebb', nbf=101),
             dict(dropout_ratio=0.5), 101, 22, 23, dropout_seed=321, weights_by_seed=True)
    run_case('unetpres8_oja_n101_eval', '', 'unetpres', UNetpRes, dict(neurons=8, rule='oja', nbf=101),
             dict(dropout_ratio=0.5), 101, 22, 24, train_mode=False, weights_by_seed=True)
    run_margin_case('margin_unetp_oja_n32', UNetp, dict(rule='oja', nbf=32), 'unetp', 32)
    run_train_case('train_unetp_hebb_n32', 'w_unetp_n32_s0', UNetp, dict(rule='hebb', nbf=32), 'unetp', 32, 4, 1e-3, 0)
    run_train_case('train_unetpres_oja_n21', 'w_unetpres4_n21_s7', UNetpRes, dict(neurons=4, dropout_ratio=0.0, rule='oja', nbf=21),
                   'unetpres', 21, 4, 1e-3, 7)


if __name__ == '__main__':
    main()
