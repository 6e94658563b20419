"""Score-network oracle: functional fp32 restatement of NCSN_LiDAR_small.forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A floating-point kernel keeps a
plain torch fp32 reference; this is it.  It is a functional rewrite driven by a
flat parameter dict (reference state_dict keys), not the reference's module
tree.  Follows /root/reference/LiDARGen/models/ncsnv2.py:484-518 and the block
semantics of models/layers.py and models/normalization.py:163-176.

`taps` (optional dict) collects named intermediates so per-layer CUDA parity
tests can localise a mismatch.
"""
import torch
import torch.nn.functional as F


def circ_conv3x3(x, w, b=None, dilation=1):
    """3x3 conv with circular padding on BOTH H and W (layers.py:37-60; torch's
    padding_mode='circular' wraps rows as well as columns)."""
    d = dilation
    xp = F.pad(x, (d, d, d, d), mode="circular")
    return F.conv2d(xp, w, b, dilation=d)


def zero_conv3x3(x, w, b=None):
    """plain zero-padded 3x3 (begin/end conv ncsnv2.py:433,436; ConvMeanPool layers.py:295)."""
    return F.conv2d(x, w, b, padding=1)


def mean_pool2(x):
    """average of the four stride-2 phases (layers.py:310-312)."""
    return (x[:, :, ::2, ::2] + x[:, :, 1::2, ::2] + x[:, :, ::2, 1::2] + x[:, :, 1::2, 1::2]) / 4.0


def instance_norm_plus(x, alpha, gamma, beta, eps=1e-5):
    """InstanceNorm2dPlus (normalization.py:163-176): per-(n,c) instance norm
    (biased variance) plus a cross-channel standardised mean term."""
    mu = x.mean(dim=(2, 3))                                   # [N, C]
    m = mu.mean(dim=-1, keepdim=True)
    v = mu.var(dim=-1, keepdim=True)                           # unbiased over channels
    mu_n = (mu - m) / torch.sqrt(v + 1e-5)
    h = F.instance_norm(x, eps=eps)
    h = h + (mu_n * alpha)[..., None, None]
    return gamma.view(1, -1, 1, 1) * h + beta.view(1, -1, 1, 1)


def _residual_block(P, pre, x, kind, dilation=None, taps=None):
    """layers.py:401-456.  kind in {'plain','down_pool','dilated'}."""
    n1 = instance_norm_plus(x, P[pre + ".normalize1.alpha"], P[pre + ".normalize1.gamma"], P[pre + ".normalize1.beta"])
    h = F.elu(n1)
    conv = (lambda t, w, b: circ_conv3x3(t, w, b, dilation)) if dilation else (lambda t, w, b: circ_conv3x3(t, w, b, 1))
    h = conv(h, P[pre + ".conv1.weight"], P[pre + ".conv1.bias"])
    if taps is not None:
        taps[pre + ".conv1"] = h
    h = instance_norm_plus(h, P[pre + ".normalize2.alpha"], P[pre + ".normalize2.gamma"], P[pre + ".normalize2.beta"])
    h = F.elu(h)
    if kind == "down_pool":
        h = mean_pool2(zero_conv3x3(h, P[pre + ".conv2.conv.weight"], P[pre + ".conv2.conv.bias"]))
        sc = mean_pool2(F.conv2d(x, P[pre + ".shortcut.conv.weight"], P[pre + ".shortcut.conv.bias"]))
    else:
        h = conv(h, P[pre + ".conv2.weight"], P[pre + ".conv2.bias"])
        if kind == "dilated":
            sc = conv(x, P[pre + ".shortcut.weight"], P[pre + ".shortcut.bias"])
        else:
            sc = x
    out = sc + h
    if taps is not None:
        taps[pre] = out
    return out


def _rcu(P, pre, x, n_blocks):
    """RCUBlock (layers.py:112-134): per block two (ELU -> bias-free conv), then += block input."""
    for b in range(1, n_blocks + 1):
        r = x
        for s in (1, 2):
            x = circ_conv3x3(F.elu(x), P[f"{pre}.{b}_{s}_conv.weight"])
        x = x + r
    return x


def _crp(P, pre, x):
    """CRPBlock (layers.py:62-83): ELU, then 2 x (maxpool5 -> conv -> accumulate)."""
    x = F.elu(x)
    path = x
    for i in (0, 1):
        path = F.max_pool2d(path, kernel_size=5, stride=1, padding=2)
        path = circ_conv3x3(path, P[f"{pre}.convs.{i}.weight"])
        x = path + x
    return x


def _refine(P, pre, xs, out_hw, start=False, end=False, taps=None):
    """RefineBlock (layers.py:214-249) with MSFBlock (layers.py:165-184)."""
    hs = [_rcu(P, f"{pre}.adapt_convs.{i}", x, 2) for i, x in enumerate(xs)]
    if not start:
        acc = None
        for i, h in enumerate(hs):
            h = circ_conv3x3(h, P[f"{pre}.msf.convs.{i}.weight"], P[f"{pre}.msf.convs.{i}.bias"])
            h = F.interpolate(h, size=out_hw, mode="bilinear", align_corners=True)
            acc = h if acc is None else acc + h
        h = acc
    else:
        h = hs[0]
    if taps is not None:
        taps[pre + ".msf"] = h
    h = _crp(P, f"{pre}.crp", h)
    if taps is not None:
        taps[pre + ".crp"] = h
    h = _rcu(P, f"{pre}.output_convs", h, 3 if end else 1)
    if taps is not None:
        taps[pre] = h
    return h


@torch.no_grad()
def score_forward(P, x, y, taps=None):
    """s(x, sigma_y).  P: dict with reference state_dict keys (float32, on x's device);
    x: [B,2,H,W] float32; y: [B] int64 noise-level labels.  ncsnv2.py:484-518."""
    B, _, H, W = x.shape
    h = 2.0 * x - 1.0
    xs = torch.linspace(0, 1, steps=W, device=x.device)
    ys = torch.linspace(0, 1, steps=H, device=x.device)
    grid = torch.stack((xs.view(1, W).expand(H, W), ys.view(H, 1).expand(H, W)), dim=0)
    h = torch.cat((h, grid.unsqueeze(0).expand(B, 2, H, W)), dim=1)
    out = zero_conv3x3(h, P["begin_conv.weight"], P["begin_conv.bias"])
    if taps is not None:
        taps["begin_conv"] = out
    l1 = _residual_block(P, "res1.1", _residual_block(P, "res1.0", out, "plain", taps=taps), "plain", taps=taps)
    l2 = _residual_block(P, "res2.1", _residual_block(P, "res2.0", l1, "down_pool", taps=taps), "plain", taps=taps)
    l3 = _residual_block(P, "res3.1", _residual_block(P, "res3.0", l2, "dilated", 2, taps=taps), "plain", 2, taps=taps)
    l4 = _residual_block(P, "res4.1", _residual_block(P, "res4.0", l3, "dilated", 4, taps=taps), "plain", 4, taps=taps)
    r1 = _refine(P, "refine1", [l4], l4.shape[2:], start=True, taps=taps)
    r2 = _refine(P, "refine2", [l3, r1], l3.shape[2:], taps=taps)
    r3 = _refine(P, "refine3", [l2, r2], l2.shape[2:], taps=taps)
    r4 = _refine(P, "refine4", [l1, r3], l1.shape[2:], end=True, taps=taps)
    o = F.elu(instance_norm_plus(r4, P["normalizer.alpha"], P["normalizer.gamma"], P["normalizer.beta"]))
    o = zero_conv3x3(o, P["end_conv.weight"], P["end_conv.bias"])
    return o / P["sigmas"][y].view(B, 1, 1, 1)


class OracleScoreNet:
    """callable(x, labels) wrapper with the reference module's call signature."""

    def __init__(self, params):
        self.params = params

    def __call__(self, x, y):
        return score_forward(self.params, x, y)
