"""ADA augmentation (`trainers/non_leaking.py`) on the libb200gan kernels.

Same call surface as the reference module:

    augment(img, p, transform_matrix=(None, None)) -> (img, (G, C))                 non_leaking.py:388-392
    random_apply_affine(img, p, G=None, antialiasing_kernel=SYM6) -> (img, G)       :314-371
    random_apply_color(img, p, C=None) -> (img, C)                                  :380-386
    sample_affine(p, size, height, width) -> G (size, 3, 3)                         :152-215
    sample_color(p, size) -> C (size, 4, 4)                                         :218-252
    AdaptiveP: the `ada_aug_p` controller of the trainer                            generator_trainer.py:333-337, 669-687

What runs where.  The transforms are drawn on the HOST exactly as in the reference (tiny fp32 matrices; the random
numbers are requested from torch's CPU generator in the reference's order -- parameter, then the Bernoulli gate, stage
by stage -- so one seed gives both implementations the same G and C).  The image work is three kernels instead of the
reference's ~20 passes: the 12x12-tap interpolating FIR (`upfirdn2d(up=2)`), ONE kernel for [sampling grid, bilinear
`grid_sample`, colour matrix] (`ops.affine_color`: the grid is an affine function of the output pixel and is evaluated in
the kernel, never stored; the colour transform is per-pixel linear and therefore commutes with the FIR that follows,
its offset divided by that filter's DC gain) and the decimating FIR (`upfirdn2d(down=2)`).  Reflection padding and the
final crop stay torch views / one small copy.
"""
import math

import torch
from torch.nn import functional as F

from . import ops

# sym6 wavelet low-pass taps (the anti-aliasing filter of the geometric transforms), non_leaking.py:9-22
SYM6 = (0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148)


# ---------------------------------------------------------------------------------------------
# random draws, in the reference's order and with the reference's torch calls (non_leaking.py:120-141)
# ---------------------------------------------------------------------------------------------
def _choice(n, values):
    return torch.tensor(values)[torch.randint(high=len(values), size=(n,))]


def _uniform(n, lo, hi):
    return torch.empty(n).uniform_(lo, hi)


def _lognormal(n, std):
    return torch.empty(n).log_normal_(mean=0, std=std)


def _normal(n, std):
    return torch.empty(n).normal_(0, std)


def _gate(n, p):
    return torch.empty(n).bernoulli_(p).view(n, 1, 1)


def _eye(n, d):
    return torch.eye(d).unsqueeze(0).repeat(n, 1, 1)


def _lin2(a, b, c, d):
    """(n,3,3) homogeneous matrices with the 2x2 block [[a, b], [c, d]]"""
    m = _eye(a.shape[0], 3)
    m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1] = a, b, c, d
    return m


def _rot2(theta):
    return _lin2(torch.cos(theta), -torch.sin(theta), torch.sin(theta), torch.cos(theta))


def _shift2(tx, ty):
    m = _eye(tx.shape[0], 3)
    m[:, 0, 2], m[:, 1, 2] = tx, ty
    return m


def _chain(stages, n, dim):
    """stages: [(draw parameter -> matrices (n,dim,dim), probability)]; each is applied with its probability on top of
    what is there: M <- (gate * T + (1 - gate) * I) @ M  (non_leaking.py:144-149)"""
    eye = _eye(n, dim)
    m = eye
    for make, prob in stages:
        t = make()
        gate = _gate(n, prob)
        m = (gate * t + (1 - gate) * eye) @ m
    return m


def sample_affine(p, size, height, width):
    """(size,3,3) geometric transforms in normalised [-1,1] coordinates: x-flip, 90-degree rotation, integer translation,
    isotropic scale, rotation, anisotropic scale, rotation, fractional translation (non_leaking.py:152-215)."""
    one = torch.ones(size)
    p_rot = 1 - math.sqrt(1 - p)

    def int_shift():
        u = _uniform(size, -0.125, 0.125)
        return _shift2(torch.round(u * width) / width, torch.round(u * height) / height)

    def iso():
        s = _lognormal(size, 0.2 * math.log(2))
        return _lin2(s, 0 * one, 0 * one, s)

    def aniso():
        s = _lognormal(size, 0.2 * math.log(2))
        return _lin2(s, 0 * one, 0 * one, 1 / s)

    def frac_shift():
        t = _normal(size, 0.125)
        return _shift2(t, t)

    return _chain([
        (lambda: _lin2(1 - 2.0 * _choice(size, (0, 1)), 0 * one, 0 * one, one), p),
        (lambda: _rot2(-math.pi / 2 * _choice(size, (0, 3))), p),
        (int_shift, p),
        (iso, p),
        (lambda: _rot2(-_uniform(size, -math.pi, math.pi)), p_rot),
        (aniso, p),
        (lambda: _rot2(-_uniform(size, -math.pi, math.pi)), p_rot),
        (frac_shift, p),
    ], size, 3)


def sample_color(p, size):
    """(size,4,4) colour transforms on homogeneous RGB: brightness, contrast, luma flip, hue rotation, saturation
    (non_leaking.py:218-252); v = (1,1,1)/sqrt(3) is the luma axis."""
    u = 1 / math.sqrt(3)
    v = torch.tensor([u, u, u, 0.0])
    vv = torch.outer(v, v)
    eye = torch.eye(4)

    def brightness():
        b = _normal(size, 0.2)
        m = _eye(size, 4)
        m[:, :3, 3] = b.view(-1, 1)
        return m

    def contrast():
        c = _lognormal(size, 0.5 * math.log(2))
        m = _eye(size, 4)
        m[:, 0, 0] = m[:, 1, 1] = m[:, 2, 2] = c
        return m

    def luma_flip():
        i = _choice(size, (0, 1))
        return eye - 2 * vv * i.view(-1, 1, 1)

    def hue():
        theta = _uniform(size, -math.pi, math.pi)
        cross = torch.tensor([(0, -u, u), (u, 0, -u), (-u, u, 0)])
        s, c = torch.sin(theta).view(-1, 1, 1), torch.cos(theta).view(-1, 1, 1)
        m = _eye(size, 4)
        m[:, :3, :3] = c * torch.eye(3) + s * cross + (1 - c) * vv[:3, :3]
        return m

    def saturation():
        i = _lognormal(size, 1 * math.log(2))
        return vv + (eye - vv) * i.view(-1, 1, 1)

    return _chain([(brightness, p), (contrast, p), (luma_flip, p), (hue, p), (saturation, p)], size, 4)


# ---------------------------------------------------------------------------------------------
# geometry of the warp (non_leaking.py:255-312, 323-357)
# ---------------------------------------------------------------------------------------------
def _padding_for(g_inv, height, width):
    """how far the inverse-transformed image corners leave the frame, in pixels, maximised over the batch:
    (x_low, x_high, y_low, y_high)"""
    corners = torch.tensor([(-1.0, -1, 1), (-1, 1, 1), (1, -1, 1), (1, 1, 1)]).t()
    ext = g_inv[:, :2, :] @ corners                                        # (n, 2, 4)
    size = torch.tensor((width, height))
    low = ((ext.min(-1).values + 1) * size).clamp(max=0).abs().ceil().max(0).values.to(torch.int64).tolist()
    high = (ext.max(-1).values * size - size).clamp(min=0).ceil().max(0).values.to(torch.int64).tolist()
    return low[0], high[0], low[1], high[1]


def _sample_and_pad(img, p, pad_k, G):
    batch, _, height, width = img.shape
    while True:
        g = sample_affine(p, batch, height, width) if G is None else G
        px1, px2, py1, py2 = _padding_for(torch.inverse(g), height, width)
        pads = (px1 + pad_k, px2 + pad_k, py1 + pad_k, py2 + pad_k)
        if max(pads[0], pads[1]) < width and max(pads[2], pads[3]) < height:           # what reflection padding accepts
            return F.pad(img, pads, mode='reflect'), g, (px1, px2, py1, py2)
        if G is not None:
            raise ValueError('the given transform moves the image further than reflection padding can cover')


def _source_pixel_map(g, w_o, h_o, w_p, h_p, w2, h2, px1, py1):
    """(n,6) float64: output pixel (ox, oy) of the warped 2x image -> source pixel of the 2x image.  The reference builds
    this as a grid tensor: linspace over the padded frame (make_grid, :341-348), inverse transform (affine_grid, :349),
    rescale to the padded frame (:350-354), and grid_sample's own un-normalisation with align_corners=False."""
    n = g.shape[0]
    x0, x1 = -2 * px1 / w_o - 1, 2 * (w_p - px1) / w_o - 1
    y0, y1 = -2 * py1 / h_o - 1, 2 * (h_p - py1) / h_o - 1
    to_grid = torch.tensor([[(x1 - x0) / (w2 - 1), 0, x0], [0, (y1 - y0) / (h2 - 1), y0], [0, 0, 1]], dtype=torch.float64)
    rescale = torch.tensor([[w_o / w_p, 0, (w_o + 2 * px1) / w_p - 1], [0, h_o / h_p, (h_o + 2 * py1) / h_p - 1], [0, 0, 1]],
                           dtype=torch.float64)
    to_pixel = torch.tensor([[w2 / 2, 0, (w2 - 1) / 2], [0, h2 / 2, (h2 - 1) / 2], [0, 0, 1]], dtype=torch.float64)
    inv = torch.inverse(g).double()                                        # fp32 inverse like the reference, then exact
    inv[:, 2, :] = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    m = to_pixel @ rescale @ inv @ to_grid
    return m[:, :2, :].reshape(n, 6).contiguous()


def _color_rows(C, dc_gain, channels):
    """(n,4,4) homogeneous colour matrices -> (n,3,4) rows (matrix | offset / dc_gain) for the kernel"""
    assert channels == 3, 'the colour transform is defined on RGB'
    rows = C[:, :3, :].clone().float()
    rows[:, :, 3] /= dc_gain
    return rows.contiguous()


def random_apply_affine(img, p, G=None, antialiasing_kernel=SYM6, C=None):
    """Geometric augmentation of a batch (non_leaking.py:314-371).  C: optional (n,4,4) colour matrices applied in the same
    kernel as the warp (see the module docstring).  Returns (img, G)."""
    taps = torch.as_tensor(antialiasing_kernel, dtype=torch.float32)
    len_k = taps.numel()
    pad_k = (len_k + 1) // 2
    k2 = torch.outer(taps, taps).to(img.device)
    h_o, w_o = img.shape[2], img.shape[3]
    img_pad, G, (px1, px2, py1, py2) = _sample_and_pad(img, p, pad_k, G)
    w_p, h_p = img_pad.shape[3] - len_k + 1, img_pad.shape[2] - len_k + 1
    img_2x = ops.upfirdn2d(img_pad, k2.flip(0, 1), up=2)
    h2, w2 = img_2x.shape[2], img_2x.shape[3]
    mat = _source_pixel_map(G, w_o, h_o, w_p, h_p, w2, h2, px1, py1).to(img.device)
    color = None if C is None else _color_rows(C, float(k2.sum()), img.shape[1]).to(img.device)
    warped = ops.affine_color(img_2x, mat, color, (h2, w2))
    down = ops.upfirdn2d(warped, k2, down=2)
    end_y = down.shape[2] if py2 + 1 == 0 else -py2 - 1
    end_x = down.shape[3] if px2 + 1 == 0 else -px2 - 1
    return down[:, :, py1:end_y, px1:end_x], G


def random_apply_color(img, p, C=None):
    """Colour augmentation alone (non_leaking.py:373-386) through the same kernel with an identity warp."""
    if C is None:
        C = sample_color(p, img.shape[0])
    n, c, h, w = img.shape
    ident = torch.tensor([1.0, 0, 0, 0, 1.0, 0], dtype=torch.float64).repeat(n, 1).to(img.device)
    return ops.affine_color(img, ident, _color_rows(C, 1.0, c).to(img.device), (h, w)), C


def augment(img, p, transform_matrix=(None, None)):
    """`non_leaking.augment` (:388-392): geometric then colour transform with probability-p stages; returns
    (img, (G, C)) so that a second batch can be given the same transforms."""
    G, C = transform_matrix
    if G is None:
        # the reference draws G (inside random_apply_affine, possibly several times) before C: keep the order of the draws
        G = _first_admissible_affine(img, p, (len(SYM6) + 1) // 2)
    if C is None:
        C = sample_color(p, img.shape[0])
    out, G = random_apply_affine(img, p, G, C=C)
    return out, (G, C)


def _first_admissible_affine(img, p, pad_k):
    """the rejection loop of the reference (:285-310) without touching the image: draw G until reflection padding covers it"""
    batch, _, height, width = img.shape
    while True:
        g = sample_affine(p, batch, height, width)
        px1, px2, py1, py2 = _padding_for(torch.inverse(g), height, width)
        if max(px1, px2) + pad_k < width and max(py1, py2) + pad_k < height:
            return g


class AdaptiveP:
    """The trainer's `ada_aug_p` controller (generator_trainer.py:333-337, 669-687): r_t = E[sign(D(real))] over at
    least 256 predictions; when the configured p is 0 the probability moves by ada_target / ada_length per image towards
    r_t == ada_target."""

    def __init__(self, p=0.0, ada_target=0.6, ada_length=500000):
        self.adaptive = not p > 0
        self.p = p if p > 0 else 0.0
        self.target, self.step = ada_target, ada_target / ada_length
        self.signs, self.count, self.r_t = 0.0, 0, 0.0

    def update(self, real_pred):
        self.signs += float(torch.sign(real_pred.detach()).sum())
        self.count += real_pred.shape[0]
        if self.count > 255:
            self.r_t = self.signs / self.count
            if self.adaptive:
                self.p = min(1.0, max(0.0, self.p + (1 if self.r_t > self.target else -1) * self.step * self.count))
            self.signs, self.count = 0.0, 0
        return self.p
