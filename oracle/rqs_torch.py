"""Op-for-op torch restatement of upstream zuko.transforms.MonotonicRQSTransform.

TEST INFRASTRUCTURE ONLY. PARITY UNPINNED: zuko is an un-vendored, un-pinned third-party
dependency of the reference (Dockerfile:7 `zuko @ git+https://github.com/probabilists/zuko@master`;
README.md:59-61 points at the valsdav/zuko fork); it is absent from /root/reference and from this
image. This file restates the published algorithm of zuko >= 1.0 (`zuko/transforms.py`) as SURVEY
App. B records it, using exactly the torch ops zuko uses (softmax -> pad -> cumsum -> exp ->
searchsorted -> gather) so that a torch user sees the same rounding. Reference call sites:
memflow/unfolding_flow/unfolding_flow.py:88-97, memflow/unfolding_flow/custom_spline_flow_ps.py:16-33.
"""
import math

import torch
import torch.nn.functional as F


class MonotonicRQSTransform:
    def __init__(self, widths, heights, derivatives, bound=5.0, slope=1e-3):
        if slope is not None:
            widths = widths / (1 + abs(2 * widths / math.log(slope)))
            heights = heights / (1 + abs(2 * heights / math.log(slope)))
            derivatives = derivatives / (1 + abs(derivatives / math.log(slope)))
        widths = F.pad(F.softmax(widths, dim=-1), (1, 0), value=0)
        heights = F.pad(F.softmax(heights, dim=-1), (1, 0), value=0)
        derivatives = F.pad(derivatives, (1, 1), value=0)
        self.horizontal = bound * (2 * torch.cumsum(widths, dim=-1) - 1)
        self.vertical = bound * (2 * torch.cumsum(heights, dim=-1) - 1)
        self.derivatives = torch.exp(derivatives)

    @property
    def bins(self):
        return self.horizontal.shape[-1] - 1

    def bin(self, k):
        mask = torch.logical_and(0 <= k, k < self.bins)
        k = k % self.bins
        k0_k1 = torch.stack((k, k + 1), dim=-1)
        k0_k1, hs, vs, ds = torch.broadcast_tensors(
            k0_k1, self.horizontal[..., :2], self.vertical[..., :2], self.derivatives[..., :2])
        x0, x1 = self.horizontal.expand(*hs.shape[:-1], -1).gather(-1, k0_k1).unbind(dim=-1)
        y0, y1 = self.vertical.expand(*vs.shape[:-1], -1).gather(-1, k0_k1).unbind(dim=-1)
        d0, d1 = self.derivatives.expand(*ds.shape[:-1], -1).gather(-1, k0_k1).unbind(dim=-1)
        s = (y1 - y0) / (x1 - x0)
        return mask, x0, x1, y0, y1, d0, d1, s

    @staticmethod
    def searchsorted(seq, value):
        seq, value = torch.broadcast_tensors(seq, value[..., None])
        seq = seq.contiguous()
        return torch.searchsorted(seq, value[..., :1].contiguous()).squeeze(dim=-1)

    def _call(self, x):
        return self.call_and_ladj(x)[0]

    def _inverse(self, y):
        k = self.searchsorted(self.vertical, y) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self.bin(k)
        y_ = mask * (y - y0)
        a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
        b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
        c = -s * y_
        z = 2 * c / (-b - (b**2 - 4 * a * c).sqrt())
        x = x0 + z * (x1 - x0)
        return torch.where(mask, x, y)

    def log_abs_det_jacobian(self, x, y=None):
        return self.call_and_ladj(x)[1]

    def call_and_ladj(self, x):
        k = self.searchsorted(self.horizontal, x) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self.bin(k)
        z = mask * (x - x0) / (x1 - x0)
        z_1mz = z * (1 - z)
        den = s + (d0 + d1 - 2 * s) * z_1mz
        y = y0 + (y1 - y0) * (s * z**2 + d0 * z_1mz) / den
        jacobian = s**2 * (2 * s * z_1mz + d0 * (1 - z) ** 2 + d1 * z**2) / den**2
        return torch.where(mask, y, x), mask * jacobian.log()
