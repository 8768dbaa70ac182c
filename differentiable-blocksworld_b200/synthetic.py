"""Synthetic multi-view inputs of the benchmark / tests: cameras on a ring looking at the origin in
the PyTorch3D convention the reference's datasets produce (X_cam = X_world @ R + T, src/dataset/dtu.py:75-124), with
a shared NDC intrinsics matrix laid out as src/dataset/dtu.py:102-106."""
import math

import torch


def ring_cameras(n_views, dist=2.75, elev_deg=25.0, fx=4.8, dtype=torch.float32, jitter=0.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    az = torch.arange(n_views, dtype=torch.float64) / max(n_views, 1) * 2 * math.pi
    el = torch.full((n_views,), elev_deg * math.pi / 180, dtype=torch.float64)
    if jitter:
        el = el + (torch.rand(n_views, generator=g, dtype=torch.float64) - 0.5) * jitter
    C = torch.stack([dist * torch.cos(el) * torch.sin(az), dist * torch.sin(el), dist * torch.cos(el) * torch.cos(az)], 1)
    zc = -C / C.norm(dim=1, keepdim=True)
    up = torch.tensor([0., 1., 0.], dtype=torch.float64)[None].expand(n_views, -1)
    xc = torch.cross(up, zc, dim=1)
    xc = xc / xc.norm(dim=1, keepdim=True)
    yc = torch.cross(zc, xc, dim=1)
    R = torch.stack([xc, yc, zc], dim=2)
    T = -torch.bmm(C[:, None], R)[:, 0]
    K = torch.zeros(4, 4, dtype=torch.float64)
    K[0, 0] = K[1, 1] = fx
    K[2, 3] = 1
    K[3, 2] = 1
    return R.to(dtype), T.to(dtype), K.to(dtype)
