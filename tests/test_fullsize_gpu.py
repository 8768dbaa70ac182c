"""GPU: the committed golden fixture, and -- at BASELINE.json's full size (49 views 400x400, 10 blocks, K=10), where the
oracle would take minutes -- size-independent properties of the render path."""
import os

import numpy as np
import pytest
import torch

from oracle import dbw_path as D
from tests.helpers import scene_to_device, render_product, split_map_grads

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_cuda_matches_committed_golden_fixture():
    from tests.golden.make_golden import small_case
    g = np.load(os.path.join(GOLD, 'render_small.npz'))
    dev = torch.device('cuda:0')
    tpl, p, R, T, K, imgs = small_case()
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF)
    fg = render_product(scene_to_device(blocks, dev), R.to(dev), T.to(dev), K, (40, 48), 1e-4, 10, z_clip=0.001,
                        detach_bary=True, faces_alpha=fa.to(dev))
    env = render_product(scene_to_device(tpl.build_env(p), dev), R.to(dev), T.to(dev), K, (40, 48), 0.0, 1, z_clip=0.001)
    for out, key in ((fg, 'fg'), (env, 'env')):
        err = (out.cpu() - torch.from_numpy(g[key])).abs()
        assert (err > 1e-4).float().mean().item() <= 1e-4 and err.max().item() < 2e-2, (key, err.max().item())


@pytest.fixture(scope='module')
def full():
    dev = torch.device('cuda:0')
    tpl = D.SceneTemplate(n_blocks=10, txt_size=256)
    p = D.init_params(10, 256, seed=227391, boxy=True)
    R, T, K = D.ring_cameras(49)
    blocks, alpha = tpl.build_blocks(p, decimate=8)
    env = tpl.build_env(p)
    return dict(dev=dev, tpl=tpl, blocks=blocks, env=env, fa=alpha.repeat_interleave(tpl.BNF), R=R.to(dev), T=T.to(dev), K=K)


def _render_blocks(full, sl=slice(None), maps=None, requires_grad=False):
    sc = scene_to_device(full['blocks'], full['dev'], requires_grad=requires_grad)
    if maps is not None:
        sc['maps'] = maps
    fa = full['fa'].to(full['dev'])
    if requires_grad:
        fa.requires_grad_(True)
    out = render_product(sc, full['R'][sl], full['T'][sl], full['K'], (400, 400), 1e-4, 10, z_clip=0.001, detach_bary=True, faces_alpha=fa)
    return out, sc, fa


def test_fullsize_ranges_and_determinism(full):
    a, _, _ = _render_blocks(full)
    b, _, _ = _render_blocks(full)
    assert a.shape == (49, 4, 400, 400) and torch.isfinite(a).all()
    assert torch.equal(a, b)                                           # forward has no atomics: bit-reproducible
    assert a[:, 3].min() >= 0 and a[:, 3].max() <= 1 + 1e-6            # coverage is a probability
    assert a[:, :3].min() >= 0 and a[:, :3].max() <= 1 + 1e-5          # convex blend of sigmoid textures over a black background
    assert 0.02 < (a[:, 3] > 0.01).float().mean().item() < 0.9         # the blocks are in view


def test_fullsize_view_sharding_is_exact(full):
    """rendering a contiguous shard of the views alone == the same slice of the full batch (what parallel.py relies on)."""
    a, _, _ = _render_blocks(full)
    for sl in (slice(0, 7), slice(7, 13), slice(43, 49)):
        b, _, _ = _render_blocks(full, sl)
        assert torch.equal(a[sl], b)


def test_fullsize_constant_texture_shift_is_coverage(full):
    """sum_k occ_k alpha_k = 1 - occ_K = A: adding c to every texel adds c*A to the blended RGB (black background)."""
    a, sc, _ = _render_blocks(full)
    b, _, _ = _render_blocks(full, maps=sc['maps'] - 0.25)
    assert (a[:, :3] - b[:, :3] - 0.25 * a[:, 3:]).abs().max().item() < 2e-6
    assert torch.equal(a[:, 3], b[:, 3])


def test_fullsize_backward_is_linear_and_shards_sum(full):
    dev = full['dev']
    g = torch.Generator().manual_seed(0)
    w1 = torch.rand(49, 4, 400, 400, generator=g).to(dev)
    w2 = torch.rand(49, 4, 400, 400, generator=g).to(dev)

    def grads(w, sl=slice(None)):
        out, sc, fa = _render_blocks(full, sl, requires_grad=True)
        (out * w[sl]).sum().backward()
        return sc['verts'].grad, sc['maps'].grad, fa.grad

    g1, g2, g12 = grads(w1), grads(w2), grads(2 * w1 - 0.5 * w2)
    for a, b, c in zip(g1, g2, g12):
        ref = 2 * a - 0.5 * b
        assert ((c - ref).norm() / ref.norm()).item() < 1e-4           # fp32 atomics: order-dependent rounding only
    # gradient of the full batch == sum of the gradients of the view shards (the all-reduce of parallel.py)
    parts = [grads(w1, sl) for sl in (slice(0, 13), slice(13, 25), slice(25, 49))]
    for i, a in enumerate(g1):
        tot = sum(pp[i] for pp in parts)
        assert ((tot - a).norm() / a.norm()).item() < 1e-4


def test_fullsize_env_pass_covers_every_pixel(full):
    sc = scene_to_device(full['env'], full['dev'])
    out = render_product(sc, full['R'], full['T'], full['K'], (400, 400), 0.0, 1, z_clip=0.001)
    assert torch.isfinite(out).all()
    assert (out[:, 3] == 1).float().mean().item() > 0.9999             # camera sits inside the background sphere
