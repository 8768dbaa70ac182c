"""CPU: the DEVICE math of the rasterization kernels (differentiable-blocksworld_b200/csrc/dbw_math.cuh) compiled for the
host (tests/host_math/dbw_math_host.cpp, g++ -ffp-contract=off) and checked against the oracle: the per-(pixel, face)
candidate test, barycentrics / depth / signed distance, their backward pieces, and the bilinear texture tap.  The CUDA
kernels inline exactly these functions, so a regression in the math shows up here without a GPU; the kernels' own
orchestration (binning, top-K, atomics, blending) is covered by the `-m gpu` parity tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pt3d

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_math', 'dbw_math_host.cpp')
BLUR = float(np.log(1. / 1e-4 - 1.) * 1e-4)        # renderer.py:51 at sigma = 1e-4


@pytest.fixture(scope='module')
def hm():
    out_dir = os.path.join(HERE, 'host_math', '_build')
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, 'libdbw_math_host.so')
    gxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([gxx, '-O2', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-Wall', '-Werror', '-Wno-unknown-pragmas', SRC, '-o', so], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _triangles(seed, n):
    """random screen-space triangles (NDC x, NDC y, view z): big, small, slivers, partly off-screen, both windings"""
    g = np.random.default_rng(seed)
    tris = []
    for i in range(n):
        c = g.uniform(-1.1, 1.1, 2)
        r = [0.9, 0.3, 0.08][i % 3]
        xy = c + g.uniform(-r, r, (3, 2))
        if i % 5 == 4:                               # sliver
            xy[2] = xy[0] + (xy[1] - xy[0]) * g.uniform(0.2, 0.8) + g.uniform(-2e-3, 2e-3, 2)
        z = g.uniform(0.5, 6.0, 3)
        tris.append(np.concatenate([xy, z[:, None]], 1).astype(np.float32))
    return tris


def _oracle(fv, H, W, blur, persp, clipb, dtype):
    t = torch.from_numpy(fv.astype(np.float64)).to(dtype)[None].clone().requires_grad_(True)
    p2f, zbuf, bary, dists = pt3d._RasterizeFaceVerts.apply(t, torch.zeros(1, dtype=torch.long), torch.ones(1, dtype=torch.long),
                                                            None, (H, W), blur, 1, persp, clipb, False)
    return t, p2f[0, ..., 0], zbuf[0, ..., 0], bary[0, ..., 0, :], dists[0, ..., 0]


def _device(hm, fv, H, W, blur, persp, clipb):
    hit = np.zeros((H, W), np.int32)
    zbuf, dists, bary = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32), np.zeros((H, W, 3), np.float32)
    hm.hm_forward(_p(np.ascontiguousarray(fv)), H, W, ctypes.c_float(blur), int(persp), int(clipb), _p(hit), _p(zbuf), _p(bary), _p(dists))
    return hit, zbuf, bary, dists


@pytest.mark.parametrize('persp,clipb', [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize('blur', [BLUR, 0.0])
def test_candidate_test_and_fragment_values_match_the_oracle(hm, persp, clipb, blur):
    H, W = 24, 40
    n_px = n_flip = 0
    for i, fv in enumerate(_triangles(11, 30)):
        hit, zbuf, bary, dists = _device(hm, fv, H, W, blur, persp, clipb)
        _, p2f32, z32, b32, d32 = (x.detach() for x in _oracle(fv, H, W, blur, persp, clipb, torch.float32))
        _, p2f, z64, b64, d64 = (x.detach() for x in _oracle(fv, H, W, blur, persp, clipb, torch.float64))
        ref32 = (p2f32 >= 0).numpy()
        # discrete decisions: identical to the fp32 oracle except where `dist >= blur` is decided by the last bit (the
        # device multiplies by reciprocals where the oracle divides)
        flips = hit.astype(bool) != ref32
        n_px += H * W; n_flip += int(flips.sum())
        if blur == 0.0:                       # the inside test itself is exact: no flips at all without a halo
            assert not flips.any()
        sliver = i % 5 == 4
        # continuous values: against the fp64 oracle for well-conditioned triangles; slivers (area ~ 1e-3 of the squared
        # edge length) amplify fp32 rounding by that ratio, so they are compared with the fp32 oracle, loosely
        both = hit.astype(bool) & (p2f >= 0).numpy() & ref32
        rz, rb, rd = (z32, b32, d32) if sliver else (z64, b64, d64)
        tol = 2e-3 if sliver else 2e-5
        for got, ref, name in ((zbuf, rz, 'zbuf'), (dists, rd, 'dists')):
            err = np.abs(got[both] - ref.numpy()[both])
            assert (err <= tol * (0.1 + np.abs(ref.numpy()[both]))).all(), (name, i, err.max(initial=0))
        rbn = rb.numpy()[both]
        assert (np.abs(bary[both] - rbn) <= tol * (1 + np.abs(rbn))).all(), i     # unclipped barycentrics are unbounded
    assert n_flip <= 2e-4 * n_px, (n_flip, n_px)


def test_backward_pieces_match_the_oracle(hm):
    H, W = 24, 40
    g = np.random.default_rng(5)
    checked = 0
    for fv in _triangles(23, 24):
        hit, *_ = _device(hm, fv, H, W, BLUR, True, True)
        t, p2f, zbuf, bary, dists = _oracle(fv, H, W, BLUR, True, True, torch.float64)
        both = hit.astype(bool) & (p2f >= 0).numpy()
        if both.sum() < 4:
            continue
        gz = (g.standard_normal((H, W)) * both).astype(np.float32)
        gb = (g.standard_normal((H, W, 3)) * both[..., None]).astype(np.float32)
        gd = (g.standard_normal((H, W)) * both * 50).astype(np.float32)
        got = np.zeros(9, np.float32)
        hm.hm_backward(_p(np.ascontiguousarray(fv)), H, W, 1, 1, _p(both.astype(np.int32)), _p(gz), _p(gb), _p(gd), _p(got))
        loss = (zbuf * torch.from_numpy(gz).double()).sum() + (bary * torch.from_numpy(gb).double()).sum() \
            + (dists * torch.from_numpy(gd).double()).sum()
        loss.backward()
        ref = t.grad[0].reshape(-1).numpy()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12)
        assert rel < 2e-3, (rel, got, ref)
        checked += 1
    assert checked >= 10


def test_bilinear_tap_matches_grid_sample(hm):
    """tex_tap == TexturesUV.sample_textures: grid = 2 uv - 1 on the H-flipped map, bilinear, align_corners=True, border"""
    g = torch.Generator().manual_seed(3)
    for Hm, Wm in ((8, 8), (5, 13)):
        tex = torch.rand(Hm, Wm, 3, generator=g, dtype=torch.float64)
        uv = torch.rand(200, 2, generator=g, dtype=torch.float64) * 1.3 - 0.15          # some taps beyond the border
        uv.requires_grad_(True)
        grid = (uv * 2 - 1)[None, None]
        out = F.grid_sample(tex.flip(0).permute(2, 0, 1)[None], grid, mode='bilinear', align_corners=True, padding_mode='border')[0, :, 0].t()
        jac = torch.stack([torch.autograd.grad(out[:, c].sum(), uv, retain_graph=True)[0] for c in range(3)], 1)   # (n, 3, 2)
        m = np.ascontiguousarray(tex.numpy().astype(np.float32))
        for i in range(uv.shape[0]):
            rgb, du, dv = (np.zeros(3, np.float32) for _ in range(3))
            hm.hm_sample(_p(m), Hm, Wm, ctypes.c_float(float(uv[i, 0].detach())), ctypes.c_float(float(uv[i, 1].detach())), _p(rgb), _p(du), _p(dv))
            assert np.abs(rgb - out[i].detach().numpy()).max() < 2e-5
            assert np.abs(du - jac[i, :, 0].numpy()).max() < 2e-3 and np.abs(dv - jac[i, :, 1].numpy()).max() < 2e-3


@pytest.mark.parametrize('persp', [True, False])
def test_z_clip_of_faces_matches_the_oracle(hm, persp):
    """clip_face (dbw_clip.cuh, the code of face_setup_kernel) vs the oracle's clip_faces (PyTorch3D clip.py semantics):
    which faces are culled / kept / cut into one or two triangles, the new vertices and the barycentric conversion rows"""
    g = np.random.default_rng(9)
    n, z_clip = 400, 0.25
    fv = g.uniform(-1.5, 1.5, (n, 3, 3)).astype(np.float32)
    fv[:, :, 2] = g.uniform(-1.0, 2.0, (n, 3)).astype(np.float32)              # a third of the vertices behind the plane
    fv[:40, :, 2] = np.abs(fv[:40, :, 2]) + z_clip + 0.1                        # some faces entirely in front
    ntri = np.zeros(n, np.int32)
    tri, conv = np.zeros((n, 2, 3, 3), np.float32), np.zeros((n, 2, 3, 3), np.float32)
    hm.hm_clip(_p(np.ascontiguousarray(fv)), n, ctypes.c_float(z_clip), int(persp), _p(ntri), _p(tri), _p(conv))
    ref = pt3d.clip_faces(torch.from_numpy(fv).double(), torch.zeros(1, dtype=torch.long), torch.tensor([n]), z_clip, persp)
    n_behind = (fv[:, :, 2] < z_clip).sum(1)
    assert (ntri == np.array([1, 2, 1, 0])[n_behind]).all()                     # 0 behind: 1, 1 behind: quad -> 2, 2 behind: 1, 3: culled
    assert set(np.unique(ntri)) == {0, 1, 2}
    u2c = ref.u2c.numpy()
    for f in range(n):
        for k in range(ntri[f]):
            c = u2c[f] + k
            assert int(ref.to_unclipped[c]) == f
            assert np.abs(tri[f, k] - ref.face_verts[c].numpy()).max() < 1e-4, (f, k)
            assert np.abs(conv[f, k] - ref.conversion[c].numpy()).max() < 1e-5, (f, k)
        if ntri[f] == 2:
            assert int(ref.neighbor[u2c[f]]) == u2c[f] + 1 and int(ref.neighbor[u2c[f] + 1]) == u2c[f]


def _oracle_queue(K, cands):
    """the per-pixel queue of oracle/raster_oracle.c (RasterizeMeshesNaiveCpu + the clipped-quad rule of clip.py):
    tuples ordered by (depth, face, distance); a half of a z-clipped quad evicts / yields to its other half by |distance|"""
    q = []
    for z, f, d, nb in cands:
        handled = False
        if nb >= 0:
            for i, (_, qf, qd) in enumerate(q):
                if qf == nb:
                    if abs(d) < abs(qd):
                        q.pop(i)
                    else:
                        handled = True
                    break
        if handled:
            continue
        q.append((z, f, d))
        q.sort()
        del q[K:]
    return q


@pytest.mark.parametrize('K', [1, 4, 10, 25])
def test_fragment_list_matches_the_oracle_queue(hm, K):
    """fraglist_offer (dbw_fraglist.cuh, the per-pixel sorted list in shared memory): ordering by (depth, slot) incl. exact
    depth ties, overflow beyond K, the mutual exclusion of the two halves of a z-clipped quad, and that every entry's
    payload (closest edge, u, v) travels with its key -- against the oracle's queue on random candidate streams"""
    g = np.random.default_rng(K)
    for trial in range(300):
        n = int(g.integers(0, 3 * K + 4))
        slots = g.permutation(200)[:n].astype(np.int32)
        pz = g.choice(np.float32([0.5, 0.75, 1.0, 1.5, 2.0, 3.0]), n) if trial % 3 == 0 else g.uniform(0.1, 5.0, n).astype(np.float32)
        sd = (g.uniform(1e-6, 9e-4, n) * g.choice([-1.0, 1.0], n)).astype(np.float32)
        nb = np.full(n, -1, np.int32)
        for i in range(0, n - 1, 5):                       # some consecutive candidates are the two halves of a quad
            if g.random() < 0.6:
                nb[i], nb[i + 1] = slots[i + 1], slots[i]
                if g.random() < 0.5:
                    pz[i + 1] = pz[i]
        pz = np.ascontiguousarray(pz, np.float32)
        out_slot, out_sd = np.zeros(K, np.int32), np.zeros(K, np.float32)
        assert hm.hm_topk(K, n, _p(pz), _p(slots), _p(sd), _p(nb), _p(out_slot), _p(out_sd)) == 0
        ref = _oracle_queue(K, [(float(pz[i]), int(slots[i]), float(sd[i]), int(nb[i])) for i in range(n)])
        got = [(int(s), float(d)) for s, d in zip(out_slot, out_sd) if s >= 0]
        assert got == [(f, d) for _, f, d in ref], (trial, got, ref)


def test_scene_vertex_math_matches_torch(hm):
    """dbw_scene_math.cuh (the per-vertex code of the fused scene-geometry kernels): 6D rotation forward / backward against
    rotation_6d_to_matrix + autograd, the parametric superquadric against the oracle's restatement of superquadric.py:10-14"""
    from oracle import dbw_path as D
    g = torch.Generator().manual_seed(4)
    n = 64
    d6 = torch.randn(n, 6, generator=g, dtype=torch.float64, requires_grad=True)
    R_ref = pt3d.rotation_6d_to_matrix(d6)
    gR = torch.randn(n, 3, 3, generator=g, dtype=torch.float64)
    (R_ref * gR).sum().backward()
    d6f = np.ascontiguousarray(d6.detach().numpy().astype(np.float32))
    R, gd6 = np.zeros((n, 9), np.float32), np.zeros((n, 6), np.float32)
    hm.hm_rot6d(_p(d6f), n, _p(R))
    hm.hm_rot6d_backward(_p(d6f), _p(np.ascontiguousarray(gR.numpy().astype(np.float32).reshape(n, 9))), n, _p(gd6))
    assert np.abs(R.reshape(n, 3, 3) - R_ref.detach().numpy()).max() < 1e-5
    ref_g = d6.grad.numpy()
    assert np.abs(gd6 - ref_g).max() <= 1e-4 * max(1.0, np.abs(ref_g).max())
    # superquadric vertices on the icosphere's (eta, omega) of the scene template
    tpl = D.SceneTemplate(n_blocks=3, txt_size=8)
    eta, omega = tpl.sq_eta.double(), tpl.sq_omega.double()
    raw = torch.tensor([[-2.0, 0.3], [0.0, 0.0], [1.5, -0.7]], dtype=torch.float64)
    eps = torch.sigmoid(raw) * 1.8 + 0.1                                  # dbw.py:349
    ref = D.parametric_sq(eta, omega, eps[:, :1], eps[:, 1:]) * 0.25
    N, Vb = eta.shape
    out, aux = np.zeros((N, Vb, 3), np.float32), np.zeros((N, Vb, 6), np.float32)
    hm.hm_superquadric(_p(np.ascontiguousarray(eta.numpy().astype(np.float32))), _p(np.ascontiguousarray(omega.numpy().astype(np.float32))),
                       _p(np.ascontiguousarray(raw.numpy().astype(np.float32))), N, Vb, ctypes.c_float(0.25), _p(out), _p(aux))
    assert np.abs(out - ref.numpy()).max() < 1e-5
    assert np.abs(aux[:, 0, 4:] - eps.numpy()).max() < 1e-6


@pytest.mark.parametrize('blur', [BLUR, 0.0])
def test_tile_binning_test_never_drops_a_face_that_reaches_the_tile(hm, blur):
    """tri_overlaps_rect (the binner's edge-line test on the halo-expanded tile) is CONSERVATIVE: every tile that holds a pixel
    the per-pixel candidate test accepts lists the face; and it is useful: it rejects most tiles of the face's bounding box
    that the face does not reach (slivers and diagonal faces are why it exists)"""
    H, W, TS = 48, 64, 8
    ntx, nty = W // TS, H // TS
    kept_needlessly = boxed = 0
    for fv in _triangles(3, 60):
        hit, *_ = _device(hm, fv, H, W, blur, True, True)
        listed = np.zeros((nty, ntx), np.int32)
        hm.hm_tile_overlap(_p(np.ascontiguousarray(fv)), H, W, TS, ctypes.c_float(blur), _p(listed))
        reached = hit.reshape(nty, TS, ntx, TS).any(axis=(1, 3))
        assert not (reached & (listed == 0)).any(), 'a tile with accepted pixels was not listed'
        kept_needlessly += int(((listed == 1) & ~reached).sum())
        boxed += int((~reached).sum())
    assert kept_needlessly < 0.25 * boxed


def test_closest_edge_recorded_by_the_forward(hm):
    """tri_dist2_edge == tri_dist2 bit for bit, and the edge it names realises the minimum (first in the order v0v1, v0v2, v1v2
    among equals -- the order tri_dist_backward resolves ties in), incl. at pixels equidistant from two edges (a vertex region)"""
    H, W = 40, 56
    ties = 0
    for fv in _triangles(9, 30):
        dist, ref = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
        edge, seg = np.zeros((H, W), np.int32), np.zeros((H, W, 3), np.float32)
        hm.hm_dist_edge(_p(np.ascontiguousarray(fv)), H, W, _p(dist), _p(ref), _p(edge), _p(seg))
        assert np.array_equal(dist, ref)
        assert np.array_equal(np.take_along_axis(seg, edge[..., None].astype(np.int64), -1)[..., 0], dist)
        assert np.array_equal(edge, np.argmin(seg, axis=-1))            # argmin takes the FIRST minimum
        ties += int((np.sort(seg, axis=-1)[..., 0] == np.sort(seg, axis=-1)[..., 1]).sum())
    assert ties > 0                                                      # vertex regions: two segments share the closest point
