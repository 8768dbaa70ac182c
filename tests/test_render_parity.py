"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): rendered RGB / silhouette within 1e-4 abs, gradients within 1e-3 rel.
Gradients are compared with the float64 oracle (the reference's own fp32 cumprod backward is
noisy where 1 - alpha is tiny), as relative L2 error per tensor plus an element-wise bound scaled by the tensor's
max magnitude."""
import pytest
import torch

from oracle import dbw_path as D
from tests.helpers import scene_to_device, render_product, split_map_grads, decision_mask

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_REL = 1e-3


def _dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# Faces that share an edge or a vertex can reach a halo pixel with depths equal to ~1e-7 relative (their clamped
# barycentrics land on the same point of the shared edge); which of them sorts first is then decided by the last bit
# of pz -- in PyTorch3D too (its CPU and CUDA kernels need not agree there).  Such near-tie flips are the only
# tolerated deviation: at most NEAR_TIE_FRAC of the values may exceed IMG_TOL, and never by more than NEAR_TIE_MAX.
NEAR_TIE_FRAC = 1e-4
NEAR_TIE_MAX = 2e-2


def _check_image(out, ref, max_bad_frac=NEAR_TIE_FRAC, max_err=NEAR_TIE_MAX):
    err = (out - ref).abs()
    bad = (err > IMG_TOL).float().mean().item()
    assert bad <= max_bad_frac and err.max().item() <= max(max_err, IMG_TOL), \
        f'max err {err.max().item():.3e}, {bad * 100:.4f}% of values above {IMG_TOL}'


def _setup(n_blocks=4, txt=32, n_views=2, seed=3, boxy=False, dtype=torch.float32, dist=2.75):
    tpl = D.SceneTemplate(n_blocks=n_blocks, txt_size=txt)
    p = D.init_params(n_blocks, txt, seed=seed, boxy=boxy, dtype=dtype)
    R, T, K = D.ring_cameras(n_views, dtype=dtype, jitter=0.3, seed=seed, dist=dist)
    return tpl, p, R, T, K


@pytest.mark.parametrize('size', [(64, 64), (48, 80), (70, 50)])
@pytest.mark.parametrize('mode', ['coarse', 'fine', 'sigmoid'])
def test_blocks_forward(size, mode):
    dev = _dev()
    tpl, p, R, T, K = _setup(boxy=(mode == 'fine'))
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF) if mode != 'fine' else None
    sigma = {'coarse': 1e-4, 'fine': 5e-6, 'sigmoid': 1e-4}[mode]
    clip_inside = mode != 'sigmoid'
    ref = D.render(blocks, R, T, K, size, sigma=sigma, faces_per_pixel=10, z_clip=0.001, detach_bary=True,
                   faces_alpha=fa, clip_inside=clip_inside, background=(0.1, 0.2, 0.3))
    sc = scene_to_device(blocks, dev)
    out = render_product(sc, R.to(dev), T.to(dev), K, size, sigma, 10, z_clip=0.001, detach_bary=True,
                         faces_alpha=None if fa is None else fa.to(dev), clip_inside=clip_inside,
                         background=(0.1, 0.2, 0.3))
    _check_image(out.cpu(), ref)


@pytest.mark.parametrize('K_', [1, 3, 10, 25, 40])
def test_blocks_forward_faces_per_pixel(K_):
    dev = _dev()
    tpl, p, R, T, K = _setup(n_blocks=6)
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF).repeat(R.shape[0])          # batch-packed (B*F,) as dbw.py:219
    ref = D.render(blocks, R, T, K, (64, 64), sigma=1e-4, faces_per_pixel=K_, z_clip=0.001, detach_bary=True, faces_alpha=fa)
    sc = scene_to_device(blocks, dev)
    out = render_product(sc, R.to(dev), T.to(dev), K, (64, 64), 1e-4, K_, z_clip=0.001, detach_bary=True, faces_alpha=fa.to(dev))
    _check_image(out.cpu(), ref)


def test_topk_ids_match_oracle():
    """bit-exact integer parity: the per-pixel z-sorted face ids equal the oracle's pix_to_face."""
    dev = _dev()
    tpl, p, R, T, K = _setup(n_blocks=5)
    blocks, alpha = tpl.build_blocks(p)
    _, frags = D.render(blocks, R, T, K, (64, 64), sigma=1e-4, faces_per_pixel=10, z_clip=0.001, return_fragments=True)
    sc = scene_to_device(blocks, dev)
    _, ids = render_product(sc, R.to(dev), T.to(dev), K, (64, 64), 1e-4, 10, z_clip=0.001, return_ids=True)
    Fn = blocks['faces'].shape[0]
    ref_ids = torch.where(frags.pix_to_face >= 0, frags.pix_to_face % Fn, frags.pix_to_face).permute(0, 3, 1, 2)
    mism = (ids.cpu().long() != ref_ids).float().mean().item()
    assert mism <= 1e-4, f'{mism * 100:.4f}% of top-K ids differ'


def _grad_parity(scene, R, T, K, size, sigma, Kf, z_clip, detach, fa, clip_inside, seed, max_ambiguous=2e-3,
                 img_bad_frac=0.0, img_max_err=IMG_TOL):
    """forward+backward of the CUDA path vs the FLOAT64 oracle.  A few pixels take a different discrete decision in
    fp32 than in fp64 (inside test / halo cut-off / K-th face / which half of a z-clipped quad); they are identified by
    comparing the kept face ids, excluded from the loss on BOTH sides, counted and bounded -- everything else must
    agree to 1e-4 (image) and 1e-3 relative (gradients)."""
    dev = _dev()
    B = R.shape[0]
    ref, frags = D.render(scene, R, T, K, size, sigma=sigma, faces_per_pixel=Kf, z_clip=z_clip, detach_bary=detach,
                          faces_alpha=fa, clip_inside=clip_inside, return_fragments=True)
    sc = scene_to_device(scene, dev, requires_grad=True)
    fa_d = None if fa is None else fa.detach().float().to(dev).requires_grad_(True)
    out, ids = render_product(sc, R.to(dev), T.to(dev), K, size, sigma, Kf, z_clip=z_clip, detach_bary=detach,
                              faces_alpha=fa_d, clip_inside=clip_inside, return_ids=True)
    mask = decision_mask(ids, frags, scene['faces'].shape[0])
    ambiguous = 1 - mask.mean().item()
    assert ambiguous <= max_ambiguous, f'{ambiguous * 100:.3f}% of pixels take a different discrete decision'
    _check_image(out.detach().cpu().double() * mask, ref.detach() * mask, max_bad_frac=img_bad_frac, max_err=img_max_err)
    gen = torch.Generator().manual_seed(seed)
    wgt = torch.rand(B, 4, *size, generator=gen, dtype=torch.float64) * mask
    scene['verts'].retain_grad()
    for m in scene['maps']:
        m.retain_grad()
    if fa is not None:
        fa.retain_grad()
    (ref * wgt).sum().backward()
    (out * wgt.to(dev).float()).sum().backward()
    gv, gv_ref = sc['verts'].grad.cpu().double(), scene['verts'].grad
    assert _rel(gv, gv_ref) < GRAD_REL, f'verts grad rel err {_rel(gv, gv_ref):.3e}'
    assert (gv - gv_ref).abs().max() <= 2e-3 * gv_ref.abs().max()
    for g, m in zip(split_map_grads(sc['maps'].grad.cpu().double(), sc['table']), scene['maps']):
        if m.grad is None:
            assert g.abs().max() == 0
        else:
            assert _rel(g, m.grad) < GRAD_REL, f'map grad rel err {_rel(g, m.grad):.3e}'
    if fa is not None:
        assert _rel(fa_d.grad.cpu().double(), fa.grad) < GRAD_REL, f'faces_alpha grad rel err {_rel(fa_d.grad.cpu().double(), fa.grad):.3e}'
    return ambiguous


def test_env_forward_backward():
    """sigma = 0, K = 1, gradient through barycentrics to the ground pose and both env textures (dbw.py:135-138)."""
    tpl, p, R, T, K = _setup(dtype=torch.float64)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    env = tpl.build_env(p)
    _grad_parity(env, R, T, K, (64, 64), 0.0, 1, 0.001, False, None, True, seed=0)


@pytest.mark.parametrize('mode', ['coarse', 'fine', 'sigmoid', 'bary'])
def test_blocks_backward(mode):
    tpl, p, R, T, K = _setup(n_blocks=4, dtype=torch.float64, boxy=(mode == 'fine'))
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF) if mode != 'fine' else None
    sigma = 5e-6 if mode == 'fine' else 1e-4
    _grad_parity(blocks, R, T, K, (64, 64), sigma, 10, 0.001, mode != 'bary', fa, mode != 'sigmoid', seed=1)


def test_blocks_backward_batch_packed_alpha():
    """faces_alpha of length B*F exactly as the reference builds it (dbw.py:219)."""
    tpl, p, R, T, K = _setup(n_blocks=3, dtype=torch.float64, n_views=3)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF).repeat(3)
    _grad_parity(blocks, R, T, K, (40, 56), 1e-4, 10, 0.001, True, fa, True, seed=4)


def test_z_clipped_faces_forward_backward():
    """camera inside the geometry so that faces straddle z = z_clip and are split: forward parity and the
    gradient through the clipped vertices / barycentric conversion.  Pixels in the Voronoi cell of a vertex shared by
    the two halves of a split quad have mathematically tied distances to both halves, so which half is kept is decided
    by rounding (in PyTorch3D too): those are the 'ambiguous' pixels here."""
    tpl, p, R, T, K = _setup(n_blocks=3, dtype=torch.float64, dist=0.45, n_views=3)
    K = K.clone(); K[0, 0] = K[1, 1] = 1.2            # wide field of view so that clipped faces are visible
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    blocks, alpha = tpl.build_blocks(p)
    env = tpl.build_env(p)
    scene = D.join_scenes([env, blocks])
    amb = _grad_parity(scene, R, T, K, (48, 48), 1e-4, 8, 0.05, False, None, True, seed=2, max_ambiguous=1e-2)
    print(f'ambiguous pixels: {amb * 100:.3f}%')


def test_verts_are_ndc_entry():
    """the kernel's differentiable input can be NDC vertices directly (SURVEY 8b recommended split)."""
    dev = _dev()
    tpl, p, R, T, K = _setup()
    blocks, alpha = tpl.build_blocks(p)
    from oracle import pt3d
    ndc = pt3d.world_to_ndc(blocks['verts'], R, T, K)
    ref = D.render(blocks, R, T, K, (64, 64), sigma=1e-4, faces_per_pixel=10, z_clip=0.001)
    sc = scene_to_device(blocks, dev)
    sc['verts'] = ndc.to(dev).contiguous()
    out = render_product(sc, None, None, K, (64, 64), 1e-4, 10, z_clip=0.001, verts_are_ndc=True)
    _check_image(out.cpu(), ref)


@pytest.mark.parametrize('faces_per_pixel', [25, 10])
def test_stress_shape_many_faces(faces_per_pixel):
    """BASELINE configs[4] shape at reduced resolution: 50 blocks (4000 faces), K = 25 -- exercises the chunked tile
    lists (more listed faces than the shared-memory list holds at once: 256 entries for K = 25, 128 for the K <= 10
    kernels) and the K = 25 register top-K."""
    dev = _dev()
    tpl = D.SceneTemplate(n_blocks=50, txt_size=16)
    p = D.init_params(50, 16, seed=7, boxy=True)
    p['T'] = p['T'] * 0.35                       # crowd the blocks so that many faces overlap the same tiles
    R, T, K = D.ring_cameras(2, jitter=0.3, seed=7, dist=2.0)
    blocks, alpha = tpl.build_blocks(p)
    fa = alpha.repeat_interleave(tpl.BNF)
    ref, fr = D.render(blocks, R, T, K, (96, 128), sigma=1e-4, faces_per_pixel=faces_per_pixel, z_clip=0.001, detach_bary=True,
                       faces_alpha=fa, return_fragments=True)
    sc = scene_to_device(blocks, dev)
    out, ids = render_product(sc, R.to(dev), T.to(dev), K, (96, 128), 1e-4, faces_per_pixel, z_clip=0.001, detach_bary=True,
                              faces_alpha=fa.to(dev), return_ids=True)
    assert (fr.pix_to_face[..., -1] >= 0).float().mean() > 0.01       # K really is exceeded somewhere
    _check_image(out.cpu(), ref, max_bad_frac=3e-4)
    mism = 1 - decision_mask(ids, fr, blocks['faces'].shape[0]).mean().item()
    assert mism < 2e-3, mism


def test_bmvs_shape_non_square_backward():
    """BASELINE configs[3] aspect (576x768 -> 72x96 here), fine phase."""
    tpl, p, R, T, K = _setup(n_blocks=5, dtype=torch.float64, n_views=2, boxy=True)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    blocks, alpha = tpl.build_blocks(p)
    _grad_parity(blocks, R, T, K, (72, 96), 5e-6, 10, 0.001, True, None, True, seed=6)


def test_flat_shaded_viz_render_matches_oracle():
    """renderer_light path: flat shading with a camera-fixed directional light, white background,
    4x supersampled hard render + 4x4 box filter (viz_purpose=True)."""
    import torch.nn.functional as F
    from dbw_b200 import Renderer, Meshes, TexturesUV, join_meshes_as_scene
    dev = _dev()
    tpl, p, R, T, K = _setup(n_blocks=3)
    blocks, alpha = tpl.build_blocks(p)
    shade = D.flat_shade_multiplier(blocks['verts'], blocks['faces'], R)
    ref = D.render(blocks, R, T, K, (128, 160), sigma=0, faces_per_pixel=1, z_clip=0.001, background=(1., 1., 1.), face_shade=shade)
    ref = F.avg_pool2d(ref, 4, 4)
    rend = Renderer((32, 40), faces_per_pixel=1, sigma=0, z_clip=0.001, cameras={'name': 'perspective'}, shading_type='flat',
                    background_color=(1, 1, 1), lights={'name': 'directional', 'direction': [[1, 0.25, -1]],
                                                        'ambient_color': [[0.7] * 3], 'diffuse_color': [[0.4] * 3],
                                                        'specular_color': [[0.] * 3]}).to(dev)
    rend.update_cameras(device=dev, K=K[None].to(dev))
    # the same scene as Meshes / TexturesUV objects (3 blocks joined)
    nb, Fn = 3, tpl.BNF
    verts = blocks['verts'].reshape(nb, -1, 3).to(dev)
    faces = tpl.block_faces[None].expand(nb, -1, -1).to(dev)
    maps = torch.stack(blocks['maps']).to(dev)
    txt = TexturesUV(maps, tpl.block_faces_uvs[None].expand(nb, -1, -1).to(dev), tpl.block_verts_uvs[None].expand(nb, -1, -1).to(dev))
    scene = join_meshes_as_scene(Meshes(verts, faces, txt))
    out = rend(scene.extend(2), R.to(dev), T.to(dev), viz_purpose=True)
    err = (out.cpu() - ref).abs()
    assert (err > 1e-4).float().mean().item() < 2e-3 and err.max().item() < 0.1, err.max().item()   # 16 hard samples per pixel


def test_render_edges_matches_oracle():
    """edge overlay: K=1 hard rasterization, signed distances thresholded at the line width."""
    from dbw_b200 import Renderer, Meshes, TexturesUV, join_meshes_as_scene
    dev = _dev()
    tpl, p, R, T, K = _setup(n_blocks=3)
    blocks, alpha = tpl.build_blocks(p)
    mask_ref, p2f_ref = D.render_edges(blocks, R, T, K, (96, 128), linewidth=2, z_clip=0.001)
    rend = Renderer((96, 128), faces_per_pixel=10, z_clip=0.001, cameras={'name': 'perspective'}).to(dev)
    rend.update_cameras(device=dev, K=K[None].to(dev))
    nb = 3
    txt = TexturesUV(torch.stack(blocks['maps']).to(dev), tpl.block_faces_uvs[None].expand(nb, -1, -1).to(dev),
                     tpl.block_verts_uvs[None].expand(nb, -1, -1).to(dev))
    scene = join_meshes_as_scene(Meshes(blocks['verts'].reshape(nb, -1, 3).to(dev), tpl.block_faces[None].expand(nb, -1, -1).to(dev), txt))
    mask, p2f = rend.render_edges(scene.extend(2), R.to(dev), T.to(dev), linewidth=2, return_pix2face=True)
    assert (p2f.cpu() != p2f_ref).float().mean().item() < 1e-4
    assert (mask.cpu() != mask_ref).float().mean().item() < 2e-4
    img = torch.rand(2, 3, 24, 32, device=dev)
    drawn = rend.draw_edges(img, scene.extend(2), R.to(dev), T.to(dev), colors=(1, 0, 0), linewidth=1)
    assert drawn.shape == img.shape and torch.isfinite(drawn).all() and (drawn - img).abs().max() > 0.1


@pytest.mark.parametrize('size,dist', [((64, 64), 2.75), ((400, 400), 2.75), ((72, 96), 0.45)])
def test_hard_single_layer_kernel_equals_generic_kernel(size, dist):
    """K = 1, sigma = 0 renders have a dedicated kernel (visible-slot lists, register-resident best fragment): image, ids and
    the backward's gradients must equal the generic kernel's bit for bit (same edge functions, same (depth, slot) order)"""
    from dbw_b200 import _lib
    dev = _dev()
    tpl, p, R, T, K = _setup(n_blocks=3, n_views=3, dist=dist)
    if dist < 1:
        K = K.clone(); K[0, 0] = K[1, 1] = 1.2
    scene = D.join_scenes([tpl.build_env(p), tpl.build_blocks(p)[0]])
    outs = []
    for generic in (1, 0):
        _lib.lib().dbw_debug_generic_kernel_only(generic)
        try:
            sc = scene_to_device(scene, dev, requires_grad=True)
            out, ids = render_product(sc, R.to(dev), T.to(dev), K, size, 0.0, 1, z_clip=0.001 if dist > 1 else 0.05, return_ids=True)
            (out * torch.linspace(0.5, 1.5, out.numel(), device=dev).view_as(out)).sum().backward()
            outs.append((out.detach().clone(), ids.clone(), sc['verts'].grad.clone(), sc['maps'].grad.clone()))
        finally:
            _lib.lib().dbw_debug_generic_kernel_only(0)
    (a, ia, gva, gma), (b, ib, gvb, gmb) = outs
    assert torch.equal(ia, ib), f'{(ia != ib).float().mean().item() * 100:.4f}% of the ids differ'
    assert torch.equal(a, b)
    assert (gva - gvb).norm() <= 1e-5 * gva.norm() and (gma - gmb).norm() <= 1e-5 * gma.norm()      # atomics: order only


def test_hard_single_layer_kernel_many_visible_faces():
    """4000 crowded faces, all on screen: the visible list takes 16 rounds of 256 slots per tile and tile lists overflow
    their 64 entries (several rounds per batch) -- image and ids still equal the generic kernel's bit for bit"""
    from dbw_b200 import _lib
    dev = _dev()
    tpl = D.SceneTemplate(n_blocks=50, txt_size=16)
    p = D.init_params(50, 16, seed=7, boxy=True)
    p['T'] = p['T'] * 0.2
    R, T, K = D.ring_cameras(2, jitter=0.3, seed=7, dist=1.6)
    blocks, _ = tpl.build_blocks(p)
    outs = []
    for generic in (1, 0):
        _lib.lib().dbw_debug_generic_kernel_only(generic)
        try:
            out, ids = render_product(scene_to_device(blocks, dev), R.to(dev), T.to(dev), K, (96, 128), 0.0, 1, z_clip=0.001, return_ids=True)
            outs.append((out.clone(), ids.clone()))
        finally:
            _lib.lib().dbw_debug_generic_kernel_only(0)
    assert (outs[0][1] >= 0).float().mean() > 0.3
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    ref = D.render(blocks, R, T, K, (96, 128), sigma=0, faces_per_pixel=1, z_clip=0.001)
    _check_image(outs[1][0].cpu(), ref, max_bad_frac=3e-4)
