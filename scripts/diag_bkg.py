import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dbw_path as D
from tests.test_baseline_shapes_gpu import _bench_model, _ring, RING_VIEWS
import dbw_b200
from dbw_b200 import _lib
dev = torch.device('cuda:0')
R, T, K = _ring(RING_VIEWS)
for generic in (1, 0):
    _lib.lib().dbw_debug_generic_kernel_only(generic)
    model = _bench_model((400, 400))
    model.opacity_noise_buffer = torch.zeros(10, device=dev)
    g = torch.Generator().manual_seed(7)
    imgs = torch.rand(3, 3, 400, 400, generator=g)
    inp = {'imgs': imgs.to(dev), 'R': R.float().to(dev), 'T': T.float().to(dev), 'K': K.float()[None].expand(3, -1, -1).to(dev)}
    losses = model(inp, None)
    losses['total'].backward()
    gb = model.texture_bkg.grad
    print('generic' if generic else 'hard   ', 'loss', losses['rgb'].item(), 'texture_bkg grad max', gb.abs().max().item(), 'nonzero', int((gb != 0).sum()),
          'texture_ground grad norm', model.texture_ground.grad.norm().item())
    # which env faces do the pixels see?
    st = model._static_arrays()
    (ev, ea, et), _ = model._scene_tensors(False)
    from dbw_b200.renderer import render_scene
    r = model.renderer_env
    with torch.no_grad():
        out, ids = render_scene(ev, st['faces_e'], st['fvu_e'], st['fmap_e'], ea, et, inp['R'], inp['T'], r.cameras.intrinsics(), r.img_size, 0.0, 1,
                                r.z_clip, False, True, (0., 0., 0.), None, True, blur_radius=0.0, return_ids=True, maps_are_texels4=True)
    F = st['faces_e'].shape[0]
    face = torch.where(ids >= F, ids - F, ids)
    print('   pixels on bkg faces', int(((face >= 0) & (face < model.bkg_n_faces)).sum()), 'empty', int((ids < 0).sum()))
_lib.lib().dbw_debug_generic_kernel_only(0)
