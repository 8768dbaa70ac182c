"""Diagnostic (GPU): where do the environment pass' vertex gradients deviate from the fp64 oracle at the cfg 5 camera?
Prints, per ground vertex with the largest deviation: CUDA vs fp64 vs fp32-oracle gradient, and whether the vertex belongs to
a z-clipped face.  python scripts/diag_ground_grad.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dbw_path as D, pt3d
from tests.helpers import scene_to_device, render_product

dev = torch.device('cuda:0')
tpl = D.SceneTemplate(n_blocks=50, txt_size=128, txt_bkg_upscale=2)
R, T, K = D.ring_cameras(256, dtype=torch.float64)
R, T = R[[17]], T[[17]]
size = (800, 800)
g = torch.Generator().manual_seed(7)
w = torch.rand(1, 4, *size, generator=g, dtype=torch.float64)
out = {}
for name, dt in (('f64', torch.float64), ('f32', torch.float32)):
    p = {k: v.clone().requires_grad_(True) for k, v in D.init_params(50, 128, txt_bkg_upscale=2, seed=227391, boxy=True, dtype=dt).items()}
    env = tpl.build_env(p, decimate=8)
    env['verts'].retain_grad()
    img = D.render(env, R.to(dt), T.to(dt), K.to(dt), size, sigma=0, faces_per_pixel=1, z_clip=0.001, detach_bary=False)
    (img * w.to(dt)).sum().backward()
    out[name] = (env['verts'].grad.double(), p['T_ground'].grad.double(), p['R_6d_ground'].grad.double(), img.detach().double(), env)
env = out['f64'][4]
sc = scene_to_device(env, dev, requires_grad=True)
img = render_product(sc, R.to(dev), T.to(dev), K, size, 0.0, 1, z_clip=0.001)
(img * w.float().to(dev)).sum().backward()
gc = sc['verts'].grad.cpu().double()
g64, g32 = out['f64'][0], out['f32'][0]
nb = tpl.bkg_verts.shape[0]
print('image max |cuda-f64| %.2e  |f32-f64| %.2e' % ((img.detach().cpu().double() - out['f64'][3]).abs().max(), (out['f32'][3] - out['f64'][3]).abs().max()))
print('ground verts grad rel: cuda %.2e  f32 %.2e' % ((gc[nb:] - g64[nb:]).norm() / g64[nb:].norm(), (g32[nb:] - g64[nb:]).norm() / g64[nb:].norm()))
print('sum over ground verts (~T_ground direction): cuda', gc[nb:].sum(0).tolist(), 'f64', g64[nb:].sum(0).tolist(), 'f32', g32[nb:].sum(0).tolist())
ndc = pt3d.world_to_ndc(env['verts'].detach(), R, T, K)[0]
faces = env['faces']
zmin = ndc[faces][:, :, 2].min(1)[0]; zmax = ndc[faces][:, :, 2].max(1)[0]
clipped_face = (zmin < 0.001) & (zmax > 0.001)
vclip = torch.zeros(len(ndc), dtype=torch.bool); vclip[faces[clipped_face].reshape(-1)] = True
dev_c = (gc - g64).norm(dim=1); dev_32 = (g32 - g64).norm(dim=1)
idx = torch.argsort(dev_c, descending=True)[:10]
for i in idx.tolist():
    print(f'v{i} ground={i >= nb} in_clipped_face={bool(vclip[i])} z={ndc[i, 2]:.3f} |g64|={g64[i].norm():.3e} dev_cuda={dev_c[i]:.3e} dev_f32={dev_32[i]:.3e}')
print('total dev: clipped-face verts %.3e, others %.3e (cuda); f32: %.3e / %.3e' % (dev_c[vclip].norm(), dev_c[~vclip].norm(), dev_32[vclip].norm(), dev_32[~vclip].norm()))
