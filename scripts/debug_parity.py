import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dbw_path as D
from tests.helpers import scene_to_device, render_product, split_map_grads
dev = torch.device('cuda:0')

def f32scene(scene):
    return {k: ([m.detach().float() for m in v] if k == 'maps' else (v.detach().float() if v.is_floating_point() else v)) for k, v in scene.items()}

# ---- z-clip case
tpl = D.SceneTemplate(n_blocks=3, txt_size=32)
p = D.init_params(3, 32, seed=3, dtype=torch.float64)
R, T, K = D.ring_cameras(3, dtype=torch.float64, jitter=0.3, seed=3, dist=0.45)
K = K.clone(); K[0, 0] = K[1, 1] = 1.2
blocks, alpha = tpl.build_blocks(p)
env = tpl.build_env(p)
scene = D.join_scenes([env, blocks])
s32 = f32scene(scene)
ref32, fr32 = D.render(s32, R.float(), T.float(), K.float(), (48, 48), sigma=1e-4, faces_per_pixel=8, z_clip=0.05, detach_bary=False, return_fragments=True)
sc = scene_to_device(scene, dev)
out, ids = render_product(sc, R.to(dev), T.to(dev), K, (48, 48), 1e-4, 8, z_clip=0.05, detach_bary=False, return_ids=True)
err = (out.cpu() - ref32).abs()
print('zclip max err', err.max().item(), 'bad', (err > 1e-4).sum().item())
Fn = scene['faces'].shape[0]
bad = (err > 1e-4).any(1).nonzero()
from oracle import pt3d
ndc = pt3d.world_to_ndc(s32['verts'], R.float(), T.float(), K.float())
fv = ndc[:, s32['faces']]
nbeh = (fv[..., 2] < 0.05).sum(-1)
print('faces by n_behind per view', [(nbeh[b] == i).sum().item() for b in range(3) for i in range(4)])
for b, y, x in bad.tolist()[:8]:
    o = fr32.pix_to_face[b, y, x]
    print('px', b, y, x, 'oracle faces', (o % Fn).tolist(), 'nbehind', [int(nbeh[b, f % Fn]) if f >= 0 else -1 for f in o.tolist()],
          'cuda slots', ids[b, :, y, x].tolist(), 'out', out[b, :, y, x].tolist(), 'ref', ref32[b, :, y, x].tolist())
    print('   oracle dists', fr32.dists[b, y, x].tolist(), 'z', fr32.zbuf[b, y, x].tolist())

# ---- env TypeError
try:
    tpl = D.SceneTemplate(n_blocks=4, txt_size=32)
    p = D.init_params(4, 32, seed=3, dtype=torch.float64)
    R, T, K = D.ring_cameras(2, dtype=torch.float64, jitter=0.3, seed=3)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    env = tpl.build_env(p)
    wgt = torch.rand(2, 4, 64, 64, dtype=torch.float64)
    ref = D.render(env, R, T, K, (64, 64), sigma=0, faces_per_pixel=1, z_clip=0.001, detach_bary=False)
    for m in env['maps']: m.retain_grad()
    env['verts'].retain_grad()
    (ref * wgt).sum().backward()
    sc = scene_to_device(env, dev, requires_grad=True)
    out = render_product(sc, R.to(dev), T.to(dev), K, (64, 64), 0.0, 1, z_clip=0.001, detach_bary=False)
    print('env img err', (out.detach().cpu().double() - ref.detach()).abs().max().item())
    (out * wgt.to(dev).float()).sum().backward()
    gv, gv_ref = sc['verts'].grad.cpu().double(), env['verts'].grad
    print('env verts grad rel', ((gv - gv_ref).norm() / gv_ref.norm()).item(), gv_ref.norm().item())
    for g, m in zip(split_map_grads(sc['maps'].grad.cpu().double(), sc['table']), env['maps']):
        print('env map grad rel', ((g - m.grad).norm() / m.grad.norm()).item())
except Exception:
    traceback.print_exc()
