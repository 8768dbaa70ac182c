import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from copy import deepcopy
import dbw_b200
from dbw_b200.dbw import DifferentiableBlocksWorld
from dbw_b200 import join_meshes_as_scene, render_scene
dev = torch.device('cuda:0')
torch.manual_seed(bench.SEED)
model = DifferentiableBlocksWorld((400, 400), **deepcopy(bench.MODEL_CFG)).to(dev); model.train()
inp = bench.synthetic_inputs(4, 400, 400, device=dev)
model._install_cameras(inp)
env = join_meshes_as_scene([model.build_bkg(world_coord=True), model.build_ground(world_coord=True)])
verts, faces = env.get_mesh_verts_faces(0)
R, T = inp['R'], inp['T']
zv = (verts[None] @ R + T[:, None])[..., 2]            # (B,V)
fz = zv[:, faces]                                       # (B,F,3)
nb = (fz < 0.001).sum(-1)
print('faces by n_behind per view:', [[int((nb[b] == i).sum()) for i in range(4)] for b in range(4)])
fvu, fmap = env.textures.scene_arrays(); maps, table = env.textures.packed_maps()
out, ids = render_scene(verts, faces, fvu, fmap, maps, table, R, T, model.renderer_env.cameras.intrinsics(), (400, 400), 0.0, 1, z_clip=0.001, return_ids=True)
F_ = faces.shape[0]
sl = ids[:, 0].long()
print('slots >= F:', int((sl >= F_).sum()), 'of', sl.numel(), ' empty:', int((sl < 0).sum()))
f = sl.clamp(min=0) % F_
nbp = torch.gather(nb, 1, f.reshape(4, -1))
print('pixels by n_behind of their face:', [int((nbp == i).sum()) for i in range(4)])
print('verts z range view0', float(zv[0].min()), float(zv[0].max()), 'ground faces visible px', int((f >= 320).sum()))
