import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dbw_path as D
from tests.helpers import scene_to_device, render_product, slots_to_clipped_idx
dev = torch.device('cuda:0')
tpl = D.SceneTemplate(n_blocks=6, txt_size=32)
p = D.init_params(6, 32, seed=3)
R, T, K = D.ring_cameras(2, jitter=0.3, seed=3)
blocks, alpha = tpl.build_blocks(p)
fa = alpha.repeat_interleave(tpl.BNF).repeat(2)
for K_ in (1, 3):
    ref, fr = D.render(blocks, R, T, K, (64, 64), sigma=1e-4, faces_per_pixel=K_, z_clip=0.001, detach_bary=True, faces_alpha=fa, return_fragments=True)
    sc = scene_to_device(blocks, dev)
    out, ids = render_product(sc, R.to(dev), T.to(dev), K, (64, 64), 1e-4, K_, z_clip=0.001, detach_bary=True, faces_alpha=fa.to(dev), return_ids=True)
    err = (out.cpu() - ref).abs()
    bad = (err > 1e-4).any(1).nonzero()
    cid = slots_to_clipped_idx(ids, fr, blocks['faces'].shape[0])
    print('K', K_, 'max err', err.max().item(), 'n bad px', len(bad), 'id mismatch entries', (cid != fr.clipped_idx).sum().item())
    # also full-K fragments to see the z's around
    _, fr10 = D.render(blocks, R, T, K, (64, 64), sigma=1e-4, faces_per_pixel=10, z_clip=0.001, return_fragments=True)
    for b, y, x in bad.tolist()[:5]:
        print(' px', b, y, x, 'oracle', fr.pix_to_face[b, y, x].tolist(), 'cuda', cid[b, y, x].tolist(), 'out', out[b, :, y, x].tolist(), 'ref', ref[b, :, y, x].tolist())
        print('   z10', [f'{v:.9g}' for v in fr10.zbuf[b, y, x].tolist()], 'f10', fr10.pix_to_face[b, y, x].tolist(), 'd10', [f'{v:.4g}' for v in fr10.dists[b, y, x].tolist()])
