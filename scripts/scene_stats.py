"""Work statistics of the bench scene (cfg 2) from the oracle's projection: how many (pixel, face) pairs each stage of the
forward rasterizer sees.  Guides kernel design; not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dbw_path as D, pt3d

def stats(n_blocks=10, txt=16, H=400, W=400, sigma=1e-4, views=(0, 7, 20), tile=(16, 8), K=10, boxy=False, nv=49):
    tpl = D.SceneTemplate(n_blocks=n_blocks, txt_size=txt)
    p = D.init_params(n_blocks, txt, seed=227391, boxy=boxy)
    R, T, Km = D.ring_cameras(nv)
    blocks, alpha = tpl.build_blocks(p)
    ndc = pt3d.world_to_ndc(blocks['verts'], R, T, Km).numpy()
    faces = blocks['faces'].numpy()
    blur = np.log(1 / 1e-4 - 1) * sigma
    r = np.sqrt(blur)
    s = min(H, W)
    xs = -((2 * (W - 1 - np.arange(W)) + 1) / W - 1) * (W / s) * -1
    # pixel centres in NDC (+x left): x = ndc(W-1-j)
    def ndc1(i, S1, S2):
        rng = 2.0 * (S1 / S2 if S1 > S2 else 1.0)
        return -rng / 2 + (rng * i + rng / 2) / S1
    px = np.array([ndc1(W - 1 - j, W, H) for j in range(W)])
    py = np.array([ndc1(H - 1 - i, H, W) for i in range(H)])
    out = []
    for v in views:
        tri = ndc[v][faces]                       # F,3,3
        ok = tri[:, :, 2].min(1) > 1e-8
        tri = tri[ok]
        x0, x1 = tri[:, :, 0].min(1) - r, tri[:, :, 0].max(1) + r
        y0, y1 = tri[:, :, 1].min(1) - r, tri[:, :, 1].max(1) + r
        inbb = ((px[None, None, :] >= x0[:, None, None]) & (px[None, None, :] <= x1[:, None, None]) &
                (py[None, :, None] >= y0[:, None, None]) & (py[None, :, None] <= y1[:, None, None]))    # F,H,W
        # exact test: inside or dist<blur
        a, b, c = tri[:, 0, :2], tri[:, 1, :2], tri[:, 2, :2]
        P = np.stack(np.meshgrid(px, py), -1)     # H,W,2
        def edge(p, a, b):
            return (p[..., 0] - a[:, None, None, 0]) * (b[:, None, None, 1] - a[:, None, None, 1]) - (p[..., 1] - a[:, None, None, 1]) * (b[:, None, None, 0] - a[:, None, None, 0])
        area = (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]) - (c[:, 1] - a[:, 1]) * (b[:, 0] - a[:, 0])
        e0, e1, e2 = edge(P[None], b, c), edge(P[None], c, a), edge(P[None], a, b)
        sg = np.sign(area)[:, None, None]
        inside = (e0 * sg > 0) & (e1 * sg > 0) & (e2 * sg > 0)
        def segd(p, a, b):
            ba = b - a; l2 = (ba ** 2).sum(-1)
            t = ((p[None] - a[:, None, None]) * ba[:, None, None]).sum(-1) / l2[:, None, None]
            t = np.clip(t, 0, 1)
            q = a[:, None, None] + t[..., None] * ba[:, None, None]
            return ((q - p[None]) ** 2).sum(-1)
        d = np.minimum(np.minimum(segd(P, a, b), segd(P, a, c)), segd(P, b, c))
        hit = inbb & (inside | (d < blur))
        cand = inbb.sum(0); kept = hit.sum(0)
        TH, TW = tile[1], tile[0]
        nty, ntx = H // TH, W // TW
        inbb_t = inbb.reshape(-1, nty, TH, ntx, TW)
        tlist = inbb_t.any((2, 4)).sum(0)          # faces listed per tile (approx: bbox contains some pixel centre)
        # warp patches 8x4
        ib = inbb.reshape(-1, H // 4, 4, W // 8, 8)
        wl = ib.any((2, 4)).sum(0)
        hb = hit.reshape(-1, H // 4, 4, W // 8, 8)
        wh = hb.any((2, 4)).sum(0)
        wmaxkept = kept.reshape(H // 4, 4, W // 8, 8).max((1, 3))
        wmaxcand = cand.reshape(H // 4, 4, W // 8, 8).max((1, 3))
        act = cand > 0
        out.append(dict(view=v, faces=int(ok.sum()), px_with_cand=float(act.mean()), cand_per_px=float(cand.mean()),
                        cand_per_active_px=float(cand[act].mean()), kept_per_px=float(kept.mean()),
                        kept_per_active=float(kept[kept > 0].mean()), px_kept_frac=float((kept > 0).mean()),
                        kept_gtK=float((kept > K).mean()), tile_list_mean=float(tlist.mean()), tile_list_nonempty=float(tlist[tlist > 0].mean()),
                        tile_list_max=int(tlist.max()), warp_list=float(wl.mean()), warp_hitfaces=float(wh.mean()),
                        warp_maxkept=float(wmaxkept.mean()), warp_maxcand=float(wmaxcand.mean()), kept_max=int(kept.max())))
    return out

if __name__ == '__main__':
    for o in stats():
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in o.items()})
