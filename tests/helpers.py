"""Shared test plumbing: move an oracle scene (dict of CPU tensors) to the device layout the C-ABI consumes and
call the product renderer on it."""
import torch


def scene_to_device(scene, dev, requires_grad=False):
    maps = [m.detach().to(dev).float().contiguous() for m in scene['maps']]
    table, off = [], 0
    for m in maps:
        table.append((off, m.shape[0], m.shape[1])); off += m.numel()
    flat = torch.cat([m.reshape(-1) for m in maps])
    out = {
        'verts': scene['verts'].detach().to(dev).float().contiguous(),
        'faces': scene['faces'].to(dev).to(torch.int32).contiguous(),
        'faces_uvs': scene['faces_verts_uvs'].detach().to(dev).float().contiguous(),
        'face_map': scene['face_map'].to(dev).to(torch.int32).contiguous(),
        'maps': flat, 'table': table,
    }
    if requires_grad:
        out['verts'].requires_grad_(True)
        out['maps'].requires_grad_(True)
    return out


def intrinsics(K):
    K = K.reshape(-1, 4, 4)[0]
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def render_product(sc, R, T, K, image_size, sigma, faces_per_pixel, z_clip=None, detach_bary=False, faces_alpha=None,
                   clip_inside=True, background=(0., 0., 0.), return_ids=False, verts_are_ndc=False):
    from dbw_b200 import render_scene
    return render_scene(sc['verts'], sc['faces'], sc['faces_uvs'], sc['face_map'], sc['maps'], sc['table'],
                        R.float() if R is not None else None, T.float() if T is not None else None, intrinsics(K),
                        image_size, sigma, faces_per_pixel, z_clip=z_clip, detach_bary=detach_bary,
                        clip_inside=clip_inside, background=background, faces_alpha=faces_alpha,
                        return_ids=return_ids, verts_are_ndc=verts_are_ndc)


def split_map_grads(flat_grad, table):
    return [flat_grad[o:o + h * w * 3].reshape(h, w, 3) for o, h, w in table]


def slots_to_clipped_idx(ids, frags, n_faces):
    """Map the CUDA kernel's per-view face slots (B,K,H,W) (slot < F: a face's first triangle, slot >= F: the second
    triangle of a z-clipped quad) to the oracle's index into its CLIPPED, batch-packed face list, layout (B,H,W,K)."""
    ids = ids.cpu().long().permute(0, 2, 3, 1)
    B = ids.shape[0]
    valid = ids >= 0
    f = ids.clamp(min=0) % n_faces
    second = (ids >= n_faces).long()
    packed = f + torch.arange(B)[:, None, None, None] * n_faces
    if frags.unclipped_to_clipped is not None:
        packed = frags.unclipped_to_clipped[packed] + second
    return torch.where(valid, packed, ids)


def decision_mask(ids, frags, n_faces):
    """(B,1,H,W) float mask of the pixels where the CUDA path and the oracle kept the same z-sorted faces."""
    same = (slots_to_clipped_idx(ids, frags, n_faces) == frags.clipped_idx).all(-1)
    return same[:, None].to(frags.zbuf.dtype)
