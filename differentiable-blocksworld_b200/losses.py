"""Parameter-only regularisers of the DBW objective (what /root/reference src/model/dbw.py:373-405 adds to the pixel
terms) as standalone functions: they never touch an image, so they stay plain PyTorch on the leaf parameters -- tiny,
view-independent, and (in a multi-GPU step) identical on every rank (parallel.py divides them by the world size)."""
import torch

from . import geometry as G

OVERLAP_N_POINTS = 1000        # samples per block for the overlap term (dbw.py:33)
OVERLAP_N_BLOCKS = 1.95        # more than ~2 blocks claiming a point is penalised (dbw.py:34)
OVERLAP_TEMPERATURE = 0.005    # softness of the inside/outside decision (dbw.py:35)

TV_NORMS = {
    'l2': lambda d: torch.norm(d, dim=-1),
    'l1': lambda d: d.abs().sum(-1),
    'l2sq': lambda d: d.pow(2).sum(-1),
}


def effective_opacities(alpha_full, coarse):
    """soft opacities while they are being learned, their 0/1 decision afterwards (dbw.py:375,402)"""
    return alpha_full if coarse else (alpha_full > 0.5).float()


def parsimony(alpha_full, coarse):
    """sqrt-opacity sparsity prior on the number of blocks (dbw.py:373-376); switched off in the fine phase"""
    if not coarse:
        return alpha_full.new_zeros(())
    return G.safe_pow(effective_opacities(alpha_full, coarse), 0.5).mean()


def _tv_map(maps, norm, wrap_u=False):
    """total variation of (M,H,W,3) maps along v and u; block maps are periodic in u (their u axis closes the sphere)"""
    dv = norm(maps[:, 1:] - maps[:, :-1])
    if wrap_u:
        du = norm(torch.roll(maps, shifts=-1, dims=2) - maps)
        return du.sum(0).mean() + dv.sum(0).mean()          # summed over blocks: every map receives the same gradient scale
    du = norm(maps[:, :, 1:] - maps[:, :, :-1])
    return dv.mean() + du.mean()


def total_variation(bkg_maps, ground_maps, block_maps, norm, coarse):
    """dbw.py:378-387: background + blocks at full weight, the ground at the phase factor (1 coarse, 0.1 fine); the whole
    term is scaled by the same factor again by the caller"""
    factor = 1.0 if coarse else 0.1
    tv = _tv_map(bkg_maps, norm)
    if block_maps is not None and len(block_maps) > 0:
        tv = tv + _tv_map(block_maps, norm, wrap_u=True)
    tv = tv + _tv_map(ground_maps, norm) * factor
    return tv * factor


def overlap(S, R, T, eps1, eps2, alpha_full, ratio, coarse, generator=None):
    """dbw.py:389-405: sample points inside every block's bounding box, count (softly, opacity-weighted) how many
    superquadrics contain each of them, penalise counts above OVERLAP_N_BLOCKS.  Off in the fine phase."""
    if not coarse:
        return alpha_full.new_zeros(())
    N = S.shape[0]
    with torch.no_grad():
        pts = torch.rand(N, OVERLAP_N_POINTS, 3, device=S.device, generator=generator) * 2 - 1
        pts = torch.bmm(pts * ratio * S[:, None], R) + T[:, None]              # into the scene frame
        pts = pts.reshape(1, -1, 3).expand(N, -1, -1)
    local = torch.bmm(pts - T[:, None], R.transpose(1, 2)) / (S[:, None] * ratio)  # every point in every block's frame
    sdf = G.superquadric_implicit(local, eps1, eps2)
    occupancy = torch.sigmoid(-sdf / OVERLAP_TEMPERATURE) * effective_opacities(alpha_full, coarse)[:, None]
    return (occupancy.sum(0) - OVERLAP_N_BLOCKS).clamp(min=0).mean()
