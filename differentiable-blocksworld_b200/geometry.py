"""Topology / UV builders and the small closed-form math of the scene model (run once at init or on (N,42) tensors):
what the reference gets from src/utils/mesh.py:78-169,210-211, src/utils/superquadric.py:10-38,
src/utils/pytorch.py:31-36, src/model/tools.py:173-207 and from PyTorch3D's ico_sphere / SubdivideMeshes /
rotation_6d_to_matrix.  Vertex and face ORDER follows PyTorch3D so that the
sq_eta / sq_omega / block_*_uvs buffers of a reference checkpoint stay meaningful."""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT_EPS = 1e-6

_ICO_A, _ICO_B = 0.5257, 0.8507          # published level-0 icosahedron constants (not re-normalised at level 0)
_ICO_VERTS = [[-_ICO_A, _ICO_B, 0], [_ICO_A, _ICO_B, 0], [-_ICO_A, -_ICO_B, 0], [_ICO_A, -_ICO_B, 0],
              [0, -_ICO_A, _ICO_B], [0, _ICO_A, _ICO_B], [0, -_ICO_A, -_ICO_B], [0, _ICO_A, -_ICO_B],
              [_ICO_B, 0, -_ICO_A], [_ICO_B, 0, _ICO_A], [-_ICO_B, 0, -_ICO_A], [-_ICO_B, 0, _ICO_A]]
_ICO_FACES = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
              [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
              [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]


def subdivide_mesh(verts, faces):
    """Loop-style 1->4 split: a new vertex at the midpoint of every unique edge (appended in sorted-edge order),
    sub-faces emitted as [corner0..], [corner1..], [corner2..], [centre] blocks."""
    V, Fn = verts.shape[0], faces.shape[0]
    a, b, c = faces.unbind(1)
    pairs = torch.cat([torch.stack([b, c], 1), torch.stack([c, a], 1), torch.stack([a, b], 1)], 0)
    lo, hi = pairs.min(1)[0], pairs.max(1)[0]
    keys, inv = torch.unique(lo * V + hi, return_inverse=True)
    mid = 0.5 * (verts[keys // V] + verts[keys % V])
    e_bc, e_ca, e_ab = (inv.view(3, Fn) + V).unbind(0)
    new_faces = torch.cat([torch.stack([a, e_ab, e_ca], 1), torch.stack([b, e_bc, e_ab], 1),
                           torch.stack([c, e_ca, e_bc], 1), torch.stack([e_bc, e_ca, e_ab], 1)], 0)
    return torch.cat([verts, mid], 0), new_faces


def ico_sphere(level=0):
    verts = torch.tensor(_ICO_VERTS, dtype=torch.float32)
    faces = torch.tensor(_ICO_FACES, dtype=torch.long)
    for _ in range(level):
        verts, faces = subdivide_mesh(verts, faces)
        verts = verts / verts.norm(p=2, dim=1, keepdim=True)
    return verts, faces


def spherical_uv(X, eps=1e-7):
    """UV in [0,1]: u from the azimuth around +Y (atan2(x, z)), v from the inclination measured from -Y."""
    r = X.norm(dim=-1).clamp(min=eps)
    y = (X[..., 1] / r).clamp(-1 + eps, 1 - eps)
    u = (torch.atan2(X[..., 0], X[..., 2]) + np.pi) / (2 * np.pi)
    v = torch.acos(-y) / np.pi
    return torch.stack([u, v], dim=-1)


def icosphere_uvs(level, eps=1e-8):
    """(faces_uvs (F,3), verts_uvs (Vt,2)) with the two fixes of the reference (mesh.py:127-169): faces crossing the
    u seam get a duplicated UV vertex shifted by +-1; pole faces get a private pole UV vertex at the mean u of their two
    other corners."""
    verts, faces = ico_sphere(level)
    uvs = spherical_uv(verts)
    faces = faces.clone()
    # seam
    fu = uvs[faces][..., 0]
    wraps = (fu - fu.roll(-1, dims=1)).abs().max(1)[0] > 0.5
    sub = fu[wraps]
    sgn = torch.sign(sub - 0.5 + eps)
    major = sgn.sum(1, keepdim=True)
    lone = sgn != major
    rows = faces[wraps]
    new_uv = torch.stack([(sub + major * lone)[lone], uvs[rows][..., 1][lone]], -1)
    rows[lone] = len(uvs) + torch.arange(int(lone.sum()))
    faces[wraps] = rows
    uvs = torch.cat([uvs, new_uv], 0)
    # poles
    fuv = uvs[faces]
    fv = fuv[..., 1]
    touches = (fv.max(1)[0] > 0.99) | (fv.min(1)[0] < 0.01)
    sub_v = fv[touches]
    at_pole = (sub_v > 0.99) | (sub_v < 0.01)
    u_mid = (fuv[touches][..., 0] * (~at_pole).float()).sum(1) / 2
    rows = faces[touches]
    new_uv = torch.stack([u_mid, sub_v[at_pole]], -1)
    rows[at_pole] = len(uvs) + torch.arange(int(at_pole.sum()))
    faces[touches] = rows
    uvs = torch.cat([uvs, new_uv], 0)
    return faces, uvs


def unit_plane():
    """primitives/plane.obj of the reference: the y=0 square [-1,1]^2, two triangles."""
    verts = torch.tensor([[1., 0., -1.], [1., 0., 1.], [-1., 0., 1.], [-1., 0., -1.]])
    faces = torch.tensor([[3, 1, 0], [3, 2, 1]], dtype=torch.long)
    return verts, faces


def load_obj(path):
    """(verts (V,3) float32, faces (F,3) int64) of a Wavefront OBJ: `v x y z` and `f a[/t[/n]] b.. c..` records, polygons
    fan-triangulated -- what the reference reads its primitives with (load_objs_as_meshes(load_textures=False),
    src/utils/mesh.py:173,211)."""
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                verts.append([float(t) for t in tok[1:4]])
            elif tok[0] == 'f':
                idx = [int(t.split('/')[0]) for t in tok[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                faces += [[idx[0], idx[k], idx[k + 1]] for k in range(1, len(idx) - 1)]
    return torch.tensor(verts, dtype=torch.float32), torch.tensor(faces, dtype=torch.long)


def unit_cube():
    """primitives/cube.obj of the reference (src/utils/mesh.py:172-173): the cube [-1,1]^3, 8 vertices, 12 triangles, in
    the file's vertex / face order (bottom, top, +x, +z, -x, -z; first triangles of the six quads, then the second ones)."""
    verts = torch.tensor([[1., -1., -1.], [1., -1., 1.], [-1., -1., 1.], [-1., -1., -1.],
                          [1., 1., -1.], [1., 1., 1.], [-1., 1., 1.], [-1., 1., -1.]])
    faces = torch.tensor([[1, 3, 0], [7, 5, 4], [4, 1, 0], [5, 2, 1], [2, 7, 3], [0, 7, 4],
                          [1, 2, 3], [7, 6, 5], [4, 5, 1], [5, 6, 2], [2, 6, 7], [0, 3, 7]], dtype=torch.long)
    return verts, faces


def cube_uvs():
    """(faces_uvs (12,3), verts_uvs (14,2)) of the cube primitive (src/utils/mesh.py:176-207): the six faces unfolded as a
    cross in a 4x8 grid of the unit square (four side quads in a row at v in [3/8, 5/8], top and bottom above / below the
    second one)."""
    q, e = 1 / 4, 1 / 8
    verts_uvs = torch.tensor([[0, 3 * e], [0, 5 * e], [q, 5 * e], [q, 3 * e], [3 * q, 3 * e], [3 * q, 5 * e], [2 * q, 5 * e],
                              [2 * q, 3 * e], [1, 3 * e], [1, 5 * e], [q, 7 * e], [2 * q, 7 * e], [q, e], [2 * q, e]], dtype=torch.float32)
    faces_uvs = torch.tensor([[1, 3, 0], [7, 5, 4], [4, 9, 8], [11, 2, 10], [2, 7, 3], [12, 7, 13],
                              [1, 2, 3], [7, 6, 5], [4, 5, 9], [11, 6, 2], [2, 6, 7], [12, 3, 7]], dtype=torch.long)
    return faces_uvs, verts_uvs


def _rot(axis, deg):
    a = math.radians(float(deg))
    R = torch.eye(3)
    if axis == 'elev':        # about X, angle with +Z in the YZ plane (sign flipped)
        c, s = math.cos(-a), math.sin(-a)
        R[1, 1], R[1, 2], R[2, 1], R[2, 2] = c, s, -s, c
    elif axis == 'azim':      # about Y
        c, s = math.cos(a), math.sin(a)
        R[0, 0], R[0, 2], R[2, 0], R[2, 2] = c, s, -s, c
    else:                     # roll, about Z
        c, s = math.cos(a), math.sin(a)
        R[0, 0], R[0, 1], R[1, 0], R[1, 1] = c, s, -s, c
    return R


def euler_world_rotation(elev, azim, roll):
    return (_rot('elev', elev) @ _rot('azim', azim) @ _rot('roll', roll))[None]


def rotation_6d_to_matrix(d6):
    x, y = d6[..., :3], d6[..., 3:]
    e1 = F.normalize(x, dim=-1)
    e2 = F.normalize(y - (e1 * y).sum(-1, keepdim=True) * e1, dim=-1)
    return torch.stack((e1, e2, torch.cross(e1, e2, dim=-1)), dim=-2)


def matrix_to_rotation_6d(M):
    return M[..., :2, :].clone().reshape(*M.shape[:-2], 6)


def random_rotations(n, generator=None):
    q = F.normalize(torch.randn(n, 4, generator=generator), dim=-1)
    r, i, j, k = q.unbind(-1)
    s = 2.0
    M = torch.stack([1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)], -1)
    return M.reshape(n, 3, 3)


def signed_pow(t, e):
    return torch.sign(t) * torch.abs(t).pow(e)


def safe_pow(t, e, eps=SQRT_EPS):
    return t.clamp(eps).pow(e)


def superquadric_points(eta, omega, eps1, eps2):
    ce, se = signed_pow(torch.cos(eta), eps1), signed_pow(torch.sin(eta), eps1)
    co, so = signed_pow(torch.cos(omega), eps2), signed_pow(torch.sin(omega), eps2)
    return torch.stack([ce * so, se, ce * co], dim=-1)


def superquadric_implicit(points, eps1, eps2):
    """as_sdf=2 variant used by the overlap regulariser (dbw.py:400, superquadric.py:17-38)."""
    pts = points.clamp(-5, 5)
    x2, y2, z2 = pts[..., 0] ** 2, pts[..., 1] ** 2, pts[..., 2] ** 2
    x, y, z = safe_pow(x2, 1 / eps2), safe_pow(y2, 1 / eps1), safe_pow(z2, 1 / eps2)
    res = safe_pow(x + z, eps2 / eps1) + y
    return safe_pow(res, eps1 / 2) - 1


def fancy_cmap():
    """The reference's block colour map (src/utils/plot.py:77-87): gold followed by seaborn's 21-colour 'hls' palette
    rotated by three entries, as a 256-level linearly interpolated lookup table.  Restated without seaborn/matplotlib:
    hls_palette(21) = hls_to_rgb(h, l=0.6, s=0.65) for h = (i/21 + 0.01) mod 1; 'gold' = (1, 0.843137, 0)."""
    import colorsys
    hues = [(i / 21.0 + 0.01) % 1.0 for i in range(21)]
    hls = [colorsys.hls_to_rgb(h, 0.6, 0.65) for h in hues]
    anchors = np.array([(1.0, 215.0 / 255.0, 0.0)] + hls[3:] + hls[:2])
    xs = np.linspace(0.0, 1.0, len(anchors))
    lut = np.stack([np.interp(np.linspace(0.0, 1.0, 256), xs, anchors[:, c]) for c in range(3)], axis=1)

    def cmap(values):
        if isinstance(values, torch.Tensor):
            values = values.detach().cpu().numpy()
        v = np.asarray(values, dtype=np.float64)
        idx = np.clip((v * 256).astype(np.int64), 0, 255)
        idx[v == 1.0] = 255
        return lut[idx]
    return cmap
