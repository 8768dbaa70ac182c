"""Drop-in for the reference's scene model on the render hot path: `DifferentiableBlocksWorld`
(src/model/dbw.py:38-462).  Same constructor kwargs (configs/*.yml parse unmodified), same parameter / buffer names
(checkpoints and the texture-prefixed Adam group of src/optimizer.py:9-14 keep working), same
`forward(inp, labels) -> dict of losses` / `predict(inp, labels)` signatures, so src/trainer.py:137-147 can drive it.
Rendering goes through the B200 kernels (renderer.py); compositing + RGB loss are one fused kernel when possible.

Visualisation helpers of SURVEY.md section 8f are covered as far as the trainer's periodic logging needs them (predict(w_edges=True),
predict_synthetic, get_arranged_block_txt); qualitative_eval / video / OBJ export remain out of scope."""
import ctypes
from collections import OrderedDict
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, geometry as G
from .renderer import Renderer, render_scene, _c, _stream
from .scene_ops import scene_geometry, texture_atlas
from .structures import Meshes, TexturesUV, join_meshes_as_scene, join_meshes_as_batch

DECIMATE_FACTOR = 8
OVERLAP_N_POINTS = 1000
OVERLAP_N_BLOCKS = 1.95
OVERLAP_TEMPERATURE = 0.005
DIRECTION_LIGHT = [1, 0.25, -1]

tv_norm_funcs = {'l2': lambda x: torch.norm(x, dim=-1), 'l1': lambda x: x.abs().sum(-1), 'l2sq': lambda x: (x ** 2).sum(-1)}


class _CompositeMSE(torch.autograd.Function):
    """rec = fg_rgb * fg_a + (1 - fg_a) * env_rgb ; loss = mean((imgs - rec)^2)   (dbw.py:223 + :366-367): one kernel
    forward, one kernel backward (which also folds in a gradient arriving at `rec` from other loss terms)."""

    @staticmethod
    def forward(ctx, fg, env, imgs, n_total_views):
        ctx.set_materialize_grads(False)          # g_rec arrives as None when `rec` is unused: no zero tensors, no sync
        B, _, H, W = fg.shape
        fg, env, imgs = fg.contiguous(), env.contiguous(), imgs.contiguous().float()
        rec = torch.empty(B, 3, H, W, device=fg.device, dtype=torch.float32)
        loss = torch.zeros((), device=fg.device, dtype=torch.float32)
        inv = 1.0 / (float(n_total_views) * 3 * H * W)
        _lib.check(_lib.lib().dbw_composite_mse(B, H, W, _c(fg), _c(env), _c(imgs), ctypes.c_float(inv), _c(rec), _c(loss),
                                                None, None, _stream()), 'dbw_composite_mse')
        ctx.save_for_backward(fg, env, imgs)
        ctx.inv = inv
        return rec, loss

    @staticmethod
    def backward(ctx, g_rec, g_loss):
        fg, env, imgs = ctx.saved_tensors
        B, _, H, W = fg.shape
        g_fg, g_env = torch.empty_like(fg), torch.empty_like(env)
        gl = g_loss.contiguous().float() if g_loss is not None else None
        gr = g_rec.contiguous().float() if g_rec is not None else None
        _lib.check(_lib.lib().dbw_composite_mse_backward(B, H, W, _c(fg), _c(env), _c(imgs), ctypes.c_float(ctx.inv), _c(gl),
                                                         _c(gr), _c(g_fg), _c(g_env), _stream()), 'dbw_composite_mse_backward')
        return g_fg, g_env, None, None


class DifferentiableBlocksWorld(nn.Module):
    name = 'dbw'

    def __init__(self, img_size, **kwargs):
        super().__init__()
        self._init_kwargs = deepcopy(kwargs)
        self._init_kwargs['img_size'] = img_size
        self._init_blocks(**kwargs.get('mesh', {}))
        self._init_renderer(img_size, **kwargs.get('renderer', {}))
        self._init_rend_optim(**kwargs.get('rend_optim', {}))
        self._init_loss(**kwargs.get('loss', {}))
        self.cur_epoch = 0
        # data-parallel context (parallel.py): this rank renders `len(inp['imgs'])` of `n_total_views` views
        self.n_total_views = None
        self.noise_generator = None
        # static topology (default): blocks dropped by the opacity filters (dbw.py:316-328) keep their slot in the mesh
        # and are DISABLED through face_map = -1 instead of being sliced out -> identical images and gradients, but no
        # host sync and fixed shapes, so the whole step can be captured in a CUDA graph (graph.py)
        self.static_topology = True
        self.opacity_noise_buffer = None          # optional pre-drawn randn (N,) used instead of drawing inside forward
        self._static_arrays = None
        # fused scene construction (scene_ops.py): mesh build + texture prep as one kernel each way
        self.fused_scene = True
        self._fused_static = None
        self.overlap_passes = False               # environment pass on a side stream, concurrent with the blocks pass
        self._env_stream = None

    @property
    def init_kwargs(self):
        return deepcopy(self._init_kwargs)

    # ------------------------------------------------------------------ construction (dbw.py:55-163)
    def _init_blocks(self, **kwargs):
        self.n_blocks = kwargs.pop('n_blocks', 1)
        self.S_world = kwargs.pop('S_world', 1)
        elev, azim, roll = kwargs.pop('R_world', [0, 0, 0])
        self.register_buffer('R_world', G.euler_world_rotation(elev, azim, roll))
        self.register_buffer('T_world', torch.Tensor(kwargs.pop('T_world', [0., 0., 0.]))[None])
        self.z_far = kwargs.pop('z_far', 10)
        self.ratio_block_scene = kwargs.pop('ratio_block_scene', 1 / 4)
        self.txt_size = kwargs.pop('txt_size', 256)
        self.txt_bkg_upscale = kwargs.pop('txt_bkg_upscale', 1)
        self.scale_min = kwargs.pop('scale_min', 0.2)
        opacity_init = kwargs.pop('opacity_init', 0.5)
        T_range = kwargs.pop('T_range', [1, 1, 1])
        T_init_mode = kwargs.pop('T_init_mode', 'gauss')
        assert len(kwargs) == 0, kwargs

        # spherical background (faces flipped to look inward) and planar ground
        bv, bf = G.ico_sphere(2)
        self.bkg = Meshes((bv * self.z_far)[None], bf.flip(1)[None])
        self.register_buffer('bkg_verts_uvs', G.spherical_uv(self.bkg.verts_packed()))
        gv, gf = G.unit_plane()
        gv = gv * torch.Tensor([self.z_far, 1, self.z_far])[None]
        for _ in range(3):
            gv, gf = G.subdivide_mesh(gv, gf)
        self.ground = Meshes(gv[None], gf[None])
        self.register_buffer('ground_verts_uvs', (gv[:, [0, 2]] / self.z_far + 1) / 2)

        # primitive blocks
        sv, sf = G.ico_sphere(1)
        N = self.n_blocks
        self.blocks = Meshes((sv * self.ratio_block_scene)[None].repeat(N, 1, 1), sf[None].repeat(N, 1, 1))
        self.sq_eps = nn.Parameter(torch.zeros(N, 2))
        verts = self.blocks.verts_padded() / self.ratio_block_scene
        self.register_buffer('sq_eta', torch.asin(verts[..., 1]))
        self.register_buffer('sq_omega', torch.atan2(verts[..., 0], verts[..., 2]))
        faces_uvs, verts_uvs = G.icosphere_uvs(1)
        p_left = abs(int(np.floor(verts_uvs.min(0)[0][0].item() * self.txt_size)))
        p_right = int(np.ceil((verts_uvs.max(0)[0][0].item() - 1) * self.txt_size))
        verts_u = (verts_uvs[..., 0] * self.txt_size + p_left) / (self.txt_size + p_left + p_right)
        self.txt_padding = p_left, p_right
        self.BNF = len(faces_uvs)
        self.register_buffer('block_faces_uvs', faces_uvs)
        self.register_buffer('block_verts_uvs', torch.stack([verts_u, verts_uvs[..., 1]], dim=-1))

        # learnable poses
        self.R_6d_ground = nn.Parameter(torch.Tensor([[1., 0., 0., 0., 1., 0.]]))
        self.T_ground = nn.Parameter(torch.Tensor([[0., -0.9 * T_range[1], 0.]]))
        S_init = (torch.rand(N, 3) + 0.5 - self.scale_min).log()
        R_6d_init = G.matrix_to_rotation_6d(G.random_rotations(N))
        if T_init_mode == 'gauss':
            T_init = torch.randn(N, 3) / 2 * torch.Tensor(T_range)
        elif T_init_mode == 'uni':
            T_init = (2 * torch.rand(N, 3) - 1) * torch.Tensor(T_range)
        else:
            raise NotImplementedError
        self.S = nn.Parameter(S_init.clone())
        self.R_6d = nn.Parameter(R_6d_init.clone())
        self.T = nn.Parameter(T_init.clone())

        # learnable opacities and textures
        self.alpha_logit = nn.Parameter(torch.logit(torch.ones(N) * opacity_init) + 1e-3)
        TS, s = self.txt_size, self.txt_bkg_upscale
        self.texture_bkg = nn.Parameter(torch.randn(1, TS * s, TS * s, 3) / 10)
        self.texture_ground = nn.Parameter(torch.randn(1, TS * s, TS * s, 3) / 10)
        self.textures = nn.Parameter(torch.randn(N, TS, TS, 3) / 10)

    def _init_rend_optim(self, **kwargs):
        self.opacity_noise = kwargs.pop('opacity_noise', False)
        self.decouple_rendering = kwargs.pop('decouple_rendering', False)
        self.coarse_learning = kwargs.pop('coarse_learning', True)
        self.decimate_txt = kwargs.pop('decimate_txt', False)
        self.decim_factor = kwargs.pop('decimate_factor', DECIMATE_FACTOR)
        self.kill_blocks = kwargs.pop('kill_blocks', False)
        assert len(kwargs) == 0, kwargs

    def _init_renderer(self, img_size, **kwargs):
        kwargs = deepcopy(kwargs)
        self.renderer = Renderer(img_size, **deepcopy(kwargs))
        kwargs['sigma'] = 5e-6
        self.renderer_fine = Renderer(img_size, **deepcopy(kwargs))
        kwargs['faces_per_pixel'] = 1
        kwargs['sigma'] = 0
        kwargs['detach_bary'] = False
        self.renderer_env = Renderer(img_size, **deepcopy(kwargs))
        kwargs['lights'] = {'name': 'directional', 'direction': [DIRECTION_LIGHT], 'ambient_color': [[0.7, 0.7, 0.7]],
                            'diffuse_color': [[0.4, 0.4, 0.4]], 'specular_color': [[0., 0., 0.]]}
        kwargs['shading_type'] = 'flat'
        kwargs['background_color'] = (1, 1, 1)
        self.renderer_light = Renderer(img_size, **deepcopy(kwargs))

    def _init_loss(self, **kwargs):
        loss_weights = {
            'rgb': kwargs.pop('rgb_weight', 1.0),
            'perceptual': kwargs.pop('perceptual_weight', 0),
            'parsimony': kwargs.pop('parsimony_weight', 0),
            'scale': kwargs.pop('scale_weight', 0),
            'tv': kwargs.pop('tv_weight', 0),
            'overlap': kwargs.pop('overlap_weight', 0),
        }
        name = kwargs.pop('name', 'mse')
        perceptual_name = kwargs.pop('perceptual_name', 'lpips')
        self.tv_norm = tv_norm_funcs[kwargs.pop('tv_type', 'l2sq')]
        assert len(kwargs) == 0, kwargs
        self.loss_weights = {k: v for k, v in loss_weights.items() if v > 0}
        self.loss_names = [f'loss_{n}' for n in list(self.loss_weights.keys()) + ['total']]
        self.loss_name = name
        self.criterion = {'mse': nn.MSELoss, 'l2': nn.MSELoss, 'l1': nn.L1Loss}[name]()
        if 'perceptual' in self.loss_weights:
            # LPIPS / VGG perceptual terms are dense-conv networks served by cuDNN: outside the render hot path
            # (SURVEY 8a row a13).  Plug any callable (imgs, rec) -> scalar via set_perceptual_loss().
            self.perceptual_loss = _make_perceptual(perceptual_name)

    def set_perceptual_loss(self, fn):
        self.perceptual_loss = fn

    def set_cur_epoch(self, epoch):
        self.cur_epoch = epoch

    def step(self):
        self.cur_epoch += 1

    def to(self, device):
        super().to(device)
        self.bkg, self.ground, self.blocks = self.bkg.to(device), self.ground.to(device), self.blocks.to(device)
        for r in (self.renderer, self.renderer_fine, self.renderer_env, self.renderer_light):
            r.to(device)
        return self

    @property
    def bkg_n_faces(self):
        return self.bkg.num_faces_per_mesh().sum().item()

    @property
    def ground_n_faces(self):
        return self.ground.num_faces_per_mesh().sum().item()

    @property
    def env_n_faces(self):
        return self.bkg_n_faces + self.ground_n_faces

    @property
    def blocks_n_faces(self):
        return self.blocks.num_faces_per_mesh().sum().item()

    def is_live(self, name):
        milestone = getattr(self, name)
        if isinstance(milestone, bool):
            return milestone
        return True if self.cur_epoch < milestone else False

    # ------------------------------------------------------------------ forward (dbw.py:198-239)
    def forward(self, inp, labels=None):
        layers = self._render_layers(inp)
        imgs = inp['imgs']
        n_total = self.n_total_views or len(imgs)
        if layers[1] is not None and isinstance(self.criterion, nn.MSELoss) and imgs.is_cuda:
            rec, mse = _CompositeMSE.apply(layers[1], layers[0], imgs, n_total)
            return self.compute_losses(imgs, rec, rgb_loss=mse)
        rec = self._composite(layers)
        return self.compute_losses(imgs, rec, rgb_loss=self.criterion(imgs, rec) * (len(imgs) / float(n_total)))

    def _install_cameras(self, inp):
        if 'K' in inp and self.renderer.cameras.K is None:
            for r in (self.renderer, self.renderer_fine, self.renderer_env, self.renderer_light):
                r.update_cameras(device=inp['imgs'].device, K=inp['K'][0:1])
                r.cameras.intrinsics()          # the one host read of K happens here, never inside a step

    def _render_layers(self, inp, filter_transparent=False):
        """(env RGBA, blocks RGBA or None) in decoupled mode, (scene RGBA, None) in joint mode."""
        B, R_tgt, T_tgt = len(inp['imgs']), inp['R'], inp['T']
        self._install_cameras(inp)
        fine_learning = not self.is_live('coarse_learning')
        filter_tsp = filter_transparent or fine_learning
        renderer = self.renderer_fine if fine_learning else self.renderer
        if self.decouple_rendering and self.static_topology and self.fused_scene and R_tgt.is_cuda:
            return self._render_layers_fused(B, R_tgt, T_tgt, filter_tsp, renderer)
        if self.decouple_rendering and self.static_topology:
            env = join_meshes_as_scene([self.build_bkg(world_coord=True), self.build_ground(world_coord=True)])
            out_env = self.renderer_env(env.extend(B), R=R_tgt, T=T_tgt)
            verts, faces, fvu, fmap, maps, table = self._blocks_scene_static(filter_tsp)
            alpha = None if filter_tsp else self._alpha.repeat_interleave(self.BNF)
            r = renderer
            out_fg = render_scene(verts, faces, fvu, fmap, maps, table, R_tgt, T_tgt, r.cameras.intrinsics(), r.img_size,
                                  r.sigma, r.faces_per_pixel, r.z_clip, r.detach_bary, r.clip_inside, r.background_color,
                                  alpha, r.perspective_correct, blur_radius=r.blur_radius)
            return out_env, out_fg
        if self.decouple_rendering:
            env = join_meshes_as_scene([self.build_bkg(world_coord=True), self.build_ground(world_coord=True)])
            out_env = self.renderer_env(env.extend(B), R=R_tgt, T=T_tgt)
            blocks = self.build_blocks(filter_transparent=filter_tsp, as_scene=True)
            if len(blocks) > 0:
                alpha = None if filter_tsp else self._alpha.repeat_interleave(self.BNF)
                out_fg = renderer(blocks.extend(B), R=R_tgt, T=T_tgt, faces_alpha=alpha)
            else:
                out_fg = torch.zeros_like(out_env)
            return out_env, out_fg
        scene = self.build_scene(filter_transparent=filter_tsp)
        if not filter_tsp:
            alpha_env = torch.ones(self.env_n_faces, device=R_tgt.device)
            alpha = torch.cat([alpha_env, self._alpha.repeat_interleave(self.BNF)], dim=0)
        else:
            alpha = None
        return renderer(scene.extend(B), R=R_tgt, T=T_tgt, faces_alpha=alpha), None

    @staticmethod
    def _composite(layers):
        first, fg = layers
        if fg is None:
            return first[:, :3]
        rec_fg, mask = fg.split([3, 1], dim=1)
        return rec_fg * mask + (1 - mask) * first[:, :3]

    def predict(self, inp, labels=None, w_edges=False, filter_transparent=False):
        rec = self._composite(self._render_layers(inp, filter_transparent))
        if w_edges:                                   # dbw.py:234-238: coloured face edges drawn over the reconstruction
            B, R_tgt, T_tgt = len(inp['imgs']), inp['R'], inp['T']
            fine_learning = not self.is_live('coarse_learning')
            filter_tsp = filter_transparent or fine_learning
            renderer = self.renderer_fine if fine_learning else self.renderer
            env = join_meshes_as_scene([self.build_bkg(world_coord=True), self.build_ground(world_coord=True)])
            blocks = self.build_blocks(filter_transparent=filter_tsp, as_scene=True)
            scene = join_meshes_as_scene([env, blocks]) if len(blocks) > 0 else env
            colors = self.get_scene_face_colors(filter_transparent=filter_tsp).repeat(B, 1)
            rec = renderer.draw_edges(rec, scene.extend(B), R_tgt, T_tgt, colors=colors)
        return rec

    def predict_synthetic(self, inp, labels=None):
        """flat-shaded render of the opaque blocks with one synthetic colour per block (dbw.py:241-248)."""
        B, R_tgt, T_tgt = len(inp['imgs']), inp['R'], inp['T']
        self._install_cameras(inp)
        blocks = self.build_blocks(filter_transparent=True, synthetic_colors=True, as_scene=True)
        if len(blocks) > 0:
            return self.renderer_light(blocks.extend(B), R=R_tgt, T=T_tgt, viz_purpose=True)[:, :3]
        return torch.ones_like(inp['imgs'])

    @torch.no_grad()
    def get_scene_face_colors(self, filter_transparent=False, w_env=True):
        """one colour per face of the scene mesh: environment faces get the first colour of the map, block k the colour at
        (k+1)/N (dbw.py:420-431)."""
        val_blocks = torch.linspace(0, 1, self.n_blocks + 1)[1:]
        if filter_transparent:
            val_blocks = val_blocks[self.get_opacities().cpu() > 0.5]
        elif self.kill_blocks:
            val_blocks = val_blocks[self.get_opacities().cpu() > 0.01]
        NFE = self.env_n_faces if w_env else 0
        values = torch.cat([torch.zeros(NFE), val_blocks.repeat_interleave(self.BNF)])
        return torch.from_numpy(G.fancy_cmap()(values.numpy())).float().to(self.bkg.device)

    # ------------------------------------------------------------------ scene construction (dbw.py:250-352)
    def _decimate(self, maps):
        sub = F.avg_pool2d(maps.permute(0, 3, 1, 2), kernel_size=self.decim_factor, stride=self.decim_factor)
        return F.interpolate(sub, scale_factor=self.decim_factor).permute(0, 2, 3, 1)

    def _to_world(self, verts):
        return (verts * self.S_world) @ self.R_world + self.T_world[:, None]

    def build_scene(self, filter_transparent=False, w_bkg=True, reduce_ground=False):
        meshes = []
        if w_bkg:
            meshes.append(self.build_bkg())
        meshes.append(self.build_ground(reduced=reduce_ground))
        blocks = self.build_blocks(filter_transparent)
        if len(blocks) > 0:
            meshes.append(blocks)
        scene = join_meshes_as_scene(meshes) if (len(meshes) - 1 + len(blocks)) > 1 else meshes[0]
        verts, faces = scene.get_mesh_verts_faces(0)
        return Meshes(self._to_world(verts[None]), faces[None], scene.textures)

    def build_bkg(self, reduced=False, world_coord=False):
        verts, faces = [t[None] for t in self.bkg.get_mesh_verts_faces(0)]
        if reduced:
            verts = verts * 3 / self.z_far
        if world_coord:
            verts = self._to_world(verts)
        maps = torch.sigmoid(self.texture_bkg)
        self._bkg_maps = maps
        if self.training and self.is_live('decimate_txt'):
            maps = self._decimate(maps)
        return Meshes(verts, faces, textures=TexturesUV(maps, faces, self.bkg_verts_uvs[None], align_corners=True))

    def build_ground(self, reduced=False, world_coord=False):
        S_ground = 1. if not reduced else torch.Tensor([3 / self.z_far, 1, 3 / self.z_far]).to(self.bkg.device)
        verts, faces = [t[None] for t in self.ground.get_mesh_verts_faces(0)]
        verts = (verts * S_ground) @ G.rotation_6d_to_matrix(self.R_6d_ground) + self.T_ground[:, None]
        if world_coord:
            verts = self._to_world(verts)
        maps = torch.sigmoid(self.texture_ground)
        self._ground_maps = maps
        if self.training and self.is_live('decimate_txt'):
            maps = self._decimate(maps)
        return Meshes(verts, faces, textures=TexturesUV(maps, faces, self.ground_verts_uvs[None], align_corners=True))

    def _fused_arrays(self):
        dev = self.alpha_logit.device
        if self._fused_static is None or self._fused_static['dev'] != dev:
            N, Vb = self.n_blocks, self.sq_eta.shape[1]
            gv, gf = self.ground.get_mesh_verts_faces(0)
            bv, bf = self.bkg.get_mesh_verts_faces(0)
            geom = {'n_blocks': N, 'verts_per_block': Vb, 'n_ground_verts': gv.shape[0], 'sq_eta': self.sq_eta.contiguous(),
                    'sq_omega': self.sq_omega.contiguous(), 'ground_verts': gv.contiguous().float(),
                    'ratio': float(self.ratio_block_scene), 'scale_min': float(self.scale_min), 'S_world': float(self.S_world),
                    'R_world': [float(x) for x in self.R_world.reshape(-1).cpu()], 'T_world': [float(x) for x in self.T_world.reshape(-1).cpu()]}
            bkg_world = self._to_world(bv[None])[0].contiguous()             # static: no learnable pose
            faces_b = (self.blocks.faces_padded() + (torch.arange(N, device=dev) * Vb)[:, None, None]).reshape(-1, 3)
            fvu_b = self.block_verts_uvs[self.block_faces_uvs][None].expand(N, -1, -1, -1).reshape(-1, 3, 2).contiguous()
            fmap_b = torch.arange(N, device=dev, dtype=torch.int32).repeat_interleave(self.BNF)
            faces_e = torch.cat([bf, gf + bv.shape[0]]).to(torch.int32).contiguous()
            fvu_e = torch.cat([self.bkg_verts_uvs[bf], self.ground_verts_uvs[gf]]).contiguous()
            fmap_e = torch.cat([torch.zeros(len(bf)), torch.ones(len(gf))]).to(dev).to(torch.int32)
            self._fused_static = {'dev': dev, 'geom': geom, 'bkg_world': bkg_world, 'faces_b': faces_b.to(torch.int32).contiguous(),
                                  'fvu_b': fvu_b, 'fmap_b': fmap_b, 'faces_e': faces_e, 'fvu_e': fvu_e, 'fmap_e': fmap_e}
        return self._fused_static

    def _render_layers_fused(self, B, R_tgt, T_tgt, filter_tsp, renderer):
        """decoupled rendering with the fused scene kernels: leaf parameters -> 2 scene kernels -> 2 render passes."""
        st = self._fused_arrays()
        N, Vb = self.n_blocks, st['geom']['verts_per_block']
        coarse_learning = self.training and self.is_live('coarse_learning')
        decim = self.decim_factor if (self.training and self.is_live('decimate_txt')) else 1
        verts = scene_geometry(self.sq_eps, self.S, self.R_6d, self.T, self.R_6d_ground, self.T_ground, st['geom'])
        # ---- environment pass (bkg sphere is static, the ground follows R_6d_ground / T_ground)
        env_verts = torch.cat([st['bkg_world'], verts[N * Vb:]])
        tb, tg = texture_atlas(self.texture_bkg, 0, 0, decim), texture_atlas(self.texture_ground, 0, 0, decim)
        env_atlas = torch.cat([tb.reshape(-1, 4), tg.reshape(-1, 4)])
        Hb = self.texture_bkg.shape[1]
        table_e = [(0, Hb, Hb), (Hb * Hb * 3, Hb, Hb)]
        re = self.renderer_env
        # the two passes are independent until compositing: the environment pass runs on a side stream so that its
        # kernels (and, through autograd, their backward) overlap the blocks pass -- each raster kernel alone leaves
        # 35-50 % of the issue slots idle (profiles/), two different ones interleave on the SMs
        cur = torch.cuda.current_stream()
        if self._env_stream is None:
            self._env_stream = torch.cuda.Stream()
        side = self._env_stream if self.overlap_passes else cur
        if side is not cur:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            out_env = render_scene(env_verts, st['faces_e'], st['fvu_e'], st['fmap_e'], env_atlas, table_e, R_tgt, T_tgt,
                                   re.cameras.intrinsics(), re.img_size, re.sigma, re.faces_per_pixel, re.z_clip, re.detach_bary,
                                   re.clip_inside, re.background_color, None, re.perspective_correct, blur_radius=re.blur_radius,
                                   maps_are_texels4=True)
        # ---- blocks pass
        alpha_logit = self.alpha_logit
        if self.opacity_noise and coarse_learning:
            alpha_logit = alpha_logit + self.opacity_noise * self._draw_opacity_noise()
        self._alpha = torch.sigmoid(alpha_logit)
        self._alpha_full = self._alpha.clone()
        fmap = st['fmap_b']
        if filter_tsp or self.kill_blocks:
            mask = torch.sigmoid(self.alpha_logit) > (0.5 if filter_tsp else 0.01)
            self._alpha_full = self._alpha_full * mask
            fmap = torch.where(mask.repeat_interleave(self.BNF), fmap, torch.full_like(fmap, -1))
        p_left, p_right = self.txt_padding
        atlas = texture_atlas(self.textures, p_left, p_right, self.decim_factor if (coarse_learning and self.is_live('decimate_txt')) else 1)
        Ht, Wt = atlas.shape[1], atlas.shape[2]
        table_b = [(i * Ht * Wt * 3, Ht, Wt) for i in range(N)]
        alpha = None if filter_tsp else self._alpha.repeat_interleave(self.BNF)
        r = renderer
        out_fg = render_scene(verts[:N * Vb], st['faces_b'], st['fvu_b'], fmap, atlas.reshape(-1, 4), table_b, R_tgt, T_tgt,
                              r.cameras.intrinsics(), r.img_size, r.sigma, r.faces_per_pixel, r.z_clip, r.detach_bary,
                              r.clip_inside, r.background_color, alpha, r.perspective_correct, blur_radius=r.blur_radius,
                              maps_are_texels4=True)
        if side is not cur:
            cur.wait_stream(side)
            out_env.record_stream(cur)
        # the regularisers of compute_losses() read these (plain torch on parameters, only built when they are used)
        self._blocks_SRT = None
        self._needs_reg_state = True
        return out_env, out_fg

    def _ensure_reg_state(self):
        """state the parameter-only regularisers read (dbw.py:313,349-351); the fused path builds it lazily."""
        if getattr(self, '_needs_reg_state', False):
            self._blocks_maps = torch.sigmoid(self.textures)
            self._bkg_maps, self._ground_maps = torch.sigmoid(self.texture_bkg), torch.sigmoid(self.texture_ground)
            self._blocks_SRT = (self.S.exp() + self.scale_min, G.rotation_6d_to_matrix(self.R_6d), self.T)
            eps1, eps2 = (self.sq_eps.sigmoid() * 1.8 + 0.1).split([1, 1], dim=-1)
            self._blocks_eps = eps1, eps2
            self._needs_reg_state = False

    def _draw_opacity_noise(self):
        if self.opacity_noise_buffer is not None:
            return self.opacity_noise_buffer
        if self.noise_generator is not None:
            return torch.randn(self.alpha_logit.shape, generator=self.noise_generator, device=self.alpha_logit.device)
        return torch.randn_like(self.alpha_logit)

    def _blocks_scene_static(self, filter_transparent):
        """build_blocks(as_scene=True) with fixed shapes: same vertices / maps / opacities, filtered blocks disabled via
        face_map = -1 (see include/dbw_render.h) instead of removed.  No host synchronisation."""
        N, dev = self.n_blocks, self.alpha_logit.device
        coarse_learning = self.training and self.is_live('coarse_learning')
        S, R, T = self.S.exp() + self.scale_min, G.rotation_6d_to_matrix(self.R_6d), self.T
        alpha_logit = self.alpha_logit
        if self.opacity_noise and coarse_learning:
            alpha_logit = alpha_logit + self.opacity_noise * self._draw_opacity_noise()
        self._alpha = torch.sigmoid(alpha_logit)
        self._alpha_full = self._alpha.clone()
        maps = torch.sigmoid(self.textures)
        verts = (self.get_blocks_verts() * S[:, None]) @ R + T[:, None]
        self._blocks_maps, self._blocks_SRT = maps, (S, R, T)
        if self._static_arrays is None or self._static_arrays[0].device != dev:
            faces = (self.blocks.faces_padded() + (torch.arange(N, device=dev) * verts.shape[1])[:, None, None]).reshape(-1, 3)
            fvu = self.block_verts_uvs[self.block_faces_uvs][None].expand(N, -1, -1, -1).reshape(-1, 3, 2).contiguous()
            fmap = torch.arange(N, device=dev, dtype=torch.int32).repeat_interleave(self.BNF)
            self._static_arrays = (faces.to(torch.int32).contiguous(), fvu, fmap)
        faces, fvu, fmap = self._static_arrays
        if filter_transparent or self.kill_blocks:
            mask = torch.sigmoid(self.alpha_logit) > (0.5 if filter_transparent else 0.01)
            self._alpha_full = self._alpha_full * mask
            fmap = torch.where(mask.repeat_interleave(self.BNF), fmap, torch.full_like(fmap, -1))
        if coarse_learning and self.is_live('decimate_txt'):
            maps = self._decimate(maps)
        p_left, p_right = self.txt_padding
        maps = F.pad(maps.permute(0, 3, 1, 2), pad=(p_left, p_right, 0, 0), mode='circular').permute(0, 2, 3, 1).contiguous()
        verts = self._to_world(verts).reshape(-1, 3)
        Ht, Wt = maps.shape[1], maps.shape[2]
        table = [(i * Ht * Wt * 3, Ht, Wt) for i in range(N)]
        return verts, faces, fvu, fmap, maps.reshape(-1), table

    def build_blocks(self, filter_transparent=False, world_coord=False, as_scene=False, synthetic_colors=False):
        coarse_learning = self.training and self.is_live('coarse_learning')
        S, R, T = self.S.exp() + self.scale_min, G.rotation_6d_to_matrix(self.R_6d), self.T
        if self.opacity_noise and coarse_learning:
            alpha_logit = self.alpha_logit + self.opacity_noise * self._draw_opacity_noise()
        else:
            alpha_logit = self.alpha_logit
        self._alpha = torch.sigmoid(alpha_logit)
        self._alpha_full = self._alpha.clone()
        maps = torch.sigmoid(self.textures)
        if synthetic_colors:
            values = torch.linspace(0, 1, self.n_blocks + 1)[1:]
            colors = torch.from_numpy(G.fancy_cmap()(values.numpy())).float().to(maps.device)
            maps = colors[:, None, None].expand(-1, self.txt_size, self.txt_size, -1)
        verts = (self.get_blocks_verts() * S[:, None]) @ R + T[:, None]
        faces = self.blocks.faces_padded()
        self._blocks_maps, self._blocks_SRT = maps, (S, R, T)

        if filter_transparent or self.kill_blocks:
            mask = torch.sigmoid(self.alpha_logit) > (0.5 if filter_transparent else 0.01)
            self._alpha_full = self._alpha_full * mask
            NB = int(mask.sum().item())
            if NB == 0:
                return Meshes([], [])
            verts, faces, maps, self._alpha = verts[mask], faces[mask], maps[mask], self._alpha[mask]
        else:
            NB = self.n_blocks

        if coarse_learning and self.is_live('decimate_txt'):
            maps = self._decimate(maps)
        p_left, p_right = self.txt_padding
        maps = F.pad(maps.permute(0, 3, 1, 2), pad=(p_left, p_right, 0, 0), mode='circular').permute(0, 2, 3, 1)
        txt = TexturesUV(maps, self.block_faces_uvs[None].expand(NB, -1, -1), self.block_verts_uvs[None].expand(NB, -1, -1),
                         align_corners=True)
        if world_coord or as_scene:
            verts = self._to_world(verts)
        blocks = Meshes(verts, faces, textures=txt)
        return join_meshes_as_scene(blocks) if as_scene else blocks

    def get_blocks_verts(self):
        eps1, eps2 = (self.sq_eps.sigmoid() * 1.8 + 0.1).split([1, 1], dim=-1)
        self._blocks_eps = eps1, eps2
        return G.superquadric_points(self.sq_eta, self.sq_omega, eps1, eps2) * self.ratio_block_scene

    # ------------------------------------------------------------------ losses (dbw.py:361-408)
    def compute_losses(self, imgs, rec, rgb_loss=None):
        losses = {k: torch.zeros((), device=imgs.device) for k in self.loss_weights}
        coarse_learning = self.is_live('coarse_learning')
        if any(k in self.loss_weights for k in ('tv', 'overlap')):
            self._ensure_reg_state()
        if 'rgb' in losses:
            losses['rgb'] = self.loss_weights['rgb'] * (rgb_loss if rgb_loss is not None else self.criterion(imgs, rec))
        if 'perceptual' in losses and self.perceptual_loss is not None:
            factor = 1 if coarse_learning else 0.1
            losses['perceptual'] = self.loss_weights['perceptual'] * factor * self.perceptual_loss(imgs, rec)
        if 'parsimony' in losses:
            factor = 1 if coarse_learning else 0
            alpha = self._alpha_full if coarse_learning else (self._alpha_full > 0.5).float()
            losses['parsimony'] = self.loss_weights['parsimony'] * factor * G.safe_pow(alpha, 0.5).mean()
        if 'tv' in losses:
            factor = 1 if coarse_learning else 0.1
            tv_loss = sum([self.tv_norm(torch.diff(self._bkg_maps, dim=k)).mean() for k in [1, 2]])
            if len(self._blocks_maps) > 0:
                dx = self.tv_norm(torch.diff(self._blocks_maps, dim=2, append=self._blocks_maps[:, :, 0:1]))
                dy = self.tv_norm(torch.diff(self._blocks_maps, dim=1))
                tv_loss += (dx.sum(0).mean() + dy.sum(0).mean())
            tv_loss += sum([self.tv_norm(torch.diff(self._ground_maps, dim=k)).mean() for k in [1, 2]]) * factor
            losses['tv'] = self.loss_weights['tv'] * factor * tv_loss
        if 'overlap' in losses:
            factor = 1 if coarse_learning else 0
            N = self.n_blocks
            with torch.no_grad():
                points = torch.rand(N, OVERLAP_N_POINTS, 3, device=rec.device, generator=self.noise_generator) * 2 - 1
                S, R, T = self._blocks_SRT
                points = (points * self.ratio_block_scene * S[:, None]) @ R + T[:, None]
                points = points.view(-1, 3)[None].expand(N, -1, -1)
            eps1, eps2 = self._blocks_eps
            points_inv = ((points - T[:, None]) @ R.transpose(1, 2)) / (S[:, None] * self.ratio_block_scene)
            sdf = G.superquadric_implicit(points_inv, eps1, eps2)
            occupancy = torch.sigmoid(-sdf / OVERLAP_TEMPERATURE)
            alpha = self._alpha_full if coarse_learning else (self._alpha_full > 0.5).float()
            occupancy = occupancy * alpha[:, None]
            losses['overlap'] = self.loss_weights['overlap'] * factor * (occupancy.sum(0) - OVERLAP_N_BLOCKS).clamp(0).mean()
        losses['total'] = sum(losses.values())
        return losses

    def get_opacities(self):
        alpha = torch.sigmoid(self.alpha_logit)
        if self.kill_blocks:
            alpha = alpha * (alpha > 0.01)
        return alpha

    @torch.no_grad()
    def get_nb_opaque_blocks(self):
        return (self.get_opacities() > 0.5).sum().item()

    @torch.no_grad()
    def get_arranged_block_txt(self):
        maps = torch.sigmoid(self.textures).permute(0, 3, 1, 2)
        ncol, nrow = 5, len(maps) // 5
        rows = [torch.cat([maps[k] for k in range(ncol * i, ncol * (i + 1))], dim=2) for i in range(nrow)]
        return torch.cat(rows, dim=1)[None]

    @torch.no_grad()
    def load_state_dict(self, state_dict, **_unused):
        state = self.state_dict()
        missing = []
        for name, param in state_dict.items():
            name = name.replace('module.', '').replace('spq_', 'sq_')         # dbw.py:444-445 backward compatibility
            if name in state:
                state[name].copy_(param.data if isinstance(param, nn.Parameter) else param)
            else:
                missing.append(name)
        if missing:
            print(f'load_state_dict: {missing} not found')

    @torch.no_grad()
    def quantitative_eval(self, loader, device, hard_inference=True):
        """PSNR of hard renders over a loader (the SSIM / LPIPS columns of dbw.py:464-493 need networks out of scope)."""
        self.eval()
        opacities = self.get_opacities()
        scene = self.build_scene(filter_transparent=True)
        tot, n = 0.0, 0
        for inp, labels in loader:
            inp = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in inp.items()}
            self._install_cameras(inp)
            N = len(inp['imgs'])
            if hard_inference:
                rec = self.renderer(scene.extend(N), inp['R'], inp['T'], viz_purpose=True)[:, :3]
            else:
                rec = self.predict(inp, labels, filter_transparent=True)
            mse = F.mse_loss(inp['imgs'], rec)
            tot += float(-10.0 * torch.log10(mse)) * N
            n += N
        return OrderedDict([('n_blocks', int((opacities > 0.5).sum())), ('PSNR', tot / max(n, 1))]
                           + [(f'alpha{k}', a.item()) for k, a in enumerate(opacities)])


def _make_perceptual(name):
    try:
        import lpips  # noqa: F401  (not in this image)
    except ImportError:
        print(f'[dbw_b200] perceptual loss "{name}" needs the `lpips` package and VGG weights, which are not available '
              f'here; the term is reported as 0 until set_perceptual_loss() installs a callable')
        return None
    if name != 'lpips':
        raise NotImplementedError(name)
    net = lpips.LPIPS(net='vgg')
    return lambda imgs, rec: net(imgs, rec, normalize=True).mean()


def create_model(cfg, img_size, **kwargs):
    """model factory with the reference's signature (src/model/__init__.py:12-17)."""
    kwargs = deepcopy(cfg['model'])
    name = kwargs.pop('name')
    assert name == 'dbw', name
    return DifferentiableBlocksWorld(img_size, **kwargs)
