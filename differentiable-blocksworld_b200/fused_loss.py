"""Decoupled render + compositing + RGB loss as ONE autograd node ("MSE fused into blend"):

    env  = render(environment)                      src/model/dbw.py:219      (renderer_env, K=1)
    fg   = render(blocks)                           src/model/dbw.py:220-221  (renderer / renderer_fine, K=10)
    rec  = fg_rgb * fg_a + (1 - fg_a) * env_rgb     src/model/dbw.py:223
    loss = mean((imgs - rec)^2)                     src/model/dbw.py:366-367  (nn.MSELoss)

The blocks pass runs with the loss epilogue of include/dbw_render.h (dbw_render_forward_loss): compositing, squared error
and the MSE gradients w.r.t. both layers happen per pixel inside the rasterizer while the blended colour is in registers.
Neither `fg` nor `rec` nor the two image-sized gradient round trips of a separate composite kernel ever touch HBM; the
backward is the two raster backward kernels reading those gradients, scaled on the fly by the upstream gradient of the loss
(dbw_render_backward_scaled).  Not usable when another loss consumes `rec` with gradient (LPIPS): the model then keeps
the separate composite kernel (dbw._CompositeMSE), which accepts a gradient on `rec`."""
import ctypes

import torch

from . import _lib
from ._lib import DbwLossEpilogue
from .renderer import _c, _stream, scene_settings

N_PARTIALS = 1024
# run the two passes' backward kernels concurrently (the environment's on a side stream: parallel branches of the step's CUDA
# graph).  One GPU, 49 views: no gain (the kernels fill the machine by themselves); 8 GPUs, ~6 views each: the launches are
# short and one kernel's ramp-up hides the other's tail -- bench.py --overlap-bwd measures it
OVERLAP_BACKWARD_PASSES = False
_side_stream = None


class ScenePass:
    """static description of one render pass: topology, UVs, map layout and the renderer's shading options"""

    def __init__(self, faces, faces_uvs, face_map, map_table_host, renderer, alpha_group=1, n_static_faces=0):
        self.faces = faces.to(torch.int32).contiguous()
        self.faces_uvs = faces_uvs.contiguous().float()
        self.face_map = face_map.to(torch.int32).contiguous()
        self.map_table_host, self.r = map_table_host, renderer
        self.alpha_group, self.n_static_faces = alpha_group, n_static_faces

    def settings(self, verts, maps, B, faces_alpha, view_rows=None):
        r = self.r
        return scene_settings(verts, self.faces, maps, self.map_table_host, B, r.cameras.intrinsics(), r.img_size, r.sigma,
                              r.faces_per_pixel, r.z_clip, r.detach_bary, r.clip_inside, r.background_color, faces_alpha,
                              r.perspective_correct, False, r.blur_radius, True, self.alpha_group, self.n_static_faces, view_rows)


def _workspace(cfg, dev):
    fwd, bwd = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(_lib.lib().dbw_workspace_bytes(ctypes.byref(cfg), ctypes.byref(fwd), ctypes.byref(bwd)), 'dbw_workspace_bytes')
    return torch.empty(fwd.value, dtype=torch.uint8, device=dev), bwd.value


class _SceneMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env_verts, env_maps, blk_verts, blk_maps, blk_alpha, R, T, imgs, env_pass, blk_pass, blk_face_map,
                cfgs, inv_count, want_rec):
        L = _lib.lib()
        dev = blk_verts.device
        B, _, H, W = imgs.shape
        f32 = lambda t: t.detach().contiguous().float()
        env_verts, env_maps, blk_verts, blk_maps = f32(env_verts), f32(env_maps), f32(blk_verts), f32(blk_maps)
        R, T, imgs = f32(R), f32(T), f32(imgs)
        fa = f32(blk_alpha) if blk_alpha is not None else None
        fmap_b = blk_face_map.to(torch.int32).contiguous()
        (cfg_e, table_e), (cfg_b, table_b) = cfgs
        ws_e, bwd_e = _workspace(cfg_e, dev)
        ws_b, bwd_b = _workspace(cfg_b, dev)
        env_rgba = torch.empty(B, 4, H, W, dtype=torch.float32, device=dev)
        _lib.check(L.dbw_render_forward_ex(ctypes.byref(cfg_e), _c(env_verts), _c(env_pass.faces), _c(env_pass.faces_uvs),
                                           _c(env_pass.face_map), _c(env_maps), _c(table_e), _c(R), _c(T), None, _c(env_rgba),
                                           None, _c(ws_e), ws_e.numel(), None, None, _stream()), 'dbw_render_forward_ex')
        g_fg = torch.empty(B, 4, H, W, dtype=torch.float32, device=dev)       # d loss / d blocks RGBA (unscaled)
        g_env = torch.empty(B, 4, H, W, dtype=torch.float32, device=dev)      # d loss / d env RGBA (unscaled)
        partials = torch.empty(N_PARTIALS, dtype=torch.float32, device=dev)
        rec = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_rec else None
        ep = DbwLossEpilogue()
        ep.env_rgba, ep.target, ep.g_env = env_rgba.data_ptr(), imgs.data_ptr(), g_env.data_ptr()
        ep.rec = rec.data_ptr() if rec is not None else None
        ep.loss_partials, ep.n_partials, ep.inv_count = partials.data_ptr(), N_PARTIALS, float(inv_count)
        _lib.check(L.dbw_render_forward_loss(ctypes.byref(cfg_b), _c(blk_verts), _c(blk_pass.faces), _c(blk_pass.faces_uvs),
                                             _c(fmap_b), _c(blk_maps), _c(table_b), _c(R), _c(T), _c(fa), _c(g_fg), None,
                                             _c(ws_b), ws_b.numel(), ctypes.byref(ep), _stream()), 'dbw_render_forward_loss')
        loss = partials.sum()
        ctx.save_for_backward(env_verts, env_maps, blk_verts, blk_maps, fa, R, T, fmap_b, table_e, table_b, ws_e, ws_b, g_fg, g_env)
        ctx.meta = (env_pass, blk_pass, cfg_e, cfg_b, bwd_e, bwd_b)
        if want_rec:
            ctx.mark_non_differentiable(rec)
            return loss, rec
        return loss

    @staticmethod
    def backward(ctx, g_loss, _g_rec=None):
        env_verts, env_maps, blk_verts, blk_maps, fa, R, T, fmap_b, table_e, table_b, ws_e, ws_b, g_fg, g_env = ctx.saved_tensors
        env_pass, blk_pass, cfg_e, cfg_b, bwd_e, bwd_b = ctx.meta
        L = _lib.lib()
        dev = blk_verts.device
        gl = g_loss.detach().contiguous().float()
        need = ctx.needs_input_grad

        # every gradient buffer of both passes (and the kernels' scratch, which they expect zeroed... they zero it themselves)
        # comes out of ONE zero-filled arena: one fill launch per step instead of five
        wants = [('gm_b', blk_maps, need[3]), ('gm_e', env_maps, need[1]), ('gv_b', blk_verts, need[2]), ('gv_e', env_verts, need[0]),
                 ('ga_b', fa, need[4] and fa is not None)]
        total = sum((t.numel() + 3) // 4 * 4 for _, t, on in wants if on)
        arena = torch.zeros(total, device=dev, dtype=torch.float32)
        g, off = {}, 0
        for name, t, on in wants:
            g[name] = None
            if on:
                g[name] = arena[off:off + t.numel()].view_as(t)
                off += (t.numel() + 3) // 4 * 4                      # float4 texel atlases stay 16-byte aligned

        def run(cfg, p, fmap, verts, maps, table, alpha, ws, grad, bwd_bytes, g_verts, g_maps, g_alpha):
            if g_verts is None and g_maps is None and g_alpha is None:
                return
            scratch = torch.empty(bwd_bytes, dtype=torch.uint8, device=dev)
            _lib.check(L.dbw_render_backward_scaled(ctypes.byref(cfg), _c(verts), _c(p.faces), _c(p.faces_uvs), _c(fmap), _c(maps),
                                                    _c(table), _c(R), _c(T), _c(alpha), None, _c(ws), ws.numel(), _c(grad),
                                                    _c(gl), _c(g_verts), _c(g_alpha), _c(g_maps), _c(scratch), scratch.numel(),
                                                    _stream()), 'dbw_render_backward_scaled')

        if OVERLAP_BACKWARD_PASSES:
            global _side_stream
            cur = torch.cuda.current_stream()
            if _side_stream is None or _side_stream.device != dev:
                _side_stream = torch.cuda.Stream(device=dev)
            _side_stream.wait_stream(cur)
            with torch.cuda.stream(_side_stream):
                run(cfg_e, env_pass, env_pass.face_map, env_verts, env_maps, table_e, None, ws_e, g_env, bwd_e, g['gv_e'], g['gm_e'], None)
            run(cfg_b, blk_pass, fmap_b, blk_verts, blk_maps, table_b, fa, ws_b, g_fg, bwd_b, g['gv_b'], g['gm_b'], g['ga_b'])
            cur.wait_stream(_side_stream)
        else:
            run(cfg_b, blk_pass, fmap_b, blk_verts, blk_maps, table_b, fa, ws_b, g_fg, bwd_b, g['gv_b'], g['gm_b'], g['ga_b'])
            run(cfg_e, env_pass, env_pass.face_map, env_verts, env_maps, table_e, None, ws_e, g_env, bwd_e, g['gv_e'], g['gm_e'], None)
        gv_e, gm_e, gv_b, gm_b, ga_b = g['gv_e'], g['gm_e'], g['gv_b'], g['gm_b'], g['ga_b']
        return gv_e, gm_e, gv_b, gm_b, ga_b, None, None, None, None, None, None, None, None, None


def scene_mse(env_verts, env_atlas, blk_verts, blk_atlas, blk_alpha, R, T, imgs, env_pass, blk_pass, blk_face_map,
              n_total_views=None, return_rec=False, view_rows=None):
    """MSE between `imgs` (B,3,H,W) and the blocks composited over the environment, both rendered from raw scene tensors
    (float4 texel atlases from scene_ops.texture_atlas).  `n_total_views`: size of the GLOBAL batch the mean runs over
    (view-sharded data parallelism); `view_rows` (B,2) int32: the [row_begin, row_end) of each view this rank owns.  Returns the loss, or (loss, rec) with a non-differentiable `rec`."""
    B, _, H, W = imgs.shape
    inv = 1.0 / (float(n_total_views or B) * 3 * H * W)
    # settings are made out here: inside Function.forward grad mode is off and the passes would not save fragment state
    cfgs = (env_pass.settings(env_verts, env_atlas, B, None, view_rows), blk_pass.settings(blk_verts, blk_atlas, B, blk_alpha, view_rows))
    return _SceneMSEFn.apply(env_verts, env_atlas, blk_verts, blk_atlas, blk_alpha, R, T, imgs, env_pass, blk_pass,
                             blk_face_map, cfgs, inv, bool(return_rec))
