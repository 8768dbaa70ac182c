/*
 * include/dbw_render.h -- C-ABI of the B200-native differentiable primitive renderer.
 *
 * The reference (monniert/differentiable-blocksworld) has no FFI: its render hot path is a Python-level seam,
 *     Renderer.forward(meshes, R, T, viz_purpose=False, **kwargs) -> (B,4,H,W)      src/model/renderer.py:84-98
 * behind which PyTorch3D's MeshRasterizer (_C.rasterize_meshes / _C.rasterize_meshes_backward) and the
 * reference's LayeredShader + layered_rgb_blend (src/model/renderer.py:219-273) run.  This header is what a
 * binding for that seam loads instead: plain pointers and sizes, no torch types.  Each entry point cites the
 * reference interface it replaces.  INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions
 *  - every pointer except `settings` is a DEVICE pointer on the current CUDA device (the *_host entry points
 *    take HOST pointers and do the copies themselves);
 *  - the caller owns every buffer; the library never allocates or frees device memory (the *_host entry
 *    points keep one grow-only per-process scratch arena, released by dbw_host_arena_release());
 *  - all work is ordered on `stream` (a cudaStream_t passed as void*); no implicit synchronisation;
 *  - gradient outputs ACCUMULATE (atomics): the caller zero-fills them;
 *  - return value 0 = success, negative = error (message via dbw_last_error(), thread-local); never throws;
 *  - dtype is float32 throughout; ids are int32.
 */
#ifndef DBW_RENDER_H
#define DBW_RENDER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBW_ABI_VERSION 7
#define DBW_MAX_FACES_PER_PIXEL 64

/*
 * Mirrors RasterizationSettings + BlendParams + the LayeredShader flags the reference configures at
 * src/model/renderer.py:29-54 and the camera it installs at src/model/dbw.py:204-208.
 */
typedef struct DbwRenderSettings {
  int32_t n_views;            /* B: cameras / images in this call (Meshes.extend(B), src/model/dbw.py:215,220)      */
  int32_t height, width;      /* img_size (H, W)                                                                    */
  int32_t faces_per_pixel;    /* K  (renderer.py:33)                                                                */
  int32_t n_verts, n_faces;   /* V, F of the (single) scene mesh shared by all views                                 */
  int32_t n_maps;             /* M texture maps (join_meshes_as_scene keeps one per sub-mesh)                        */
  int32_t alpha_view_stride;  /* faces_alpha layout: 0 = (F,) shared by all views; F = (B*F,) batch-packed (dbw.py:219) */
  float fx, fy, px, py;       /* PerspectiveCameras K in NDC (src/dataset/dtu.py:102-106): K[0,0], K[1,1], K[0,2], K[1,2] */
  float sigma;                /* BlendParams.sigma (renderer.py:31); 0 = hard                                        */
  float blur_radius;          /* log(1/1e-4 - 1) * sigma (renderer.py:51)                                            */
  float z_clip;               /* z_clip_value (renderer.py:35,46); < 0 disables clipping                             */
  float proj_eps;             /* eps passed to the cameras (renderer.py:20,94) = 1e-8                                 */
  float background[3];        /* BlendParams.background_color (renderer.py:32)                                       */
  int32_t clip_inside;        /* LayeredShader clip_inside (renderer.py:41)                                          */
  int32_t perspective_correct;/* renderer.py:34,46 (None -> True for perspective cameras)                            */
  int32_t clip_barycentric;   /* clip_barycentric_coords=True (renderer.py:46)                                       */
  int32_t detach_bary;        /* LayeredShader detach_bary (renderer.py:43,222-223): no gradient through barycentrics */
  int32_t verts_are_ndc;      /* 1: `verts` is (B,V,3) = (x_ndc, y_ndc, z_view) and R/T/K are ignored                 */
  int32_t n_map_floats;       /* total floats in `maps` (= 3 * sum_m H_m*W_m): sizes the library's float4 texel scratch    */
  int32_t maps_are_texels4;   /* 1: `maps` (and `g_maps`) are already RGB+pad float4 texel atlases (dbw_texture_prep_*);
                                 DbwMapDesc.offset keeps its meaning (3 * first texel index)                              */
  int32_t save_fragment_state;/* 1: the forward also keeps one 16 B record {triangle slot | closest edge, u, v, signed squared
                                 distance} per kept fragment (+ a per-pixel count byte) in its workspace; dbw_render_backward*
                                 streams these instead of re-deriving geometry and REQUIRES them.  Costs
                                 B*H*W*(16 K + 1) bytes of workspace ADDRESS space (only fragments that exist are touched)  */
  int32_t alpha_group;        /* faces sharing one opacity entry (0/1: one per face).  faces_alpha / g_faces_alpha then have
                                 n_faces / alpha_group entries per view: a DBW block is 80 faces with ONE opacity
                                 (src/model/dbw.py:219 repeat_interleave's it per face; that is alpha_group = 1)            */
  int32_t n_static_faces;     /* faces [0, n_static_faces) have constant vertices (the background sphere of the environment,
                                 src/model/dbw.py:267-280): the backward skips their vertex gradient                        */
  const int32_t* view_rows;   /* DEVICE (B,2) int32 or NULL: [row_begin, row_end) of each view that this call renders / differentiates.
                                 Rows outside are left untouched in every output (and contribute nothing to a fused loss):
                                 a data-parallel step shards at (view, row band) granularity, so 49 views split evenly
                                 over 8 ranks                                                                   */
} DbwRenderSettings;

/* Texture table entry: map m lives at maps[offset .. offset + height*width*3), row-major (H, W, 3). */
typedef struct DbwMapDesc { int32_t offset, height, width, reserved; } DbwMapDesc;

int dbw_abi_version(void);
/* Test hook: 1 = hard single-layer renders (K = 1, sigma = 0) go through the generic raster kernel instead of their dedicated one. */
void dbw_debug_generic_kernel_only(int on);
/* sizeof(DbwRenderSettings) as compiled into the library: lets a binding verify its mirror of the struct. */
size_t dbw_sizeof_settings(void);
const char* dbw_last_error(void);

/* Bytes of the two caller-provided device scratch buffers:
 *   *fwd_bytes: workspace written by dbw_render_forward and read again by dbw_render_backward (keep it alive);
 *   *bwd_bytes: scratch used only inside dbw_render_backward. */
int dbw_workspace_bytes(const DbwRenderSettings* settings, size_t* fwd_bytes, size_t* bwd_bytes);

/*
 * Forward of Renderer.forward (src/model/renderer.py:84-98): projection -> z-clip -> rasterize (top-K nearest
 * faces per pixel within the blur halo) -> UV interpolation + bilinear texture fetch -> layered soft blend.
 *   verts        (V,3) world-space vertices   [or (B,V,3) NDC when settings->verts_are_ndc]
 *   faces        (F,3) int32 vertex ids
 *   faces_uvs    (F,3,2) per-face-vertex UVs = verts_uvs[faces_uvs] of TexturesUV (src/model/dbw.py:280,295,342)
 *   face_map     (F) int32 texture map of each face; a NEGATIVE entry disables the face (it is never rasterized and gets
 *                no gradient) -- how a static-topology caller drops low-opacity blocks (src/model/dbw.py:316-328)
 *   maps, map_table  packed maps + (M) DbwMapDesc (device)
 *   R (B,3,3), T (B,3)  row-vector convention X_view = X_world @ R + T (src/dataset/dtu.py:75-124)
 *   faces_alpha  NULL, or per-face opacity (src/model/dbw.py:219, renderer.py:258-260), see alpha_view_stride / alpha_group
 *   out_rgba     (B,4,H,W)  RGB + coverage, NCHW (renderer.py:268)
 *   topk_ids     (B,K,H,W) int32 or NULL: z-sorted face slots kept per pixel (-1 = empty) = fragments.pix_to_face modulo
 *                the slot convention; a diagnostic / visualisation output (render_edges), NOT what the backward reads --
 *                that is the workspace's fragment records (settings->save_fragment_state)
 */
int dbw_render_forward(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                       const float* faces_uvs, const int32_t* face_map, const float* maps,
                       const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                       float* out_rgba, int32_t* topk_ids, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Same, with the two optional extras the reference's visualisation paths need:
 *   face_shade  (B,F,3) or NULL: per-view per-face colour multiplier = PyTorch3D flat shading with ambient + directional
 *               diffuse light, as configured for renderer_light (src/model/dbw.py:139-143, renderer.py:87-97,195-205)
 *   out_dists   (B,K,H,W) or NULL: signed squared NDC distance of each kept fragment to its face's nearest edge
 *               (< 0 inside, -1 for empty slots) = fragments.dists, what render_edges thresholds (renderer.py:134-146)
 */
int dbw_render_forward_ex(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                          const float* faces_uvs, const int32_t* face_map, const float* maps,
                          const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                          float* out_rgba, int32_t* topk_ids, void* workspace, size_t workspace_bytes,
                          const float* face_shade, float* out_dists, void* stream);

/*
 * Backward of the same (replaces autograd through layered_rgb_blend, grid_sample, interpolate_face_attributes
 * and _C.rasterize_meshes_backward).  grad_rgba (B,4,H,W).  Outputs accumulate:
 *   g_verts        (V,3)  [or (B,V,3) when verts_are_ndc]
 *   g_faces_alpha  same shape as faces_alpha (may be NULL when faces_alpha is NULL)
 *   g_maps         same packing as maps (may be NULL: no texture gradient)
 * Requires the forward to have run with save_fragment_state = 1 on the same workspace; `topk_ids` is ignored (may be
 * NULL; kept in the signature since ABI 1).
 */
int dbw_render_backward(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                        const float* faces_uvs, const int32_t* face_map, const float* maps,
                        const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                        const int32_t* topk_ids, const void* workspace, size_t workspace_bytes,
                        const float* grad_rgba, float* g_verts, float* g_faces_alpha, float* g_maps,
                        void* bwd_scratch, size_t bwd_scratch_bytes, void* stream);

/*
 * Fused compositing + RGB loss of the decoupled scene (src/model/dbw.py:223 and :366-367):
 *   rec = rgb_fg * a_fg + (1 - a_fg) * rgb_env ;  loss = weight * mean((imgs - rec)^2)
 * fg, env: (B,4,H,W) outputs of dbw_render_forward; imgs (B,3,H,W).  Writes rec (B,3,H,W) (may be NULL), adds
 * the loss into *loss_sum (device scalar, caller zero-fills), and -- when g_fg / g_env are not NULL -- writes the
 * gradients of `loss` w.r.t. fg and env ((B,4,H,W), overwritten).  `inv_count` = weight / (B_total*3*H*W).
 */
int dbw_composite_mse(int32_t n_views, int32_t height, int32_t width, const float* fg, const float* env,
                      const float* imgs, float inv_count, float* rec, float* loss_sum, float* g_fg, float* g_env,
                      void* stream);

/* Backward of dbw_composite_mse: d(g_loss * loss + <g_rec, rec>) / d fg, d env, written to g_fg / g_env ((B,4,H,W)).
 * g_loss: device scalar (may be NULL = 0); g_rec: (B,3,H,W) gradient arriving at `rec` from other losses (may be NULL). */
int dbw_composite_mse_backward(int32_t n_views, int32_t height, int32_t width, const float* fg, const float* env,
                               const float* imgs, float inv_count, const float* g_loss, const float* g_rec, float* g_fg,
                               float* g_env, void* stream);

/*
 * Fused loss epilogue ("MSE fused into blend").  dbw_render_forward_loss is dbw_render_forward for the
 * LAST layer of a decoupled render (the blocks, src/model/dbw.py:219-223) with the compositing over the already rendered
 * environment and the RGB loss (dbw.py:366-367, nn.MSELoss) done in the rasterizer's epilogue, per pixel, while the blended
 * colour is still in registers:
 *     rec = rgb * a + (1 - a) * rgb_env ;  loss = sum((rec - target)^2) * inv_count
 * Instead of the image, `out_rgba` receives d loss / d (this layer's RGBA) and `g_env` d loss / d env_rgba (alpha plane 0),
 * i.e. exactly the grad_rgba inputs of the two dbw_render_backward calls (scaled there by the upstream gradient of the
 * loss through `grad_scale`): no composite kernel, no image round trip through HBM in either direction.
 */
typedef struct DbwLossEpilogue {
  const float* env_rgba;      /* (B,4,H,W) render of the layer behind (dbw_render_forward output)                          */
  const float* target;        /* (B,3,H,W) ground-truth images                                                             */
  float* g_env;               /* (B,4,H,W) out: d loss / d env_rgba                                                        */
  float* rec;                 /* (B,3,H,W) out: composited image, or NULL                                                  */
  float* loss_partials;       /* (n_partials) out: zero-filled by the call, then partial sums; loss = their sum            */
  int32_t n_partials;         /* power of two, >= 1                                                                        */
  float inv_count;            /* weight / (B_total * 3 * H * W): the mean of nn.MSELoss over the GLOBAL batch              */
} DbwLossEpilogue;

int dbw_render_forward_loss(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                            const float* faces_uvs, const int32_t* face_map, const float* maps,
                            const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                            float* out_g_rgba, int32_t* topk_ids, void* workspace, size_t workspace_bytes,
                            const DbwLossEpilogue* epilogue, void* stream);

/* dbw_render_backward with grad_rgba multiplied by the device scalar *grad_scale (NULL = 1) as it is read. */
int dbw_render_backward_scaled(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                               const float* faces_uvs, const int32_t* face_map, const float* maps,
                               const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                               const int32_t* topk_ids, const void* workspace, size_t workspace_bytes,
                               const float* grad_rgba, const float* grad_scale, float* g_verts, float* g_faces_alpha,
                               float* g_maps, void* bwd_scratch, size_t bwd_scratch_bytes, void* stream);

/*
 * Fused scene construction (src/model/dbw.py:267-352), so that a training step needs no eager tensor ops between the
 * leaf parameters and the rasterizer.  All pointers are device pointers.
 */
typedef struct DbwSceneGeometry {
  int32_t n_blocks, verts_per_block, n_ground_verts, reserved;
  const float* sq_eta;        /* (N,Vb) buffers sq_eta / sq_omega (dbw.py:86-87)                                           */
  const float* sq_omega;
  const float* sq_eps;        /* (N,2)  dbw.py:84, eps = sigmoid(.)*1.8 + 0.1 (dbw.py:349)                                 */
  const float* S;             /* (N,3)  scale = exp(S) + scale_min (dbw.py:299)                                            */
  const float* R_6d;          /* (N,6)  rotation_6d_to_matrix (dbw.py:299)                                                 */
  const float* T;             /* (N,3)                                                                                     */
  const float* ground_verts;  /* (Vg,3) static plane vertices (dbw.py:76-78); may be NULL when n_ground_verts == 0         */
  const float* R_6d_ground;   /* (6)    dbw.py:99,285                                                                      */
  const float* T_ground;      /* (3)    dbw.py:100                                                                         */
  float ratio_block_scene, scale_min, S_world;
  float R_world[9], T_world[3];   /* world transform (v * S_world) @ R_world + T_world (dbw.py:264,344)                    */
} DbwSceneGeometry;

/* verts_out (N*Vb + Vg, 3): world-space vertices of the N superquadric blocks, then of the ground plane. */
int dbw_scene_geometry_forward(const DbwSceneGeometry* g, float* verts_out, void* stream);
/* g_verts (N*Vb + Vg, 3) -> gradients of the leaf parameters (written, not accumulated). */
int dbw_scene_geometry_backward(const DbwSceneGeometry* g, const float* g_verts, float* g_sq_eps, float* g_S, float* g_R_6d,
                                float* g_T, float* g_R_6d_ground, float* g_T_ground, void* stream);
/* The layout the two render passes consume without a concatenation: verts_out (n_env_static + Vg + N*Vb, 3) = the static
 * environment vertices (the background sphere in world space, dbw.py:271-274; copied), the ground, then the blocks -- the
 * environment pass takes the first n_env_static + Vg rows, the blocks pass the rest.  The backward takes the two gradient
 * blocks where they lie (any two device pointers: no gather). */
int dbw_scene_geometry_forward_env(const DbwSceneGeometry* g, const float* env_static_verts, int32_t n_env_static,
                                   float* verts_out, void* stream);
int dbw_scene_geometry_backward_parts(const DbwSceneGeometry* g, const float* g_block_verts, const float* g_ground_verts,
                                      float* g_sq_eps, float* g_S, float* g_R_6d, float* g_T, float* g_R_6d_ground,
                                      float* g_T_ground, void* stream);

/* Block opacities of a step (src/model/dbw.py:300-316 with static shapes), one launch each way:
 *   alpha[k]      = sigmoid(alpha_logit[k] + noise_scale * noise[k])      (noise may be NULL)
 *   keep[k]       = keep_threshold < 0  ||  sigmoid(alpha_logit[k]) > keep_threshold      (0.5: hard filter, 0.01: kill_blocks)
 *   alpha_kept[k] = keep[k] ? alpha[k] : 0                                (optional)
 *   face_map_out[k*faces_per_block + j] = keep[k] ? face_map_in[..] : -1  (optional; -1 = never rasterized)
 * backward: g_alpha_logit = (g_alpha + keep * g_alpha_kept) * alpha * (1 - alpha)   (either gradient may be NULL). */
int dbw_opacity_forward(const float* alpha_logit, const float* noise, float noise_scale, float keep_threshold,
                        const int32_t* face_map_in, int32_t n_blocks, int32_t faces_per_block, float* alpha, float* alpha_kept,
                        int32_t* face_map_out, void* stream);
int dbw_opacity_backward(const float* alpha_logit, const float* noise, float noise_scale, float keep_threshold,
                         const float* g_alpha, const float* g_alpha_kept, int32_t n_blocks, float* g_alpha_logit, void* stream);

/* textures (M,TS,TS,3) logits -> atlas (M, TS, p_left+TS+p_right) float4 texels: sigmoid, optional `decimate`xdecimate box
 * filter (1 = off, 8 = dbw.py:331-334), circular padding along u (dbw.py:339-341).  Backward: g_atlas -> g_textures. */
/* One texture stack of a multi-stack texture preparation: (n_maps, txt_size, txt_size, 3) logits -> float4 texel atlas
 * (n_maps, txt_size, p_left + txt_size + p_right) [forward: `atlas` is written; backward: `atlas` is the atlas GRADIENT that is read
 * and `g_textures` receives d/d logits]. */
/* `stage` splits the preparation at the CELL COLOURS c = box_mean(sigmoid(logits)), (n_maps, txt_size/decimate, txt_size/decimate, 3)
 * floats -- the smallest tensor on the way (64x smaller than the textures while they are decimated), where a data-parallel
 * trainer sums gradients over ranks (parallel.GradSumPoint):
 *   DBW_TEX_STAGE_FUSED   logits -> atlas                                     (backward: atlas gradient -> logit gradient)
 *   DBW_TEX_STAGE_CELLS   logits -> cells, written to `atlas`                 (backward: `atlas` = CELL gradient -> logit gradient)
 *   DBW_TEX_STAGE_EXPAND  cells (passed as `textures`) -> atlas               (backward: atlas gradient -> cell gradient, written
 *                                                                              to `g_textures`; `textures` is not read)
 * CELLS then EXPAND equals FUSED bit for bit, forward and backward. */
#define DBW_MAX_TEX_JOBS 4
#define DBW_TEX_STAGE_FUSED 0
#define DBW_TEX_STAGE_CELLS 1
#define DBW_TEX_STAGE_EXPAND 2
typedef struct DbwTexJob {
  const float* textures; float* atlas; float* g_textures;
  int32_t n_maps, txt_size, p_left, p_right, decimate, stage;
} DbwTexJob;
/* Several stacks in ONE launch (the step's three: background, ground, blocks; `jobs` is a HOST array of n_jobs <= 4). */
int dbw_texture_prep_forward_multi(const DbwTexJob* jobs, int32_t n_jobs, void* stream);
int dbw_texture_prep_backward_multi(const DbwTexJob* jobs, int32_t n_jobs, void* stream);
int dbw_texture_prep_forward(const float* textures, int32_t n_maps, int32_t txt_size, int32_t p_left, int32_t p_right,
                             int32_t decimate, float* atlas_out, void* stream);
int dbw_texture_prep_backward(const float* textures, int32_t n_maps, int32_t txt_size, int32_t p_left, int32_t p_right,
                              int32_t decimate, const float* g_atlas, float* g_textures, void* stream);

/*
 * Host-buffer variants (the end-to-end call a reference-side binding makes with numpy / CPU tensors): every
 * pointer is a HOST pointer; the library stages through its own device arena on the current device, runs the
 * device entry points above on `stream`, copies results back and synchronises the stream before returning.
 */
int dbw_render_forward_host(const DbwRenderSettings* settings, const float* verts, const int32_t* faces,
                            const float* faces_uvs, const int32_t* face_map, const float* maps, size_t maps_floats,
                            const DbwMapDesc* map_table, const float* R, const float* T, const float* faces_alpha,
                            float* out_rgba, void* stream);
void dbw_host_arena_release(void);

/* Number of kernel launches issued by this library since process start (for bench.py's gpu_launches). */
uint64_t dbw_launch_count(void);

/* Optional live timing of the two rasterization kernels with CUDA events on the launch stream (bench.py's roofline):
 * kind 0 = raster forward, 1 = raster backward; K filters by faces_per_pixel (0 = any).  dbw_timing_read synchronises
 * on the recorded events and returns their summed duration and count; dbw_timing_reset frees them. */
void dbw_timing_enable(int on);
int dbw_timing_read(int kind, int K, double* total_ms, int* count);
void dbw_timing_reset(void);

/*
 * The step's one gradient exchange (views shard with no data-path collective; ONE all-reduce(SUM) of the flat
 * parameter-gradient bucket) as a hand-written all-reduce over NVLink peer memory -- a plain kernel launch on `stream`,
 * hence capturable inside the step's CUDA graph.  One process per GPU of one node:
 *   dbw_comm_create      allocates this rank's arena (bucket of max_floats + inbox of the same size + control; cudaMalloc)
 *   dbw_comm_buffer      the arena's bucket: the caller keeps its gradients THERE (nothing is staged or copied)
 *   dbw_comm_ipc_handle  64-byte cudaIpcMemHandle of the arena, to be all-gathered by the caller (torch.distributed)
 *   dbw_comm_connect     opens the peers' arenas: all_handles = world x 64 bytes in rank order
 *   dbw_comm_all_reduce  in-place SUM of the bucket's first n_floats over the ranks (n_floats % 4 == 0; buf must be the
 *                        pointer dbw_comm_buffer returned).  Push protocol: slices are pushed to their owners, summed there
 *                        in rank order (bit-identical results everywhere) and pushed back -- only posted stores cross NVLink
 *   dbw_comm_error       0, or which barrier timed out (a peer did not arrive within ~30 s: results are garbage, no hang)
 */
int dbw_comm_create(int32_t world, int32_t rank, size_t max_floats, void** comm_out);
int dbw_comm_buffer(void* comm, float** out);
int dbw_comm_ipc_handle(void* comm, void* out_handle64);
int dbw_comm_connect(void* comm, const void* all_handles);
int dbw_comm_all_reduce(void* comm, float* buf, size_t n_floats, void* stream);
int dbw_comm_error(void* comm, int32_t* out);
int dbw_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* DBW_RENDER_H */
