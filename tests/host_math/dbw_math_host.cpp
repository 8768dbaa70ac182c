// Host build of the DEVICE math of the rasterization kernels (differentiable-blocksworld_b200/csrc/dbw_math.cuh), for the
// CPU test suite: the very functions the CUDA kernels inline -- edge functions, barycentrics, perspective correction,
// clipping, point-triangle distance and their backward pieces -- compiled with g++ behind a small shim and driven over an
// image exactly the way raster_forward_kernel / raster_backward_kernel drive them per (pixel, face).  Test infrastructure
// only (tests/test_device_math_host.py compares it with the oracle); built with -ffp-contract=off so that the
// explicitly rounded intrinsics (__fmul_rn, ...) are plain IEEE operations, as on the device.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __restrict__
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float4 __ldg(const float4* p) { return *p; }

#include "../../differentiable-blocksworld_b200/csrc/dbw_math.cuh"
#include "../../differentiable-blocksworld_b200/csrc/dbw_clip.cuh"
#include "../../differentiable-blocksworld_b200/csrc/dbw_fraglist.cuh"
#include "../../differentiable-blocksworld_b200/csrc/dbw_scene_math.cuh"

// record of one triangle as face_setup's write_slot packs it (dbw_render.cu, write_slot): reciprocal of the eps-shifted area
// and of the squared edge lengths (-1 = degenerate edge)
static TriGeom make_tri(const float* fv) {
  const float x0 = fv[0], y0 = fv[1], z0 = fv[2], x1 = fv[3], y1 = fv[4], z1 = fv[5], x2 = fv[6], y2 = fv[7], z2 = fv[8];
  const f2 a = {x0, y0}, b = {x1, y1}, c = {x2, y2};
  const float area = edge_nc(c, a, b);
  const float l01 = (x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0), l02 = (x2 - x0) * (x2 - x0) + (y2 - y0) * (y2 - y0);
  const float l12 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
  const float4 r0 = make_float4(x0, y0, x1, y1), r1 = make_float4(x2, y2, z0, z1);
  const float4 r2 = make_float4(z2, __int_as_float(0), __int_as_float(-1), __int_as_float(0));
  const float4 r3 = make_float4(1.f / __fadd_rn(area, DBW_KEPS), l01 <= DBW_KEPS ? -1.f : 1.f / l01,
                                l02 <= DBW_KEPS ? -1.f : 1.f / l02, l12 <= DBW_KEPS ? -1.f : 1.f / l12);
  return unpack_tri(r0, r1, r2, r3);
}

extern "C" {

// One triangle (fv = 9 floats: NDC x, NDC y, view z per vertex), K = 1, over an H x W image: the candidate test of
// raster_forward_kernel's inner loop.  Outputs per pixel: hit (0/1), zbuf, bary (3), signed squared distance -- -1 where
// the face does not reach the pixel, like fragments of the reference rasterizer.
void hm_forward(const float* fv, int H, int W, float blur, int persp, int clipb, int* hit, float* zbuf, float* bary, float* dists) {
  const TriGeom t = make_tri(fv);
  const float sqrt_blur = sqrtf(blur);
  const float xmin = fminf(fminf(fv[0], fv[3]), fv[6]) - sqrt_blur, xmax = fmaxf(fmaxf(fv[0], fv[3]), fv[6]) + sqrt_blur;
  const float ymin = fminf(fminf(fv[1], fv[4]), fv[7]) - sqrt_blur, ymax = fmaxf(fmaxf(fv[1], fv[4]), fv[7]) + sqrt_blur;
  for (int yi = 0; yi < H; ++yi)
    for (int xi = 0; xi < W; ++xi) {
      const int o = yi * W + xi;
      hit[o] = 0; zbuf[o] = -1.f; dists[o] = -1.f; bary[o * 3] = bary[o * 3 + 1] = bary[o * 3 + 2] = -1.f;
      const f2 p = {pix_to_ndc(W - 1 - xi, W, H), pix_to_ndc(H - 1 - yi, H, W)};
      if (p.x > xmax || p.x < xmin || p.y > ymax || p.y < ymin) continue;
      const Edges ed = eval_edges(p, t);
      if (!ed.inside && blur == 0.f) continue;
      const float dist = tri_dist2(p, t);
      if (!ed.inside && dist >= blur) continue;
      const Bary b = bary_from_edges(ed, t, persp != 0, clipb != 0);
      if (b.pz < 0.f) continue;
      hit[o] = 1; zbuf[o] = b.pz; dists[o] = b.inside ? -dist : dist;
      bary[o * 3] = b.bc.x; bary[o * 3 + 1] = b.bc.y; bary[o * 3 + 2] = b.bc.z;
    }
}

// Gradient of sum(grad_zbuf * zbuf + <grad_bary, bary> + grad_dists * dists) over the hit pixels w.r.t. the 9 floats of
// the triangle, composed from the device backward pieces in the order raster_backward_kernel composes them.
void hm_backward(const float* fv, int H, int W, int persp, int clipb, const int* hit, const float* grad_zbuf,
                 const float* grad_bary, const float* grad_dists, float* g_fv) {
  const TriGeom t = make_tri(fv);
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int yi = 0; yi < H; ++yi)
    for (int xi = 0; xi < W; ++xi) {
      const int o = yi * W + xi;
      if (!hit[o]) continue;
      const f2 p = {pix_to_ndc(W - 1 - xi, W, H), pix_to_ndc(H - 1 - yi, H, W)};
      const Bary b = eval_bary(p, t, persp != 0, clipb != 0);
      const float gz = grad_zbuf[o];
      f3 gb = {gz * t.z0 + grad_bary[o * 3], gz * t.z1 + grad_bary[o * 3 + 1], gz * t.z2 + grad_bary[o * 3 + 2]};
      float gz0 = gz * b.bc.x, gz1 = gz * b.bc.y, gz2 = gz * b.bc.z;
      if (clipb) gb = clip_backward(b.bp, gb);
      if (persp) gb = persp_backward(b.b0, t.z0, t.z1, t.z2, gb, gz0, gz1, gz2);
      f2 g0 = {0.f, 0.f}, g1 = {0.f, 0.f}, g2 = {0.f, 0.f};
      bary_backward(p, t, gb, g0, g1, g2);
      tri_dist_backward(p, t, b.inside ? -grad_dists[o] : grad_dists[o], g0, g1, g2);
      acc[0] += g0.x; acc[1] += g0.y; acc[2] += gz0; acc[3] += g1.x; acc[4] += g1.y; acc[5] += gz1;
      acc[6] += g2.x; acc[7] += g2.y; acc[8] += gz2;
    }
  for (int i = 0; i < 9; ++i) g_fv[i] = (float)acc[i];
}

// bilinear tap of tex_tap (TexturesUV.sample_textures semantics) on an H x W x 3 map: colour, and d colour / d (u, v)
void hm_sample(const float* map, int H, int W, float u, float v, float* rgb, float* d_du, float* d_dv) {
  const TexTap t = tex_tap(u, v, 0, H, W);
  const int idx[4] = {t.i00, t.i01, t.i10, t.i11};
  const float w[4] = {t.w00, t.w01, t.w10, t.w11};
  float c[4][3];
  for (int k = 0; k < 4; ++k) for (int ch = 0; ch < 3; ++ch) c[k][ch] = idx[k] >= 0 ? map[idx[k] * 3 + ch] : 0.f;
  const float fx0 = (float)t.x0, fy0 = (float)t.y0;
  const float ex = fx0 + 1.f - t.ix, wx = t.ix - fx0, ey = fy0 + 1.f - t.iy, wy = t.iy - fy0;
  for (int ch = 0; ch < 3; ++ch) {
    rgb[ch] = c[0][ch] * w[0] + c[1][ch] * w[1] + c[2][ch] * w[2] + c[3][ch] * w[3];
    const float gix = (c[1][ch] - c[0][ch]) * ey + (c[3][ch] - c[2][ch]) * wy;      // as raster_backward_kernel's barycentric path
    const float giy = (c[2][ch] - c[0][ch]) * ex + (c[3][ch] - c[1][ch]) * wx;
    d_du[ch] = gix * t.mx; d_dv[ch] = giy * t.my;
  }
}

// z-clip of n faces (fv: n x 9 floats) as face_setup_kernel does it: per face the number of triangles (0/1/2), their vertices
// (n x 2 x 9) and the barycentric conversion matrices (n x 2 x 9, identity for an untouched face)
void hm_clip(const float* fv, int n, float z_clip, int persp, int* ntri, float* tri, float* conv) {
  for (int f = 0; f < n; ++f) {
    float a[3][3];
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) a[i][c] = fv[f * 9 + i * 3 + c];
    ClipResult r;
    clip_face(a, z_clip, persp != 0, r);
    ntri[f] = r.ntri;
    for (int k = 0; k < 2; ++k)
      for (int i = 0; i < 9; ++i) {
        tri[(f * 2 + k) * 9 + i] = k < r.ntri ? r.tri[k][i] : 0.f;
        conv[(f * 2 + k) * 9 + i] = k < r.ntri ? (r.clipped ? r.conv[k][i] : (i % 4 == 0 ? 1.f : 0.f)) : 0.f;
      }
  }
}

}  // extern "C"

extern "C" {
// a stream of n candidates of one pixel offered to the sorted shared-memory fragment list of raster_forward_kernel
// (fraglist_offer, dbw_fraglist.cuh), laid out as on the device: this pixel's column inside a [K][stride] array
int hm_topk(int K, int n, const float* pz, const int* slot, const float* sd, const int* neighbor, int* out_slot, float* out_sd) {
  const int stride = 7, col = 3;                 // any column of a wider array: neighbours must stay untouched
  float4* A = new float4[(size_t)K * stride];
  float* V = new float[(size_t)K * stride];
  for (int i = 0; i < K * stride; ++i) { A[i] = make_float4(-7.f, -7.f, -7.f, -7.f); V[i] = -7.f; }
  int cnt = 0;
  for (int i = 0; i < n; ++i)
    cnt = fraglist_offer(A + col, V + col, stride, cnt, K, pz[i], slot[i], i % 3, sd[i], fabsf(sd[i]), neighbor[i], 0.25f * i, 0.5f * i);
  int bad = 0;
  for (int k = 0; k < K; ++k) {
    const float4 e = A[k * stride + col];
    out_slot[k] = k < cnt ? (__float_as_int(e.y) & DBW_FRAG_SLOT_MASK) : -1;
    out_sd[k] = k < cnt ? e.z : 0.f;
    if (k < cnt) {                               // the payload travels with its key: edge id, u, v of the candidate that made it
      int src = -1;
      for (int i = 0; i < n; ++i) if (slot[i] == out_slot[k]) src = i;
      if (src < 0 || ((__float_as_int(e.y) >> DBW_FRAG_EDGE_SHIFT) & 3) != src % 3 || e.w != 0.25f * src || V[k * stride + col] != 0.5f * src || e.x != pz[src]) bad = 1;
    }
  }
  for (int i = 0; i < K * stride; ++i) if (i % stride != col && (A[i].x != -7.f || V[i] != -7.f)) bad = 2;
  delete[] A; delete[] V;
  return bad;
}
}  // extern "C"

extern "C" {
// rotation_6d_to_matrix and its backward (dbw_scene_math.cuh), n rotations
void hm_rot6d(const float* d6, int n, float* R) { for (int i = 0; i < n; ++i) rot6d(d6 + i * 6, R + i * 9); }
void hm_rot6d_backward(const float* d6, const float* gR, int n, float* gd6) {
  for (int i = 0; i < n; ++i) rot6d_backward(d6 + i * 6, gR + i * 9, gd6 + i * 6);
}
// unit-scale superquadric vertices of N blocks x Vb vertices (local_vertex): out (N, Vb, 3), aux (N, Vb, 6)
void hm_superquadric(const float* sq_eta, const float* sq_omega, const float* sq_eps, int N, int Vb, float ratio, float* out, float* aux) {
  GeomParams P;
  memset(&P, 0, sizeof(P));
  P.n_blocks = N; P.verts_per_block = Vb; P.sq_eta = sq_eta; P.sq_omega = sq_omega; P.sq_eps = sq_eps; P.ratio = ratio;
  for (int b = 0; b < N; ++b)
    for (int v = 0; v < Vb; ++v) local_vertex(P, b, v, out + (b * Vb + v) * 3, aux + (b * Vb + v) * 6);
}
}  // extern "C"

extern "C" {
// The tile binner's conservative triangle / rectangle test (tri_overlaps_rect), driven as raster_forward_kernel drives it:
// for every TS x TS pixel tile of an H x W image, the rectangle spanned by the tile's pixel centres, expanded by sqrt(blur).
// out[tile] = 1 if the face would be listed for the tile.
void hm_tile_overlap(const float* fv, int H, int W, int TS, float blur, int* out) {
  const float sb = sqrtf(blur);
  const f2 v0 = {fv[0], fv[1]}, v1 = {fv[3], fv[4]}, v2 = {fv[6], fv[7]};
  const int ntx = (W + TS - 1) / TS, nty = (H + TS - 1) / TS;
  for (int ty = 0; ty < nty; ++ty)
    for (int tx = 0; tx < ntx; ++tx) {
      const int xa = tx * TS, xb = (tx + 1) * TS - 1 < W - 1 ? (tx + 1) * TS - 1 : W - 1;
      const int ya = ty * TS, yb = (ty + 1) * TS - 1 < H - 1 ? (ty + 1) * TS - 1 : H - 1;
      // pixel xi maps to NDC pix_to_ndc(W - 1 - xi): decreasing in xi
      const float x_lo = pix_to_ndc(W - 1 - xb, W, H), x_hi = pix_to_ndc(W - 1 - xa, W, H);
      const float y_lo = pix_to_ndc(H - 1 - yb, H, W), y_hi = pix_to_ndc(H - 1 - ya, H, W);
      out[ty * ntx + tx] = tri_overlaps_rect(v0, v1, v2, x_lo - sb, x_hi + sb, y_lo - sb, y_hi + sb) ? 1 : 0;
    }
}

// tri_dist2_edge over an image: the distance (must equal tri_dist2 bit for bit), the edge it names, and the three segment
// distances, so that the test can check the edge realises the minimum with tri_dist_backward's tie order
void hm_dist_edge(const float* fv, int H, int W, float* dist, float* dist_ref, int* edge, float* seg) {
  const TriGeom t = make_tri(fv);
  for (int yi = 0; yi < H; ++yi)
    for (int xi = 0; xi < W; ++xi) {
      const int o = yi * W + xi;
      const f2 p = {pix_to_ndc(W - 1 - xi, W, H), pix_to_ndc(H - 1 - yi, H, W)};
      int e = -1;
      dist[o] = tri_dist2_edge(p, t, e);
      dist_ref[o] = tri_dist2(p, t);
      edge[o] = e;
      seg[o * 3] = seg_dist2(p, t.v0, t.v1, t.il01); seg[o * 3 + 1] = seg_dist2(p, t.v0, t.v2, t.il02); seg[o * 3 + 2] = seg_dist2(p, t.v1, t.v2, t.il12);
    }
}
}  // extern "C"
