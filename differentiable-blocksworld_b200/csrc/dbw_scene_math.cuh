// dbw_scene_math.cuh -- per-vertex math of the fused scene-geometry kernels (dbw_scene.cu): signed power, 6D rotation
// (forward / backward) and the unit-scale superquadric vertex (src/utils/superquadric.py:10-14, src/model/dbw.py:299-352).
// Plain C++ apart from the __device__ markers: tests/host_math compiles it for the CPU and checks it against torch.
#pragma once
#include <math.h>

struct Mat3 { float m[9]; };

__device__ __forceinline__ float spow(float x, float e) {        // sign(x) * |x|^e   (utils/pytorch.py:31-32)
  const float a = fabsf(x);
  if (a == 0.f) return 0.f;
  return copysignf(powf(a, e), x);
}

// rotation_6d_to_matrix: rows b1 = a1/|a1|, b2 = normalize(a2 - (b1.a2) b1), b3 = b1 x b2
__device__ __forceinline__ void rot6d(const float* d6, float* R) {
  const float n1 = fmaxf(sqrtf(d6[0] * d6[0] + d6[1] * d6[1] + d6[2] * d6[2]), 1e-12f);
  const float b1x = d6[0] / n1, b1y = d6[1] / n1, b1z = d6[2] / n1;
  const float d = b1x * d6[3] + b1y * d6[4] + b1z * d6[5];
  const float ux = d6[3] - d * b1x, uy = d6[4] - d * b1y, uz = d6[5] - d * b1z;
  const float n2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);
  const float b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  R[0] = b1x; R[1] = b1y; R[2] = b1z; R[3] = b2x; R[4] = b2y; R[5] = b2z;
  R[6] = b1y * b2z - b1z * b2y; R[7] = b1z * b2x - b1x * b2z; R[8] = b1x * b2y - b1y * b2x;
}

__device__ __forceinline__ void rot6d_backward(const float* d6, const float* gR, float* gd6) {
  const float n1 = fmaxf(sqrtf(d6[0] * d6[0] + d6[1] * d6[1] + d6[2] * d6[2]), 1e-12f);
  const float b1[3] = {d6[0] / n1, d6[1] / n1, d6[2] / n1};
  const float d = b1[0] * d6[3] + b1[1] * d6[4] + b1[2] * d6[5];
  const float u[3] = {d6[3] - d * b1[0], d6[4] - d * b1[1], d6[5] - d * b1[2]};
  const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  float gb1[3] = {gR[0], gR[1], gR[2]}, gb2[3] = {gR[3], gR[4], gR[5]};
  const float gb3[3] = {gR[6], gR[7], gR[8]};
  // b3 = b1 x b2 :  d/db1 = b2 x g3 ,  d/db2 = g3 x b1
  gb1[0] += b2[1] * gb3[2] - b2[2] * gb3[1]; gb1[1] += b2[2] * gb3[0] - b2[0] * gb3[2]; gb1[2] += b2[0] * gb3[1] - b2[1] * gb3[0];
  gb2[0] += gb3[1] * b1[2] - gb3[2] * b1[1]; gb2[1] += gb3[2] * b1[0] - gb3[0] * b1[2]; gb2[2] += gb3[0] * b1[1] - gb3[1] * b1[0];
  // b2 = u / |u|
  const float gdot2 = gb2[0] * b2[0] + gb2[1] * b2[1] + gb2[2] * b2[2];
  const float gu[3] = {(gb2[0] - gdot2 * b2[0]) / n2, (gb2[1] - gdot2 * b2[1]) / n2, (gb2[2] - gdot2 * b2[2]) / n2};
  // u = a2 - (b1.a2) b1
  const float gub1 = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
  gd6[3] = gu[0] - gub1 * b1[0]; gd6[4] = gu[1] - gub1 * b1[1]; gd6[5] = gu[2] - gub1 * b1[2];
  gb1[0] += -d * gu[0] - gub1 * d6[3]; gb1[1] += -d * gu[1] - gub1 * d6[4]; gb1[2] += -d * gu[2] - gub1 * d6[5];
  // b1 = a1 / |a1|
  const float gdot1 = gb1[0] * b1[0] + gb1[1] * b1[1] + gb1[2] * b1[2];
  gd6[0] = (gb1[0] - gdot1 * b1[0]) / n1; gd6[1] = (gb1[1] - gdot1 * b1[1]) / n1; gd6[2] = (gb1[2] - gdot1 * b1[2]) / n1;
}

struct GeomParams {
  int n_blocks, verts_per_block, n_ground_verts;
  const float* sq_eta; const float* sq_omega;        // (N, Vb)
  const float* sq_eps; const float* S; const float* R6; const float* T;   // (N,2) (N,3) (N,6) (N,3)
  const float* ground_verts; const float* R6g; const float* Tg;          // (Vg,3) (6) (3)
  float ratio, scale_min, S_world;
  Mat3 R_world; float T_world[3];
};

// unit-scale vertex of primitive `prim` (block: parametric superquadric * ratio; ground: its static plane vertex)
__device__ __forceinline__ void local_vertex(const GeomParams& P, int prim, int v, float* u, float* aux /*ce,se,co,so,e1,e2*/) {
  if (prim < P.n_blocks) {
    const float e1 = 1.8f / (1.f + expf(-P.sq_eps[prim * 2])) + 0.1f;
    const float e2 = 1.8f / (1.f + expf(-P.sq_eps[prim * 2 + 1])) + 0.1f;
    const float eta = P.sq_eta[prim * P.verts_per_block + v], om = P.sq_omega[prim * P.verts_per_block + v];
    const float ce = spow(cosf(eta), e1), se = spow(sinf(eta), e1), co = spow(cosf(om), e2), so = spow(sinf(om), e2);
    u[0] = ce * so * P.ratio; u[1] = se * P.ratio; u[2] = ce * co * P.ratio;
    aux[0] = ce; aux[1] = se; aux[2] = co; aux[3] = so; aux[4] = e1; aux[5] = e2;
  } else {
    u[0] = P.ground_verts[v * 3]; u[1] = P.ground_verts[v * 3 + 1]; u[2] = P.ground_verts[v * 3 + 2];
  }
}
