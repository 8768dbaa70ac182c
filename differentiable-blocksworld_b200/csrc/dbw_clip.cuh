// dbw_clip.cuh -- z-clipping of one projected face into 0, 1 or 2 triangles (PyTorch3D clip.py: clip_faces with z_clip_value
// only), used by face_setup_kernel (dbw_render.cu).  Plain C++ apart from the __device__ markers, so
// that tests/host_math can compile it for the CPU and check it against the oracle's clip_faces.
#pragma once
#include "dbw_math.cuh"

struct ClipResult {
  int ntri;            // 0, 1 or 2 triangles
  bool clipped;
  int i1;              // isolated vertex
  float w2, w3;
  float tri[2][9];     // (x,y,z) x 3
  float conv[2][9];    // rows = barycentrics of the clipped triangle's vertices in the original face
};

__device__ __forceinline__ void lerp_clip(const float* p1, const float* p2, float w, bool persp, float* out) {
  if (persp) {
    const float q1x = p1[0] * p1[2], q1y = p1[1] * p1[2], q2x = p2[0] * p2[2], q2y = p2[1] * p2[2];
    const float Px = q1x * (1.f - w) + q2x * w, Py = q1y * (1.f - w) + q2y * w, Pz = p1[2] * (1.f - w) + p2[2] * w;
    out[0] = Px / Pz; out[1] = Py / Pz; out[2] = Pz;
  } else {
    out[0] = p1[0] * (1.f - w) + p2[0] * w; out[1] = p1[1] * (1.f - w) + p2[1] * w; out[2] = p1[2] * (1.f - w) + p2[2] * w;
  }
}

__device__ __forceinline__ void set3(float* d, const float* s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }

__device__ void clip_face(const float a[3][3], float z_clip, bool persp, ClipResult& r) {
  int nb = 0, behind[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { behind[i] = (z_clip >= 0.f) && (a[i][2] < z_clip); nb += behind[i]; }
  r.clipped = false; r.i1 = 0; r.w2 = r.w3 = 0.f;
  if (nb == 0) {
    r.ntri = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) set3(&r.tri[0][i * 3], a[i]);
    return;
  }
  if (nb == 3) { r.ntri = 0; return; }
  r.clipped = true;
  int i1 = 0;
  if (nb == 2) { for (int i = 0; i < 3; ++i) if (!behind[i]) i1 = i; }   // the single vertex in front
  else         { for (int i = 0; i < 3; ++i) if (behind[i]) i1 = i; }    // the single vertex behind
  const int i2 = (i1 + 1) % 3, i3 = (i1 + 2) % 3;
  const float* p1 = a[i1]; const float* p2 = a[i2]; const float* p3 = a[i3];
  const float w2 = (p1[2] - z_clip) / (p1[2] - p2[2]);
  const float w3 = (p1[2] - z_clip) / (p1[2] - p3[2]);
  float p4[3], p5[3], b1[3] = {0, 0, 0}, b2[3] = {0, 0, 0}, b3[3] = {0, 0, 0}, b4[3] = {0, 0, 0}, b5[3] = {0, 0, 0};
  lerp_clip(p1, p2, w2, persp, p4);
  lerp_clip(p1, p3, w3, persp, p5);
  b1[i1] = 1.f; b2[i2] = 1.f; b3[i3] = 1.f;
  b4[i1] = 1.f - w2; b4[i2] = w2;
  b5[i1] = 1.f - w3; b5[i3] = w3;
  r.i1 = i1; r.w2 = w2; r.w3 = w3;
  if (nb == 2) {          // (p4, p5, p1)
    r.ntri = 1;
    set3(&r.tri[0][0], p4); set3(&r.tri[0][3], p5); set3(&r.tri[0][6], p1);
    set3(&r.conv[0][0], b4); set3(&r.conv[0][3], b5); set3(&r.conv[0][6], b1);
  } else {                // (p4, p2, p5) and (p5, p2, p3)
    r.ntri = 2;
    set3(&r.tri[0][0], p4); set3(&r.tri[0][3], p2); set3(&r.tri[0][6], p5);
    set3(&r.conv[0][0], b4); set3(&r.conv[0][3], b2); set3(&r.conv[0][6], b5);
    set3(&r.tri[1][0], p5); set3(&r.tri[1][3], p2); set3(&r.tri[1][6], p3);
    set3(&r.conv[1][0], b5); set3(&r.conv[1][3], b2); set3(&r.conv[1][6], b3);
  }
}
