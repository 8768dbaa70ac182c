// dbw_topk.cuh -- the K nearest fragments of one pixel, kept sorted in registers (raster_forward_kernel, dbw_render.cu).
// Order = the tuple order of PyTorch3D's CPU rasterizer queue (depth, then face index) folded into
// one 64-bit key; the two halves of a z-clipped quad exclude each other (Appendix A3).  Plain C++ apart from the
// __device__ markers: tests/host_math compiles it for the CPU and checks it against the oracle's queue.
#pragma once
#include "dbw_math.cuh"

// depth bits (positive floats order like their bit patterns) in the high word, triangle slot in the low word
__device__ __forceinline__ unsigned long long make_key(float pz, int slot) {
  return ((unsigned long long)__float_as_uint(pz + 0.f) << 32) | (unsigned)slot;
}

// Offer one candidate (depth pz, triangle slot, signed squared distance sd, |sd| = dist, slot of the other half of its
// z-clipped quad or -1) to the sorted list key[] / dk[] (empty entries: key = ~0).  Every index is a compile-time constant
// of a fully unrolled loop, so that key[] / dk[] stay in registers.
template <int K>
__device__ __forceinline__ void topk_offer(unsigned long long (&key)[K], float (&dk)[K], float pz, int slot, float sd, float dist,
                                           int neighbor) {
  if (neighbor >= 0) {
    // the other half of a z-clipped quad: only the half with the smaller |dist| may stay (A3)
    int found = -1; float d_found = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if ((int)(unsigned)key[k] == neighbor && key[k] != ~0ull) { found = k; d_found = dk[k]; }
    if (found >= 0) {
      if (dist < fabsf(d_found)) {
#pragma unroll
        for (int q = 0; q < K - 1; ++q) if (q >= found) { key[q] = key[q + 1]; dk[q] = dk[q + 1]; }
        key[K - 1] = ~0ull; dk[K - 1] = 0.f;
      } else return;
    }
  }
  const unsigned long long nk = make_key(pz, slot);
  if (nk >= key[K - 1]) return;
  // sorted insertion with static indexing
#pragma unroll
  for (int k = K - 1; k >= 1; --k) {
    const bool up = nk < key[k - 1];
    const bool here = !up && nk < key[k];
    key[k] = up ? key[k - 1] : (here ? nk : key[k]);
    dk[k] = up ? dk[k - 1] : (here ? sd : dk[k]);
  }
  if (nk < key[0]) { key[0] = nk; dk[0] = sd; }
}
