// dbw_fraglist.cuh -- the K nearest fragments of one pixel as a SORTED LIST IN SHARED MEMORY (raster_forward_kernel).
//
// Round 1 kept the list in registers (64-bit keys, fully unrolled sorted insertion: ~110 instructions per insertion and
// 3K registers per thread, which capped the K=10 kernel at 38 % occupancy and made K=25 spill).  The list now lives in
// shared memory, one column per thread (entry k of thread t at index k*NT + t: conflict-free), sorted on insertion; an
// insertion is one 128-bit + one 32-bit shared load/store pair per shifted entry, the register cost is the entry count.
// Each entry carries everything shading needs -- no geometry is re-derived for the kept fragments:
//     A[k] = (pz, bits, signed squared distance, u)      V[k] = v
//     bits = triangle slot (20 bits) | closest edge (2 bits: 0 = v0v1, 1 = v0v2, 2 = v1v2) << 20 | outside (sd >= 0) << 22
//            | texture map of the face (9 bits) << 23
// Order = the tuple order of PyTorch3D's CPU rasterizer queue (depth, then face index); the two
// halves of a z-clipped quad exclude each other (Appendix A3).  Plain C++ apart from the __device__ markers:
// tests/host_math compiles it for the CPU and checks it against the oracle's queue.
#pragma once
#include "dbw_math.cuh"

#define DBW_FRAG_SLOT_MASK 0x000fffff
#define DBW_FRAG_EDGE_SHIFT 20
#define DBW_FRAG_OUTSIDE_BIT (1 << 22)
#define DBW_FRAG_MAP_SHIFT 23
#define DBW_FRAG_MAX_MAPS 512

// (pz, slot) < key of entry e ?   depths are >= 0, so their bit patterns order like the values
__device__ __forceinline__ bool frag_key_less(unsigned pz_bits, int slot, float4 e) {
  const unsigned eb = __float_as_uint(e.x);
  return pz_bits < eb || (pz_bits == eb && slot < (__float_as_int(e.y) & DBW_FRAG_SLOT_MASK));
}

// Offer one candidate to the list of `n` entries (capacity K); returns the new entry count.
// A / V point at THIS thread's column; consecutive entries are `stride` elements apart.
__device__ __forceinline__ int fraglist_offer(float4* A, float* V, int stride, int n, int K, float pz, int slot, int edge,
                                              float sd, float dist, int neighbor, float u, float v, int map_id = 0) {
  const unsigned pzb = __float_as_uint(pz + 0.f);
  if (neighbor >= 0) {
    // the other half of a z-clipped quad: only the half with the smaller |dist| may stay (A3)
    for (int i = 0; i < n; ++i) {
      const float4 e = A[i * stride];
      if ((__float_as_int(e.y) & DBW_FRAG_SLOT_MASK) != neighbor) continue;
      if (!(dist < fabsf(e.z))) return n;
      for (int q = i; q < n - 1; ++q) { A[q * stride] = A[(q + 1) * stride]; V[q * stride] = V[(q + 1) * stride]; }
      --n;
      break;
    }
  }
  if (n == K) {
    if (!frag_key_less(pzb, slot, A[(K - 1) * stride])) return n;
    n = K - 1;                                   // the farthest entry is dropped
  }
  int i = n;
  while (i > 0) {
    const float4 e = A[(i - 1) * stride];
    if (!frag_key_less(pzb, slot, e)) break;
    A[i * stride] = e; V[i * stride] = V[(i - 1) * stride];
    --i;
  }
  A[i * stride] = make_float4(pz, __int_as_float(slot | (edge << DBW_FRAG_EDGE_SHIFT) | (sd < 0.f ? 0 : DBW_FRAG_OUTSIDE_BIT) |
                                                 (map_id << DBW_FRAG_MAP_SHIFT)), sd, u);
  V[i * stride] = v;
  return n + 1;
}
