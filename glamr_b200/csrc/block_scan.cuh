// CTA-wide in-place inclusive prefix scan over a strided global array (forward or reverse), any length.
#pragma once
#include "common.cuh"

namespace glamr {

constexpr int kScanThreads = 512;

// Inclusive scan (forward or reverse) of `count` floats at data[k * stride], in place, by one CTA of kScanThreads.
__device__ __forceinline__ void block_scan_inplace(float* data, int count, int stride, bool reverse, float* smem /*[kScanThreads/32 + 1]*/) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kScanThreads / 32;
  float carry = 0.0f;
  for (int base = 0; base < count; base += kScanThreads) {
    const int k = base + tid;
    const int idx = reverse ? (count - 1 - k) : k;
    float v = (k < count) ? data[(size_t)idx * stride] : 0.0f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) smem[wid] = v;
    __syncthreads();
    if (wid == 0) {
      float w = (lane < NW) ? smem[lane] : 0.0f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      if (lane < NW) smem[lane] = w;
    }
    __syncthreads();
    const float prefix = (wid > 0 ? smem[wid - 1] : 0.0f) + carry;
    v += prefix;
    if (k < count) data[(size_t)idx * stride] = v;
    const float total = smem[NW - 1];
    __syncthreads();
    carry += total;
  }
}


}  // namespace glamr
