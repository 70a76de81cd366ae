// Learned-prior inference on sm_100a: the motion infiller (CVAE, transformer encoder/decoder over 50-frame windows,
// motion_infiller/models/motion_infiller_vae.py:22-123,252-421,564-632) and the trajectory predictor (CVAE, MLP +
// 2-layer bidirectional LSTM, traj_pred/models/traj_pred_vae.py:20-92,202-333; lib/models/{mlp,rnn,pos_encoding}.py).
// Every Linear (QKV / out-proj / FFN / MLP / LSTM input projections) runs on the tensor cores: tcgen05.mma kind::tf32
// with a 3xTF32 split and the accumulator in TMEM (gemm_tf32x3_tcgen05_kernel); LayerNorm, softmax attention (S <= 64,
// shared memory) and the LSTM recurrence (W_hh resident in registers + shared memory for the whole sequence) are FP32
// SIMT kernels.  Outputs match the reference's fp32 networks to <= 1e-4 (tests/golden/nets.npz).
// Weights are addressed by their reference state-dict names so Lightning checkpoints map 1:1.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "block_scan.cuh"
#include "glamr_math.cuh"

namespace glamr {

// ------------------------------------------------------------------------------------------------ SGEMM  Y = act(X W^T + b)
// X [M,K] (row stride ldx), W [N,K], Y [M,N] (row stride ldy).  64x64x16 tiles, 256 threads, 4x4 micro-tiles.
constexpr int GT = 64, GK = 16;
template <int ACT>   // 0 none, 1 relu
__global__ void __launch_bounds__(256) gemm_bias_act_kernel(int M, int N, int K, const float* __restrict__ X, int ldx,
                                                            const float* __restrict__ W, const float* __restrict__ bias,
                                                            const float* __restrict__ bias2, float* __restrict__ Y, int ldy) {
  __shared__ float Xs[GK][GT + 4];
  __shared__ float Ws[GK][GT + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int tx = tid & 15, ty = tid >> 4;       // 16 x 16 threads, each 4 (m) x 4 (n)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const int m = m0 + lr, n = n0 + lr;
      Xs[lk + q][lr] = (m < M && k < K) ? X[(size_t)m * ldx + k] : 0.0f;
      Ws[lk + q][lr] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.0f) + (bias2 ? bias2[n] : 0.0f);
      if (ACT == 1) v = fmaxf(v, 0.0f);
      Y[(size_t)m * ldy + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 GEMM (3xTF32)
// Y = act(X W^T + b) on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 with the accumulator in TMEM.
// FP32 accuracy is kept with the 3xTF32 split  x = hi + lo (hi = tf32(x), lo = tf32(x - hi)):
//   X W^T ~= Xhi Whi^T + Xlo Whi^T + Xhi Wlo^T     (error ~2^-21 relative: the 1e-4 parity bar of the infilled pose holds)
// CTA = 128 threads, tile 128 (M) x 128 (N), K step 32.  Both operands are K-major (X [M,K] and W [N,K] row-major), written
// by the CTA into shared memory in the canonical no-swizzle UMMA layout (8-row x 16-byte core matrices: element (r,k) at
// ((k/4)*128 + r)*16 + (k%4)*4 bytes => LBO = 2048 B between K groups, SBO = 128 B between 8-row groups), made visible to
// the async proxy with fence.proxy.async, multiplied by one elected thread (12 MMAs per K step), completion tracked with
// tcgen05.commit -> mbarrier; the epilogue reads the 128x128 fp32 accumulator with tcgen05.ld (32x32b.x32), adds the
// bias, applies the activation and stores.
constexpr int TCM = 128, TCK = 32;                              // N tile (NT) is a template parameter: 128, or 32 for one-tile-high problems
constexpr int kTcATileFloats = TCM * TCK;                       // 4096 floats = 16 KB (hi or lo of the X tile)

__device__ __forceinline__ float4 split_tf32_4(const float4 v, float4& lo) {
  float4 hi;
  split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
  return hi;
}

#ifdef GLAMR_EXPERIMENT
__device__ int g_tc_dbg = 0;   // experiment switches (tools/tc_gemm_exp.py, env GLAMR_TC_DEBUG); not in the release build
#endif
constexpr int kTcThreads = 256;
constexpr int kTcStages = 2;
template <int NT>
struct TcCfg {
  static constexpr int kBTileFloats = NT * TCK;
  static constexpr int kStageFloats = 2 * kTcATileFloats + 2 * kBTileFloats;            // Xhi | Xlo | Whi | Wlo
  static constexpr int kXVec = (TCM * TCK / 4) / kTcThreads;                            // 16-byte loads per thread and K step
  static constexpr int kWVec = (NT * TCK / 4) / kTcThreads;
  static constexpr size_t kSmemBytes = (size_t)kTcStages * kStageFloats * sizeof(float) + 64;
  static_assert(kWVec >= 1, "W tile smaller than one 16-byte load per thread");
};

// one K step (32 columns) of the 128-row X tile and the NT-row W tile, in flight in registers
template <int NT>
struct TcRegs {
  float4 x[TcCfg<NT>::kXVec], w[TcCfg<NT>::kWVec];
};
template <bool VEC>
__device__ __forceinline__ float4 tc_load_row4(const float* __restrict__ P, int ld, int row, int nrows, int k, int K) {
  if (VEC) {
    if (row < nrows && k < K) return __ldg(reinterpret_cast<const float4*>(P + (size_t)row * ld + k));
    return make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float a[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) a[q] = (row < nrows && k + q < K) ? P[(size_t)row * ld + k + q] : 0.0f;
  return make_float4(a[0], a[1], a[2], a[3]);
}
template <int NT, bool VEC, bool WIMG = false>
__device__ __forceinline__ void tc_load_tiles(TcRegs<NT>& r, int tid, int M, int N, int K, const float* __restrict__ X, int ldx,
                                              const float* __restrict__ W, int m0, int n0, int k0) {
#pragma unroll
  for (int i = 0; i < TcCfg<NT>::kXVec; ++i) {
    const int idx = tid + i * kTcThreads;
    r.x[i] = tc_load_row4<VEC>(X, ldx, m0 + (idx & (TCM - 1)), M, k0 + (idx / TCM) * 4, K);      // row, 16-byte K group (0..7)
  }
  if (!WIMG) {
#pragma unroll
    for (int i = 0; i < TcCfg<NT>::kWVec; ++i) {
      const int idx = tid + i * kTcThreads;
      r.w[i] = tc_load_row4<VEC>(W, K, n0 + (idx & (NT - 1)), N, k0 + (idx / NT) * 4, K);
    }
  }
}
// split into tf32 hi / lo and store as K-major 8x16-byte core matrices (the layout umma_desc_kmajor_noswizzle describes)
template <int NT, bool WIMG = false>
__device__ __forceinline__ void tc_store_tiles(float* st, int tid, const TcRegs<NT>& r) {
  float* Ahi = st;
  float* Alo = st + kTcATileFloats;
  float* Bhi = st + 2 * kTcATileFloats;
  float* Blo = Bhi + TcCfg<NT>::kBTileFloats;
#pragma unroll
  for (int i = 0; i < TcCfg<NT>::kXVec; ++i) {
    const int idx = tid + i * kTcThreads;
    const int off = ((idx / TCM) * TCM + (idx & (TCM - 1))) * 4;
    float4 lo;
    const float4 hi = split_tf32_4(r.x[i], lo);
    *reinterpret_cast<float4*>(Ahi + off) = hi;
    *reinterpret_cast<float4*>(Alo + off) = lo;
  }
  if (!WIMG) {
#pragma unroll
    for (int i = 0; i < TcCfg<NT>::kWVec; ++i) {
      const int idx = tid + i * kTcThreads;
      const int off = ((idx / NT) * NT + (idx & (NT - 1))) * 4;
      float4 lo;
      const float4 hi = split_tf32_4(r.w[i], lo);
      *reinterpret_cast<float4*>(Bhi + off) = hi;
      *reinterpret_cast<float4*>(Blo + off) = lo;
    }
  }
}

// ---- weights as a pre-split operand image -------------------------------------------------------------------------
// A weight matrix W [N,K] is constant between glamr_net_set_tensor calls: it is split into tf32 hi / lo and tiled ONCE into the
// shared-memory image the MMA reads, [N tile][K step of 32][hi | lo][8 K groups][NT rows][4] (zero padded), so that the kernel
// fetches the W half of a pipeline stage with one 1-D bulk TMA copy instead of loading, splitting and storing it with all threads
// on every launch.
template <int NT>
__global__ void build_w_image_kernel(const float* __restrict__ W, int N, int K, int ksteps, float* __restrict__ img) {
  const size_t total = (size_t)((N + NT - 1) / NT) * ksteps * NT * TCK;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e % TCK);
    const int r = (int)((e / TCK) % NT);
    const size_t tile = e / ((size_t)TCK * NT);              // tn * ksteps + ks
    const int ks = (int)(tile % ksteps), tn = (int)(tile / ksteps);
    const int n = tn * NT + r, k = ks * TCK + kk;
    const float v = (n < N && k < K) ? W[(size_t)n * K + k] : 0.0f;
    float hi, lo;
    split_tf32(v, hi, lo);
    float* q = img + tile * (2 * TcCfg<NT>::kBTileFloats) + ((size_t)(kk >> 2) * NT + r) * 4 + (kk & 3);
    q[0] = hi;
    q[TcCfg<NT>::kBTileFloats] = lo;
  }
}

// 256 threads; two shared-memory stages: while the tensor core works on stage s (12 UTCHMMA per K step, tracked by
// tcgen05.commit -> mbarrier[s]) all threads split and store K step it+1 into stage s^1 and already have the global
// loads of step it+2 in flight in registers, so the L2 latency never sits on the critical path of these small GEMMs.
// NT = 128: 128x128 tiles (one CTA per SM).  NT = 32: 128x32 tiles for problems one tile high (M <= 128, a single
// 120-frame window): 4x more CTAs, each with a quarter of the W traffic, split work and epilogue.
// WIMG: W points at the pre-split operand image of the weight (build_w_image_kernel) and arrives by bulk TMA (wbar[s]).
template <int ACT, bool VEC, int NT, bool WIMG>
__global__ void __launch_bounds__(kTcThreads) gemm_tf32x3_tcgen05_kernel(int M, int N, int K, const float* __restrict__ X, int ldx,
                                                                         const float* __restrict__ W, const float* __restrict__ bias,
                                                                         const float* __restrict__ bias2, float* __restrict__ Y, int ldy) {
  using Cfg = TcCfg<NT>;
  extern __shared__ __align__(128) unsigned char tc_smem[];
  float* stage0 = reinterpret_cast<float*>(tc_smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage0 + kTcStages * Cfg::kStageFloats);   // [2] MMA completion per stage
  uint64_t* wbar = bar + 2;                                                              // [2] weight image landed (WIMG)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TCM, n0 = blockIdx.x * NT;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const int nk = (K + TCK - 1) / TCK;
  constexpr uint32_t kWBytes = 2 * Cfg::kBTileFloats * sizeof(float);
  const float* wimg = W + (size_t)blockIdx.x * nk * (2 * Cfg::kBTileFloats);            // this column tile's K steps, contiguous
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    mbar_fence_init();
    if (WIMG) {
      mbar_expect_tx(&wbar[0], kWBytes);
      tma_bulk_g2s(stage0 + 2 * kTcATileFloats, wimg, kWBytes, &wbar[0]);
    }
  }
  const int dbg = GLAMR_DBG(g_tc_dbg);
  TcRegs<NT> regs;
  tc_load_tiles<NT, VEC, WIMG>(regs, tid, M, N, K, X, ldx, W, m0, n0, 0);
  tc_store_tiles<NT, WIMG>(stage0, tid, regs);
  if (nk > 1) tc_load_tiles<NT, VEC, WIMG>(regs, tid, M, N, K, X, ldx, W, m0, n0, TCK);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, K-major A/B, N>>3 [17,23), M>>4 [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);

  for (int it = 0; it < nk; ++it) {
    const int s = it & 1;
    float* st = stage0 + s * Cfg::kStageFloats;
    if (tid == 0) {
      if (WIMG) mbar_wait(&wbar[s], (it >> 1) & 1);                       // this stage's weight image has landed
#pragma unroll
      for (int k8 = 0; k8 < ((dbg & 1) ? 0 : TCK / 8); ++k8) {           // one tf32 MMA consumes K = 8 (two 16-byte K groups)
        const size_t koa = (size_t)k8 * 2 * TCM * 4, kob = (size_t)k8 * 2 * NT * 4;   // floats
        const uint64_t dah = umma_desc_kmajor_noswizzle(st + koa, TCM), dal = umma_desc_kmajor_noswizzle(st + kTcATileFloats + koa, TCM);
        const uint64_t dbh = umma_desc_kmajor_noswizzle(st + 2 * kTcATileFloats + kob, NT);
        const uint64_t dbl = umma_desc_kmajor_noswizzle(st + 2 * kTcATileFloats + Cfg::kBTileFloats + kob, NT);
        umma_tf32(tmem_d, dah, dbh, idesc, (it > 0 || k8 > 0) ? 1u : 0u);
        umma_tf32(tmem_d, dal, dbh, idesc, 1u);
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
      }
      // arrive on mbarrier[s] when every MMA issued so far has completed (implies fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[s])) : "memory");
    }
    if (it + 1 < nk) {
      // stage s^1 was consumed by the MMAs of step it-1: wait for their commit, then refill it while step `it` computes
      if (it >= 1) mbar_wait(&bar[s ^ 1], ((it - 1) >> 1) & 1);
      if (WIMG && tid == 0) {                                            // stage s^1 is free: fetch the weight image of step it+1
        mbar_expect_tx(&wbar[s ^ 1], kWBytes);
        tma_bulk_g2s(stage0 + (s ^ 1) * Cfg::kStageFloats + 2 * kTcATileFloats, wimg + (size_t)(it + 1) * (2 * Cfg::kBTileFloats), kWBytes, &wbar[s ^ 1]);
      }
      if (!(dbg & 2)) tc_store_tiles<NT, WIMG>(stage0 + (s ^ 1) * Cfg::kStageFloats, tid, regs);
      if (it + 2 < nk && !(dbg & 8)) tc_load_tiles<NT, VEC, WIMG>(regs, tid, M, N, K, X, ldx, W, m0, n0, (it + 2) * TCK);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
  }
  mbar_wait(&bar[(nk - 1) & 1], ((nk - 1) >> 1) & 1);      // all MMAs done: the accumulator is final
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: a warp may read TMEM lanes (= rows) 32 (w % 4) .. +31; 32-column chunks alternate between warps 0-3 and 4-7
  if (!(dbg & 4)) {
    const int wq = warp & 3, wh = warp >> 2;
    const int m = m0 + wq * 32 + lane;
#pragma unroll 1
    for (int cc = wh; cc < NT / 32; cc += 2) {
      if (n0 + cc * 32 >= N) break;
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cc * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
          "%28, %29, %30, %31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int nb = n0 + cc * 32;
      if (m < M && nb + 32 <= N && (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
        float4* yrow = reinterpret_cast<float4*>(Y + (size_t)m * ldy + nb);          // this lane's 128 contiguous bytes
        const bool bvec = ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(bias2)) & 15) == 0 && (nb & 3) == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float bsum[4] = {0.f, 0.f, 0.f, 0.f};
          if (bvec) {
            if (bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias + nb + j)); bsum[0] += t.x; bsum[1] += t.y; bsum[2] += t.z; bsum[3] += t.w; }
            if (bias2) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias2 + nb + j)); bsum[0] += t.x; bsum[1] += t.y; bsum[2] += t.z; bsum[3] += t.w; }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) bsum[q] = (bias ? __ldg(bias + nb + j + q) : 0.0f) + (bias2 ? __ldg(bias2 + nb + j + q) : 0.0f);
          }
          float o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q] = __uint_as_float(v[j + q]) + bsum[q];
            if (ACT == 1) o[q] = fmaxf(o[q], 0.0f);
          }
          yrow[j >> 2] = make_float4(o[0], o[1], o[2], o[3]);
        }
      } else if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nb + j;
          if (n < N) {
            float o = __uint_as_float(v[j]) + (bias ? bias[n] : 0.0f) + (bias2 ? bias2[n] : 0.0f);
            if (ACT == 1) o = fmaxf(o, 0.0f);
            Y[(size_t)m * ldy + n] = o;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(NT));
}

static int g_gemm_mode = 1;   // 1 = tcgen05 3xTF32 for the transformer (default), 0 = FP32 SIMT everywhere (A/B verification)
// The trajectory predictor (MLP + LSTM, M = T*B rows, outputs integrated over T frames by the trajectory codec) stays on the
// FP32 SIMT GEMM: its matrices are launch-latency sized and its per-frame heading error accumulates through the prefix sum.
struct ScopedFp32Gemm {
  int saved;
  ScopedFp32Gemm() : saved(g_gemm_mode) { g_gemm_mode = 0; }
  ~ScopedFp32Gemm() { g_gemm_mode = saved; }
};

// ---- cache of weight operand images, keyed by (device pointer, N, K, tile width); cleared when a network's tensors change
struct WImgKey {
  const float* w; int N, K, NT;
  bool operator<(const WImgKey& o) const { return w != o.w ? w < o.w : (N != o.N ? N < o.N : (K != o.K ? K < o.K : NT < o.NT)); }
};
static std::map<WImgKey, float*> g_wimg;
static int g_wimg_enabled = -1;     // GLAMR_NET_WIMG=0|1 (default: GLAMR_DEFAULT_NET_WIMG)
static void wimg_clear() {
  for (auto& kv : g_wimg) cudaFree(kv.second);
  g_wimg.clear();
}
template <int NT>
static int wimg_get(cudaStream_t s, const float* W, int N, int K, const float** out) {
  const WImgKey key{W, N, K, NT};
  auto it = g_wimg.find(key);
  if (it != g_wimg.end()) { *out = it->second; return GLAMR_OK; }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  if (cap != cudaStreamCaptureStatusNone) { *out = nullptr; return GLAMR_OK; }      // never allocate while a graph is being captured
  const int ksteps = (K + TCK - 1) / TCK, tiles = (N + NT - 1) / NT;
  float* img = nullptr;
  GLAMR_CUDA_TRY(cudaMalloc(&img, (size_t)tiles * ksteps * 2 * TcCfg<NT>::kBTileFloats * sizeof(float)));
  const size_t total = (size_t)tiles * ksteps * NT * TCK;
  build_w_image_kernel<NT><<<(unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, s>>>(W, N, K, ksteps, img);
  GLAMR_LAUNCH_CHECK();
  g_wimg[key] = img;
  *out = img;
  return GLAMR_OK;
}

template <int NT>
static int gemm_tc_launch(cudaStream_t s, int M, int N, int K, const float* X, int ldx, const float* W, const float* b, const float* b2,
                          float* Y, int ldy, int act) {
  static bool attr = false;
  constexpr size_t smem = TcCfg<NT>::kSmemBytes;
  if (!attr) {
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<0, true, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<1, true, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<0, false, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<1, false, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<0, true, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<1, true, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<0, false, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<1, false, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  if (g_wimg_enabled < 0) {
    const char* e = getenv("GLAMR_NET_WIMG");
    g_wimg_enabled = e ? atoi(e) : GLAMR_DEFAULT_NET_WIMG;
  }
  dim3 grid((N + NT - 1) / NT, (M + TCM - 1) / TCM);
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)X | (uintptr_t)W) % 16 == 0);
  const float* img = nullptr;
  if (g_wimg_enabled) {
    const int rc = wimg_get<NT>(s, W, N, K, &img);
    if (rc) return rc;
  }
  auto go = [&](auto kernel, const float* wptr) {
    kernel<<<grid, kTcThreads, smem, s>>>(M, N, K, X, ldx, wptr, b, b2, Y, ldy);
  };
  if (img) {
    const bool xvec = (K % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
    if (xvec) { if (act == 1) go(gemm_tf32x3_tcgen05_kernel<1, true, NT, true>, img); else go(gemm_tf32x3_tcgen05_kernel<0, true, NT, true>, img); }
    else { if (act == 1) go(gemm_tf32x3_tcgen05_kernel<1, false, NT, true>, img); else go(gemm_tf32x3_tcgen05_kernel<0, false, NT, true>, img); }
  } else if (vec) {
    if (act == 1) go(gemm_tf32x3_tcgen05_kernel<1, true, NT, false>, W); else go(gemm_tf32x3_tcgen05_kernel<0, true, NT, false>, W);
  } else {
    if (act == 1) go(gemm_tf32x3_tcgen05_kernel<1, false, NT, false>, W); else go(gemm_tf32x3_tcgen05_kernel<0, false, NT, false>, W);
  }
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}


// ------------------------------------------------------------------------------------------------ skinny GEMM (M <= 256)
// The prior runs at batch 1-4: its linear layers are [50 B x K] x [K x N] with K, N <= 512 -- a few MFLOP each, ~450 of them
// in dependent order per sequence, so the figure of merit is the LATENCY of one launch, not its throughput.  One warp owns an
// 8 x 4 output tile and splits K across its lanes: every lane issues all its 16-byte loads of the 8 X rows and 4 W rows at once
// (no shared memory, no block barrier, one global round trip), accumulates 32 partial dot products in FP32 and the warp folds
// them with a 31-shuffle transpose-reduction that leaves output (r, c) on lane 4 r + c.  ~2 us per launch against ~12 us for the
// one-tile tcgen05 kernel at M = 50 (profiles/init_breakdown_r02d_p1.txt), exact FP32 FMA arithmetic.
constexpr int kSkinnyMaxM = 256;
template <int ACT, bool VEC>
__global__ void __launch_bounds__(128) gemm_skinny_kernel(int M, int N, int K, const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ bias2, float* __restrict__ Y, int ldy) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int m0 = blockIdx.y * 8, n0 = (blockIdx.x * 4 + wid) * 4;
  if (n0 >= N) return;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
  const float* xr[8];
  const float* wr[4];
#pragma unroll
  for (int r = 0; r < 8; ++r) xr[r] = X + (size_t)min(m0 + r, M - 1) * ldx;      // rows / columns past the edge are clamped, never stored
#pragma unroll
  for (int c = 0; c < 4; ++c) wr[c] = W + (size_t)min(n0 + c, N - 1) * K;
  if (VEC) {
#pragma unroll 2
    for (int k = 4 * lane; k < K; k += 128) {
      float4 xv[8], wv[4];
#pragma unroll
      for (int r = 0; r < 8; ++r) xv[r] = *reinterpret_cast<const float4*>(xr[r] + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(wr[c] + k);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          acc[r * 4 + c] = fmaf(xv[r].w, wv[c].w, fmaf(xv[r].z, wv[c].z, fmaf(xv[r].y, wv[c].y, fmaf(xv[r].x, wv[c].x, acc[r * 4 + c]))));
    }
  } else {
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      float xv[8], wv[4];
#pragma unroll
      for (int r = 0; r < 8; ++r) xv[r] = xr[r][k];
#pragma unroll
      for (int c = 0; c < 4; ++c) wv[c] = wr[c][k];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r * 4 + c] = fmaf(xv[r], wv[c], acc[r * 4 + c]);
    }
  }
  // transpose-reduce: after the step with offset o a lane keeps the half of its values whose index has bit o equal to its own lane bit
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = up ? acc[i + o] : acc[i];
      const float send = up ? acc[i] : acc[i + o];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  const int m = m0 + (lane >> 2), n = n0 + (lane & 3);
  if (m < M && n < N) {
    float v = acc[0] + (bias ? bias[n] : 0.0f) + (bias2 ? bias2[n] : 0.0f);
    if (ACT == 1) v = fmaxf(v, 0.0f);
    Y[(size_t)m * ldy + n] = v;
  }
}

static int gemm_skinny_launch(cudaStream_t s, int M, int N, int K, const float* X, int ldx, const float* W, const float* b, const float* b2,
                              float* Y, int ldy, int act) {
  const dim3 grid((N + 15) / 16, (M + 7) / 8);
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)X | (uintptr_t)W) % 16 == 0);
  if (vec) {
    if (act == 1) gemm_skinny_kernel<1, true><<<grid, 128, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
    else gemm_skinny_kernel<0, true><<<grid, 128, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
  } else {
    if (act == 1) gemm_skinny_kernel<1, false><<<grid, 128, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
    else gemm_skinny_kernel<0, false><<<grid, 128, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
  }
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}

static int g_skinny = -1;       // GLAMR_NET_SKINNY=0 sends the small problems to the tile kernels again (A/B runs)
static int gemm(cudaStream_t s, int M, int N, int K, const float* X, int ldx, const float* W, const float* b, const float* b2, float* Y,
                int ldy, int act) {
  if (g_skinny < 0) {
    const char* e = getenv("GLAMR_NET_SKINNY");
    g_skinny = e ? atoi(e) : 1;
  }
  if (g_skinny && M <= kSkinnyMaxM) return gemm_skinny_launch(s, M, N, K, X, ldx, W, b, b2, Y, ldy, act);
  if (g_gemm_mode == 1) {
    static int ntile = -1;     // GLAMR_TC_NTILE = 32 | 128 forces one tile shape (experiments); default: by problem height
    if (ntile < 0) {
#ifdef GLAMR_EXPERIMENT
      if (const char* e = getenv("GLAMR_TC_DEBUG")) {
        const int v = atoi(e);
        GLAMR_CUDA_TRY(cudaMemcpyToSymbol(g_tc_dbg, &v, sizeof(int)));
      }
#endif
      const char* t = getenv("GLAMR_TC_NTILE");
      ntile = t ? atoi(t) : 0;
    }
    const bool narrow = ntile == 32 || (ntile != 128 && M <= TCM);
    return narrow ? gemm_tc_launch<32>(s, M, N, K, X, ldx, W, b, b2, Y, ldy, act) : gemm_tc_launch<128>(s, M, N, K, X, ldx, W, b, b2, Y, ldy, act);
  }
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT);
  if (act == 1)
    gemm_bias_act_kernel<1><<<grid, 256, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
  else
    gemm_bias_act_kernel<0><<<grid, 256, 0, s>>>(M, N, K, X, ldx, W, b, b2, Y, ldy);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}

// ------------------------------------------------------------------------------------------------ LayerNorm(x + r), D = 256
__global__ void __launch_bounds__(128) add_layernorm_kernel(int M, const float* __restrict__ X, const float* __restrict__ R,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ Y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  float v[8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    v[i] = X[(size_t)row * 256 + c] + R[(size_t)row * 256 + c];
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    Y[(size_t)row * 256 + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

// ------------------------------------------------------------------------------------------------ attention, head dim 32
// Q rows (tq * B + b), K/V rows (tk * B + b); row strides ldq / ldkv; 8 heads of 32.  key_mask [B,Sk] (1 = ignore) or NULL.
// grid (B, 8 heads, ceil(Sq / 8)): a CTA stages the head's K / V once and its 4 warps take 2 queries each, so that the kernel is
// one short dependent chain deep at the batch sizes of the inference (B = 1-4) instead of Sq / 4 of them.
constexpr int kAttnQChunk = 8;
__global__ void __launch_bounds__(128) attention_kernel(int B, int Sq, int Sk, const float* __restrict__ Q, int ldq,
                                                        const float* __restrict__ K, const float* __restrict__ V, int ldkv,
                                                        const uint8_t* __restrict__ key_mask, float* __restrict__ O, int ldo) {
  __shared__ float Ks[64][33];
  __shared__ float Vs[64][33];
  __shared__ float Ps[4][64];
  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int e = tid; e < Sk * 32; e += 128) {
    const int j = e >> 5, d = e & 31;
    Ks[j][d] = K[((size_t)j * B + b) * ldkv + h * 32 + d];
    Vs[j][d] = V[((size_t)j * B + b) * ldkv + h * 32 + d];
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;   // 1/sqrt(32)
  const int q_end = min(Sq, ((int)blockIdx.z + 1) * kAttnQChunk);
  for (int q = blockIdx.z * kAttnQChunk + w; q < q_end; q += 4) {
    const float qd = Q[((size_t)q * B + b) * ldq + h * 32 + lane] * scale;   // torch scales q before q k^T
    float sc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = lane + 32 * r;
      float a = 0.0f;
#pragma unroll
      for (int d = 0; d < 32; ++d) a = fmaf(__shfl_sync(0xffffffffu, qd, d), (j < Sk) ? Ks[j][d] : 0.0f, a);
      const bool dead = (j >= Sk) || (key_mask && key_mask[(size_t)b * Sk + j]);
      sc[r] = dead ? -INFINITY : a;
    }
    float mx = fmaxf(sc[0], sc[1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = (sc[0] == -INFINITY) ? 0.0f : expf(sc[0] - mx), e1 = (sc[1] == -INFINITY) ? 0.0f : expf(sc[1] - mx);
    const float den = warp_sum(e0 + e1);
    Ps[w][lane] = e0 / den;
    Ps[w][lane + 32] = e1 / den;
    __syncwarp();
    float o = 0.0f;
    for (int j = 0; j < Sk; ++j) o = fmaf(Ps[w][j], Vs[j][lane], o);      // in key order (the reference's softmax(QK^T) V row sum)
    O[((size_t)q * B + b) * ldo + h * 32 + lane] = o;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ positional encoding
// out[row] = [ src_row (in_dim) | PE(pos) (256) ] with the 'original' sinusoid (lib/models/pos_encoding.py:27-32);
// row = t * B + b, pos = t + pos_offset; src row = (src_bcast_t ? b : row) -> lets z be repeated over time.
__global__ void pe_concat_kernel(int rows, int B, int in_dim, const float* __restrict__ src, int src_ld, int src_bcast_t,
                                 int token_mode, int pos_offset, float* __restrict__ out) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int t = row / B, b = row - t * B;
  const int od = in_dim + 256;
  const float* s = token_mode ? (src + (size_t)t * src_ld) : (src + (size_t)(src_bcast_t ? b : row) * src_ld);
  for (int c = threadIdx.x; c < od; c += blockDim.x) {
    float v;
    if (c < in_dim) {
      v = s[c];
    } else {
      const int e = c - in_dim;
      const float mul = expf((float)(e & ~1) * (-9.210340371976184f / 256.0f));   // exp(2i * -ln(1e4)/256)
      const float a = (float)(t + pos_offset) * mul;
      v = (e & 1) ? cosf(a) : sinf(a);
    }
    out[(size_t)row * od + c] = v;
  }
}

// rows t*B+b of out [T2,B,D] = cat(a[:Ta], b_[:Tb]) along time
__global__ void concat_time_kernel(int Ta, int Tb, int B, int D, const float* __restrict__ a, const float* __restrict__ b_,
                                   float* __restrict__ out) {
  const size_t total = (size_t)(Ta + Tb) * B * D;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / D;
    const int t = (int)(row / B);
    out[e] = (t < Ta) ? a[e] : b_[e - (size_t)Ta * B * D];
  }
}

// z = mu + eps * exp(0.5 logvar)  (lib/utils/dist.py:8-26); eps may be NULL (-> mu) or broadcast over the batch
__global__ void sample_z_kernel(int B, int nz, const float* __restrict__ mu, int ld_mu, const float* __restrict__ logvar, int ld_lv,
                                const float* __restrict__ eps, int eps_ld, float* __restrict__ z) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * nz) return;
  const int b = e / nz, k = e - b * nz;
  const float ep = eps ? eps[(size_t)b * eps_ld + k] : 0.0f;
  z[e] = mu[(size_t)b * ld_mu + k] + ep * expf(0.5f * logvar[(size_t)b * ld_lv + k]);
}

__global__ void mean_time_kernel(int T, int B, int D, const float* __restrict__ X, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * D) return;
  float s = 0.0f;
  for (int t = 0; t < T; ++t) s += X[(size_t)t * B * D + e];
  out[e] = s / (float)T;
}

// [z (nz, per batch) | context row] -> rows of width nz + D
__global__ void concat_z_kernel(int rows, int B, int nz, int D, const float* __restrict__ z, const float* __restrict__ ctx,
                                float* __restrict__ out) {
  const size_t total = (size_t)rows * (nz + D);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / (nz + D);
    const int c = (int)(e - row * (nz + D));
    const int b = (int)(row % B);
    out[e] = (c < nz) ? z[(size_t)b * nz + c] : ctx[row * D + (c - nz)];
  }
}

// frame 0 of the predicted local trajectory: xy := init_xy (or 0), heading vec := init (or (0,1))
// (traj_pred_vae.py:319-329)
__global__ void traj_first_frame_kernel(int B, float* __restrict__ local /*[T,B,11]*/, const float* __restrict__ init_xy,
                                        const float* __restrict__ init_heading) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* l = local + (size_t)b * 11;
  l[0] = init_xy ? init_xy[b * 2] : 0.0f;
  l[1] = init_xy ? init_xy[b * 2 + 1] : 0.0f;
  l[9] = init_heading ? cosf(init_heading[b]) : 0.0f;
  l[10] = init_heading ? sinf(init_heading[b]) : 1.0f;
}

__global__ void quat_rows_to_aa_kernel(int n, const float* __restrict__ q, float* __restrict__ aa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float qq[4] = {q[i * 4], q[i * 4 + 1], q[i * 4 + 2], q[i * 4 + 3]}, a[3];
  quat_to_aa(qq, a);
  aa[i * 3] = a[0]; aa[i * 3 + 1] = a[1]; aa[i * 3 + 2] = a[2];
}

// ------------------------------------------------------------------------------------------------ LSTM recurrence
// nn.LSTMCell loop of lib/models/rnn.py:45-61 for one direction of one layer, hidden 128, gate order i,f,g,o.
// grid (B, 2 directions), 512 threads = one gate row each.  The recurrent matrix W_hh [512,128] lives on chip for the
// whole sequence: columns 0..63 of a thread's row in registers, columns 64..127 in shared memory ([k][row], conflict
// free).  xproj [T,B,2,512] already holds W_ih x_t + b_ih + b_hh.  out [T,B,256] = [h_fwd | h_bwd].
constexpr int LH = 128, LG = 4 * LH, LREG = 64;
__global__ void __launch_bounds__(LG, 1) lstm_recurrence_kernel(int T, int B, const float* __restrict__ xproj,
                                                                const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                                                                float* __restrict__ out) {
  extern __shared__ float smem[];
  float* Wsm = smem;                       // [LH - LREG][LG]
  float* hs = Wsm + (LH - LREG) * LG;      // [LH]
  float* gs = hs + LH;                     // [LG]
  const int b = blockIdx.x, dir = blockIdx.y, r = threadIdx.x;
  const float* W = (dir == 0 ? whh_f : whh_b) + (size_t)r * LH;
  float wreg[LREG];
#pragma unroll
  for (int k = 0; k < LREG; ++k) wreg[k] = W[k];
  for (int k = 0; k < LH - LREG; ++k) Wsm[k * LG + r] = W[LREG + k];
  float c = 0.0f;
  if (r < LH) hs[r] = 0.0f;
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    float a = xproj[(((size_t)t * B + b) * 2 + dir) * LG + r];
#pragma unroll
    for (int k4 = 0; k4 < LREG / 4; ++k4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hs + 4 * k4);
      a = fmaf(wreg[4 * k4], h4.x, a);
      a = fmaf(wreg[4 * k4 + 1], h4.y, a);
      a = fmaf(wreg[4 * k4 + 2], h4.z, a);
      a = fmaf(wreg[4 * k4 + 3], h4.w, a);
    }
#pragma unroll 8
    for (int k = 0; k < LH - LREG; ++k) a = fmaf(Wsm[k * LG + r], hs[LREG + k], a);
    gs[r] = a;
    __syncthreads();
    if (r < LH) {
      const float ig = 1.0f / (1.0f + expf(-gs[r]));
      const float fg = 1.0f / (1.0f + expf(-gs[LH + r]));
      const float gg = tanhf(gs[2 * LH + r]);
      const float og = 1.0f / (1.0f + expf(-gs[3 * LH + r]));
      c = fg * c + ig * gg;
      const float h = og * tanhf(c);
      hs[r] = h;
      out[((size_t)t * B + b) * (2 * LH) + dir * LH + r] = h;
    }
    __syncthreads();
  }
}

}  // namespace glamr

// =================================================================================================== C ABI
using namespace glamr;

struct glamr_net {
  std::map<std::string, std::pair<float*, size_t>> t;
  std::vector<void*> allocs;
};

namespace {
const float* W(const glamr_net* n, const std::string& name, size_t expect, int* err) {
  auto it = n->t.find(name);
  if (it == n->t.end() || (expect && it->second.second != expect)) { *err = 1; return nullptr; }
  return it->second.first;
}

struct AttnW { const float *in_w, *in_b, *out_w, *out_b; };
struct EncLayer { AttnW sa; const float *l1w, *l1b, *l2w, *l2b, *n1g, *n1b, *n2g, *n2b; };
struct DecLayer { AttnW sa, ca; const float *l1w, *l1b, *l2w, *l2b, *n1g, *n1b, *n2g, *n2b, *n3g, *n3b; };

AttnW attn_w(const glamr_net* n, const std::string& p, int* e) {
  return {W(n, p + ".in_proj_weight", 768 * 256, e), W(n, p + ".in_proj_bias", 768, e), W(n, p + ".out_proj.weight", 256 * 256, e),
          W(n, p + ".out_proj.bias", 256, e)};
}
EncLayer enc_layer(const glamr_net* n, const std::string& p, int* e) {
  EncLayer L;
  L.sa = attn_w(n, p + ".self_attn", e);
  L.l1w = W(n, p + ".linear1.weight", 512 * 256, e); L.l1b = W(n, p + ".linear1.bias", 512, e);
  L.l2w = W(n, p + ".linear2.weight", 256 * 512, e); L.l2b = W(n, p + ".linear2.bias", 256, e);
  L.n1g = W(n, p + ".norm1.weight", 256, e); L.n1b = W(n, p + ".norm1.bias", 256, e);
  L.n2g = W(n, p + ".norm2.weight", 256, e); L.n2b = W(n, p + ".norm2.bias", 256, e);
  return L;
}
DecLayer dec_layer(const glamr_net* n, const std::string& p, int* e) {
  DecLayer L;
  L.sa = attn_w(n, p + ".self_attn", e);
  L.ca = attn_w(n, p + ".multihead_attn", e);
  L.l1w = W(n, p + ".linear1.weight", 512 * 256, e); L.l1b = W(n, p + ".linear1.bias", 512, e);
  L.l2w = W(n, p + ".linear2.weight", 256 * 512, e); L.l2b = W(n, p + ".linear2.bias", 256, e);
  L.n1g = W(n, p + ".norm1.weight", 256, e); L.n1b = W(n, p + ".norm1.bias", 256, e);
  L.n2g = W(n, p + ".norm2.weight", 256, e); L.n2b = W(n, p + ".norm2.bias", 256, e);
  L.n3g = W(n, p + ".norm3.weight", 256, e); L.n3b = W(n, p + ".norm3.bias", 256, e);
  return L;
}

struct Arena {
  float* base; size_t cap, used;
  float* take(size_t n) { size_t o = used; used += (n + 63) & ~(size_t)63; return used <= cap ? base + o : nullptr; }
};

int layernorm(cudaStream_t s, int M, const float* X, const float* R, const float* g, const float* b, float* Y) {
  add_layernorm_kernel<<<(M + 3) / 4, 128, 0, s>>>(M, X, R, g, b, Y);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}

// multi-head attention block: out = out_proj(attn(q_src, kv_src)); q_src [Sq*B,256], kv_src [Sk*B,256]
int mha(cudaStream_t s, Arena& A, const AttnW& w, int B, int Sq, int Sk, const float* q_src, const float* kv_src, const uint8_t* mask,
        float* out) {
  if (Sk > 64) return GLAMR_EUNSUPPORTED;
  const int Mq = Sq * B, Mk = Sk * B;
  const size_t mark = A.used;
  float* q = A.take((size_t)Mq * 256);
  float* kv = A.take((size_t)Mk * 512);
  float* att = A.take((size_t)Mq * 256);
  if (!q || !kv || !att) return GLAMR_ENOSPACE;
  int rc;
  if ((rc = gemm(s, Mq, 256, 256, q_src, 256, w.in_w, w.in_b, nullptr, q, 256, 0))) return rc;
  if ((rc = gemm(s, Mk, 512, 256, kv_src, 256, w.in_w + 256 * 256, w.in_b + 256, nullptr, kv, 512, 0))) return rc;
  attention_kernel<<<dim3(B, 8, (Sq + kAttnQChunk - 1) / kAttnQChunk), 128, 0, s>>>(B, Sq, Sk, q, 256, kv, kv + 256, 512, mask, att, 256);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, Mq, 256, 256, att, 256, w.out_w, w.out_b, nullptr, out, 256, 0))) return rc;
  A.used = mark;
  return GLAMR_OK;
}

int ffn(cudaStream_t s, Arena& A, int M, const float* x, const float* l1w, const float* l1b, const float* l2w, const float* l2b, float* out) {
  const size_t mark = A.used;
  float* h = A.take((size_t)M * 512);
  if (!h) return GLAMR_ENOSPACE;
  int rc;
  if ((rc = gemm(s, M, 512, 256, x, 256, l1w, l1b, nullptr, h, 512, 1))) return rc;
  if ((rc = gemm(s, M, 256, 512, h, 512, l2w, l2b, nullptr, out, 256, 0))) return rc;
  A.used = mark;
  return GLAMR_OK;
}

// nn.TransformerEncoderLayer, post-norm, relu, eval mode (dropout = identity); x updated in place
int encoder_layer(cudaStream_t s, Arena& A, const EncLayer& L, int B, int S, float* x, const uint8_t* mask) {
  const int M = S * B;
  const size_t mark = A.used;
  float* t = A.take((size_t)M * 256);
  if (!t) return GLAMR_ENOSPACE;
  int rc;
  if ((rc = mha(s, A, L.sa, B, S, S, x, x, mask, t))) return rc;
  if ((rc = layernorm(s, M, x, t, L.n1g, L.n1b, x))) return rc;
  if ((rc = ffn(s, A, M, x, L.l1w, L.l1b, L.l2w, L.l2b, t))) return rc;
  if ((rc = layernorm(s, M, x, t, L.n2g, L.n2b, x))) return rc;
  A.used = mark;
  return GLAMR_OK;
}
// nn.TransformerDecoderLayer (no tgt mask), memory [Sm*B,256] with key padding mask
int decoder_layer(cudaStream_t s, Arena& A, const DecLayer& L, int B, int S, int Sm, float* x, const float* mem, const uint8_t* mem_mask) {
  const int M = S * B;
  const size_t mark = A.used;
  float* t = A.take((size_t)M * 256);
  if (!t) return GLAMR_ENOSPACE;
  int rc;
  if ((rc = mha(s, A, L.sa, B, S, S, x, x, nullptr, t))) return rc;
  if ((rc = layernorm(s, M, x, t, L.n1g, L.n1b, x))) return rc;
  if ((rc = mha(s, A, L.ca, B, S, Sm, x, mem, mem_mask, t))) return rc;
  if ((rc = layernorm(s, M, x, t, L.n2g, L.n2b, x))) return rc;
  if ((rc = ffn(s, A, M, x, L.l1w, L.l1b, L.l2w, L.l2b, t))) return rc;
  if ((rc = layernorm(s, M, x, t, L.n3g, L.n3b, x))) return rc;
  A.used = mark;
  return GLAMR_OK;
}
}  // namespace

// Y[M,N] = act(X[M,K] W[N,K]^T + bias) -- stand-alone entry for the GEMM used by every Linear of the prior networks
// (nn.Linear in lib/models/mlp.py:32-41, nn.MultiheadAttention projections, FFN).  mode: 1 tcgen05 3xTF32, 0 FP32 SIMT.
extern "C" int glamr_linear_forward(int M, int N, int K, const float* X, const float* W, const float* bias, int relu, float* Y, int mode,
                                    void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !X || !W || !Y) return GLAMR_EINVAL;
  const int saved = g_gemm_mode;
  g_gemm_mode = mode;
  const int rc = gemm((cudaStream_t)stream, M, N, K, X, K, W, bias, nullptr, Y, N, relu ? 1 : 0);
  g_gemm_mode = saved;
  return rc;
}
extern "C" int glamr_net_set_gemm_mode(int mode) {
  if (mode != 0 && mode != 1) return GLAMR_EINVAL;
  g_gemm_mode = mode;
  return GLAMR_OK;
}

extern "C" int glamr_net_create(glamr_net** out) {
  if (!out) return GLAMR_EINVAL;
  *out = new glamr_net();
  return GLAMR_OK;
}
extern "C" int glamr_net_destroy(glamr_net* n) {
  if (!n) return GLAMR_OK;
  cudaDeviceSynchronize();
  wimg_clear();                              // operand images are keyed by weight pointers that are about to be freed
  for (void* p : n->allocs) cudaFree(p);
  delete n;
  return GLAMR_OK;
}
// upload one named parameter (HOST pointer, float32)
extern "C" int glamr_net_set_tensor(glamr_net* n, const char* name, const float* host, size_t numel) {
  if (!n || !name || !host || numel == 0) return GLAMR_EINVAL;
  void* p = nullptr;
  GLAMR_CUDA_TRY(cudaMalloc(&p, numel * sizeof(float)));
  GLAMR_CUDA_TRY(cudaMemcpy(p, host, numel * sizeof(float), cudaMemcpyHostToDevice));
  n->allocs.push_back(p);
  if (n->t.count(name)) { cudaDeviceSynchronize(); wimg_clear(); }      // a weight was replaced: cached operand images are stale
  n->t[name] = {(float*)p, numel};
  return GLAMR_OK;
}

extern "C" size_t glamr_infiller_workspace_floats(int B) { return (size_t)B * 50 * 256 * 16 + 65536; }

// One 50-frame window of MotionInfillerVAE.inference_one_step (mode 'infer', sample_num 1), B sequences.
//   in_pose [50,B,69]  key_pad_mask [B,50] (1 = frame not usable as key)  eps [B or 1,128] (eps_rows = B or 1) or NULL
//   out_pose [40,B,69] = [first 10 input frames | 30 decoded frames]
extern "C" int glamr_infiller_window_forward(const glamr_net* n, int B, const float* in_pose, const uint8_t* key_pad_mask,
                                             const float* eps, int eps_rows, float* out_pose, float* workspace,
                                             size_t workspace_floats, void* stream) {
  if (!n || B <= 0 || !in_pose || !key_pad_mask || !out_pose || !workspace) return GLAMR_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  int e = 0;
  const std::string ce = "context_encoder.", dd = "data_decoder.";
  const float* in_fc_w = W(n, ce + "in_fc.weight", 256 * 69, &e), * in_fc_b = W(n, ce + "in_fc.bias", 256, &e);
  const float* cpe_w = W(n, ce + "pos_enc.fc.weight", 256 * 512, &e), * cpe_b = W(n, ce + "pos_enc.fc.bias", 256, &e);
  EncLayer enc[2] = {enc_layer(n, ce + "temporal_net.layers.0", &e), enc_layer(n, ce + "temporal_net.layers.1", &e)};
  const float* dpe_w = W(n, dd + "pos_enc.fc.weight", 256 * 384, &e), * dpe_b = W(n, dd + "pos_enc.fc.bias", 256, &e);
  DecLayer dec[2] = {dec_layer(n, dd + "temporal_net.layers.0", &e), dec_layer(n, dd + "temporal_net.layers.1", &e)};
  const float* om0w = W(n, dd + "out_mlp.affine_layers.0.weight", 512 * 256, &e), * om0b = W(n, dd + "out_mlp.affine_layers.0.bias", 512, &e);
  const float* om1w = W(n, dd + "out_mlp.affine_layers.1.weight", 256 * 512, &e), * om1b = W(n, dd + "out_mlp.affine_layers.1.bias", 256, &e);
  const float* ofw = W(n, dd + "out_fc.weight", 69 * 256, &e), * ofb = W(n, dd + "out_fc.bias", 69, &e);
  const float* ppe_w = W(n, dd + "prior_pos_enc.fc.weight", 256 * 512, &e), * ppe_b = W(n, dd + "prior_pos_enc.fc.bias", 256, &e);
  DecLayer pri = dec_layer(n, dd + "prior_temporal_net.layers.0", &e);
  const float* mu_tok = W(n, dd + "mu_token", 256, &e), * lv_tok = W(n, dd + "logvar_token", 256, &e);
  const float* pmw = W(n, dd + "p_z_mu_net.weight", 128 * 256, &e), * pmb = W(n, dd + "p_z_mu_net.bias", 128, &e);
  const float* plw = W(n, dd + "p_z_logvar_net.weight", 128 * 256, &e), * plb = W(n, dd + "p_z_logvar_net.bias", 128, &e);
  if (e) return GLAMR_EINVAL;
  Arena A{workspace, workspace_floats, 0};
  const int S = 50, Sc = 30, M = S * B, Mc = Sc * B;
  float* x = A.take((size_t)M * 256);
  float* cat = A.take((size_t)M * 512);
  float* tok = A.take(2 * 256);
  float* px = A.take((size_t)2 * B * 256);
  float* mu = A.take((size_t)B * 128);
  float* lv = A.take((size_t)B * 128);
  float* z = A.take((size_t)B * 128);
  float* dx = A.take((size_t)Mc * 256);
  float* h1 = A.take((size_t)Mc * 512);
  float* h2 = A.take((size_t)Mc * 256);
  float* dec_out = A.take((size_t)Mc * 69);
  if (!dec_out) return GLAMR_ENOSPACE;
  int rc;
  // ---- context encoder (motion_infiller_vae.py:92-123)
  if ((rc = gemm(s, M, 256, 69, in_pose, 69, in_fc_w, in_fc_b, nullptr, x, 256, 0))) return rc;
  pe_concat_kernel<<<M, 128, 0, s>>>(M, B, 256, x, 256, 0, 0, 0, cat);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, M, 256, 512, cat, 512, cpe_w, cpe_b, nullptr, x, 256, 0))) return rc;
  for (int l = 0; l < 2; ++l)
    if ((rc = encoder_layer(s, A, enc[l], B, S, x, key_pad_mask))) return rc;
  // ---- learned prior over z (:354-362): two tokens attend to the context
  GLAMR_CUDA_TRY(cudaMemcpyAsync(tok, mu_tok, 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  GLAMR_CUDA_TRY(cudaMemcpyAsync(tok + 256, lv_tok, 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  pe_concat_kernel<<<2 * B, 128, 0, s>>>(2 * B, B, 256, tok, 256, 0, 1, 0, cat);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, 2 * B, 256, 512, cat, 512, ppe_w, ppe_b, nullptr, px, 256, 0))) return rc;
  if ((rc = decoder_layer(s, A, pri, B, 2, S, px, x, key_pad_mask))) return rc;
  if ((rc = gemm(s, B, 128, 256, px, 256, pmw, pmb, nullptr, mu, 128, 0))) return rc;
  if ((rc = gemm(s, B, 128, 256, px + (size_t)B * 256, 256, plw, plb, nullptr, lv, 128, 0))) return rc;
  sample_z_kernel<<<(B * 128 + 127) / 128, 128, 0, s>>>(B, 128, mu, 128, lv, 128, eps, eps_rows == 1 ? 0 : 128, z);
  GLAMR_LAUNCH_CHECK();
  // ---- decoder (:383-395): z repeated over the 30 current frames, PE offset 10
  pe_concat_kernel<<<Mc, 128, 0, s>>>(Mc, B, 128, z, 128, 1, 0, 10, cat);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, Mc, 256, 384, cat, 384, dpe_w, dpe_b, nullptr, dx, 256, 0))) return rc;
  for (int l = 0; l < 2; ++l)
    if ((rc = decoder_layer(s, A, dec[l], B, Sc, S, dx, x, key_pad_mask))) return rc;
  if ((rc = gemm(s, Mc, 512, 256, dx, 256, om0w, om0b, nullptr, h1, 512, 1))) return rc;
  if ((rc = gemm(s, Mc, 256, 512, h1, 512, om1w, om1b, nullptr, h2, 256, 1))) return rc;
  if ((rc = gemm(s, Mc, 69, 256, h2, 256, ofw, ofb, nullptr, dec_out, 69, 0))) return rc;
  concat_time_kernel<<<64, 256, 0, s>>>(10, Sc, B, 69, in_pose, dec_out, out_pose);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}


// ------------------------------------------------------------------------------------------------ all windows of a sequence
// motion_infiller_vae.py:618-632: the autoregressive sweep of 50-frame windows with stride 30.  Window i reads frames
// [30 i, 30 i + 50) of the running pose (zeros past the end), masks keys that are invisible or past the end (the 10 past frames
// are always usable), and its first min(40, T - 30 i) output frames replace the running pose.  One call for the whole sweep: the
// window staging / commit are two small kernels instead of a dozen framework ops per window.
__global__ void window_prep_kernel(int T, int B, int s0, const float* __restrict__ pose, const uint8_t* __restrict__ key_pad_all,
                                   float* __restrict__ win, uint8_t* __restrict__ kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = B * 69;
  if (i < 50 * row) {
    const int t = i / row;
    win[i] = (s0 + t < T) ? pose[(size_t)(s0 + t) * row + (i - t * row)] : 0.0f;
  }
  if (i < B * 50) {
    const int b = i / 50, t = i - b * 50;
    kp[i] = t < 10 ? 0 : ((s0 + t < T) ? key_pad_all[(size_t)b * T + s0 + t] : 1);
  }
}
__global__ void window_commit_kernel(int B, int s0, int nfr, const float* __restrict__ out, float* __restrict__ pose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = B * 69;
  if (i < nfr * row) pose[(size_t)s0 * row + i] = out[i];
}

extern "C" size_t glamr_infiller_sequence_workspace_floats(int B) {
  return glamr_infiller_workspace_floats(B) + (size_t)90 * B * 69 + (size_t)(B * 50 + 3) / 4 + 192;
}

//   pose_io [T,B,69]: the input body pose, overwritten with the infilled pose   key_pad_all [B,T] uint8 (1 = frame invisible)
//   eps [n_windows][eps_rows][128] with eps_rows in {1, B}, n_windows = ceil((T - 10) / 30)
extern "C" int glamr_infiller_forward(const glamr_net* n, int T, int B, float* pose_io, const uint8_t* key_pad_all, const float* eps,
                                      int eps_rows, float* workspace, size_t workspace_floats, void* stream) {
  if (!n || T <= 10 || B <= 0 || !pose_io || !key_pad_all || !eps || !workspace || (eps_rows != 1 && eps_rows != B)) return GLAMR_EINVAL;
  if (workspace_floats < glamr_infiller_sequence_workspace_floats(B)) return GLAMR_ENOSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* win = workspace;
  float* out = win + (size_t)50 * B * 69;
  uint8_t* kp = reinterpret_cast<uint8_t*>(out + (size_t)40 * B * 69);
  float* rest = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(kp + (size_t)B * 50) + 255) & ~(uintptr_t)255);
  const size_t rest_floats = workspace_floats - (size_t)(rest - workspace);
  const int nwin = (T - 10 + 29) / 30;
  const int prep_blocks = (50 * B * 69 + 255) / 256;
  for (int i = 0; i < nwin; ++i) {
    const int s0 = i * 30;
    window_prep_kernel<<<prep_blocks, 256, 0, s>>>(T, B, s0, pose_io, key_pad_all, win, kp);
    GLAMR_LAUNCH_CHECK();
    const int rc = glamr_infiller_window_forward(n, B, win, kp, eps + (size_t)i * eps_rows * 128, eps_rows, out, rest, rest_floats, stream);
    if (rc) return rc;
    const int nfr = (s0 + 40 < T ? s0 + 40 : T) - s0;
    window_commit_kernel<<<(nfr * B * 69 + 255) / 256, 256, 0, s>>>(B, s0, nfr, out, pose_io);
    GLAMR_LAUNCH_CHECK();
  }
  return GLAMR_OK;
}

extern "C" size_t glamr_trajpred_workspace_floats(int T, int B) { return (size_t)T * B * (512 + 256 + 2 * 512 + 256 + 384 + 512 + 256 + 64) + (size_t)B * 2048 + 65536; }

// TrajPredVAE.inference (multi_step False, sample_num 1): joint positions -> local trajectory -> global trajectory.
//   in_joint_pos [T,B,69]   eps [B or 1,128] or NULL   init_xy [B,2] / init_heading [B] or NULL
//   out_local_traj [T,B,11]  out_trans [T,B,3]  out_orient_aa [T,B,3]
extern "C" int glamr_trajpred_forward(const glamr_net* n, int T, int B, const float* in_joint_pos, const float* eps, int eps_rows,
                                      const float* init_xy, const float* init_heading, float* out_local_traj, float* out_trans,
                                      float* out_orient_aa, float* workspace, size_t workspace_floats, void* stream);

extern "C" int glamr_traj_local2global(int T, int B, const float* local_traj, int local_heading, float* trans, float* orient_q,
                                       float* scratch, void* stream);

extern "C" int glamr_trajpred_forward(const glamr_net* n, int T, int B, const float* in_joint_pos, const float* eps, int eps_rows,
                                      const float* init_xy, const float* init_heading, float* out_local_traj, float* out_trans,
                                      float* out_orient_aa, float* workspace, size_t workspace_floats, void* stream) {
  if (!n || T <= 0 || B <= 0 || !in_joint_pos || !out_local_traj || !out_trans || !out_orient_aa || !workspace) return GLAMR_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  ScopedFp32Gemm fp32_only;
  int e = 0;
  const std::string ce = "context_encoder.", dd = "data_decoder.";
  const float* im0w = W(n, ce + "in_mlp.affine_layers.0.weight", 512 * 69, &e), * im0b = W(n, ce + "in_mlp.affine_layers.0.bias", 512, &e);
  const float* im1w = W(n, ce + "in_mlp.affine_layers.1.weight", 256 * 512, &e), * im1b = W(n, ce + "in_mlp.affine_layers.1.bias", 256, &e);
  const float *wih[2][2], *whh[2][2], *bih[2][2], *bhh[2][2];
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      const std::string p = ce + "temporal_net." + std::to_string(l) + (d == 0 ? ".rnn_f." : ".rnn_b.");
      wih[l][d] = W(n, p + "weight_ih", 512 * 256, &e); whh[l][d] = W(n, p + "weight_hh", 512 * 128, &e);
      bih[l][d] = W(n, p + "bias_ih", 512, &e); bhh[l][d] = W(n, p + "bias_hh", 512, &e);
    }
  const float* cm0w = W(n, ce + "out_mlp.affine_layers.0.weight", 512 * 256, &e), * cm0b = W(n, ce + "out_mlp.affine_layers.0.bias", 512, &e);
  const float* cm1w = W(n, ce + "out_mlp.affine_layers.1.weight", 256 * 512, &e), * cm1b = W(n, ce + "out_mlp.affine_layers.1.bias", 256, &e);
  const float* pm0w = W(n, dd + "prior_mlp.affine_layers.0.weight", 512 * 256, &e), * pm0b = W(n, dd + "prior_mlp.affine_layers.0.bias", 512, &e);
  const float* pm1w = W(n, dd + "prior_mlp.affine_layers.1.weight", 256 * 512, &e), * pm1b = W(n, dd + "prior_mlp.affine_layers.1.bias", 256, &e);
  const float* pzw = W(n, dd + "p_z_net.weight", 256 * 256, &e), * pzb = W(n, dd + "p_z_net.bias", 256, &e);
  const float* dm0w = W(n, dd + "out_mlp.affine_layers.0.weight", 512 * 384, &e), * dm0b = W(n, dd + "out_mlp.affine_layers.0.bias", 512, &e);
  const float* dm1w = W(n, dd + "out_mlp.affine_layers.1.weight", 256 * 512, &e), * dm1b = W(n, dd + "out_mlp.affine_layers.1.bias", 256, &e);
  const float* ofw = W(n, dd + "out_fc.weight", 11 * 256, &e), * ofb = W(n, dd + "out_fc.bias", 11, &e);
  if (e) return GLAMR_EINVAL;
  Arena A{workspace, workspace_floats, 0};
  const int M = T * B;
  float* h512 = A.take((size_t)M * 512);
  float* x = A.take((size_t)M * 256);
  float* xp = A.take((size_t)M * 2 * 512);
  float* y = A.take((size_t)M * 256);
  float* cat = A.take((size_t)M * 384);
  float* hm = A.take((size_t)B * 256);
  float* hp = A.take((size_t)B * 512);
  float* hq = A.take((size_t)B * 256);
  float* pz = A.take((size_t)B * 256);
  float* z = A.take((size_t)B * 128);
  float* oq = A.take((size_t)M * 4);
  float* sc = A.take((size_t)M * 3);
  if (!sc) return GLAMR_ENOSPACE;
  int rc;
  static bool attr = false;
  const size_t lstm_smem = ((size_t)(LH - LREG) * LG + LH + LG) * sizeof(float);
  if (!attr) {
    GLAMR_CUDA_TRY(cudaFuncSetAttribute(lstm_recurrence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem));
    attr = true;
  }
  // ---- context encoder (traj_pred_vae.py:72-92)
  if ((rc = gemm(s, M, 512, 69, in_joint_pos, 69, im0w, im0b, nullptr, h512, 512, 1))) return rc;
  if ((rc = gemm(s, M, 256, 512, h512, 512, im1w, im1b, nullptr, x, 256, 1))) return rc;
  for (int l = 0; l < 2; ++l) {
    for (int d = 0; d < 2; ++d)   // xproj[t][b][d][:] = W_ih x + b_ih + b_hh
      if ((rc = gemm(s, M, 512, 256, x, 256, wih[l][d], bih[l][d], bhh[l][d], xp + d * 512, 1024, 0))) return rc;
    lstm_recurrence_kernel<<<dim3(B, 2), LG, lstm_smem, s>>>(T, B, xp, whh[l][0], whh[l][1], y);
    GLAMR_LAUNCH_CHECK();
    float* tmp = x; x = y; y = tmp;
  }
  if ((rc = gemm(s, M, 512, 256, x, 256, cm0w, cm0b, nullptr, h512, 512, 1))) return rc;
  if ((rc = gemm(s, M, 256, 512, h512, 512, cm1w, cm1b, nullptr, y, 256, 1))) return rc;      // y = context
  // ---- prior + z (:281-297)
  mean_time_kernel<<<(B * 256 + 127) / 128, 128, 0, s>>>(T, B, 256, y, hm);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, B, 512, 256, hm, 256, pm0w, pm0b, nullptr, hp, 512, 1))) return rc;
  if ((rc = gemm(s, B, 256, 512, hp, 512, pm1w, pm1b, nullptr, hq, 256, 1))) return rc;
  if ((rc = gemm(s, B, 256, 256, hq, 256, pzw, pzb, nullptr, pz, 256, 0))) return rc;
  sample_z_kernel<<<(B * 128 + 127) / 128, 128, 0, s>>>(B, 128, pz, 256, pz + 128, 256, eps, eps_rows == 1 ? 0 : 128, z);
  GLAMR_LAUNCH_CHECK();
  // ---- decoder (:298-333)
  concat_z_kernel<<<256, 256, 0, s>>>(M, B, 128, 256, z, y, cat);
  GLAMR_LAUNCH_CHECK();
  if ((rc = gemm(s, M, 512, 384, cat, 384, dm0w, dm0b, nullptr, h512, 512, 1))) return rc;
  if ((rc = gemm(s, M, 256, 512, h512, 512, dm1w, dm1b, nullptr, x, 256, 1))) return rc;
  if ((rc = gemm(s, M, 11, 256, x, 256, ofw, ofb, nullptr, out_local_traj, 11, 0))) return rc;
  traj_first_frame_kernel<<<(B + 127) / 128, 128, 0, s>>>(B, out_local_traj, init_xy, init_heading);
  GLAMR_LAUNCH_CHECK();
  if ((rc = glamr_traj_local2global(T, B, out_local_traj, 1, out_trans, oq, sc, stream))) return rc;
  quat_rows_to_aa_kernel<<<(M + 127) / 128, 128, 0, s>>>(M, oq, out_orient_aa);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}
