// One global-optimisation iteration of GLAMR on sm_100a (global_recon/models/global_recon_model.py:547-570):
//   traj_forward    (1 CTA / person: trajectory codec with two block-wide prefix scans)
//   cam_forward     (1 thread / frame)
//   pose_prep + lbs + joints_finalize   (smpl_kernels.cu: the full SMPL evaluation for every frame-person)
//   frame_residuals (1 thread / frame-person: projection, residuals, analytic gradients, warp-shuffle reductions)
//   camera_backward (1 thread / frame), camera_scatter (mode 3)
//   traj_backward   (1 CTA / person: reverse scans, variable gradients, regularisers)
//   reduce          (loss partials, fixed-camera gradient)   [-> optional NCCL allreduce by the caller]
//   adam            (loss terms + torch.optim.Adam update, step count on device => CUDA-graph capturable)
// No atomics: every sum is a fixed-order tree, so iterations are bit-reproducible.
#include <stdlib.h>
#include <string.h>

#include "globalopt_frames.cuh"
#include "smpl_model.cuh"
#include "block_scan.cuh"

namespace glamr {

constexpr int kFrameThreads = 128;

__device__ void block_reduce_terms(const TermAcc& acc, double* out /*[NUM_TERMS]*/, double* smem /*[warps][NUM_TERMS]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) {
    const double s = warp_sum(acc.v[k]);
    if (lane == 0) smem[wid * GLAMR_NUM_TERMS + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < GLAMR_NUM_TERMS) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += smem[w * GLAMR_NUM_TERMS + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(kScanThreads) traj_forward_kernel(OptCtx c) {
  __shared__ float sm[kScanThreads / 32 + 1];
  const int p = blockIdx.x;
  const glamr_person_t& ps = c.pb.persons[p];
  const int len = ps.len, T = c.pb.T;
  const size_t n0 = (size_t)p * T + ps.start;
  for (int i = threadIdx.x; i < len; i += kScanThreads) traj_pre(c, p, i);
  __syncthreads();
  block_scan_inplace(c.sc.heading + n0, len, 1, false, sm);
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += kScanThreads) traj_mid(c, p, i);
  __syncthreads();
  block_scan_inplace(c.sc.xy + 2 * n0, len, 2, false, sm);
  block_scan_inplace(c.sc.xy + 2 * n0 + 1, len, 2, false, sm);
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += kScanThreads) traj_post(c, p, t);
}

__global__ void __launch_bounds__(kFrameThreads) cam_forward_kernel(OptCtx c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < c.pb.T) cam_forward(c, t);
}

__global__ void __launch_bounds__(kFrameThreads) frame_residuals_kernel(OptCtx c, double* partial) {
  __shared__ double sm[(kFrameThreads / 32) * GLAMR_NUM_TERMS];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = c.pb.P * c.pb.T;
  TermAcc acc;
  acc.clear();
  if (n < N) {
    const int p = n / c.pb.T, t = n - p * c.pb.T;
    if (p >= c.pb.p_begin && p < c.pb.p_end) {
      frame_residuals(c, p, t, acc);
    } else {
      for (int k = 0; k < 3; ++k) { c.sc.g_orient[(size_t)n * 3 + k] = 0.0f; c.sc.g_trans[(size_t)n * 3 + k] = 0.0f; }
      for (int k = 0; k < 12; ++k) c.sc.g_cam[(size_t)n * 12 + k] = 0.0f;
    }
  }
  block_reduce_terms(acc, partial + (size_t)blockIdx.x * GLAMR_NUM_TERMS, sm);
}

__global__ void __launch_bounds__(kFrameThreads) camera_backward_kernel(OptCtx c, double* partial) {
  __shared__ double sm[(kFrameThreads / 32) * GLAMR_NUM_TERMS];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  TermAcc acc;
  acc.clear();
  if (t < c.pb.T) camera_backward(c, t, acc);
  block_reduce_terms(acc, partial + (size_t)blockIdx.x * GLAMR_NUM_TERMS, sm);
}

__global__ void __launch_bounds__(kFrameThreads) camera_scatter_kernel(OptCtx c) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < c.pb.T) camera_scatter_to_persons(c, s);
}

__global__ void __launch_bounds__(kScanThreads) traj_backward_kernel(OptCtx c, double* partial) {
  __shared__ float sm[kScanThreads / 32 + 1];
  __shared__ double smd[(kScanThreads / 32) * GLAMR_NUM_TERMS];
  const int p = blockIdx.x;
  const glamr_person_t& ps = c.pb.persons[p];
  const int len = ps.len, T = c.pb.T;
  const size_t n0 = (size_t)p * T + ps.start;
  TermAcc acc;
  acc.clear();
  for (int t = threadIdx.x; t < T; t += kScanThreads) traj_back_pre(c, p, t, acc);
  __syncthreads();
  block_scan_inplace(c.sc.g_xy + 2 * n0, len, 2, true, sm);
  block_scan_inplace(c.sc.g_xy + 2 * n0 + 1, len, 2, true, sm);
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += kScanThreads) traj_back_mid(c, p, i, acc);
  __syncthreads();
  block_scan_inplace(c.sc.g_head + n0, len, 1, true, sm);
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += kScanThreads) traj_back_post(c, p, i, acc);
  block_reduce_terms(acc, partial + (size_t)blockIdx.x * GLAMR_NUM_TERMS, smd);
}

// loss partials -> un-normalised term sums (reduce_buf tail); fixed camera: sum the per-frame gradients over T
__global__ void __launch_bounds__(256) reduce_kernel(OptCtx c, const double* partial, int n_slots, float* reduce_buf) {
  __shared__ double sm[8 * 16];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < GLAMR_NUM_TERMS) {
    double s = 0.0;
    for (int k = 0; k < n_slots; ++k) s += partial[(size_t)k * GLAMR_NUM_TERMS + tid];
    reduce_buf[c.pb.n_params + tid] = (float)s;
  }
  if (c.pb.cam_mode == GLAMR_CAM_FIXED) {
    double a[9];
    for (int k = 0; k < 9; ++k) a[k] = 0.0;
    for (int t = tid; t < c.pb.T; t += blockDim.x)
      for (int k = 0; k < 9; ++k) a[k] += (double)c.sc.g_cam_fix[(size_t)t * 12 + k];
    for (int k = 0; k < 9; ++k) {
      const double s = warp_sum(a[k]);
      if (lane == 0) sm[wid * 16 + k] = s;
    }
    __syncthreads();
    if (tid < 9) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w * 16 + tid];
      const int off = (tid < 6) ? c.pb.off_cam_rot + tid : c.pb.off_cam_trans + (tid - 6);
      reduce_buf[off] = (float)s;
    }
  }
}

struct AdamState {
  float* m;
  float* v;
  double* beta_pow;   // [2] running beta1^t, beta2^t ; beta_pow[2] holds the step count (as a double)
};

__global__ void __launch_bounds__(256) losses_kernel(OptCtx c, const float* __restrict__ reduce_buf, float* __restrict__ loss_terms,
                                                     const double* step_count, int hist_stride) {
  if (threadIdx.x == 0) {
    if (hist_stride > 0) loss_terms += (size_t)step_count[0] * hist_stride;
    double total = 0.0;
    for (int k = 0; k < GLAMR_NUM_TERMS; ++k) {
      float val = 0.0f;
      if (c.pb.term_enabled[k]) {
        val = reduce_buf[c.pb.n_params + k] / c.pb.term_norm[k];
        if (!c.pb.term_monitor[k]) total += (double)val * (double)c.pb.term_weight[k];
      }
      loss_terms[k] = val;
    }
    loss_terms[GLAMR_NUM_TERMS] = (float)total;
  }
}

__global__ void __launch_bounds__(256) adam_kernel(OptCtx c, float* __restrict__ theta, const float* __restrict__ reduce_buf, float lr,
                                                   AdamState ad, double lr_d) {
  // every thread derives the same bias corrections from the running powers; block 0 advances them afterwards
  const double b1 = ad.beta_pow[0] * 0.9, b2 = ad.beta_pow[1] * 0.999;
  const float bc1 = (float)(1.0 - b1);
  const float bc2s = (float)sqrt(1.0 - b2);
  const float step_size = (float)(lr_d / (1.0 - b1));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.pb.n_params; i += gridDim.x * blockDim.x) {
    if (!c.pb.active[i]) continue;
    const float g = reduce_buf[i];
    float m = ad.m[i], v = ad.v[i];
    m = m + 0.1f * (g - m);
    v = v * 0.999f + 0.001f * g * g;
    const float denom = sqrtf(v) / bc2s + 1e-8f;
    theta[i] = theta[i] - step_size * (m / denom);
    ad.m[i] = m;
    ad.v[i] = v;
  }
  (void)bc1;
  (void)lr;
}
__global__ void adam_advance_kernel(AdamState ad) {
  ad.beta_pow[0] *= 0.9;
  ad.beta_pow[1] *= 0.999;
  ad.beta_pow[2] += 1.0;
}

}  // namespace glamr

// =================================================================================================== C ABI
using namespace glamr;

struct glamr_opt {
  SmplDev smpl;
  glamr_problem_t pb;
  OptScratch sc;
  SmplWorkspace ws;
  AdamState adam;
  double* partial;
  int n_slots, slots_res, slots_cam;
  void* arena;
  size_t arena_bytes;
  float gs[GLAMR_NUM_TERMS];
  int timing;                 // != 0: bracket the LBS kernel with events (bench / roofline only, not graph-capturable)
  cudaEvent_t ev_lbs0, ev_lbs1;
};

extern "C" size_t glamr_sizeof_person(void) { return sizeof(glamr_person_t); }
extern "C" size_t glamr_sizeof_problem(void) { return sizeof(glamr_problem_t); }

static void compute_gs(glamr_opt* st) {
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) {
    const glamr_problem_t& pb = st->pb;
    st->gs[k] = (pb.term_enabled[k] && !pb.term_monitor[k] && pb.term_norm[k] != 0.0f) ? pb.term_weight[k] / pb.term_norm[k] : 0.0f;
  }
}
static OptCtx make_ctx(const glamr_opt* st, const float* theta, float* grad) {
  OptCtx c;
  c.pb = st->pb;
  c.sc = st->sc;
  c.sc.grad = grad;
  c.theta = theta;
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) c.gs[k] = st->gs[k];
  return c;
}

extern "C" int glamr_opt_create(glamr_opt_t** out, const glamr_smpl_t* smpl, const glamr_problem_t* pb) {
  if (!out || !smpl || !pb || pb->P <= 0 || pb->T <= 0 || pb->J <= 0 || pb->n_params <= 0) return GLAMR_EINVAL;
  if (pb->J != smpl->dev.n_map) return GLAMR_EINVAL;
  glamr_opt* st = (glamr_opt*)calloc(1, sizeof(glamr_opt));
  if (!st) return GLAMR_EINVAL;
  st->smpl = smpl->dev;
  st->pb = *pb;
  compute_gs(st);
  const size_t N = (size_t)pb->P * pb->T, T = pb->T, J = pb->J;
  st->slots_res = (int)((N + kFrameThreads - 1) / kFrameThreads);
  st->slots_cam = (int)((T + kFrameThreads - 1) / kFrameThreads);
  st->n_slots = st->slots_res + st->slots_cam + pb->P;
  // one arena for all scratch (floats), doubles first for alignment
  size_t floats = 0;
  auto take = [&](size_t nfl) { size_t o = floats; floats += (nfl + 63) & ~(size_t)63; return o; };
  const size_t o_partial = take((size_t)st->n_slots * GLAMR_NUM_TERMS * 2);
  const size_t o_beta = take(4);
  const size_t o_heading = take(N), o_xy = take(2 * N), o_tl = take(11 * N), o_ob = take(3 * N), o_tb = take(3 * N),
               o_ow = take(3 * N), o_tw = take(3 * N), o_cam = take(12 * T), o_caminv = take(12 * T), o_camd6 = take(6 * T),
               o_jw = take(N * J * 3), o_kp = take(N * J * 2), o_ociw = take(3 * N), o_tciw = take(3 * N), o_go = take(3 * N),
               o_gt = take(3 * N), o_gcam = take(12 * N), o_gcf = take(12 * T), o_gxy = take(2 * N), o_gh = take(N),
               o_m = take(pb->n_params), o_v = take(pb->n_params);
  const size_t o_ws = take(smpl_workspace_floats((int)N, smpl->dev.S));
  st->arena_bytes = floats * sizeof(float);
  cudaError_t e = cudaMalloc(&st->arena, st->arena_bytes);
  if (e != cudaSuccess) { free(st); return (int)e; }
  e = cudaMemset(st->arena, 0, st->arena_bytes);
  if (e != cudaSuccess) { cudaFree(st->arena); free(st); return (int)e; }
  float* b = (float*)st->arena;
  st->partial = (double*)(b + o_partial);
  st->adam.beta_pow = (double*)(b + o_beta);
  st->sc.heading = b + o_heading; st->sc.xy = b + o_xy; st->sc.traj_local = b + o_tl; st->sc.orient_base = b + o_ob;
  st->sc.trans_base = b + o_tb; st->sc.orient_world = b + o_ow; st->sc.trans_world = b + o_tw; st->sc.cam = b + o_cam;
  st->sc.cam_inv = b + o_caminv; st->sc.cam_d6 = b + o_camd6; st->sc.joints_world = b + o_jw; st->sc.kp_pred = b + o_kp;
  st->sc.orient_ciw = b + o_ociw; st->sc.trans_ciw = b + o_tciw; st->sc.g_orient = b + o_go; st->sc.g_trans = b + o_gt;
  st->sc.g_cam = b + o_gcam; st->sc.g_cam_fix = b + o_gcf; st->sc.g_xy = b + o_gxy; st->sc.g_head = b + o_gh;
  st->sc.grad = nullptr;
  st->adam.m = b + o_m; st->adam.v = b + o_v;
  st->ws = smpl_carve_workspace(b + o_ws, (int)N, smpl->dev.S);
  const double one[3] = {1.0, 1.0, 0.0};
  e = cudaMemcpy(st->adam.beta_pow, one, sizeof(one), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(st->arena); free(st); return (int)e; }
  *out = st;
  return GLAMR_OK;
}

extern "C" int glamr_opt_kernel_timing(glamr_opt_t* st, int enable) {
  if (!st) return GLAMR_EINVAL;
  if (enable && !st->ev_lbs0) {
    GLAMR_CUDA_TRY(cudaEventCreate(&st->ev_lbs0));
    GLAMR_CUDA_TRY(cudaEventCreate(&st->ev_lbs1));
  }
  st->timing = enable;
  return GLAMR_OK;
}

extern "C" int glamr_opt_last_lbs_ms(glamr_opt_t* st, float* ms) {
  if (!st || !ms || !st->ev_lbs0) return GLAMR_EINVAL;
  GLAMR_CUDA_TRY(cudaEventSynchronize(st->ev_lbs1));
  GLAMR_CUDA_TRY(cudaEventElapsedTime(ms, st->ev_lbs0, st->ev_lbs1));
  return GLAMR_OK;
}

extern "C" int glamr_opt_destroy(glamr_opt_t* st) {
  if (!st) return GLAMR_OK;
  if (st->ev_lbs0) { cudaEventDestroy(st->ev_lbs0); cudaEventDestroy(st->ev_lbs1); }
  cudaFree(st->arena);
  free(st);
  return GLAMR_OK;
}

extern "C" int glamr_opt_set_problem(glamr_opt_t* st, const glamr_problem_t* pb, int reset_adam, void* stream) {
  if (!st || !pb) return GLAMR_EINVAL;
  if (pb->P != st->pb.P || pb->T != st->pb.T || pb->J != st->pb.J || pb->n_params != st->pb.n_params) return GLAMR_EINVAL;
  st->pb = *pb;
  compute_gs(st);
  if (reset_adam) {
    cudaStream_t s = (cudaStream_t)stream;
    GLAMR_CUDA_TRY(cudaMemsetAsync(st->adam.m, 0, sizeof(float) * pb->n_params, s));
    GLAMR_CUDA_TRY(cudaMemsetAsync(st->adam.v, 0, sizeof(float) * pb->n_params, s));
    static const double one[3] = {1.0, 1.0, 0.0};
    GLAMR_CUDA_TRY(cudaMemcpyAsync(st->adam.beta_pow, one, sizeof(one), cudaMemcpyHostToDevice, s));
  }
  return GLAMR_OK;
}

extern "C" size_t glamr_opt_reduce_count(const glamr_opt_t* st) { return st ? (size_t)st->pb.n_params + GLAMR_NUM_TERMS : 0; }

extern "C" int glamr_opt_backward(glamr_opt_t* st, const float* theta, float* reduce_buf, void* stream) {
  if (!st || !theta || !reduce_buf) return GLAMR_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const glamr_problem_t& pb = st->pb;
  const int N = pb.P * pb.T;
  OptCtx c = make_ctx(st, theta, reduce_buf);
  GLAMR_CUDA_TRY(cudaMemsetAsync(reduce_buf, 0, sizeof(float) * ((size_t)pb.n_params + GLAMR_NUM_TERMS), s));
  traj_forward_kernel<<<pb.P, kScanThreads, 0, s>>>(c);
  GLAMR_LAUNCH_CHECK();
  cam_forward_kernel<<<st->slots_cam, kFrameThreads, 0, s>>>(c);
  GLAMR_LAUNCH_CHECK();
  // SMPL for the persons this rank owns (global_recon_model.py:517-524)
  const int n_begin = pb.p_begin * pb.T, n_end = pb.p_end * pb.T;
  if (n_end > n_begin) {
    SmplWorkspace w = st->ws;
    // the workspace is indexed by the global frame-person index; kernels take [0, n) so offset the pointers
    SmplWorkspace wo = w;
    wo.A += (size_t)n_begin * kNJ * 12; wo.pf += (size_t)n_begin * kPFPad; wo.jposed += (size_t)n_begin * kNJ * 3;
    wo.vcompact += (size_t)n_begin * st->smpl.S * 3; wo.root_raw += (size_t)n_begin * 3;
    const int nn = n_end - n_begin;
    int rc;
    if ((rc = launch_pose_prep(st->smpl, nn, st->sc.orient_world + (size_t)n_begin * 3, pb.smpl_pose_all + (size_t)n_begin * 69,
                               pb.smpl_beta_all + (size_t)n_begin * kNB, 1, wo, s))) return rc;
    if (st->timing) GLAMR_CUDA_TRY(cudaEventRecord(st->ev_lbs0, s));
    if ((rc = launch_lbs(st->smpl, 0, nn, pb.smpl_beta_all + (size_t)n_begin * kNB, wo, nullptr, s))) return rc;
    if (st->timing) GLAMR_CUDA_TRY(cudaEventRecord(st->ev_lbs1, s));
    if ((rc = launch_joints_finalize(st->smpl, nn, 0, st->sc.trans_world + (size_t)n_begin * 3,
                                     pb.scale_all ? pb.scale_all + n_begin : nullptr, wo,
                                     st->sc.joints_world + (size_t)n_begin * pb.J * 3, s))) return rc;
  }
  frame_residuals_kernel<<<st->slots_res, kFrameThreads, 0, s>>>(c, st->partial);
  GLAMR_LAUNCH_CHECK();
  camera_backward_kernel<<<st->slots_cam, kFrameThreads, 0, s>>>(c, st->partial + (size_t)st->slots_res * GLAMR_NUM_TERMS);
  GLAMR_LAUNCH_CHECK();
  if (pb.cam_mode == GLAMR_CAM_FROM_PERSONS) {
    camera_scatter_kernel<<<st->slots_cam, kFrameThreads, 0, s>>>(c);
    GLAMR_LAUNCH_CHECK();
  }
  traj_backward_kernel<<<pb.P, kScanThreads, 0, s>>>(c, st->partial + (size_t)(st->slots_res + st->slots_cam) * GLAMR_NUM_TERMS);
  GLAMR_LAUNCH_CHECK();
  reduce_kernel<<<1, 256, 0, s>>>(c, st->partial, st->n_slots, reduce_buf);
  GLAMR_LAUNCH_CHECK();
  (void)N;
  return GLAMR_OK;
}

extern "C" int glamr_opt_losses(glamr_opt_t* st, const float* reduce_buf, float* loss_terms, void* stream) {
  if (!st || !reduce_buf || !loss_terms) return GLAMR_EINVAL;
  OptCtx c = make_ctx(st, nullptr, nullptr);
  losses_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c, reduce_buf, loss_terms, st->adam.beta_pow + 2, 0);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}

extern "C" int glamr_opt_apply(glamr_opt_t* st, float* theta, const float* reduce_buf, double lr, float* loss_terms,
                               int loss_hist_stride, void* stream) {
  if (!st || !theta || !reduce_buf) return GLAMR_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  OptCtx c = make_ctx(st, theta, nullptr);
  if (loss_terms) {
    losses_kernel<<<1, 32, 0, s>>>(c, reduce_buf, loss_terms, st->adam.beta_pow + 2, loss_hist_stride);
    GLAMR_LAUNCH_CHECK();
  }
  const int blocks = (st->pb.n_params + 255) / 256;
  adam_kernel<<<blocks < 592 ? blocks : 592, 256, 0, s>>>(c, theta, reduce_buf, (float)lr, st->adam, lr);
  GLAMR_LAUNCH_CHECK();
  adam_advance_kernel<<<1, 1, 0, s>>>(st->adam);
  GLAMR_LAUNCH_CHECK();
  return GLAMR_OK;
}

extern "C" int glamr_opt_read(glamr_opt_t* st, int what, const float** ptr, size_t* count) {
  if (!st || !ptr || !count) return GLAMR_EINVAL;
  const size_t N = (size_t)st->pb.P * st->pb.T, T = st->pb.T, J = st->pb.J;
  switch (what) {
    case GLAMR_R_ORIENT_WORLD: *ptr = st->sc.orient_world; *count = 3 * N; break;
    case GLAMR_R_TRANS_WORLD: *ptr = st->sc.trans_world; *count = 3 * N; break;
    case GLAMR_R_ORIENT_BASE: *ptr = st->sc.orient_base; *count = 3 * N; break;
    case GLAMR_R_TRANS_BASE: *ptr = st->sc.trans_base; *count = 3 * N; break;
    case GLAMR_R_KP_PRED: *ptr = st->sc.kp_pred; *count = N * J * 2; break;
    case GLAMR_R_ORIENT_CAM_IN_WORLD: *ptr = st->sc.orient_ciw; *count = 3 * N; break;
    case GLAMR_R_TRANS_CAM_IN_WORLD: *ptr = st->sc.trans_ciw; *count = 3 * N; break;
    case GLAMR_R_CAM_POSE: *ptr = st->sc.cam; *count = 12 * T; break;
    case GLAMR_R_CAM_POSE_INV: *ptr = st->sc.cam_inv; *count = 12 * T; break;
    case GLAMR_R_JOINTS_WORLD: *ptr = st->sc.joints_world; *count = N * J * 3; break;
    case GLAMR_R_TRAJ_LOCAL: *ptr = st->sc.traj_local; *count = 11 * N; break;
    case GLAMR_R_SMPL_A: *ptr = st->ws.A; *count = N * kNJ * 12; break;
    default: return GLAMR_EINVAL;
  }
  return GLAMR_OK;
}
