// TEST INFRASTRUCTURE ONLY: sequential host driver for the frame functions of glamr_b200/csrc/globalopt_frames.cuh.
// Mirrors the kernel sequence of globalopt_kernels.cu (prefix scans done with plain loops) so that the analytic
// backward can be validated against torch autograd on the GPU-less build box.  Never used by the product.
#include <math.h>
#include <string.h>

#include <vector>

#include "../../glamr_b200/csrc/globalopt_frames.cuh"

using namespace glamr;

struct EmuHandle {
  glamr_problem_t pb;
  std::vector<float> buf[24];
  OptScratch sc;
  std::vector<float> m, v;
  double beta_pow[3];
};

static OptCtx make_ctx(EmuHandle* h, const float* theta, float* grad) {
  OptCtx c;
  c.pb = h->pb;
  c.sc = h->sc;
  c.sc.grad = grad;
  c.theta = theta;
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) {
    const glamr_problem_t& pb = h->pb;
    c.gs[k] = (pb.term_enabled[k] && !pb.term_monitor[k] && pb.term_norm[k] != 0.0f) ? pb.term_weight[k] / pb.term_norm[k] : 0.0f;
  }
  return c;
}

extern "C" {

size_t glamr_host_sizeof_problem() { return sizeof(glamr_problem_t); }
size_t glamr_host_sizeof_person() { return sizeof(glamr_person_t); }

int glamr_host_emu_create(EmuHandle** out, const glamr_problem_t* pb) {
  EmuHandle* h = new EmuHandle();
  h->pb = *pb;
  const size_t N = (size_t)pb->P * pb->T, T = pb->T, J = pb->J;
  int i = 0;
  auto take = [&](size_t n) { h->buf[i].assign(n, 0.0f); return h->buf[i++].data(); };
  h->sc.heading = take(N); h->sc.xy = take(2 * N); h->sc.traj_local = take(11 * N); h->sc.orient_base = take(3 * N);
  h->sc.trans_base = take(3 * N); h->sc.orient_world = take(3 * N); h->sc.trans_world = take(3 * N); h->sc.cam = take(12 * T);
  h->sc.cam_inv = take(12 * T); h->sc.cam_d6 = take(6 * T); h->sc.joints_world = take(N * J * 3); h->sc.kp_pred = take(N * J * 2);
  h->sc.orient_ciw = take(3 * N); h->sc.trans_ciw = take(3 * N); h->sc.g_orient = take(3 * N); h->sc.g_trans = take(3 * N);
  h->sc.g_cam = take(12 * N); h->sc.g_cam_fix = take(12 * T); h->sc.g_xy = take(2 * N); h->sc.g_head = take(N);
  h->sc.grad = nullptr;
  h->m.assign(pb->n_params, 0.0f);
  h->v.assign(pb->n_params, 0.0f);
  h->beta_pow[0] = h->beta_pow[1] = 1.0; h->beta_pow[2] = 0.0;
  *out = h;
  return 0;
}
int glamr_host_emu_destroy(EmuHandle* h) { delete h; return 0; }
int glamr_host_emu_set_problem(EmuHandle* h, const glamr_problem_t* pb, int reset_adam) {
  h->pb = *pb;
  if (reset_adam) {
    std::fill(h->m.begin(), h->m.end(), 0.0f);
    std::fill(h->v.begin(), h->v.end(), 0.0f);
    h->beta_pow[0] = h->beta_pow[1] = 1.0; h->beta_pow[2] = 0.0;
  }
  return 0;
}

static void scan(float* d, int count, int stride, bool reverse) {
  float run = 0.0f;
  for (int k = 0; k < count; ++k) {
    const int idx = reverse ? count - 1 - k : k;
    run += d[(size_t)idx * stride];
    d[(size_t)idx * stride] = run;
  }
}

// trajectory + camera forward: fills orient_world / trans_world / cam
int glamr_host_emu_forward_pose(EmuHandle* h, const float* theta) {
  OptCtx c = make_ctx(h, theta, nullptr);
  const glamr_problem_t& pb = h->pb;
  for (int p = 0; p < pb.P; ++p) {
    const glamr_person_t& ps = pb.persons[p];
    const size_t n0 = (size_t)p * pb.T + ps.start;
    for (int i = 0; i < ps.len; ++i) traj_pre(c, p, i);
    scan(c.sc.heading + n0, ps.len, 1, false);
    for (int i = 0; i < ps.len; ++i) traj_mid(c, p, i);
    scan(c.sc.xy + 2 * n0, ps.len, 2, false);
    scan(c.sc.xy + 2 * n0 + 1, ps.len, 2, false);
    for (int t = 0; t < pb.T; ++t) traj_post(c, p, t);
  }
  for (int t = 0; t < pb.T; ++t) cam_forward(c, t);
  return 0;
}

int glamr_host_emu_buffer(EmuHandle* h, int what, float** ptr, size_t* count) {
  const size_t N = (size_t)h->pb.P * h->pb.T, T = h->pb.T, J = h->pb.J;
  switch (what) {
    case GLAMR_R_ORIENT_WORLD: *ptr = h->sc.orient_world; *count = 3 * N; break;
    case GLAMR_R_TRANS_WORLD: *ptr = h->sc.trans_world; *count = 3 * N; break;
    case GLAMR_R_ORIENT_BASE: *ptr = h->sc.orient_base; *count = 3 * N; break;
    case GLAMR_R_TRANS_BASE: *ptr = h->sc.trans_base; *count = 3 * N; break;
    case GLAMR_R_KP_PRED: *ptr = h->sc.kp_pred; *count = N * J * 2; break;
    case GLAMR_R_ORIENT_CAM_IN_WORLD: *ptr = h->sc.orient_ciw; *count = 3 * N; break;
    case GLAMR_R_TRANS_CAM_IN_WORLD: *ptr = h->sc.trans_ciw; *count = 3 * N; break;
    case GLAMR_R_CAM_POSE: *ptr = h->sc.cam; *count = 12 * T; break;
    case GLAMR_R_CAM_POSE_INV: *ptr = h->sc.cam_inv; *count = 12 * T; break;
    case GLAMR_R_JOINTS_WORLD: *ptr = h->sc.joints_world; *count = N * J * 3; break;
    case GLAMR_R_TRAJ_LOCAL: *ptr = h->sc.traj_local; *count = 11 * N; break;
    case 100: *ptr = h->sc.g_orient; *count = 3 * N; break;      // test-only: dL/d smpl_orient_world of the last backward
    default: return -1;
  }
  return 0;
}

// residuals + backward; joints_world must have been filled by the caller (SMPL is evaluated by the oracle in tests)
int glamr_host_emu_backward(EmuHandle* h, const float* theta, float* reduce_buf) {
  const glamr_problem_t& pb = h->pb;
  memset(reduce_buf, 0, sizeof(float) * ((size_t)pb.n_params + GLAMR_NUM_TERMS));
  OptCtx c = make_ctx(h, theta, reduce_buf);
  TermAcc acc;
  acc.clear();
  for (int p = 0; p < pb.P; ++p)
    for (int t = 0; t < pb.T; ++t) {
      const size_t n = (size_t)p * pb.T + t;
      if ((int)n >= pb.n_begin && (int)n < pb.n_end) frame_residuals(c, p, t, acc);
      else {
        for (int k = 0; k < 3; ++k) { c.sc.g_orient[n * 3 + k] = 0; c.sc.g_trans[n * 3 + k] = 0; }
        for (int k = 0; k < 12; ++k) c.sc.g_cam[n * 12 + k] = 0;
      }
    }
  for (int t = 0; t < pb.T; ++t) camera_backward(c, t, acc);
  if (pb.cam_mode == GLAMR_CAM_FROM_PERSONS)
    for (int s = 0; s < pb.T; ++s) camera_scatter_to_persons(c, s);
  for (int p = 0; p < pb.P; ++p) {
    const glamr_person_t& ps = pb.persons[p];
    const size_t n0 = (size_t)p * pb.T + ps.start;
    for (int t = 0; t < pb.T; ++t) traj_back_pre(c, p, t, acc);
    scan(c.sc.g_xy + 2 * n0, ps.len, 2, true);
    scan(c.sc.g_xy + 2 * n0 + 1, ps.len, 2, true);
    for (int i = 0; i < ps.len; ++i) traj_back_mid(c, p, i, acc);
    scan(c.sc.g_head + n0, ps.len, 1, true);
    for (int i = 0; i < ps.len; ++i) traj_back_post(c, p, i, acc);
  }
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) reduce_buf[pb.n_params + k] = (float)acc.v[k];
  if (pb.cam_mode == GLAMR_CAM_FIXED) {
    double a[9] = {0};
    for (int t = 0; t < pb.T; ++t)
      for (int k = 0; k < 9; ++k) a[k] += c.sc.g_cam_fix[(size_t)t * 12 + k];
    for (int k = 0; k < 9; ++k) reduce_buf[(k < 6) ? pb.off_cam_rot + k : pb.off_cam_trans + (k - 6)] = (float)a[k];
  }
  return 0;
}

int glamr_host_emu_losses(EmuHandle* h, const float* reduce_buf, float* loss_terms) {
  const glamr_problem_t& pb = h->pb;
  double total = 0.0;
  for (int k = 0; k < GLAMR_NUM_TERMS; ++k) {
    float val = 0.0f;
    if (pb.term_enabled[k]) {
      val = reduce_buf[pb.n_params + k] / pb.term_norm[k];
      if (!pb.term_monitor[k]) total += (double)val * (double)pb.term_weight[k];
    }
    loss_terms[k] = val;
  }
  loss_terms[GLAMR_NUM_TERMS] = (float)total;
  return 0;
}

int glamr_host_emu_adam(EmuHandle* h, float* theta, const float* reduce_buf, double lr) {
  const glamr_problem_t& pb = h->pb;
  const double b1 = h->beta_pow[0] * 0.9, b2 = h->beta_pow[1] * 0.999;
  const float bc2s = (float)sqrt(1.0 - b2);
  const float step_size = (float)(lr / (1.0 - b1));
  for (int i = 0; i < pb.n_params; ++i) {
    if (!pb.active[i]) continue;
    const float g = reduce_buf[i];
    float m = h->m[i], v = h->v[i];
    m = m + 0.1f * (g - m);
    v = v * 0.999f + 0.001f * g * g;
    const float denom = sqrtf(v) / bc2s + 1e-8f;
    theta[i] = theta[i] - step_size * (m / denom);
    h->m[i] = m; h->v[i] = v;
  }
  h->beta_pow[0] = b1; h->beta_pow[1] = b2; h->beta_pow[2] += 1.0;
  return 0;
}
}
