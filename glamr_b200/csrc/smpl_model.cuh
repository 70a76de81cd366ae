// Device-resident SMPL constants, re-tiled for the kernels of smpl_kernels.cu (layout is documented in DESIGN.md).
#pragma once
#include "common.cuh"

namespace glamr {

struct SmplDev {
  // dense constants
  const float* pd_tiles;     // [54][207][384]  posedirs, vertex-tile major: one contiguous 13,824 B block per
                             //                 (tile, 9-row chunk) so a CTA streams its slab with 1-D bulk TMA
  const float* v_template;   // [6912][3]  (padded with zeros)
  const float* shapedirs;    // [6912][30] ([v][c][l] as in the model file)
  const float* j_template;   // [24][3]    J_regressor @ v_template
  const float* j_shapedirs;  // [24][3][10] J_regressor @ shapedirs   (J(beta) = j_template + j_shapedirs . beta)
  // K-sparse skinning weights (K = max non-zeros per vertex, rows zero-padded)
  const float* skin_w;       // [6912][K]
  const uint8_t* skin_j;     // [6912][K]
  int K;
  // vertices that feed a picked joint or an extra-regressor row ("support"), compacted
  const int32_t* compact_of_vertex;  // [6912] index into the support list or -1
  int S;                     // support size
  const int32_t* reg_ptr;    // [n_extra+1]  CSR over J_regressor_extra rows
  const int32_t* reg_ci;     // [nnz] compact vertex index
  const float* reg_w;        // [nnz]
  const int32_t* pick_ci;    // [n_picks] compact vertex index of each picked vertex
  const int32_t* joint_map;  // [n_map] into [24 | n_picks | n_extra]
  int n_extra, n_picks, n_map;
  int parents[kNJ];
  int level[kNJ];
  int n_levels;
};

struct SmplWorkspace {
  float* A;         // [n/32][24][32][12] relative joint transforms (3x4 row-major per joint), tile-major
  float* pf;        // [n/32][23][9][32] pose feature (R_j - I), j = 1..23, tile-major for bulk TMA (208 floats per frame reserved)
  float* jposed;    // [n][24][3]   posed LBS joints
  float* vcompact;  // [n][S][3]    skinned support vertices
  float* root_raw;  // [n][3]       un-rooted joint 0 (for vertex re-rooting)
};

inline size_t smpl_workspace_floats(int n, int S) {
  const size_t n32 = ((size_t)n + 31) / 32 * 32;   // the pose feature is tile-major over whole 32-frame tiles
  return (size_t)n * (kNJ * 3 + (size_t)S * 3 + 3) + n32 * (kPFPad + kNJ * 12) + 64;
}
inline SmplWorkspace smpl_carve_workspace(void* base, int n, int S) {
  SmplWorkspace w;
  float* p = (float*)base;
  w.A = p; p += ((size_t)n + 31) / 32 * 32 * kNJ * 12;
  w.pf = p; p += ((size_t)n + 31) / 32 * 32 * kPFPad;
  w.jposed = p; p += (size_t)n * kNJ * 3;
  w.vcompact = p; p += (size_t)n * S * 3;
  w.root_raw = p;
  return w;
}

#if defined(__CUDACC__)
// un-rooted joint `idx` of [24 LBS | picks | extra regressed] for local frame-person f  (lib/models/smpl.py:299-301)
__device__ __forceinline__ void raw_joint(const SmplDev& m, const SmplWorkspace& w, int f, int idx, float* o) {
  if (idx < kNJ) {
    const float* p = w.jposed + ((size_t)f * kNJ + idx) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
  } else if (idx < kNJ + m.n_picks) {
    const float* p = w.vcompact + ((size_t)f * m.S + m.pick_ci[idx - kNJ]) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
  } else {
    const int r = idx - kNJ - m.n_picks;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int e = m.reg_ptr[r]; e < m.reg_ptr[r + 1]; ++e) {
      const float* p = w.vcompact + ((size_t)f * m.S + m.reg_ci[e]) * 3;
      const float wt = m.reg_w[e];
      a0 = fmaf(wt, p[0], a0); a1 = fmaf(wt, p[1], a1); a2 = fmaf(wt, p[2], a2);
    }
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}
#endif

// launches (smpl_kernels.cu); all asynchronous on `s`
// orient may be NULL (zeros).  use_betas == 0 -> rest joints from the template only (SMPL.get_joints).
int launch_pose_prep(const SmplDev& m, int n, const float* orient, const float* body_pose, const float* betas,
                     int use_betas, const SmplWorkspace& w, cudaStream_t s, bool pdl = false);
// n_begin..n_end: frame-person range to skin.  vertices may be NULL.
int launch_lbs(const SmplDev& m, int n_begin, int n_end, const float* betas, const SmplWorkspace& w, float* vertices,
               cudaStream_t s, bool pdl = false);
int launch_joints_finalize(const SmplDev& m, int n, int orig_joints, const float* root_trans, const float* root_scale,
                           const SmplWorkspace& w, float* joints, cudaStream_t s);
int launch_reroot_vertices(int n, const float* root_raw, const float* root_trans, const float* root_scale,
                           float* vertices, cudaStream_t s);

}  // namespace glamr

struct glamr_smpl {
  glamr::SmplDev dev;
  void* allocs[16];
  int n_allocs;
};
