// Row-wise dispatch over the primitives of glamr_math.cuh: one row of in0 (and in1) -> one row of out, plus the
// vector-Jacobian product.  The CUDA library exposes it as glamr_rowop_fwd / glamr_rowop_vjp (c_api.cu); the host
// test harness (tests/host_harness) compiles the same dispatch with g++.
#pragma once
#include "glamr_math.cuh"

namespace glamr {

enum RowOp : int {
  ROP_AA_TO_ROTMAT = 0,     // [3] -> [9]        konia angle_axis_to_rotation_matrix
  ROP_RODRIGUES_SMPLX = 1,  // [3] -> [9]        smplx batch_rodrigues
  ROP_ROT6D_TO_ROTMAT = 2,  // [6] -> [9]
  ROP_ROTMAT_TO_QUAT = 3,   // [9] -> [4]
  ROP_QUAT_TO_AA = 4,       // [4] -> [3]
  ROP_AA_TO_QUAT = 5,       // [3] -> [4]
  ROP_QUAT_MUL = 6,         // [4],[4] -> [4]
  ROP_ROTMAT_TO_AA = 7,     // [9] -> [3]
  ROP_QUAT_TO_ROTMAT = 8,   // [4] -> [9]        (forward only)
  ROP_SAFE_ATAN2 = 9,       // [2]=(y,x) -> [1]
  ROP_PROJECT = 10,         // [3]=X, in1 [9]=K -> [2]   (vjp w.r.t. X only)
  ROP_MAT3_MUL = 11,        // [9],[9] -> [9]
  ROP_COUNT = 12
};

GLAMR_HD void rowop_dims(int op, int& d0, int& d1, int& dout) {
  d1 = 0;
  switch (op) {
    case ROP_AA_TO_ROTMAT: case ROP_RODRIGUES_SMPLX: d0 = 3; dout = 9; break;
    case ROP_ROT6D_TO_ROTMAT: d0 = 6; dout = 9; break;
    case ROP_ROTMAT_TO_QUAT: d0 = 9; dout = 4; break;
    case ROP_QUAT_TO_AA: d0 = 4; dout = 3; break;
    case ROP_AA_TO_QUAT: d0 = 3; dout = 4; break;
    case ROP_QUAT_MUL: d0 = 4; d1 = 4; dout = 4; break;
    case ROP_ROTMAT_TO_AA: d0 = 9; dout = 3; break;
    case ROP_QUAT_TO_ROTMAT: d0 = 4; dout = 9; break;
    case ROP_SAFE_ATAN2: d0 = 2; dout = 1; break;
    case ROP_PROJECT: d0 = 3; d1 = 9; dout = 2; break;
    case ROP_MAT3_MUL: d0 = 9; d1 = 9; dout = 9; break;
    default: d0 = 0; dout = 0; break;
  }
}

GLAMR_HD void rowop_fwd(int op, const float* a, const float* b, float* o) {
  switch (op) {
    case ROP_AA_TO_ROTMAT: aa_to_rotmat(a, o); break;
    case ROP_RODRIGUES_SMPLX: rodrigues_smplx(a, o); break;
    case ROP_ROT6D_TO_ROTMAT: rot6d_to_rotmat(a, o); break;
    case ROP_ROTMAT_TO_QUAT: rotmat_to_quat(a, o); break;
    case ROP_QUAT_TO_AA: quat_to_aa(a, o); break;
    case ROP_AA_TO_QUAT: aa_to_quat(a, o); break;
    case ROP_QUAT_MUL: quat_mul(a, b, o); break;
    case ROP_ROTMAT_TO_AA: rotmat_to_aa(a, o); break;
    case ROP_QUAT_TO_ROTMAT: quat_to_rotmat(a, o); break;
    case ROP_SAFE_ATAN2: o[0] = safe_atan2(a[0], a[1]); break;
    case ROP_PROJECT: project(b, a, o); break;
    case ROP_MAT3_MUL: mat3_mul(a, b, o); break;
    default: break;
  }
}

// ga / gb may be nullptr when that input has no gradient
GLAMR_HD void rowop_vjp(int op, const float* a, const float* b, const float* g, float* ga, float* gb) {
  switch (op) {
    case ROP_AA_TO_ROTMAT: aa_to_rotmat_vjp(a, g, ga); break;
    case ROP_RODRIGUES_SMPLX: rodrigues_smplx_vjp(a, g, ga); break;
    case ROP_ROT6D_TO_ROTMAT: rot6d_to_rotmat_vjp(a, g, ga); break;
    case ROP_ROTMAT_TO_QUAT: rotmat_to_quat_vjp(a, g, ga); break;
    case ROP_QUAT_TO_AA: quat_to_aa_vjp(a, g, ga); break;
    case ROP_AA_TO_QUAT: aa_to_quat_vjp(a, g, ga); break;
    case ROP_QUAT_MUL: quat_mul_vjp(a, b, g, ga, gb); break;
    case ROP_ROTMAT_TO_AA: rotmat_to_aa_vjp(a, g, ga); break;
    case ROP_SAFE_ATAN2: safe_atan2_vjp(a[0], a[1], g[0], ga[0], ga[1]); break;
    case ROP_PROJECT: project_vjp(b, a, g, ga); break;
    case ROP_MAT3_MUL:
      if (ga) mat3_mult(g, b, ga);   // g * B^T
      if (gb) mat3_tmul(a, g, gb);   // A^T * g
      break;
    default: break;
  }
}

}  // namespace glamr
