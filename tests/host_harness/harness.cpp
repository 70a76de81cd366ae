// TEST INFRASTRUCTURE ONLY.  Compiles the device-function headers of glamr_b200/csrc with g++ so that the math
// (forward values and hand-derived VJPs) can be checked against torch autograd on the GPU-less build box.
// The product never loads this library; its compute path is the CUDA library and fails loudly without it.
#include "../../glamr_b200/csrc/rowops.cuh"
#include "../../glamr_b200/csrc/eval_math.cuh"

extern "C" {

int glamr_host_rowop_dims(int op, int* d0, int* d1, int* dout) {
  glamr::rowop_dims(op, *d0, *d1, *dout);
  return (*dout > 0) ? 0 : 1;
}

int glamr_host_rowop_fwd(int op, int n, const float* in0, const float* in1, float* out) {
  int d0, d1, dout;
  glamr::rowop_dims(op, d0, d1, dout);
  if (dout == 0) return 1;
  for (int i = 0; i < n; ++i) glamr::rowop_fwd(op, in0 + (long)i * d0, in1 ? in1 + (long)i * d1 : nullptr, out + (long)i * dout);
  return 0;
}

int glamr_host_rowop_vjp(int op, int n, const float* in0, const float* in1, const float* gout, float* gin0, float* gin1) {
  int d0, d1, dout;
  glamr::rowop_dims(op, d0, d1, dout);
  if (dout == 0) return 1;
  for (int i = 0; i < n; ++i)
    glamr::rowop_vjp(op, in0 + (long)i * d0, in1 ? in1 + (long)i * d1 : nullptr, gout + (long)i * dout,
                     gin0 ? gin0 + (long)i * d0 : nullptr, gin1 ? gin1 + (long)i * d1 : nullptr);
  return 0;
}

int glamr_host_procrustes(int n, int J, const float* S1, const float* S2, float* out) {
  for (int f = 0; f < n; ++f) glamr::procrustes_frame(J, S1 + (long)f * J * 3, S2 + (long)f * J * 3, out + (long)f * J * 3);
  return 0;
}
}
