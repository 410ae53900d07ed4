// tests/host_math.cpp -- TEST HARNESS: runs the __host__ __device__ kernel arithmetic of csrc/ba_math.cuh on the CPU
// so that it can be compared with the oracle without a GPU.  Not part of the product library.
#include "../sfm-toy-library_b200/csrc/ba_math.cuh"
extern "C" {
void host_obs_eval(const double* cam, const double* X, double f, double ox, double oy, double* r, double* Jc, double* Jp, double* Jf) {
    CamDerived d; cam_derive(cam, d); obs_eval(d, X, f, ox, oy, r, Jc, Jp, Jf);
}
void host_obs_residual(const double* cam, const double* X, double f, double ox, double oy, double* r) {
    CamDerived d; cam_derive(cam, d); obs_residual(d, X, f, ox, oy, r);
}
int host_chol3_inverse(const double* U, double* M) { return chol3_inverse(U, M) ? 1 : 0; }
}
