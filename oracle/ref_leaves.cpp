// TEST INFRASTRUCTURE.  C entry points over leaf functions of the REAL reference, compiled from /root/reference by `make -C oracle ref`
// into oracle/_ref/libg2o_ref_leaves.so (together with this file; see the Makefile for the list of reference sources):
//   robust kernels  g2o/core/robust_kernel_impl.cpp:50-181, constructed by name through the reference's own RobustKernelFactory
//   dq/dR           g2o/types/slam3d/dquat2mat.cpp:35-85 + dquat2mat_maxima_generated.cpp
//   normalize_theta g2o/stuff/misc.h:114-127
//   sampleGaussian  g2o/stuff/sampler.cpp:31-45 (one static std::normal_distribution shared by every engine - the noise source of create_sphere)
// tests/test_reference_leaves.py checks the oracle's restatements (and, on the GPU, the device functions through them) against these.
#include <cstring>

#include "g2o/core/robust_kernel.h"
#include "g2o/core/robust_kernel_factory.h"
#include "g2o/stuff/misc.h"
#include "g2o/stuff/sampler.h"
#include "g2o/types/slam3d/dquat2mat.h"

extern "C" {

// rho[3] = (rho, rho', rho'') of the kernel `name` ("Huber", "Cauchy", ... as registered by G2O_REGISTER_ROBUST_KERNEL) at squared error e2;
// returns 0, or -1 when the reference's factory does not know the name
int ref_robustify(const char* name, double delta, double e2, double* rho) {
  g2o::AbstractRobustKernelCreator* creator = g2o::RobustKernelFactory::instance()->creator(name);
  if (!creator) return -1;
  g2o::RobustKernel* k = creator->construct();
  k->setDelta(delta);
  g2o::Vector3 r;
  k->robustify(e2, r);
  rho[0] = r[0]; rho[1] = r[1]; rho[2] = r[2];
  delete k;
  return 0;
}

// R9: rotation matrix, column-major; out27: the 3 x 9 matrix dq_dR, column-major
void ref_dq_dR(const double* R9, double* out27) {
  Eigen::Matrix<number_t, 3, 9, Eigen::ColMajor> D;
  g2o::internal::compute_dq_dR(D, R9[0], R9[1], R9[2], R9[3], R9[4], R9[5], R9[6], R9[7], R9[8]);
  std::memcpy(out27, D.data(), sizeof(double) * 27);
}

double ref_normalize_theta(double theta) { return g2o::normalize_theta(theta); }

// out[i] = sampleGaussian(&engine[which[i]]) for two default-seeded std::mt19937 engines, as the two GaussianSampler objects of
// create_sphere.cpp:117-132 hold them (GaussianSampler() : _generator(new std::mt19937), stuff/sampler.h:47-56).  The static
// distribution inside sampler.cpp keeps its saved value across calls and across engines; call once per process for a clean sequence.
void ref_sample_gaussian_two_engines(const int* which, int n, double* out) {
  std::mt19937 engine[2];
  for (int i = 0; i < n; ++i) out[i] = g2o::sampleGaussian(&engine[which[i] & 1]);
}

}  // extern "C"
