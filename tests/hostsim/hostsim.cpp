// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Runs the per-sample device algorithms of csrc/rbd_device.cuh ON THE CPU (they are __host__ __device__ templates)
// so that the math of the kernels can be checked against the oracle in the CPU-only test tier, before any GPU time
// is spent.  The shipped librbd_b200.so does not contain this file and has no CPU path.
#include <string>
#include <vector>

#include "../../rigidbodydynamics/jl_b200/csrc/rbd_device.cuh"
#include "../../rigidbodydynamics/jl_b200/csrc/rbd_model.h"

using namespace rbd;

namespace {
template <class T> const ModelDev<T>& dev(const HostModel& m);
template <> const ModelDev<float>& dev<float>(const HostModel& m) { return m.dev32; }
template <> const ModelDev<double>& dev<double>(const HostModel& m) { return m.dev64; }

template <class T>
void run_dynamics(const HostModel& hm, int64_t B, const T* q, const T* v, const T* tau, const T* wext, T* vd, T* qd) {
  const ModelDev<T>& M = dev<T>(hm);
  std::vector<T> stash(M.nrows + 64);
  for (int64_t b = 0; b < B; ++b) {
    AbaIO<T> io;
    io.q = {q + b, B}; io.v = {v + b, B};
    io.tau = {tau ? tau + b : nullptr, B}; io.wext = {wext ? wext + b : nullptr, B};
    io.vd = {vd + b, B, true}; io.qd = {qd ? qd + b : nullptr, B, true};
    Stash<T, 1> st{stash.data()};
    if (hm.general) aba_sample<T, 1, true>(M, io, st);
    else aba_sample<T, 1, false>(M, io, st);
  }
}
}  // namespace

extern "C" {
int hostsim_info(const rbd_model_desc* d, int* nrows, int* general, int* nslots, int* order) {
  HostModel hm; std::string err;
  int rc = build_host_model(d, hm, err);
  if (rc) return rc;
  *nrows = hm.dev64.nrows; *general = hm.general; *nslots = hm.nslots;
  for (int i = 0; i < hm.nb; ++i) order[i] = hm.order[i];
  return 0;
}
int hostsim_dynamics(const rbd_model_desc* d, int dtype, int64_t B, const void* q, const void* v, const void* tau,
                     const void* wext, void* vd, void* qd) {
  HostModel hm; std::string err;
  int rc = build_host_model(d, hm, err);
  if (rc) return rc;
  if (dtype == 0) run_dynamics<float>(hm, B, (const float*)q, (const float*)v, (const float*)tau, (const float*)wext, (float*)vd, (float*)qd);
  else run_dynamics<double>(hm, B, (const double*)q, (const double*)v, (const double*)tau, (const double*)wext, (double*)vd, (double*)qd);
  return 0;
}
}
