// librbd_b200.so -- C ABI (include/rbd_b200.h) over the sm_100a kernels.
//
// Kernel design (DESIGN.md has the long version):
//   * one THREAD owns one sample; a block is one or more independent warps; there are no block-level barriers;
//   * the grid is persistent: blocks_per_SM x SM_count blocks loop over groups of NT consecutive samples;
//   * inputs / outputs are rows x batch with the batch index fastest, so lane l of a warp touches element b0 + l of a
//     row: every global access of a warp is one fully-used 128-byte line (fp32) / two lines (fp64);
//   * the per-sample working set that must survive between the three passes lives in shared memory, laid out
//     [row][lane] so a warp's access to a row hits 32 distinct banks;
//   * the flattened mechanism is a __grid_constant__ kernel parameter: it is read through the constant bank with a
//     warp-uniform index, i.e. as uniform-register operands, not as per-thread loads.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../../include/rbd_b200.h"
#include "rbd_device.cuh"
#include "rbd_model.h"

using namespace rbd;

// ------------------------------------------------------------------------------------------------------------------
// handle, errors
// ------------------------------------------------------------------------------------------------------------------
struct rbd_model {
  HostModel hm;
  // staging for the *_host entry points (allocated on first use, owned by the handle)
  std::mutex host_mu;
  void* d_stage[3] = {nullptr, nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

thread_local std::string g_err;
thread_local rbd_launch_info g_launch = {0, 0, 0, 0, 0, 0.f};

int fail(int status, const std::string& msg) { g_err = msg; return status; }
int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return RBD_ECUDA;
}
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail_cuda(e_, #expr); } while (0)

struct DeviceProps { int sms = 0; int max_smem_optin = 0; int smem_per_sm = 0; bool ok = false; };
int get_props(DeviceProps& p) {
  static std::mutex mu;
  static DeviceProps cache[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && cache[dev].ok) { p = cache[dev]; return RBD_OK; }
  DeviceProps q;
  CUDA_TRY(cudaDeviceGetAttribute(&q.sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&q.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&q.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  q.ok = true;
  if (dev >= 0 && dev < 64) cache[dev] = q;
  p = q;
  return RBD_OK;
}

template <class T> const ModelDev<T>& dev_model(const HostModel& m);
template <> const ModelDev<float>& dev_model<float>(const HostModel& m) { return m.dev32; }
template <> const ModelDev<double>& dev_model<double>(const HostModel& m) { return m.dev64; }

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
template <class T> struct AbaArgs {
  const T* q; const T* v; const T* tau; const T* wext;
  T* vd; T* qd;
  int64_t ld, B;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class T, int NT, bool GENERAL>
__global__ void __launch_bounds__(NT) aba_kernel(const __grid_constant__ ModelDev<T> M, const AbaArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sh = reinterpret_cast<T*>(smem_raw);
  const Stash<T, NT> st{sh + threadIdx.x};
  const int64_t ngroups = (a.B + NT - 1) / NT;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    {  // pull the NEXT group's input lines into L2 while this group is being computed (one 128-byte line per row)
      const int64_t gn = g + gridDim.x;
      if (gn < ngroups) {
        const int64_t bn = gn * NT + (threadIdx.x & ~31);
        const int lane = threadIdx.x & 31;
        for (int r = lane; r < M.nq; r += 32) prefetch_l2(a.q + (int64_t)r * a.ld + bn);
        for (int r = lane; r < M.nv; r += 32) prefetch_l2(a.v + (int64_t)r * a.ld + bn);
        if (a.tau) for (int r = lane; r < M.nv; r += 32) prefetch_l2(a.tau + (int64_t)r * a.ld + bn);
      }
    }
    const int64_t b = g * NT + threadIdx.x;
    const bool active = b < a.B;
    const int64_t bl = active ? b : a.B - 1;     // inactive lanes recompute the last sample, stores are masked
    AbaIO<T> io;
    io.q = {a.q + bl, a.ld};
    io.v = {a.v + bl, a.ld};
    io.tau = {a.tau ? a.tau + bl : nullptr, a.ld};
    io.wext = {a.wext ? a.wext + bl : nullptr, a.ld};
    io.vd = {a.vd + bl, a.ld, active};
    io.qd = {a.qd ? a.qd + bl : nullptr, a.ld, active};
    aba_sample<T, NT, GENERAL>(M, io, st);
  }
}

template <class K> int configure(K kernel, int nt, size_t smem, const DeviceProps& p, int& blocks_per_sm) {
  if ((int)smem > p.max_smem_optin) return fail(RBD_EUNSUPPORTED, "model working set exceeds shared memory per block");
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, nt, smem));
  if (blocks_per_sm < 1) return fail(RBD_EUNSUPPORTED, "kernel does not fit on an SM");
  return RBD_OK;
}

template <class T, bool GENERAL>
int launch_aba(const HostModel& hm, const AbaArgs<T>& a, cudaStream_t stream) {
  constexpr int NT = 32;
  DeviceProps p;
  if (int rc = get_props(p)) return rc;
  const ModelDev<T>& M = dev_model<T>(hm);
  const size_t smem = (size_t)M.nrows * NT * sizeof(T);
  auto kernel = aba_kernel<T, NT, GENERAL>;
  int bps = 0;
  if (int rc = configure(kernel, NT, smem, p, bps)) return rc;
  const int64_t ngroups = (a.B + NT - 1) / NT;
  const int grid = (int)std::min<int64_t>(ngroups, (int64_t)bps * p.sms);
  kernel<<<grid, NT, smem, stream>>>(M, a);
  CUDA_TRY(cudaGetLastError());
  g_launch.kernels_launched += 1;
  g_launch.grid = grid; g_launch.block = NT; g_launch.smem_bytes = (int)smem; g_launch.blocks_per_sm = bps;
  return RBD_OK;
}

template <class T>
int dynamics_t(const rbd_model* model, int64_t B, int64_t ld, const void* q, const void* v, const void* tau,
               const void* wext, void* vd, void* qd, cudaStream_t stream) {
  AbaArgs<T> a{(const T*)q, (const T*)v, (const T*)tau, (const T*)wext, (T*)vd, (T*)qd, ld, B};
  if (wext) return fail(RBD_EUNSUPPORTED, "external wrenches: not implemented yet");
  return model->hm.general ? launch_aba<T, true>(model->hm, a, stream) : launch_aba<T, false>(model->hm, a, stream);
}

int check_common(const rbd_model* model, int32_t dtype, int64_t B, int64_t ld) {
  if (!model) return fail(RBD_EINVAL, "model handle is NULL");
  if (dtype != RBD_F32 && dtype != RBD_F64) return fail(RBD_EINVAL, "dtype must be RBD_F32 or RBD_F64");
  if (B < 0 || ld < B) return fail(RBD_EDIM, "batch size / leading dimension mismatch (need ld >= B >= 0)");
  return RBD_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t rbd_version(void) { return RBD_B200_VERSION; }
const char* rbd_last_error(void) { return g_err.c_str(); }
const char* rbd_status_string(int32_t s) {
  switch (s) {
    case RBD_OK: return "RBD_OK";
    case RBD_EINVAL: return "RBD_EINVAL";
    case RBD_EDIM: return "RBD_EDIM";
    case RBD_ELOOP: return "RBD_ELOOP";
    case RBD_ESTALE: return "RBD_ESTALE";
    case RBD_ECUDA: return "RBD_ECUDA";
    case RBD_EUNSUPPORTED: return "RBD_EUNSUPPORTED";
    case RBD_ENOMEM: return "RBD_ENOMEM";
  }
  return "RBD_?";
}

int32_t rbd_model_create(const rbd_model_desc* desc, rbd_model** out) {
  if (!out) return fail(RBD_EINVAL, "rbd_model_create: out is NULL");
  *out = nullptr;
  rbd_model* m = new (std::nothrow) rbd_model();
  if (!m) return fail(RBD_ENOMEM, "out of memory");
  std::string err;
  int rc = build_host_model(desc, m->hm, err);
  if (rc != RBD_OK) { delete m; return fail(rc, err); }
  *out = m;
  return RBD_OK;
}

int32_t rbd_model_destroy(rbd_model* m) {
  if (!m) return RBD_OK;
  for (int i = 0; i < 3; ++i) {
    if (m->d_stage[i]) cudaFree(m->d_stage[i]);
    if (m->streams[i]) cudaStreamDestroy(m->streams[i]);
  }
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  delete m;
  return RBD_OK;
}

int32_t rbd_model_get_info(const rbd_model* m, rbd_model_info* info) {
  if (!m || !info) return fail(RBD_EINVAL, "rbd_model_get_info: NULL argument");
  std::memset(info, 0, sizeof(*info));
  info->nb = m->hm.nb; info->nq = m->hm.nq; info->nv = m->hm.nv;
  info->stash_rows = m->hm.dev64.nrows;
  info->max_branch_depth = m->hm.nslots;
  info->general_path = m->hm.general ? 1 : 0;
  info->modcount = m->hm.modcount;
  for (int i = 0; i < m->hm.nb; ++i) {
    info->qstart[i] = m->hm.qstart[i];
    info->vstart[i] = m->hm.vstart[i];
    info->eval_order[i] = m->hm.order[i];
  }
  return RBD_OK;
}

int32_t rbd_model_check_modcount(const rbd_model* m, int64_t modcount) {
  if (!m) return fail(RBD_EINVAL, "model handle is NULL");
  if (m->hm.modcount != modcount)
    return fail(RBD_ESTALE, "ModificationCountMismatch: the Mechanism was modified after the model handle was created");
  return RBD_OK;
}

int32_t rbd_get_launch_info(rbd_launch_info* info) {
  if (!info) return fail(RBD_EINVAL, "rbd_get_launch_info: NULL argument");
  *info = g_launch;
  return RBD_OK;
}

int32_t rbd_dynamics(const rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q, const void* v,
                     const void* tau, const void* wext, void* vd_out, void* qd_out, void* stream) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  if (!q || !v || !vd_out) return fail(RBD_EINVAL, "rbd_dynamics: q, v and vd_out must not be NULL");
  g_launch = {0, 0, 0, 0, 0, 0.f};
  if (B == 0) return RBD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == RBD_F32 ? dynamics_t<float>(model, B, ld, q, v, tau, wext, vd_out, qd_out, s)
                          : dynamics_t<double>(model, B, ld, q, v, tau, wext, vd_out, qd_out, s);
}

int32_t rbd_inverse_dynamics(const rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q,
                             const void* v, const void* vd, const void* wext, void* tau_out, void* stream) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)v; (void)vd; (void)wext; (void)tau_out; (void)stream;
  return fail(RBD_EUNSUPPORTED, "rbd_inverse_dynamics: not implemented yet");
}

int32_t rbd_dynamics_bias(const rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q,
                          const void* v, const void* wext, void* c_out, void* stream) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)v; (void)wext; (void)c_out; (void)stream;
  return fail(RBD_EUNSUPPORTED, "rbd_dynamics_bias: not implemented yet");
}

int32_t rbd_mass_matrix(const rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q, void* M_out,
                        void* stream) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)M_out; (void)stream;
  return fail(RBD_EUNSUPPORTED, "rbd_mass_matrix: not implemented yet");
}

// ---- host-pointer variants: chunked H2D -> kernel -> D2H pipeline over three internal streams -----------------------
namespace {
constexpr int64_t kChunk = 1 << 16;

int ensure_staging(rbd_model* m, size_t bytes_per_stream) {
  for (int i = 0; i < 3; ++i)
    if (!m->streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking));
  if (!m->ev0) { CUDA_TRY(cudaEventCreate(&m->ev0)); CUDA_TRY(cudaEventCreate(&m->ev1)); }
  if (m->stage_bytes >= bytes_per_stream) return RBD_OK;
  for (int i = 0; i < 3; ++i) {
    if (m->d_stage[i]) { cudaFree(m->d_stage[i]); m->d_stage[i] = nullptr; }
    CUDA_TRY(cudaMalloc(&m->d_stage[i], bytes_per_stream));
  }
  m->stage_bytes = bytes_per_stream;
  return RBD_OK;
}

// copy rows x C block between a host array with leading dimension ld and a dense device tile (leading dimension C)
int copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int rows, cudaMemcpyKind kind,
              cudaStream_t s) {
  CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, s));
  return RBD_OK;
}
}  // namespace

int32_t rbd_dynamics_host(rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q, const void* v,
                          const void* tau, const void* wext, void* vd_out, void* qd_out) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  if (!q || !v || !vd_out) return fail(RBD_EINVAL, "rbd_dynamics_host: q, v and vd_out must not be NULL");
  if (wext) return fail(RBD_EUNSUPPORTED, "external wrenches: not implemented yet");
  g_launch = {0, 0, 0, 0, 0, 0.f};
  if (B == 0) return RBD_OK;
  std::lock_guard<std::mutex> lk(model->host_mu);
  const HostModel& hm = model->hm;
  const size_t es = dtype == RBD_F32 ? 4 : 8;
  const int64_t C = std::min<int64_t>(kChunk, B);
  const int rows_in = hm.nq + hm.nv + (tau ? hm.nv : 0);
  const int rows_out = hm.nv + (qd_out ? hm.nq : 0);
  if (int rc = ensure_staging(model, (size_t)(rows_in + rows_out) * C * es)) return rc;
  int launches = 0;
  rbd_launch_info last = g_launch;
  int nchunk = 0;
  for (int64_t b0 = 0; b0 < B; b0 += C, ++nchunk) {
    const int64_t n = std::min<int64_t>(C, B - b0);
    const int si = nchunk % 3;
    cudaStream_t s = model->streams[si];
    char* base = (char*)model->d_stage[si];
    char* dq = base;
    char* dv = dq + (size_t)hm.nq * C * es;
    char* dtau = dv + (size_t)hm.nv * C * es;
    char* dvd = dtau + (size_t)(tau ? hm.nv : 0) * C * es;
    char* dqd = dvd + (size_t)hm.nv * C * es;
    const size_t off = (size_t)b0 * es;
    if (int rc = copy_rows(dq, C * es, (const char*)q + off, ld * es, n * es, hm.nq, cudaMemcpyHostToDevice, s)) return rc;
    if (int rc = copy_rows(dv, C * es, (const char*)v + off, ld * es, n * es, hm.nv, cudaMemcpyHostToDevice, s)) return rc;
    if (tau)
      if (int rc = copy_rows(dtau, C * es, (const char*)tau + off, ld * es, n * es, hm.nv, cudaMemcpyHostToDevice, s)) return rc;
    int rc = dtype == RBD_F32
                 ? dynamics_t<float>(model, n, C, dq, dv, tau ? dtau : nullptr, nullptr, dvd, qd_out ? dqd : nullptr, s)
                 : dynamics_t<double>(model, n, C, dq, dv, tau ? dtau : nullptr, nullptr, dvd, qd_out ? dqd : nullptr, s);
    if (rc) return rc;
    launches += 1;
    last = g_launch;
    if (int rc2 = copy_rows((char*)vd_out + off, ld * es, dvd, C * es, n * es, hm.nv, cudaMemcpyDeviceToHost, s)) return rc2;
    if (qd_out)
      if (int rc2 = copy_rows((char*)qd_out + off, ld * es, dqd, C * es, n * es, hm.nq, cudaMemcpyDeviceToHost, s)) return rc2;
  }
  for (int i = 0; i < 3; ++i) CUDA_TRY(cudaStreamSynchronize(model->streams[i]));
  g_launch = last;
  g_launch.kernels_launched = launches;
  return RBD_OK;
}

int32_t rbd_inverse_dynamics_host(rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q,
                                  const void* v, const void* vd, const void* wext, void* tau_out) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)v; (void)vd; (void)wext; (void)tau_out;
  return fail(RBD_EUNSUPPORTED, "rbd_inverse_dynamics_host: not implemented yet");
}
int32_t rbd_dynamics_bias_host(rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q,
                               const void* v, const void* wext, void* c_out) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)v; (void)wext; (void)c_out;
  return fail(RBD_EUNSUPPORTED, "rbd_dynamics_bias_host: not implemented yet");
}
int32_t rbd_mass_matrix_host(rbd_model* model, int32_t dtype, int64_t B, int64_t ld, const void* q, void* M_out) {
  if (int rc = check_common(model, dtype, B, ld)) return rc;
  (void)q; (void)M_out;
  return fail(RBD_EUNSUPPORTED, "rbd_mass_matrix_host: not implemented yet");
}

}  // extern "C"
