// Model-specialised kernels: compile (rbd_jit.cpp), load (cudaLibraryLoadData) and launch.  Host-only C++ on the CUDA runtime.
//
// The kernels come as a pair, exactly like the generic ones in rbd_b200.cu: single-warp shared-memory blocks on the caller's
// stream and one Tensor-Memory CTA per SM on the handle's side stream, both claiming groups of 32 samples from one counter.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rbd_handle.h"
#include "rbd_jit.h"

namespace rbd {
namespace {

struct Env {
  bool jit = true, no_tmem = false, packed = true;
  int64_t min_batch = 1 << 15;      // below this a missing cubin is not compiled on the fly (generic kernels serve the call)
  const char* only = nullptr;       // RBD_ONLY=smem|tmem: launch one kernel of the pair (profiling aid)
  int smem_blocks = 0;
  Env() {
    if (const char* e = getenv("RBD_JIT")) jit = e[0] != '0';
    no_tmem = getenv("RBD_NO_TMEM") != nullptr;
    if (const char* e = getenv("RBD_JIT_PACKED")) packed = e[0] != '0';
    if (const char* e = getenv("RBD_JIT_MIN_BATCH")) min_batch = atoll(e);
    only = getenv("RBD_ONLY");
    if (const char* e = getenv("RBD_SMEM_BLOCKS")) smem_blocks = atoi(e);
  }
};
const Env& env() { static const Env e; return e; }

struct Props { int sms = 0, max_smem_optin = 0; };
cudaError_t props(Props& p) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  static std::mutex mu;
  static Props cache[64];
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 64 && cache[dev].sms) { p = cache[dev]; return cudaSuccess; }
  if ((e = cudaDeviceGetAttribute(&p.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  if (dev < 64) cache[dev] = p;
  return cudaSuccess;
}

}  // namespace

cudaError_t pair_begin(rbd_model* m, cudaStream_t stream, PairCtx& ctx) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  unsigned slot;
  {
    std::lock_guard<std::mutex> lk(m->side_mu);
    if (m->side_device != dev) {       // first use (or the caller moved to another device): (re)create the per-device resources
      if (m->side_stream) {
        cudaStreamDestroy(m->side_stream);
        for (int i = 0; i < kEventRing; ++i) { cudaEventDestroy(m->fork_ev[i]); cudaEventDestroy(m->join_ev[i]); }
        cudaFree(m->counters);
        m->side_stream = nullptr; m->counters = nullptr;
      }
      if ((e = cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
      for (int i = 0; i < kEventRing; ++i) {
        if ((e = cudaEventCreateWithFlags(&m->fork_ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&m->join_ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
      if ((e = cudaMalloc((void**)&m->counters, kCounterRing * sizeof(unsigned long long))) != cudaSuccess) return e;
      m->side_device = dev;
    }
    slot = m->next_call++;
  }
  ctx.side = m->side_stream;
  ctx.fork = m->fork_ev[slot % kEventRing];
  ctx.join = m->join_ev[slot % kEventRing];
  ctx.counter = m->counters + (slot % kCounterRing);
  return cudaMemsetAsync(ctx.counter, 0, sizeof(unsigned long long), stream);
}

void spec_release(rbd_model* m) {
  std::lock_guard<std::mutex> lk(m->spec_mu);
  for (auto& kv : m->spec) if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  m->spec.clear();
}

// The arithmetic mode is the library's choice, not the caller's: fp32 programs run packed (two samples per thread) unless
// RBD_JIT_PACKED=0.
SpecKey spec_resolve(SpecKey key) {
  key.packed = !key.f64 && env().packed;
  return key;
}

// Scalar straight-line programs beyond this size are slower than the generic kernels (instruction supply; measured on Atlas:
// fp32 ABA 12.7 k nodes 1.03x, fp64 0.83x, RNEA 5.8 k nodes 1.46x, 7-DoF arm 2.2 k nodes 2.7x): estimated from the body count.
bool spec_worthwhile(const HostModel& hm, const SpecKey& key) {
  if (key.packed) return true;
  const int est = hm.nb * (key.algo == SPEC_ABA ? 410 : 190);
  return est <= 9000;
}

int spec_prepare(rbd_model* m, const SpecKey& key_in, bool load_on_device, std::string& err) {
  const SpecKey key = spec_resolve(key_in);
  std::lock_guard<std::mutex> lk(m->spec_mu);
  SpecEntry& se = m->spec[spec_key_bits(key)];
  if (se.state == 1) return RBD_OK;
  if (se.state == -1 && load_on_device) { err = se.why; return RBD_EUNSUPPORTED; }
  std::vector<char> cubin;
  bool cached = false;
  const SpecTuning tune = spec_default_tuning(m->hm, key);
  if (!jit_get_cubin(m->hm, key, tune, cubin, true, &cached, nullptr, err)) {
    se.state = -1; se.why = err;
    return RBD_EUNSUPPORTED;
  }
  if (!load_on_device) return RBD_OK;      // precompile only (CPU build step): the cubin is in the cache now
  cudaError_t e = cudaLibraryLoadData(&se.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&se.k_smem, se.lib, "rbd_jit_smem");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&se.k_tmem, se.lib, "rbd_jit_tmem");
  if (e == cudaSuccess && key.packed) e = cudaLibraryGetKernel(&se.k_smem32, se.lib, "rbd_jit_smem32");
  cudaFuncAttributes fa{};
  if (e == cudaSuccess) { e = cudaFuncGetAttributes(&fa, (const void*)se.k_smem); se.regs_smem = fa.numRegs; }
  if (e == cudaSuccess) { e = cudaFuncGetAttributes(&fa, (const void*)se.k_tmem); se.regs_tmem = fa.numRegs; }
  se.rows = spec_stash_rows(m->hm, key);
  se.smem_warps = tune.smem_warps;
  const size_t smem = (size_t)std::max(1, se.rows) * 32 * ((key.f64 || key.packed) ? 8 : 4) * se.smem_warps;
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && se.k_smem32) e = cudaFuncSetAttribute((const void*)se.k_smem32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && se.k_smem32) e = cudaFuncSetAttribute((const void*)se.k_smem32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_smem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_tmem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) {
    cudaGetLastError();
    se.state = -1;
    se.why = err = std::string("loading the specialised cubin: ") + cudaGetErrorString(e);
    if (se.lib) { cudaLibraryUnload(se.lib); se.lib = nullptr; }
    return RBD_EUNSUPPORTED;
  }
  se.from_cache = cached;
  se.state = 1;
  return RBD_OK;
}

int spec_try_launch(rbd_model* m, const SpecKey& key_in, const SpecLaunchArgs& a, cudaStream_t stream, bool& used,
                    rbd_launch_info& li, std::string& err) {
  used = false;
  const Env& ev = env();
  if (!ev.jit) return RBD_OK;
  const SpecKey key = spec_resolve(key_in);
  if (!spec_worthwhile(m->hm, key)) return RBD_OK;
  const int rows = spec_stash_rows(m->hm, key);
  if (rows > 256) return RBD_OK;                     // one warp's share of Tensor Memory: 256 fp32 rows (or 256 two-column rows with 4 warps)
  Props p;
  if (cudaError_t e = props(p)) { err = cudaGetErrorString(e); return RBD_ECUDA; }
  const size_t es = (key.f64 || key.packed) ? 8 : 4;       // bytes per stash row and lane
  if ((size_t)rows * 32 * es > (size_t)p.max_smem_optin - 64) return RBD_OK;
  SpecEntry* se = nullptr;
  {
    std::lock_guard<std::mutex> lk(m->spec_mu);
    se = &m->spec[spec_key_bits(key)];
  }
  if (se->state == -1) return RBD_OK;
  if (se->state == 0) {
    if (a.B < ev.min_batch) {          // small call and nothing loaded yet: use a cached cubin if there is one, never compile
      std::vector<char> probe;
      std::string e2;
      bool cached = false;
      if (!jit_get_cubin(m->hm, key, spec_default_tuning(m->hm, key), probe, false, &cached, nullptr, e2)) return RBD_OK;
    }
    std::string e2;
    if (spec_prepare(m, key, true, e2) != RBD_OK) return RBD_OK;      // generic kernels take over; the reason is kept in the entry
  }
  const int W = se->smem_warps;
  const int group = key.packed ? 64 : 32;
  const size_t smem = (size_t)std::max(1, rows) * 32 * es * W;
  const int tm_warps = (key.f64 || key.packed) ? 4 : 8;
  // packed mode reads / writes one 64-bit word per pair: needs 8-byte aligned pairs; an odd batch leaves one sample for the
  // 32-bit-I/O kernel
  const size_t fsz = key.f64 ? 8 : 4;
  bool aligned = true;
  if (key.packed) {
    const uintptr_t bits = (uintptr_t)a.q | (uintptr_t)a.v | (uintptr_t)a.in2 | (uintptr_t)a.o0 | (uintptr_t)a.o1;
    aligned = (bits & 7) == 0 && (a.ld & 1) == 0;
  }
  const int64_t Bmain = key.packed ? (aligned ? (a.B & ~(int64_t)1) : 0) : a.B;
  const int64_t Btail = a.B - Bmain;
  auto off = [&](const void* ptr, int64_t n) -> const void* { return ptr ? (const char*)ptr + (size_t)n * fsz : nullptr; };
  struct KArgs { const void* q; const void* v; const void* in2; void* o0; void* o1; long long ld, B; unsigned long long* counter; };
  int launched = 0;
  cudaError_t e = cudaSuccess;
  if (Bmain > 0) {
    const int64_t ngroups = (Bmain + group - 1) / group;
    int bps = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void*)se->k_smem, 32 * W, smem)) != cudaSuccess) {
      err = cudaGetErrorString(e); return RBD_ECUDA;
    }
    if (bps < 1) return RBD_OK;
    // Register-file room for the Tensor-Memory CTA next to the shared-memory CTAs.  The pair only pays when shared memory (not
    // the register file) limits the single kernel's residency, and when there is enough work for both kernels' warps.
    const int rs = ((se->regs_smem + 7) / 8) * 8 * 32 * W, rt = ((se->regs_tmem + 7) / 8) * 8 * 32 * tm_warps;
    int bps_pair = std::max(1, std::min(bps, (65536 - rt) / rs));
    bool pair = !ev.no_tmem && rt + rs <= 65536 && (bps_pair * W + tm_warps) * 100 >= bps * W * 115;
    if (ev.smem_blocks > 0) bps_pair = std::max(1, std::min(bps_pair, ev.smem_blocks));
    if (pair && ngroups < (int64_t)(bps_pair * W + tm_warps / 2) * p.sms) pair = false;
    PairCtx ctx;
    if ((e = pair_begin(m, stream, ctx)) != cudaSuccess) { err = cudaGetErrorString(e); return RBD_ECUDA; }
    KArgs ka = {a.q, a.v, a.in2, a.o0, a.o1, (long long)a.ld, (long long)Bmain, ctx.counter};
    void* params[] = {&ka};
    if (pair) {
      const bool do_s = !ev.only || ev.only[0] == 's', do_t = !ev.only || ev.only[0] == 't';
      e = cudaEventRecord(ctx.fork, stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx.side, ctx.fork, 0);
      if (e == cudaSuccess && do_s) { e = cudaLaunchKernel((const void*)se->k_smem, dim3(bps_pair * p.sms), dim3(32 * W), params, smem, stream); ++launched; }
      if (e == cudaSuccess && do_t) { e = cudaLaunchKernel((const void*)se->k_tmem, dim3(p.sms), dim3(32 * tm_warps), params, 0, ctx.side); ++launched; }
      if (e == cudaSuccess) e = cudaEventRecord(ctx.join, ctx.side);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ctx.join, 0);
      li.grid = bps_pair * p.sms; li.blocks_per_sm = bps_pair;
    } else {
      const int grid = (int)std::min<int64_t>((ngroups + W - 1) / W, (int64_t)bps * p.sms);
      e = cudaLaunchKernel((const void*)se->k_smem, dim3(grid), dim3(32 * W), params, smem, stream);
      ++launched;
      li.grid = grid; li.blocks_per_sm = bps;
    }
  }
  if (e == cudaSuccess && Btail > 0) {
    const int64_t ngroups = (Btail + group - 1) / group;
    int bps = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void*)se->k_smem32, 32 * W, smem);
    PairCtx ctx;
    if (e == cudaSuccess) e = pair_begin(m, stream, ctx);
    if (e == cudaSuccess) {
      KArgs ka = {off(a.q, Bmain), off(a.v, Bmain), off(a.in2, Bmain), (void*)off(a.o0, Bmain), (void*)off(a.o1, Bmain),
                  (long long)a.ld, (long long)Btail, ctx.counter};
      void* params[] = {&ka};
      const int grid = (int)std::min<int64_t>((ngroups + W - 1) / W, (int64_t)std::max(1, bps) * p.sms);
      e = cudaLaunchKernel((const void*)se->k_smem32, dim3(grid), dim3(32 * W), params, smem, stream);
      ++launched;
      if (Bmain == 0) { li.grid = grid; li.blocks_per_sm = bps; }
    }
  }
  if (e != cudaSuccess) { err = std::string("specialised kernel launch: ") + cudaGetErrorString(e); return RBD_ECUDA; }
  li.kernels_launched += launched;
  li.block = 32 * W; li.smem_bytes = (int)smem;
  used = true;
  return RBD_OK;
}

}  // namespace rbd
