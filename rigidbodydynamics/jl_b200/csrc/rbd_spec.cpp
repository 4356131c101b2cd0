// Model-specialised kernels: compile (rbd_jit.cpp), load (cudaLibraryLoadData) and launch.  Host-only C++ on the CUDA runtime.
//
// The kernels come as a pair, exactly like the generic ones in rbd_b200.cu: single-warp shared-memory blocks on the caller's
// stream and one Tensor-Memory CTA per SM on the handle's side stream, both claiming groups of 32 samples from one counter.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rbd_handle.h"
#include "rbd_jit.h"

namespace rbd {
namespace {

struct Env {
  bool jit = true, no_tmem = false, force_pair = false, no_gate = false;
  int64_t min_batch = 1 << 15;      // below this a missing cubin is not compiled on the fly (generic kernels serve the call)
  int smem_blocks = 0;
  int variant = 0;                  // RBD_JIT_VARIANT=1 (pair) / 2 (unified): skip the tuning
  const char* only = nullptr;       // RBD_ONLY=smem|tmem: launch one kernel of the pair (ncu captures: the profiler serialises them)
  Env() {
    if (const char* e = getenv("RBD_JIT")) jit = e[0] != '0';
    no_tmem = getenv("RBD_NO_TMEM") != nullptr;
    force_pair = getenv("RBD_FORCE_PAIR") != nullptr;       // experiments only
    no_gate = getenv("RBD_NO_GATE") != nullptr;
    if (const char* e = getenv("RBD_JIT_MIN_BATCH")) min_batch = atoll(e);
    if (const char* e = getenv("RBD_SMEM_BLOCKS")) smem_blocks = atoi(e);
    if (const char* e = getenv("RBD_JIT_VARIANT")) variant = atoi(e);
    only = getenv("RBD_ONLY");
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
      if ((e = cudaMalloc((void**)&m->counters, kCounterRing * 2 * sizeof(unsigned long long))) != cudaSuccess) return e;
      m->side_device = dev;
    }
    slot = m->next_call++;
  }
  ctx.side = m->side_stream;
  ctx.fork = m->fork_ev[slot % kEventRing];
  ctx.join = m->join_ev[slot % kEventRing];
  ctx.counter = m->counters + 2 * (slot % kCounterRing);
  ctx.flag = reinterpret_cast<int*>(ctx.counter + 1);
  return cudaMemsetAsync(ctx.counter, 0, 2 * sizeof(unsigned long long), stream);
}

void spec_release(rbd_model* m) {
  std::lock_guard<std::mutex> lk(m->spec_mu);
  for (auto& kv : m->spec) if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  m->spec.clear();
}

// Straight-line programs far beyond the instruction caches run at the SM's instruction-fetch rate (~1.4 instr/clk on B200)
// instead of its issue rate; measured against the generic (looping, cache-resident) kernels on Atlas: fp32 ABA 12.7 k nodes
// 1.03x, fp32 RNEA 5.8 k nodes 1.46x, 7-DoF arm 2.2 k nodes 2.7x, but fp64 ABA 0.83x (half the resident warps).  fp32 programs
// are therefore always specialised, fp64 programs only below an estimated size.
bool spec_worthwhile(const HostModel& hm, const SpecKey& key) {
  if (!key.f64) return true;
  if (key.algo == SPEC_CRBA) return !getenv("RBD_JIT_NO_CRBA64");   // small stash (2 rows per body): fp64 keeps its resident warps
  if (key.algo == SPEC_KIN) return true;                             // stash = the pending branch nodes only
  const int est = hm.nb * (key.algo == SPEC_ABA ? 410 : 190);
  return est <= 6500;
}

int spec_prepare(rbd_model* m, const SpecKey& key, bool load_on_device, std::string& err) {
  std::lock_guard<std::mutex> lk(m->spec_mu);
  SpecEntry& se = m->spec[spec_key_bits(key)];
  if (se.state == 1) return RBD_OK;
  if (se.state == -1 && load_on_device) { err = se.why; return RBD_EUNSUPPORTED; }
  std::vector<char> cubin;
  bool cached = false;
  if (!jit_get_cubin(m->hm, key, cubin, true, &cached, nullptr, err)) {
    se.state = -1; se.why = err;
    return RBD_EUNSUPPORTED;
  }
  if (!load_on_device) return RBD_OK;      // precompile only (CPU build step): the cubin is in the cache now
  cudaError_t e = cudaLibraryLoadData(&se.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&se.k_smem, se.lib, "rbd_jit_smem");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&se.k_tmem, se.lib, "rbd_jit_tmem");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&se.k_uni, se.lib, "rbd_jit_uni");
  cudaFuncAttributes fa{};
  if (e == cudaSuccess) { e = cudaFuncGetAttributes(&fa, (const void*)se.k_smem); se.regs_smem = fa.numRegs; }
  if (e == cudaSuccess) { e = cudaFuncGetAttributes(&fa, (const void*)se.k_tmem); se.regs_tmem = fa.numRegs; }
  if (e == cudaSuccess) { e = cudaFuncGetAttributes(&fa, (const void*)se.k_uni); se.regs_uni = fa.numRegs; }
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_tmem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  se.rows = spec_stash_rows(m->hm, key);
  const size_t smem = (size_t)std::max(1, se.rows) * 32 * (key.f64 ? 8 : 4);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_smem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  se.uni_sw = spec_uni_smem_warps(m->hm, key);
  if (e == cudaSuccess && se.uni_sw > 0) e = cudaFuncSetAttribute((const void*)se.k_uni, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * se.uni_sw));
  if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)se.k_uni, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
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

int spec_try_launch(rbd_model* m, const SpecKey& key, const SpecLaunchArgs& a, cudaStream_t stream, bool& used,
                    rbd_launch_info& li, const int** gate, std::string& err) {
  used = false;
  if (gate) *gate = nullptr;
  const Env& ev = env();
  if (!ev.jit || !spec_worthwhile(m->hm, key)) return RBD_OK;
  if (a.ld * (key.f64 ? 8 : 4) >= (1ll << 32)) return RBD_OK;   // the generated code forms row offsets as 32 x 32 -> 64-bit products of the byte stride
  const int rows = spec_stash_rows(m->hm, key);
  if (rows > 256) return RBD_OK;                     // one warp's share of Tensor Memory: 256 fp32 rows (128 x 2 columns in fp64)
  Props p;
  if (cudaError_t e = props(p)) { err = cudaGetErrorString(e); return RBD_ECUDA; }
  const size_t es = key.f64 ? 8 : 4;
  const size_t smem = (size_t)std::max(1, rows) * 32 * es;
  if (smem > (size_t)p.max_smem_optin) return RBD_OK;
  SpecEntry* se = nullptr;
  {
    std::lock_guard<std::mutex> lk(m->spec_mu);
    se = &m->spec[spec_key_bits(key)];
  }
  if (key.algo == SPEC_KIN) {        // the 64-bit map key carries only a hash of the jacobian path: use the entry only for ITS path
    std::lock_guard<std::mutex> lk(m->spec_mu);
    if (!se->kin_set) { std::memcpy(se->kin_sign, key.kin_sign, sizeof se->kin_sign); se->kin_set = true; }
    else if (std::memcmp(se->kin_sign, key.kin_sign, sizeof se->kin_sign) != 0) return RBD_OK;
  }
  if (se->state == -1) return RBD_OK;
  if (se->state == 0) {
    if (a.B < ev.min_batch) {          // small call and nothing loaded yet: use a cached cubin if there is one, never compile
      std::vector<char> probe;
      std::string e2;
      bool cached = false;
      if (!jit_get_cubin(m->hm, key, probe, false, &cached, nullptr, e2)) return RBD_OK;
    }
    std::string e2;
    if (spec_prepare(m, key, true, e2) != RBD_OK) {      // generic kernels take over; the reason is kept in the entry
      if (getenv("RBD_JIT_VERBOSE")) fprintf(stderr, "[rbd_b200] specialised kernels unavailable: %s\n", e2.c_str());
      return RBD_OK;
    }
  }
  const int64_t ngroups = (a.B + 31) / 32;
  const int tm_warps = key.f64 ? 4 : 8;
  int bps = 0;
  if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void*)se->k_smem, 32, smem)) {
    err = cudaGetErrorString(e); return RBD_ECUDA;
  }
  if (bps < 1) return RBD_OK;
  // Three ways to fill an SM (rbd_jit_kernels.cuh): shared-memory blocks alone, the kernel pair, the unified CTA.  The two
  // Tensor-Memory variants are used when they put at least 15 % more warps on the SM than shared memory alone and there is work
  // for them; which of the two is decided by timing both once, on the first call large enough to tell (not while capturing).
  if (ev.smem_blocks > 0) bps = std::max(1, std::min(bps, ev.smem_blocks));
  const int rs = ((se->regs_smem + 7) / 8) * 8 * 32, rt = ((se->regs_tmem + 7) / 8) * 8 * 32 * tm_warps;
  const int bps_pair = std::max(1, std::min(bps, (65536 - rt) / rs));
  const bool pair_ok = !ev.no_tmem && rt + rs <= 65536 && (bps_pair + tm_warps) * 100 >= bps * 115 &&
                       ngroups >= (int64_t)(bps_pair + tm_warps / 2) * p.sms;
  const int uni_warps = se->uni_sw > 0 ? se->uni_sw + tm_warps : 0;
  const bool uni_ok = !ev.no_tmem && uni_warps * 100 >= bps * 115 && ((se->regs_uni + 7) / 8) * 8 * 32 * uni_warps <= 65536 &&
                      ngroups >= (int64_t)(uni_warps - tm_warps / 2) * p.sms;
  struct KArgs {
    const void* q; const void* v; const void* in2; void* o0; void* o1; long long ld, B; unsigned long long* counter; int* flag;
    void* peers[8]; long long peer_ld, peer_col0; void* mc; int npeers;
    void* ko[8];
  };
  int launched = 0;
  int* last_flag = nullptr;
  auto run = [&](int variant) -> cudaError_t {        // 0 = shared memory alone, 1 = pair, 2 = unified
    PairCtx ctx;
    cudaError_t e = pair_begin(m, stream, ctx);
    if (e != cudaSuccess) return e;
    KArgs ka = {a.q, a.v, a.in2, a.o0, a.o1, (long long)a.ld, (long long)a.B, ctx.counter, ctx.flag, {}, (long long)a.peer_ld,
                (long long)a.peer_col0, a.mc, a.npeers};
    for (int i = 0; i < a.npeers && i < 8; ++i) ka.peers[i] = a.peers[i];
    for (int i = 0; i < 8; ++i) ka.ko[i] = a.ko[i];
    void* params[] = {&ka};
    if (variant == 2) {
      e = cudaLaunchKernel((const void*)se->k_uni, dim3(p.sms), dim3(32 * uni_warps), params, smem * se->uni_sw, stream);
      ++launched;
      li.grid = p.sms; li.block = 32 * uni_warps; li.blocks_per_sm = 1; li.smem_bytes = (int)(smem * se->uni_sw);
    } else if (variant == 1) {
      e = cudaEventRecord(ctx.fork, stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx.side, ctx.fork, 0);
      if (e == cudaSuccess && (!ev.only || ev.only[0] == 's')) { e = cudaLaunchKernel((const void*)se->k_smem, dim3(bps_pair * p.sms), dim3(32), params, smem, stream); ++launched; }
      if (e == cudaSuccess && (!ev.only || ev.only[0] == 't')) { e = cudaLaunchKernel((const void*)se->k_tmem, dim3(p.sms), dim3(32 * tm_warps), params, 0, ctx.side); ++launched; }
      if (e == cudaSuccess) e = cudaEventRecord(ctx.join, ctx.side);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ctx.join, 0);
      li.grid = bps_pair * p.sms; li.block = 32; li.blocks_per_sm = bps_pair; li.smem_bytes = (int)smem;
    } else {
      const int grid = (int)std::min<int64_t>(ngroups, (int64_t)bps * p.sms);
      e = cudaLaunchKernel((const void*)se->k_smem, dim3(grid), dim3(32), params, smem, stream);
      ++launched;
      li.grid = grid; li.block = 32; li.blocks_per_sm = bps; li.smem_bytes = (int)smem;
    }
    last_flag = ctx.flag;
    return e;
  };
  cudaError_t e = cudaSuccess;
  int variant = 0;
  if (pair_ok && uni_ok) {
    variant = ev.variant ? ev.variant : se->choice;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (variant == 0 && cap == cudaStreamCaptureStatusNone && a.B >= (int64_t)p.sms * 32 * 64) {
      // tune: each variant once to warm up, then timed; outputs are simply overwritten with the same values
      float best = 0.f;
      cudaEvent_t t0, t1;
      cudaEventCreate(&t0); cudaEventCreate(&t1);
      for (int v = 1; v <= 2 && e == cudaSuccess; ++v) {
        e = run(v);
        if (e == cudaSuccess) e = cudaEventRecord(t0, stream);
        if (e == cudaSuccess) e = run(v);
        if (e == cudaSuccess) e = cudaEventRecord(t1, stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(t1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t0, t1);
        if (e == cudaSuccess && (variant == 0 || ms < best)) { best = ms; variant = v; }
      }
      cudaEventDestroy(t0); cudaEventDestroy(t1);
      if (e == cudaSuccess) se->choice = variant;
      launched = 0;
    }
    if (variant == 0) variant = 1;
  } else if (uni_ok) variant = 2;
  else if (pair_ok) variant = 1;
  if (ev.force_pair && uni_warps) variant = 2;
  if (e == cudaSuccess) e = run(variant);
  if (e != cudaSuccess) { err = std::string("specialised kernel launch: ") + cudaGetErrorString(e); return RBD_ECUDA; }
  li.kernels_launched += launched;
  if (gate && !key.f64 && !ev.no_gate) *gate = last_flag;
  used = true;
  return RBD_OK;
}

}  // namespace rbd
