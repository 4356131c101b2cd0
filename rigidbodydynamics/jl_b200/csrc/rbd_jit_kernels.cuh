// Persistent kernels around the generated per-sample functions (see rbd_jit_prelude.cuh).  Same structure as the generic pair
// in rbd_b200.cu: single-warp shared-memory blocks plus one Tensor-Memory CTA per SM, all claiming groups of 32 consecutive
// samples from one atomic counter; the next group's input lines are prefetched into L2 while the current one is computed.
// Warps never synchronise with each other.  (Tried and dropped: multi-warp CTAs walking the program as a convoy behind CTA
// barriers every 64 ... 1024 statements, to share instruction fetches -- no measurable change)
#pragma once

#define RBD_QUEUE_LOOP(CALL)                                                                        \
  const int lane = threadIdx.x & 31;                                                                \
  const long long ngroups = (a.B + 31) / 32;                                                        \
  long long g = rbd_claim_group(a.counter);                                                         \
  while (g < ngroups) {                                                                             \
    const long long gn = rbd_claim_group(a.counter);                                                \
    if (gn < ngroups) {                                                                             \
      const long long bn = gn * 32;                                                                 \
      rbd_prefetch_rows(a.q, RBD_SPEC_NQ, a.ld, bn);                                                \
      rbd_prefetch_rows(a.v, RBD_SPEC_NV, a.ld, bn);                                                \
      if (RBD_SPEC_HAS_IN2) rbd_prefetch_rows(a.in2, RBD_SPEC_NV, a.ld, bn);                        \
    }                                                                                               \
    const long long b = g * 32 + lane;                                                              \
    const bool active = b < a.B;                                                                    \
    const long long bl = active ? b : a.B - 1;   /* inactive lanes recompute the last sample, stores are masked */ \
    CALL(a.q + bl, a.v + bl, RBD_SPEC_HAS_IN2 ? a.in2 + bl : (const rbd_f*)0, a.o0 + bl,             \
         RBD_SPEC_HAS_OUT1 ? a.o1 + bl : (rbd_f*)0, a.ld, active, a.flag);                           \
    g = gn;                                                                                         \
  }

extern "C" __global__ void __launch_bounds__(32, RBD_SPEC_F64 ? 1 : 16) rbd_jit_smem(const RbdJitArgs a) {
  extern __shared__ __align__(16) unsigned char rbd_smem_raw[];
  volatile rbd_f* sh = reinterpret_cast<volatile rbd_f*>(rbd_smem_raw) + threadIdx.x;
#define RBD_CALL_SMEM(q_, v_, i_, o0_, o1_, ld_, act_, fl_) rbd_spec_smem(q_, v_, i_, o0_, o1_, ld_, act_, fl_, sh)
  RBD_QUEUE_LOOP(RBD_CALL_SMEM)
}

// CTA of NW = 8 (fp32) / 4 (fp64) warps over all 512 TMEM columns: warp w uses lane quadrant w % 4 and, in the 8-warp CTA,
// column half w / 4 (256 fp32 rows per warp; fp64 rows take two columns).
#define RBD_TM_WARPS (RBD_SPEC_F64 ? 4 : 8)
extern "C" __global__ void __launch_bounds__(32 * RBD_TM_WARPS, RBD_SPEC_F64 ? 1 : 2) rbd_jit_tmem(const RbdJitArgs a) {
  __shared__ unsigned tm_slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&tm_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm_base = tm_slot;
  const unsigned w = threadIdx.x >> 5;
  const unsigned tm = tm_base + (((w & 3u) * 32u) << 16) + (w >> 2) * 256u;
#define RBD_CALL_TMEM(q_, v_, i_, o0_, o1_, ld_, act_, fl_) rbd_spec_tmem(q_, v_, i_, o0_, o1_, ld_, act_, fl_, tm)
  {
    RBD_QUEUE_LOOP(RBD_CALL_TMEM)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
}
