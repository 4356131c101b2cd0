// Persistent kernels around the generated per-sample functions (see rbd_jit_prelude.cuh): warps claim groups of 32 consecutive
// samples from an atomic counter; the next group's input lines are prefetched into L2 while the current one is computed; warps
// never synchronise with each other.  Three ways to fill an SM with warps (rbd_spec.cpp picks one per model and entry point):
//   rbd_jit_smem alone       single-warp blocks, stash in shared memory -- models whose stash is small;
//   rbd_jit_smem + _tmem     the kernel pair of the generic path: shared-memory blocks and a Tensor-Memory CTA side by side;
//   rbd_jit_uni              one CTA per SM, half of its warps' stashes in shared memory, half in Tensor Memory, one program.
// Which of the last two is faster depends on the program (Atlas forward dynamics: pair 1.03 vs 0.93 G evals/s; Atlas inverse
// dynamics: 2.23 vs 2.72 G; 31-joint chain: 0.66 vs 0.81 G), so the first large call times both once and keeps the winner.  (Tried and dropped: multi-warp CTAs walking the program as a convoy behind CTA
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
      if (RBD_SPEC_USES_V) rbd_prefetch_rows(a.v, RBD_SPEC_NV, a.ld, bn);   /* not when the program never reads v */ \
      if (RBD_SPEC_HAS_IN2) rbd_prefetch_rows(a.in2, RBD_SPEC_NV, a.ld, bn);                        \
    }                                                                                               \
    const long long b = g * 32 + lane;                                                              \
    const bool active = b < a.B;                                                                    \
    const long long bl = active ? b : a.B - 1;   /* inactive lanes recompute the last sample, stores are masked */ \
    CALL(a.q + bl, a.v + bl, RBD_SPEC_HAS_IN2 ? a.in2 + bl : (const rbd_f*)0, a.o0 + bl,             \
         RBD_SPEC_HAS_OUT1 ? a.o1 + bl : (rbd_f*)0, a.ld, active, a.flag, a, bl);                    \
    g = gn;                                                                                         \
  }

extern "C" __global__ void __launch_bounds__(32, RBD_SPEC_F64 ? 1 : 16) rbd_jit_smem(const RbdJitArgs a) {
  extern __shared__ __align__(16) unsigned char rbd_smem_raw[];
  volatile rbd_f* sh = reinterpret_cast<volatile rbd_f*>(rbd_smem_raw) + threadIdx.x;
#define RBD_CALL_SMEM(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_) rbd_spec_smem(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_, sh)
  RBD_QUEUE_LOOP(RBD_CALL_SMEM)
}

// Tensor-Memory CTA of the kernel PAIR (rbd_jit_smem blocks + one of these per SM, two launches on two streams sharing the
// work queue): 8 (fp32) / 4 (fp64) warps over all 512 TMEM columns -- warp w uses lane quadrant w % 4 and, with 8 warps, column
// half w / 4 (256 fp32 rows per warp; fp64 rows take two columns).
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
#define RBD_CALL_TMEM(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_) rbd_spec_tmem(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_, tm)
  {
    RBD_QUEUE_LOOP(RBD_CALL_TMEM)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
}

// Unified CTA, one per SM: RBD_UNI_SW warps with their stash in shared memory ([row][lane] slices of the dynamic shared memory)
// followed by RBD_TM_WARPS warps with theirs in Tensor Memory (same layout as above).  All warps run the SAME per-sample
// function (flavour UNI): one instruction stream per SM instead of two.
#ifndef RBD_UNI_SW
#define RBD_UNI_SW RBD_TM_WARPS
#endif
extern "C" __global__ void __launch_bounds__(32 * (RBD_UNI_SW + RBD_TM_WARPS), 1) rbd_jit_uni(const RbdJitArgs a) {
  extern __shared__ __align__(16) unsigned char rbd_smem_raw[];
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
  const bool use_tm = w >= RBD_UNI_SW;
  const unsigned wt = use_tm ? w - RBD_UNI_SW : 0u;            // RBD_UNI_SW is a multiple of 4, so wt % 4 == w % 4 = this warp's lane quadrant
  const unsigned tm = tm_base + (((w & 3u) * 32u) << 16) + (wt >> 2) * 256u;
  const unsigned ws = use_tm ? 0u : w;
  volatile rbd_f* sh = reinterpret_cast<volatile rbd_f*>(rbd_smem_raw) + ws * (RBD_SPEC_ROWS * 32u) + (threadIdx.x & 31u);
#define RBD_CALL_UNI(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_) rbd_spec_uni(q_, v_, i_, o0_, o1_, ld_, act_, fl_, pa_, pb_, tm, sh, use_tm)
  {
    RBD_QUEUE_LOOP(RBD_CALL_UNI)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
}
