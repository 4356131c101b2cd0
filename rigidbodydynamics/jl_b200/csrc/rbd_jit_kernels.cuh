// Persistent kernels around the generated per-sample functions (see rbd_jit_prelude.cuh).
//
// Structure: a shared-memory CTA of RBD_SMEM_WARPS warps (each warp owns a [rows][32] slice of the dynamic shared memory) and a
// Tensor-Memory CTA over all 512 TMEM columns (8 warps for scalar fp32 -- one column per row; 4 warps for fp64 and packed fp32 --
// two columns per row), both claiming groups of 32 x RBD_WIDTH consecutive samples -- one group per warp per iteration -- from
// one atomic counter; the next iteration's input lines are prefetched into L2 while the current one is computed.
// Claims are per CTA (every warp of a CTA runs the same number of iterations) so that the optional RBD_CONVOY() barriers inside
// the sample function are safe.
#pragma once

#ifndef RBD_SMEM_WARPS
#define RBD_SMEM_WARPS 8
#endif
#define RBD_TM_WARPS ((RBD_SPEC_F64 || RBD_WIDTH == 2) ? 4 : 8)
#define RBD_GROUP (32 * RBD_WIDTH)

#define RBD_CONVOY_LOOP(NW, BODY)                                                                   \
  __shared__ long long s_claim[2];                                                                  \
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;                                          \
  const long long ngroups = (a.B + RBD_GROUP - 1) / RBD_GROUP;                                      \
  if (threadIdx.x == 0) s_claim[0] = (long long)atomicAdd(a.counter, (unsigned long long)(NW));     \
  __syncthreads();                                                                                  \
  long long g0 = s_claim[0];                                                                        \
  int par = 0;                                                                                      \
  while (g0 < ngroups) {                                                                            \
    if (threadIdx.x == 0) s_claim[par ^ 1] = (long long)atomicAdd(a.counter, (unsigned long long)(NW)); \
    __syncthreads();                                                                                \
    const long long gn0 = s_claim[par ^ 1];                                                         \
    par ^= 1;                                                                                       \
    if (gn0 + w < ngroups) {                                                                        \
      const long long bn = (gn0 + w) * RBD_GROUP;                                                   \
      rbd_prefetch_rows(a.q, RBD_SPEC_NQ, a.ld, bn);                                                \
      rbd_prefetch_rows(a.v, RBD_SPEC_NV, a.ld, bn);                                                \
      if (RBD_SPEC_HAS_IN2) rbd_prefetch_rows(a.in2, RBD_SPEC_NV, a.ld, bn);                        \
    }                                                                                               \
    const long long b = (g0 + w) * RBD_GROUP + RBD_WIDTH * lane;                                    \
    BODY                                                                                            \
    g0 = gn0;                                                                                       \
  }

// aligned I/O: the host guarantees an even sample count (packed mode), so a thread's samples are all real or all idle; idle
// lanes / warps recompute the last group's samples with their stores masked
#define RBD_BODY_ALIGNED(FN, STASH)                                                                 \
    const bool active = b < a.B;                                                                    \
    const long long bl = active ? b : a.B - RBD_WIDTH;                                              \
    FN(a.q + bl, a.v + bl, RBD_SPEC_HAS_IN2 ? a.in2 + bl : (const rbd_f*)0, a.o0 + bl,               \
       RBD_SPEC_HAS_OUT1 ? a.o1 + bl : (rbd_f*)0, a.ld, active, STASH);

extern "C" __global__ void __launch_bounds__(32 * RBD_SMEM_WARPS, (RBD_SPEC_F64 || RBD_WIDTH == 2) ? 1 : (RBD_SMEM_WARPS >= 8 ? 2 : (512 / (32 * RBD_SMEM_WARPS))))
rbd_jit_smem(const RbdJitArgs a) {
  extern __shared__ __align__(16) unsigned char rbd_smem_raw[];
#if RBD_WIDTH == 2
  volatile unsigned long long* sh = reinterpret_cast<volatile unsigned long long*>(rbd_smem_raw) + (threadIdx.x >> 5) * (RBD_SPEC_ROWS * 32) + (threadIdx.x & 31);
#else
  volatile rbd_v* sh = reinterpret_cast<volatile rbd_v*>(rbd_smem_raw) + (threadIdx.x >> 5) * (RBD_SPEC_ROWS * 32) + (threadIdx.x & 31);
#endif
  RBD_CONVOY_LOOP(RBD_SMEM_WARPS, RBD_BODY_ALIGNED(rbd_spec_smem, sh))
}

#if RBD_WIDTH == 2
// Packed mode, any alignment and any sample count: one 32-bit access per half.  Serves arrays whose pairs are not 8-byte aligned
// (odd leading dimension / base) and the last sample of an odd batch; same arithmetic, so results do not depend on which kernel
// a sample went through.
extern "C" __global__ void __launch_bounds__(32 * RBD_SMEM_WARPS, 1) rbd_jit_smem32(const RbdJitArgs a) {
  extern __shared__ __align__(16) unsigned char rbd_smem_raw[];
  volatile unsigned long long* sh = reinterpret_cast<volatile unsigned long long*>(rbd_smem_raw) + (threadIdx.x >> 5) * (RBD_SPEC_ROWS * 32) + (threadIdx.x & 31);
#define RBD_BODY_IO32                                                                               \
    const bool active = b < a.B, active1 = b + 1 < a.B;                                             \
    const long long bl = active ? b : a.B - 1;                                                      \
    const int d1 = active1 ? 1 : 0;                                                                 \
    rbd_spec_smem32(a.q + bl, a.v + bl, RBD_SPEC_HAS_IN2 ? a.in2 + bl : (const rbd_f*)0, a.o0 + bl,  \
                    RBD_SPEC_HAS_OUT1 ? a.o1 + bl : (rbd_f*)0, a.ld, active, d1, active1, sh);
  RBD_CONVOY_LOOP(RBD_SMEM_WARPS, RBD_BODY_IO32)
}
#endif

// Warp w of the Tensor-Memory CTA uses lane quadrant w % 4 and, in the 8-warp CTA, column half w / 4.
extern "C" __global__ void __launch_bounds__(32 * RBD_TM_WARPS, (RBD_SPEC_F64 || RBD_WIDTH == 2) ? 1 : 2) rbd_jit_tmem(const RbdJitArgs a) {
  __shared__ unsigned tm_slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&tm_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm_base = tm_slot;
  const unsigned tm = tm_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16) + (threadIdx.x >> 7) * 256u;
  {
    RBD_CONVOY_LOOP(RBD_TM_WARPS, RBD_BODY_ALIGNED(rbd_spec_tmem, tm))
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
}
