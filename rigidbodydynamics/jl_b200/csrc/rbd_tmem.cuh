// Tensor Memory as a per-thread scratchpad (sm_100a only).
//
// Each SM has 256 KB of TMEM (512 columns x 128 lanes x 32 bit) next to its 228 KB of shared memory.  It exists to hold
// tcgen05.mma accumulators, but tcgen05.ld / tcgen05.st move 32-bit words between registers and TMEM with the "32x32b"
// shape: lane l of warp w touches TMEM lane 32*(w % 4) + l, one column per register.  That is exactly the access pattern of
// this library's stash (one private scalar per thread and row), so a CTA of 128 threads that allocates N columns gets N
// private words per thread, and a CTA of 256 threads that allocates all 512 columns gets 256 per thread (warps 4-7 share the
// lane quadrants of warps 0-3 and use the upper half of the columns) -- a second on-chip home for the per-sample working
// set, doubling the number of samples an SM can keep in flight for this latency-bound kernel.
#pragma once
#include <stdint.h>

namespace rbd {

#if defined(__CUDACC__)
struct StashTM {
  uint32_t base;     // TMEM address of row 0 for this warp: (lane quadrant << 16) | first column

  // tcgen05.st is asynchronous: a later tcgen05.ld of the same word needs tcgen05.wait::st in between.  The algorithms
  // call fence_st() exactly where a thread re-reads what it wrote (pass boundaries, pending slots, parent rows).
  __device__ __forceinline__ void fence_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  template <int N> __device__ __forceinline__ void ldv(int row, float* out) const {
    uint32_t r[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[k]) : "r"(base + (uint32_t)(row + k)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = __uint_as_float(r[k]);
  }
  __device__ __forceinline__ float ld(int row) const {
    float v;
    ldv<1>(row, &v);
    return v;
  }
  __device__ __forceinline__ void st(int row, float v) const {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + (uint32_t)row), "r"(__float_as_uint(v)) : "memory");
  }
  __device__ __forceinline__ void add(int row, float v) const { fence_st(); st(row, ld(row) + v); }
  __device__ __forceinline__ const StashTM& slots() const { return *this; }
};

// fp64 stash in Tensor Memory: one row = two adjacent 32-bit columns (tcgen05.ld/st .x2).
struct StashTM64 {
  uint32_t base;
  __device__ __forceinline__ void fence_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  template <int N> __device__ __forceinline__ void ldv(int row, double* out) const {
    uint32_t lo[N], hi[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo[k]), "=r"(hi[k]) : "r"(base + 2u * (uint32_t)(row + k)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = __hiloint2double((int)hi[k], (int)lo[k]);
  }
  __device__ __forceinline__ double ld(int row) const {
    double v;
    ldv<1>(row, &v);
    return v;
  }
  __device__ __forceinline__ void st(int row, double v) const {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(base + 2u * (uint32_t)row), "r"((uint32_t)__double2loint(v)), "r"((uint32_t)__double2hiint(v)) : "memory");
  }
  __device__ __forceinline__ void add(int row, double v) const { fence_st(); st(row, ld(row) + v); }
  __device__ __forceinline__ const StashTM64& slots() const { return *this; }
};
// Dual{Float64,1} stash split over both on-chip memories: the VALUE part of a row in shared memory ([row][thread], 8 bytes), the
// PARTIAL part in Tensor Memory (two 32-bit columns per row).  A 128-thread CTA then needs rows x 1 KB of shared memory and
// 2 x rows TMEM columns -- Atlas: 213 KB + 426 columns = 4 resident warps/SM instead of the 2 that fit in shared memory alone.
#if defined(RBD_DUAL_TYPES)
struct StashDualTM {
  double* pv;        // shared memory, already offset by the thread index; row stride kThreads
  uint32_t base;     // TMEM address of row 0 for this warp
  static constexpr int kThreads = 128;
  __device__ __forceinline__ void fence_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  template <int N> __device__ __forceinline__ void ldv(int row, Dual64* out) const {
    uint32_t lo[N], hi[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo[k]), "=r"(hi[k]) : "r"(base + 2u * (uint32_t)(row + k)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = Dual64(pv[(row + k) * kThreads], __hiloint2double((int)hi[k], (int)lo[k]));
  }
  __device__ __forceinline__ Dual64 ld(int row) const {
    Dual64 v;
    ldv<1>(row, &v);
    return v;
  }
  __device__ __forceinline__ void st(int row, const Dual64& x) const {
    pv[row * kThreads] = x.v;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(base + 2u * (uint32_t)row), "r"((uint32_t)__double2loint(x.d)), "r"((uint32_t)__double2hiint(x.d)) : "memory");
  }
  __device__ __forceinline__ void add(int row, const Dual64& x) const { fence_st(); st(row, ld(row) + x); }
  __device__ __forceinline__ const StashDualTM& slots() const { return *this; }
};
#endif
template <class T> struct StashTMFor;
template <> struct StashTMFor<float> { using type = StashTM; static constexpr int kColsPerRow = 1; };
template <> struct StashTMFor<double> { using type = StashTM64; static constexpr int kColsPerRow = 2; };

// Allocate `cols` (power of two >= 32) TMEM columns for the CTA; every thread returns the base address.
template <uint32_t COLS> __device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t* smem_slot) {
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *smem_slot;
}
template <uint32_t COLS> __device__ __forceinline__ void tmem_free_cta(uint32_t addr) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
#endif

}  // namespace rbd
