// Access macros of the generated per-sample function for one variant (no include guard: included once per variant).
//   RBD_FLAVOR_SMEM / RBD_FLAVOR_TMEM : home of the stash;   RBD_IO32 (packed mode only): global I/O with one 32-bit access per
//   half (unaligned arrays, odd tail) instead of one 64-bit access per pair.
#undef RBD_STASH_ARG
#undef RBD_SLD
#undef RBD_SST
#undef RBD_SFENCE
#undef RBD_TM_REG
#undef RBD_TM_LD
#undef RBD_TM_WAIT_LD
#undef RBD_TM_VAL
#undef RBD_IO_ARGS
#undef RBD_LDG
#undef RBD_STG

#if RBD_WIDTH == 1
#define RBD_IO_ARGS const rbd_f* __restrict__ q, const rbd_f* __restrict__ v, const rbd_f* __restrict__ in2, rbd_f* __restrict__ o0, \
                    rbd_f* __restrict__ o1, const long long ld, const bool active
#define RBD_LDG(p, r) __ldg((p) + (long long)(r) * ld)
#define RBD_STG(p, r, val_) do { if (active) (p)[(long long)(r) * ld] = (val_); } while (0)
#elif !defined(RBD_IO32)
// pointers address the FIRST sample of the thread's pair; the pair is adjacent and 8-byte aligned
#define RBD_IO_ARGS const rbd_f* __restrict__ q, const rbd_f* __restrict__ v, const rbd_f* __restrict__ in2, rbd_f* __restrict__ o0, \
                    rbd_f* __restrict__ o1, const long long ld, const bool active
#define RBD_LDG(p, r) __ldg(reinterpret_cast<const unsigned long long*>((p) + (long long)(r) * ld))
#define RBD_STG(p, r, val_) do { if (active) *reinterpret_cast<unsigned long long*>((p) + (long long)(r) * ld) = (val_); } while (0)
#else
// second sample at offset d1 (0 when the thread only has one real sample); stores masked per half
#define RBD_IO_ARGS const rbd_f* __restrict__ q, const rbd_f* __restrict__ v, const rbd_f* __restrict__ in2, rbd_f* __restrict__ o0, \
                    rbd_f* __restrict__ o1, const long long ld, const bool active, const int d1, const bool active1
#define RBD_LDG(p, r) rbd_pack2(__ldg((p) + (long long)(r) * ld), __ldg((p) + (long long)(r) * ld + d1))
#define RBD_STG(p, r, val_) do { float lo_, hi_; rbd_unpack2(val_, lo_, hi_); if (active) (p)[(long long)(r) * ld] = lo_; \
                                 if (active1) (p)[(long long)(r) * ld + d1] = hi_; } while (0)
#endif

#if defined(RBD_FLAVOR_SMEM)
// shared memory, [row][lane] of one warp: conflict-free, constant offsets.  volatile: with every row a compile-time constant the
// compiler would otherwise forward each pass-1 store to its pass-2 load THROUGH REGISTERS (and spill them) -- the stash exists to
// get those values out of the register file.
#if RBD_WIDTH == 2
#define RBD_STASH_ARG volatile unsigned long long* sh      /* one 64-bit access per pair */
#define RBD_SLD(r) sh[(r) * 32]
#define RBD_SST(r, val_) sh[(r) * 32] = (val_)
#else
#define RBD_STASH_ARG volatile rbd_v* sh
#define RBD_SLD(r) sh[(r) * 32]
#define RBD_SST(r, val_) sh[(r) * 32] = (val_)
#endif
#define RBD_SFENCE()
#elif defined(RBD_FLAVOR_TMEM)
// Tensor Memory (rbd_tmem.cuh): lane l of warp w owns TMEM lane 32 (w % 4) + l; one 32-bit column per fp32 row, two per fp64 row
// or packed fp32 pair
#define RBD_STASH_ARG const unsigned tm
#define RBD_SFENCE() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define RBD_TM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#if RBD_SPEC_F64 || RBD_WIDTH == 2
struct rbd_u2 { unsigned lo, hi; };
#define RBD_TM_REG rbd_u2
#define RBD_TM_LD(u, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(u.lo), "=r"(u.hi) : "r"(tm + 2u * (r)) : "memory")
#if RBD_SPEC_F64
#define RBD_TM_VAL(u) __hiloint2double((int)u.hi, (int)u.lo)
#define RBD_SST(r, val_) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tm + 2u * (r)), "r"((unsigned)__double2loint(val_)), "r"((unsigned)__double2hiint(val_)) : "memory")
#else
#define RBD_TM_VAL(u) (((unsigned long long)u.hi << 32) | u.lo)
#define RBD_SST(r, val_) asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {lo, hi}; }" ::"r"(tm + 2u * (r)), "l"(val_) : "memory")
#endif
#else
#define RBD_TM_REG unsigned
#define RBD_TM_LD(u, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(u) : "r"(tm + (r)) : "memory")
#define RBD_TM_VAL(u) __uint_as_float(u)
#define RBD_SST(r, val_) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + (r)), "r"(__float_as_uint(val_)) : "memory")
#endif
#endif
