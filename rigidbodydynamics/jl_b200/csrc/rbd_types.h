// Device-side model layout shared by the host preprocessing (rbd_model.cpp) and the kernels (rbd_device.cuh).
//
// A Mechanism is flattened ONCE (rbd_model_create) into `ModelDev<T>`: bodies in depth-first preorder, each
// with its parent link, canonicalised tree transform, canonicalised inertia and the bookkeeping the three
// passes need.  The struct is passed to the kernels as a __grid_constant__ parameter, so every field is read
// through the constant bank with a warp-uniform index (the body loop index is uniform across the warp).
#pragma once
#include <stdint.h>

namespace rbd {

constexpr int kMaxBodies = 64;

// Joint kinds == rbd_joint_type codes (include/rbd_b200.h) == reference joint types (src/joint_types/*.jl).
enum Kind : int32_t {
  K_REV = 0, K_PRIS = 1, K_FIXED = 2, K_PLANAR = 3, K_QFLOAT = 4, K_SPQFLOAT = 5, K_QSPH = 6, K_SINCOS = 7
};

inline int kind_nq(int k) { const int n[8] = {1, 1, 0, 3, 7, 6, 4, 2}; return n[k]; }
inline int kind_nv(int k) { const int n[8] = {1, 1, 0, 3, 6, 6, 3, 1}; return n[k]; }

// Per-body flags (all warp-uniform).
enum BodyFlags : int32_t {
  F_FIRST_CHILD = 1,   // parent == this - 1 in preorder: parent<->child hand-over stays in registers
  F_SLOT_INIT = 2,     // not first child, and the LAST child of its parent in preorder: initialises the parent's slot
  F_HAS_PENDING = 4,   // has >= 2 children: owns a pending slot (inward accumulation / outward (v, a) save)
  F_LEAF = 8,          // no children
  F_ROOT_CHILD = 16,   // parent is the world: nothing is propagated inward
  // Revolute bodies whose constant tree rotation is a pure z-rotation (joint axis parallel to the parent's) or the cyclic
  // permutation P (P v = (v_z, v_x, v_y): joint axis along the parent frame's +x) times a z-rotation: the whole joint
  // transform is then  E(q) = [P] Rz(q + qoff)  and the spatial transforms of the ABA passes use sparse z-rotations.
  F_ZPAR = 32,
  F_ZPERP = 64,
  F_ZERO_R = 128,      // fast-class body whose frame origin coincides with its parent's (pt == 0): no origin shift inward
};

// Shared-memory "stash" rows per sample.  One row = one scalar per sample (lane); rows are private to a thread.
constexpr int kRowsOneDof = 6;      // v between passes 1 and 2, then U~ (5 non-unit entries) and u~
constexpr int kSlotRowsAba = 27;    // pending articulated inertia (21) + bias force (6); reused for (v, a) = 12
constexpr int kSlotRowsRnea = 12;   // (v, a) outward; 6 for the inward wrench
constexpr int kSlotRowsCrba = 10;   // composite rigid-body inertia (m, h, J)

template <class T> struct BodyDev {
  T Rt[9];            // canonicalised joint_to_predecessor rotation (joint frame -> parent body frame), row-major
  T pt[3];            // ... translation
  T m;                // body inertia in the (canonicalised) body frame: mass,
  T h[3];             //   cross_part = m * com,
  T J[6];             //   moment about the frame origin: xx xy xz yy yz zz
  T qoff;             // F_ZPAR / F_ZPERP bodies: constant z-rotation folded into the joint angle; 0 otherwise
  int32_t kind;       // Kind
  int32_t parent;     // preorder index of the parent body, -1 = world
  int32_t qrow, vrow; // first row of this joint in q / v (reference order!)
  int32_t row0;       // first stash row of this body
  int32_t oslot;      // index of this body's own pending slot (valid if F_HAS_PENDING), else -1
  int32_t pslot;      // index of the parent's pending slot (valid if !F_FIRST_CHILD && !F_ROOT_CHILD), else -1
  int32_t flags;
  int32_t refidx;     // joint index in the reference order (row 6*refidx of wext)
};

template <class T> struct ModelDev {
  int32_t nb, nq, nv;
  int32_t nrows;      // total ABA stash rows per sample (body rows + pending slots)
  int32_t slot_base;  // first ABA stash row of pending slot 0 (slot s starts at slot_base + s * kSlotRowsAba)
  int32_t nslots;
  T g[3];             // gravitational acceleration, root frame
  T pad_;
  BodyDev<T> body[kMaxBodies];
};

}  // namespace rbd
