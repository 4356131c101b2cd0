// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See rbd_oracle.hpp for scope, citations and parity pinning.
//
// C entry points over the templated restatement so that tests (ctypes) and bench.py's CPU-baseline leg
// can run it on the same structure-of-arrays batches the GPU library takes: every array is [rows][B] with
// the batch index fastest (element (k, b) at x[k*B + b]).
#include "rbd_oracle.hpp"

#include <algorithm>
#include <thread>

using namespace rbdo;

namespace {

template <class F> void parallel_for(int64_t B, int nthreads, F f) {
  if (nthreads <= 1 || B < 2) { f(0, B, 0); return; }
  nthreads = (int)std::min<int64_t>(nthreads, B);
  std::vector<std::thread> th;
  int64_t chunk = (B + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    int64_t lo = t * chunk, hi = std::min<int64_t>(B, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

// The batch arrays are rows x batch with the batch fastest (the GPU layout).  Walking ONE sample through them touches a
// different page per row; the drivers below therefore move TILES of kTile consecutive samples (contiguous row segments)
// into a sample-major local buffer, evaluate, and move results back the same way, so the CPU baseline is not penalised by
// a layout chosen for the GPU.
constexpr int64_t kTile = 128;
template <class T> inline void tile_in(const T* src, int64_t B, int64_t b0, int64_t nt, int n, T* dst /*[nt][n]*/) {
  if (!src) return;
  for (int k = 0; k < n; ++k) {
    const T* row = src + (int64_t)k * B + b0;
    for (int64_t j = 0; j < nt; ++j) dst[j * n + k] = row[j];
  }
}
template <class T> inline void tile_out(T* dst, int64_t B, int64_t b0, int64_t nt, int n, const T* src /*[nt][n]*/) {
  if (!dst) return;
  for (int k = 0; k < n; ++k) {
    T* row = dst + (int64_t)k * B + b0;
    for (int64_t j = 0; j < nt; ++j) row[j] = src[j * n + k];
  }
}

template <class T>
int dynamics_t(const Model& m, int64_t B, const T* q, const T* v, const T* tau, const T* wext, T* vd, T* qd,
               int algo, int nthreads) {
  std::vector<int> status(std::max(1, nthreads), 0);
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int tid) {
    Workspace<T> w(m);
    const int nw = m.nb * 6;
    std::vector<T> ql(kTile * m.nq), vl(kTile * m.nv), tl(kTile * m.nv), wl(kTile * nw), vdl(kTile * m.nv), qdl(kTile * m.nq);
    for (int64_t b0 = lo; b0 < hi; b0 += kTile) {
      const int64_t nt = std::min<int64_t>(kTile, hi - b0);
      tile_in(q, B, b0, nt, m.nq, ql.data());
      tile_in(v, B, b0, nt, m.nv, vl.data());
      tile_in(tau, B, b0, nt, m.nv, tl.data());
      tile_in(wext, B, b0, nt, nw, wl.data());
      for (int64_t j = 0; j < nt; ++j) {
        const T* tj = tau ? &tl[j * m.nv] : nullptr;
        const T* wj = wext ? &wl[j * nw] : nullptr;
        bool ok;
        if (algo == 0) {
          ok = dynamics(w, &ql[j * m.nq], &vl[j * m.nv], tj, wj, &vdl[j * m.nv], qd ? &qdl[j * m.nq] : nullptr);
        } else {
          ok = aba(w, &ql[j * m.nq], &vl[j * m.nv], tj, wj, &vdl[j * m.nv]);
          if (qd) configuration_derivative(m, &ql[j * m.nq], &vl[j * m.nv], &qdl[j * m.nq]);
        }
        if (!ok) status[tid] = 1;
      }
      tile_out(vd, B, b0, nt, m.nv, vdl.data());
      tile_out(qd, B, b0, nt, m.nq, qdl.data());
    }
  });
  for (int s : status) if (s) return 1;
  return 0;
}

template <class T>
int inverse_dynamics_t(const Model& m, int64_t B, const T* q, const T* v, const T* vd, const T* wext, T* tau, int nthreads) {
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int) {
    Workspace<T> w(m);
    const int nw = m.nb * 6;
    std::vector<T> ql(kTile * m.nq), vl(kTile * m.nv), vdl(kTile * m.nv), wl(kTile * nw), tl(kTile * m.nv);
    for (int64_t b0 = lo; b0 < hi; b0 += kTile) {
      const int64_t nt = std::min<int64_t>(kTile, hi - b0);
      tile_in(q, B, b0, nt, m.nq, ql.data());
      tile_in(v, B, b0, nt, m.nv, vl.data());
      tile_in(vd, B, b0, nt, m.nv, vdl.data());
      tile_in(wext, B, b0, nt, nw, wl.data());
      for (int64_t j = 0; j < nt; ++j) {
        const T* wj = wext ? &wl[j * nw] : nullptr;
        if (vd) inverse_dynamics(w, &ql[j * m.nq], &vl[j * m.nv], &vdl[j * m.nv], wj, &tl[j * m.nv]);
        else dynamics_bias(w, &ql[j * m.nq], &vl[j * m.nv], wj, &tl[j * m.nv]);
      }
      tile_out(tau, B, b0, nt, m.nv, tl.data());
    }
  });
  return 0;
}

// per-body caches of inverse_dynamics!: `accelerations` (spatial_accelerations! :387-417) and the joint wrenches left in
// `jointwrenchesout` by joint_wrenches_and_torques! (:442-459), both in the root frame, [6*nb x B]
template <class T>
int inverse_dynamics_bodies_t(const Model& m, int64_t B, const T* q, const T* v, const T* vd, const T* wext, T* acc, T* jw) {
  Workspace<T> w(m);
  const int nw = m.nb * 6;
  std::vector<T> ql(m.nq), vl(m.nv), vdl(m.nv), wl(nw), tl(m.nv);
  for (int64_t b = 0; b < B; ++b) {
    for (int k = 0; k < m.nq; ++k) ql[k] = q[(size_t)k * B + b];
    for (int k = 0; k < m.nv; ++k) { vl[k] = v[(size_t)k * B + b]; vdl[k] = vd ? vd[(size_t)k * B + b] : T(0); }
    for (int k = 0; k < nw; ++k) wl[k] = wext ? wext[(size_t)k * B + b] : T(0);
    if (vd) inverse_dynamics(w, ql.data(), vl.data(), vdl.data(), wext ? wl.data() : nullptr, tl.data());
    else dynamics_bias(w, ql.data(), vl.data(), wext ? wl.data() : nullptr, tl.data());
    for (int i = 0; i < m.nb; ++i)
      for (int k = 0; k < 3; ++k) {
        if (acc) { acc[(size_t)(6 * i + k) * B + b] = w.accel[i].ang[k]; acc[(size_t)(6 * i + 3 + k) * B + b] = w.accel[i].lin[k]; }
        if (jw) { jw[(size_t)(6 * i + k) * B + b] = w.wrench[i].ang[k]; jw[(size_t)(6 * i + 3 + k) * B + b] = w.wrench[i].lin[k]; }
      }
  }
  return 0;
}

// contact_dynamics!(result, state)                                       mechanism_algorithms.jl:680-723
// with SoftContactModel{HuntCrossleyModel, ViscoelasticCoulombModel}    contact.jl:104-118, :130-146, :152-206
// and HalfSpace3D environment primitives                                 contact.jl:219-239
//   body[p]: tree joint whose successor carries point p;  loc [np][3] in that body's frame;  hc [np][3] = k, lambda, n;
//   fr [np][3] = mu, k, b;  hs [nh][6] = point, outward normal;  state s / derivative sd [3*np*nh x B] (pair (p, h) at rows
//   3 (p nh + h) ..), s is reset where the pair is not in contact (:716 reset!);  wr [6 nb x B] = result.contactwrenches.
template <class T>
int contact_dynamics_t(const Model& m, int64_t B, const T* q, const T* v, int np, const int* body, const double* loc,
                       const double* hc, const double* fr, int nh, const double* hs, T* s, T* sd, T* wr) {
  Workspace<T> w(m);
  std::vector<T> ql(m.nq), vl(m.nv);
  std::vector<S6<T>> cw(m.nb);
  for (int64_t b = 0; b < B; ++b) {
    for (int k = 0; k < m.nq; ++k) ql[k] = q[(size_t)k * B + b];
    for (int k = 0; k < m.nv; ++k) vl[k] = v[(size_t)k * B + b];
    update_transforms(w, ql.data());
    update_twists(w, vl.data());
    for (int i = 0; i < m.nb; ++i) cw[i] = S6<T>();                                        // :688 zero(Wrench)
    for (int p = 0; p < np; ++p) {
      const int i = body[p];
      const S6<T>& tw = w.twist[i];                                                      // :693 twist_wrt_world
      const V3<T> lp(T(loc[3 * p]), T(loc[3 * p + 1]), T(loc[3 * p + 2]));
      const V3<T> pt = w.T_root[i].R * lp + w.T_root[i].p;                                // :698 body_to_root * location
      const V3<T> vel = cross(tw.ang, pt) + tw.lin;                                       // :699 point_velocity
      for (int h = 0; h < nh; ++h) {
        const V3<T> hp(T(hs[6 * h]), T(hs[6 * h + 1]), T(hs[6 * h + 2]));
        V3<T> n(T(hs[6 * h + 3]), T(hs[6 * h + 4]), T(hs[6 * h + 5]));
        n = n * (T(1) / T(std::sqrt(dot(n, n))));                                          // contact.jl:225 normalize
        const size_t row = (size_t)3 * (p * nh + h);
        const T sep = dot(pt - hp, n);                                                    // contact.jl:237 separation
        V3<T> xd;
        if (sep <= T(0)) {                                                                // :710, contact.jl:238 point_inside
          const T z = -sep, zd = -dot(vel, n);                                            // contact.jl:108-109
          const T zn = T(std::pow(z, T(hc[3 * p + 2])));
          T fn = T(hc[3 * p + 1]) * zn * zd + T(hc[3 * p]) * zn;                          // contact.jl:143-146
          if (fn < T(0)) fn = T(0);                                                       // contact.jl:110 max(., 0)
          const V3<T> vt = vel + zd * n;                                                  // contact.jl:113 tangential velocity
          const T mu = T(fr[3 * p]), kf = T(fr[3 * p + 1]), bf = T(fr[3 * p + 2]);
          const V3<T> x = s ? V3<T>(s[(row + 0) * B + b], s[(row + 1) * B + b], s[(row + 2) * B + b]) : V3<T>();
          V3<T> ft = -(kf * x) - bf * vt;                                                 // contact.jl:188 fstick
          const T n2 = dot(ft, ft), m2 = (mu * fn) * (mu * fn);
          if (n2 > m2) ft = ft * T(std::sqrt(m2 / n2));                                   // contact.jl:191-197 Coulomb cone
          xd = (-(kf * x) - ft) * (T(1) / bf);                                            // contact.jl:205 state derivative
          const V3<T> f = fn * n + ft;                                                    // contact.jl:117
          cw[i] = cw[i] + S6<T>{cross(pt, f), f};                                         // :714 Wrench(point, force)
        } else if (s) {
          for (int k = 0; k < 3; ++k) s[(row + k) * B + b] = T(0);                        // :716 reset!
        }
        if (sd) { sd[(row + 0) * B + b] = xd.x; sd[(row + 1) * B + b] = xd.y; sd[(row + 2) * B + b] = xd.z; }   // :717 zero!
      }
    }
    for (int i = 0; i < m.nb; ++i)
      for (int k = 0; k < 3; ++k) { wr[(size_t)(6 * i + k) * B + b] = cw[i].ang[k]; wr[(size_t)(6 * i + 3 + k) * B + b] = cw[i].lin[k]; }
  }
  return 0;
}

template <class T> int mass_matrix_t(const Model& m, int64_t B, const T* q, T* M, int nthreads) {
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int) {
    Workspace<T> w(m);
    const int nm = m.nv * m.nv;
    std::vector<T> ql(kTile * m.nq), Ml((size_t)kTile * nm);
    for (int64_t b0 = lo; b0 < hi; b0 += kTile) {
      const int64_t nt = std::min<int64_t>(kTile, hi - b0);
      tile_in(q, B, b0, nt, m.nq, ql.data());
      for (int64_t j = 0; j < nt; ++j) {
        update_transforms(w, &ql[j * m.nq]);
        update_motion_subspaces(w);
        update_spatial_inertias(w);
        update_crb_inertias(w);
        mass_matrix(w, &Ml[j * nm]);
      }
      tile_out(M, B, b0, nt, nm, Ml.data());
    }
  });
  return 0;
}

template <class T>
int kinematics_t(const Model& m, int64_t B, const T* q, const T* v, const signed char* sign, T* tr, T* com, T* ke, T* pe,
                 T* mom, T* mrb, T* A, T* J, int nthreads) {
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int) {
    Workspace<T> w(m);
    const int nt12 = 12 * m.nb, n6v = 6 * m.nv;
    std::vector<T> ql(m.nq), vl(m.nv), trl(nt12), Al(n6v), Jl(n6v);
    T c3[3], k1[1], p1[1], h6[6], b6[6];
    for (int64_t b = lo; b < hi; ++b) {
      for (int k = 0; k < m.nq; ++k) ql[k] = q[(int64_t)k * B + b];
      if (v) for (int k = 0; k < m.nv; ++k) vl[k] = v[(int64_t)k * B + b];
      KinOut<T> o;
      o.transforms = tr ? trl.data() : nullptr; o.com = com ? c3 : nullptr; o.ke = ke ? k1 : nullptr; o.pe = pe ? p1 : nullptr;
      o.momentum = mom ? h6 : nullptr; o.mrb = mrb ? b6 : nullptr; o.A = A ? Al.data() : nullptr; o.J = J ? Jl.data() : nullptr;
      kinematics(w, ql.data(), v ? vl.data() : nullptr, sign, o);
      auto put = [&](T* dst, const T* src, int n) { if (dst) for (int k = 0; k < n; ++k) dst[(int64_t)k * B + b] = src[k]; };
      put(tr, trl.data(), nt12); put(com, c3, 3); put(ke, k1, 1); put(pe, p1, 1); put(mom, h6, 6); put(mrb, b6, 6);
      put(A, Al.data(), n6v); put(J, Jl.data(), n6v);
    }
  });
  return 0;
}

// Dual{Float64,6} arrays: [rows][B][7] doubles (value, 6 partials) -- Julia's memory layout of Matrix{Dual}(B, n)
int dynamics_dual6(const Model& m, int64_t B, const double* q, const double* v, const double* tau, double* vd, int algo,
                   int nthreads) {
  using D = DualN<6>;
  std::vector<int> status(std::max(1, nthreads), 0);
  auto gatherd = [](const double* src, int64_t B, int64_t b, int n, D* dst) {
    for (int k = 0; k < n; ++k) {
      const double* e = src + ((int64_t)k * B + b) * 7;
      dst[k].v = e[0];
      for (int i = 0; i < 6; ++i) dst[k].d[i] = e[1 + i];
    }
  };
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int tid) {
    Workspace<D> w(m);
    std::vector<D> ql(m.nq), vl(m.nv), tl(m.nv), vdl(m.nv);
    for (int64_t b = lo; b < hi; ++b) {
      gatherd(q, B, b, m.nq, ql.data());
      gatherd(v, B, b, m.nv, vl.data());
      if (tau) gatherd(tau, B, b, m.nv, tl.data());
      bool ok = algo == 0 ? dynamics(w, ql.data(), vl.data(), tau ? tl.data() : nullptr, (const D*)nullptr, vdl.data(), (D*)nullptr)
                          : aba(w, ql.data(), vl.data(), tau ? tl.data() : nullptr, (const D*)nullptr, vdl.data());
      if (!ok) status[tid] = 1;
      for (int k = 0; k < m.nv; ++k) {
        double* e = vd + ((int64_t)k * B + b) * 7;
        e[0] = vdl[k].v;
        for (int i = 0; i < 6; ++i) e[1 + i] = vdl[k].d[i];
      }
    }
  });
  for (int s : status) if (s) return 1;
  return 0;
}

}  // namespace

extern "C" {

void* rbdo_model_create(int nb, const int* parent, const int* jtype, const double* X_tree, const double* jparam,
                        const double* inertia, const double* gravity) {
  Model* m = new Model();
  m->nb = nb;
  m->parent.assign(parent, parent + nb);
  m->jtype.assign(jtype, jtype + nb);
  m->qstart.resize(nb);
  m->vstart.resize(nb);
  int nq = 0, nv = 0;
  for (int i = 0; i < nb; ++i) {
    m->qstart[i] = nq; m->vstart[i] = nv;
    nq += joint_nq(jtype[i]); nv += joint_nv(jtype[i]);
  }
  m->nq = nq; m->nv = nv;
  m->X_tree.assign(X_tree, X_tree + 12 * (size_t)nb);
  m->jparam.assign(jparam, jparam + 9 * (size_t)nb);
  m->inertia.assign(inertia, inertia + 13 * (size_t)nb);
  for (int k = 0; k < 3; ++k) m->gravity[k] = gravity[k];
  m->finalize();
  return m;
}
void rbdo_model_destroy(void* m) { delete static_cast<Model*>(m); }
int rbdo_nq(void* m) { return static_cast<Model*>(m)->nq; }
int rbdo_nv(void* m) { return static_cast<Model*>(m)->nv; }

// dtype: 0 = float32, 1 = float64.  algo: 0 = reference path (RNEA bias + CRBA + Cholesky), 1 = world-frame ABA.
int rbdo_dynamics(void* mp, int dtype, int64_t B, const void* q, const void* v, const void* tau, const void* wext,
                  void* vd, void* qd, int algo, int nthreads) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return dynamics_t<float>(m, B, (const float*)q, (const float*)v, (const float*)tau, (const float*)wext, (float*)vd, (float*)qd, algo, nthreads);
  return dynamics_t<double>(m, B, (const double*)q, (const double*)v, (const double*)tau, (const double*)wext, (double*)vd, (double*)qd, algo, nthreads);
}
int rbdo_dynamics_dual6(void* mp, int64_t B, const double* q, const double* v, const double* tau, double* vd, int algo,
                        int nthreads) {
  return dynamics_dual6(*static_cast<Model*>(mp), B, q, v, tau, vd, algo, nthreads);
}
// simulate(): nsteps RK4 Munthe-Kaas steps, constant torques, fp64; q [nq][B], v [nv][B] updated in place
int rbdo_integrate(void* mp, int64_t B, double* q, double* v, const double* tau, double dt, int nsteps, int nthreads) {
  const Model& m = *static_cast<Model*>(mp);
  std::vector<int> status(std::max(1, nthreads), 0);
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int tid) {
    Workspace<double> w(m);
    std::vector<double> ql(m.nq), vl(m.nv), tl(m.nv);
    for (int64_t b = lo; b < hi; ++b) {
      for (int k = 0; k < m.nq; ++k) ql[k] = q[(int64_t)k * B + b];
      for (int k = 0; k < m.nv; ++k) { vl[k] = v[(int64_t)k * B + b]; if (tau) tl[k] = tau[(int64_t)k * B + b]; }
      for (int s = 0; s < nsteps; ++s)
        if (!integrate_step(w, ql.data(), vl.data(), tau ? tl.data() : nullptr, dt)) status[tid] = 1;
      for (int k = 0; k < m.nq; ++k) q[(int64_t)k * B + b] = ql[k];
      for (int k = 0; k < m.nv; ++k) v[(int64_t)k * B + b] = vl[k];
    }
  });
  for (int s : status) if (s) return 1;
  return 0;
}
// vd == NULL  =>  dynamics_bias
int rbdo_inverse_dynamics(void* mp, int dtype, int64_t B, const void* q, const void* v, const void* vd, const void* wext,
                          void* tau, int nthreads) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return inverse_dynamics_t<float>(m, B, (const float*)q, (const float*)v, (const float*)vd, (const float*)wext, (float*)tau, nthreads);
  return inverse_dynamics_t<double>(m, B, (const double*)q, (const double*)v, (const double*)vd, (const double*)wext, (double*)tau, nthreads);
}
int rbdo_inverse_dynamics_bodies(void* mp, int dtype, int64_t B, const void* q, const void* v, const void* vd, const void* wext,
                                 void* acc, void* jw) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return inverse_dynamics_bodies_t<float>(m, B, (const float*)q, (const float*)v, (const float*)vd, (const float*)wext, (float*)acc, (float*)jw);
  return inverse_dynamics_bodies_t<double>(m, B, (const double*)q, (const double*)v, (const double*)vd, (const double*)wext, (double*)acc, (double*)jw);
}
int rbdo_contact_dynamics(void* mp, int dtype, int64_t B, const void* q, const void* v, int np, const int* body, const double* loc,
                          const double* hc, const double* fr, int nh, const double* hs, void* s, void* sd, void* wr) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return contact_dynamics_t<float>(m, B, (const float*)q, (const float*)v, np, body, loc, hc, fr, nh, hs, (float*)s, (float*)sd, (float*)wr);
  return contact_dynamics_t<double>(m, B, (const double*)q, (const double*)v, np, body, loc, hc, fr, nh, hs, (double*)s, (double*)sd, (double*)wr);
}
int rbdo_mass_matrix(void* mp, int dtype, int64_t B, const void* q, void* M, int nthreads) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return mass_matrix_t<float>(m, B, (const float*)q, (float*)M, nthreads);
  return mass_matrix_t<double>(m, B, (const double*)q, (double*)M, nthreads);
}

// kinematics by-products; every output pointer may be NULL; v may be NULL when no velocity-dependent output is requested
int rbdo_kinematics(void* mp, int dtype, int64_t B, const void* q, const void* v, const signed char* sign, void* tr, void* com,
                    void* ke, void* pe, void* mom, void* mrb, void* A, void* J, int nthreads) {
  const Model& m = *static_cast<Model*>(mp);
  if (dtype == 0) return kinematics_t<float>(m, B, (const float*)q, (const float*)v, sign, (float*)tr, (float*)com, (float*)ke, (float*)pe, (float*)mom, (float*)mrb, (float*)A, (float*)J, nthreads);
  return kinematics_t<double>(m, B, (const double*)q, (const double*)v, sign, (double*)tr, (double*)com, (double*)ke, (double*)pe, (double*)mom, (double*)mrb, (double*)A, (double*)J, nthreads);
}

}  // extern "C"
