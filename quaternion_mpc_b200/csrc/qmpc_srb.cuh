// qmpc_srb.cuh — kernel "srb": structure-exploiting AL-iLQR for the quaternion single-rigid-body
// models (QuatMpc 4 feet / 2-contact model).  Same algorithm and decisions as the dense kernel and
// the CPU oracle, but every product is written for the sparsity the model actually has:
//
//   error state  e = [dp, phi, dv, dw] (4 blocks of 3),  input u = [f_1..f_NF]
//   A_k = [[I,0,hI,0],[0,Aff,0,Afw],[0,0,I,0],[0,0,0,I]]        only Aff, Afw (3x3) vary with k
//   B_k = M_k W,  W = [[I/m ...],[I^-1 skew(r_1) ...]] (6 x NU, constant over horizon AND iterations)
//                 M_k = [[h h/2 I,0],[0,Cf],[hI,0],[0,hI]] (12 x 6), only Cf (3x3) varies
//   cone rows act on one foot each  =>  R + rho J^T D J =: D is block-diagonal (3x3 per foot)
//
// so  Quu = D + W^T S W,  Qux = W^T T,  Qu = g + W^T s  with the 6-dim wrench-space quantities
// S = M^T P M, T = M^T P A, s = M^T p, all formed from 3x3 block products.  Quu (NU x NU) is then
// factored with the same Cholesky as the reference algorithm: R = 1e-6 makes Quu ill-conditioned
// (cond ~1e8, the internal-force directions), and a Woodbury/6x6 route was measured to lose ~5 digits
// to cancellation there (1.7e-4 N worst-case GRF deviation vs the oracle) - the explicit Cholesky
// keeps the worst case at the 1e-7 N level.  The value update uses P = Qxx + Qux^T K, whose error is
// insensitive to the ill-conditioned directions because Qux lies in range(W^T).
// ~5.5 kFMA per knot instead of ~28 kFMA dense; no A_k / B_k storage.
//
// Math restated from: legged_ctrl/src/utils/AltroUtils.cpp:78-110 (midpoint chain rule),
// :395-439 (continuous Jacobian), :153-168 (error-state projection), QuatMpc.cpp:194-215 (cones).
#pragma once
#include "qmpc_dense.cuh"

namespace qmpc {

template <int NF>
struct SrbLayout {
  static constexpr int NU = 3 * NF, NC = 6 * NF;
  QMPC_HD static size_t X(int N) { return 0; }
  QMPC_HD static size_t Xn(int N) { return X(N) + (size_t)(N + 1) * 13; }
  QMPC_HD static size_t U(int N) { return Xn(N) + (size_t)(N + 1) * 13; }
  QMPC_HD static size_t Un(int N) { return U(N) + (size_t)N * NU; }
  QMPC_HD static size_t mu(int N) { return Un(N) + (size_t)N * NU; }
  QMPC_HD static size_t Y(int N) { return mu(N) + (size_t)N * NC; }
  QMPC_HD static size_t P(int N) { return Y(N) + (size_t)(N + 1) * 12; }     // packed upper triangle, 78
  QMPC_HD static size_t pv(int N) { return P(N) + (size_t)(N + 1) * 78; }
  QMPC_HD static size_t lin(int N) { return pv(N) + (size_t)(N + 1) * 12; }  // Aff, Afw, Cf
  QMPC_HD static size_t K(int N) { return lin(N) + (size_t)N * 27; }         // NU x 12
  QMPC_HD static size_t d(int N) { return K(N) + (size_t)N * NU * 12; }
  QMPC_HD static size_t total(int N) { return d(N) + (size_t)N * NU; }
};

struct KnotLin {
  double Aff[9], Afw[9], Cf[9];
};

// Omega(w) = d(qdot)/dq = 0.5 [[0,-w^T],[w,-skew(w)]]   (AltroUtils.cpp:408-410)
QMPC_HD inline void fill_omega(const double* w, double* Om) {
  Om[0] = 0;            Om[1] = -0.5 * w[0];  Om[2] = -0.5 * w[1];  Om[3] = -0.5 * w[2];
  Om[4] = 0.5 * w[0];   Om[5] = 0;            Om[6] = 0.5 * w[2];   Om[7] = -0.5 * w[1];
  Om[8] = 0.5 * w[1];   Om[9] = -0.5 * w[2];  Om[10] = 0;           Om[11] = 0.5 * w[0];
  Om[12] = 0.5 * w[2];  Om[13] = 0.5 * w[1];  Om[14] = -0.5 * w[0]; Om[15] = 0;
}

// the three state-dependent 3x3 blocks of the error-state linearisation at knot (x,u) -> xn
template <int NF>
QMPC_HD void srb_linearize(const QuatModel<NF>& m, const double* x, const double* u, const double* xn, double hd,
                           double hh, KnotLin& L) {
  double xd[13], qm[4], wm[3];
  m.ct_dyn(x, u, xd);
#pragma unroll
  for (int i = 0; i < 4; ++i) qm[i] = x[3 + i] + hh * xd[3 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) wm[i] = x[10 + i] + hh * xd[10 + i];
  double G[12], Gm[12], Gn[12], Om[16], Omm[16];
  quat_G(x + 3, G);
  quat_G(qm, Gm);
  quat_G(xn + 3, Gn);
  fill_omega(x + 10, Om);
  fill_omega(wm, Omm);
  double Aqq[16], Aqw[12], T1[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = 0;
#pragma unroll
      for (int l = 0; l < 4; ++l) s += Omm[4 * i + l] * Om[4 * l + j];
      Aqq[4 * i + j] = hd * (hh * s + Omm[4 * i + j]) + (i == j ? 1.0 : 0.0);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int l = 0; l < 4; ++l) s += Omm[4 * i + l] * (0.5 * G[3 * l + j]);
      Aqw[3 * i + j] = hd * (hh * s + 0.5 * Gm[3 * i + j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int l = 0; l < 4; ++l) s += Aqq[4 * i + l] * G[3 * l + j];
      T1[3 * i + j] = s;
    }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s1 += Gn[3 * i + a] * T1[3 * i + b];
        s2 += Gn[3 * i + a] * Aqw[3 * i + b];
        s3 += Gn[3 * i + a] * (0.5 * Gm[3 * i + b]);
      }
      L.Aff[3 * a + b] = s1;
      L.Afw[3 * a + b] = s2;
      L.Cf[3 * a + b] = hd * (hh * s3);
    }
}

// y = A^T v  (12)
QMPC_HD inline void srb_At_vec(const KnotLin& L, double hd, const double* v, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    y[i] = v[i];
    y[3 + i] = L.Aff[i] * v[3] + L.Aff[3 + i] * v[4] + L.Aff[6 + i] * v[5];
    y[6 + i] = hd * v[i] + v[6 + i];
    y[9 + i] = L.Afw[i] * v[3] + L.Afw[3 + i] * v[4] + L.Afw[6 + i] * v[5] + v[9 + i];
  }
}
// t = M^T v  (6)
QMPC_HD inline void srb_Mt_vec(const KnotLin& L, double hd, double hh, const double* v, double* t) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t[i] = hd * hh * v[i] + hd * v[6 + i];
    t[3 + i] = L.Cf[i] * v[3] + L.Cf[3 + i] * v[4] + L.Cf[6 + i] * v[5] + hd * v[9 + i];
  }
}
// r = W^T t  (NU) ; W[F rows] = I/m, W[tau rows c][3i+b] = IS_i[c][b]
template <int NF>
QMPC_HD inline void srb_Wt_vec(const QuatModel<NF>& m, const double* t, double* r) {
#pragma unroll
  for (int i = 0; i < NF; ++i)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      r[3 * i + b] = m.inv_mass * t[b] + m.IS[9 * i + b] * t[3] + m.IS[9 * i + 3 + b] * t[4] + m.IS[9 * i + 6 + b] * t[5];
}

// One Riccati step.  In: P,pv of knot k+1 (full 12x12 row-major / 12).  Out: P,pv of knot k, the
// gains K (NU x 12), d (NU).  Returns false if Quu is not positive definite.
template <int NF>
QMPC_HD bool srb_backward_step(const QuatModel<NF>& m, const KnotLin& L, double hd, double hh, const double* lx,
                               const double* lxx /*144*/, const double* g /*NU: lu+gu*/, const double* Dblk /*NF*9*/,
                               double* P, double* pv, double* K, double* d, double* dphi0) {
  constexpr int NU = 3 * NF;
  const double c1 = hd * hh;
  double PA[144], T[72];
  // PA = P A
#pragma unroll 2
  for (int i = 0; i < 12; ++i) {
    const double* Pi = P + 12 * i;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      PA[12 * i + j] = Pi[j];
      PA[12 * i + 3 + j] = Pi[3] * L.Aff[j] + Pi[4] * L.Aff[3 + j] + Pi[5] * L.Aff[6 + j];
      PA[12 * i + 6 + j] = hd * Pi[j] + Pi[6 + j];
      PA[12 * i + 9 + j] = Pi[3] * L.Afw[j] + Pi[4] * L.Afw[3 + j] + Pi[5] * L.Afw[6 + j] + Pi[9 + j];
    }
  }
  // T = M^T PA (6 x 12)
#pragma unroll 2
  for (int j = 0; j < 12; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      T[12 * i + j] = c1 * PA[12 * i + j] + hd * PA[12 * (6 + i) + j];
      T[12 * (3 + i) + j] = L.Cf[i] * PA[36 + j] + L.Cf[3 + i] * PA[48 + j] + L.Cf[6 + i] * PA[60 + j] +
                            hd * PA[12 * (9 + i) + j];
    }
  // S = M^T P M (6 x 6) via PM (12 x 6)
  double S[36];
  {
    double PM[72];
#pragma unroll 2
    for (int i = 0; i < 12; ++i) {
      const double* Pi = P + 12 * i;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        PM[6 * i + c] = c1 * Pi[c] + hd * Pi[6 + c];
        PM[6 * i + 3 + c] = Pi[3] * L.Cf[c] + Pi[4] * L.Cf[3 + c] + Pi[5] * L.Cf[6 + c] + hd * Pi[9 + c];
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        S[6 * i + c] = c1 * PM[6 * i + c] + hd * PM[6 * (6 + i) + c];
        S[6 * (3 + i) + c] = L.Cf[i] * PM[18 + c] + L.Cf[3 + i] * PM[24 + c] + L.Cf[6 + i] * PM[30 + c] +
                             hd * PM[6 * (9 + i) + c];
      }
  }
  double s[6], Qx[12], Qu[NU];
  srb_Mt_vec(L, hd, hh, pv, s);
  srb_At_vec(L, hd, pv, Qx);
#pragma unroll
  for (int a = 0; a < 12; ++a) Qx[a] += lx[a];
  srb_Wt_vec(m, s, Qu);
#pragma unroll
  for (int a = 0; a < NU; ++a) Qu[a] += g[a];

  // Quu = D + W^T (S W) ; Qux = W^T T
  double Quu[NU * NU], Qux[NU * 12];
  {
    double SW[6 * NU];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int b = 0; b < 3; ++b)
          SW[NU * r + 3 * f + b] = m.inv_mass * S[6 * r + b] + S[6 * r + 3] * m.IS[9 * f + b] +
                                   S[6 * r + 4] * m.IS[9 * f + 3 + b] + S[6 * r + 5] * m.IS[9 * f + 6 + b];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int j = 0; j < NU; ++j)
          Quu[NU * (3 * f + a) + j] = m.inv_mass * SW[NU * a + j] + m.IS[9 * f + a] * SW[NU * 3 + j] +
                                      m.IS[9 * f + 3 + a] * SW[NU * 4 + j] + m.IS[9 * f + 6 + a] * SW[NU * 5 + j];
#pragma unroll
        for (int j = 0; j < 12; ++j)
          Qux[12 * (3 * f + a) + j] = m.inv_mass * T[12 * a + j] + m.IS[9 * f + a] * T[36 + j] +
                                      m.IS[9 * f + 3 + a] * T[48 + j] + m.IS[9 * f + 6 + a] * T[60 + j];
#pragma unroll
        for (int b = 0; b < 3; ++b) Quu[NU * (3 * f + a) + 3 * f + b] += Dblk[9 * f + 3 * a + b];
      }
  }
  // Cholesky of Quu and the NE+1 solves (same algorithm as the reference statement)
  if (!chol<NU>(Quu)) return false;
#pragma unroll 1
  for (int c = 0; c <= 12; ++c) {
    double rhs[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) rhs[i] = c < 12 ? Qux[12 * i + c] : Qu[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double t = rhs[i];
#pragma unroll
      for (int l = 0; l < i; ++l) t -= Quu[NU * i + l] * rhs[l];
      rhs[i] = t / Quu[NU * i + i];
    }
#pragma unroll
    for (int i = NU - 1; i >= 0; --i) {
      double t = rhs[i];
#pragma unroll
      for (int l = i + 1; l < NU; ++l) t -= Quu[NU * l + i] * rhs[l];
      rhs[i] = t / Quu[NU * i + i];
    }
    if (c < 12) {
#pragma unroll
      for (int i = 0; i < NU; ++i) K[12 * i + c] = -rhs[i];
    } else {
#pragma unroll
      for (int i = 0; i < NU; ++i) d[i] = -rhs[i];
    }
  }
  {
    double t = 0;
#pragma unroll
    for (int i = 0; i < NU; ++i) t += Qu[i] * d[i];
    *dphi0 += t;
  }
  // P <- lxx + A^T PA + Qux^T K (upper triangle, mirrored) ; pv <- Qx + Qux^T d
  double Pn[144];
#pragma unroll 2
  for (int j = 0; j < 12; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Pn[12 * i + j] = PA[12 * i + j];
      Pn[12 * (3 + i) + j] = L.Aff[i] * PA[36 + j] + L.Aff[3 + i] * PA[48 + j] + L.Aff[6 + i] * PA[60 + j];
      Pn[12 * (6 + i) + j] = hd * PA[12 * i + j] + PA[12 * (6 + i) + j];
      Pn[12 * (9 + i) + j] = L.Afw[i] * PA[36 + j] + L.Afw[3 + i] * PA[48 + j] + L.Afw[6 + i] * PA[60 + j] +
                             PA[12 * (9 + i) + j];
    }
#pragma unroll 1
  for (int a = 0; a < 12; ++a) {
#pragma unroll
    for (int b = a; b < 12; ++b) {
      double t = 0, t2 = 0;
#pragma unroll
      for (int l = 0; l < NU; ++l) { t += Qux[12 * l + a] * K[12 * l + b]; t2 += Qux[12 * l + b] * K[12 * l + a]; }
      const double v = 0.5 * ((Pn[12 * a + b] + lxx[12 * a + b] + t) + (Pn[12 * b + a] + lxx[12 * b + a] + t2));
      P[12 * a + b] = v;
      P[12 * b + a] = v;
    }
    double t = 0;
#pragma unroll
    for (int l = 0; l < NU; ++l) t += Qux[12 * l + a] * d[l];
    pv[a] = Qx[a] + t;
  }
  return true;
}

// AL gradient (NU) and D blocks = diag(R) + rho J_a^T J_a per foot (NF*9)
template <int NF>
QMPC_HD void srb_al_terms(const QuatModel<NF>& m, const QmpcConfig& cfg, int k, const double* u, const GVec& mu_k, double rho,
                          double* gu, double* Dblk) {
  double Hb[9 * NF];
  al_terms(m, k, u, mu_k, rho, gu, Hb);
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int a = 0; a < 9; ++a) Dblk[9 * f + a] = Hb[9 * f + a];
    Dblk[9 * f] += cfg.r_weights[3 * f];
    Dblk[9 * f + 4] += cfg.r_weights[3 * f + 1];
    Dblk[9 * f + 8] += cfg.r_weights[3 * f + 2];
  }
}

template <int NF>
QMPC_HD void srb_solve_one(const QmpcConfig& cfg, const SolverOpts& o, const QmpcProblem* in,
                           const unsigned char* sched, QmpcWarmStart* warm, QmpcResult* out,
                           double* ws, int pid, size_t stride) {
  using M = QuatModel<NF>;
  using L = SrbLayout<NF>;
  constexpr int NX = 13, NE = 12, NU = M::NU, NC = M::NC;
  const int N = o.N;
  const float h = o.h;
  const double hd = (double)h, hh = (double)(h / 2);
  double* base = ws + pid;
  GVec X{base + L::X(N) * stride, stride}, Xn{base + L::Xn(N) * stride, stride};
  GVec U{base + L::U(N) * stride, stride}, Un{base + L::Un(N) * stride, stride};
  const GVec gmu{base + L::mu(N) * stride, stride}, gY{base + L::Y(N) * stride, stride};
  const GVec gP{base + L::P(N) * stride, stride}, gpv{base + L::pv(N) * stride, stride};
  const GVec glin{base + L::lin(N) * stride, stride}, gK{base + L::K(N) * stride, stride};
  const GVec gd{base + L::d(N) * stride, stride};

  M m;
  double x0[NX];
  {
    QmpcProblem prob = in[pid];
    m.setup(cfg, prob, sched ? sched + (size_t)pid * QMPC_MAX_HORIZON : nullptr, x0);
  }
  double rho = o.penalty_initial;
  for (int i = 0; i < N * NC; ++i) gmu[i] = 0.0;

  {
    double x[NX], xn[NX], u0[NU];
    for (int i = 0; i < NX; ++i) { x[i] = x0[i]; X[i] = x0[i]; }
    for (int i = 0; i < NU; ++i) u0[i] = m.uref_at(0, i);   // SetInput(u_traj_ref.at(0)), QuatMpc.cpp:253
    const QmpcWarmStart* wsrc = (warm && warm[pid].valid) ? warm + pid : nullptr;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (wsrc) {
        const double* wr = warm_row(wsrc, k, N);
        for (int i = 0; i < NU; ++i) u0[i] = wr[i];
      }
      st<NU>(U.off(k * NU), u0);
      mid_dyn(m, x, u0, h, xn);
      for (int i = 0; i < NX; ++i) { x[i] = xn[i]; X[(k + 1) * NX + i] = xn[i]; }
    }
  }
  double viol = 0;
  double phi = merit(m, cfg, o, X, U, gmu, rho, &viol);
  int status = QMPC_STATUS_MAX_ITERATIONS, iters = 0;
  double cost_decrease = INFINITY;
  if (!isfinite(phi)) status = QMPC_STATUS_NONFINITE;

#pragma unroll 1
  for (int it = 0; it < o.iterations_max && status == QMPC_STATUS_MAX_ITERATIONS; ++it) {
    // ---------------- linearise every knot (27 doubles each)
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      double x[NX], u[NU], xnx[NX];
      KnotLin Lk;
      ld<NX>(x, X.off(k * NX));
      ld<NU>(u, U.off(k * NU));
      ld<NX>(xnx, X.off((k + 1) * NX));
      srb_linearize(m, x, u, xnx, hd, hh, Lk);
      st<9>(glin.off(k * 27), Lk.Aff);
      st<9>(glin.off(k * 27 + 9), Lk.Afw);
      st<9>(glin.off(k * 27 + 18), Lk.Cf);
    }

    if (it > 0) {
      // ---------------- stationarity with the Riccati duals of the accepted step
      double rx = 0, ru = 0;
      {
        double x[NX], lx[NE], hphi;
        ld<NX>(x, X.off(N * NX));
        cost_expand(m, cfg, N, x, lx, &hphi);
        for (int a = 0; a < NE; ++a) {
          double v = fabs(lx[a] - gY[N * NE + a]);
          if (v > rx) rx = v;
        }
      }
#pragma unroll 1
      for (int k = 0; k < N; ++k) {
        double x[NX], u[NU], lx[NE], hphi, yn[NE], gu[NU], Hb[9 * NF], Aty[NE], t6[6], Bty[NU];
        KnotLin Lk;
        ld<NX>(x, X.off(k * NX));
        ld<NU>(u, U.off(k * NU));
        ld<NE>(yn, gY.off((k + 1) * NE));
        ld<9>(Lk.Aff, glin.off(k * 27));
        ld<9>(Lk.Afw, glin.off(k * 27 + 9));
        ld<9>(Lk.Cf, glin.off(k * 27 + 18));
        cost_expand(m, cfg, k, x, lx, &hphi);
        al_terms(m, k, u, gmu.off(k * NC), rho, gu, Hb);
        srb_At_vec(Lk, hd, yn, Aty);
        srb_Mt_vec(Lk, hd, hh, yn, t6);
        srb_Wt_vec(m, t6, Bty);
        for (int a = 0; a < NE; ++a) {
          double v = fabs(lx[a] + Aty[a] - gY[k * NE + a]);
          if (v > rx) rx = v;
        }
        for (int a = 0; a < NU; ++a) {
          double v = fabs(cfg.r_weights[a] * (u[a] - m.uref_at(k, a)) + gu[a] + Bty[a]);
          if (v > ru) ru = v;
        }
      }
      double stat = rx > ru ? rx : ru;
      if (stat < o.tol_stationarity && viol < o.tol_primal_feasibility) {
        status = QMPC_STATUS_SUCCESS;
        break;
      }
      if (fabs(cost_decrease) < o.tol_cost_intermediate || stat < o.tol_stationarity) {
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          double u[NU], c[NC];
          ld<NU>(u, U.off(k * NU));
          cone_eval(m, k, u, c);
          for (int i = 0; i < NC; ++i) {
            double est = gmu[k * NC + i] + rho * c[i];
            gmu[k * NC + i] = est > 0 ? est : 0;
          }
        }
        double r = rho * o.penalty_scaling;
        rho = r < o.penalty_max ? r : o.penalty_max;
        phi = merit(m, cfg, o, X, U, gmu, rho, &viol);
      }
    }

    // ---------------- Riccati backward pass (structured)
    double dphi0 = 0;
    bool bp_ok = true;
    {
      double P[144], pv[NE];
      {
        double x[NX], hphi;
        ld<NX>(x, X.off(N * NX));
        cost_expand(m, cfg, N, x, pv, &hphi);
        cost_hessian<M>(cfg, x, hphi, P);
        int idx = 0;
        for (int a = 0; a < NE; ++a)
          for (int b = a; b < NE; ++b) gP[N * 78 + idx++] = P[12 * a + b];
        st<NE>(gpv.off(N * NE), pv);
      }
#pragma unroll 1
      for (int k = N - 1; k >= 0; --k) {
        double x[NX], u[NU], lx[NE], hphi, lxx[144], gu[NU], Dblk[9 * NF], g[NU];
        double Kk[NU * 12], dk[NU];
        KnotLin Lk;
        ld<NX>(x, X.off(k * NX));
        ld<NU>(u, U.off(k * NU));
        ld<9>(Lk.Aff, glin.off(k * 27));
        ld<9>(Lk.Afw, glin.off(k * 27 + 9));
        ld<9>(Lk.Cf, glin.off(k * 27 + 18));
        cost_expand(m, cfg, k, x, lx, &hphi);
        cost_hessian<M>(cfg, x, hphi, lxx);
        srb_al_terms(m, cfg, k, u, gmu.off(k * NC), rho, gu, Dblk);
        for (int i = 0; i < NU; ++i) g[i] = cfg.r_weights[i] * (u[i] - m.uref_at(k, i)) + gu[i];
        if (!srb_backward_step(m, Lk, hd, hh, lx, lxx, g, Dblk, P, pv, Kk, dk, &dphi0)) {
          bp_ok = false;
          break;
        }
        st<NU * 12>(gK.off(k * NU * 12), Kk);
        st<NU>(gd.off(k * NU), dk);
        int idx = 0;
        for (int a = 0; a < NE; ++a)
          for (int b = a; b < NE; ++b) gP[k * 78 + idx++] = P[12 * a + b];
        st<NE>(gpv.off(k * NE), pv);
      }
    }
    if (!bp_ok) { status = QMPC_STATUS_BACKWARD_FAILED; break; }

    // ---------------- forward pass with back-tracking line search
    double alpha = 1.0, phin = 0, violn = 0;
    bool accepted = false;
#pragma unroll 1
    for (int ls = 0; ls < o.ls_iters_max; ++ls) {
      double x[NX], xn[NX];
      for (int i = 0; i < NX; ++i) { x[i] = x0[i]; Xn[i] = x0[i]; }
#pragma unroll 1
      for (int k = 0; k < N; ++k) {
        double xb[NX], dx[NE], u[NU];
        ld<NX>(xb, X.off(k * NX));
        state_diff<M>(x, xb, dx);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double t = 0;
#pragma unroll
          for (int l = 0; l < NE; ++l) t += gK[k * NU * 12 + i * 12 + l] * dx[l];
          u[i] = U[k * NU + i] + alpha * gd[k * NU + i] + t;
          Un[k * NU + i] = u[i];
        }
        mid_dyn(m, x, u, h, xn);
        for (int i = 0; i < NX; ++i) { x[i] = xn[i]; Xn[(k + 1) * NX + i] = xn[i]; }
      }
      phin = merit(m, cfg, o, Xn, Un, gmu, rho, &violn);
      if (isfinite(phin) && phin <= phi + o.ls_c1 * alpha * dphi0) { accepted = true; break; }
      alpha *= o.ls_decrease;
    }
    iters = it + 1;
    if (!accepted) { status = QMPC_STATUS_LINESEARCH_FAILED; break; }
    cost_decrease = phi - phin;
    // Riccati duals of the accepted step: y_k = P_k dx_k + p_k
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
      double xnw[NX], xb[NX], dx[NE], y[NE];
      ld<NX>(xnw, Xn.off(k * NX));
      ld<NX>(xb, X.off(k * NX));
      state_diff<M>(xnw, xb, dx);
      ld<NE>(y, gpv.off(k * NE));
      int idx = 0;
#pragma unroll
      for (int a = 0; a < NE; ++a)
#pragma unroll
        for (int b = a; b < NE; ++b) {
          const double v = gP[k * 78 + idx++];
          y[a] += v * dx[b];
          if (b != a) y[b] += v * dx[a];
        }
      st<NE>(gY.off(k * NE), y);
    }
    { GVec t = X; X = Xn; Xn = t; }
    { GVec t = U; U = Un; Un = t; }
    phi = phin;
    viol = violn;
  }

  QmpcResult r;
  double u0[NU];
  ld<NU>(u0, U);
  m.write_result(u0, r);
  r.max_violation = viol;
  r.iterations = iters;
  r.status = status;
  out[pid] = r;
  if (warm) {
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < 12; ++i) warm[pid].u[k][i] = i < NU ? U[k * NU + i] : 0.0;
    warm[pid].valid = status != QMPC_STATUS_NONFINITE;
  }
}

#ifdef __CUDACC__
template <int NF>
__global__ void __launch_bounds__(64)
qmpc_srb_kernel(QmpcConfig cfg, SolverOpts o, const QmpcProblem* __restrict__ in,
                const unsigned char* __restrict__ sched, QmpcWarmStart* __restrict__ warm,
                QmpcResult* __restrict__ out,
                double* __restrict__ ws, int batch, size_t stride) {
  const int pid = blockIdx.x * blockDim.x + threadIdx.x;
  if (pid >= batch) return;
  srb_solve_one<NF>(cfg, o, in, sched, warm, out, ws, pid, stride);
}
#endif

}  // namespace qmpc
