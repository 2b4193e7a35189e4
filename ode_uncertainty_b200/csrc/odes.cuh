// ODE right-hand sides (plugins) as device templates: one source serves the primal evaluation
// (T = double) and the forward-mode tangent evaluation (T = Dual<K>).
//
// State layout: the reference's x[N, D] flattened row-major (x.flatten(), used by the EKF at
// src/filters/sqrt_ekf.py:147-163); for second-order systems positions come first, then
// velocities.  Parameter layout: constructor-keyword order of the reference builder
// (`ODEBuilder.params` insertion order, src/ode/ode.py:18-23).
//
// Every struct exposes
//   NX           flattened state dimension n
//   NP           number of scalar parameters
//   defaults()   reference default parameter values
//   rhs<T,PT>(t, x, th, dx)
#pragma once
#include "dual.cuh"
#include "gdual.cuh"

namespace odeu {

// ---------------------------------------------------------------------------------------------
// Lorenz-63 (reference: src/ode/lorenz.py:30-54; params sigma, beta, rho :14-16)
struct OdeLorenz {
  static constexpr int NX = 3;
  static constexpr int NP = 3;
  template <class T, class PT>
  ODEU_HD static void rhs(double, const T* x, const PT* th, T* dx) {
    const PT& sigma = th[0];
    const PT& beta = th[1];
    const PT& rho = th[2];
    dx[0] = sigma * (x[1] - x[0]);
    dx[1] = x[0] * (rho - x[2]) - x[1];
    dx[2] = x[0] * x[1] - beta * x[2];
  }
};

// Van der Pol (reference: src/ode/van_der_pol.py:22-46; flat state [x, dx/dt])
struct OdeVanDerPol {
  static constexpr int NX = 2;
  static constexpr int NP = 1;
  template <class T, class PT>
  ODEU_HD static void rhs(double, const T* x, const PT* th, T* dx) {
    dx[0] = x[1];
    dx[1] = th[0] * (1.0 - sqr(x[0])) * x[1] - x[0];
  }
};

// Lotka-Volterra (reference: src/ode/lotka_volterra.py:31-54; params alpha, beta, gamma, delta)
struct OdeLotkaVolterra {
  static constexpr int NX = 2;
  static constexpr int NP = 4;
  template <class T, class PT>
  ODEU_HD static void rhs(double, const T* x, const PT* th, T* dx) {
    dx[0] = th[0] * x[0] - th[1] * x[0] * x[1];
    dx[1] = -th[2] * x[1] + th[3] * x[0] * x[1];
  }
};

// Pendulum (reference: src/ode/pendulum.py:22-46; `-9.81 / length * sin(x)`)
struct OdePendulum {
  static constexpr int NX = 2;
  static constexpr int NP = 1;
  template <class T, class PT>
  ODEU_HD static void rhs(double, const T* x, const PT* th, T* dx) {
    dx[0] = x[1];
    dx[1] = (-9.81 / th[0]) * d_sin(x[0]);
  }
};

// Linearly coupled anharmonic oscillators (reference: src/ode/lcao.py:35-63).  D positions then
// D velocities; the coupling term is `flip(x_prev)`, so it works for any D.
template <int D>
struct OdeLCAO {
  static constexpr int NX = 2 * D;
  static constexpr int NP = 3;
  template <class T, class PT>
  ODEU_HD static void rhs(double, const T* x, const PT* th, T* dx) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      dx[i] = x[D + i];
      dx[D + i] = -th[0] * x[i] - th[1] * cube(x[i]) - th[2] * x[D - 1 - i];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Hodgkin-Huxley (reference: src/ode/hodgkin_huxley.py; rate functions :12-27, steady states
// :29-36, gate equations :38-44, currents :46-58, models full/reduced-1/reduced-4 :125-249).
// Expressions keep the reference's parenthesisation (notably tau_u); divisions by
// literals are kept as divisions so the primal matches an IEEE evaluation of the same formula.
// Parameter order: C, A, g_Na, E_Na, g_K, E_K, g_leak, E_leak, V_T, g_M, tau_max, g_L, E_Ca,
// g_T, V_x (constructor order :64-81).
namespace hh {
enum { P_C = 0, P_A, P_gNa, P_ENa, P_gK, P_EK, P_gleak, P_Eleak, P_VT, P_gM, P_taumax, P_gL,
       P_ECa, P_gT, P_Vx, NPAR };

template <class T> ODEU_HD T a_m(const T& u) { return -0.32 * (u - 13.0) / (d_exp(-(u - 13.0) / 4.0) - 1.0); }
template <class T> ODEU_HD T b_m(const T& u) { return 0.28 * (u - 40.0) / (d_exp((u - 40.0) / 5.0) - 1.0); }
template <class T> ODEU_HD T a_n(const T& u) { return -0.032 * (u - 15.0) / (d_exp(-(u - 15.0) / 5.0) - 1.0); }
template <class T> ODEU_HD T b_n(const T& u) { return 0.5 * d_exp(-(u - 10.0) / 40.0); }
template <class T> ODEU_HD T a_h(const T& u) { return 0.128 * d_exp(-(u - 17.0) / 18.0); }
template <class T> ODEU_HD T b_h(const T& u) { return 4.0 / (1.0 + d_exp(-(u - 40.0) / 5.0)); }
template <class T> ODEU_HD T a_q(const T& V) { return 0.055 * (-27.0 - V) / (d_exp((-27.0 - V) / 3.8) - 1.0); }
template <class T> ODEU_HD T b_q(const T& V) { return 0.94 * d_exp((-75.0 - V) / 17.0); }
template <class T> ODEU_HD T a_r(const T& V) { return 0.000457 * d_exp((-13.0 - V) / 50.0); }
template <class T> ODEU_HD T b_r(const T& V) { return 0.0065 / (d_exp((-15.0 - V) / 28.0) + 1.0); }
template <class T> ODEU_HD T p_inf(const T& V) { return 1.0 / (1.0 + d_exp(-(V + 35.0) / 10.0)); }

// stimulus window (reference :52): 210e-6 for 10 <= t <= 90
ODEU_HD double I_in(double t) { return (t >= 10.0 && t <= 90.0) ? 210.0 * 1e-6 : 0.0; }

// MODEL: 0 = full (8 states), 1 = reduced-1 (7), 4 = reduced-4 (4)
template <int MODEL> struct dim;
template <> struct dim<0> { static constexpr int value = 8; };
template <> struct dim<1> { static constexpr int value = 7; };
template <> struct dim<4> { static constexpr int value = 4; };

// One compartment.  `th` is indexed through `stride` so the multi-compartment layout
// (param-major, compartment-minor) can reuse it.
template <int MODEL, class T, class PT>
ODEU_HD void rhs_single(double t, const T* x, const PT* th, int stride, T* dx) {
  const T& V = x[0];
  const PT& V_T = th[P_VT * stride];
  const T u = V - V_T;  // V - V_T, shared by the m/h/n rates ("V - V_T - 13.0" etc.)
  dx[1] = a_m(u) * (1.0 - x[1]) - b_m(u) * x[1];
  dx[2] = a_h(u) * (1.0 - x[2]) - b_h(u) * x[2];
  dx[3] = a_n(u) * (1.0 - x[3]) - b_n(u) * x[3];
  T I = th[P_gNa * stride] * cube(x[1]) * x[2] * (th[P_ENa * stride] - V);
  I = I + th[P_gK * stride] * pow4(x[3]) * (th[P_EK * stride] - V);
  I = I + th[P_gleak * stride] * (th[P_Eleak * stride] - V);
  if (MODEL != 4) {
    const T tau_p = th[P_taumax * stride] /
                    (3.3 * d_exp((V + 35.0) / 20.0) + d_exp(-(V + 35.0) / 20.0));
    dx[4] = (p_inf(V) - x[4]) / tau_p;
    dx[5] = a_q(V) * (1.0 - x[5]) - b_q(V) * x[5];
    dx[6] = a_r(V) * (1.0 - x[6]) - b_r(V) * x[6];
    I = I + th[P_gM * stride] * x[4] * (th[P_EK * stride] - V);
    I = I + th[P_gL * stride] * sqr(x[5]) * x[6] * (th[P_ECa * stride] - V);
  }
  if (MODEL == 0) {
    const PT& V_x = th[P_Vx * stride];
    const T w = V + V_x;
    const T tau_u = (30.8 + (211.4 + d_exp((w + 113.2) / 5.0))) /
                    (3.7 * (1.0 + d_exp((w + 84.0) / 3.2)));
    const T u_inf = 1.0 / (1.0 + d_exp((w + 81.0) / 4.0));
    const T s_inf = 1.0 / (1.0 + d_exp(-(w + 57.0) / 6.2));
    dx[7] = (u_inf - x[7]) / tau_u;
    I = I + th[P_gT * stride] * sqr(s_inf) * x[7] * (th[P_ECa * stride] - V);
  }
  // f_V (reference :53-58): (sum of currents + I_in / A) / C
  dx[0] = (I + I_in(t) / th[P_A * stride]) / th[P_C * stride];
}
// ---------------------------------------------------------------------------------------------
// Row interface (used by the row-parallel kernel, ekf_rows.cuh): ONE equation of the system
// together with its non-zero partial derivatives.  Every gate equation depends on two states
// only (the membrane potential and the gate itself), the voltage equation on the states of its
// own compartment (+ the neighbours' potentials), so the Jacobian of the right-hand side has
// DIM + 2 (DIM - 1) (+ coupling) entries per compartment instead of DIM^2, and the 2 `exp` of a
// gate are evaluated once per stage instead of once per tangent column.
// The rate functions above are reused with a one-lane dual in V; S is the scalar type of the
// filter (double, or value + parameter direction for the gradient kernel).
template <class S> struct ParSingle {
  const S* th; int st;
  ODEU_HD const S& operator()(int k) const { return th[k * st]; }
};
template <class S, int NC> struct ParMulti {   // flat layout of OdeMultiHH (see below)
  const S* th; int st; int g;
  ODEU_HD const S& operator()(int k) const {
    return k == P_C ? th[(NC - 1) * st] : th[(NC + (k - 1) * NC + g) * st];
  }
};
template <class S> ODEU_HD GDual<S, 1> seed1(const S& v) {
  GDual<S, 1> r; r.v = v; r.d[0] = S(1.0); return r;
}

// gate q (1 = m, 2 = h, 3 = n, 4 = p, 5 = q, 6 = r, 7 = u): f = d x_q / dt, dV = df/dV, dX = df/dx_q
template <int MODEL, class S, class PA>
ODEU_HD void gate_row(int q, const S& V, const S& xq, const PA& par, S& f, S& dV, S& dX) {
  using D1 = GDual<S, 1>;
  D1 al, be;
  bool rate_form = true;
  switch (q) {
    case 1: { const D1 u = seed1<S>(V - par(P_VT)); al = a_m(u); be = b_m(u); break; }
    case 2: { const D1 u = seed1<S>(V - par(P_VT)); al = a_h(u); be = b_h(u); break; }
    case 3: { const D1 u = seed1<S>(V - par(P_VT)); al = a_n(u); be = b_n(u); break; }
    case 5: { const D1 v = seed1<S>(V); al = a_q(v); be = b_q(v); break; }
    case 6: { const D1 v = seed1<S>(V); al = a_r(v); be = b_r(v); break; }
    case 4: {   // dp/dt = (p_inf(V) - p) / tau_p(V)
      const D1 v = seed1<S>(V);
      const D1 tau = par(P_taumax) / (3.3 * d_exp((v + 35.0) / 20.0) + d_exp(-(v + 35.0) / 20.0));
      const D1 fd = (p_inf(v) - xq) / tau;
      f = fd.v; dV = fd.d[0]; dX = -(1.0 / tau.v);
      rate_form = false;
      break;
    }
    default: {  // 7: du/dt = (u_inf(V + V_x) - u) / tau_u(V + V_x)   (full model only)
      const D1 w = seed1<S>(V + par(P_Vx));
      const D1 tau = (30.8 + (211.4 + d_exp((w + 113.2) / 5.0))) / (3.7 * (1.0 + d_exp((w + 84.0) / 3.2)));
      const D1 uinf = 1.0 / (1.0 + d_exp((w + 81.0) / 4.0));
      const D1 fd = (uinf - xq) / tau;
      f = fd.v; dV = fd.d[0]; dX = -(1.0 / tau.v);
      rate_form = false;
      break;
    }
  }
  if (rate_form) {   // dx/dt = alpha (1 - x) - beta x
    const S om = 1.0 - xq;
    f = al.v * om - be.v * xq;
    dV = al.d[0] * om - be.d[0] * xq;
    dX = -(al.v + be.v);
  }
}

// voltage equation of one compartment: xc[k * xst] = state k of the compartment;
// df[k] = d f_V / d x_k, k < dim<MODEL>
template <int MODEL, class S, class PA>
ODEU_HD void v_row(double t, const S* xc, int xst, const PA& par, S& f, S* df) {
  const S& V = xc[0];
  const S& m = xc[xst];
  const S& h = xc[2 * xst];
  const S& nn = xc[3 * xst];
  const S& gNa = par(P_gNa);
  const S& gK = par(P_gK);
  const S& gl = par(P_gleak);
  const S dNa = par(P_ENa) - V;
  const S dK = par(P_EK) - V;
  const S m2 = m * m;
  const S t1 = gNa * (m2 * m);
  const S t2 = t1 * h;
  S I = t2 * dNa;
  S dv = -t2;
  df[1] = (3.0 * (gNa * m2)) * h * dNa;
  df[2] = t1 * dNa;
  const S n2 = nn * nn;
  const S u1 = gK * (n2 * n2);
  I = I + u1 * dK;
  dv = dv - u1;
  df[3] = (4.0 * (gK * (n2 * nn))) * dK;
  I = I + gl * (par(P_Eleak) - V);
  dv = dv - gl;
  if (MODEL != 4) {
    const S& p = xc[4 * xst];
    const S& qq = xc[5 * xst];
    const S& rr = xc[6 * xst];
    const S& gM = par(P_gM);
    const S& gL = par(P_gL);
    const S dCa = par(P_ECa) - V;
    const S w1 = gM * p;
    I = I + w1 * dK;
    dv = dv - w1;
    df[4] = gM * dK;
    const S v1 = gL * (qq * qq);
    const S v2 = v1 * rr;
    I = I + v2 * dCa;
    dv = dv - v2;
    df[5] = (2.0 * (gL * qq)) * rr * dCa;
    df[6] = v1 * dCa;
    if (MODEL == 0) {
      const S& uu = xc[7 * xst];
      const S& gT = par(P_gT);
      const GDual<S, 1> w = seed1<S>(V + par(P_Vx));
      const GDual<S, 1> si = 1.0 / (1.0 + d_exp(-(w + 57.0) / 6.2));
      const S z1 = gT * (si.v * si.v);
      const S z2 = z1 * uu;
      I = I + z2 * dCa;
      df[7] = z1 * dCa;
      dv = dv + ((gT * (2.0 * (si.v * si.d[0]))) * uu * dCa - z2);
    }
  }
  const S iC = 1.0 / par(P_C);
  f = (I + I_in(t) / par(P_A)) * iC;
  df[0] = dv * iC;
#pragma unroll
  for (int k = 1; k < dim<MODEL>::value; ++k) df[k] = df[k] * iC;
}
}  // namespace hh

template <int MODEL>
struct OdeHodgkinHuxley {
  static constexpr int NX = hh::dim<MODEL>::value;
  static constexpr int NP = hh::NPAR;
  template <class T, class PT>
  ODEU_HD static void rhs(double t, const T* x, const PT* th, T* dx) {
    hh::rhs_single<MODEL>(t, x, th, 1, dx);
  }
  // ---- row interface (ekf_rows.cuh): row r = g * ROW_CLASSES + q
  static constexpr int ROW_CLASSES = NX;
  static constexpr int ROW_GROUPS = 1;
  static constexpr int VN = NX + (NX & 1);            // voltage-row slots, padded: row offsets stay even
  static constexpr int NNZ = VN + 2 * (NX - 1);
  ODEU_HD static constexpr int row_ndep(int q) { return q == 0 ? NX : 2; }
  ODEU_HD static constexpr int row_off(int, int q) { return q == 0 ? 0 : VN + 2 * (q - 1); }
  ODEU_HD static constexpr int row_dep(int, int q, int k) { return q == 0 ? k : (k == 0 ? 0 : q); }
  template <class S>
  ODEU_HD static void row(int q, int, double t, const S* xs, int xst, const S* th, int tst, S& f, S* df) {
    const hh::ParSingle<S> par{th, tst};
    if (q == 0) hh::v_row<MODEL>(t, xs, xst, par, f, df);
    else hh::gate_row<MODEL>(q, xs[0], xs[q * xst], par, f, df[0], df[1]);
  }
};

// Multi-compartment chain (reference: src/ode/hodgkin_huxley.py:358-401).  Flat parameter
// layout = reference dict order: coupling_coeffs[NC-1], C[1], then A, g_Na, ..., V_x each [NC].
// The membrane capacitance C is shared (shape [1], broadcast, :393-396).
template <int MODEL, int NC>
struct OdeMultiHH {
  static constexpr int DIM = hh::dim<MODEL>::value;
  static constexpr int NX = NC * DIM;
  static constexpr int NP = (NC - 1) + 1 + 14 * NC;
  template <class T, class PT>
  ODEU_HD static void rhs(double t, const T* x, const PT* th, T* dx) {
    const PT* cc = th;            // [NC-1]
    const PT& Cm = th[NC - 1];    // shared capacitance
    const PT* per = th + NC;      // 14 arrays of [NC], param-major
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      // local parameter view in single-compartment order: slot 0 = C, slots 1.. = per-comp
      PT loc[hh::NPAR];
      loc[hh::P_C] = Cm;
#pragma unroll
      for (int k = 1; k < hh::NPAR; ++k) loc[k] = per[(k - 1) * NC + c];
      hh::rhs_single<MODEL>(t, x + c * DIM, loc, 1, dx + c * DIM);
    }
    // tridiagonal voltage coupling G V / C (reference :374-383, 396)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T acc = T(0.0);
      bool any = false;
      if (c > 0) { acc = cc[c - 1] * (x[(c - 1) * DIM] - x[c * DIM]); any = true; }
      if (c + 1 < NC) {
        T r = cc[c] * (x[(c + 1) * DIM] - x[c * DIM]);
        acc = any ? acc + r : r;
      }
      dx[c * DIM] = dx[c * DIM] + acc / Cm;
    }
  }
  // ---- row interface (ekf_rows.cuh): row r = g * DIM + q, g = compartment, q = equation.
  // The voltage row depends on the DIM states of its compartment and on the potentials of the
  // neighbouring compartments (NBR slots; a chain end points at its own potential with a zero
  // coefficient, like the zero entries of the reference's dense `G @ V`, :374-383).
  static constexpr int ROW_CLASSES = DIM;
  static constexpr int ROW_GROUPS = NC;
  static constexpr int NBR = NC == 1 ? 0 : (NC == 2 ? 1 : 2);
  static constexpr int VN = (DIM + NBR + 1) / 2 * 2;  // voltage-row slots, padded: row offsets stay even
  static constexpr int PER = VN + 2 * (DIM - 1);
  static constexpr int NNZ = NC * PER;
  ODEU_HD static constexpr int row_ndep(int q) { return q == 0 ? DIM + NBR : 2; }
  ODEU_HD static constexpr int row_off(int g, int q) {
    return g * PER + (q == 0 ? 0 : VN + 2 * (q - 1));
  }
  ODEU_HD static constexpr int nbr_comp(int g, int s) {   // s-th neighbour slot of compartment g
    return NC == 2 ? 1 - g : (s == 0 ? (g > 0 ? g - 1 : g) : (g + 1 < NC ? g + 1 : g));
  }
  ODEU_HD static constexpr int row_dep(int g, int q, int k) {
    return q == 0 ? (k < DIM ? g * DIM + k : nbr_comp(g, k - DIM) * DIM) : (k == 0 ? g * DIM : g * DIM + q);
  }
  template <class S>
  ODEU_HD static void row(int q, int g, double t, const S* xs, int xst, const S* th, int tst, S& f, S* df) {
    const hh::ParMulti<S, NC> par{th, tst, g};
    const S* xc = xs + (long long)g * DIM * xst;
    if (q != 0) {
      hh::gate_row<MODEL>(q, xc[0], xc[q * xst], par, f, df[0], df[1]);
      return;
    }
    hh::v_row<MODEL>(t, xc, xst, par, f, df);
    const S iC = 1.0 / par(hh::P_C);
#pragma unroll
    for (int s = 0; s < NBR; ++s) {
      const int nb = nbr_comp(g, s);
      // coupling coefficient between g and nb (cc[min(g, nb)]); none at a chain end
      const bool real = nb != g;
      const S cc = real ? th[(nb < g ? nb : g) * tst] : S(0.0);
      f = f + cc * (xs[(long long)nb * DIM * xst] - xc[0]) * iC;
      const S e = cc * iC;
      df[0] = df[0] - e;
      df[DIM + s] = e;
    }
  }
};

}  // namespace odeu
