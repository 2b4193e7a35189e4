"""ORACLE (test infrastructure, never shipped, never on the product path).

Instrumented op-counter ("commit an instrumented op-counter in the oracle that
reproduces [the algorithmic flop constants] within +-10 %").  One EKF-RK step of the reference
algorithm (`src/solvers/rksolver.py:113-155`, `src/utils.py:72-79`, `src/filters/sqrt_ekf.py:92-197,
337-376`, `src/utils.py:109-128`) is executed on scalars that count every arithmetic operation,
in the cheapest standard formulation that yields the same outputs (full covariance, n tangent
columns T = J P_sqrt, P = T T^T symmetric + diagonal Q, Cholesky-based measurement update with
a Joseph-form covariance) so the figure cannot be inflated by the square-root/QR route.

Counting convention: `+ - * / exp sin log sqrt` = 1 each (an FMA is therefore 2);
negation, abs, comparisons, selects and multiplications by a structural zero = 0.
Forward-mode tangents are counted per column with the textbook dual-number rules
(d(ab) = da b + a db = 3 ops, constant * dual = 1 op, d exp(a) = exp(a) da = 1 op, ...).

The constants bench.py and tools/bench_c3.py divide by are checked against these counts in
tests/test_opcount.py.  Run `python -m oracle.opcount` for the table.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


class Counter:
    def __init__(self):
        self.primal = 0
        self.tangent = 0
        self.exp = 0

    @property
    def total(self):
        return self.primal + self.tangent

    def snapshot(self):
        return (self.primal, self.tangent, self.exp)


CNT = Counter()


def _is_num(x):
    return isinstance(x, Num)


class Num:
    """Counted scalar with K forward-mode tangent lanes (`t is None`: a constant w.r.t. the seeds)."""
    __slots__ = ("v", "t")

    def __init__(self, v: float, t: Optional[np.ndarray] = None):
        self.v = float(v)
        self.t = t

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def lift(x) -> "Num":
        return x if _is_num(x) else Num(x)

    @staticmethod
    def _tan(k: int, lanes: np.ndarray):
        CNT.tangent += k * lanes.shape[0]

    # -- arithmetic ----------------------------------------------------------------------------
    def __neg__(self):
        return Num(-self.v, None if self.t is None else -self.t)

    def __abs__(self):
        return Num(abs(self.v), None if self.t is None else np.sign(self.v) * self.t)

    def __add__(self, o):
        if not _is_num(o):
            if o == 0.0:
                return self
            CNT.primal += 1
            return Num(self.v + o, self.t)
        CNT.primal += 1
        if self.t is None or o.t is None:
            return Num(self.v + o.v, self.t if o.t is None else o.t)
        self._tan(1, self.t)
        return Num(self.v + o.v, self.t + o.t)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-o if _is_num(o) else -float(o))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if not _is_num(o):
            if o == 0.0:
                return Num(0.0)
            if o == 1.0:
                return self
            CNT.primal += 1
            if self.t is not None:
                self._tan(1, self.t)
            return Num(self.v * o, None if self.t is None else self.t * o)
        CNT.primal += 1
        if self.t is None and o.t is None:
            return Num(self.v * o.v)
        if self.t is None or o.t is None:
            a, b = (self, o) if o.t is None else (o, self)      # a dual, b constant
            self._tan(1, a.t)
            return Num(a.v * b.v, a.t * b.v)
        self._tan(3, self.t)
        return Num(self.v * o.v, self.t * o.v + self.v * o.t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if not _is_num(o):
            CNT.primal += 1
            if self.t is not None:
                self._tan(1, self.t)
            return Num(self.v / o, None if self.t is None else self.t / o)
        CNT.primal += 1
        q = self.v / o.v
        if self.t is None and o.t is None:
            return Num(q)
        if o.t is None:
            self._tan(1, self.t)
            return Num(q, self.t / o.v)
        if self.t is None:
            self._tan(2, o.t)                                    # -(q / b) db
            return Num(q, -(q / o.v) * o.t)
        self._tan(3, self.t)                                     # (da - q db) / b
        return Num(q, (self.t - q * o.t) / o.v)

    def __rtruediv__(self, o):
        return Num(o) / self

    def __pow__(self, k: int):
        assert isinstance(k, int) and k >= 1
        r = self
        for _ in range(k - 1):
            r = r * self
        return r

    def __float__(self):
        return self.v


def _unary(fn, dfn):
    def f(x):
        x = Num.lift(x)
        CNT.primal += 1
        v = fn(x.v)
        if x.t is None:
            return Num(v)
        Num._tan(1, x.t)
        return Num(v, dfn(x.v, v) * x.t)
    return f


def exp(x):
    CNT.exp += 1
    return _exp(x)


_exp = _unary(math.exp, lambda a, v: v)
sin = _unary(math.sin, lambda a, v: math.cos(a))      # cos counted inside the tangent op
log = _unary(math.log, lambda a, v: 1.0 / a)
sqrt = _unary(math.sqrt, lambda a, v: 0.5 / v)


# --------------------------------------------------------------------------------------------------
# Tableaux (src/solvers/rkf45.py:10-34, dopri65.py:10-72, bs32.py:10-32, heun_euler.py:10-30)
def tableau(name: str):
    from oracle.ref_torch import TABLEAUX        # oracle -> oracle import only
    A, b, c = TABLEAUX[name]
    return A.numpy(), b.numpy(), c.numpy()


# --------------------------------------------------------------------------------------------------
# Right-hand sides on counted scalars: (t, x list, params dict) -> list
def rhs_lorenz(t, x, p):  # src/ode/lorenz.py:46-52
    return [p["sigma"] * (x[1] - x[0]), x[0] * (p["rho"] - x[2]) - x[1], x[0] * x[1] - p["beta"] * x[2]]


def rhs_van_der_pol(t, x, p):  # src/ode/van_der_pol.py:38-44
    return [x[1], p["damping"] * (1.0 - x[0] ** 2) * x[1] - x[0]]


def rhs_lotka_volterra(t, x, p):  # src/ode/lotka_volterra.py:47-52
    return [p["alpha"] * x[0] - p["beta"] * x[0] * x[1], -p["gamma"] * x[1] + p["delta"] * x[0] * x[1]]


# Hodgkin-Huxley reduced-1 (src/ode/hodgkin_huxley.py:12-58, 125-249); rate expressions with the
# reference's parenthesisation
def _hh_reduced1(t, x, p, coupling):
    V, m, h, n, pp, q, r = x
    VT = p["V_T"]
    a_m = -0.32 * (V - VT - 13.0) / (exp(-(V - VT - 13.0) / 4.0) - 1.0)
    b_m = 0.28 * (V - VT - 40.0) / (exp((V - VT - 40.0) / 5.0) - 1.0)
    a_n = -0.032 * (V - VT - 15.0) / (exp(-(V - VT - 15.0) / 5.0) - 1.0)
    b_n = 0.5 * exp(-(V - VT - 10.0) / 40.0)
    a_h = 0.128 * exp(-(V - VT - 17.0) / 18.0)
    b_h = 4.0 / (1.0 + exp(-(V - VT - 40.0) / 5.0))
    a_q = 0.055 * (-27.0 - V) / (exp((-27.0 - V) / 3.8) - 1.0)
    b_q = 0.94 * exp((-75.0 - V) / 17.0)
    a_r = 0.000457 * exp((-13.0 - V) / 50.0)
    b_r = 0.0065 / (exp((-15.0 - V) / 28.0) + 1.0)
    p_inf = 1.0 / (1.0 + exp(-(V + 35.0) / 10.0))
    tau_p = p["tau_max"] / (3.3 * exp((V + 35.0) / 20.0) + exp(-(V + 35.0) / 20.0))
    dm = a_m * (1.0 - m) - b_m * m
    dh = a_h * (1.0 - h) - b_h * h
    dn = a_n * (1.0 - n) - b_n * n
    dp = (p_inf - pp) / tau_p
    dq = a_q * (1.0 - q) - b_q * q
    dr = a_r * (1.0 - r) - b_r * r
    I_Na = p["g_Na"] * m ** 3 * h * (p["E_Na"] - V)
    I_K = p["g_K"] * n ** 4 * (p["E_K"] - V)
    I_leak = p["g_leak"] * (p["E_leak"] - V)
    I_M = p["g_M"] * pp * (p["E_K"] - V)
    I_L = p["g_L"] * q ** 2 * r * (p["E_Ca"] - V)
    I_in = 210.0e-6 if 10.0 <= t <= 90.0 else 0.0
    dV = (I_Na + I_K + I_leak + I_M + I_L + I_in / p["A"]) / p["C"]
    if coupling is not None:
        dV = dV + coupling / p["C"]
    return [dV, dm, dh, dn, dp, dq, dr]


def rhs_hh_r1(t, x, p):
    return _hh_reduced1(t, x, p, None)


def make_rhs_multi_hh_r1(C: int) -> Callable:  # src/ode/hodgkin_huxley.py:358-401
    def rhs(t, x, p):
        cc = p["coupling_coeffs"]
        V = [x[7 * c] for c in range(C)]
        out = []
        for c in range(C):
            coup = 0.0
            if c > 0:
                coup = coup + cc[c - 1] * (V[c - 1] - V[c])      # tridiagonal G V (rows sum to 0)
            if c + 1 < C:
                coup = coup + cc[c] * (V[c + 1] - V[c])
            pc = {k: (v[c] if isinstance(v, (list, tuple)) else v) for k, v in p.items() if k != "coupling_coeffs"}
            out += _hh_reduced1(t, x[7 * c:7 * c + 7], pc, coup)
        return out
    return rhs


HH_DEFAULTS = dict(C=1.0, A=8.3e-5, g_Na=25.0, E_Na=53.0, g_K=7.0, E_K=-107.0, g_leak=0.1,
                   E_leak=-70.0, V_T=-60.0, g_M=0.01, tau_max=4e3, g_L=0.01, E_Ca=120.0)


# --------------------------------------------------------------------------------------------------
def rk_step(rhs, params, tab, h: float, t: float, x: List[Num]):
    """x+ (row b[1] propagates) and eps = |x+(b0) - x+(b1)|; structural zeros of the tableau skipped;
    the products h*a_ij, h*b_ij are pre-scaled constants."""
    A, b, c = tab
    S, n = len(c), len(x)
    ks: List[List[Num]] = []
    for i in range(S):
        xi = []
        for d in range(n):
            acc = x[d]
            for j in range(i):
                if A[i, j] != 0.0:
                    acc = acc + ks[j][d] * float(h * A[i, j])
            xi.append(acc)
        ks.append(rhs(t + h * c[i], xi, params))
    xn, eps = [], []
    for d in range(n):
        acc = x[d]
        for j in range(S):
            if b[1, j] != 0.0:
                acc = acc + ks[j][d] * float(h * b[1, j])
        xn.append(acc)
        e = Num(0.0)
        first = True
        for j in range(S):
            dj = float(h * (b[0, j] - b[1, j]))
            if dj != 0.0:
                if first:
                    tmp = ks[j][d]
                    e = Num(tmp.v * dj)          # first term is a product only, tangent-free:
                    CNT.primal += 1              # eps carries no tangent (stop_gradient-free but unused)
                    first = False
                else:
                    tmpv = ks[j][d].v
                    e = Num(e.v + tmpv * dj)
                    CNT.primal += 2
        eps.append(abs(e))
    return xn, eps


def _mat(rows, cols, fill=0.0):
    return [[Num(fill) for _ in range(cols)] for _ in range(rows)]


def count_step(rhs, params, tableau_name: str, x0: Sequence[float], h: float = 0.01, t: float = 20.0,
               H: Optional[np.ndarray] = None, R: Optional[np.ndarray] = None, seed: int = 7) -> Dict[str, int]:
    """Counts one predict (+ correct + NLL term if H is given).  Returns the per-phase counts."""
    global CNT
    rng = np.random.default_rng(seed)
    n = len(x0)
    tab = tableau(tableau_name)
    P_sqrt = np.tril(rng.normal(size=(n, n))) * 1e-3 + 1e-3 * np.eye(n)
    out: Dict[str, int] = {}

    # ---- F_rk: primal stages, x+, eps ---------------------------------------------------------------
    CNT = Counter()
    rk_step(rhs, params, tab, h, t, [Num(v) for v in x0])
    out["rk"] = CNT.total
    # c_f alone
    CNT = Counter()
    rhs(t, [Num(v) for v in x0], params)
    out["c_f"], out["exp_per_rhs"] = CNT.primal, CNT.exp

    # ---- F_tan: T = J P_sqrt as n dense tangent columns (src/utils.py:72-79) ----------------------------
    CNT = Counter()
    xs = [Num(x0[d], P_sqrt[d, :].copy()) for d in range(n)]
    xn, eps = rk_step(rhs, params, tab, h, t, xs)
    out["tan"] = CNT.tangent
    CNT = Counter()
    rhs(t, [Num(x0[d], P_sqrt[d, :1].copy()) for d in range(n)], params)
    out["c_J"] = CNT.tangent
    T = np.stack([xn[d].t if xn[d].t is not None else np.zeros(n) for d in range(n)])

    # ---- F_cov: P = T T^T (lower triangle) + diag((s eps)^2), Diagonal cov-update -------------------------
    CNT = Counter()
    Tn = [[Num(T[i, k]) for k in range(n)] for i in range(n)]
    P = _mat(n, n)
    for i in range(n):
        for j in range(i + 1):
            acc = Tn[i][0] * Tn[j][0]
            for k in range(1, n):
                acc = acc + Tn[i][k] * Tn[j][k]
            P[i][j] = P[j][i] = acc
    for i in range(n):
        e = Num(eps[i].v)
        P[i][i] = P[i][i] + e * e
    out["cov"] = CNT.total

    # ---- F_corr: dense H, Cholesky of S, gain, Joseph update, NLL (sqrt_ekf.py:337-376, utils.py:109-128) ---
    if H is not None:
        L = H.shape[0]
        CNT = Counter()
        Hn = [[Num(H[l, k]) for k in range(n)] for l in range(L)]
        y = [Num(v) for v in rng.normal(size=L)]
        xv = [Num(v.v) for v in xn]

        def dot(a, bb):
            acc = a[0] * bb[0]
            for k in range(1, len(a)):
                acc = acc + a[k] * bb[k]
            return acc

        def dense(v):                      # counts even numerically-zero entries of a DENSE operand
            return Num(v.v if v.v != 0.0 else 1e-300)

        Hd = [[dense(v) for v in row] for row in Hn]
        W = [[dot(Hd[l], [P[k][j] for k in range(n)]) for j in range(n)] for l in range(L)]   # H P
        Sm = _mat(L, L)
        for a in range(L):
            for bb in range(a + 1):
                Sm[a][bb] = Sm[bb][a] = dot(W[a], Hd[bb]) + float(R[a, bb])
        Ls = _mat(L, L)                    # Cholesky
        for j in range(L):
            acc = Sm[j][j]
            for k in range(j):
                acc = acc - Ls[j][k] * Ls[j][k]
            Ls[j][j] = sqrt(acc)
            for i in range(j + 1, L):
                acc = Sm[i][j]
                for k in range(j):
                    acc = acc - Ls[i][k] * Ls[j][k]
                Ls[i][j] = acc / Ls[j][j]
        K = _mat(n, L)                     # K = W^T S^-1: forward + backward substitution per state row
        for i in range(n):
            z = [None] * L
            for a in range(L):
                acc = W[a][i]
                for k in range(a):
                    acc = acc - Ls[a][k] * z[k]
                z[a] = acc / Ls[a][a]
            for a in reversed(range(L)):
                acc = z[a]
                for k in range(a + 1, L):
                    acc = acc - Ls[k][a] * K[i][k]
                K[i][a] = acc / Ls[a][a]
        d = [y[l] - dot(Hd[l], xv) for l in range(L)]             # innovation
        for i in range(n):
            xv[i] = xv[i] + dot(K[i], d)
        # Joseph: P+ = (I-KH) P (I-KH)^T + K R K^T = P - K W - (K W)^T + K (S) K^T, lower triangle
        KS = [[dot(K[i], [Sm[k][a] for k in range(L)]) for a in range(L)] for i in range(n)]
        G = [[KS[i][a] - W[a][i] for a in range(L)] for i in range(n)]     # K S - P H^T
        for i in range(n):
            for j in range(i + 1):
                acc = P[i][j]
                for a in range(L):
                    acc = acc + G[i][a] * K[j][a]
                    acc = acc - K[i][a] * W[a][j]
                P[i][j] = P[j][i] = acc
        # NLL term: 0.5 |Ls^-1 d|^2 + L/2 log 2pi + sum log Ls_aa
        z = [None] * L
        for a in range(L):
            acc = d[a]
            for k in range(a):
                acc = acc - Ls[a][k] * z[k]
            z[a] = acc / Ls[a][a]
        nl = dot(z, z) * 0.5 + (0.5 * L * math.log(2 * math.pi))
        for a in range(L):
            nl = nl + log(Ls[a][a])
        out["corr"] = CNT.total
    out["predict"] = out["rk"] + out["tan"] + out["cov"]
    out["step"] = out["predict"] + out.get("corr", 0)
    return out


def count_grad_step(step_counts: Dict[str, int], p: int) -> int:
    """SURVEY 8(d) upper-bound model for the forward-mode gradient over p parameters:
    F_grad = p (2 F_tan + F_rk + 2 F_cov + 2 F_corr), on the COUNTED phases."""
    c = step_counts
    return c["step"] + p * (2 * c["tan"] + c["rk"] + 2 * c["cov"] + 2 * c.get("corr", 0))


# The configurations of SURVEY 8(d) ------------------------------------------------------------------
def configs() -> Dict[str, Dict[str, int]]:
    res = {}
    res["lorenz"] = count_step(rhs_lorenz, dict(sigma=10.0, beta=8.0 / 3.0, rho=28.0), "RKF45",
                               [1.3, -2.0, 21.0], H=np.eye(3), R=1e-3 * np.eye(3))
    res["van_der_pol"] = count_step(rhs_van_der_pol, dict(damping=5.0), "RKF45", [1.7, 0.4],
                                    H=np.eye(2), R=1e-3 * np.eye(2))
    x_hh = [-65.0, 0.05, 0.6, 0.3, 0.05, 0.01, 0.3]
    H1 = np.zeros((1, 7)); H1[0, 0] = 1.0
    res["hh_1comp_r1"] = count_step(rhs_hh_r1, HH_DEFAULTS, "RKF45", x_hh, H=H1, R=0.1 * np.eye(1))
    p2 = {k: [v, v] for k, v in HH_DEFAULTS.items() if k != "C"}
    p2["C"] = 1.0
    p2["coupling_coeffs"] = [0.5]
    H2 = np.zeros((2, 14)); H2[0, 0] = 1.0; H2[1, 7] = 1.0
    res["hh_2comp_r1"] = count_step(make_rhs_multi_hh_r1(2), p2, "RKF45", x_hh + x_hh, H=H2, R=0.1 * np.eye(2))
    res["hh_1comp_r1"]["grad"] = count_grad_step(res["hh_1comp_r1"], 9)
    res["hh_2comp_r1"]["grad"] = count_grad_step(res["hh_2comp_r1"], 12)
    return res


# the constants bench.py / tools/bench_c3.py divide by
SURVEY_CONSTANTS = {
    "lorenz": {"predict": 777, "step": 1155},
    "van_der_pol": {"predict": 380, "step": 501},
    "hh_2comp_r1": {"predict": 37.7e3, "step": 40.7e3, "grad": 0.99e6},
    "hh_1comp_r1": {"predict": 9.6e3, "step": 10.0e3, "grad": 0.18e6},
}

if __name__ == "__main__":
    for name, c in configs().items():
        ref = SURVEY_CONSTANTS[name]
        print(f"{name:14s} c_f={c['c_f']:4d} ({c['exp_per_rhs']} exp) c_J={c['c_J']:4d}  rk={c['rk']:6d} tan={c['tan']:7d} "
              f"cov={c['cov']:5d} corr={c.get('corr', 0):5d}  predict={c['predict']:7d} ({c['predict']/ref['predict']:.3f}x) "
              f"step={c['step']:7d} ({c['step']/ref['step']:.3f}x)"
              + (f" grad={c['grad']:8d} ({c['grad']/ref['grad']:.3f}x)" if "grad" in c else ""))
