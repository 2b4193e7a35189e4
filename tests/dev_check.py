"""Developer scratch check: host-compiled kernel source vs the torch oracle (not a test)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))  # lives under tests/: it imports the oracle
import numpy as np, torch
from oracle import ref_torch as R
import util as U
from ode_uncertainty_b200 import _native as N

def case(ode_name, ode_id, tab, solver_id, T, L_sel=None, variant=0, disable=False, Qw=None, gamma=0.0, cov='diagonal', cov_id=0, scale=1.0, x0=None, h=0.01, Rvar=1e-3, t0=0.0, backend="hostemu", every=1):
    ode, params, shape = R.ODES[ode_name]
    n = shape[0]*shape[1]
    x0 = torch.tensor(x0, dtype=torch.float64).reshape(shape)
    P0s = torch.eye(n)*1e-12
    Q = torch.zeros(n,n) if Qw is None else torch.diag(torch.tensor(Qw))
    if L_sel is not None:
        H = torch.eye(n)[L_sel]
        L = H.shape[0]
        Rs = torch.eye(L)*Rvar**0.5
        xs,_ = R.run_rk(ode, params, tab, h, t0, x0, T)
        rng = np.random.default_rng(8)
        ys = (xs[1:].reshape(T,-1) @ H.T) + torch.tensor(rng.normal(0, Rvar**0.5, (T,L)))
        flags, ymap = U.sync_times_all(T)
        if every > 1:
            flags = (np.arange(1, T+1) % every == 0).astype(np.uint8)
    else:
        H = torch.eye(n); L=0; Rs = torch.zeros(0,0); ys=torch.zeros(1,0); flags=np.zeros(T,dtype=np.uint8); ymap=np.zeros(T,dtype=np.int64)
    st = R.init_state(t0, x0, P0s, Q, gamma**0.5, Rs)
    t=time.time()
    traj, nll, quirks = R.run_filter(ode, params, tab, h, cov, scale, disable, st, H, ys, flags, ymap, T, 1)
    t_or=time.time()-t
    plan = U.make_plan(ode_id=ode_id, solver_id=solver_id, step_size=h, ode_variant=variant, cov_fn_id=cov_id, cov_scale=scale, disable_cov_update=disable)
    kw = dict(t0=t0, P0_sqrt=P0s.numpy(), Q_sqrt=Q.numpy(), gamma_sqrt=gamma**0.5, save_interval=1)
    if L>0: kw.update(H=H.numpy(), R_sqrt=Rs.numpy(), ys=ys.numpy(), correct_flags=flags, xy_index_map=ymap)
    out = U.run_ekf(backend, plan, x0.reshape(1,n).numpy(), T, **kw)
    tr = out["traj"]
    print(ode_name, tab, "T",T,"L",L, "oracle %.1fs"%t_or, quirks)
    for k in ("t","x","eps","P","y_hat","S"):
        a = tr[k][:,0] if k!="t" else tr[k]
        b = traj[k].numpy()
        if b.size==0: continue
        err = np.abs(a-b).reshape(len(b),-1).max(1)/np.maximum(np.abs(b).reshape(len(b),-1).max(1),1e-300)
        print("  %-6s max rel err per-step: %.2e (at %d)  final %.2e"%(k, err.max(), err.argmax(), err[-1]))
    print("  nll", out["nll"][0], float(nll), abs(out["nll"][0]-float(nll))/max(abs(float(nll)),1e-300))

if __name__ == "__main__":
    be = sys.argv[1] if len(sys.argv) > 1 else "hostemu"
    case("Lorenz", N.ODE_LORENZ, "RKF45", N.SOLVER_RKF45, 300, x0=[1.,1.,1.], backend=be)
    case("Lorenz", N.ODE_LORENZ, "RKF45", N.SOLVER_RKF45, 300, L_sel=[0,1,2], x0=[1.,1.,1.], backend=be)
    case("VanDerPol", N.ODE_VAN_DER_POL, "Dopri65", N.SOLVER_DOPRI65, 200, L_sel=[0], x0=[2.,10.], t0=10.0, backend=be)
    case("LotkaVolterra", N.ODE_LOTKA_VOLTERRA, "BS32", N.SOLVER_BS32, 200, L_sel=[0], x0=[1.,1.], disable=True, Qw=[1.0,1.0], gamma=1e-5, Rvar=0.1, backend=be)
    case("LotkaVolterra", N.ODE_LOTKA_VOLTERRA, "HeunEuler", N.SOLVER_HEUN_EULER, 100, L_sel=[0], x0=[1.,1.], Qw=[1.0,1.0], gamma=1e-5, Rvar=0.1, backend=be, every=3)
    case("HodgkinHuxley/reduced-1", N.ODE_HODGKIN_HUXLEY, "RKF45", N.SOLVER_RKF45, 100, L_sel=[0], variant=1, x0=R.hh_initial_value("reduced-1", -70.0, R.ODES["HodgkinHuxley/reduced-1"][1]).flatten().tolist(), disable=True, Qw=[1.0]*7, gamma=1e-2, Rvar=0.1, backend=be, t0=9.5)
