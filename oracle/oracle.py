"""ctypes wrapper of the CPU oracle (oracle/adjoint_oracle.c).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.  Parity status of the oracle
itself: "parity unpinned" at the bit level (no Julia here, no golden vectors upstream); it is pinned
by the reference's own test relations in tests/test_oracle_relations.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

FAM = {"lv": 0, "lorenz": 1, "robertson": 2, "sde_lv": 3, "mlp": 4, "sde_linear": 5, "ball": 6, "relax": 7}
SA = {"interpolating": 0, "gauss": 1, "quadrature": 2, "backsolve": 3, "gauss_kronrod": 4}
ST = {"tsit5_fixed": 0, "rosenbrock23": 1, "em": 2, "euler_heun": 3, "tsit5_adaptive": 4}
FAM_DIMS = {"lv": (2, 4, 0), "lorenz": (3, 3, 0), "robertson": (3, 3, 0), "sde_lv": (2, 6, 2), "ball": (2, 2, 0), "relax": (1, 2, 0)}


class OracleCfg(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("sensealg", C.c_int32), ("stepper", C.c_int32), ("cost_kind", C.c_int32),
        ("d", C.c_int32), ("P", C.c_int32), ("m", C.c_int32), ("K", C.c_int32),
        ("N", C.c_int64),
        ("t0", C.c_double), ("t1", C.c_double), ("dt", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
        ("quad_abstol", C.c_double), ("quad_reltol", C.c_double),
        ("cost_a", C.c_double), ("cost_b", C.c_double),
        ("shared_p", C.c_int32), ("no_start", C.c_int32), ("checkpointing", C.c_int32),
        ("backsolve_ckpt_every_step", C.c_int32), ("mlp_hidden", C.c_int32), ("cont_cost", C.c_int32),
        ("cont_a", C.c_double), ("cont_b", C.c_double),
        ("n_events", C.c_int32), ("_pad", C.c_int32),
        ("ev_times", C.c_void_p), ("ev_scale", C.c_void_p), ("ev_shift", C.c_void_p),
        ("ev_pscale", C.c_void_p), ("ev_pshift", C.c_void_p),
        ("cost_av", C.c_void_p), ("cost_bv", C.c_void_p), ("cont_av", C.c_void_p), ("cont_bv", C.c_void_p),
        ("dgdp_c", C.c_void_p), ("dgdp_e", C.c_void_p), ("cdgdp_c", C.c_void_p), ("cdgdp_e", C.c_void_p),
        ("cc_on", C.c_int32), ("cc_idx", C.c_int32), ("cc_dir", C.c_int32), ("cc_pcomp", C.c_int32), ("cc_pparam", C.c_int32), ("cc_found", C.c_int32),
        ("cc_level", C.c_double), ("cc_psign", C.c_double), ("cc_scale", C.c_void_p), ("cc_shift", C.c_void_p),
        ("cc_lparam", C.c_int32), ("cc_acomp", C.c_int32), ("cc_aparam", C.c_int32), ("cc_qcomp", C.c_int32),
        ("cc_lcoef", C.c_double), ("cc_acoef", C.c_double), ("cc_qcoef", C.c_double),
        ("ev_acomp", C.c_void_p), ("ev_aparam", C.c_void_p), ("ev_acoef", C.c_void_p),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "adjoint_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        assert _lib.oracle_sizeof_cfg() == C.sizeof(OracleCfg)
        _lib.oracle_ensemble_gradient.restype = C.c_int
        _lib.oracle_ensemble_loss.restype = C.c_int
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_cfg(family, sensealg, stepper, N, saveat, t0, t1, dt=0.0, abstol=1e-6, reltol=1e-3, quad_abstol=1e-10,
             quad_reltol=1e-10, cost=("explicit",), shared_p=True, no_start=False, checkpointing=True,
             ckpt_every_step=False, d=None, P=None, mlp_hidden=0, cont_cost=None, events=None, cost_vec=None, cont_vec=None, crossing=None, event_padd=None):
    if family == "mlp":
        d = 2
        H = mlp_hidden
        P = 2 * H + H + H * H + H + 2 * H + 2
        m = 0
    elif family == "sde_linear":
        P, m = 2, d
    else:
        d, P, m = FAM_DIMS[family]
    cfg = OracleCfg()
    cfg.family, cfg.sensealg, cfg.stepper = FAM[family], SA[sensealg], ST[stepper]
    cfg.cost_kind = 0 if cost[0] == "explicit" else 1
    cfg.cost_a, cfg.cost_b = (cost[1], cost[2]) if cost[0] == "affine" else (0.0, 0.0)
    cfg.d, cfg.P, cfg.m, cfg.K, cfg.N = d, P, m, len(saveat), N
    cfg.t0, cfg.t1, cfg.dt, cfg.abstol, cfg.reltol = t0, t1, dt, abstol, reltol
    cfg.quad_abstol, cfg.quad_reltol = quad_abstol, quad_reltol
    cfg.shared_p, cfg.no_start, cfg.checkpointing = int(shared_p), int(no_start), int(checkpointing)
    cfg.backsolve_ckpt_every_step, cfg.mlp_hidden = int(ckpt_every_step), mlp_hidden
    if cont_cost is not None:      # continuous cost g(u) = a/2 |u|^2 + b sum(u)
        cfg.cont_cost, cfg.cont_a, cfg.cont_b = 1, float(cont_cost[0]), float(cont_cost[1])
    # per-component cost coefficients (a[d], b[d], c[P] or None, e[P] or None): dgdu = a .* u + b, dgdp = c .* p + e
    keep = []

    def _vec(x, n):
        if x is None:
            return None
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))
        keep.append(v)
        return v.ctypes.data
    if cost_vec is not None:
        a, b, c, e = (tuple(cost_vec) + (None, None))[:4]
        cfg.cost_kind = 1
        cfg.cost_av, cfg.cost_bv, cfg.dgdp_c, cfg.dgdp_e = _vec(a, d), _vec(b, d), _vec(c, P), _vec(e, P)
    if cont_vec is not None:
        a, b, c, e = (tuple(cont_vec) + (None, None))[:4]
        cfg.cont_cost = 1
        cfg.cont_av, cfg.cont_bv, cfg.cdgdp_c, cfg.cdgdp_e = _vec(a, d), _vec(b, d), _vec(c, P), _vec(e, P)
    cfg._keep_cost = keep
    # continuous callback: crossing = dict(idx, level=0, direction=-1, scale=None, shift=None, pcomp=-1, pparam=0, psign=-1,
    #                                       lparam=-1, lcoef=0 (level += lcoef * p[lparam]), acomp=-1, aparam=0, acoef=0 (u[acomp] += acoef * p[aparam]))
    cfg.cc_lparam, cfg.cc_acomp, cfg.cc_qcomp = -1, -1, -1
    if crossing is not None:
        cfg.cc_qcomp, cfg.cc_qcoef = int(crossing.get("qcomp", -1)), float(crossing.get("qcoef", 1.0))    # u[qcomp] <- qcoef * u[qcomp]^2
        cfg.cc_lparam, cfg.cc_lcoef = int(crossing.get("lparam", -1)), float(crossing.get("lcoef", 0.0))
        cfg.cc_acomp, cfg.cc_aparam, cfg.cc_acoef = int(crossing.get("acomp", -1)), int(crossing.get("aparam", 0)), float(crossing.get("acoef", 0.0))
        cfg.cc_on, cfg.cc_idx, cfg.cc_dir = 1, int(crossing["idx"]), int(crossing.get("direction", -1))
        cfg.cc_level = float(crossing.get("level", 0.0))
        cfg.cc_pcomp, cfg.cc_pparam, cfg.cc_psign = int(crossing.get("pcomp", -1)), int(crossing.get("pparam", 0)), float(crossing.get("psign", -1.0))
        cfg.cc_scale, cfg.cc_shift = _vec(crossing.get("scale", 1.0), d), _vec(crossing.get("shift", 0.0), d)
    if events is not None:         # preset-time events: (times[E], scale[E, d], shift[E, d]), u <- scale * u + shift
        et, es, ec = (np.ascontiguousarray(x, dtype=np.float64) for x in events[:3])
        assert es.shape == (len(et), d) and ec.shape == (len(et), d)
        cfg._keep = [et, es, ec]   # the struct holds raw pointers
        cfg.n_events = len(et)
        cfg.ev_times, cfg.ev_scale, cfg.ev_shift = et.ctypes.data, es.ctypes.data, ec.ctypes.data
        if len(events) == 5 and events[3] is not None:     # parameter-changing affect p <- pscale * p + pshift
            ps, pc = (np.ascontiguousarray(x, dtype=np.float64) for x in events[3:5])
            assert ps.shape == (len(et), P) and pc.shape == (len(et), P)
            cfg._keep += [ps, pc]
            cfg.ev_pscale, cfg.ev_pshift = ps.ctypes.data, pc.ctypes.data
    if event_padd is not None:     # (comp[E], param[E], coef[E]): u[comp[e]] += coef[e] * p[param[e]] at event e (comp < 0: none)
        ac, ak = (np.ascontiguousarray(x, dtype=np.int32) for x in event_padd[:2])
        af = np.ascontiguousarray(event_padd[2], dtype=np.float64)
        assert len(ac) == len(ak) == len(af) == cfg.n_events
        cfg._keep_padd = [ac, ak, af]
        cfg.ev_acomp, cfg.ev_aparam, cfg.ev_acoef = ac.ctypes.data, ak.ctypes.data, af.ctypes.data
    return cfg


def gradient(cfg, saveat, u0, p, dLdu=None, dW=None, want_saved=True, nthreads=0):
    """One gradient evaluation.  u0[d,N], p[P] or [P,N], dLdu[K,d,N], dW[S,m,N] (all float64, C order).
    Returns dict(saved[K,d,N], du0[d,N], dp[P] or [P,N], steps[N])."""
    N, d, P, K = cfg.N, cfg.d, cfg.P, cfg.K
    saveat = np.ascontiguousarray(saveat, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(d, N)
    p = np.ascontiguousarray(p, dtype=np.float64)
    assert p.size == (P if cfg.shared_p else P * N)
    if dLdu is not None:
        dLdu = np.ascontiguousarray(dLdu, dtype=np.float64)
        assert dLdu.shape == (K, d, N)
    if dW is not None:
        dW = np.ascontiguousarray(dW, dtype=np.float64)
    saved = np.zeros((K, d, N)) if want_saved else None
    du0 = np.zeros((d, N))
    dp = np.zeros(P if cfg.shared_p else (P, N))
    steps = np.zeros(N, dtype=np.int32)
    rc = lib().oracle_ensemble_gradient(C.byref(cfg), _ptr(saveat), _ptr(u0), _ptr(p), _ptr(dW), _ptr(dLdu),
                                        _ptr(saved), _ptr(du0), _ptr(dp), _ptr(steps), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_ensemble_gradient failed rc={rc}")
    return {"saved": saved, "du0": du0, "dp": dp, "steps": steps}


def forward(cfg, saveat, u0, p, dW=None, nthreads=0):
    N, d, K = cfg.N, cfg.d, cfg.K
    saveat = np.ascontiguousarray(saveat, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(d, N)
    p = np.ascontiguousarray(p, dtype=np.float64)
    if dW is not None:
        dW = np.ascontiguousarray(dW, dtype=np.float64)
    saved = np.zeros((K, d, N))
    rc = lib().oracle_ensemble_gradient(C.byref(cfg), _ptr(saveat), _ptr(u0), _ptr(p), _ptr(dW), None,
                                        _ptr(saved), None, None, None, C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle forward failed rc={rc}")
    return saved


def loss(cfg, saveat, u0, p, dW=None, nthreads=0):
    """Per-member loss of the affine cost family: L = sum_k sum_j (a/2 u^2 + b u)."""
    N, d = cfg.N, cfg.d
    saveat = np.ascontiguousarray(saveat, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(d, N)
    p = np.ascontiguousarray(p, dtype=np.float64)
    if dW is not None:
        dW = np.ascontiguousarray(dW, dtype=np.float64)
    out = np.zeros(N)
    rc = lib().oracle_ensemble_loss(C.byref(cfg), _ptr(saveat), _ptr(u0), _ptr(p), _ptr(dW), _ptr(out), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle loss failed rc={rc}")
    return out


def event_list(cfg, u0, p, max_events=64):
    """Event times and the states left / right of every event of ONE member's hybrid forward solve (continuous callback)."""
    d = cfg.d
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(d)
    p = np.ascontiguousarray(p, dtype=np.float64)
    t, um, up = np.zeros(max_events), np.zeros((max_events, d)), np.zeros((max_events, d))
    n = lib().oracle_event_list(C.byref(cfg), _ptr(u0), _ptr(p), C.c_int(max_events), _ptr(t), _ptr(um), _ptr(up))
    if n < 0:
        raise RuntimeError(f"oracle_event_list failed rc={n}")
    n = min(n, max_events)
    return t[:n], um[:n], up[:n]


def family_eval(family, u, p, lam, ito=False, mlp_hidden=0):
    d = len(u)
    cfg = make_cfg(family, "interpolating", "tsit5_fixed", 1, [0.0], 0.0, 1.0, dt=0.1, d=d, mlp_hidden=mlp_hidden)
    u, p, lam = (np.ascontiguousarray(x, dtype=np.float64) for x in (u, p, lam))
    f, jtl, ftl = np.zeros(d), np.zeros(d), np.zeros(cfg.P)
    rc = lib().oracle_family_eval(C.byref(cfg), C.c_int(int(ito)), _ptr(u), _ptr(p), _ptr(lam), _ptr(f), _ptr(jtl), _ptr(ftl))
    assert rc == 0, rc
    return f, jtl, ftl


def tsit5_tableau():
    c, a, bt = np.zeros(7), np.zeros((7, 6)), np.zeros(7)
    lib().oracle_tsit5_tableau(_ptr(c), _ptr(a), _ptr(bt))
    return c, a, bt


def tsit5_btheta(theta):
    b = np.zeros(7)
    lib().oracle_tsit5_btheta(C.c_double(theta), _ptr(b))
    return b
