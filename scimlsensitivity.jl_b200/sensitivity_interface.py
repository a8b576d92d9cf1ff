"""Direct interface: adjoint_sensitivities(sol, alg; t, dgdu_discrete, sensealg, ...) -> (du0, dp).

Mirrors /root/reference/src/sensitivity_interface.jl:373-526 (and the Gauss / Quadrature twins
src/gauss_adjoint.jl:766-870, src/quadrature_adjoint.jl:510-633) for an ensemble solution produced by
`solve(EnsembleProblem, alg, EnsembleB200(); ...)`: same keyword names, same defaults, same return convention
(`du0`, and `dp` as a ROW: shape (1, P) for shared parameters, like the reference's `dp'`,
src/sensitivity_interface.jl:503-507).  For an ensemble `du0` is [d, N]; with per-member parameters `dp` is [P, N].
"""
import numpy as np

from .problems import AdjointSensitivityParameterCompatibilityError, AffineCost, QuadraticRunningCost
from .sensitivity_algorithms import (B200Adjoint, BacksolveAdjoint, GaussAdjoint, GaussKronrodAdjoint, InterpolatingAdjoint,
                                     QuadratureAdjoint, sensealg_name)


def _check_params(p):
    """src/sensitivity_interface.jl:438-443"""
    if p is None:
        raise ValueError("Your model does not have parameters, and thus it is impossible to calculate the derivative "
                         "of the solution with respect to the parameters.")
    if hasattr(p, "dtype"):
        if "float" not in str(p.dtype):
            raise AdjointSensitivityParameterCompatibilityError()
        return
    try:
        arr = np.asarray(p)
    except Exception:
        raise AdjointSensitivityParameterCompatibilityError()
    if arr.dtype.kind != "f":
        raise AdjointSensitivityParameterCompatibilityError()


def adjoint_sensitivities(sol, alg=None, *, sensealg=None, t=None, dgdu_discrete=None, dgdp_discrete=None,
                          dgdu_continuous=None, dgdp_continuous=None, g=None, no_start=False,
                          abstol=1.0e-6, reltol=1.0e-3, checkpoints=None, corfunc_analytical=None, callback=None,
                          row_dp=True, **kwargs):
    """dgdu_discrete: an `AffineCost` (evaluated in-kernel) or an array Delta[K, d, N] of cotangents at `t`."""
    if sensealg is None:
        sensealg = InterpolatingAdjoint()          # reference default (src/sensitivity_interface.jl:375)
    inner = sensealg.inner if isinstance(sensealg, B200Adjoint) else sensealg
    if not isinstance(inner, (BacksolveAdjoint, InterpolatingAdjoint, QuadratureAdjoint, GaussAdjoint, GaussKronrodAdjoint)):
        raise TypeError("adjoint_sensitivities: sensealg must be one of the continuous adjoints")
    # events: the forward solution carries them (PresetTimeCallback passed to solve, like the reference's tracked callbacks
    # inside sol.prob.kwargs); a different callback for the reverse pass alone is meaningless
    if callback is not None and getattr(sol.engine, "events", None) is None:
        raise NotImplementedError("pass the PresetTimeCallback to solve(); other callbacks/events are not carried on the "
                                  "B200 path (SURVEY.md App. E): delegate to the reference implementation")
    if getattr(sol.engine, "events", None) is not None and isinstance(inner, QuadratureAdjoint):
        raise NotImplementedError("QuadratureAdjoint does not support callbacks")
    if dgdp_continuous is not None or g is not None:
        raise NotImplementedError("dgdp_continuous / g are not built on the B200 path (named cost families only)")
    if dgdu_continuous is not None and not isinstance(dgdu_continuous, QuadraticRunningCost):
        raise NotImplementedError("dgdu_continuous must be a QuadraticRunningCost on the B200 path")
    if dgdp_discrete is not None:
        raise NotImplementedError("dgdp_discrete is not built on the B200 path yet")
    if dgdu_discrete is None and dgdu_continuous is None:
        # src/interpolating_adjoint.jl:321-326
        raise ValueError("Either `dgdu_discrete`, `dgdp_discrete`, `dgdu_continuous`, `dgdp_continuous`, or `g` "
                         "must be specified.")
    _check_params(sol.p)
    eng = sol.engine
    if eng is None:
        raise RuntimeError("solution carries no live device handle")
    ts = sol.t if t is None else np.asarray(t, dtype=np.float64)
    if dgdu_discrete is None:
        ts = np.zeros(0)                                   # continuous cost only: no jumps (discrete = false in the reference)
    name = sensealg_name(inner)
    # Backsolve through the direct interface: checkpoints default to sol.t = every forward step
    # (src/sensitivity_interface.jl:433); pass `checkpoints=ts` for the rrule behaviour.
    every = False
    checkpointing = True
    if isinstance(inner, BacksolveAdjoint):
        checkpointing = inner.checkpointing
        every = checkpoints is None
    cost = dgdu_discrete if isinstance(dgdu_discrete, AffineCost) else None
    if dgdu_discrete is None:
        cost = AffineCost(0.0, 0.0)
    eng.set_reverse(name, cost=cost, no_start=no_start, checkpointing=checkpointing, ckpt_every_step=every, t=ts)
    if dgdu_continuous is not None or getattr(eng, "_cont_on", False):
        eng.handle.set_continuous_cost(dgdu_continuous is not None, getattr(dgdu_continuous, "a", 0.0), getattr(dgdu_continuous, "b", 0.0))
        eng._cont_on = dgdu_continuous is not None
    # adjoint solve tolerances are keywords of adjoint_sensitivities (src/sensitivity_interface.jl:432; used by the adaptive
    # steppers only); quadgk tolerances come from the sensealg (src/quadrature_adjoint.jl:517)
    is_quad = isinstance(inner, QuadratureAdjoint)
    if getattr(eng, "adaptive", False) or is_quad:
        eng.handle.set_tolerances(abstol if getattr(eng, "adaptive", False) else 0.0, reltol if getattr(eng, "adaptive", False) else 0.0,
                                  inner.abstol if is_quad else 0.0, inner.reltol if is_quad else 0.0)
    du0, dp = eng.reverse(None if cost is not None else dgdu_discrete)
    from .distributed import allreduce_dp
    dp = allreduce_dp(dp, eng)
    if row_dp and eng.shared_p:
        dp = dp.reshape(1, -1)
    return du0, dp
