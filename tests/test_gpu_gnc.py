"""GncOptimizer outer loop on the GPU path against the CPU oracle, both driven by the same
restatement of gnc_optimizer.h:53-130 (tests/gnc_driver.py) on config A (bundle_adjustment example: Barron-robust
InverseRangeLandmarkLinearGncFactor with the convexity parameter mu in the Values)."""
import numpy as np
import pytest

from symforce_b200 import capi, desc as D, problems as P
from tests import oracle_capi as O
from tests.gnc_driver import gnc_optimize

pytestmark = pytest.mark.gpu

GNC = dict(mu_initial=0.0, mu_step=0.33, mu_max=0.99, gnc_update_min_reduction=1e-3)  # test/symforce_gnc_test.cc:14-21
COST_TOL = 1e-8


@pytest.mark.parametrize("gnc", [GNC, dict(GNC, mu_step=0.5), dict(GNC, mu_initial=0.5, mu_step=0.2)])
def test_gnc_matches_oracle(gnc):
    prob = P.ba_example()
    g, o = capi.SfxProblem(prob), O.OracleProblem(prob)
    vg = np.array(prob.values, dtype=np.float64, copy=True)
    vo = vg.copy()
    st_g, sched_g = gnc_optimize(g, vg, prob.meta["mu_off"], prob.params, gnc)
    st_o, sched_o = gnc_optimize(o, vo, prob.meta["mu_off"], prob.params, gnc)
    ig, io = g.iterations(), o.iterations()
    assert sched_g == sched_o and len(sched_g) > 1
    assert (st_g.status, st_g.n_iterations, st_g.best_index) == (st_o.status, st_o.n_iterations, st_o.best_index)
    assert len(ig) == len(io)
    for a, b in zip(ig, io):
        assert a.iteration == b.iteration and a.update_accepted == b.update_accepted
        assert abs(a.new_error - b.new_error) <= COST_TOL * abs(b.new_error)
        assert abs(a.current_lambda - b.current_lambda) <= 1e-12 * abs(b.current_lambda)
    assert np.allclose(vg, vo, rtol=1e-7, atol=1e-9)
    g.close()


def test_continue_requires_a_preceding_optimize():
    prob = P.ba_example()
    g = capi.SfxProblem(prob)
    with pytest.raises(RuntimeError, match="rc=1"):
        g.optimize_continue(3)
    g.optimize(3)
    g.linearize()  # any other entry point invalidates the LM state that a continuation needs
    with pytest.raises(RuntimeError, match="rc=1"):
        g.optimize_continue(3)
    with pytest.raises(RuntimeError, match="rc=1"):
        g.relax_damping_to_initial()
    g.close()


def test_reference_gnc_known_answers():
    """The reference's own GNC test (test/symforce_gnc_test.cc:23-86) on the GPU path -- gnc_factors::BarronFactor as a
    device kind, the sample data of the test reproduced from mt19937(42) -- and parity with the oracle on it."""
    prob = P.gnc_test()
    g, o = capi.SfxProblem(prob), O.OracleProblem(prob)
    vg = np.array(prob.values, dtype=np.float64, copy=True)
    vo = vg.copy()
    st_g, sched_g = gnc_optimize(g, vg, prob.meta["mu_off"], prob.params, GNC)
    st_o, sched_o = gnc_optimize(o, vo, prob.meta["mu_off"], prob.params, GNC)
    x0 = prob.meta["x_off"]
    x_gnc = vg[x0:x0 + 5]
    assert len(g.iterations()) == 9                      # CHECK(gnc_stats.iterations.size() == 9)
    assert np.linalg.norm(x_gnc) < 0.1                   # CHECK(gnc_optimized_x.norm() < 0.1)
    assert st_g.status == D.STATUS_SUCCESS
    g2 = capi.SfxProblem(prob)                           # the plain optimization with u = 0
    g2.optimize()
    x_regular = g2.best_values()[x0:x0 + 5]
    assert np.linalg.norm(x_gnc) * 5 < np.linalg.norm(x_regular)
    # parity with the oracle
    assert sched_g == sched_o
    for a, b in zip(g.iterations(), o.iterations()):
        assert a.iteration == b.iteration and a.update_accepted == b.update_accepted
        assert abs(a.new_error - b.new_error) <= COST_TOL * abs(b.new_error)
    assert np.allclose(vg, vo, rtol=1e-7, atol=1e-9)
    g.close()
    g2.close()
