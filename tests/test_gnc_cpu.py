"""GncOptimizer outer loop on the CPU oracle: control flow of gnc_optimizer.h:53-130 on the
bundle_adjustment example's GNC factors (config A)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from symforce_b200 import desc as D, problems as P  # noqa: E402
from tests import oracle_capi as O  # noqa: E402
from tests.gnc_driver import gnc_optimize  # noqa: E402

GNC = dict(mu_initial=0.0, mu_step=0.33, mu_max=0.99, gnc_update_min_reduction=1e-3)  # test/symforce_gnc_test.cc:14-21


def test_gnc_outer_loop_on_oracle():
    prob = P.ba_example()
    o = O.OracleProblem(prob)
    values = np.array(prob.values, dtype=np.float64, copy=True)
    st, schedule = gnc_optimize(o, values, prob.meta["mu_off"], prob.params, GNC)
    its = o.iterations()
    # mu was stepped 0 -> 0.33 -> 0.66 -> 0.99 and every stage converged
    assert schedule == [0.0, 0.33, 0.66, 0.99]
    assert st.status == D.STATUS_SUCCESS
    assert st.n_iterations == len(its)
    # one record with iteration == -1 (the counter is not reset between stages), numbering continues
    assert [it.iteration for it in its] == list(range(-1, len(its) - 1))
    # damping is relaxed, not reset: lambda never exceeds the initial value at the start of a stage
    assert its[st.best_index].new_error < its[0].new_error
    # plain Optimizer behaviour when mu_step <= 0 (gnc_optimizer.h:56, symforce.lcm:215-217)
    o2 = O.OracleProblem(prob)
    v2 = np.array(prob.values, dtype=np.float64, copy=True)
    st2, sched2 = gnc_optimize(o2, v2, prob.meta["mu_off"], prob.params, dict(GNC, mu_step=0.0))
    o3 = O.OracleProblem(prob)
    st3 = o3.optimize()
    assert sched2 == [0.0] and st2.n_iterations == st3.n_iterations and st2.status == st3.status
    assert np.array_equal(v2, o3.best_values())


def test_reference_gnc_known_answers_on_oracle():
    """test/symforce_gnc_test.cc:23-86 (the reference's own GNC test): 9 iteration records, |x| < 0.1, five times
    closer to zero than the plain optimization with mu = 0, SUCCESS."""
    prob = P.gnc_test()
    o = O.OracleProblem(prob)
    values = np.array(prob.values, dtype=np.float64, copy=True)
    st, schedule = gnc_optimize(o, values, prob.meta["mu_off"], prob.params, GNC)
    x_gnc = values[prob.meta["x_off"]:prob.meta["x_off"] + 5]
    assert len(o.iterations()) == 9
    assert st.status == D.STATUS_SUCCESS
    assert np.linalg.norm(x_gnc) < 0.1
    o2 = O.OracleProblem(prob)  # regular_optimized_values.Set('u', 0.0); sym::Optimize(params, factors, values)
    o2.optimize()
    v2 = o2.best_values()
    x_regular = v2[prob.meta["x_off"]:prob.meta["x_off"] + 5]
    assert np.linalg.norm(x_gnc) * 5 < np.linalg.norm(x_regular)


def test_cpp_gnc_example_builds_the_fixture_data():
    """examples/gnc_test.cc draws its samples with sym::Random on std::mt19937(42) like the reference test; the values it
    hands to the optimizer are the committed fixture the oracle / GPU tests use (host-only run: no GPU needed)."""
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples"), "_build/gnc_test"])
    out = subprocess.run([os.path.join(ROOT, "examples", "_build", "gnc_test")], capture_output=True, text=True,
                         env=dict(os.environ, GNC_TEST_PRINT_VALUES="1"), timeout=60)
    assert out.returncode == 0, out.stderr
    got = np.array([float(x) for x in out.stdout.split()])
    want = P.gnc_test().values
    assert got.shape[0] == want.shape[0] - 1  # the Values of the test get `u` later, from GncOptimizer::Optimize
    assert np.array_equal(got, want[:-1])
