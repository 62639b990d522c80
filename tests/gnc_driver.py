"""
Test helper: the outer loop of sym::GncOptimizer::Optimize (symforce/opt/gnc_optimizer.h:53-130) restated over the
entry points both libraries export (sfx_* = the CUDA product, orc_* = the CPU oracle):

    values[mu] = mu_initial; early_exit_min_reduction = gnc_update_min_reduction while mu is being stepped;
    Reset + IterateToConvergence; then, while iterations remain and the last stage ended with SUCCESS:
    mu += mu_step (clamped to mu_max, restoring the early-exit threshold), RelaxDampingToInitial,
    ResetState(values) + IterateToConvergence.  The iteration budget is the TOTAL over all stages and the
    iteration records accumulate.
"""
import copy


def gnc_optimize(h, values, mu_off, params, gnc, num_iterations=-1):
    """h: capi.SfxProblem or oracle_capi.OracleProblem; values: float64 buffer (modified in place);
    gnc: dict(mu_initial, mu_step, mu_max, gnc_update_min_reduction).  Returns (stats, mu_schedule)."""
    if num_iterations < 0:
        num_iterations = params.iterations
    updating = gnc["mu_initial"] < gnc["mu_max"] and gnc["mu_step"] > 0.0
    values[mu_off] = gnc["mu_initial"]
    p = copy.copy(params)
    early_exit = params.early_exit_min_reduction
    if updating:
        p.early_exit_min_reduction = gnc["gnc_update_min_reduction"]
    h.update_params(p)
    h.set_values(values)
    st = h.optimize(num_iterations)
    values[:] = h.best_values()
    schedule = [float(values[mu_off])]
    while st.n_iterations < num_iterations:
        if st.status != 1 or not updating:  # SUCCESS == 1
            break
        values[mu_off] = values[mu_off] + gnc["mu_step"]
        h.relax_damping_to_initial()
        if values[mu_off] >= gnc["mu_max"]:
            values[mu_off] = gnc["mu_max"]
            p.early_exit_min_reduction = early_exit
            h.update_params(p)
            updating = False
        schedule.append(float(values[mu_off]))
        h.set_values(values)
        st = h.optimize_continue(num_iterations - st.n_iterations)
        values[:] = h.best_values()
    return st, schedule
