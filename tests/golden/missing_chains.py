"""Data sets with missing genotype calls used by the chain goldens: read by make_golden.py
(which runs the unmodified reference on them) and by tests/test_gpu_chain.py (which runs the CUDA path on the same,
regenerated, data).  Arguments of bmagwa_b200.synth.write_dataset."""

MISSING_CHAINS = {
    "chainD": dict(n=203, m_g=300, m_e=2, seed=10, miss_rate=0.03, e_qg=5, var_qg=20, use_individual_tau2=1, do_n_iter=4000,
                   n_rao=100, n_rao_burnin=20, delay_rejection=10, seeds="1234"),
    "chainE": dict(n=150, m_g=200, m_e=0, seed=11, miss_rate=0.12, e_qg=4, var_qg=12, use_individual_tau2=0, do_n_iter=3000,
                   n_rao=50, n_rao_burnin=10, delay_rejection=5, seeds="99"),
}
