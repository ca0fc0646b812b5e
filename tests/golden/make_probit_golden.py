"""Posterior inclusion probabilities of the PROBIT model on the reference's small_modelspace fixture (n = 30, 2^7 models) from
an INDEPENDENT sampler -- plain numpy, a different algorithm from the product's (single-site Gibbs on the inclusion
indicators with the coefficients integrated out, instead of multi-step Metropolis-Hastings moves with delayed rejection
driven by a Rao-Blackwellised proposal), written from the model definition alone:

    y_i = 1[z_i > 0],  z = b0 + X_gamma beta + eps,  eps ~ N(0, 1)                      (Albert & Chib; sigma2 = 1)
    beta_j = alpha eta_j,  eta_j ~ N(0, tau2),  alpha ~ N(mu_alpha, 1),  tau2 ~ Scaled-Inv-chi2(nu, s2)   (src/prior.cpp:30-141,
        shared tau2, parameter expansion),  b0 flat (inv_tau2_e_const_val = 0)
    gamma ~ beta-binomial with (g_a, g_b) from (e_qg, var_qg)                          (src/prior.hpp:144-167, 273-290)

The probit mode has no reference counterpart, so this chain is what pins it: tests/test_gpu_chain.py requires
the device chain's inclusion probabilities within Monte-Carlo error of tests/golden/probit_pips.json.

    python tests/golden/make_probit_golden.py        # ~2 min; writes probit_pips.json and data/small_modelspace_cc.y
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

N, M_G = 30, 7
PRIOR = dict(e_qg=2.0, var_qg=2.0, nu_tau2=3.0, s2_tau2=0.02, mu_alpha=1.0)
SWEEPS, BURN = 150000, 5000


def load_data():
    from oracle import cpu
    d = os.path.join(HERE, "data")
    bed = np.fromfile(os.path.join(d, "small_modelspace.bed"), dtype=np.uint8)[3:].copy()
    cpu.recode_minor(bed, N, M_G)
    X = np.stack([cpu.decode_column(bed, N, j, 0) for j in range(M_G)], axis=1)
    fam = [l.split() for l in open(os.path.join(d, "small_modelspace.fam")) if l.strip()]
    # case / control labels with a signal the 30 individuals can show: liability on SNPs 1 and 4 (the fixture's own phenotype,
    # dichotomised, leaves the posterior at the prior)
    rs = np.random.default_rng(7)
    liab = 1.3 * (X[:, 1] - X[:, 1].mean()) - 1.3 * (X[:, 4] - X[:, 4].mean()) + 0.6 * rs.normal(size=N)
    labels = (liab > np.median(liab)).astype(np.float64)
    with open(os.path.join(d, "small_modelspace_cc.y"), "w") as fh:
        for f, v in zip(fam, labels):
            fh.write("%s %s %d\n" % (f[0], f[1], int(v)))
    return X, labels


def betabin_params(e_q, var_q, m_g):
    z = (var_q - e_q * (1 - e_q)) / ((m_g - 1) * e_q)
    g_a = (z - 1) / (1 - m_g * z / e_q)
    return g_a, (m_g / e_q - 1) * g_a


def log_ml(XtX, Xtz, zz, idx, prec):
    """log marginal likelihood of z (up to a model-independent constant): intercept flat, SNP coefficients N(0, 1/prec)."""
    cols = [0] + [1 + j for j in idx]
    A = XtX[np.ix_(cols, cols)].copy()
    k = len(idx)
    A[np.arange(1, k + 1), np.arange(1, k + 1)] += prec
    L = np.linalg.cholesky(A)
    v = np.linalg.solve(L, Xtz[cols])
    return 0.5 * k * np.log(prec) - np.log(np.diag(L)).sum() - 0.5 * (zz - v @ v)


def main():
    from scipy.special import ndtr, ndtri
    X, labels = load_data()
    rs = np.random.default_rng(20121101)
    g_a, g_b = betabin_params(PRIOR["e_qg"], PRIOR["var_qg"], M_G)
    nu, s2, mu_a = PRIOR["nu_tau2"], PRIOR["s2_tau2"], PRIOR["mu_alpha"]
    D = np.concatenate([np.ones((N, 1)), X], axis=1)
    XtX = D.T @ D
    case = labels > 0.5
    gamma = np.zeros(M_G, dtype=bool)
    alpha, tau2 = 1.0, nu * s2 / max(0.1, nu - 2)
    z = np.where(case, 0.8, -0.8)
    incl = np.zeros(M_G)
    rb = np.zeros(M_G)
    size_hist = np.zeros(M_G + 1)
    kept = 0
    for sweep in range(SWEEPS):
        prec = 1.0 / (alpha * alpha * tau2)
        Xtz, zz = D.T @ z, z @ z
        # 1. inclusion indicators, coefficients integrated out
        for j in rs.permutation(M_G):
            others = [i for i in range(M_G) if gamma[i] and i != j]
            L_others = len(others)
            l1 = log_ml(XtX, Xtz, zz, sorted(others + [j]), prec) + np.log(g_a + L_others)
            l0 = log_ml(XtX, Xtz, zz, others, prec) + np.log(g_b + M_G - L_others - 1)
            p1 = 1.0 / (1.0 + np.exp(l0 - l1))
            gamma[j] = rs.uniform() < p1
            if sweep >= BURN:
                rb[j] += p1
        idx = [i for i in range(M_G) if gamma[i]]
        k = len(idx)
        # 2. coefficients given the model
        cols = [0] + [1 + j for j in idx]
        A = XtX[np.ix_(cols, cols)].copy()
        A[np.arange(1, k + 1), np.arange(1, k + 1)] += prec
        L = np.linalg.cholesky(A)
        mean = np.linalg.solve(L.T, np.linalg.solve(L, Xtz[cols]))
        beta = mean + np.linalg.solve(L.T, rs.normal(size=k + 1))
        b0, bg = beta[0], beta[1:]
        eta = bg / alpha
        # 3. tau2 | eta  (scaled inverse chi-square)
        nu_n = nu + k
        tau2 = (nu * s2 + eta @ eta) / rs.chisquare(nu_n)
        # 4. alpha | eta, z, b0
        if k > 0:
            xe = X[:, idx] @ eta
            var = 1.0 / (1.0 + xe @ xe)
            alpha = np.sqrt(var) * rs.normal() + var * (mu_a + xe @ (z - b0))
        else:
            alpha = rs.normal() + mu_a
        bg = alpha * eta
        # 5. latent phenotype | coefficients, labels
        mu = b0 + (X[:, idx] @ bg if k else 0.0)
        u = rs.uniform(size=N)
        z = np.where(case, mu - ndtri(u * ndtr(mu)), mu + ndtri(u * ndtr(-mu)))
        if sweep >= BURN:
            incl += gamma
            size_hist[k] += 1
            kept += 1
    out = dict(prior=PRIOR, sweeps=SWEEPS, burn=BURN, n=N, m_g=M_G, labels=[int(v) for v in labels],
               pip_frequency=[float(v) for v in incl / kept], pip_rao_blackwell=[float(v) for v in rb / kept],
               model_size_distribution=[float(v) for v in size_hist / kept])
    json.dump(out, open(os.path.join(HERE, "probit_pips.json"), "w"), indent=1)
    print("PIP (frequency)      ", np.round(incl / kept, 4))
    print("PIP (Rao-Blackwell)  ", np.round(rb / kept, 4))
    print("model size distribution", np.round(size_hist / kept, 4))


if __name__ == "__main__":
    main()
