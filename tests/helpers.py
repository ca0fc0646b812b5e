"""Shared test helpers: fixture inis and dataset loading through the oracle."""
import os

import numpy as np

from oracle import cpu

SMALL_INI = """[datafiles]
file_fam = {d}/small_modelspace.fam
file_g = {d}/small_modelspace.bed
recode_g_to_minor_allele_count = {recode}
[sizes]
n = 30
m_g = 7
m_e = 0
[sampler]
type = PMV
do_n_iter = {do_n_iter}
n_rao = {n_rao}
n_rao_burnin = {n_rao_burnin}
adaptation = 0
delay_rejection = {delay_rejection}
verbosity = 0
thin = {thin}
save_beta = 0
n_sample_tau2_and_missing = 10
max_move_size = 7
max_SNP_neighborhood_size = 3
[thread]
n_threads = 1
basename = {out}/chain
seeds = {seed}
[model]
types = A
[prior]
e_qg = 2
var_qg = 2
use_individual_tau2 = {indiv}
R2mode_sigma2 = 0.20
nu_sigma2 = 1
nu_tau2_A = 3
s2_tau2_A = 0.02
mu_alpha = 1.0
inv_tau2_e_const_val = 0
inv_tau2_e_val = 1
"""

PLINK_INI = """[datafiles]
file_fam = {d}/plinktest.fam
file_g = {d}/plinktest.bed
recode_g_to_minor_allele_count = 0
[sizes]
n = 5
m_g = 10
m_e = 0
[sampler]
type = PMV
do_n_iter = 100
n_rao = 50
n_rao_burnin = 1
verbosity = 0
thin = 10
n_sample_tau2_and_missing = 10
max_move_size = 5
max_SNP_neighborhood_size = 2
[thread]
n_threads = 1
basename = {out}/chain
seeds = 1245
[model]
types = A
[prior]
e_qg = 2
var_qg = 4
use_individual_tau2 = 0
s2_sigma2 = 1
nu_sigma2 = 4
nu_tau2_A = 4
s2_tau2_A = 1
mu_alpha = 2.0
inv_tau2_e_const_val = 0
inv_tau2_e_val = 1
"""


def read_fam_y(path):
    return np.array([float(l.split()[5]) for l in open(path) if l.strip()])


def load_small(d, recode=True):
    bed = cpu.read_bed(os.path.join(d, "small_modelspace.bed"), 30, 7)
    if recode:
        cpu.recode_minor(bed, 30, 7)
    return bed, read_fam_y(os.path.join(d, "small_modelspace.fam"))


def load_plink(d):
    return cpu.read_bed(os.path.join(d, "plinktest.bed"), 5, 10), read_fam_y(os.path.join(d, "plinktest.fam"))


def read_chain(base):
    def rd(name, dt):
        return np.fromfile(base + "_" + name + ".dat", dtype=dt)
    return dict(jumpdistance=rd("jumpdistance", np.uint8), move_type=rd("move_type", np.uint8),
                move_size=rd("move_size", np.uint8), modelsize=rd("modelsize", np.uint32), loci=rd("loci", np.uint32),
                log_likelihood=rd("log_likelihood", np.float64), log_prior=rd("log_prior", np.float64),
                sigma2=rd("sigma2", np.float64), pve=rd("pve", np.float64), alpha=rd("alpha", np.float64),
                rao=rd("rao", np.float64))
