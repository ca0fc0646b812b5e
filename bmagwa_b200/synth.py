"""Seeded synthetic GWAS data in the reference's input formats.

Mirrors the bundled example of the reference (testdata/testdata.sim: null SNPs
plus 20 causal SNPs, allele frequency U(0.02, 0.95), 0.02 of variance each;
testdata/testdata.e: two null U(0,1) covariates) as specified in SURVEY.md
section 8(d): per SNP j, f_j ~ U(0.02, 0.95), g_ij ~ Binomial(2, f_j) i.i.d.;
the last `n_causal` SNPs are causal with beta_j = sqrt(h2_each / (2 f_j (1 - f_j)));
y = sum_j beta_j (g_ij - 2 f_j) + N(0, resid_var).

Files written: PLINK SNP-major .bed (magic 6C 1B 01; 4 individuals per byte,
individual i in bits 2(i%4) of byte i/4; codes 00 -> 0, 10 -> 1, 11 -> 2,
01 -> missing -- the coding read by /root/reference/src/data.cpp:36,245-273),
.fam, .y, .e and an .ini with the keys of testdata/testdata.ini.
"""
import os

import numpy as np

# genotype value (0,1,2, -1 = missing) -> PLINK 2-bit code as the reference decodes it
_CODE_OF_VALUE = np.array([0b00, 0b10, 0b11, 0b01], dtype=np.uint8)  # index -1 -> 0b01


def pack_bed_payload(G):
    """G: (n, m) integer array with values in {-1,0,1,2} -> uint8 payload (m * ceil(n/4))."""
    n, m = G.shape
    B = (n + 3) // 4
    codes = _CODE_OF_VALUE[G.T.astype(np.int64)]  # (m, n); -1 indexes the last entry
    pad = np.zeros((m, 4 * B - n), dtype=np.uint8)
    codes = np.concatenate([codes, pad], axis=1).reshape(m, B, 4)
    payload = codes[:, :, 0] | (codes[:, :, 1] << 2) | (codes[:, :, 2] << 4) | (codes[:, :, 3] << 6)
    return np.ascontiguousarray(payload.astype(np.uint8)).reshape(-1)


def make_genotypes(n, m_g, seed=20121101, miss_rate=0.0, chunk=4096):
    """Returns (payload uint8[m_g*ceil(n/4)], freqs float64[m_g]); generated SNP-chunk-wise."""
    rng = np.random.default_rng(seed)
    B = (n + 3) // 4
    f = rng.uniform(0.02, 0.95, size=m_g)
    payload = np.empty(m_g * B, dtype=np.uint8)
    for lo in range(0, m_g, chunk):
        hi = min(m_g, lo + chunk)
        g = rng.binomial(2, f[lo:hi][None, :], size=(n, hi - lo)).astype(np.int8)
        if miss_rate > 0:
            g[rng.random(size=g.shape) < miss_rate] = -1
        payload[lo * B:hi * B] = pack_bed_payload(g)
    return payload, f


def unpack_payload(payload, n, m_g):
    """uint8 payload -> (n, m_g) int8 additive values with -1 for missing (numpy restatement of data.cpp:40-54)."""
    B = (n + 3) // 4
    p = payload.reshape(m_g, B)
    codes = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(m_g, 4 * B)[:, :n]
    lut = np.array([0, -1, 1, 2], dtype=np.int8)
    return lut[codes].T


def make_phenotype(payload, f, n, m_g, seed=20121101, n_causal=20, h2_each=0.02, resid_var=0.6, binary=False):
    rng = np.random.default_rng(seed + 1)
    n_causal = min(n_causal, m_g)
    B = (n + 3) // 4
    causal = np.arange(m_g - n_causal, m_g)
    G = unpack_payload(payload[(m_g - n_causal) * B:], n, n_causal).astype(np.float64)
    G[G < 0] = 0.0
    fc = f[causal]
    beta = np.sqrt(h2_each / (2.0 * fc * (1.0 - fc)))
    liab = (G - 2.0 * fc[None, :]) @ beta + rng.normal(0.0, np.sqrt(resid_var), size=n)
    y = (liab > 0).astype(np.float64) if binary else liab
    return y, causal, beta


INI_TEMPLATE = """[datafiles]
file_fam = {base}.fam
file_g = {base}.bed
file_e = {base}.e
file_y = {base}.y
recode_g_to_minor_allele_count = {recode}

[sizes]
n = {n}
m_g = {m_g}
m_e = {m_e}

[sampler]
type = PMV
do_n_iter = {do_n_iter}
n_rao = {n_rao}
n_rao_burnin = {n_rao_burnin}
adaptation = 0
verbosity = {verbosity}
thin = {thin}
n_sample_tau2_and_missing = {n_sample_tau2_and_missing}
delay_rejection = {delay_rejection}
p_move_size = 0.2
p_move_size_nbc = 0.25
p_move_size_nbs = 0.7
max_SNP_neighborhood_size = 20
adapt_p_move_size = 1
p_move_size_acpt_goal = 0
max_move_size = {max_move_size}
flat_proposal_dist = 0
save_beta = {save_beta}

[thread]
n_threads = {n_threads}
basename = {outbase}
seeds = {seeds}

[model]
types = {types}

[prior]
use_individual_tau2 = {use_individual_tau2}
type_A = 1
type_H = 1
type_D = 1
type_R = 1
type_AH = 1
e_qg = {e_qg}
var_qg = {var_qg}
R2mode_sigma2 = 0.2
nu_sigma2 = 1
nu_tau2_A = 5
s2_tau2_A = 0.05
nu_tau2_H = 5
s2_tau2_H = 0.05
nu_tau2_D = 5
s2_tau2_D = 0.05
nu_tau2_R = 5
s2_tau2_R = 0.05
mu_alpha = 1
inv_tau2_e_const_val = 0
inv_tau2_e_val = 0.001
"""


def write_dataset(directory, name, n, m_g, m_e=2, seed=20121101, miss_rate=0.0, binary=False,
                  n_causal=20, **ini_overrides):
    """Writes name.{bed,fam,y,e,ini} under `directory`; returns a dict describing the dataset."""
    os.makedirs(directory, exist_ok=True)
    base = os.path.join(directory, name)
    payload, f = make_genotypes(n, m_g, seed=seed, miss_rate=miss_rate)
    y, causal, beta = make_phenotype(payload, f, n, m_g, seed=seed, binary=binary, n_causal=n_causal)
    rng = np.random.default_rng(seed + 2)
    E = rng.uniform(0.0, 1.0, size=(n, m_e))
    with open(base + ".bed", "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        fh.write(payload.tobytes())
    ids = [("per%d" % i, "per%d" % i) for i in range(n)]
    with open(base + ".fam", "w") as fh:
        for (fid, iid), yi in zip(ids, y):
            fh.write("%s %s 0 0 1 %.10g\n" % (fid, iid, yi))
    with open(base + ".y", "w") as fh:
        for (fid, iid), yi in zip(ids, y):
            fh.write("%s %s %.17g\n" % (fid, iid, yi))
    with open(base + ".e", "w") as fh:
        for (fid, iid), row in zip(ids, E):
            fh.write("%s %s %s\n" % (fid, iid, " ".join("%.17g" % v for v in row)))
    cfg = dict(base=base, recode=1, n=n, m_g=m_g, m_e=m_e, do_n_iter=1000, n_rao=500, n_rao_burnin=1000,
               verbosity=100000, thin=10, n_sample_tau2_and_missing=10, delay_rejection=10, max_move_size=20,
               save_beta=0, n_threads=1, outbase=os.path.join(directory, "chain"), seeds="1234",
               use_individual_tau2=1, e_qg=min(20, max(1, m_g // 4)), var_qg=300 if m_g > 100 else 2, types="A")
    cfg.update(ini_overrides)
    ini_path = base + ".ini"
    with open(ini_path, "w") as fh:
        fh.write(INI_TEMPLATE.format(**cfg))
    return dict(ini=ini_path, base=base, n=n, m_g=m_g, m_e=m_e, payload=payload, freqs=f, y=y, E=E,
                causal=causal, beta=beta, cfg=cfg)
