"""Synthetic allele-count inputs of the shapes BASELINE.json names.

Pure host-side input generation (NumPy RNG only, no likelihood arithmetic): a random rooted
clone tree over K clusters, Dirichlet clonal prevalences per sample, cellular prevalence =
subtree sum, and per cluster `m` mutations with depth ~ Poisson(d), alt ~ Binomial(depth, ccf/2)
at genotype (major, minor, normal) = (1, 1, 2), error rate 1e-3, tumour content 1.
"""

import numpy as np


def simulate_counts(n_clusters, n_samples, muts_per_cluster, depth, seed=0, error_rate=1e-3):
    rng = np.random.default_rng(seed)
    K, S, m = n_clusters, n_samples, muts_per_cluster
    parent = np.full(K, -1, dtype=np.int64)
    for k in range(1, K):
        parent[k] = rng.integers(0, k)
    clonal = rng.dirichlet(np.ones(K), size=S)  # [S,K]
    cellular = clonal.copy()
    for k in range(K - 1, 0, -1):
        cellular[:, parent[k]] += cellular[:, k]
    M = K * m
    cluster_of = np.repeat(np.arange(K), m)
    ccf = cellular[:, cluster_of].T  # [M,S]
    total = rng.poisson(depth, size=(M, S)).astype(np.int64)
    total = np.maximum(total, 1)
    alt = rng.binomial(total, np.clip(ccf / 2.0, 0.0, 1.0)).astype(np.int64)
    ref = total - alt
    # genotype table of get_major_cn_prior(1, 1, 2) (reference data/pyclone.py:269-307): one genotype
    cn = np.zeros((M, S, 1, 3), dtype=np.int64)
    cn[..., 0, :] = (2, 2, 2)
    mu = np.zeros((M, S, 1, 3))
    mu[..., 0, :] = (error_rate, error_rate, min(1 - error_rate, 0.5))
    log_pi = np.zeros((M, S, 1))
    n_geno = np.ones((M, S), dtype=np.int32)
    return {
        "ref": ref, "alt": alt, "tumour_content": np.ones((M, S)), "cn": cn, "mu": mu, "log_pi": log_pi,
        "n_geno": n_geno, "cluster_off": np.arange(0, M + 1, m, dtype=np.int64), "parent": parent,
        "cellular": cellular,
    }
