"""Synthetic data points for oracle / golden runs (test infrastructure): the count recipe of
phyclone_b200.simulate pushed through the NumPy likelihood restatement."""

from phyclone_b200.simulate import simulate_counts

from . import numerics as onum


def make_synthetic(n_clusters, n_samples, grid_size, muts_per_cluster, depth, seed=0, density="beta-binomial",
                   precision=400):
    c = simulate_counts(n_clusters, n_samples, muts_per_cluster, depth, seed)
    grids = onum.likelihood_grid(c["ref"], c["alt"], c["tumour_content"], c["cn"], c["mu"], c["log_pi"], c["n_geno"],
                                 grid_size, density, precision)
    c["values"] = onum.cluster_sum(grids, c["cluster_off"])
    return c
