"""TSV -> data points (reference data/pyclone.py:19-436).

The pandas part (reading, filtering, ordering) follows the reference's rules; the arithmetic -- the
per-(mutation, sample, grid point) binomial / beta-binomial log-likelihood, the per-cluster sum and the
outlier marginal -- runs on the device in three batched launches instead of one Numba call per mutation.
`--assign-loss-prob` (a load-time permutation heuristic) is out of scope.
"""

from collections import OrderedDict

import numpy as np
import pandas as pd

from .. import ops
from ..utils.exceptions import MajorCopyNumberError
from .base import DataPoint


def get_major_cn_prior(major_cn, minor_cn, normal_cn, error_rate=1e-3):
    """data/pyclone.py:269-307: genotype table (cn[C,3], mu[C,3], log_pi[C]) of one (mutation, sample)."""
    total_cn = major_cn + minor_cn
    if major_cn < minor_cn:
        raise MajorCopyNumberError(major_cn, minor_cn)
    cn, mu = [], []
    for x in range(1, major_cn + 1):
        cn.append((normal_cn, normal_cn, total_cn))
        mu.append((error_rate, error_rate, min(1 - error_rate, x / total_cn)))
    mutation_after_cn = (normal_cn, total_cn, total_cn)
    if mutation_after_cn not in cn:
        cn.append(mutation_after_cn)
        mu.append((error_rate, error_rate, min(1 - error_rate, 1 / total_cn)))
    n = len(cn)
    return np.array(cn, dtype=np.int64), np.array(mu, dtype=np.float64), np.zeros(n) - np.log(n)


def compute_outlier_prob(outlier_prob, cluster_size):
    """data/pyclone.py:154-160"""
    if outlier_prob == 0:
        return outlier_prob, np.log(1.0)
    return np.log(outlier_prob) * cluster_size, np.log1p(-outlier_prob) * cluster_size


def load_pyclone_table(file_name):
    """data/pyclone.py:164-266 up to the per-mutation objects: returns (mutations, samples, arrays) with the
    arrays shaped [M, S, ...] in sorted-mutation / sorted-sample order."""
    df = pd.read_table(file_name)
    if len(df.columns) == 1:
        df = pd.read_csv(file_name)
    df["sample_id"] = df["sample_id"].astype(str)
    df = df.loc[df["major_cn"] > 0]
    samples = sorted(df["sample_id"].unique())
    size = df.groupby(df["mutation_id"])["sample_id"].transform("size")
    df = df.loc[size == len(samples)].copy()
    if "error_rate" not in df.columns:
        df.loc[:, "error_rate"] = 1e-3
    if "tumour_content" not in df.columns:
        df.loc[:, "tumour_content"] = 1.0
    df = df.sort_values(by="mutation_id", ascending=True)
    mutations = list(OrderedDict.fromkeys(df["mutation_id"]))
    M, S = len(mutations), len(samples)
    recs = {(m, s): row for m, s, row in zip(df["mutation_id"], df["sample_id"], df.itertuples(index=False))}
    tables = {}
    cmax = 1
    for key, row in recs.items():
        t = tables[key] = get_major_cn_prior(int(row.major_cn), int(row.minor_cn), int(row.normal_cn),
                                             error_rate=float(row.error_rate))
        cmax = max(cmax, len(t[0]))
    out = {
        "ref": np.zeros((M, S), dtype=np.int64), "alt": np.zeros((M, S), dtype=np.int64), "t": np.zeros((M, S)),
        "cn": np.zeros((M, S, cmax, 3), dtype=np.int64), "mu": np.zeros((M, S, cmax, 3)),
        "log_pi": np.full((M, S, cmax), -np.inf), "n_geno": np.zeros((M, S), dtype=np.int32),
    }
    for i, m in enumerate(mutations):
        for j, s in enumerate(samples):
            row, (cn, mu, log_pi) = recs[(m, s)], tables[(m, s)]
            c = len(cn)
            out["ref"][i, j], out["alt"][i, j], out["t"][i, j] = row.ref_counts, row.alt_counts, row.tumour_content
            out["cn"][i, j, :c], out["mu"][i, j, :c], out["log_pi"][i, j, :c], out["n_geno"][i, j] = cn, mu, log_pi, c
    return mutations, samples, out


def load_data(file_name, rng=None, low_loss_prob=0.0001, high_loss_prob=0.4, assign_loss_prob=False,
              cluster_file=None, density="beta-binomial", grid_size=101, outlier_prob=1e-4, precision=400):
    """data/pyclone.py:19-107 -> (list of DataPoint, samples)."""
    if assign_loss_prob:
        raise NotImplementedError("--assign-loss-prob is outside the B200 hot path (DESIGN.md section 8)")
    if density == "beta-binomial" and precision is None:
        raise Exception("Precision must be set when using Beta-Binomial.")
    mutations, samples, a = load_pyclone_table(file_name)
    grids = ops.likelihood_grid(a["ref"], a["alt"], a["t"], a["cn"], a["mu"], a["log_pi"], a["n_geno"], grid_size,
                                density, precision)
    if cluster_file is None:
        probs = [compute_outlier_prob(outlier_prob, 1)] * len(mutations)
        marg = ops.outlier_marginal(grids)
        data = [DataPoint(i, grids[i], name=m, outlier_prob=probs[i][0], outlier_prob_not=probs[i][1],
                          outlier_marginal_prob=float(marg[i])) for i, m in enumerate(mutations)]
        return data, samples
    cluster_df = pd.read_csv(cluster_file, sep="\t")
    if "outlier_prob" not in cluster_df.columns:
        cluster_df.loc[:, "outlier_prob"] = outlier_prob
    if outlier_prob == 0:
        cluster_df.loc[:, "outlier_prob"] = outlier_prob
    else:
        cluster_df.loc[cluster_df["outlier_prob"] == 0, "outlier_prob"] = outlier_prob
    cluster_df = cluster_df[["mutation_id", "cluster_id", "outlier_prob"]].drop_duplicates()
    cluster_sizes = cluster_df["cluster_id"].value_counts().to_dict()
    clusters = cluster_df.set_index("mutation_id")["cluster_id"].to_dict()
    cluster_outlier_probs = cluster_df.set_index("cluster_id")["outlier_prob"].to_dict()
    members = {}
    for i, m in enumerate(mutations):
        members.setdefault(clusters[m], []).append(i)
    order = sorted(members)
    idx = np.concatenate([members[c] for c in order])
    off = np.cumsum([0] + [len(members[c]) for c in order])
    values = ops.cluster_sum(grids[idx], off)
    marg = ops.outlier_marginal(values)
    data = []
    for k, c in enumerate(order):
        op, opn = compute_outlier_prob(cluster_outlier_probs[c], cluster_sizes[c])
        data.append(DataPoint(k, values[k], name="{}".format(c), outlier_prob=op, outlier_prob_not=opn,
                              outlier_marginal_prob=float(marg[k])))
    return data, samples
