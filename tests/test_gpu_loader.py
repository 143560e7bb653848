"""GPU: TSV -> DataPoints (a1-a4 of SURVEY.md 8a through the file loader) against the reference's grids."""

import numpy as np
import pytest

from oracle import numerics as onum
from tests.parity import assert_close

pytestmark = pytest.mark.gpu


def _write_tsv(tmp_path, golden):
    spec, a, b, t = golden["lik2_cnspec"], golden["lik2_a"], golden["lik2_b"], golden["lik2_t"]
    M, S = a.shape
    rows = ["mutation_id\tsample_id\tref_counts\talt_counts\tmajor_cn\tminor_cn\tnormal_cn\ttumour_content"]
    for m in reversed(range(M)):  # shuffled on purpose: the loader sorts by mutation id
        for s in range(S):
            rows.append("m%02d\ts%d\t%d\t%d\t%d\t%d\t%d\t%r" % (m, s, a[m, s], b[m, s], spec[m, s, 0], spec[m, s, 1],
                                                              spec[m, s, 2], float(t[m, s])))
    p = tmp_path / "data.tsv"
    p.write_text("\n".join(rows) + "\n")
    cl = ["mutation_id\tsample_id\tcluster_id"]
    assign = [2, 0, 1, 1, 0, 2, 2, 2, 0, 1, 0, 0]
    for m in range(M):
        for s in range(S):
            cl.append("m%02d\ts%d\t%d" % (m, s, assign[m]))
    c = tmp_path / "clusters.tsv"
    c.write_text("\n".join(cl) + "\n")
    return str(p), str(c), assign


def test_load_data_unclustered_and_clustered(tmp_path, golden):
    from phyclone_b200.data.pyclone import MajorCopyNumberError, get_major_cn_prior, load_data

    path, cpath, assign = _write_tsv(tmp_path, golden)
    data, samples = load_data(path, grid_size=101, density="beta-binomial", precision=400, outlier_prob=0)
    assert samples == ["s0", "s1", "s2"] and [d.name for d in data] == ["m%02d" % i for i in range(12)]
    got = np.array([d.value for d in data])
    assert_close(got, golden["lik2_bb101"], what="beta-binomial grids from TSV")
    assert all(d.outlier_prob == 0 and d.outlier_prob_not == 0.0 for d in data)
    data, _ = load_data(path, grid_size=101, density="binomial", precision=400, outlier_prob=1e-4)
    assert_close(np.array([d.value for d in data]), golden["lik2_bin101"], what="binomial grids from TSV")
    assert data[0].outlier_prob == pytest.approx(np.log(1e-4)) and data[0].outlier_prob_not == pytest.approx(np.log1p(-1e-4))
    ref_marg = [onum.outlier_marginal(v) for v in golden["lik2_bin101"]]
    assert_close(np.array([d.outlier_marginal_prob for d in data]), np.array(ref_marg), rtol=1e-12)
    # clustered: value = sum of member grids in sorted-mutation order, clusters in sorted id order
    data, _ = load_data(path, cluster_file=cpath, grid_size=101, precision=400, outlier_prob=1e-3)
    assert [d.name for d in data] == ["0", "1", "2"]
    for k, d in enumerate(data):
        mem = [m for m in range(12) if assign[m] == k]
        ref = np.sum(golden["lik2_bb101"][mem], axis=0)
        assert_close(d.value, ref, rtol=1e-12, what="cluster %d" % k)
        assert d.outlier_prob == pytest.approx(np.log(1e-3) * len(mem))
    with pytest.raises(MajorCopyNumberError):
        get_major_cn_prior(1, 2, 2)
    with pytest.raises(Exception):
        load_data(path, density="beta-binomial", precision=None)
