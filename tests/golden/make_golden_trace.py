"""Write a small trace file and its post-processing outputs with the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden_trace.py
mixing.tsv + mixing_clusters.tsv, one chain of 12 iterations with 20 particles, then the reference's own
`create_main_run_output`, `write_map_results` (both map types) and `write_topology_report` with an archive.
Outputs (committed): trace_mixing.pkl.gz, trace_mixing_map_{joint,freq}.tsv/.nwk, trace_mixing_topologies.tsv,
trace_mixing_topologies.tar.gz, trace_mixing_consensus_*.tsv/.nwk.  Graph container = oracle/refshim/rustworkx (so tree["graph"] is a plain list).
"""

import io
import os
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle", "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

from phyclone.data.pyclone import load_data  # noqa: E402
from phyclone.process_trace.process_trace import (create_main_run_output, write_consensus_results,  # noqa: E402
                                                  write_map_results, write_topology_report)
from phyclone.run import run_phyclone_chain  # noqa: E402

DATA = "/root/reference/examples/data/mixing.tsv"
CLUSTERS = "/root/reference/examples/data/mixing_clusters.tsv"


def main():
    rng = np.random.default_rng(0)
    with redirect_stdout(io.StringIO()):
        data, samples = load_data(DATA, rng, 1e-4, 0.4, False, cluster_file=CLUSTERS, density="beta-binomial",
                                  grid_size=101, outlier_prob=0, precision=400)
        res = run_phyclone_chain(1, True, 1.0, data, float("inf"), 12, 20, 1, 1, 0, 10**9, "semi-adapted", 0.5,
                                 np.random.default_rng(1), samples, 1, 0, 0.0)
        for e in res["trace"]:
            e["time"] = 0.0
        p = lambda n: os.path.join(HERE, "trace_mixing" + n)  # noqa: E731
        create_main_run_output(CLUSTERS, p(".pkl.gz"), {0: res})
        write_map_results(p(".pkl.gz"), p("_map_joint.tsv"), p("_map_joint.nwk"))
        write_map_results(p(".pkl.gz"), p("_map_freq.tsv"), p("_map_freq.nwk"), map_type="frequency")
        write_topology_report(p(".pkl.gz"), p("_topologies.tsv"), p("_topologies.tar.gz"), top_trees=3)
        write_consensus_results(p(".pkl.gz"), p("_consensus_joint.tsv"), p("_consensus_joint.nwk"))
        write_consensus_results(p(".pkl.gz"), p("_consensus_counts.tsv"), p("_consensus_counts.nwk"), weight_type="counts")
        write_consensus_results(p(".pkl.gz"), p("_consensus_counts70.tsv"), p("_consensus_counts70.nwk"),
                                consensus_threshold=0.7, weight_type="counts")
    for f in sorted(os.listdir(HERE)):
        if f.startswith("trace_mixing"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
