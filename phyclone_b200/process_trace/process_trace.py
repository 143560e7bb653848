"""Trace consumers (reference process_trace/process_trace.py): main run output, MAP results table, topology
report.  Trees are rebuilt lazily in a device tile store (`Tree.from_dict(d, store)` records the node refreshes
and nothing runs until a value is read); the MAP prevalence assignment of all requested trees is one batched
device launch (`map.map_trees`)."""

import os
import tarfile
import tempfile
from sys import maxsize

import numpy as np
import pandas as pd

from ..tiles import TileStore
from ..tree.tree import Tree
from ..utils.math import exp_normalize
from .consensus import get_consensus_tree, tree_from_consensus
from .map import map_trees
from .trace_io import read_trace, write_trace


def create_main_run_output(cluster_file, out_file, results):
    """process_trace.py:386-393"""
    for chain_result in results.values():
        if cluster_file is not None:
            chain_result["clusters"] = pd.read_csv(cluster_file, sep="\t")[["mutation_id", "cluster_id"]].drop_duplicates()
    write_trace(results, out_file)


def store_for(results, device=None, store=None):
    """Tile store holding the data points' grids of a loaded trace."""
    if store is not None:
        return store
    data = results[0]["data"]
    return TileStore(np.array([dp.value for dp in data]), device=device)


def print_string_to_file(str_to_write, filename):
    with open(filename, "w") as f:
        print(str_to_write, file=f)


def write_map_results(in_file, out_table_file, out_tree_file, map_type="joint-likelihood", device=None, store=None):
    """process_trace.py:19-61: the table and Newick string of ONE sampled tree -- the one with the highest joint
    likelihood, or (map_type "frequency") the best-scoring visit of the most frequently sampled topology."""
    results = read_trace(in_file)
    first = results[0]
    store = store_for(results, device, store)
    if map_type == "frequency":
        ranking = create_topology_dataframe(create_topology_dict_from_trace(results, store).values())
        top = ranking.sort_values(by="count", ascending=False).iloc[0]
        chain_num, map_iter = top["chain_num"], top["iter"]
    else:
        chain_num, map_iter, best = 0, 0, float("-inf")
        for c, chain_results in results.items():
            for i, entry in enumerate(chain_results["trace"]):
                if entry["log_p_one"] > best:  # strict: the first of equal maxima wins, as in the reference
                    chain_num, map_iter, best = c, i, entry["log_p_one"]
    tree = Tree.from_dict(results[chain_num]["trace"][map_iter]["tree"], store)
    table = get_clone_table(first["data"], first["samples"], tree, clusters=first.get("clusters", None))
    _create_results_output_files(out_table_file, out_tree_file, table, tree)
    return table


def create_topology_dict_from_trace(trace, store):
    """process_trace.py:64-71.  Rebuilding a tree only records its node refreshes (memoised per distinct
    subtree); nothing is evaluated unless a value is read, so counting topologies costs no kernel time."""
    topologies = dict()
    top = lambda: store.book.counters()["top"]  # noqa: E731 -- arena slots handed out so far
    permanent = store.arena.alloc.permanent
    # every rebuilt tree takes arena slots for its distinct node ops, and only the trees kept in the dictionary are
    # read again: when the slots in use are many times what the kept trees need, start a new arena generation and
    # re-derive the kept trees in it (a long trace would otherwise double the arena up to its cap).  Only if the
    # store held nothing but its permanent tiles on entry -- a caller's own live trees must not be invalidated.
    compact_ok = hasattr(store, "book") and top() == permanent
    seen_entries = 0
    for chain_num, chain_result in trace.items():
        for i, x in enumerate(chain_result["trace"]):
            curr_tree = Tree.from_dict(x["tree"], store)
            count_topology(topologies, x, i, curr_tree, chain_num)
            seen_entries += 1
            if compact_ok and seen_entries % TOPOLOGY_COMPACT_EVERY == 0:
                kept = sum(2 * (len(t._lr) + len(info["topology"]._lr)) for t, info in topologies.items())
                if top() - permanent > max(TOPOLOGY_COMPACT_MIN_SLOTS, TOPOLOGY_COMPACT_FACTOR * kept):
                    topologies = _compact_topologies(topologies, store)
    return topologies


TOPOLOGY_COMPACT_EVERY = 256      # trace entries between checks
TOPOLOGY_COMPACT_MIN_SLOTS = 8192  # never compact below this many slots in use
TOPOLOGY_COMPACT_FACTOR = 8  # ... nor below this multiple of the kept trees' own slots


def _compact_topologies(topologies, store):
    """New arena generation holding only the trees of `topologies` (same insertion order, same trees structurally)."""
    store.flush()
    store.reset()
    fresh = dict()
    for key, info in topologies.items():
        new_key = Tree.from_dict(key.to_dict(), store)
        info["topology"] = new_key if info["topology"] is key else Tree.from_dict(info["topology"].to_dict(), store)
        fresh[new_key] = info
    return fresh


def count_topology(topologies, x, i, x_top, chain_num=0):
    """process_trace.py:146-166: one visit of topology `x_top` (trace entry `x`, position `i` in its chain)."""
    seen = topologies.get(x_top)
    if seen is None:
        log_mult = x_top.multiplicity
        topologies[x_top] = dict(topology=x_top, count=1, log_p_joint_max=x["log_p_one"], iter=i, chain_num=chain_num,
                                 multiplicity=np.exp(log_mult), log_multiplicity=log_mult)
        return
    seen["count"] += 1
    if x["log_p_one"] > seen["log_p_joint_max"]:
        seen.update(log_p_joint_max=x["log_p_one"], iter=i, topology=x_top, chain_num=chain_num)


def create_topology_dataframe(topologies):
    """process_trace.py:174-182"""
    for topology in topologies:
        tree = topology["topology"]
        topology["topology"] = tree.to_newick_string()
    df = pd.DataFrame(topologies)
    df = df.sort_values(by="log_p_joint_max", ascending=False, ignore_index=True)
    df.insert(0, "topology_id", "t_" + df.index.astype(str))
    return df


def write_topology_report(in_file, out_file, topologies_archive=None, top_trees=float("inf"), device=None, store=None):
    """process_trace.py:74-111"""
    if top_trees == maxsize:
        top_trees = float("inf")
    results = read_trace(in_file)
    store = store_for(results, device, store)
    topologies_dict = create_topology_dict_from_trace(results, store)
    trees = list(topologies_dict.keys())
    topology_df = create_topology_dataframe(topologies_dict.values())
    topology_df.to_csv(out_file, index=False, sep="\t")
    if topologies_archive is not None:
        create_topologies_archive(topology_df, results, top_trees, topologies_dict, topologies_archive, trees)
    return topology_df


def create_topologies_archive(topology_df, results, top_trees, topologies_dict, topologies_archive, trees=None):
    """process_trace.py:114-143; the MAP assignment of every selected topology is ONE device launch."""
    filename_template = "{}_results_table.tsv"
    nwk_template = "{}.nwk"
    clusters = results[0].get("clusters", None)
    data = results[0]["data"]
    samples = results[0]["samples"]
    selected = []
    for tree, values in topologies_dict.items():
        row = topology_df.loc[
            (topology_df["topology"] == values["topology"]) & (topology_df["count"] == values["count"])
            & (topology_df["log_p_joint_max"] == values["log_p_joint_max"]) & (topology_df["iter"] == values["iter"])
            & (topology_df["chain_num"] == values["chain_num"])
        ]
        assert len(row) == 1
        topology_id = row["topology_id"].values[0]
        if int(topology_id[2:]) >= top_trees:
            continue
        selected.append((topology_id, tree))
    assignments = map_trees([t for _, t in selected])
    with tarfile.open(topologies_archive, "w:gz") as archive:
        with tempfile.TemporaryDirectory() as tmp_dir:
            for (topology_id, tree), assignment in zip(selected, assignments):
                table = get_clone_table(data, samples, tree, clusters=clusters, assignment=assignment)
                filename = filename_template.format(topology_id)
                filepath = os.path.join(tmp_dir, filename)
                table.to_csv(filepath, index=False, sep="\t")
                archive.add(filepath, arcname=str(os.path.join(topology_id, filename)))
                nwk_filename = nwk_template.format(topology_id)
                nwk_path = os.path.join(tmp_dir, nwk_filename)
                print_string_to_file(tree.to_newick_string(), nwk_path)
                archive.add(nwk_path, arcname=str(os.path.join(topology_id, nwk_filename)))


def write_consensus_results(in_file, out_table_file, out_tree_file, consensus_threshold=0.5,
                            weight_type="joint-likelihood", device=None, store=None):
    """process_trace.py:185-233: consensus tree of the trace (by counts, or weighted by each topology's best joint
    likelihood times its count), its MAP prevalences and the clone table."""
    results = read_trace(in_file)
    data = results[0]["data"]
    store = store_for(results, device, store)
    trees, probs = [], None
    if weight_type == "counts":
        for chain_results in results.values():
            trees.extend(Tree.from_dict(x["tree"], store) for x in chain_results["trace"])
    else:
        logs = []
        for tree, info in create_topology_dict_from_trace(results, store).items():
            trees.append(tree)
            logs.append(info["log_p_joint_max"] + np.log(info["count"]))
        probs, _ = exp_normalize(np.array(logs))
    children, idxs, _ = get_consensus_tree(trees, data=data, threshold=consensus_threshold, weighted=probs is not None,
                                           log_p_list=probs)
    tree = tree_from_consensus(data, children, idxs, store)
    table = get_clone_table(data, results[0]["samples"], tree, clusters=results[0].get("clusters", None))
    _create_results_output_files(out_table_file, out_tree_file, table, tree)
    return table


def _create_results_output_files(out_table_file, out_tree_file, table, tree):
    table.to_csv(out_table_file, index=False, sep="\t")
    print_string_to_file(tree.to_newick_string(), out_tree_file)


def get_clone_table(data, samples, tree, clusters=None, assignment=None):
    """process_trace.py:294-320: one row per (mutation, sample) with the MAP cellular prevalence (`ccf`) and clonal
    prevalence of the mutation's clone; -1 for mutations outside every clone.  Rows are ordered by clone, then sample,
    then as in the labels table -- the order the reference's group-by loop produces."""
    labels = get_labels_table(data, tree, clusters=clusters)
    ccfs, clonal_prevs = map_trees([tree])[0] if assignment is None else assignment
    column_of = {name: k for k, name in enumerate(samples)}
    rows = labels.loc[labels.index.repeat(len(samples))].copy()
    rows["sample_id"] = list(samples) * len(labels)
    rows = rows.sort_values(by=["clone_id", "sample_id"], kind="stable", ignore_index=True)
    in_tree = rows["clone_id"].isin(list(ccfs)).to_numpy()
    if in_tree.any():
        ccf, prev = np.full(len(rows), -1.0), np.full(len(rows), -1.0)
        for k, (clone, sample) in enumerate(zip(rows["clone_id"], rows["sample_id"])):
            if in_tree[k]:
                ccf[k] = ccfs[clone][column_of[sample]]
                prev[k] = clonal_prevs[clone][column_of[sample]]
    else:  # nothing but outliers: the reference's columns stay integer
        ccf = prev = np.full(len(rows), -1, dtype=np.int64)
    rows["ccf"], rows["clonal_prev"] = ccf, prev
    return rows


def get_labels_table(data, tree, clusters=None):
    """process_trace.py:323-383: (mutation, clone) pairs of the tree; mutations the tree does not hold are listed under
    the outlier clone.  With a cluster table the tree's data points are clusters and every cluster expands to its
    mutations."""
    outlier = tree.outlier_node_name
    clone_of = tree.labels
    if clusters is None:
        placed = [(data[idx].name, clone) for idx, clone in clone_of.items()]
        taken = {name for name, _ in placed}
        placed += [(x.name, outlier) for x in data if x.name not in taken]
        table = pd.DataFrame(placed, columns=["mutation_id", "clone_id"])
        return table.sort_values(by=["clone_id", "mutation_id"])
    members = {cid: grp["mutation_id"].unique() for cid, grp in clusters.groupby("cluster_id")}
    placed = []
    for idx, clone in clone_of.items():
        cid = int(data[idx].name)
        placed += [(mut, clone, cid) for mut in members[cid]]
    taken = {mut for mut, _, _ in placed}
    rest = clusters.loc[~clusters["mutation_id"].isin(taken)]
    placed += [(mut, outlier, cid) for mut, cid in zip(rest["mutation_id"], rest["cluster_id"])]
    table = pd.DataFrame(placed, columns=["mutation_id", "clone_id", "cluster_id"])
    return table.sort_values(by=["clone_id", "cluster_id", "mutation_id"])
