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
    """process_trace.py:19-61"""
    results = read_trace(in_file)
    data = results[0]["data"]
    store = store_for(results, device, store)
    chain_num = 0
    if map_type == "frequency":
        topologies = create_topology_dict_from_trace(results, store)
        df = create_topology_dataframe(topologies.values())
        df = df.sort_values(by="count", ascending=False)
        map_iter = df["iter"].iloc[0]
        chain_num = df["chain_num"].iloc[0]
    else:
        map_iter = 0
        map_val = float("-inf")
        for curr_chain_num, chain_results in results.items():
            for i, x in enumerate(chain_results["trace"]):
                if x["log_p_one"] > map_val:
                    map_iter = i
                    map_val = x["log_p_one"]
                    chain_num = curr_chain_num
    tree = Tree.from_dict(results[chain_num]["trace"][map_iter]["tree"], store)
    clusters = results[0].get("clusters", None)
    table = get_clone_table(data, results[0]["samples"], tree, clusters=clusters)
    _create_results_output_files(out_table_file, out_tree_file, table, tree)
    return table


def create_topology_dict_from_trace(trace, store):
    """process_trace.py:64-71.  Rebuilding a tree only records its node refreshes (memoised per distinct
    subtree); nothing is evaluated unless a value is read, so counting topologies costs no kernel time."""
    topologies = dict()
    for chain_num, chain_result in trace.items():
        for i, x in enumerate(chain_result["trace"]):
            curr_tree = Tree.from_dict(x["tree"], store)
            count_topology(topologies, x, i, curr_tree, chain_num)
    return topologies


def count_topology(topologies, x, i, x_top, chain_num=0):
    """process_trace.py:146-166"""
    if x_top in topologies:
        topology = topologies[x_top]
        topology["count"] += 1
        curr_log_p_one = x["log_p_one"]
        if curr_log_p_one > topology["log_p_joint_max"]:
            topology["log_p_joint_max"] = curr_log_p_one
            topology["iter"] = i
            topology["topology"] = x_top
            topology["chain_num"] = chain_num
    else:
        log_mult = x_top.multiplicity
        topologies[x_top] = {
            "topology": x_top, "count": 1, "log_p_joint_max": x["log_p_one"], "iter": i, "chain_num": chain_num,
            "multiplicity": np.exp(log_mult), "log_multiplicity": log_mult,
        }


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
    """process_trace.py:294-320"""
    labels = get_labels_table(data, tree, clusters=clusters)
    ccfs, clonal_prev_dict = map_trees([tree])[0] if assignment is None else assignment
    samples_idx_dict = {k: v for v, k in enumerate(samples)}
    labels["sample_id"] = [samples] * len(labels)
    labels = labels.explode("sample_id")
    grouped = labels.groupby(["clone_id", "sample_id"])
    df_list = []
    for name, group in grouped:
        clone_id, sample_id = name
        group = group.copy()
        if clone_id in ccfs:
            group["ccf"] = ccfs[clone_id][samples_idx_dict[sample_id]]
            group["clonal_prev"] = clonal_prev_dict[clone_id][samples_idx_dict[sample_id]]
        else:
            group["ccf"] = -1
            group["clonal_prev"] = -1
        df_list.append(group)
    return pd.concat(df_list, ignore_index=True)


def get_labels_table(data, tree, clusters=None):
    """process_trace.py:323-383"""
    df_records_list = []
    clone_muts = set()
    outlier_node_name = tree.outlier_node_name
    tree_labels = tree.labels
    if clusters is None:
        for idx in tree_labels:
            df_records_list.append({"mutation_id": data[idx].name, "clone_id": tree_labels[idx]})
            clone_muts.add(data[idx].name)
        for x in data:
            if x.name not in clone_muts:
                df_records_list.append({"mutation_id": x.name, "clone_id": outlier_node_name})
        df = pd.DataFrame(df_records_list)
        df = df.sort_values(by=["clone_id", "mutation_id"])
    else:
        clusters_grouped = clusters.groupby("cluster_id")
        for idx in tree_labels:
            cluster_id = int(data[idx].name)
            clone_id = tree_labels[idx]
            muts_set = clusters_grouped.get_group(cluster_id)["mutation_id"].unique()
            df_records_list.extend({"mutation_id": mut, "clone_id": clone_id, "cluster_id": cluster_id} for mut in muts_set)
            clone_muts.update(muts_set)
        missing_muts_df = clusters.loc[~clusters["mutation_id"].isin(clone_muts)].copy()
        missing_muts_df["clone_id"] = outlier_node_name
        df_records_list.extend(missing_muts_df.to_dict("records"))
        df = pd.DataFrame(df_records_list)
        df = df.sort_values(by=["clone_id", "cluster_id", "mutation_id"])
    return df
