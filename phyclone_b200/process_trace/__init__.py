"""Post-processing of traces (reference phyclone/process_trace): the MAP prevalence assignment runs on the device,
trace files keep the reference's gzip+pickle schema."""
from .map import get_map_node_ccfs_and_clonal_prev_dicts, map_trees  # noqa: F401
