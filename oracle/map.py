"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MAP prevalence assignment
(process_trace/map.py:5-173): max-plus fold of the children with arg-max back-pointers, prefix max, and the
top-down walk from the last grid index.  The product never imports this module.

Pinned by tests/golden/map_*.npz, generated from the unmodified reference by tests/golden/make_golden_map.py.
"""

import numpy as np


def fold_max(child_row, prev_row):
    """_compute_log_D_n (map.py:117-134): result[i] = max_{j<=i} child[j] + prev[i-j]; the `>=` scan keeps the
    LARGEST maximising j."""
    G = len(prev_row)
    choice = np.zeros(G, dtype=np.int64)
    result = np.empty(G)
    for i in range(G):
        vals = child_row[: i + 1] + prev_row[i::-1]
        k = int(np.argmax(vals[::-1]))  # first maximum of the reversed vector = last maximum of the scan
        choice[i] = i - k
        result[i] = vals[i - k]
    return choice, result


def log_d_max(children):
    """compute_log_D (map.py:96-114): D starts at ZERO (not -inf) and folds every child in list order."""
    log_d = np.zeros(children[0].shape)
    choices = []
    for child in children:
        ch = np.zeros(child.shape, dtype=np.int64)
        new = np.empty(child.shape)
        for s in range(child.shape[0]):
            ch[s], new[s] = fold_max(child[s], log_d[s])
        log_d = new
        choices.append(ch)
    return choices, log_d


def log_s_max(children):
    """compute_log_S (map.py:67-93): running maximum; the strict `>` keeps the EARLIEST maximising index."""
    d_choice, log_d = log_d_max(children)
    S, G = log_d.shape
    log_s = np.maximum.accumulate(log_d, axis=1)
    s_choice = np.zeros((S, G), dtype=np.int64)
    for s in range(S):
        for j in range(1, G):
            s_choice[s, j] = j if log_d[s, j] > log_s[s, j - 1] else s_choice[s, j - 1]
    return d_choice, s_choice, log_s


def map_tree(node_log_p, children, post_order):
    """node_log_p [N,S,G]; children[v] = ordered child list; post_order children-first, root last.
    Returns (max_idx [N,S], log_R_max [N,S,G])  (compute_max_likelihood :48-64, set_max_assignment :137-166)."""
    N, S, G = node_log_p.shape
    r_max = np.zeros((N, S, G))
    d_choice, s_choice = {}, {}
    for v in post_order:
        if not children[v]:
            r_max[v] = node_log_p[v]
        else:
            d_choice[v], s_choice[v], log_s = log_s_max([r_max[c] for c in children[v]])
            r_max[v] = node_log_p[v] + log_s
    max_idx = np.zeros((N, S), dtype=np.int64)
    max_idx[post_order[-1]] = G - 1
    for v in reversed(post_order):
        if not children[v]:
            continue
        tot = [int(s_choice[v][d, max_idx[v, d]]) for d in range(S)]
        for i in range(len(children[v]) - 1, -1, -1):
            c = children[v][i]
            for d in range(S):
                max_idx[c, d] = d_choice[v][i][d, tot[d]]
                tot[d] -= max_idx[c, d]
    return max_idx, r_max


def ccfs_and_clonal_prevs(max_idx, children, G):
    """get_map_ccfs (:169-173) and get_map_clonal_prev (:36-45)."""
    ccf = max_idx / (G - 1)
    prev = ccf.copy()
    for v, kids in enumerate(children):
        for c in kids:
            prev[v] -= ccf[c]
    return ccf, prev
