"""Order in which an SMC sweep visits the data points of the current tree (reference smc/utils.py:16-119): a random
linear extension of "descendants before ancestors" -- the points of a subtree come before the points of its root
node -- with the outliers shuffled in anywhere.  Host-side combinatorics; what matters for parity is the sequence of
`rng.shuffle` calls, which is the reference's."""

from itertools import repeat


def interleave_lists(lists, rng):
    """A uniformly random merge of the lists that keeps the order inside each (smc/utils.py:103-119): shuffle one
    marker per element, then deal the lists out in marker order."""
    markers = []
    for which, lst in enumerate(lists):
        markers.extend(repeat(which, len(lst)))
    rng.shuffle(markers)
    heads = [iter(lst) for lst in lists]  # the reference pops from the front of each list: O(n) per element
    return [next(heads[which]) for which in markers]


def _subtree_order(tree, rng, node):
    """Points below `node` in a random compatible order, then the node's own points in random order."""
    below = interleave_lists([_subtree_order(tree, rng, child) for child in tree.get_children(node)], rng)
    own = tree.get_data(node)
    rng.shuffle(own)
    below.extend(own)
    return below


def _log_factorial(n):
    from ..utils.math import cached_log_factorial

    return cached_log_factorial(n)


def _log_multinomial(sizes):
    """utils/math.py:159-177"""
    if len(sizes) == 0:
        return 0
    result = _log_factorial(sum(sizes))
    for x in sizes:
        result -= _log_factorial(x)
    return result


class RootPermutationDistribution(object):
    @staticmethod
    def subtree_terms(tree, node):
        """(log number of compatible orders of the subtree under `node`, data points in it): smc/utils.py:41-62"""
        count, sizes = 0, []
        for child in tree.get_children(node):
            c, n = RootPermutationDistribution.subtree_terms(tree, child)
            count += c
            sizes.append(n)
        own = tree.get_data_len(node)
        count += _log_multinomial(sizes)
        count += _log_factorial(own)
        return count, own + sum(sizes)

    @staticmethod
    def log_count(tree, source=None):
        """smc/utils.py:18-62"""
        if source is not None:
            return RootPermutationDistribution.subtree_terms(tree, source)[0]
        from ..utils.math import log_binomial_coefficient

        count, sizes = 0, []
        for node in tree.roots:
            c, n = RootPermutationDistribution.subtree_terms(tree, node)
            count += c
            sizes.append(n)
        count += _log_multinomial(sizes)
        count += log_binomial_coefficient(len(tree.data), len(tree.outliers))
        return count

    @staticmethod
    def log_pdf(tree):
        """smc/utils.py:64-66"""
        return -RootPermutationDistribution.log_count(tree)

    @staticmethod
    def sample(tree, rng, source=None):
        if source is not None:
            return _subtree_order(tree, rng, source)
        in_tree = interleave_lists([_subtree_order(tree, rng, root) for root in tree.roots], rng)
        outliers = list(tree.outliers)
        rng.shuffle(outliers)
        return interleave_lists([in_tree, outliers], rng)
