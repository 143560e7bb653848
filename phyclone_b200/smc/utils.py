"""Data-point order for SMC (reference smc/utils.py:16-119).  Host-side combinatorics, no arithmetic."""

from itertools import repeat


class RootPermutationDistribution(object):
    @staticmethod
    def sample(tree, rng, source=None):
        if source is None:
            sigma = [RootPermutationDistribution.sample(tree, rng, source=node) for node in tree.roots]
            sigma = interleave_lists(sigma, rng)
            outliers = list(tree.outliers)
            rng.shuffle(outliers)
            return interleave_lists([sigma, outliers], rng)
        child_sigma = [RootPermutationDistribution.sample(tree, rng, source=child) for child in tree.get_children(source)]
        sigma = interleave_lists(child_sigma, rng)
        source_sigma = tree.get_data(source)
        rng.shuffle(source_sigma)
        sigma.extend(source_sigma)
        return sigma


def interleave_lists(lists, rng):
    sentinels = []
    for i, lst in enumerate(lists):
        sentinels.extend(repeat(i, len(lst)))
    rng.shuffle(sentinels)
    heads = [iter(lst) for lst in lists]  # the reference pops from the front of each list: O(n) per element
    return [next(heads[idx]) for idx in sentinels]
