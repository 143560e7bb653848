"""`rustworkx.visit` stand-in (see package docstring: test infrastructure only)."""


class DFSVisitor:
    def discover_vertex(self, v, t):
        return

    def finish_vertex(self, v, t):
        return

    def tree_edge(self, e):
        return

    def back_edge(self, e):
        return

    def forward_or_cross_edge(self, e):
        return


class BFSVisitor:
    def discover_vertex(self, v):
        return

    def finish_vertex(self, v):
        return

    def tree_edge(self, e):
        return

    def non_tree_edge(self, e):
        return

    def gray_target_edge(self, e):
        return

    def black_target_edge(self, e):
        return


class StopSearch(Exception):
    pass


class PruneSearch(Exception):
    pass
