"""Drop-in for the reference's mctsnode.py: Node as a read-only view of the search tree that lives in the GPU's
flat structure-of-arrays store (szb_tree_export).  Field names and the PUCT formula follow mctsnode.py:7-63;
selection / expansion / backup themselves run in the CUDA kernels (k_select / k_finish)."""
import math

import torch

from . import chess_compat as chess


class Node:
    def __init__(self, game=None, args=None, state=None, parent=None, action_taken=None, prior=0, color=chess.WHITE,
                 search_scope_game=None):
        self.game, self.args, self.parent = game, args, parent
        self.action_taken, self.prior, self.color = action_taken, prior, color
        self.children = []
        self.visit_count = 0
        self.value_sum = .0
        self.value = .0

    def is_fully_expanded(self):
        return len(self.children)

    def get_ucb(self, vc, vsum, prior):
        q_value = 1 - (vsum / (vc + 1e-6) + 1) / 2
        return q_value + self.args['C'] * (math.sqrt(self.visit_count) / (vc + 1)) * prior

    def select(self):
        """the child the device kernel would descend into next (same fp32 arithmetic, first maximum wins)"""
        vc = torch.tensor([c.visit_count for c in self.children])
        vsum = torch.tensor([c.value_sum for c in self.children])
        prior = torch.tensor([c.prior for c in self.children])
        return self.children[torch.argmax(self.get_ucb(vc, vsum, prior)).item()]

    def expand(self, policy):
        raise RuntimeError("Node is a read-only view: expansion happens on the GPU (MCTS0.search)")

    def backpropagate(self, value):
        raise RuntimeError("Node is a read-only view: backup happens on the GPU (MCTS0.search)")


def tree_view(export, args, root_color, decode_move):
    """Builds Node objects from szb_tree_export arrays.  decode_move(index, color) -> Move."""
    nodes = []
    n_nodes = len(export["node_first"])
    root = Node(args=args, color=root_color)
    root.visit_count, root.value_sum = export["root_visits"], export["root_value_sum"]
    by_index = {0: root}
    order = list(range(n_nodes))
    for i in order:                      # parents are created before their children (creation order)
        node = by_index[i]
        first, count = int(export["node_first"][i]), int(export["node_count"][i])
        for e in range(first, first + count):
            child = Node(args=args, parent=node, color=not node.color, prior=float(export["edge_prior"][e]),
                         action_taken=decode_move(int(export["edge_move"][e]), node.color))
            child.visit_count = int(export["edge_visits"][e])
            child.value_sum = float(export["edge_value_sum"][e])
            node.children.append(child)
            c = int(export["edge_child"][e])
            if c >= 0:
                by_index[c] = child
        nodes.append(node)
    return root
