"""Phylogeny tree used by the per-node prototype head (host side).

Mirrors the public surface of the reference tree class (`util/node.py:16-491` in
harishB97/PIPNet) that the hot path and its callers touch -- child/label maps,
`nodes_with_children()` order, leaf-descendant sets, `set_num_protos`, class-loss
weights and the joint leaf distribution -- but is written from scratch: descendant
sets are computed once in a single post-order sweep and every traversal is iterative.

Semantics that matter for parity (all cited against the reference):
  * `add_children` sorts the names passed in one call and labels them with the next
    free indices (`util/node.py:73-81`).
  * `nodes_with_children()` is breadth-first from the receiver, parents before
    children, siblings in insertion order (`util/node.py:174-185`).  This order is the
    column order of the flat prototype axis used by the CUDA kernels.
  * `set_num_protos` follows `util/node.py:43-71` (per-child mode, min-protos split
    mode, per-descendant mode).
  * `set_loss_weightage_using_descendants_count` gives `min(d)/d_c` (`util/node.py:37-41`).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn.functional as F


def split_value(total: int, parts: int) -> List[int]:
    """Split `total` into `parts` near-equal integers, larger ones first
    (same contract as `util/node.py:9-14`)."""
    q, r = divmod(total, parts)
    return [q + (1 if i < r else 0) for i in range(parts)]


class Node:
    def __init__(self, name: str, parent: Optional["Node"] = None, label: Optional[int] = None):
        self.name = name
        self.parent = parent
        self.label = label
        self.children: List[Node] = []
        self.children_to_labels: Dict[str, int] = {}
        self.weights = None
        self.num_protos = 0
        self.num_protos_per_child: Optional[Dict[str, int]] = None
        # filled by assign_all_descendents()
        self.descendents: set = set()
        self.leaf_descendents: set = set()
        self.leaf_descendents_of_child: Dict[str, set] = defaultdict(set)

    # ------------------------------------------------------------------ construction
    def add_children(self, names, labels=None):
        if not isinstance(names, list):
            names = [names]
        if labels is None:
            first = len(self.children)
            labels = list(range(first, first + len(names)))
        names.sort()
        for nm, lb in zip(names, labels):
            self.children.append(Node(nm, parent=self, label=lb))
            self.children_to_labels[nm] = lb

    def add_children_to(self, name, children):
        self.get_node(name).add_children(children)

    # ------------------------------------------------------------------ queries
    def num_children(self) -> int:
        return len(self.children)

    def is_leaf(self) -> bool:
        return not self.children

    def has_logits(self) -> bool:
        return len(self.children) > 1

    def children_names(self) -> List[str]:
        return [c.name for c in self.children]

    def get_child(self, name):
        for c in self.children:
            if c.name == name:
                return c
        return None

    def _bfs(self) -> Iterable["Node"]:
        frontier = [self]
        while frontier:
            nxt = []
            for n in frontier:
                yield n
                nxt.extend(n.children)
            frontier = nxt

    def get_node(self, name):
        for n in self._bfs():
            if n.name == name:
                return n
        print("node for " + name + " not found")
        return None

    def get_node_attr(self, name, attr):
        return getattr(self.get_node(name), attr)

    def set_node_attr(self, name, attr, value):
        return setattr(self.get_node(name), attr, value)

    def nodes_with_children(self) -> List["Node"]:
        return [n for n in self._bfs() if n.children]

    def nodes_without_children(self) -> List["Node"]:
        return [n for n in self._bfs() if not n.has_logits()]

    def classes_with_children(self) -> List[str]:
        return [n.name for n in self.nodes_with_children()]

    def class_to_num_children(self) -> Dict[str, int]:
        return {n.name: n.num_children() for n in self._bfs()}

    def leaves(self) -> List["Node"]:
        return [n for n in self._bfs() if not n.children]

    # ------------------------------------------------------------------ descendants
    def assign_all_descendents(self):
        """One post-order sweep that fills descendents / leaf_descendents /
        leaf_descendents_of_child for every node below (and including) `self`.
        Equivalent to `util/node.py:238-261` + `:207-236`; a leaf maps to itself."""
        order = list(self._bfs())
        for n in reversed(order):
            if not n.children:
                n.descendents = set()
                n.leaf_descendents = {n.name}
                n.leaf_descendents_of_child = defaultdict(set)
                continue
            desc, leafs, per_child = set(), set(), defaultdict(set)
            for c in n.children:
                desc.add(c.name)
                desc |= c.descendents
                leafs |= c.leaf_descendents
                per_child[c.name] = set(c.leaf_descendents)
            n.descendents, n.leaf_descendents, n.leaf_descendents_of_child = desc, leafs, per_child

    assign_all_leaf_descendents = assign_all_descendents

    def is_descendent(self, name) -> bool:
        return name in self.descendents

    def num_descendents(self) -> int:
        return len(self.descendents)

    def num_leaf_descendents(self) -> int:
        return len(self.leaf_descendents)

    def closest_descendent_for(self, name) -> "Node":
        """Child of `self` on the path to `name` (`util/node.py:278-282`)."""
        for c in self.children:
            if c.name == name:
                return c
        for c in self.children:
            if name in c.descendents:
                return c
        raise KeyError(f"{name} is not below {self.name}")

    # ------------------------------------------------------------------ prototypes / weights
    def set_num_protos(self, num_protos_per_descendant, num_protos_per_child, min_protos=0, split_protos=False):
        if num_protos_per_child > 0:
            self.num_protos_per_child = {}
            total = 0
            for c in self.children:
                k = max(num_protos_per_child, num_protos_per_descendant * c.num_leaf_descendents())
                self.num_protos_per_child[c.name] = k
                total += k
            self.num_protos = total
            return
        want = self.num_leaf_descendents() * num_protos_per_descendant
        self.num_protos = max(min_protos, want)
        if not split_protos:
            raise NotImplementedError()
        self.num_protos_per_child = {}
        if min_protos > want:
            for c, k in zip(self.children, split_value(min_protos, self.num_children())):
                self.num_protos_per_child[c.name] = k
        elif min_protos < want:
            for c in self.children:
                self.num_protos_per_child[c.name] = len(self.leaf_descendents_of_child[c.name]) * num_protos_per_descendant
        # min_protos == want leaves the map empty, exactly like the reference (util/node.py:62-68)

    def set_loss_weightage_using_descendants_count(self):
        counts = [len(self.leaf_descendents_of_child[c.name]) for c in self.children]
        self.num_descendants_of_each_child = counts
        self.weights = min(counts) / torch.tensor(counts, requires_grad=False)

    def set_loss_weightage(self, class_size_count):
        counts = [sum(class_size_count[l] for l in self.leaf_descendents_of_child[c.name]) for c in self.children]
        self.num_images_of_each_child = counts
        self.weights = min(counts) / torch.tensor(counts, requires_grad=False)

    # ------------------------------------------------------------------ joint distribution
    def names_of_joint_distribution(self):
        if len(self.children) == 1:
            return [self.children[0].name]
        if not self.children:
            return [self.name]
        return [c.names_of_joint_distribution() for c in self.children]

    def unwrap_names_of_joint(self, names):
        flat, stack = [], [iter(names)]
        while stack:
            try:
                item = next(stack[-1])
            except StopIteration:
                stack.pop()
                continue
            if isinstance(item, list):
                stack.append(iter(item))
            else:
                flat.append(item)
        return flat

    def distribution_over_furthest_descendents(self, net, batch_size, out, leave_out_classes=None,
                                               apply_overspecificity_mask=False, device='cuda', softmax_tau=1):
        """Per-sample probability of every leaf below `self` in depth-first child order
        (`util/node.py:300-385`): product along the path of
        softmax(log1p(out[node]**2)/tau)[:, child].  PyTorch formulation kept for API
        parity; the fused path uses `ops.joint_leaf_distribution` instead."""
        if leave_out_classes:
            raise NotImplementedError("leave_out_classes is outside the B200 hot path (SURVEY.md section 8)")
        if apply_overspecificity_mask:
            raise NotImplementedError("overspecificity mask is outside the B200 hot path (SURVEY.md section 8)")
        if self.is_leaf():
            return torch.ones(batch_size, 1, device=device)
        probs = F.softmax(torch.log1p(out[self.name] ** 2) / softmax_tau, 1)
        cols = []
        for i, c in enumerate(self.children):
            sub = c.distribution_over_furthest_descendents(net=net, batch_size=batch_size, out=out, device=device,
                                                           softmax_tau=softmax_tau)
            cols.append(probs[:, i].view(batch_size, 1) * sub)
        return torch.cat(cols, 1)

    # ------------------------------------------------------------------ printing
    def _print(self, depth=0):
        lines, stack = [], [(self, depth)]
        while stack:
            n, d = stack.pop()
            lines.append('\t' * d + n.name)
            for c in reversed(n.children):
                stack.append((c, d + 1))
        return '\n'.join(lines) + '\n'

    def __str__(self):
        return self._print()
