"""Phylogeny fixtures.

The reference loads trees from Newick files that are NOT in the repository
(`configs/*.yaml` point outside it; `util/phylo_utils.py:64-101` needs ete3).  The three
real CUB trees below were recovered from notebook cell outputs of the reference
(`view_topk.ipynb` cell 2, `node_metrics.ipynb` cell 2, `test.ipynb` cell 6; see
SURVEY.md section 4); the larger ones (cub190, fish38, inat_bird) are seeded synthetic
random binary trees with the right leaf count.

Children are attached one at a time in the listed order, which is what
`construct_phylo_tree` does (`util/phylo_utils.py:85-86`), so child label == list index.
"""
from __future__ import annotations

import random
from typing import Dict, List

from .node import Node

# parent -> ordered children.  Names not appearing as keys are leaves.
CUB08 = {
    'root': ['016+181'],
    '016+181': ['cub_016', '181+097'],
    '181+097': ['181+161', '097+122'],
    '181+161': ['cub_181', '161+165'],
    '161+165': ['cub_161', 'cub_165'],
    '097+122': ['097+011', '122+113'],
    '097+011': ['cub_097', 'cub_011'],
    '122+113': ['cub_122', 'cub_113'],
}

CUB18 = {
    'root': ['052+053', '004+086'],
    '052+053': ['cub_052', '053+050'],
    '053+050': ['cub_053', '050+051'],
    '050+051': ['cub_050', 'cub_051'],
    '004+086': ['004+032', '086+045'],
    '004+032': ['cub_004', '032+033'],
    '032+033': ['cub_032', '033+031'],
    '033+031': ['cub_033', 'cub_031'],
    '086+045': ['cub_086', '045+101'],
    '045+101': ['045+003', '101+023'],
    '045+003': ['cub_045', '003+002'],
    '003+002': ['cub_003', '002+001'],
    '002+001': ['cub_002', 'cub_001'],
    '101+023': ['101+100', '023+025'],
    '101+100': ['cub_101', 'cub_100'],
    '023+025': ['cub_023', '025+024'],
    '025+024': ['cub_025', 'cub_024'],
}

CUB27 = {
    'root': ['113+001+068', 'cub_090'],
    '113+001+068': ['113+060', '001+052', 'cub_068'],
    '113+060': ['113+187', '060+071'],
    '113+187': ['113+037', '187+079'],
    '113+037': ['113+030', '037+077'],
    '113+030': ['113+085', '030+156'],
    '113+085': ['113+194', 'cub_085'],
    '113+194': ['113+118', '194+019'],
    '113+118': ['113+034', 'cub_118'],
    '113+034': ['113+016', 'cub_034'],
    '113+016': ['113+165', 'cub_016'],
    '113+165': ['113+011', '165+181'],
    '113+011': ['113+122', '011+097'],
    '113+122': ['cub_113', 'cub_122'],
    '011+097': ['cub_011', 'cub_097'],
    '165+181': ['165+161', 'cub_181'],
    '165+161': ['cub_165', 'cub_161'],
    '194+019': ['cub_194', 'cub_019'],
    '030+156': ['cub_030', 'cub_156'],
    '037+077': ['cub_037', 'cub_077'],
    '187+079': ['cub_187', 'cub_079'],
    '060+071': ['060+143', 'cub_071'],
    '060+143': ['cub_060', 'cub_143'],
    '001+052': ['001+033', 'cub_052'],
    '001+033': ['cub_001', 'cub_033'],
}

NAMED: Dict[str, Dict[str, List[str]]] = {'cub08': CUB08, 'cub18': CUB18, 'cub27': CUB27}


def synthetic_edges(num_leaves: int, seed: int = 0, prefix: str = 'sp') -> Dict[str, List[str]]:
    """Seeded random strictly-binary tree over `num_leaves` leaves: repeatedly merge two
    random roots of a forest (a coalescent).  N leaves -> N-1 internal nodes."""
    rng = random.Random(seed)
    width = max(3, len(str(num_leaves)))
    forest = [f'{prefix}_{i:0{width}d}' for i in range(num_leaves)]
    edges: Dict[str, List[str]] = {}
    k = 0
    while len(forest) > 1:
        a = forest.pop(rng.randrange(len(forest)))
        b = forest.pop(rng.randrange(len(forest)))
        name = 'root' if not forest else f'n{k:05d}'
        k += 1
        edges[name] = [a, b]
        forest.append(name)
    if num_leaves == 1:
        edges['root'] = forest
    return edges


def build_tree(edges: Dict[str, List[str]], node_cls=Node):
    """Materialise `edges` with any Node-compatible class (ours or the reference's
    `util.node.Node`, which is how the golden generator builds identical trees)."""
    root = node_cls('root')
    frontier = [root]
    while frontier:
        nxt = []
        for n in frontier:
            for child in edges.get(n.name, []):
                n.add_children([child])
                nxt.append(n.children[-1])
        frontier = nxt
    root.assign_all_descendents()
    return root


def get_tree(name: str, node_cls=Node):
    """`cub08|cub18|cub27` (real) or `synth<L>[:seed]` e.g. `synth190`, `synth38:3`."""
    if name in NAMED:
        return build_tree(NAMED[name], node_cls)
    if name.startswith('synth'):
        body = name[len('synth'):]
        leaves, _, seed = body.partition(':')
        return build_tree(synthetic_edges(int(leaves), int(seed or 0)), node_cls)
    raise KeyError(name)


def leaf_names(root) -> List[str]:
    """Sorted leaf names == `ImageFolder.class_to_idx` order (label i <-> names[i])."""
    return sorted(root.leaf_descendents)
