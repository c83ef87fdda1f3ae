"""Text-line merge: groups OCR'd lines into text regions.

Behaviour of manga_translator/textline_merge/__init__.py (`merge_bboxes_text_region` :110-181, `split_text_region` :10-84,
`dispatch` :183-208), restated without shapely/networkx: the pairwise merge predicate and the polygon primitives come from
`geometry.py`, the minimum spanning tree is a plain Kruskal whose edge order reproduces networkx's (complete graph in
`itertools.combinations` order, stable sort by weight), connected components are found with a union-find.
Pinned by the reference's own known-answer tests (tests/golden/textline_merge.json, extracted by oracle/make_merge_golden.py)."""
import itertools
from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Sequence, Set, Tuple

import numpy as np

from .geometry import Quadrilateral, can_merge_matrix, can_merge_region, polygon_distance


class _DisjointSets:
    def __init__(self, items: Iterable[int]):
        self.parent: Dict[int, int] = {i: i for i in items}

    def find(self, x: int) -> int:
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
        return x

    def union(self, a: int, b: int) -> bool:
        ra, rb = self.find(a), self.find(b)
        if ra == rb:
            return False
        self.parent[ra] = rb
        return True


def _components(nodes: Sequence[int], edges: Iterable[Tuple[int, int]]) -> List[List[int]]:
    """Connected components, each listed in `nodes` order, components ordered by their first node."""
    ds = _DisjointSets(nodes)
    for u, v in edges:
        ds.union(u, v)
    groups: Dict[int, List[int]] = {}
    for n in nodes:
        groups.setdefault(ds.find(n), []).append(n)
    return list(groups.values())


def _spanning_tree(nodes: Sequence[int], bboxes: Sequence[Quadrilateral]) -> List[Tuple[int, int, float]]:
    """Minimum spanning tree of the complete graph over `nodes` weighted by the reading-order distance, heaviest edge first."""
    cand = [(u, v, bboxes[u].distance(bboxes[v])) for u, v in itertools.combinations(nodes, 2)]
    ds = _DisjointSets(nodes)
    tree = [e for e in sorted(cand, key=lambda e: e[2]) if ds.union(e[0], e[1])]        # sorted() is stable
    return sorted(tree, key=lambda e: e[2], reverse=True)


def split_text_region(bboxes: Sequence[Quadrilateral], region: Iterable[int], gamma: float = 0.5, sigma: float = 2.0) -> List[Set[int]]:
    """Splits a merge candidate at spanning-tree edges that are outliers (`textline_merge/__init__.py:10-84`)."""
    idx = list(region)
    if len(idx) == 1:
        return [set(idx)]
    if len(idx) == 2:
        a, b = bboxes[idx[0]], bboxes[idx[1]]
        fs = max(a.font_size, b.font_size)
        if a.distance(b) < (1 + gamma) * fs and abs(a.angle - b.angle) < 0.2 * np.pi:
            return [set(idx)]
        return [{idx[0]}, {idx[1]}]
    tree = _spanning_tree(idx, bboxes)
    dists = [e[2] for e in tree]
    fontsize = float(np.mean([bboxes[i].font_size for i in idx]))
    d_std, d_mean = float(np.std(dists)), float(np.mean(dists))
    std_threshold = max(0.3 * fontsize + 5, 5)
    b1, b2 = bboxes[tree[0][0]], bboxes[tree[0][1]]
    gap = polygon_distance(b1.pts, b2.pts)
    alignment = min(abs(b1.centroid[0] - b2.centroid[0]), abs(b1.centroid[1] - b2.centroid[1]))
    longest_ok = dists[0] <= d_mean + d_std * sigma or dists[0] <= fontsize * (1 + gamma)
    spread_ok = d_std < std_threshold or (gap == 0 and alignment < 5)
    if longest_ok and spread_ok:
        return [set(idx)]
    out: List[Set[int]] = []
    for comp in _components(idx, [(u, v) for u, v, _ in tree[1:]]):            # cut the most deviating edge, recurse
        out.extend(split_text_region(bboxes, comp, gamma, sigma))
    return out


def _majority_direction(lines: Sequence[Quadrilateral]) -> str:
    top = Counter(q.direction for q in lines).most_common(2)
    if len(top) == 1 or top[0][1] != top[1][1]:
        return top[0][0]
    best, direction = -100.0, top[0][0]                  # tie: the most elongated line decides
    for q in lines:
        if q.aspect_ratio > best:
            best, direction = q.aspect_ratio, q.direction
        if 1.0 / q.aspect_ratio > best:
            best, direction = 1.0 / q.aspect_ratio, q.direction
    return direction


def merge_text_regions(bboxes: Sequence[Quadrilateral], width: int, height: int, engine=None):
    """Yields (line indices in reading order, fg colour, bg colour, direction) per text region (`:110-181`).  With `engine` the O(n^2)
    pair predicate runs on the device (`mitb_op_textline_pairs`); the graph work stays here."""
    n = len(bboxes)
    nodes = list(range(n))
    if engine is not None and n >= 2:
        adj = can_merge_matrix(bboxes, engine, aspect_ratio_tol=1.3, font_size_ratio_tol=2, char_gap_tolerance=1, char_gap_tolerance2=3)
        edges = [(int(u), int(v)) for u, v in np.argwhere(np.triu(adj, 1))]
    else:
        edges = [(u, v) for u, v in itertools.combinations(nodes, 2)
                 if can_merge_region(bboxes[u], bboxes[v], aspect_ratio_tol=1.3, font_size_ratio_tol=2, char_gap_tolerance=1,
                                     char_gap_tolerance2=3)]
    regions: List[Set[int]] = []
    for comp in _components(nodes, edges):
        regions.extend(split_text_region(bboxes, comp))
    for reg in regions:
        members = list(reg)
        lines = [bboxes[i] for i in members]
        fg = tuple(int(round(float(np.mean([getattr(q, c) for q in lines])))) for c in ("fg_r", "fg_g", "fg_b"))
        bg = tuple(int(round(float(np.mean([getattr(q, c) for q in lines])))) for c in ("bg_r", "bg_g", "bg_b"))
        direction = _majority_direction(lines)
        if direction == "h":
            members = sorted(members, key=lambda i: bboxes[i].centroid[1])
        elif direction == "v":
            members = sorted(members, key=lambda i: -bboxes[i].centroid[0])
        yield members, fg, bg, direction


@dataclass
class TextRegion:
    """What the reference packs into a TextBlock at this point (`dispatch` :183-208); rendering fields are not our concern."""
    lines: List[np.ndarray]
    texts: List[str]
    font_size: int
    angle: float
    prob: float
    fg_color: Tuple[int, int, int]
    bg_color: Tuple[int, int, int]
    direction: str
    line_indices: List[int] = field(default_factory=list)


def dispatch(textlines: Sequence[Quadrilateral], width: int, height: int, engine=None) -> List[TextRegion]:
    total_area = sum(q.area for q in textlines)
    out: List[TextRegion] = []
    for members, fg, bg, direction in merge_text_regions(textlines, width, height, engine):
        lines = [textlines[i] for i in members]
        logp = sum(np.log(q.prob) * q.area for q in lines) / total_area      # normalised by the area of ALL lines, as the reference
        angle = float(np.rad2deg(np.mean([q.angle for q in lines])) - 90)
        if abs(angle) < 3:
            angle = 0.0
        out.append(TextRegion([q.pts for q in lines], [q.text for q in lines], int(min(q.font_size for q in lines)), angle,
                              float(np.exp(logp)), fg, bg, direction, list(members)))
    return out
