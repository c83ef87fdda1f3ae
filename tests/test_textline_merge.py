"""Text-line merge against the reference's own known-answer tests: tests/golden/textline_merge.json holds the
quadrilaterals and expected groupings of manga_translator's test/test_textline_merge.py (extracted by oracle/make_merge_golden.py)."""
import json
import os

import numpy as np
import pytest

from mit_b200.host import textline_merge
from mit_b200.host.geometry import Quadrilateral

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "textline_merge.json")
CASES = json.load(open(GOLDEN))["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_merge_matches_reference_known_answers(case):
    quads = [Quadrilateral(np.array(l), "", 1) for l in case["lines"]]
    regions = textline_merge.dispatch(quads, case["width"], case["height"])
    got = {tuple(sorted(r.line_indices)) for r in regions}
    want = {tuple(c) for c in case["expected"]}
    assert got == want
    assert sorted(i for r in regions for i in r.line_indices) == list(range(len(quads)))      # a partition of the lines


def test_merge_region_fields_and_ordering():
    # three stacked horizontal lines + one far-away vertical line
    lines = [[[100, 100], [400, 100], [400, 140], [100, 140]], [[100, 150], [400, 150], [400, 190], [100, 190]],
             [[100, 200], [380, 200], [380, 240], [100, 240]], [[900, 100], [940, 100], [940, 500], [900, 500]]]
    quads = [Quadrilateral(np.array(l), f"t{i}", 0.9, 10 * i, 0, 0, 255, 255, 250) for i, l in enumerate(lines)]
    for q in quads:
        q.assigned_direction = q.direction
    regions = textline_merge.dispatch(quads[::-1], 1000, 600)          # shuffled input order
    by_size = sorted(regions, key=lambda r: -len(r.lines))
    assert [len(r.lines) for r in by_size] == [3, 1]
    block = by_size[0]
    assert block.direction == "h" and block.texts == ["t0", "t1", "t2"]                      # top to bottom
    assert block.font_size == 40 and block.angle == 0.0
    assert block.fg_color == (10, 0, 0) and block.bg_color == (255, 255, 250)
    assert 0 < block.prob <= 1 and by_size[1].direction == "v"
    assert textline_merge.dispatch([], 10, 10) == []
