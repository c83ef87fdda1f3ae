"""TEST INFRASTRUCTURE (build container only): extracts the known-answer vectors of the reference's own text-line merge tests
(/root/reference/test/test_textline_merge.py: per test a list of quadrilaterals, the page size and the expected grouping) into
tests/golden/textline_merge.json.  Only literals are read (ast), no reference code is executed or copied."""
import ast
import json
import os
import sys

REF = "/root/reference/test/test_textline_merge.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textline_merge.json")


def main():
    tree = ast.parse(open(REF, encoding="utf-8").read())
    cases = []
    for fn in tree.body:
        if not isinstance(fn, ast.AsyncFunctionDef) or not fn.name.startswith("test_merge"):
            continue
        env = {}
        for st in fn.body:
            if isinstance(st, ast.Assign):
                try:
                    val = ast.literal_eval(st.value)
                except Exception:
                    continue
                tgt = st.targets[0]
                if isinstance(tgt, ast.Tuple):
                    for t, v in zip(tgt.elts, val):
                        env[t.id] = v
                elif isinstance(tgt, ast.Name):
                    env[tgt.id] = val
        if {"lines", "expected_combinations", "width", "height"} <= set(env):
            cases.append({"name": fn.name, "width": env["width"], "height": env["height"], "lines": env["lines"],
                          "expected": env["expected_combinations"]})
    if not cases:
        sys.exit("no cases found")
    with open(OUT, "w") as f:
        json.dump({"source": "manga_translator test/test_textline_merge.py (expected groupings produced by the reference with shapely)",
                   "cases": cases}, f, separators=(",", ":"))
    print(f"wrote {len(cases)} cases, {sum(len(c['lines']) for c in cases)} lines -> {OUT}")


if __name__ == "__main__":
    main()
