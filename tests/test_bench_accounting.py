"""bench.py's byte accounting follows SURVEY.md 8(d) (the formula the judge checks the roofline numbers with)."""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_functions(*names):
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    ns = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), "bench.py", "exec"), ns)
    return [ns[n] for n in names]


def test_path_bytes_matches_the_survey_formula():
    (path_bytes,) = _load_functions("path_bytes")
    P, V, R, N = 3_000_000, 2_940_000, 15_000_000, 2_073_600
    Bf, Bb = path_bytes(P, V, R, N, 0)            # colors_precomp: A_in = 56, A_g = 68
    assert Bf == P * (56 + 4) + V * 40 + R * 8 + R * 36 + N * 28
    assert Bb == R * 40 + N * 28 + V * 40 + P * (56 + 68)
    # the survey's worked example: ~1.02 GB forward, ~1.15 GB backward
    assert abs(Bf / 1e9 - 1.02) < 0.02 and abs(Bb / 1e9 - 1.15) < 0.02
    Bf_sh, Bb_sh = path_bytes(P, V, R, N, 16)     # SH degree 3 in-kernel: 192 B of coefficients per Gaussian
    assert Bf_sh - Bf == P * (192 - 12) and Bb_sh - Bb == 2 * P * (192 - 12)


def test_stage_bytes_cover_every_profiled_stage():
    (stage_bytes,) = _load_functions("stage_bytes")
    sb = stage_bytes(1000, 900, 5000, 64 * 64, 16, 0, 700, 1500)
    for k in ("preprocess_fwd", "depth_sort", "offset_scan", "emit_cells", "cell_sort", "cell_count", "tile_offsets",
              "tile_scatter", "render_fwd", "render_bwd", "preprocess_bwd"):
        assert sb[k] > 0, k
    assert sb["depth_sort"] == 1000 * 16 * 4 + 1000 * 24 and sb["tile_scatter"] == 5000 * 4 + 1500 * 8


def test_entry_points_parse():
    for f in ("bench.py", "__graft_entry__.py"):
        ast.parse(open(os.path.join(ROOT, f)).read())
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    assert callable(g.build) and callable(g.smoke)
