"""Golden vectors for the motion-estimation row (F3): executes the reference's OWN
`_process_block_row` (src/IPP_DCT.py:207-246), taken from the unmodified source file at run
time (nothing is copied into this repository), on small seeded frame pairs.

    python oracle/make_golden_me.py        (needs /root/reference; writes tests/golden/ref_me_*.npz)
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/IPP_DCT.py"


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("_three_step_search", "_process_block_row"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns["_process_block_row"]


def frame_pair(rng, h, w, kind):
    base = rng.integers(0, 256, size=(h + 32, w + 32), dtype=np.uint8)
    if kind == "shift":          # global displacement plus noise: a unique minimum almost everywhere
        dy, dx = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
        ref = base[16:16 + h, 16:16 + w].copy()
        cur = base[16 + dy:16 + dy + h, 16 + dx:16 + dx + w].astype(np.int16) + rng.integers(-2, 3, size=(h, w))
        return ref, np.clip(cur, 0, 255).astype(np.uint8)
    if kind == "flat":           # constant areas: every candidate ties, the scan order decides
        ref = np.full((h, w), 90, dtype=np.uint8)
        cur = np.full((h, w), 93, dtype=np.uint8)
        ref[h // 3:h // 2, w // 4:w // 2] = 200
        cur[h // 3 + 2:h // 2 + 2, w // 4 - 3:w // 2 - 3] = 200
        return ref, cur
    ref = base[:h, :w].copy()    # unrelated noise
    cur = base[16:16 + h, 16:16 + w].copy()
    return ref, cur


def main():
    row_fn = reference_functions()
    out = os.path.join(ROOT, "tests", "golden")
    rng = np.random.default_rng(77)
    cases = [("shift", 64, 96, 16, 8), ("flat", 48, 80, 16, 8), ("noise", 40, 56, 8, 4), ("shift", 70, 100, 16, 6),
             ("flat", 64, 64, 32, 8), ("noise", 32, 48, 4, 3)]
    for n, (kind, h, w, bs, sr) in enumerate(cases):
        ref, cur = frame_pair(rng, h, w, kind)
        mv = np.zeros((h // bs, w // bs, 2), dtype=np.float32)
        for i in range(0, h - bs + 1, bs):                       # src/IPP_DCT.py:357-371
            ri, row = row_fn((ref, cur, i, bs, sr, w, False))
            for col, v in enumerate(row):
                mv[ri // bs, col] = v
        mvf = np.zeros((h // bs, w // bs, 2), dtype=np.float32)    # the same pair with --fast (three-step search)
        for i in range(0, h - bs + 1, bs):
            ri, row = row_fn((ref, cur, i, bs, sr, w, True))
            for col, v in enumerate(row):
                mvf[ri // bs, col] = v
        np.savez_compressed(os.path.join(out, f"ref_me_{n}_{kind}.npz"), ref=ref, cur=cur, bs=bs, sr=sr, mv=mv, mv_fast=mvf)
        print(kind, h, w, bs, sr, "->", mv.reshape(-1, 2)[:4].tolist(), "| fast:", mvf.reshape(-1, 2)[:4].tolist())


if __name__ == "__main__":
    sys.exit(main())
