#!/usr/bin/env python
"""Golden vectors for SURVEY 8(f) N1 from the reference's own draw submission block (src/engine/src/sc_vk.cpp:1841-1912,
compiled into oracle/_ref by oracle/ref_shim/scref_renderer.cpp):  python tests/golden/make_renderer_golden.py
Writes tests/golden/renderer_sort.npz: draw ids + material table in, submission order keys and bind points out.
(std::sort is unstable: the ORDER OF EQUAL KEYS is not part of the golden, the key sequence and the bind points are.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
from oracle_bind import DRAW_ITEM_DTYPE, ref_renderer_submit  # noqa: E402

rng = np.random.default_rng(20261018)
n, n_mat, n_mesh = 6000, 23, 9
d = np.zeros(n, DRAW_ITEM_DTYPE)
d["entity"] = np.arange(n)
d["meshId"] = rng.integers(0, n_mesh + 2, n)
d["materialId"] = rng.integers(0, n_mat + 2, n)
mp = rng.integers(0, 2, n_mat).astype(np.uint32)
mp[[2, 11]] = 0xFFFFFFFF
order, binds = ref_renderer_submit(d, mp, n_mesh)
np.savez_compressed(ROOT / "tests" / "golden" / "renderer_sort.npz", mesh=d["meshId"], material=d["materialId"],
                    material_pipeline=mp, mesh_count=np.uint32(n_mesh),
                    out_mesh=d["meshId"][order], out_material=d["materialId"][order], out_binds=binds,
                    out_kept=np.sort(order))
print("submitted", len(order), "of", n, "draws;", int((binds != 0).sum()), "bind points")
