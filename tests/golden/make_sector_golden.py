#!/usr/bin/env python
"""Generates tests/golden/sector_files.npz from the REFERENCE ITSELF: .scsector files written by the reference's own
sc_world::WriteSectorFile (tools/shared/world_format.cpp:76-176) in format versions 1, 3 and 4 (three INST record
layouts), and what its own sc_world::ReadSectorFile (:178-334) reads back from them. Run in the build container:

    python tests/golden/make_sector_golden.py

The fixture holds the raw file bytes (derived data, written by the reference) and the reader's output."""
import ctypes as C
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "sc-gameengine_b200"))
import oracle_bind  # noqa: E402
from scenarios import random_trs  # noqa: E402


def f(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    L = oracle_bind.ref_lib()
    L.screfWriteSectorFile.restype = C.c_int
    L.screfWriteSectorFile.argtypes = [C.c_char_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32]
    L.screfReadSectorInstances.restype = C.c_int
    L.screfReadSectorInstances.argtypes = [C.c_char_p, C.c_uint32] + [C.c_void_p] * 5
    L.screfHashAssetPath.restype = C.c_uint64
    L.screfHashAssetPath.argtypes = [C.c_char_p]
    rng = np.random.default_rng(4242)
    names = [b"meshes/cube", b"meshes/triangle", b"materials/unlit", b"materials/checker", b"materials/test"]
    ids = np.array([L.screfHashAssetPath(n) for n in names], np.uint64)
    out = {"asset_names": np.array([n.decode() for n in names]), "asset_ids": ids}
    with tempfile.TemporaryDirectory() as d:
        for k, (version, n, extra, xz) in enumerate([(4, 300, 1, (3, -2)), (3, 41, 0, (-7, 5)), (1, 17, 1, (0, 0)), (4, 0, 1, (1, 1))]):
            path = str(Path(d) / f"s{k}.scsector").encode()
            trs = random_trs(rng, n, spread=64.0)
            mesh = ids[rng.integers(0, 2, n)]
            mat = ids[2 + rng.integers(0, 3, n)]
            if n > 5:
                mesh[3] = 0                      # assetId 0 -> handle 0 (sc_world_partition.cpp:748-749)
                mat[4] = np.uint64(0x1234567890)  # unknown id -> the default material (:774-775)
            iid = rng.integers(1, 1 << 62, n, dtype=np.uint64)
            model = rng.integers(0, 1 << 62, n, dtype=np.uint64)
            tags = rng.integers(0, 1 << 31, n, dtype=np.uint32)
            assert L.screfWriteSectorFile(path, version, xz[0], xz[1], n, f(iid), f(model), f(mesh), f(mat), f(trs), f(tags), extra)
            raw = np.frombuffer(Path(path.decode()).read_bytes(), np.uint8).copy()
            cap = max(n, 1)
            oxz = np.zeros(2, np.int32); oid = np.zeros(cap, np.uint64); omesh = np.zeros(cap, np.uint64)
            omat = np.zeros(cap, np.uint64); otrs = np.zeros((cap, 9), np.float32)
            got = L.screfReadSectorInstances(path, cap, f(oxz), f(oid), f(omesh), f(omat), f(otrs))
            assert got == n, (got, n)
            out[f"f{k}_bytes"], out[f"f{k}_version"], out[f"f{k}_xz"] = raw, np.uint32(version), oxz
            out[f"f{k}_id"], out[f"f{k}_mesh"], out[f"f{k}_mat"], out[f"f{k}_trs"] = oid[:n], omesh[:n], omat[:n], otrs[:n]
            print(f"file {k}: version {version}, {n} instances, {raw.size} bytes")
    out["n_files"] = np.uint32(4)
    np.savez_compressed(HERE / "sector_files.npz", **out)
    print("written", HERE / "sector_files.npz")


if __name__ == "__main__":
    main()
