"""The C-ABI library loads and exports every symbol include/scgpu.h declares; without a GPU it fails loudly
(no CPU fallback). No compute calls are made here."""
import ctypes as C
import re
from pathlib import Path

import pytest

import scgpu

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "scgpu.h").read_text()
    return sorted(set(re.findall(r"SCGPU_API\s+[\w\s\*]+?\b(scgpu\w+)\s*\(", text)))


def test_header_declares_the_survey_boundary():
    names = declared_symbols()
    for must in ("scgpuGetApiVersion", "scgpuCreate", "scgpuDestroy", "scgpuSpawn", "scgpuDespawn", "scgpuSetLocal",
                 "scgpuSetParent", "scgpuSetViews", "scgpuSetViewPlanes", "scgpuUpdate", "scgpuGetCounts",
                 "scgpuReadVisible", "scgpuReadDrawItems", "scgpuReadWorld", "scgpuLastError"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(scgpu.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libscgpu.so lacks {n}"
    assert sorted(scgpu.SYMBOLS) == names, "python binding and header disagree"


def test_struct_layouts_match_the_reference_records():
    assert C.sizeof(scgpu.SceneDesc) == 32
    assert scgpu.DRAW_ITEM_DTYPE.itemsize == 80  # sc::DrawItem, sc_ecs.h:159-165
    assert scgpu.DRAW_ITEM_DTYPE.fields["model"][1] == 16
    lib = scgpu.load_library()
    assert lib.scgpuGetApiVersion() == 1


def test_no_cpu_fallback():
    """Without a CUDA device scgpuCreate must fail with a message, never fall back to a host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(scgpu.ScGpuError) as ei:
        scgpu.Scene(1024, max_views=1)
    assert "no CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)
    lib = scgpu.load_library()
    assert lib.scgpuCreate(None) is None
    assert b"NULL" in lib.scgpuLastError(None)


def test_product_never_links_the_oracle():
    """libscgpu.so must not depend on anything under oracle/ (ldd-level check) and the package must not import it."""
    import subprocess
    out = subprocess.run(["ldd", str(scgpu.LIB_PATH)], capture_output=True, text=True).stdout
    assert "scoracle" not in out and "scref" not in out
    for py in (ROOT / "sc-gameengine_b200" / "scgpu").glob("*.py"):
        assert "oracle" not in py.read_text().replace("the oracle", "").replace("oracle restates", ""), py


def test_sector_file_info_matches_the_reference_reader():
    """SURVEY 8(f) N3, host half: the chunk walk over .scsector images written by the reference's WriteSectorFile
    (tests/golden/sector_files.npz, made by make_sector_golden.py) yields the coordinate, version and instance count
    the reference's ReadSectorFile read; malformed images are refused. No GPU involved."""
    import ctypes as C
    import numpy as np
    import scgpu
    lib = scgpu.load_library()
    g = np.load(Path(__file__).resolve().parent / "golden" / "sector_files.npz")
    for k in range(int(g["n_files"])):
        raw = np.ascontiguousarray(g[f"f{k}_bytes"])
        xz = np.zeros(2, np.int32)
        ver, cnt = C.c_uint32(0), C.c_uint32(0)
        assert lib.scgpuSectorFileInfo(raw.ctypes.data_as(C.c_void_p), raw.size, xz.ctypes.data_as(C.c_void_p), C.byref(ver), C.byref(cnt)) == 1
        assert (list(xz), ver.value, cnt.value) == (list(g[f"f{k}_xz"]), int(g[f"f{k}_version"]), len(g[f"f{k}_id"]))
    raw = np.ascontiguousarray(g["f0_bytes"]).copy()
    bad = raw.copy(); bad[0] ^= 0xFF
    assert lib.scgpuSectorFileInfo(bad.ctypes.data_as(C.c_void_p), bad.size, None, None, None) == 0      # magic
    assert lib.scgpuSectorFileInfo(raw.ctypes.data_as(C.c_void_p), 200, None, None, None) == 0           # cut short
    assert lib.scgpuSectorFileInfo(raw.ctypes.data_as(C.c_void_p), 8, None, None, None) == 0
