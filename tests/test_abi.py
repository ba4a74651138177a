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
    assert lib.scgpuGetApiVersion() == 2


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


_FUZZ_CHILD = r'''
import ctypes as C, mmap, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import scgpu
lib = scgpu.load_library()
libc = C.CDLL(None, use_errno=True)
libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
PAGE = mmap.PAGESIZE
g = np.load(sys.argv[2])
rng = np.random.default_rng(17)
accepted = refused = 0
for f in range(int(g["n_files"])):
    base = np.ascontiguousarray(g[f"f{f}_bytes"]).astype(np.uint8)
    pages = (len(base) + PAGE - 1) // PAGE + 1
    mm = mmap.mmap(-1, pages * PAGE)
    addr = C.addressof(C.c_char.from_buffer(mm))
    assert libc.mprotect(addr + (pages - 1) * PAGE, PAGE, 0) == 0      # PROT_NONE guard behind the image
    end = (pages - 1) * PAGE
    for it in range(1500):
        raw = base.copy()
        kind = it % 5
        if kind == 0:
            raw = raw[: int(rng.integers(0, len(raw) + 1))]
        elif kind == 1:
            raw[int(rng.integers(0, len(raw)))] = rng.integers(0, 256)
        elif kind == 2:                                                  # a header / chunk field replaced
            at = 4 * int(rng.integers(0, 12))
            v = int(rng.integers(0, 1 << 32)) if rng.random() < 0.5 else int(rng.integers(0, 400))
            raw[at:at + 4] = np.frombuffer(np.uint32(v).tobytes(), np.uint8)
        elif kind == 3:                                                  # truncated AND a field replaced
            raw = raw[: int(rng.integers(16, len(raw) + 1))].copy()
            at = 4 * int(rng.integers(0, min(12, len(raw) // 4)))
            raw[at:at + 4] = np.frombuffer(np.uint32(int(rng.integers(0, 1 << 32))).tobytes(), np.uint8)
        else:
            raw = rng.integers(0, 256, int(rng.integers(0, 200)), dtype=np.uint8)
            if len(raw) >= 4 and it % 2:
                raw[:4] = np.frombuffer(b"SECT", np.uint8)
        n = len(raw)
        mm[end - n:end] = raw.tobytes()                                  # image flush against the guard page
        xz = (C.c_int32 * 2)()
        ver, cnt = C.c_uint32(0), C.c_uint32(0)
        ok = lib.scgpuSectorFileInfo(C.c_void_p(addr + end - n), n, xz, C.byref(ver), C.byref(cnt))
        if ok:
            accepted += 1
            assert n >= 16
        else:
            refused += 1
print("FUZZ OK", accepted, refused)
'''


def test_sector_file_parser_survives_corrupt_images():
    """.scsector images are untrusted input: truncated, bit-rotted and random images placed flush against a PROT_NONE
    page must be parsed or refused without one byte read past the image (a stray read kills the child process). The
    reference's own reader is no yardstick here — it accepts cut-off files (unread records stay value-initialised) and
    dies with std::bad_alloc on a corrupt instance count (world_format.cpp:213-216); the library refuses both."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", _FUZZ_CHILD, str(ROOT / "sc-gameengine_b200"),
                        str(Path(__file__).resolve().parent / "golden" / "sector_files.npz")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FUZZ OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    accepted, refused = (int(x) for x in r.stdout.split("FUZZ OK")[1].split())
    assert accepted > 100 and refused > 100
