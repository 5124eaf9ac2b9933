"""The C-ABI library loads on a CPU-only box and exports every symbol include/gq_b200.h declares.
No compute kernels run here (no GPU): only argument validation paths that return before any
CUDA call, and the host-side stream-K planner."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "gq_b200.h"
LIB = ROOT / "vq-vae-from-gaussian-vae_b200" / "gq_b200" / "libgq_b200.so"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(gq_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    assert LIB.exists(), "libgq_b200.so is not built (python vq-vae-from-gaussian-vae_b200/build_ext.py)"
    return ctypes.CDLL(str(LIB))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ["gq_scores_f32", "gq_codebook_create", "gq_codebook_destroy", "gq_quantize", "gq_quantize_exact",
                 "gq_quantize_workspace_bytes", "gq_dequant", "gq_kl_fwd", "gq_kl_bwd", "gq_kl_workspace_bytes",
                 "gq_abi_version", "gq_last_error", "gq_launch_count", "gq_index_width", "gq_pack_indices",
                 "gq_unpack_indices", "gq_dequant_packed", "gq_index_histogram", "gq_usage_entropy"]:
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in gq_b200.h but not exported"


def test_python_binding_matches_header(lib):
    import gq_b200._lib as b
    assert sorted(b.PROTOTYPES) == declared_functions()


def test_abi_version(lib):
    ver = int(re.search(r"#define GQ_B200_ABI_VERSION (\d+)", HEADER.read_text()).group(1))
    lib.gq_abi_version.restype = ctypes.c_int
    assert lib.gq_abi_version() == ver


def test_argument_validation_without_gpu(lib):
    lib.gq_last_error.restype = ctypes.c_char_p
    f = lib.gq_scores_f32
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_double, ctypes.c_void_p]
    assert f(None, None, None, None, 0, 4, 4, 1.0, None) == -1          # dim <= 0
    assert b"bad sizes" in lib.gq_last_error()
    assert f(None, None, None, None, 16, 0, 4, 1.0, None) == 0          # empty problem is a no-op
    assert f(None, None, None, None, 16, 4, 4, 1.0, None) == -1         # null pointers
    d = lib.gq_dequant
    d.restype = ctypes.c_int
    d.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    assert d(None, None, None, 1, 1, 1, 0, None, None) == -1
    c = lib.gq_codebook_create
    c.restype = ctypes.c_int
    c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    h = ctypes.c_void_p()
    assert c(None, None, 16, 16, 1.0, None, ctypes.byref(h)) == -1       # null codebook
    assert c(ctypes.c_void_p(16), None, 16, 3, 1.0, None, ctypes.byref(h)) == -1   # unsupported group
    assert b"group 3 unsupported" in lib.gq_last_error()
    q = lib.gq_quantize
    q.restype = ctypes.c_int
    q.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    assert q(None, None, None, None, 4, None, None, None, 0, None, None) == -1    # null handle


def pieces_of(cta, tiles, spc, total):
    """Python mirror of next_piece() in csrc/gq_tc.cu."""
    step, end = cta * spc, min((cta + 1) * spc, total)
    out = []
    while step < end:
        unit, t0 = divmod(step, tiles)
        t1 = min(tiles, t0 + (end - step))
        out.append((unit, t0, t1, cta - (unit * tiles) // spc))
        step += t1 - t0
    return out


@pytest.mark.parametrize("N,M", [(65536, 65536), (65536, 1024), (65536, 1), (4096, 300), (1000, 7), (262144, 4096),
                                 (65536, 1048576), (4096, 257), (128, 256), (65536, 37 * 256 + 5)])
@pytest.mark.parametrize("sms", [148, 132, 1])
@pytest.mark.parametrize("acc16", [0, 1], ids=["fp32acc-256x128", "fp16acc-128x256"])
def test_streamk_plan_covers_every_step_once(lib, N, M, sms, acc16):
    lib.gq_debug_plan.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    out = (ctypes.c_int64 * 8)()
    lib.gq_debug_plan(N, M, sms, acc16, out)
    units, tiles, ctas, spc, nseg, m_pad, unit_rows, tile_cols = list(out)
    assert (unit_rows, tile_cols) == ((128, 256) if acc16 else (256, 128))
    assert m_pad == -(-M // 256) * 256 and units == m_pad // unit_rows and tiles == -(-N // tile_cols)
    assert 1 <= ctas <= sms
    total = units * tiles
    assert (ctas - 1) * spc < total <= ctas * spc
    seen = {}
    for c in range(ctas):
        for unit, t0, t1, seg in pieces_of(c, tiles, spc, total):
            assert 0 <= seg < nseg and 0 <= t0 < t1 <= tiles
            for t in range(t0, t1):
                assert (unit, t) not in seen
                seen[(unit, t)] = seg
    assert len(seen) == total
    # segments of one unit are numbered 0..k-1 without gaps (rescore_kernel relies on it)
    for u in {0, units // 2, units - 1}:
        segs = sorted({seen[(u, t)] for t in range(tiles)})
        assert segs == list(range(len(segs)))
        first, last = (u * tiles) // spc, (u * tiles + tiles - 1) // spc
        assert len(segs) == last - first + 1


def test_index_stream_entry_points_validate_without_gpu(lib):
    lib.gq_last_error.restype = ctypes.c_char_p
    w = lib.gq_index_width
    w.restype, w.argtypes = ctypes.c_int, [ctypes.c_int64]
    assert [w(n) for n in (1, 4096, 65536, 65537, 1 << 18)] == [2, 2, 2, 4, 4]
    pk = lib.gq_pack_indices
    pk.restype, pk.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    assert pk(None, 0, 65536, None, None, None) == 0            # empty stream is a no-op
    assert pk(None, 8, 65536, None, None, None) == -1           # null pointers
    assert pk(None, 8, 0, None, None, None) == -1               # no codebook
    assert pk(None, 8, 1 << 31, None, None, None) == -1         # does not fit 32 bits
    up = lib.gq_unpack_indices
    up.restype, up.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    assert up(None, 0, 65536, None, None) == 0 and up(None, 3, 65536, None, None) == -1
    h = lib.gq_index_histogram
    h.restype, h.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    assert h(None, 4, 16, None, None, None) == -1 and b"hist is null" in lib.gq_last_error()
    u = lib.gq_usage_entropy
    u.restype, u.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    assert u(None, 16, None, None) == -1
    d = lib.gq_dequant_packed
    d.restype = ctypes.c_int
    d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    assert d(None, None, None, 1, 1, 1, 0, None, None) == -1 and b"handle is null" in lib.gq_last_error()
    m = lib.gq_codebook_set_mode
    m.restype, m.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
    assert m(None, 0) == -1 and b"handle is null" in lib.gq_last_error()
    u2 = lib.gq_codebook_set_usage_histogram
    u2.restype, u2.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]
    assert u2(None, None) == -1 and b"handle is null" in lib.gq_last_error()
