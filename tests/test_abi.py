"""CPU: the C-ABI library loads and exports every symbol include/stdadk.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "stdadk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stdadk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "st_dadk_b200", "libstdadk.so"))
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stdadk.h but not exported"
    lib.stdadk_version.restype = ctypes.c_int
    assert lib.stdadk_version() == 100
    lib.stdadk_image_floats.restype = ctypes.c_size_t
    lib.stdadk_image_floats.argtypes = [ctypes.c_int64, ctypes.c_int64]
    assert lib.stdadk_image_floats(130, 33) == 2 * 2 * 4096


def test_ctypes_structs_match_header_sizes():
    from st_dadk_b200 import _lib as L
    L.lib()   # raises on any sizeof mismatch between the ctypes structs and the compiled header
    assert ctypes.sizeof(L.Basis) == 32 and ctypes.sizeof(L.Points) == 64
    assert ctypes.sizeof(L.Layer) == 56 and ctypes.sizeof(L.Dropout) == 32
    assert ctypes.sizeof(L.Head) == 24 + 8 + 32 + 16 + 24
    assert ctypes.sizeof(L.FwdArgs) == 8 + 64 + 8 + 56 + 32 + 48 + 16


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle (or /root/reference)."""
    bad = []
    for base in ("st_dadk_b200", "stnf", "scripts"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    s = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M) or "/root/reference" in s:
                        if not f.endswith("selftest.py"):
                            bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_sass_is_blackwell_native():
    """The compiled kernels use the sm_100a tensor-core / TMEM / TMA instructions and no legacy tensor path:
    `tcgen05.mma` -> UTC*MMA, `tcgen05.ld` -> LDTM, bulk async copies -> UBLKCP, mbarrier -> SYNCS, packed FP32 ->
    FFMA2 (B200_PROFILING.md, "What proves a Blackwell-native kernel").  Skipped when cuobjdump is not installed."""
    import shutil
    import subprocess
    import pytest
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    import __graft_entry__ as g
    g.build()
    lib = os.path.join(ROOT, "st_dadk_b200", "libstdadk.so")
    arch = subprocess.run([exe, "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in arch, arch
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True).stdout
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTC[A-Z]*MMA\b") >= 50        # tcgen05.mma in forward, backward, wgrad, knot-grad, fused kernels
    assert count(r"\bLDTM\b") >= 20                # accumulators read back from tensor memory
    assert count(r"\bUBLKCP\b") >= 50              # TMA 1-D bulk copies of operand slabs and knot tables
    assert count(r"\bSYNCS\b") >= 100              # mbarrier pipelines
    assert count(r"\bFFMA2\b") >= 100              # packed FP32 epilogue math
    assert count(r"\bHMMA\b") == 0 and count(r"\b[HQI]GMMA\b") == 0      # no mma.sync / wgmma
    for kernel in ("layer_fwd_kernel", "layer_bwd_kernel", "wgrad_kernel", "knotgrad_kernel", "predict_fused_kernel",
                   "sparse_spatial_fwd_kernel", "sparse_spatial_wgrad_kernel", "adamw_ema_kernel", "sqnorm_kernel",
                   "peer_allreduce_kernel", "pack_images_kernel"):
        assert kernel in sass, kernel
