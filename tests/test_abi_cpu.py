"""The C-ABI library loads without a GPU and exports every symbol include/kicp.h declares; host-side helpers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _build()
    from kinematic_icp_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "kicp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kicp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = C.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libkicp_b200.so does not export %s" % name
    assert declared == {s[0] for s in _capi.SYMBOLS}, "python binding and header disagree"


def test_every_device_kernel_is_the_librarys_own():
    """No library kernels on any path of the product: the only device functions in libkicp_b200.so are this repo's own `k_*` kernels —
    no cub:: / thrust:: instantiations (the selects, reductions and scans of the frame path are kicp_scan.cuh fused into the kernels
    that decide) — and the library links neither cuBLAS nor any other device library beside the CUDA runtime."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    _build()
    from kinematic_icp_b200 import _capi
    listing = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels = re.findall(r"Function : (\S+)", listing)
    assert len(kernels) >= 20
    foreign = [k for k in kernels if not re.match(r"_Z\d+k_[a-z0-9_]+", k)]
    assert not foreign, foreign
    needed = subprocess.run(["readelf", "-d", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    libs = re.findall(r"\(NEEDED\)\s+Shared library: \[(.*?)\]", needed)
    assert not [l for l in libs if re.search(r"cublas|cusolver|cusparse|cufft|curand|nvrtc|cudnn", l)], libs


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every struct of include/kicp.h as gcc lays it out == the ctypes mirror (and the header is plain C)."""
    import subprocess
    from kinematic_icp_b200 import _capi
    structs = {"kicp_reg_params": _capi.RegParams, "kicp_reg_result": _capi.RegResult, "kicp_profile": _capi.Profile,
               "kicp_frame_input": _capi.FrameInput, "kicp_frame_params": _capi.FrameParams}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "kicp.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname + ".sizeof"]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
    assert C.sizeof(_capi.RegParams) == 24
    assert C.sizeof(_capi.RegResult) == 7 * 8 + 8 + 8 + 4 + 4 + 64 * 8 * 8 + 64 * 2 * 8
    assert _capi.KICP_MAX_ITERATIONS == 64 and _capi.KICP_DTYPE_F32 == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product fails loudly instead of computing on the host."""
    import kinematic_icp_b200 as kb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(kb.KicpError) as e:
        kb.Context(0)
    assert e.value.status == kb._capi.KICP_ERR_CUDA


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "kinematic-icp_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "kicp_oracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), os.path.join(base, f)


def test_shard_range_partitions_the_scan():
    import kinematic_icp_b200 as kb
    for n in (0, 1, 7, 261675):
        for g in (1, 2, 4, 8):
            r = [kb.shard_range(n, g, k) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(g - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_null_arguments_are_rejected_before_any_device_work():
    """Every entry point validates its arguments first: KICP_ERR_INVALID for null handles, with or without a GPU."""
    _build()
    from kinematic_icp_b200 import _capi
    L = _capi.lib()
    INVALID = _capi.KICP_ERR_INVALID
    z7 = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    pts = np.zeros((4, 3))
    out = np.zeros((4, 3))
    m = C.c_int64()
    p = _capi.RegParams(10, 1, 1e-3, 0.0)
    dp = _capi.dp
    assert L.kicp_map_create(None, 1.0, 100.0, 20, C.byref(C.c_void_p())) == INVALID
    assert L.kicp_map_reserve(None, 1000) == INVALID
    assert L.kicp_map_clear(None) == INVALID
    assert L.kicp_map_add_points(None, dp(pts), 4) == INVALID
    assert L.kicp_map_update_pose(None, dp(pts), 4, dp(z7)) == INVALID
    assert L.kicp_map_nearest(None, dp(pts), 4, dp(out), dp(np.zeros(4))) == INVALID
    assert L.kicp_register(None, dp(pts), 4, dp(z7), dp(z7), 1.0, C.byref(p), dp(np.zeros(7)), None) == INVALID
    assert L.kicp_register_sharded(None, dp(pts), 4, dp(z7), dp(z7), 1.0, C.byref(p), dp(np.zeros(7)), None) == INVALID
    assert L.kicp_voxel_downsample(None, dp(pts), 4, 0.5, dp(out), 4, C.byref(m)) == INVALID
    assert L.kicp_preprocess(None, dp(pts), 4, None, 0, dp(z7), dp(z7), 100.0, 0.0, 0, dp(out), 4, C.byref(m)) == INVALID
    fi, fp = _capi.FrameInput(), _capi.FrameParams()
    fp.voxel_size = 1.0
    assert L.kicp_register_frame(None, C.byref(fi), dp(z7), dp(z7), dp(z7), dp(z7), 1.0, C.byref(fp), dp(np.zeros(7)), None, 0, None, None,
                                 0, None, None) == INVALID
    assert L.kicp_frame_clouds(None, None, None, None, None) == INVALID
    assert L.kicp_ctx_set_option(None, b"assoc_variant", 1) == INVALID
    assert L.kicp_scan_create(None, 16, C.byref(C.c_void_p())) == INVALID
    assert L.kicp_comm_p2p_handle(None, None) == INVALID
    assert L.kicp_status_string(INVALID) and L.kicp_status_string(_capi.KICP_WARN_NO_CORRESPONDENCES)
