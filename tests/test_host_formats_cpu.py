"""TUM trajectory format of OfflineNode::writePosesInTumFormat (offline_node.cpp:76-97): 'timestamp x y z qx qy qz qw',
fixed notation, 6 decimals."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tum_writer_format(tmp_path):
    import __graft_entry__ as g
    g.build()
    L = C.CDLL(os.path.join(ROOT, "tests", "hooks", "_build", "libkicp_facade_hooks.so"))
    c_dp = C.POINTER(C.c_double)
    L.kfac_write_tum.argtypes = [C.c_char_p, c_dp, c_dp, C.c_int64]
    stamps = np.array([1700000000.123456789, 1700000000.2])
    poses = np.array([[0.0, 0.0, 0.3826834324, 0.9238795325, 1.5, -2.25, 0.0], [0, 0, 0, 1, 10.123456789, 0, 0.5]])
    path = tmp_path / "poses.tum"
    assert L.kfac_write_tum(str(path).encode(), stamps.ctypes.data_as(c_dp), poses.ctypes.data_as(c_dp), 2) == 0
    lines = path.read_text().splitlines()
    assert lines[0] == "1700000000.123457 1.500000 -2.250000 0.000000 0.000000 0.000000 0.382683 0.923880"
    assert lines[1] == "1700000000.200000 10.123457 0.000000 0.500000 0.000000 0.000000 0.000000 1.000000"


def _facade():
    import __graft_entry__ as g
    g.build()
    return C.CDLL(os.path.join(ROOT, "tests", "hooks", "_build", "libkicp_facade_hooks.so"))


def _decode(L, rec, fields, width=None, height=1):
    """fields: [(name, offset, datatype, count)] -> (layout5 or None, stamps)"""
    n = len(rec)
    raw = np.ascontiguousarray(rec).view(np.uint8).reshape(-1)
    L.kfac_pc2_decode.restype = C.c_int64
    L.kfac_pc2_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p]
    names = ",".join(f[0] for f in fields).encode()
    offs = np.array([f[1] for f in fields], dtype=np.uint32)
    types = np.array([f[2] for f in fields], dtype=np.uint8)
    counts = np.array([f[3] for f in fields], dtype=np.uint32)
    layout = np.zeros(5, dtype=np.int32)
    stamps = np.zeros(max(n, 1))
    k = L.kfac_pc2_decode(raw.ctypes.data, width or n, height, rec.dtype.itemsize, names, offs.ctypes.data, types.ctypes.data,
                          counts.ctypes.data, len(fields), layout.ctypes.data, stamps.ctypes.data)
    return (None, None) if k < 0 else (layout, stamps[:k].copy())


UINT32, FLOAT32, FLOAT64 = 6, 7, 8


def test_pointcloud2_layout_and_timestamp_fields():
    """kicp/pointcloud2.hpp against the message handling of RosUtils.cpp:30-39 and TimeStampHandler.cpp:42-104."""
    L = _facade()
    n = 7
    # an Ouster-like record: x y z pad intensity t(uint32 ns since sweep start) ring
    dt = np.dtype({"names": ["x", "y", "z", "intensity", "t", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2"],
                   "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
    rec = np.zeros(n, dtype=dt)
    rec["t"] = np.arange(n) * 1000
    fields = [("x", 0, FLOAT32, 1), ("y", 4, FLOAT32, 1), ("z", 8, FLOAT32, 1), ("intensity", 16, FLOAT32, 1), ("t", 20, UINT32, 1),
              ("ring", 24, 4, 1)]
    layout, stamps = _decode(L, rec, fields)
    assert layout.tolist() == [1, 32, 0, 4, 8]  # KICP_DTYPE_F32, point_step, x/y/z offsets
    assert np.array_equal(stamps, np.arange(n) * 1000.0)  # small integers are seconds as they are (<= 10 digits)
    # absolute nanoseconds (19 digits) are converted to seconds; float64 seconds pass through
    dt2 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("timestamp", "<f8")])
    rec2 = np.zeros(n, dtype=dt2)
    rec2["timestamp"] = 1.7e18 + np.arange(n) * 1e5
    layout, stamps = _decode(L, rec2, [("x", 0, FLOAT32, 1), ("y", 4, FLOAT32, 1), ("z", 8, FLOAT32, 1), ("timestamp", 12, FLOAT64, 1)])
    assert layout.tolist() == [1, 20, 0, 4, 8] and np.allclose(stamps, 1.7e9 + np.arange(n) * 1e-4, rtol=0, atol=1e-6)
    rec2["timestamp"] = 1.7e9 + np.arange(n) * 0.01
    _, stamps = _decode(L, rec2, [("x", 0, FLOAT32, 1), ("y", 4, FLOAT32, 1), ("z", 8, FLOAT32, 1), ("timestamp", 12, FLOAT64, 1)])
    assert np.array_equal(stamps, rec2["timestamp"])
    # the LAST matching field wins; a field with count 0 means "no stamps"; float32 stamps are widened
    dt3 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("time", "<f4"), ("stamps", "<f4")])
    rec3 = np.zeros(n, dtype=dt3)
    rec3["time"], rec3["stamps"] = 1.0, np.linspace(0, 0.1, n, dtype=np.float32)
    f3 = [("x", 0, FLOAT32, 1), ("y", 4, FLOAT32, 1), ("z", 8, FLOAT32, 1), ("time", 12, FLOAT32, 1), ("stamps", 16, FLOAT32, 1)]
    _, stamps = _decode(L, rec3, f3)
    assert np.array_equal(stamps, rec3["stamps"].astype(np.float64))
    _, stamps = _decode(L, rec3, f3[:4] + [("stamps", 16, FLOAT32, 0)])
    assert len(stamps) == 0
    _, stamps = _decode(L, rec3, f3[:3])
    assert len(stamps) == 0  # no time field: de-skewing disabled
    # errors: unsupported stamp type, x/y/z not float32, missing axis
    assert _decode(L, rec3, f3[:3] + [("t", 12, 5, 1)]) == (None, None)
    assert _decode(L, rec3, [("x", 0, FLOAT64, 1)] + f3[1:]) == (None, None)
    assert _decode(L, rec3, f3[1:]) == (None, None)


def test_sweep_timing_matches_timestamp_handler():
    """TimeStampHandler::ProcessTimestamps (utils/TimeStampHandler.cpp:108-139) in double seconds."""
    L = _facade()
    c_dp = C.POINTER(C.c_double)
    L.kfac_process_timestamps.argtypes = [c_dp, C.c_int64, C.c_double, c_dp, c_dp]

    def process(stamps, header, last):
        s = np.array(stamps, dtype=np.float64)
        lp, be = np.array([last]), np.zeros(2)
        L.kfac_process_timestamps(s.ctypes.data_as(c_dp), len(s), header, lp.ctypes.data_as(c_dp), be.ctypes.data_as(c_dp))
        return s, float(lp[0]), be
    raw = 100.0 + np.array([0.02, 0.0, 0.1, 0.05])
    # header stamps the END of the sweep (== max stamp): the interval ends at the header
    s, last, be = process(raw, 100.1, 99.9)
    assert np.allclose(s, [0.2, 0.0, 1.0, 0.5], atol=1e-12) and s.min() == 0.0 and s.max() == 1.0
    assert be.tolist() == [99.9, 100.1] and last == 100.1
    # header stamps the BEGINNING: the sweep duration is added
    s2, last, be = process(raw, 100.0, 100.1 - 0.1)
    assert np.array_equal(s2, s) and abs(be[1] - 100.1) < 1e-9 and last == be[1] and be[0] == 100.0
    # relative stamps (e.g. Ouster's t in seconds from the sweep start) with an end-stamped header are treated as begin-stamped
    s3, last, be = process([0.0, 0.05, 0.1], 50.0, 49.9)
    assert np.allclose(s3, [0.0, 0.5, 1.0]) and abs(be[1] - 50.1) < 1e-12
    # no stamps: the interval is [last processed, header]
    s4, last, be = process([], 7.5, 7.4)
    assert len(s4) == 0 and be.tolist() == [7.4, 7.5] and last == 7.5


def test_replay_harness_binary_is_built_and_has_no_cpu_fallback(tmp_path):
    """kicp_replay (the product's offline_node counterpart) is part of the build; without a CUDA device it refuses to run."""
    import subprocess
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(ROOT, "kinematic-icp_b200", "bin", "kicp_replay")
    assert os.path.exists(exe)
    usage = subprocess.run([exe], capture_output=True, text=True)
    assert usage.returncode == 2 and "usage:" in usage.stderr
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        (tmp_path / "empty.kseq").write_bytes(b"KSEQ1\0\0\0" + b"\0" * 120)
        run = subprocess.run([exe, str(tmp_path / "empty.kseq"), str(tmp_path / "o.tum")], capture_output=True, text=True)
        assert run.returncode == 1 and "CUDA" in run.stderr and not (tmp_path / "o.tum").exists()
