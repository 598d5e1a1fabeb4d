"""TUM trajectory format of OfflineNode::writePosesInTumFormat (offline_node.cpp:76-97): 'timestamp x y z qx qy qz qw',
fixed notation, 6 decimals."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tum_writer_format(tmp_path):
    import __graft_entry__ as g
    g.build()
    L = C.CDLL(os.path.join(ROOT, "kinematic-icp_b200", "lib", "libkinematic_icp_b200.so"))
    c_dp = C.POINTER(C.c_double)
    L.kfac_write_tum.argtypes = [C.c_char_p, c_dp, c_dp, C.c_int64]
    stamps = np.array([1700000000.123456789, 1700000000.2])
    poses = np.array([[0.0, 0.0, 0.3826834324, 0.9238795325, 1.5, -2.25, 0.0], [0, 0, 0, 1, 10.123456789, 0, 0.5]])
    path = tmp_path / "poses.tum"
    assert L.kfac_write_tum(str(path).encode(), stamps.ctypes.data_as(c_dp), poses.ctypes.data_as(c_dp), 2) == 0
    lines = path.read_text().splitlines()
    assert lines[0] == "1700000000.123457 1.500000 -2.250000 0.000000 0.000000 0.000000 0.382683 0.923880"
    assert lines[1] == "1700000000.200000 10.123457 0.000000 0.500000 0.000000 0.000000 0.000000 1.000000"
