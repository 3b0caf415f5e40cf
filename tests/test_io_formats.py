"""SAC / sac2bin `.bin` readers and writers of the C library (no GPU needed),
checked against an independent struct-level parser.  The SAC cases run where the reference's example
data exists (/root/reference, i.e. the build container); the `.bin` and writer round trips run anywhere."""
import ctypes as C
import os
import struct
import time

import numpy as np
import pytest

from oracle.gen_golden import REAL_DIR, read_sac_py

HAVE_REAL = os.path.isdir(REAL_DIR)


def _data_to_numpy(pkg, d):
    mtr, n = d.hdr.mtr, d.hdr.max
    x = np.ctypeslib.as_array(d.sigall, (mtr, n)).copy()
    t = np.ctypeslib.as_array(d.time, (mtr,)).copy()
    lag0 = np.ctypeslib.as_array(d.lag0, (mtr,)).copy()
    return x, t, lag0


@pytest.mark.skipif(not HAVE_REAL, reason="reference example data not present")
def test_sac_list_loader_matches_struct_parser(pkg, lib, tmp_path):
    files = sorted(os.listdir(REAL_DIR))[:25]
    lst = tmp_path / "filelist.txt"
    lst.write_text("\n".join(os.path.join(REAL_DIR, f) for f in files) + "\n")
    d = pkg.Data()
    assert lib.tspws_b200_load_sac_list(str(lst).encode(), C.byref(d), 0) == 0
    try:
        x, t, lag0 = _data_to_numpy(pkg, d)
        assert (d.hdr.mtr, d.hdr.max) == (25, 16501)
        assert d.hdr.dt == 4.0 and d.hdr.beg == -33000.0
        for i, f in enumerate(files):
            ref = read_sac_py(os.path.join(REAL_DIR, f))
            np.testing.assert_array_equal(x[i], ref["data"])
        # file names carry year.jday: the start times must land on that UTC day
        for i, f in enumerate(files):
            year, jday = int(f.split("_")[-1][:4]), int(f.split("_")[-1][5:8])
            tm = time.gmtime(int(t[i]))
            assert (tm.tm_year, tm.tm_yday) == (year, jday)
        assert d.hdr.sta2 == b"ccgn"          # kstnm of the example files
    finally:
        lib.tspws_b200_free_data(C.byref(d))


def test_bin_round_trip_and_layout(pkg, lib, tmp_path):
    """save_bin writes t_ccheader | time_t[nseq] | float lag0[nseq] | float data (reference src/sac2bin.c:192-195)."""
    mtr, n = 7, 301
    rng = np.random.default_rng(1)
    x = rng.standard_normal((mtr, n)).astype(np.float32)
    times = (1262304000 + 86400 * np.arange(mtr)).astype(np.int64)
    lag0 = rng.standard_normal(mtr).astype(np.float32)
    d = pkg.Data()
    d.sigall = x.ctypes.data_as(C.POINTER(C.c_float))
    d.time = times.ctypes.data_as(C.POINTER(C.c_long))
    d.lag0 = lag0.ctypes.data_as(C.POINTER(C.c_float))
    d.hdr.max, d.hdr.mtr, d.hdr.dt, d.hdr.beg = n, mtr, 0.5, -75.0
    d.hdr.sta1, d.hdr.sta2, d.hdr.net1, d.hdr.chn2 = b"ECH", b"CAN", b"G", b"00Z"
    d.hdr.stla, d.hdr.evlo = 12.5, -3.25
    path = str(tmp_path / "pair.bin").encode()
    assert lib.tspws_b200_save_bin(path, C.byref(d), b"ccgn") == 0
    raw = open(path, "rb").read()
    assert len(raw) == 116 + 8 * mtr + 4 * mtr + 4 * mtr * n
    assert raw[:4] == b"ccgn" and raw[16:19] == b"ECH" and raw[48:51] == b"CAN"
    nlags, nseq, tlength, lag1, lag2 = struct.unpack("<IIfff", raw[96:116])
    assert (nlags, nseq, lag1) == (n, mtr, -75.0) and abs(lag2 - (-75.0 + 0.5 * (n - 1))) < 1e-4
    np.testing.assert_array_equal(np.frombuffer(raw[116:116 + 8 * mtr], "<i8"), times)
    np.testing.assert_array_equal(np.frombuffer(raw[116 + 12 * mtr:], "<f4").reshape(mtr, n), x)
    e = pkg.Data()
    assert lib.tspws_b200_load_bin(path, C.byref(e), 0) == 0
    try:
        y, t2, l2 = _data_to_numpy(pkg, e)
        np.testing.assert_array_equal(y, x)
        np.testing.assert_array_equal(t2, times)
        np.testing.assert_array_equal(l2, lag0)
        assert (e.hdr.max, e.hdr.mtr) == (n, mtr) and abs(e.hdr.dt - 0.5) < 1e-6 and e.hdr.beg == -75.0
        assert e.hdr.sta1 == b"ECH" and e.hdr.sta2 == b"CAN" and abs(e.hdr.stla - 12.5) < 1e-6
    finally:
        lib.tspws_b200_free_data(C.byref(e))
    assert lib.tspws_b200_load_bin(b"/nonexistent/file.bin", C.byref(pkg.Data()), 0) == -2


def test_sac_writer_reader_round_trip(pkg, lib, tmp_path):
    n = 501
    y = np.sin(np.arange(n) * 0.1).astype(np.float32)
    h = pkg.Hdr()
    h.max, h.mtr, h.dt, h.beg = n, 1, 4.0, -1000.0
    h.sta2, h.net2, h.chn2, h.sta1 = b"CAN", b"G", b"00Z", b"ECH"
    path = str(tmp_path / "tl_test.sac").encode()
    assert lib.tspws_b200_write_sac(path, y.ctypes.data_as(C.POINTER(C.c_float)), C.byref(h), n, -1000.0, b"t-lin", 499.0) == 0
    back = read_sac_py(path.decode())
    assert back["npts"] == n and back["delta"] == 4.0 and back["b"] == -1000.0
    np.testing.assert_array_equal(back["data"], y)
    raw = open(path, "rb").read()
    assert len(raw) == 632 + 4 * n
    assert raw[624:629] == b"t-lin" and raw[440:443] == b"CAN" and raw[608:609] == b"G"
    assert struct.unpack("<f", raw[4 * 40:4 * 41])[0] == 499.0 and struct.unpack("<i", raw[4 * 76:4 * 77])[0] == 6
    buf = np.zeros(n, np.float32)
    dt, beg = C.c_float(), C.c_float()
    assert lib.tspws_b200_read_sac(path, buf.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(dt), C.byref(beg), None) == n
    np.testing.assert_array_equal(buf, y)
    assert lib.tspws_b200_read_sac(b"/nonexistent.sac", None, 0, None, None, None) == -1


def test_rows_bin_writer_layout(pkg, lib, tmp_path):
    """Resampled stacks as the reference's wrbin writes them (src/ts_pws1f.c:766-823): 116-byte header,
    time_t[M] holding the traces behind every row, float data[M][npts] -- and no lag0 block."""
    import ctypes as C
    x = np.random.default_rng(5).standard_normal((4, 301)).astype(np.float32)
    hdr = pkg.Hdr()
    hdr.max, hdr.mtr, hdr.dt, hdr.beg = 301, 77, 0.5, -75.0
    rows = (C.POINTER(C.c_float) * 4)(*[x[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(4)])
    k = np.array([40, 41, 42, 43], dtype=np.uint32)
    lib.tspws_b200_save_rows_bin.restype = C.c_int
    lib.tspws_b200_save_rows_bin.argtypes = [C.c_char_p, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.POINTER(pkg.Hdr),
                                             C.c_float, C.c_char_p, C.c_void_p]
    path = str(tmp_path / "rows.bin").encode()
    n0, npts = 150, 151                                   # a folded stack keeps its second half
    assert lib.tspws_b200_save_rows_bin(path, rows, 4, n0, npts, C.byref(hdr), C.c_float(0.0), b"ts_pws", k.ctypes.data) == 0
    raw = open(path, "rb").read()
    assert len(raw) == 116 + 4 * 8 + 4 * npts * 4
    assert raw[:6] == b"ts_pws"
    nlags, nseq = np.frombuffer(raw[96:104], dtype=np.uint32)
    tlength, lag1, lag2 = np.frombuffer(raw[104:116], dtype=np.float32)
    assert (nlags, nseq) == (npts, 4) and lag1 == 0.0 and tlength == lag2 == 0.5 * (npts - 1)
    np.testing.assert_array_equal(np.frombuffer(raw[116:116 + 32], dtype=np.int64), k)
    np.testing.assert_array_equal(np.frombuffer(raw[148:], dtype=np.float32).reshape(4, npts), x[:, n0:])
    assert lib.tspws_b200_save_rows_bin(None, rows, 4, n0, npts, C.byref(hdr), C.c_float(0.0), b"x", None) != 0
