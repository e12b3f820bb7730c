"""The NumPy oracle's integer / bit-exact pieces cross-checked against the plain-C restatement (oracle/g4r_oracle.c)."""
import ctypes as C
import os
import subprocess
import numpy as np
import gru4rec_oracle as orc

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle')


def _lib():
    so = os.path.join(ORACLE_DIR, 'libg4r_oracle.so')
    if not os.path.exists(so):
        subprocess.check_call(['make', '-C', ORACLE_DIR, '-s'])
    return C.CDLL(so)


def test_k2_c_vs_numpy():
    lib = _lib()
    rs = np.random.RandomState(0)
    for n_d, alpha in ((1, 0.5), (7, 1.0), (1000, 0.75), (37483, 0.0)):
        P = orc.sampling_cdf(rs.randint(1, 500, size=n_d), alpha).astype(np.float32)
        x = rs.rand(50000).astype(np.float32)
        x[:min(n_d, 200)] = P[:min(n_d, 200)]
        x[200:204] = [0.0, 1.0, 1.5, np.nextafter(np.float32(1), np.float32(0))]
        y = np.empty(len(x), dtype=np.int64)
        lib.k2_searchsorted(P.ctypes.data_as(C.c_void_p), C.c_longlong(n_d), x.ctypes.data_as(C.c_void_p), C.c_longlong(len(x)), y.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(y, orc.searchsorted_k2(P, x))


def test_mrg_c_vs_numpy():
    lib = _lib()
    ref = orc.MRGStreams(12345)
    n = 20011
    st = ref.substreams(ref.n_streams(n))
    st_c = np.ascontiguousarray(st.astype(np.int32))
    out_c = np.empty(n, dtype=np.float32)
    for call in range(2):
        u = ref.uniform_from_state(st, n)
        lib.mrg31k3p_fill(st_c.ctypes.data_as(C.c_void_p), C.c_longlong(st_c.shape[0]), out_c.ctypes.data_as(C.c_void_p), C.c_longlong(n))
        np.testing.assert_array_equal(out_c, u)
        np.testing.assert_array_equal(st_c.astype(np.int64), st)
