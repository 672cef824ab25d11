"""CPU suite: NetCDF I/O of COO arrays (reference slib/spsparse/netcdf.hpp:23-138, tests/test_netcdf.cpp).

include/spsparse/netcdf.hpp runs here over the minimal classic-format netCDF of include/spsparse_b200/mini_netcdf/ (the image has
no netCDF library).  Three independent checks of that format code:
(1) tests/cpp/netcdf_test.cpp selftest: the reference's write / read(alloc) / read(no alloc) round trip and the API subset;
(2) the CDF-5 file ncio_spsparse writes, byte for byte against a file assembled HERE from the published grammar
    (header = magic numrecs dim_list gatt_list var_list, big-endian, 4-byte padding; CDF-5: 64-bit counts and offsets);
(3) scipy.io.netcdf_file -- an unrelated implementation of CDF-1/2 -- reads what the C++ writer wrote (classic, classic64)
    and the C++ reader reads what scipy wrote, record (unlimited) variables included.
(4) the REFERENCE's own tests/test_netcdf.cpp, compiled unmodified against these headers (oracle/_ref/dropin_test_netcdf).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    return g.build_cpp_test("netcdf_test")


def run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def test_selftest(exe, tmp_path):
    out = run(exe, "selftest", str(tmp_path))
    assert "0 failure(s)" in out, out


# ---- (2) the CDF-5 grammar, restated --------------------------------------------------------------------------------------------
NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
NC_DOUBLE, NC_INT64, NC_UINT64 = 6, 10, 11


def pad4(b):
    return b + b"\0" * (-len(b) % 4)


def name5(s):
    return struct.pack(">q", len(s)) + pad4(s.encode())


def cdf5_spsparse(vname, shape, idx, val):
    """The file ncio_spsparse writes for one array (netcdf.hpp:93-106), as CDF-5 bytes."""
    n, rank = len(val), len(shape)
    dims = [(vname + ".size", n), (vname + ".rank", rank)]
    hdr = b"CDF\x05" + struct.pack(">q", 0)
    hdr += struct.pack(">iq", NC_DIMENSION, len(dims)) + b"".join(name5(nm) + struct.pack(">q", ln) for nm, ln in dims)
    hdr += struct.pack(">iq", 0, 0)                                                  # no global attributes
    shape_att = struct.pack(">iq", NC_ATTRIBUTE, 1) + name5("shape") + struct.pack(">iq", NC_UINT64, rank) + struct.pack(f">{rank}Q", *shape)
    absent = struct.pack(">iq", 0, 0)
    variables = [  # name, dimids, attribute list, type, vsize
        (vname + ".info", [], shape_att, NC_INT64, 8),
        (vname + ".indices", [0, 1], absent, NC_INT64, 8 * n * rank),
        (vname + ".vals", [0], absent, NC_DOUBLE, 8 * n),
    ]
    def var_header(nm, dimids, atts, typ, vsize, begin):
        return name5(nm) + struct.pack(">q", len(dimids)) + b"".join(struct.pack(">q", d) for d in dimids) + atts + struct.pack(">iqq", typ, vsize, begin)
    fixed = len(hdr) + 12 + sum(len(var_header(*v, 0)) for v in variables)
    begins, off = [], fixed
    for v in variables:
        begins.append(off)
        off += v[4]
    hdr += struct.pack(">iq", NC_VARIABLE, len(variables)) + b"".join(var_header(*v, b) for v, b in zip(variables, begins))
    assert len(hdr) == fixed
    data = struct.pack(">q", 0)                                                      # <v>.info: never written, reads as 0
    data += np.asarray(idx, dtype=">i8").reshape(n, rank).tobytes() + np.asarray(val, dtype=">f8").tobytes()
    return hdr + data


FIXED_IDX = [[1, 2], [3, 3], [4, 5], [1, 2], [0, 0]]
FIXED_VAL = [2.0, 6.0, 1.0, -0.25, 0.0]


def test_cdf5_file_byte_for_byte(exe, tmp_path):
    p = tmp_path / "arr1.nc"
    run(exe, "write", str(p))
    want = cdf5_spsparse("arr1", (5, 6), FIXED_IDX, FIXED_VAL)
    got = p.read_bytes()
    assert got == want, (len(got), len(want), next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), None))


def test_reader_takes_a_file_assembled_from_the_grammar(exe, tmp_path):
    """... and the reverse: bytes assembled here (another array, rank 1) are read back by the C++ side."""
    p = tmp_path / "v.nc"
    p.write_bytes(cdf5_spsparse("v", (9,), [[8], [0], [3]], [1.5, -2.0, 1e300]))
    out = run(exe, "dump", str(p))
    assert "var v.indices int64 v.size v.rank\nvalues 8 0 3\n" in out and "var v.vals double v.size\nvalues 1.5 -2 1.0000000000000001e+300\n" in out, out


# ---- (3) against scipy's netCDF-3 implementation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,version", [("classic", 1), ("classic64", 2)])
def test_scipy_reads_what_the_writer_wrote(exe, tmp_path, fmt, version):
    from scipy.io import netcdf_file
    p = tmp_path / f"{fmt}.nc"
    run(exe, "write-classic", str(p), fmt)
    assert p.read_bytes()[:4] == b"CDF" + bytes([version])
    with netcdf_file(str(p), "r", mmap=False) as f:
        assert f.version_byte == version
        assert dict(f.dimensions) == {"n": 5, "rank": 2}
        assert f.variables["indices"].data.tolist() == FIXED_IDX and f.variables["indices"].data.dtype == np.dtype(">i4")
        assert f.variables["vals"].data.tolist() == FIXED_VAL
        assert f.variables["tag"].data.tolist() == [-7, 300]
        assert f.variables["indices"]._attributes["shape"].tolist() == [5, 6]
        assert f.history == b"written by netcdf_test"


@pytest.mark.parametrize("version", [1, 2])
def test_reader_takes_what_scipy_wrote(exe, tmp_path, version):
    from scipy.io import netcdf_file
    p = tmp_path / f"scipy{version}.nc"
    rng = np.random.default_rng(5 + version)
    a = rng.integers(-1000, 1000, (4, 3)).astype(np.int32)
    b = rng.standard_normal(7)
    c = rng.integers(-100, 100, (5, 3)).astype(np.int16)    # record variable: 3 shorts per record, padded
    d = rng.standard_normal(5).astype(np.float32)          # second record variable
    with netcdf_file(str(p), "w", version=version) as f:
        f.createDimension("t", None)
        f.createDimension("y", 4)
        f.createDimension("x", 3)
        f.createDimension("m", 7)
        f.title = "from scipy"
        va = f.createVariable("a", "i4", ("y", "x")); va[:] = a; va.units = "counts"
        vb = f.createVariable("b", "f8", ("m",)); vb[:] = b
        vc = f.createVariable("c", "i2", ("t", "x")); vc[:] = c
        vd = f.createVariable("d", "f4", ("t",)); vd[:] = d
    out = run(exe, "dump", str(p))
    vals = {}
    lines = out.splitlines()
    for i, ln in enumerate(lines):
        if ln.startswith("var "):
            vals[ln.split()[1]] = (ln.split()[2:], np.array([float(x) for x in lines[i + 1].split()[1:]]))
    assert "dim t 5 1" in out and "dim y 4 0" in out
    assert vals["a"][0] == ["int", "y", "x"] and np.array_equal(vals["a"][1], a.ravel().astype(float))
    assert vals["b"][0] == ["double", "m"] and np.array_equal(vals["b"][1], b)
    assert vals["c"][0] == ["short", "t", "x"] and np.array_equal(vals["c"][1], c.ravel().astype(float))
    assert vals["d"][0] == ["float", "t"] and np.array_equal(vals["d"][1], d.astype(float))


def test_single_record_variable_is_not_padded(exe, tmp_path):
    """The classic format's special case: with exactly one record variable, records are packed without padding."""
    from scipy.io import netcdf_file
    p = tmp_path / "onerec.nc"
    c = np.arange(15, dtype=np.int16).reshape(5, 3) - 7
    with netcdf_file(str(p), "w", version=1) as f:
        f.createDimension("t", None)
        f.createDimension("x", 3)
        vc = f.createVariable("c", "i2", ("t", "x")); vc[:] = c
    out = run(exe, "dump", str(p))
    assert "values " + " ".join(str(int(x)) for x in c.ravel()) in out, out


# ---- (4) the reference's own test source -----------------------------------------------------------------------------------------
def test_reference_netcdf_test_against_our_headers(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_test_netcdf")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_test_netcdf not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, cwd=str(tmp_path))
    assert r.returncode == 0 and " 0 failed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
