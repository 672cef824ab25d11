"""NetCDF files of COO arrays for the Python mirror: the layout of the reference's ncio_spsparse (slib/spsparse/netcdf.hpp:86-138)

    dimensions  <v>.size, <v>.rank
    variables   <v>.info    int64 scalar, attribute "shape" uint64[rank]
                <v>.indices int64 [<v>.size, <v>.rank]
                <v>.vals    double [<v>.size]

read and written with numpy only (the image has no netCDF library): the classic format family, written as CDF-5 (the 64-bit
types need it), read as CDF-1 / CDF-2 / CDF-5 -- the same files include/spsparse/netcdf.hpp writes and reads through
include/spsparse_b200/mini_netcdf (byte-identical for the same array: tests/test_netcdf_cpu.py).  Host I/O only.

    write_spsparse(path, {"A": (shape, [rows, cols], vals), ...})
    shape, idx, val = read_spsparse(path, "A")          # idx: list of int64 arrays, one per dimension
"""
from __future__ import annotations

import struct

import numpy as np

NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
_TYPES = {1: ">i1", 2: "S1", 3: ">i2", 4: ">i4", 5: ">f4", 6: ">f8", 7: ">u1", 8: ">u2", 9: ">u4", 10: ">i8", 11: ">u8"}
NC_DOUBLE, NC_INT64, NC_UINT64 = 6, 10, 11


def _pad4(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 4)


# ---- writer (CDF-5) ---------------------------------------------------------------------------------------------------------------
def _name(s: str) -> bytes:
    return struct.pack(">q", len(s)) + _pad4(s.encode())


def write_spsparse(path: str, arrays: dict) -> None:
    """arrays: {name: (shape, [index array per dimension], values)} -- written in the order given, as ncio_spsparse would
    write them one after the other into one file."""
    dims, variables, blobs = [], [], []
    for vname, (shape, idx, val) in arrays.items():
        rank = len(shape)
        val = np.ascontiguousarray(val, dtype=np.float64)
        n = len(val)
        if len(idx) != rank or any(len(x) != n for x in idx):
            raise ValueError(f"{vname}: {rank} index arrays of {n} entries expected")
        if n == 0 or rank == 0:
            raise ValueError(f"{vname}: the classic format has no zero-length fixed dimension (an empty array cannot be stored)")
        d0 = len(dims)
        dims += [(vname + ".size", n), (vname + ".rank", rank)]
        shape_att = (struct.pack(">iq", NC_ATTRIBUTE, 1) + _name("shape") + struct.pack(">iq", NC_UINT64, rank) +
                     struct.pack(f">{rank}Q", *[int(s) for s in shape]))
        absent = struct.pack(">iq", 0, 0)
        variables += [(vname + ".info", [], shape_att, NC_INT64, 8),
                      (vname + ".indices", [d0, d0 + 1], absent, NC_INT64, 8 * n * rank),
                      (vname + ".vals", [d0], absent, NC_DOUBLE, 8 * n)]
        blobs += [struct.pack(">q", 0),
                  np.stack([np.asarray(x, dtype=np.int64) for x in idx], axis=1).astype(">i8").tobytes(),
                  val.astype(">f8").tobytes()]
    hdr = b"CDF\x05" + struct.pack(">q", 0)
    hdr += struct.pack(">iq", NC_DIMENSION, len(dims)) + b"".join(_name(nm) + struct.pack(">q", ln) for nm, ln in dims)
    hdr += struct.pack(">iq", 0, 0)   # no global attributes

    def var_header(nm, dimids, atts, typ, vsize, begin):
        return (_name(nm) + struct.pack(">q", len(dimids)) + b"".join(struct.pack(">q", d) for d in dimids) + atts +
                struct.pack(">iqq", typ, vsize, begin))
    fixed = len(hdr) + 12 + sum(len(var_header(*v, 0)) for v in variables)
    begins, off = [], fixed
    for v in variables:
        begins.append(off)
        off += v[4]
    hdr += struct.pack(">iq", NC_VARIABLE, len(variables)) + b"".join(var_header(*v, b) for v, b in zip(variables, begins))
    with open(path, "wb") as f:
        f.write(hdr)
        for b in blobs:
            f.write(b)


# ---- reader (CDF-1 / CDF-2 / CDF-5, fixed-size variables) ------------------------------------------------------------------------------
class _Reader:
    def __init__(self, raw: bytes):
        if raw[:4] == b"\x89HDF":
            raise ValueError("a netCDF-4/HDF5 file; this reader takes the classic formats (convert with `nccopy -k cdf5`)")
        if raw[:3] != b"CDF" or raw[3] not in (1, 2, 5):
            raise ValueError("not a netCDF classic (CDF-1/2/5) file")
        self.raw, self.at, self.v = raw, 4, raw[3]

    def u32(self):
        (x,) = struct.unpack_from(">I", self.raw, self.at)
        self.at += 4
        return x

    def u64(self):
        (x,) = struct.unpack_from(">Q", self.raw, self.at)
        self.at += 8
        return x

    def non_neg(self):
        return self.u64() if self.v == 5 else self.u32()

    def name(self):
        k = self.non_neg()
        s = self.raw[self.at:self.at + k].decode()
        self.at += (k + 3) & ~3
        return s

    def att_list(self):
        tag, cnt = self.u32(), self.non_neg()
        out = {}
        if tag == 0 and cnt == 0:
            return out
        if tag != NC_ATTRIBUTE:
            raise ValueError("attribute list expected")
        for _ in range(cnt):
            nm, typ, ne = self.name(), self.u32(), self.non_neg()
            dt = np.dtype(_TYPES[typ])
            out[nm] = np.frombuffer(self.raw, dtype=dt, count=ne, offset=self.at)
            self.at += (ne * dt.itemsize + 3) & ~3
        return out


def read_header(raw: bytes):
    """-> (dims: [(name, length)], global attributes, variables: {name: (dimids, attributes, dtype, begin)})"""
    r = _Reader(raw)
    r.non_neg()   # numrecs
    tag, cnt = r.u32(), r.non_neg()
    dims = []
    if not (tag == 0 and cnt == 0):
        if tag != NC_DIMENSION:
            raise ValueError("dimension list expected")
        dims = [(r.name(), r.non_neg()) for _ in range(cnt)]
    gatts = r.att_list()
    tag, cnt = r.u32(), r.non_neg()
    variables = {}
    if not (tag == 0 and cnt == 0):
        if tag != NC_VARIABLE:
            raise ValueError("variable list expected")
        for _ in range(cnt):
            nm = r.name()
            dimids = [r.non_neg() for _ in range(r.non_neg())]
            atts = r.att_list()
            typ = r.u32()
            r.non_neg()   # vsize
            begin = r.u32() if r.v == 1 else r.u64()
            variables[nm] = (dimids, atts, np.dtype(_TYPES[typ]), begin)
    return dims, gatts, variables


def read_spsparse(path: str, vname: str):
    """-> (shape tuple, [int64 index array per dimension], float64 values) of the array stored under `vname`."""
    raw = np.fromfile(path, dtype=np.uint8).tobytes()
    dims, _, variables = read_header(raw)
    for need in (".info", ".indices", ".vals"):
        if vname + need not in variables:
            raise KeyError(f"{path}: no variable {vname + need}")
    shape = tuple(int(x) for x in variables[vname + ".info"][1]["shape"])
    dimids, _, dt, begin = variables[vname + ".indices"]
    if len(dimids) != 2 or any(dims[d][1] == 0 for d in dimids):
        raise ValueError(f"{vname}.indices: a fixed [size, rank] variable expected")
    n, rank = dims[dimids[0]][1], dims[dimids[1]][1]
    if rank != len(shape):
        raise ValueError(f"{vname}: rank {rank} in the file, shape attribute of length {len(shape)}")
    ind = np.frombuffer(raw, dtype=dt, count=n * rank, offset=begin).reshape(n, rank).astype(np.int64)
    _, _, dtv, beginv = variables[vname + ".vals"]
    val = np.frombuffer(raw, dtype=dtv, count=n, offset=beginv).astype(np.float64)
    return shape, [np.ascontiguousarray(ind[:, k]) for k in range(rank)], val
