"""TEST INFRASTRUCTURE — ctypes view of oracle/db_reader.c, the plain-C restatement of the reference's elodin-db READER
(libs/db/src/append_log.rs:84-143, time_series.rs:40-139) plus the directory walk of DB::open (libs/db/src/lib.rs:592-672).
Only tests/ may import this; the product (elodin_b200/db_sink.py) has its own writer and reader."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdb_reader.so")
_lib = None


class _Log(C.Structure):
    _fields_ = [("committed_len", C.c_uint64), ("head_len", C.c_uint64), ("extra", C.c_uint8 * 8), ("data", C.c_void_p), ("len", C.c_uint64)]


class _Series(C.Structure):
    _fields_ = [("index", _Log), ("data", _Log), ("rows", C.c_uint64), ("element_size", C.c_uint64), ("index_extra", C.c_int64)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "db_reader.c")):
            subprocess.run(["make", "-s", "-C", _HERE, "libdb_reader.so"], check=True)
        L = C.CDLL(_SO)
        sp = C.POINTER(_Series)
        L.orc_series_open.argtypes = [C.c_char_p, sp]
        L.orc_series_close.argtypes = [sp]
        L.orc_series_close.restype = None
        for name, res in (("orc_series_rows", C.c_uint64), ("orc_series_index_rows", C.c_uint64), ("orc_series_element_size", C.c_uint64),
                          ("orc_series_index_extra", C.c_int64), ("orc_series_start_timestamp", C.c_int64),
                          ("orc_series_timestamps", C.POINTER(C.c_int64)), ("orc_series_data", C.POINTER(C.c_uint8))):
            getattr(L, name).argtypes = [sp]
            getattr(L, name).restype = res
        L.orc_series_get.argtypes = [sp, C.c_int64]
        L.orc_series_get.restype = C.c_int64
        L.orc_series_get_nearest.argtypes = [sp, C.c_int64]
        L.orc_series_get_nearest.restype = C.c_int64
        L.orc_series_range.argtypes = [sp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


class Series:
    """TimeSeries::open + its accessors (time_series.rs:40-139)."""

    def __init__(self, path: str):
        self._L = lib()
        self._s = _Series()
        rc = self._L.orc_series_open(path.encode(), C.byref(self._s))
        if rc:
            raise OSError(f"cannot open time series {path!r} (code {rc})")
        p = C.byref(self._s)
        self.rows = int(self._L.orc_series_rows(p))
        self.index_rows = int(self._L.orc_series_index_rows(p))
        self.element_size = int(self._L.orc_series_element_size(p))
        self.index_extra = int(self._L.orc_series_index_extra(p))
        self.start_timestamp = int(self._L.orc_series_start_timestamp(p))
        self.timestamps = np.ctypeslib.as_array(self._L.orc_series_timestamps(p), (self.index_rows,)).copy() if self.index_rows else np.zeros(0, np.int64)
        n = self.rows * self.element_size
        self.data = bytes(np.ctypeslib.as_array(self._L.orc_series_data(p), (n,))) if n else b""

    def row(self, i: int) -> bytes:
        return self.data[i * self.element_size:(i + 1) * self.element_size]

    def get(self, timestamp: int):
        i = int(self._L.orc_series_get(C.byref(self._s), int(timestamp)))
        return None if i < 0 else self.row(i)

    def get_nearest(self, timestamp: int):
        i = int(self._L.orc_series_get_nearest(C.byref(self._s), int(timestamp)))
        return None if i < 0 else (int(self.timestamps[i]), self.row(i))

    def get_range(self, start: int, end: int):
        i0, i1 = C.c_uint64(), C.c_uint64()
        if not self._L.orc_series_range(C.byref(self._s), int(start), int(end), C.byref(i0), C.byref(i1)):
            return None
        return self.timestamps[i0.value:i1.value], self.data[i0.value * self.element_size:i1.value * self.element_size]

    def latest(self):
        return None if not self.rows else (int(self.timestamps[self.rows - 1]), self.row(self.rows - 1))

    def close(self):
        if self._s is not None:
            self._L.orc_series_close(C.byref(self._s))
            self._s = None


def open_db(path: str):
    """The directory walk of DB::open (lib.rs:592-672): db_state must exist; every sub-directory except msgs / assets /
    simulation_source must be named by a decimal ComponentId; a directory without `schema` is skipped (entity-level
    metadata); the rest are time series.  Returns ({component_id: Series}, {component_id: has_metadata}, last_updated,
    start_timestamp) with the two time bounds computed as the reference does (timestamp 0 and i64::MAX excluded)."""
    if not os.path.exists(os.path.join(path, "db_state")):
        raise FileNotFoundError("MissingDbState")
    series, has_meta = {}, {}
    last_updated, start = -(1 << 63), (1 << 63) - 1
    for name in os.listdir(path):
        p = os.path.join(path, name)
        if not os.path.isdir(p) or name in ("msgs", "assets", "simulation_source"):
            continue
        if not name.isdigit():
            raise ValueError("InvalidComponentId: " + name)
        cid = int(name)
        has_meta[cid] = os.path.exists(os.path.join(p, "metadata"))
        if not os.path.exists(os.path.join(p, "schema")):
            continue
        s = Series(p)
        series[cid] = s
        if s.latest() is not None:
            last_updated = max(last_updated, s.latest()[0])
        if s.start_timestamp not in (0, (1 << 63) - 1):
            start = min(start, s.start_timestamp)
    return series, has_meta, last_updated, start
