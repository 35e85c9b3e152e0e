"""Minimal zarr-v2 directory store for the embedding container (micro_sam/util.py:684-747, :902-1041, :1044-1096).

The reference keeps image embeddings in a zarr group: `features` is a float32 dataset (2-D / 3-D inputs) or a sub-group with
one dataset per tile (tiled inputs), the embedding signature lives in the group attributes.  `zarr` is not installed in this
image, so this module writes that layout itself -- `.zgroup` / `.zarray` / `.zattrs` JSON plus one raw little-endian C-order
file per chunk (`compressor: null`, key "i.j.k") -- which any zarr v2 reader (zarr-python `zarr.open(path)`) opens as is.
Only what the embedding path needs is implemented: whole-chunk reads and writes (`ds[z]`, `ds[z] = x`, `ds[:]`), attributes,
`require_group`, dataset creation with and without data.  Missing chunks read as the fill value 0, which is what the
reference's resume logic (`np.count_nonzero(features[z]) != 0`, util.py:990) relies on.  Metadata and chunk files are written
atomically (temp file + rename), so several ranks may fill one container concurrently.
"""
from __future__ import annotations

import json
import os
import tempfile
import threading
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np


def _atomic_write(path: str, data: bytes) -> None:
    d = os.path.dirname(path)
    fd, tmp = tempfile.mkstemp(dir=d, prefix=".tmp-")
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise


def _jsonable(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    if isinstance(v, dict):
        return {str(k): _jsonable(x) for k, x in v.items()}
    return v


class Attrs:
    """`.zattrs` of a group or array; JSON semantics like zarr's (tuples come back as lists)."""

    def __init__(self, path: Optional[str]):
        self._path = path
        self._lock = threading.Lock()
        self._mem: Dict[str, Any] = {}

    def _load(self) -> Dict[str, Any]:
        if self._path is None:
            return self._mem
        if not os.path.exists(self._path):
            return {}
        with open(self._path, "r") as f:
            return json.load(f)

    def _store(self, d: Dict[str, Any]) -> None:
        if self._path is None:
            self._mem = d
        else:
            _atomic_write(self._path, json.dumps(d, indent=1).encode())

    def __getitem__(self, k):
        return self._load()[k]

    def __setitem__(self, k, v):
        with self._lock:
            d = dict(self._load())
            d[k] = json.loads(json.dumps(_jsonable(v)))
            self._store(d)

    def __contains__(self, k):
        return k in self._load()

    def get(self, k, default=None):
        return self._load().get(k, default)

    def keys(self):
        return self._load().keys()

    def asdict(self):
        return dict(self._load())

    def update(self, other):
        with self._lock:
            d = dict(self._load())
            d.update(json.loads(json.dumps(_jsonable(dict(other)))))
            self._store(d)


class Array:
    """One zarr-v2 array: float32 (or any numpy dtype), C order, uncompressed chunks."""

    def __init__(self, path: Optional[str], meta: Optional[Dict[str, Any]] = None):
        self._path = path
        if meta is None:
            with open(os.path.join(path, ".zarray")) as f:
                meta = json.load(f)
        self.shape: Tuple[int, ...] = tuple(meta["shape"])
        self.chunks: Tuple[int, ...] = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.attrs = Attrs(None if path is None else os.path.join(path, ".zattrs"))
        self._mem: Dict[str, np.ndarray] = {}
        for s, c in zip(self.shape, self.chunks):
            if s % c != 0 and c != s:
                raise ValueError(f"zarr_store: shape {self.shape} must be a multiple of the chunks {self.chunks}")

    @property
    def ndim(self) -> int:
        return len(self.shape)

    # ---- chunk IO
    def _grid(self):
        return tuple(s // c for s, c in zip(self.shape, self.chunks))

    def _read_chunk(self, idx: Tuple[int, ...]) -> np.ndarray:
        key = ".".join(str(i) for i in idx)
        if self._path is None:
            return self._mem.get(key, np.zeros(self.chunks, self.dtype))
        p = os.path.join(self._path, key)
        if not os.path.exists(p):
            return np.zeros(self.chunks, self.dtype)
        return np.fromfile(p, dtype=self.dtype).reshape(self.chunks)

    def _write_chunk(self, idx: Tuple[int, ...], data: np.ndarray) -> None:
        key = ".".join(str(i) for i in idx)
        data = np.ascontiguousarray(data, dtype=self.dtype).reshape(self.chunks)
        if self._path is None:
            self._mem[key] = data.copy()
        else:
            _atomic_write(os.path.join(self._path, key), data.tobytes())

    def _chunk_range(self, index) -> Tuple[Tuple[slice, ...], bool]:
        """Normalise `index` (int / slice(None) / tuple of those) to per-axis slices aligned with the chunk grid."""
        if not isinstance(index, tuple):
            index = (index,)
        if any(i is Ellipsis for i in index):
            raise IndexError("zarr_store: Ellipsis is not supported")
        index = index + (slice(None),) * (self.ndim - len(index))
        sl, squeeze = [], []
        for ax, (i, s, c) in enumerate(zip(index, self.shape, self.chunks)):
            if isinstance(i, (int, np.integer)):
                i = int(i) + (s if i < 0 else 0)
                if not 0 <= i < s:
                    raise IndexError(f"index {i} out of bounds for axis {ax} with size {s}")
                if c != 1:
                    raise IndexError("zarr_store: integer indexing needs a chunk size of 1 along that axis")
                sl.append(slice(i, i + 1))
                squeeze.append(ax)
            elif isinstance(i, slice):
                a, b, st = i.indices(s)
                if st != 1 or a % c != 0 or (b % c != 0 and b != s):
                    raise IndexError("zarr_store: slices must be aligned with the chunk grid")
                sl.append(slice(a, b))
            else:
                raise IndexError(f"zarr_store: unsupported index {i!r}")
        return tuple(sl), tuple(squeeze)

    def __getitem__(self, index) -> np.ndarray:
        sl, squeeze = self._chunk_range(index)
        out = np.zeros(tuple(s.stop - s.start for s in sl), self.dtype)
        for cidx in np.ndindex(*[(s.stop - s.start + c - 1) // c for s, c in zip(sl, self.chunks)]):
            gidx = tuple(s.start // c + k for s, c, k in zip(sl, self.chunks, cidx))
            dst = tuple(slice(k * c, (k + 1) * c) for k, c in zip(cidx, self.chunks))
            out[dst] = self._read_chunk(gidx)
        return out.squeeze(axis=squeeze) if squeeze else out

    def __setitem__(self, index, value) -> None:
        sl, squeeze = self._chunk_range(index)
        value = np.asarray(value, dtype=self.dtype)
        full = tuple(s.stop - s.start for s in sl)
        if squeeze and value.ndim == self.ndim - len(squeeze):
            value = np.expand_dims(value, squeeze)
        value = np.broadcast_to(value, full)
        for cidx in np.ndindex(*[(f + c - 1) // c for f, c in zip(full, self.chunks)]):
            gidx = tuple(s.start // c + k for s, c, k in zip(sl, self.chunks, cidx))
            src = tuple(slice(k * c, (k + 1) * c) for k, c in zip(cidx, self.chunks))
            self._write_chunk(gidx, value[src])

    def __len__(self):
        return self.shape[0]


class Group:
    """A zarr-v2 group on a directory (path) or in memory (path=None, the reference's `zarr.group()`)."""

    def __init__(self, path: Optional[str] = None):
        self._path = path
        self._children: Dict[str, Any] = {}
        if path is not None:
            os.makedirs(path, exist_ok=True)
            zg = os.path.join(path, ".zgroup")
            if not os.path.exists(zg):
                _atomic_write(zg, json.dumps({"zarr_format": 2}).encode())
        self.attrs = Attrs(None if path is None else os.path.join(path, ".zattrs"))

    def _child_path(self, name: str) -> Optional[str]:
        return None if self._path is None else os.path.join(self._path, name)

    def __contains__(self, name: str) -> bool:
        if name in self._children:
            return True
        p = self._child_path(name)
        return p is not None and (os.path.exists(os.path.join(p, ".zarray")) or os.path.exists(os.path.join(p, ".zgroup")))

    def __getitem__(self, name: str):
        if name in self._children:
            return self._children[name]
        p = self._child_path(name)
        if p is not None and os.path.exists(os.path.join(p, ".zarray")):
            self._children[name] = Array(p)
        elif p is not None and os.path.exists(os.path.join(p, ".zgroup")):
            self._children[name] = Group(p)
        else:
            raise KeyError(name)
        return self._children[name]

    def keys(self) -> Iterable[str]:
        names = set(self._children)
        if self._path is not None:
            for n in os.listdir(self._path):
                if not n.startswith(".") and os.path.isdir(os.path.join(self._path, n)):
                    names.add(n)
        return sorted(names)

    def require_group(self, name: str) -> "Group":
        if name in self:
            g = self[name]
            if not isinstance(g, Group):
                raise ValueError(f"{name} exists and is not a group")
            return g
        g = Group(self._child_path(name))
        self._children[name] = g
        return g

    def create_dataset(self, name: str, shape=None, dtype="float32", chunks=None, data=None) -> Array:
        """zarr's group.create_dataset (util.py:685-707).  An existing dataset of identical geometry is re-opened (several
        ranks create the same datasets); anything else is an error."""
        if data is not None:
            data = np.asarray(data)
            shape = data.shape if shape is None else tuple(shape)
            dtype = data.dtype
        shape = tuple(int(s) for s in shape)
        chunks = shape if chunks is None else tuple(int(c) for c in chunks)
        dt = np.dtype(dtype)
        meta = {"zarr_format": 2, "shape": list(shape), "chunks": list(chunks), "dtype": dt.newbyteorder("<").str if dt.itemsize > 1 else dt.str,
                "compressor": None, "fill_value": 0, "order": "C", "filters": None}
        p = self._child_path(name)
        if name in self:
            ds = self[name]
            if not isinstance(ds, Array) or ds.shape != shape or ds.chunks != chunks:
                raise ValueError(f"dataset {name} exists with a different geometry")
        else:
            if p is not None:
                os.makedirs(p, exist_ok=True)
                _atomic_write(os.path.join(p, ".zarray"), json.dumps(meta, indent=1).encode())
            ds = Array(p, meta)
            self._children[name] = ds
        if data is not None:
            ds[(slice(None),) * len(shape)] = data
        return ds


def open_group(path=None, mode: str = "a") -> Group:
    """zarr.open(save_path, mode="a") / zarr.group() for path=None (util.py:1183-1196)."""
    if mode not in ("a", "r", "r+", "w"):
        raise ValueError(f"unsupported mode {mode!r}")
    if path is not None and mode in ("r", "r+") and not os.path.exists(os.path.join(str(path), ".zgroup")):
        raise FileNotFoundError(path)
    return Group(None if path is None else str(path))
