"""numpy_dataview -- host mirror of microscopes/common/recarray/_dataview.pyx:60-112.

Wraps a 1-D structured (optionally masked) numpy array exactly like the
reference: the record bytes are the AoS rows of row_major_dataview
(include/microscopes/common/recarray/dataview.hpp:194-217), the mask is one
bool per element.  ``to_device(ctx)`` hands both to msb_dataview_create.
"""
import ctypes as C
import hashlib

import numpy as np
import numpy.ma as ma

from . import _lib

# microscopes/common/_dataview.pyx get_c_type: numpy scalar dtype -> primitive_type
_NP_TO_PRIM = {
    np.dtype(np.bool_): _lib.TYPE_B, np.dtype(np.int8): _lib.TYPE_I8, np.dtype(np.uint8): _lib.TYPE_U8,
    np.dtype(np.int16): _lib.TYPE_I16, np.dtype(np.uint16): _lib.TYPE_U16, np.dtype(np.int32): _lib.TYPE_I32,
    np.dtype(np.uint32): _lib.TYPE_U32, np.dtype(np.int64): _lib.TYPE_I64, np.dtype(np.uint64): _lib.TYPE_U64,
    np.dtype(np.float32): _lib.TYPE_F32, np.dtype(np.float64): _lib.TYPE_F64,
}


def get_c_types(dtype):
    """structured dtype -> list of RuntimeType; sub-array fields become vectors."""
    types = []
    for name in dtype.names:
        ft = dtype.fields[name][0]
        if ft.subdtype is not None:
            base, shape = ft.subdtype
            if len(shape) != 1:
                raise ValueError("only 1-D sub-array fields are supported")
            types.append(_lib.RuntimeType(_NP_TO_PRIM[np.dtype(base)], int(shape[0]), 1))
        else:
            if ft not in _NP_TO_PRIM:
                raise ValueError("unsupported field type: %s" % ft)
            types.append(_lib.RuntimeType(_NP_TO_PRIM[ft], 1, 0))
    return types


class numpy_dataview(object):
    def __init__(self, npd):
        if npd is None:
            raise ValueError("npd cannot be None")
        if len(npd.shape) != 1:
            raise ValueError("1D (structural) arrays only")
        if len(npd.dtype) == 0:
            raise ValueError("structural arrays only")
        self._n = npd.shape[0]
        # the reference takes the dtype as numpy lays it out; row_major_dataview assumes
        # packed fields (running sum of sizes), so insist on an unpadded record
        packed = np.dtype([(n, npd.dtype.fields[n][0]) for n in npd.dtype.names])
        if packed.itemsize != npd.dtype.itemsize:
            raise ValueError("record dtype must be packed (align=False)")
        if hasattr(npd, "mask"):
            self._data = np.ascontiguousarray(npd.data)
            mask = np.ascontiguousarray(ma.getmaskarray(npd))
            self._mask = mask
        else:
            self._data = np.ascontiguousarray(npd)
            self._mask = None
        self._types = get_c_types(self._data.dtype)
        self._pi = None
        self._dev = {}

    # -- reference API ----------------------------------------------------
    def size(self):
        return self._n

    def __len__(self):
        return self._n

    def types(self):
        return list(self._types)

    def __iter__(self):
        order = range(self._n) if self._pi is None else self._pi
        for i in order:
            yield self.get(int(i))

    def get(self, idx):
        """row idx as a numpy record, masked if any element is (abstract_dataview.next)"""
        if idx < 0 or idx >= self._n:
            raise RuntimeError("invalid position")
        if self._mask is None or not any(np.any(self._mask[idx][n]) for n in self._mask.dtype.names):
            return self._data[idx]
        return ma.array(self._data[idx:idx + 1], mask=self._mask[idx:idx + 1])[0]

    def permute(self, rng):
        """Fisher-Yates order for iteration (recarray/dataview.cpp:141-145, util.hpp:85-94)"""
        pi = np.arange(self._n)
        for i in range(self._n - 1, 0, -1):
            j = int(rng.integers(0, i + 1))
            pi[i], pi[j] = pi[j], pi[i]
        self._pi = pi

    def reset_permutation(self):
        self._pi = None

    def digest(self, h=None):
        h = hashlib.sha1() if h is None else h
        h.update((type(self).__module__ + "." + type(self).__name__).encode())
        if self._mask is not None:
            raise NotImplementedError("masked arrays digest not implemented")
        h.update(str(self._data.dtype).encode())
        h.update(self._data.view(np.uint8))
        return h

    def __reduce__(self):
        if self._mask is None:
            return (numpy_dataview, (self._data,))
        return (numpy_dataview, (ma.array(self._data, mask=self._mask),))

    # -- device side --------------------------------------------------------
    def raw(self):
        """(record bytes, mask bytes or None) exactly as the C ABI reads them"""
        data = self._data.view(np.uint8).reshape(self._n, -1) if self._n else np.zeros((0, 1), np.uint8)
        mask = None
        if self._mask is not None:
            mask = self._mask.view(np.uint8).reshape(self._n, -1) if self._n else np.zeros((0, 1), np.uint8)
        return data, mask

    def to_device(self, ctx):
        key = id(ctx)
        if key not in self._dev:
            self._dev[key] = device_dataview(ctx, self)
        return self._dev[key]


class device_dataview(object):
    """Owner of an msb_dataview handle."""

    def __init__(self, ctx, view=None, *, data=None, mask=None, n=None, types=None, on_device=False):
        lib = _lib.load()
        self._ctx = ctx
        if view is not None:
            data_np, mask_np = view.raw()
            self._keep = (data_np, mask_np)
            types = view.types()
            n = view.size()
            dptr = data_np.ctypes.data if n else None
            mptr = mask_np.ctypes.data if (mask_np is not None and n) else None
        else:
            dptr, mptr = data, mask
        arr = (_lib.RuntimeType * len(types))(*types)
        h = C.c_void_p()
        _lib.check(lib.msb_dataview_create(ctx.handle, dptr, mptr, n, arr, len(types), 1 if on_device else 0, C.byref(h)))
        ctx.synchronize()
        self._h = h
        self._n = n
        self._types = list(types)
        ctx._adopt(self)

    @property
    def handle(self):
        return self._h

    def size(self):
        return self._n

    def rowsize(self):
        a, b = C.c_size_t(), C.c_size_t()
        _lib.check(_lib.load().msb_dataview_rowsize(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def upload(self, data, mask=None):
        """replace the records in place from host memory (address or uint8 array); same n, same types"""
        dptr = data.ctypes.data if hasattr(data, "ctypes") else data
        mptr = mask.ctypes.data if hasattr(mask, "ctypes") else mask
        _lib.check(_lib.load().msb_dataview_upload(self._h, dptr, mptr))

    def permute(self, seed):
        """row_major_dataview::permute (dataview.cpp:141-145): a Fisher-Yates iteration order from the Philox stream;
        get_row_bytes(i) then returns record pi[i] until reset_permutation()"""
        _lib.check(_lib.load().msb_dataview_permute(self._h, int(seed)))

    def reset_permutation(self):
        _lib.check(_lib.load().msb_dataview_reset_permutation(self._h))

    def permutation(self):
        """pi as an array (the identity when no permutation is set)"""
        n = self.size() if hasattr(self, "size") else self._n
        out = np.zeros(n, np.uint64)
        _lib.check(_lib.load().msb_dataview_permutation(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        return out

    def get_row_bytes(self, idx):
        rs, ms = self.rowsize()
        row = np.zeros(rs, np.uint8)
        msk = np.zeros(max(ms, 1), np.uint8)
        _lib.check(_lib.load().msb_dataview_get_row(self._h, idx, row.ctypes.data, msk.ctypes.data))
        return row, msk[:ms]

    def close(self):
        if self._h:
            _lib.load().msb_dataview_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
