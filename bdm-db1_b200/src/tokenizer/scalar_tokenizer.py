"""Gato mu-law scalar tokenizer (API of src/tokenizer/scalar_tokenizer.py:20-63), computed by the host C++ library
(libdb1_host.so: db1_discretize / db1_decode) so the integer result is reproducible bit-for-bit."""
import ctypes as C

import numpy as np
import torch

from db1_sm100 import _lib


class ContinuousScalarTokenizer:
    def __init__(self, num_continuous_bin: int = 1024, mu: float = 100.0, M: float = 256.0):
        self.num_continuous_bin = num_continuous_bin
        self.mu = mu
        self.M = M

    def discretize(self, x, is_action: bool):
        """float scalars -> int32 bin ids in [0, num_continuous_bin); mu-law companding unless `is_action`."""
        if isinstance(x, torch.Tensor):
            arr = x.detach().cpu().numpy()
        else:
            arr = np.asarray(x)
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        out = np.empty(arr.shape, dtype=np.int32)
        rc = _lib.hostlib().db1_discretize(arr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                           C.c_longlong(arr.size), int(bool(is_action)), self.num_continuous_bin,
                                           C.c_float(self.mu), C.c_float(self.M))
        if rc != 0:
            raise _lib.Db1Error("db1_discretize failed rc=%d" % rc)
        return torch.from_numpy(out)

    def decode(self, x, is_action: bool):
        if isinstance(x, torch.Tensor):
            arr = x.detach().cpu().numpy()
        else:
            arr = np.asarray(x)
        if arr.size and (arr.max() >= self.num_continuous_bin or arr.min() < 0):
            print("Warning of exceeded range of discrete number to recontruct, by default values will be cliped, "
                  "min: {}, max:{}".format(arr.min(), arr.max()))
        arr = np.ascontiguousarray(arr, dtype=np.int32)
        out = np.empty(arr.shape, dtype=np.float32)
        rc = _lib.hostlib().db1_decode(arr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                       C.c_longlong(arr.size), int(bool(is_action)), self.num_continuous_bin,
                                       C.c_float(self.mu), C.c_float(self.M))
        if rc != 0:
            raise _lib.Db1Error("db1_decode failed rc=%d" % rc)
        return torch.from_numpy(out)
