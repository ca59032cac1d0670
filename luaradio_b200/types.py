"""Sample types and Vector, binary-compatible with the reference's CStruct types.

ComplexFloat32 == `struct {float real, imag}` (radio/types/complexfloat32.lua:19-24) == numpy complex64
== CUDA float2; Float32 == `struct {float value}` (radio/types/float32.lua:17-21) == numpy float32.
Vector mirrors radio/core/vector.lua:19-37,108-136: contiguous, `.data/.length/.size`, grow-only resize.
"""
import numpy as np


class DataType:
    def __init__(self, type_name, dtype):
        self.type_name = type_name
        self.dtype = np.dtype(dtype)
        self.size = self.dtype.itemsize

    def vector(self, num=0):
        return Vector(self, num)

    def vector_from_array(self, arr):
        v = Vector(self, 0)
        v._buf = np.ascontiguousarray(np.asarray(arr).astype(self.dtype))
        v.length = v._buf.shape[0]
        return v

    def __repr__(self):
        return self.type_name


ComplexFloat32 = DataType("ComplexFloat32", np.complex64)
Float32 = DataType("Float32", np.float32)


def type_of(array):
    a = np.asarray(array)
    if a.dtype == np.complex64 or np.iscomplexobj(a):
        return ComplexFloat32
    return Float32


class Vector:
    """Contiguous typed sample vector; `.data` is a numpy view of the first `.length` elements."""

    def __init__(self, data_type, num=0):
        self.data_type = data_type
        self._buf = np.zeros(int(num), dtype=data_type.dtype)   # zero-filled like Vector.new (vector.lua:31-32)
        self.length = int(num)

    @property
    def data(self):
        return self._buf[:self.length]

    @property
    def size(self):
        return self.length * self.data_type.size

    def resize(self, num):
        num = int(num)
        if num > self._buf.shape[0]:
            nb = np.zeros(num, dtype=self.data_type.dtype)
            nb[:self.length] = self._buf[:self.length]
            self._buf = nb
        self.length = num
        return self

    def append(self, elem):
        self.resize(self.length + 1)
        self._buf[self.length - 1] = elem
        return self

    def ctypes_ptr(self):
        return self._buf.ctypes.data

    def __len__(self):
        return self.length

    @classmethod
    def cast(cls, array):
        """Zero-copy view of a numpy array (vector.lua:48-65 Vector.cast)."""
        a = np.ascontiguousarray(array)
        dt = ComplexFloat32 if a.dtype == np.complex64 else Float32
        if a.dtype != dt.dtype:
            a = a.astype(dt.dtype)
        v = cls(dt, 0)
        v._buf = a
        v.length = a.shape[0]
        return v
