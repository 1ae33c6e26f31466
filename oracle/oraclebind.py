"""oracle/oraclebind.py -- TEST INFRASTRUCTURE (ctypes binding of our CPU restatement,
oracle/_build/libmars_oracle.so).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs import this."""
import ctypes as C
import os
import subprocess
import numpy as np

from .refbind import TensorDesc, LayerDesc, DET_DTYPE, DTYPE_SIZE, tensor_numel  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libmars_oracle.so")
BOX_DTYPE = np.dtype([("x0", "<f4"), ("y0", "<f4"), ("x1", "<f4"), ("y1", "<f4"),
                      ("confidence", "<f4"), ("class_id", "<i4")])

_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.mo_load.restype = C.c_int
        L.mo_load.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]
        L.mo_free.argtypes = [C.c_void_p]
        L.mo_run.restype = C.c_int
        L.mo_run.argtypes = [C.c_void_p]
        L.mo_run_layer.restype = C.c_int
        L.mo_run_layer.argtypes = [C.c_void_p, C.c_uint32]
        L.mo_set_depthwise_mode.argtypes = [C.c_void_p, C.c_int]
        L.mo_arena.restype = C.c_void_p
        L.mo_arena.argtypes = [C.c_void_p]
        for f in ("mo_arena_size", "mo_weights_size", "mo_buffer_size"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        L.mo_num_buffers.restype = C.c_int
        L.mo_num_buffers.argtypes = [C.c_void_p]
        for f in ("mo_num_layers", "mo_num_tensors"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        L.mo_tensor_desc.restype = C.POINTER(TensorDesc)
        L.mo_tensor_desc.argtypes = [C.c_void_p, C.c_uint32]
        L.mo_layer_desc.restype = C.POINTER(LayerDesc)
        L.mo_layer_desc.argtypes = [C.c_void_p, C.c_uint32]
        for f in ("mo_tensor_offset", "mo_tensor_alloc"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p, C.c_uint32]
        for f in ("mo_input_index", "mo_output_index"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.mo_parse_output.restype = C.c_int
        L.mo_parse_output.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
        L.mo_nms.restype = C.c_int
        L.mo_nms.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.mo_nms_corner.restype = C.c_int
        L.mo_nms_corner.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.mo_iou_corner.restype = C.c_float
        L.mo_iou_corner.argtypes = [C.c_void_p, C.c_void_p]
        L.mo_scale_detections.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mo_decode_anchor_grid.restype = C.c_int
        L.mo_decode_anchor_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                            C.c_void_p, C.c_int, C.c_int]
        L.mo_vec_dot_f32.restype = C.c_float
        L.mo_vec_dot_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for f in ("mo_vec_add_f32", "mo_vec_sub_f32", "mo_vec_mul_f32"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.mo_vec_relu_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.mo_matmul_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
        _lib = L
    return _lib


class OracleModel:
    """A .mars model loaded into the restatement oracle with its own zeroed arena."""

    def __init__(self, model, arena_bytes=8 << 20, depthwise=False):
        L = lib()
        blob = model if isinstance(model, (bytes, bytearray)) else open(model, "rb").read()
        self._blob = bytes(blob)
        self.h = C.c_void_p()
        err = L.mo_load(self._blob, len(self._blob), arena_bytes, C.byref(self.h))
        if err != 0:
            raise RuntimeError("oracle mo_load failed: %d" % err)
        if depthwise:
            L.mo_set_depthwise_mode(self.h, 1)
        self.arena_bytes = L.mo_arena_size(self.h)
        self.weights_size = L.mo_weights_size(self.h)
        self.buffer_size = L.mo_buffer_size(self.h)
        self.num_buffers = L.mo_num_buffers(self.h)
        self.num_layers = L.mo_num_layers(self.h)
        self.num_tensors = L.mo_num_tensors(self.h)

    def arena(self) -> np.ndarray:
        p = lib().mo_arena(self.h)
        return np.ctypeslib.as_array((C.c_uint8 * self.arena_bytes).from_address(p))

    def tensor_desc(self, idx):
        return lib().mo_tensor_desc(self.h, idx).contents

    def layer_desc(self, idx):
        return lib().mo_layer_desc(self.h, idx).contents

    def tensor_offset(self, idx):
        return lib().mo_tensor_offset(self.h, idx)

    def input_index(self, i=0):
        return lib().mo_input_index(self.h, i)

    def output_index(self, i=0):
        return lib().mo_output_index(self.h, i)

    def set_input(self, data: np.ndarray, i=0):
        raw = np.ascontiguousarray(data).view(np.uint8).ravel()
        off = self.tensor_offset(self.input_index(i))
        self.arena()[off: off + raw.size] = raw

    def output_bytes(self, i=0) -> np.ndarray:
        idx = self.output_index(i)
        d = self.tensor_desc(idx)
        n = tensor_numel(d) * DTYPE_SIZE.get(d.dtype, 1)
        off = self.tensor_offset(idx)
        return self.arena()[off: off + n]

    def run(self):
        err = lib().mo_run(self.h)
        if err != 0:
            raise RuntimeError("oracle mo_run failed: %d" % err)

    def run_layer(self, i):
        return lib().mo_run_layer(self.h, i)

    def close(self):
        if self.h:
            lib().mo_free(self.h)
            self.h = None


def parse_output(out_i8: np.ndarray, npred: int, scale: float, maxd=1000) -> np.ndarray:
    dets = np.zeros(maxd, dtype=DET_DTYPE)
    buf = np.ascontiguousarray(out_i8.view(np.int8))
    n = lib().mo_parse_output(buf.ctypes.data, npred, scale, dets.ctypes.data, maxd)
    return dets[:n].copy()


def nms(dets: np.ndarray, thresh=0.45) -> np.ndarray:
    d = np.ascontiguousarray(dets.copy())
    n = lib().mo_nms(d.ctypes.data, len(d), thresh)
    return d[:n].copy()


def nms_corner(boxes: np.ndarray, thresh=0.45) -> np.ndarray:
    d = np.ascontiguousarray(boxes.copy())
    n = lib().mo_nms_corner(d.ctypes.data, len(d), thresh)
    return d[:n].copy()
