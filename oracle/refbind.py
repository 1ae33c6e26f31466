"""oracle/refbind.py -- TEST INFRASTRUCTURE (ctypes binding of oracle/_ref/libmars_ref.so).

The .so is the reference's own portable C path (see oracle/build_ref.sh).  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libmars_ref.so")
REF_MODELS = os.path.join(HERE, "_ref", "models")


class TensorDesc(C.Structure):
    _pack_ = 1
    _fields_ = [("id", C.c_uint32), ("name", C.c_char * 60), ("dtype", C.c_int32),
                ("format", C.c_int32), ("ndims", C.c_uint32), ("shape", C.c_int32 * 6),
                ("data_offset", C.c_uint64), ("data_size", C.c_uint64),
                ("scale", C.c_float), ("zero_point", C.c_int32)]


class RuntimeTensor(C.Structure):
    _fields_ = [("desc", TensorDesc), ("vaddr", C.c_void_p), ("paddr", C.c_void_p),
                ("alloc_size", C.c_size_t), ("is_external", C.c_bool)]


class LayerDesc(C.Structure):
    _pack_ = 1
    _fields_ = [("id", C.c_uint32), ("type", C.c_int32), ("num_inputs", C.c_uint32),
                ("num_outputs", C.c_uint32), ("input_tensor_ids", C.c_uint32 * 4),
                ("output_tensor_ids", C.c_uint32 * 4), ("params", C.c_uint8 * 64)]


class RuntimeLayer(C.Structure):
    _fields_ = [("desc", LayerDesc), ("is_executed", C.c_bool)]


class Header(C.Structure):
    _pack_ = 1
    _fields_ = [("magic", C.c_uint32), ("version_major", C.c_uint16), ("version_minor", C.c_uint16),
                ("flags", C.c_uint32), ("num_layers", C.c_uint32), ("num_tensors", C.c_uint32),
                ("num_inputs", C.c_uint32), ("num_outputs", C.c_uint32),
                ("weights_offset", C.c_uint64), ("weights_size", C.c_uint64),
                ("input_tensor_ids", C.c_uint32 * 4), ("output_tensor_ids", C.c_uint32 * 4)]


class Model(C.Structure):
    _fields_ = [("header", Header), ("tensors", C.POINTER(RuntimeTensor)),
                ("layers", C.POINTER(RuntimeLayer)), ("ddr_base", C.c_void_p),
                ("ddr_paddr", C.c_void_p), ("ddr_size", C.c_size_t), ("oram_base", C.c_void_p),
                ("oram_paddr", C.c_void_p), ("oram_size", C.c_size_t), ("weights", C.c_void_p),
                ("weights_size", C.c_size_t), ("total_inference_us", C.c_uint64),
                ("inference_count", C.c_uint32)]


assert C.sizeof(TensorDesc) == 124 and C.sizeof(LayerDesc) == 112 and C.sizeof(Header) == 76

DET_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"),
                      ("conf", "<f4"), ("cls", "<i4")])
DTYPE_SIZE = {0: 4, 1: 4, 2: 2, 3: 1, 4: 1, 5: 1}


def fnv1a64(buf) -> str:
    """FNV-1a 64-bit over a bytes-like object (the hash SURVEY §6.2 records)."""
    h = 0xCBF29CE484222325
    data = np.frombuffer(memoryview(buf).cast("B"), dtype=np.uint8)
    # vectorising FNV is not possible (sequential); do it in chunks in pure python ints
    for b in data.tobytes():
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def bind_mars_api(lib):
    """Declare the mars_* API on a loaded library (same ABI for reference and product)."""
    PM = C.POINTER(Model)
    lib.mars_load_file.restype = C.c_int
    lib.mars_load_file.argtypes = [C.c_char_p, C.POINTER(PM)]
    lib.mars_load_memory.restype = C.c_int
    lib.mars_load_memory.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(PM)]
    lib.mars_free.restype = None
    lib.mars_free.argtypes = [PM]
    lib.mars_run.restype = C.c_int
    lib.mars_run.argtypes = [PM]
    for f in (lib.mars_get_input, lib.mars_get_output):
        f.restype = C.POINTER(RuntimeTensor)
        f.argtypes = [PM, C.c_int]
    for f in (lib.mars_get_num_inputs, lib.mars_get_num_outputs):
        f.restype = C.c_int
        f.argtypes = [PM]
    lib.mars_get_error_string.restype = C.c_char_p
    lib.mars_get_error_string.argtypes = [C.c_int]
    lib.mars_print_summary.restype = None
    lib.mars_print_summary.argtypes = [PM]
    lib.mars_vec_add_f32.restype = None
    lib.mars_vec_add_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mars_vec_dot_f32.restype = C.c_float
    lib.mars_vec_dot_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mars_matmul_f32.restype = None
    lib.mars_matmul_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    return lib


def tensor_numel(desc) -> int:
    n = 1
    for i in range(desc.ndims):
        n *= desc.shape[i]
    return n


class RefRuntime:
    """One loaded model inside the reference library (one live model per process-arena)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " missing: run oracle/build_ref.sh")
            lib = bind_mars_api(C.CDLL(REF_SO))
            lib.oracle_ref_arena_config.restype = C.c_int
            lib.oracle_ref_arena_config.argtypes = [C.c_size_t]
            lib.oracle_ref_arena.restype = C.c_void_p
            lib.oracle_ref_arena_size.restype = C.c_size_t
            lib.oracle_ref_run_layer.restype = C.c_int
            lib.oracle_ref_run_layer.argtypes = [C.POINTER(Model), C.c_uint32]
            lib.oracle_ref_parse_output.restype = C.c_int
            lib.oracle_ref_parse_output.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
            lib.oracle_ref_nms.restype = C.c_int
            lib.oracle_ref_nms.argtypes = [C.c_void_p, C.c_int, C.c_float]
            if hasattr(lib, "oracle_ref_cpp_preprocess"):
                lib.oracle_ref_cpp_preprocess.restype = C.c_int
                lib.oracle_ref_cpp_preprocess.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
            if hasattr(lib, "oracle_ref_load_image"):
                lib.oracle_ref_load_image.restype = C.c_int
                lib.oracle_ref_load_image.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
            lib.oracle_ref_cpp_nms.restype = C.c_int
            lib.oracle_ref_cpp_nms.argtypes = [C.c_void_p, C.c_int, C.c_float]
            lib.oracle_ref_cpp_iou.restype = C.c_float
            lib.oracle_ref_cpp_iou.argtypes = [C.c_void_p, C.c_void_p]
            lib.oracle_ref_cpp_scale_detections.restype = None
            lib.oracle_ref_cpp_scale_detections.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
            lib.oracle_ref_cpp_sigmoid.restype = C.c_float
            lib.oracle_ref_cpp_sigmoid.argtypes = [C.c_float]
            lib.oracle_ref_cpp_anchors.restype = None
            lib.oracle_ref_cpp_anchors.argtypes = [C.c_void_p, C.c_void_p]
            cls._lib = lib
        return cls._lib

    def __init__(self, model, arena_bytes=8 << 20):
        """model: path or bytes.  arena_bytes: 8 MiB = the reference's literal."""
        lib = self.lib()
        if lib.oracle_ref_arena_config(arena_bytes) != 0:
            raise MemoryError("arena")
        self.arena_bytes = arena_bytes
        self.m = C.POINTER(Model)()
        if isinstance(model, (bytes, bytearray, memoryview)):
            self._blob = bytes(model)
            err = lib.mars_load_memory(self._blob, len(self._blob), C.byref(self.m))
        else:
            err = lib.mars_load_file(os.fsencode(model), C.byref(self.m))
        if err != 0:
            raise RuntimeError("reference mars_load failed: %d" % err)

    # -- views -------------------------------------------------------------
    def arena(self) -> np.ndarray:
        lib = self.lib()
        p = lib.oracle_ref_arena()
        return np.ctypeslib.as_array((C.c_uint8 * self.arena_bytes).from_address(p))

    def tensor(self, rt, nbytes=None) -> np.ndarray:
        n = rt.alloc_size if nbytes is None else nbytes
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(rt.vaddr))

    def input(self, i=0):
        return self.lib().mars_get_input(self.m, i).contents

    def output(self, i=0):
        return self.lib().mars_get_output(self.m, i).contents

    def output_bytes(self, i=0) -> np.ndarray:
        o = self.output(i)
        n = tensor_numel(o.desc) * DTYPE_SIZE.get(o.desc.dtype, 1)
        return self.tensor(o, n)

    def set_input(self, data: np.ndarray, i=0):
        t = self.input(i)
        raw = np.ascontiguousarray(data).view(np.uint8).ravel()
        assert raw.size <= t.alloc_size
        self.tensor(t)[: raw.size] = raw

    def run(self):
        err = self.lib().mars_run(self.m)
        if err != 0:
            raise RuntimeError("reference mars_run failed: %d" % err)

    def run_layer(self, i):
        return self.lib().oracle_ref_run_layer(self.m, i)

    def close(self):
        if self.m:
            self.lib().mars_free(self.m)
            self.m = None


def ref_parse_output(out_i8: np.ndarray, npred: int, scale: float, maxd=1000) -> np.ndarray:
    lib = RefRuntime.lib()
    dets = np.zeros(maxd, dtype=DET_DTYPE)
    buf = np.ascontiguousarray(out_i8.view(np.int8))
    n = lib.oracle_ref_parse_output(buf.ctypes.data, npred, scale, dets.ctypes.data, maxd)
    return dets[:n].copy()


def ref_nms(dets: np.ndarray, thresh=0.45) -> np.ndarray:
    lib = RefRuntime.lib()
    d = np.ascontiguousarray(dets.copy())
    n = lib.oracle_ref_nms(d.ctypes.data, len(d), thresh)
    return d[:n].copy()


def ref_load_image(rgb: np.ndarray, tw: int, th: int, nhwc: bool) -> np.ndarray:
    """the reference's own load_image() (src/mars/mars_yolo_test.c:40-77) on an [h, w, 3] uint8 frame: the frame goes
    through a binary PPM file so that the unmodified function (stb_image decode included) runs.  Returns tw*th*3 int8."""
    import tempfile
    lib = RefRuntime.lib()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.zeros(tw * th * 3, dtype=np.int8)
    ow, oh = C.c_int(0), C.c_int(0)
    with tempfile.NamedTemporaryFile(suffix=".ppm") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(rgb.tobytes())
        f.flush()
        rc = lib.oracle_ref_load_image(f.name.encode(), tw, th, 1 if nhwc else 0, out.ctypes.data, C.byref(ow), C.byref(oh))
    if rc != 0 or (ow.value, oh.value) != (w, h):
        raise RuntimeError("reference load_image failed (rc %d, %dx%d)" % (rc, ow.value, oh.value))
    return out


def ref_preprocess_rgba(rgb: np.ndarray) -> np.ndarray:
    """the reference's own load_and_preprocess_image() (examples/yolo_detect.cpp:72-130) on an [h, w, 3] uint8 frame, through a
    binary PPM file.  Returns the 640*640*4 uint8 RGBA frame."""
    import tempfile
    lib = RefRuntime.lib()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.zeros(640 * 640 * 4, dtype=np.uint8)
    ow, oh = C.c_int(0), C.c_int(0)
    with tempfile.NamedTemporaryFile(suffix=".ppm") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(rgb.tobytes())
        f.flush()
        rc = lib.oracle_ref_cpp_preprocess(f.name.encode(), out.ctypes.data, C.byref(ow), C.byref(oh))
    if rc != 0 or (ow.value, oh.value) != (w, h):
        raise RuntimeError("reference load_and_preprocess_image failed (rc %d)" % rc)
    return out


def pattern_p0(nbytes: int) -> np.ndarray:
    """int8 test pattern of reference src/mars/mars_test.c:82-85: p[i] = i % 127."""
    return (np.arange(nbytes, dtype=np.int64) % 127).astype(np.int8)


def pattern_f32(nfloats: int) -> np.ndarray:
    """float pattern of reference src/mars/mars_test.c:76-80: (i % 256) / 255."""
    return ((np.arange(nfloats, dtype=np.int64) % 256).astype(np.float32) / np.float32(255.0)).astype(np.float32)
