"""capi.py -- ctypes binding of libmars_b200.so (the C-ABI boundary).

Mirrors the reference's public interface for the path: the ten mars_* functions of
include/mars_runtime.h:79-138, mars_math.h, the nna_* bring-up calls, plus the additive
mars_b200_* batch entry points (include/mars_b200.h).  No torch types cross this boundary:
buffers are numpy arrays or raw addresses (e.g. `tensor.data_ptr()` of pinned memory).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MARS_B200_LIB") or os.path.join(HERE, "lib", "libmars_b200.so")  # (override: tuning builds, -DMARS_TC_TUNING)


class MarsLibraryMissing(RuntimeError):
    pass


class MarsError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__("%s failed: %d (%s)%s" % (what, code, ERR_NAMES.get(code, "?"), (": " + detail) if detail else ""))


ERR_NAMES = {0: "OK", -1: "INVALID_MAGIC", -2: "VERSION_MISMATCH", -3: "ALLOC_FAILED", -4: "INVALID_FILE",
             -5: "NNA_INIT_FAILED", -6: "LAYER_FAILED", -7: "INVALID_TENSOR", -8: "INVALID_LAYER"}


class TensorDesc(C.Structure):
    _pack_ = 1
    _fields_ = [("id", C.c_uint32), ("name", C.c_char * 60), ("dtype", C.c_int32), ("format", C.c_int32),
                ("ndims", C.c_uint32), ("shape", C.c_int32 * 6), ("data_offset", C.c_uint64),
                ("data_size", C.c_uint64), ("scale", C.c_float), ("zero_point", C.c_int32)]


class RuntimeTensor(C.Structure):
    _fields_ = [("desc", TensorDesc), ("vaddr", C.c_void_p), ("paddr", C.c_void_p), ("alloc_size", C.c_size_t),
                ("is_external", C.c_bool)]


class LayerDesc(C.Structure):
    _pack_ = 1
    _fields_ = [("id", C.c_uint32), ("type", C.c_int32), ("num_inputs", C.c_uint32), ("num_outputs", C.c_uint32),
                ("input_tensor_ids", C.c_uint32 * 4), ("output_tensor_ids", C.c_uint32 * 4), ("params", C.c_uint8 * 64)]


class RuntimeLayer(C.Structure):
    _fields_ = [("desc", LayerDesc), ("is_executed", C.c_bool)]


class Header(C.Structure):
    _pack_ = 1
    _fields_ = [("magic", C.c_uint32), ("version_major", C.c_uint16), ("version_minor", C.c_uint16),
                ("flags", C.c_uint32), ("num_layers", C.c_uint32), ("num_tensors", C.c_uint32),
                ("num_inputs", C.c_uint32), ("num_outputs", C.c_uint32), ("weights_offset", C.c_uint64),
                ("weights_size", C.c_uint64), ("input_tensor_ids", C.c_uint32 * 4), ("output_tensor_ids", C.c_uint32 * 4)]


class Model(C.Structure):
    _fields_ = [("header", Header), ("tensors", C.POINTER(RuntimeTensor)), ("layers", C.POINTER(RuntimeLayer)),
                ("ddr_base", C.c_void_p), ("ddr_paddr", C.c_void_p), ("ddr_size", C.c_size_t),
                ("oram_base", C.c_void_p), ("oram_paddr", C.c_void_p), ("oram_size", C.c_size_t),
                ("weights", C.c_void_p), ("weights_size", C.c_size_t), ("total_inference_us", C.c_uint64),
                ("inference_count", C.c_uint32)]


assert C.sizeof(TensorDesc) == 124 and C.sizeof(LayerDesc) == 112 and C.sizeof(Header) == 76

DET_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"), ("conf", "<f4"), ("cls", "<i4")])
BOX_DTYPE = np.dtype([("x0", "<f4"), ("y0", "<f4"), ("x1", "<f4"), ("y1", "<f4"), ("confidence", "<f4"), ("class_id", "<i4")])

PM = C.POINTER(Model)

# every symbol include/*.h declares: name -> (restype, argtypes)
SIGNATURES = {
    # include/mars_runtime.h
    "mars_load_file": (C.c_int, [C.c_char_p, C.POINTER(PM)]),
    "mars_load_memory": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(PM)]),
    "mars_free": (None, [PM]),
    "mars_get_input": (C.POINTER(RuntimeTensor), [PM, C.c_int]),
    "mars_get_output": (C.POINTER(RuntimeTensor), [PM, C.c_int]),
    "mars_run": (C.c_int, [PM]),
    "mars_get_error_string": (C.c_char_p, [C.c_int]),
    "mars_get_num_inputs": (C.c_int, [PM]),
    "mars_get_num_outputs": (C.c_int, [PM]),
    "mars_print_summary": (None, [PM]),
    # include/mars_math.h
    "mars_vec_add_f32": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mars_vec_dot_f32": (C.c_float, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mars_matmul_f32": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    # include/mxu_ops.h
    "mxu_init": (None, [C.c_void_p]),
    "mxu_is_initialized": (C.c_int, []),
    "mxu_mul_f32": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mxu_add_f32": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mxu_sub_f32": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mxu_relu_f32": (None, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "conv2d_int8_mxu": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]),
    "conv2d_int8_nhwc_mxu": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]),
    "conv2d_float32_mxu": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    # include/nna.h, include/nna_memory.h
    "nna_init": (C.c_int, []),
    "nna_deinit": (None, []),
    "nna_get_hw_info": (C.c_int, [C.c_void_p]),
    "nna_is_ready": (C.c_int, []),
    "nna_get_version": (C.c_char_p, []),
    "nna_lock": (C.c_int, []),
    "nna_unlock": (C.c_int, []),
    "nna_device_get_ddr": (C.c_void_p, []),
    "nna_device_get_ddr_pbase": (C.c_uint32, []),
    "nna_device_get_oram": (C.c_void_p, []),
    "nna_device_get_fd": (C.c_int, []),
    "nna_device_get_memfd": (C.c_int, []),
    "nna_device_get_nndma_io": (C.c_void_p, []),
    "nna_device_get_nndma_desram": (C.c_void_p, []),
    "nna_malloc": (C.c_void_p, [C.c_size_t]),
    "nna_memalign": (C.c_void_p, [C.c_size_t, C.c_size_t]),
    "nna_calloc": (C.c_void_p, [C.c_size_t, C.c_size_t]),
    "nna_free": (None, [C.c_void_p]),
    "nna_oram_malloc": (C.c_void_p, [C.c_size_t]),
    "nna_oram_free": (None, [C.c_void_p]),
    "nna_oram_get_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nna_cache_flush": (None, [C.c_void_p, C.c_size_t]),
    "nna_cache_invalidate": (None, [C.c_void_p, C.c_size_t]),
    # include/mars_b200.h
    "mars_b200_version": (C.c_char_p, []),
    "mars_b200_device_count": (C.c_int, []),
    "mars_b200_set_device": (C.c_int, [C.c_int]),
    "mars_b200_set_arena_bytes": (None, [C.c_size_t]),
    "mars_b200_last_error": (C.c_char_p, []),
    "mars_b200_arena_upload": (C.c_int, [PM, C.c_int]),
    "mars_b200_arena_download": (C.c_int, [PM, C.c_int, C.c_void_p, C.c_size_t]),
    "mars_b200_arena_clear": (C.c_int, [PM]),
    "mars_b200_run_layer": (C.c_int, [PM, C.c_uint32]),
    "mars_b200_describe": (C.c_size_t, [PM, C.c_char_p, C.c_size_t]),
    "mars_b200_set_opt_level": (None, [PM, C.c_int]),
    "mars_b200_set_depthwise_mode": (None, [PM, C.c_int]),
    "mars_b200_set_strict": (None, [C.c_int]),
    "mars_b200_submit_run_batch": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mars_b200_group_load": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mars_b200_group_free": (None, [C.c_void_p]),
    "mars_b200_group_size": (C.c_int, [C.c_void_p]),
    "mars_b200_group_model": (PM, [C.c_void_p, C.c_int]),
    "mars_b200_group_detect_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float]),
    "mars_b200_set_f32_mode": (None, [PM, C.c_int]),
    "mars_b200_set_batch": (C.c_int, [PM, C.c_int]),
    "mars_b200_get_batch": (C.c_int, [PM]),
    "mars_b200_input_bytes": (C.c_size_t, [PM]),
    "mars_b200_output_bytes": (C.c_size_t, [PM]),
    "mars_b200_upload_inputs": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "mars_b200_download_outputs": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "mars_b200_preprocess_batch": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "mars_b200_run_resident": (C.c_int, [PM, C.c_int, C.c_int]),
    "mars_b200_detect_resident": (C.c_int, [PM, C.c_int, C.c_int, C.c_float]),
    "mars_b200_step_resident": (C.c_int, [PM, C.c_int, C.c_int, C.c_float, C.c_int]),
    "mars_b200_enqueue_step_resident": (C.c_int, [PM, C.c_int, C.c_int, C.c_float, C.c_int]),
    "mars_b200_download_detections": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "mars_b200_detect_batch": (C.c_int, [PM, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float]),
    "mars_b200_submit_batch": (C.c_int, [PM, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float]),
    "mars_b200_wait_batch": (C.c_int, [PM, C.c_int]),
    "mars_b200_run_batch": (C.c_int, [PM, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mars_b200_detections_device": (None, [PM, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mars_b200_launch_count": (C.c_uint64, [PM]),
    "mars_b200_compute_stream": (C.c_void_p, [PM]),
    "mars_b200_last_gpu_ms": (C.c_float, [PM]),
    "mars_b200_set_profile": (None, [PM, C.c_int]),
    "mars_b200_num_ops": (C.c_int, [PM]),
    "mars_b200_op_info": (C.c_int, [PM, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mars_b200_geometry": (None, [PM, C.c_void_p]),
    "mars_b200_tensor_offset": (C.c_size_t, [PM, C.c_uint32]),
    "mars_yolo_parse_output": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]),
    "mars_yolo_nms": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "mars_b200_letterbox": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_letterbox_rgba": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_resize_taps": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mars_b200_plan_describe": (C.c_size_t, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]),
    "mars_b200_requant_fit": (C.c_int, [C.c_float, C.c_longlong, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "mars_b200_requant_ref": (C.c_int, [C.c_int, C.c_float]),
    "mars_b200_nmhwsoib2_size": (C.c_size_t, [C.c_int] * 4),
    "mars_b200_ndhwc32_size": (C.c_size_t, [C.c_int] * 4),
    "mars_b200_pack_weights_nmhwsoib2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_unpack_weights_nmhwsoib2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_nchw_to_ndhwc32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_ndhwc32_to_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mars_b200_pack_weights_nmhwsoib2_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mars_b200_unpack_weights_nmhwsoib2_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mars_b200_nchw_to_ndhwc32_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mars_b200_ndhwc32_to_nchw_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mars_yolo_nms_boxes": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "mars_yolo_scale_detections": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mars_yolo_decode_anchor_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int]),
}

_lib = None


def lib():
    """Load libmars_b200.so and declare every entry point.  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MarsLibraryMissing(LIB_PATH + " is missing: run `make -C thingino-accel_b200/csrc` "
                                     "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError = header/library drift
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _addr(buf):
    """numpy array / int address / object with data_ptr() -> raw address"""
    if buf is None:
        return None
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    raise TypeError(type(buf))


def _numel(desc):
    n = 1
    for i in range(min(desc.ndims, 6)):
        n *= max(desc.shape[i], 0)
    return n


class MarsModel:
    """A loaded .mars model (mars_load_memory ... mars_free), plus the batch entry points."""

    def __init__(self, model, arena_bytes=0, device=None, batch=1):
        L = lib()
        if device is not None:
            if L.mars_b200_set_device(device) != 0:
                raise MarsError(-5, "mars_b200_set_device", L.mars_b200_last_error().decode())
        L.mars_b200_set_arena_bytes(arena_bytes)
        self.m = PM()
        if isinstance(model, (bytes, bytearray, memoryview)):
            self._blob = bytes(model)
            err = L.mars_load_memory(self._blob, len(self._blob), C.byref(self.m))
        else:
            err = L.mars_load_file(os.fsencode(model), C.byref(self.m))
        L.mars_b200_set_arena_bytes(0)
        if err != 0:
            raise MarsError(err, "mars_load", L.mars_b200_last_error().decode())
        g = (C.c_size_t * 5)()
        L.mars_b200_geometry(self.m, g)
        self.weights_size, self.buffer_size, self.num_buffers, self.slot_stride, self.arena_bytes = [int(v) for v in g]
        if batch != 1:
            self.set_batch(batch)

    # ---- reference-shaped accessors ------------------------------------------------
    @property
    def header(self):
        return self.m.contents.header

    def _check(self, err, what):
        if err != 0:
            raise MarsError(err, what, lib().mars_b200_last_error().decode())

    def input(self, i=0):
        p = lib().mars_get_input(self.m, i)
        return p.contents if p else None

    def output(self, i=0):
        p = lib().mars_get_output(self.m, i)
        return p.contents if p else None

    def tensor_view(self, rt, nbytes=None):
        n = rt.alloc_size if nbytes is None else nbytes
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(rt.vaddr))

    def set_input(self, data, i=0):
        t = self.input(i)
        raw = np.ascontiguousarray(data).view(np.uint8).ravel()
        assert raw.size <= t.alloc_size
        self.tensor_view(t)[: raw.size] = raw

    def output_bytes(self, i=0):
        o = self.output(i)
        es = {0: 4, 1: 4, 2: 2}.get(o.desc.dtype, 1)
        return self.tensor_view(o, _numel(o.desc) * es)

    def run(self):
        self._check(lib().mars_run(self.m), "mars_run")

    def mirror(self):
        """host mirror of the whole arena (ddr_base)"""
        return np.ctypeslib.as_array((C.c_uint8 * self.arena_bytes).from_address(self.m.contents.ddr_base))

    def tensor_offset(self, idx):
        return int(lib().mars_b200_tensor_offset(self.m, idx))

    # ---- B200 additions ------------------------------------------------------------
    def set_batch(self, n):
        self._check(lib().mars_b200_set_batch(self.m, n), "mars_b200_set_batch")

    def set_opt_level(self, lvl):
        lib().mars_b200_set_opt_level(self.m, lvl)

    def set_f32_mode(self, mode):
        lib().mars_b200_set_f32_mode(self.m, mode)

    def set_depthwise_mode(self, mode):
        lib().mars_b200_set_depthwise_mode(self.m, mode)

    def arena_upload(self, slot=0):
        self._check(lib().mars_b200_arena_upload(self.m, slot), "arena_upload")

    def arena_download(self, slot=0):
        n = min(self.arena_bytes, self.weights_size + self.num_buffers * self.buffer_size)
        out = np.zeros(n, dtype=np.uint8)
        self._check(lib().mars_b200_arena_download(self.m, slot, out.ctypes.data, n), "arena_download")
        return out

    def arena_clear(self):
        self._check(lib().mars_b200_arena_clear(self.m), "arena_clear")

    def run_layer(self, i):
        return lib().mars_b200_run_layer(self.m, i)

    def describe(self):
        n = lib().mars_b200_describe(self.m, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().mars_b200_describe(self.m, buf, n + 1)
        return buf.value.decode()

    @property
    def input_bytes(self):
        return int(lib().mars_b200_input_bytes(self.m))

    @property
    def out_bytes(self):
        return int(lib().mars_b200_output_bytes(self.m))

    def upload_inputs(self, first, n, host, stride=0):
        self._check(lib().mars_b200_upload_inputs(self.m, first, n, _addr(host), stride), "upload_inputs")

    def preprocess(self, first, frames):
        """frames: [n, h, w, 3] uint8 (host) -> letterboxed into input tensor 0 of slots [first, first+n) on the device
        (reference src/mars/mars_yolo_test.c:40-77); returns the GPU milliseconds of the copy + kernels"""
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        n, h, w, c = f.shape
        assert c == 3
        self._check(lib().mars_b200_preprocess_batch(self.m, first, n, f.ctypes.data, h * w * 3, w, h), "preprocess_batch")
        return float(lib().mars_b200_last_gpu_ms(self.m))

    def download_outputs(self, first, n, host=None, stride=0):
        if host is None:
            host = np.zeros((n, self.out_bytes), dtype=np.uint8)
        self._check(lib().mars_b200_download_outputs(self.m, first, n, _addr(host), stride), "download_outputs")
        return host

    def run_resident(self, first, n):
        self._check(lib().mars_b200_run_resident(self.m, first, n), "run_resident")
        return float(lib().mars_b200_last_gpu_ms(self.m))

    def detect_resident(self, first, n, thresh=0.45):
        self._check(lib().mars_b200_detect_resident(self.m, first, n, thresh), "detect_resident")
        return float(lib().mars_b200_last_gpu_ms(self.m))

    def step_resident(self, first, n, thresh=0.45, with_detect=True):
        self._check(lib().mars_b200_step_resident(self.m, first, n, thresh, 1 if with_detect else 0), "step_resident")
        return float(lib().mars_b200_last_gpu_ms(self.m))

    def enqueue_step_resident(self, first, n, thresh=0.45, with_detect=True):
        """step_resident without the host synchronisation: the work is on compute_stream() when this returns"""
        self._check(lib().mars_b200_enqueue_step_resident(self.m, first, n, thresh, 1 if with_detect else 0), "enqueue_step_resident")

    def compute_stream(self):
        """cudaStream_t handle (int) of the stream the model's kernels run on"""
        return int(lib().mars_b200_compute_stream(self.m) or 0)

    def download_detections(self, first, n, maxd=1000):
        dets = np.zeros((n, maxd), dtype=DET_DTYPE)
        counts = np.zeros(n, dtype=np.int32)
        self._check(lib().mars_b200_download_detections(self.m, first, n, dets.ctypes.data, counts.ctypes.data, maxd),
                    "download_detections")
        return dets, counts

    def detect_batch(self, n, inputs, in_stride, dets, counts, maxd=1000, thresh=0.45):
        self._check(lib().mars_b200_detect_batch(self.m, n, _addr(inputs), in_stride, _addr(dets), _addr(counts), maxd, thresh),
                    "detect_batch")

    def submit_batch(self, pool, n, inputs, in_stride, dets, counts, maxd=1000, thresh=0.45):
        """queue a batch on half `pool` of the slot pool; buffers must stay alive until wait_batch(pool)"""
        self._check(lib().mars_b200_submit_batch(self.m, pool, n, _addr(inputs), in_stride, _addr(dets), _addr(counts), maxd, thresh),
                    "submit_batch")

    def submit_run_batch(self, pool, n, inputs, in_stride, outputs, out_stride):
        """queue a batch without the YOLO post-process (raw output tensors read back); wait with wait_batch(pool)"""
        self._check(lib().mars_b200_submit_run_batch(self.m, pool, n, _addr(inputs), in_stride, _addr(outputs), out_stride), "submit_run_batch")

    def wait_batch(self, pool):
        self._check(lib().mars_b200_wait_batch(self.m, pool), "wait_batch")

    def run_batch(self, n, inputs, in_stride, outputs, out_stride):
        self._check(lib().mars_b200_run_batch(self.m, n, _addr(inputs), in_stride, _addr(outputs), out_stride), "run_batch")

    def detections_device(self):
        """(dets_ptr, counts_ptr, stride) of the device-resident detection records"""
        d, c, s = C.c_void_p(), C.c_void_p(), C.c_int()
        lib().mars_b200_detections_device(self.m, C.byref(d), C.byref(c), C.byref(s))
        return d.value, c.value, s.value

    @property
    def last_gpu_ms(self):
        return float(lib().mars_b200_last_gpu_ms(self.m))

    @property
    def launch_count(self):
        return int(lib().mars_b200_launch_count(self.m))

    def set_profile(self, on):
        lib().mars_b200_set_profile(self.m, 1 if on else 0)

    def op_profile(self):
        """list of dicts, one per device op, with accumulated CUDA-event milliseconds"""
        out = []
        info = (C.c_int32 * 16)()
        ms, calls, n = C.c_double(), C.c_uint64(), C.c_uint64()
        for i in range(lib().mars_b200_num_ops(self.m)):
            lib().mars_b200_op_info(self.m, i, info, C.byref(ms), C.byref(calls), C.byref(n))
            out.append(dict(op=i, kind=info[0], layer=info[1], impl=info[2], mode=info[3], ic=info[4], oc=info[5], oh=info[6],
                            ow=info[7], kh=info[8], kw=info[9], fused=info[10], ih=info[11], iw=info[12], ms=ms.value,
                            calls=calls.value, n=n.value))
        return out

    def close(self):
        if self.m:
            lib().mars_free(self.m)
            self.m = PM()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- pre-process on host arrays ------------------------------------------------------
def letterbox(rgb, tw, th, nhwc=False):
    """[h, w, 3] uint8 frame -> tw*th*3 int8 network input (runs on the device; reference load_image minus the decode)"""
    f = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = f.shape
    out = np.zeros(tw * th * 3, dtype=np.int8)
    if lib().mars_b200_letterbox(f.ctypes.data, w, h, tw, th, 1 if nhwc else 0, out.ctypes.data) != 0:
        raise RuntimeError("mars_b200_letterbox failed: " + lib().mars_b200_last_error().decode())
    return out


def letterbox_rgba(rgb, tw=640, th=640):
    """[h, w, 3] uint8 frame -> tw*th*4 uint8 RGBA frame (reference examples/yolo_detect.cpp:72-130 minus the decode)"""
    f = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = f.shape
    out = np.zeros(tw * th * 4, dtype=np.uint8)
    if lib().mars_b200_letterbox_rgba(f.ctypes.data, w, h, tw, th, out.ctypes.data) != 0:
        raise RuntimeError("mars_b200_letterbox_rgba failed: " + lib().mars_b200_last_error().decode())
    return out


def resize_taps(in_size, out_size):
    """host half of the pre-processing (no GPU): (start[out+1], src[], w[]) tap lists of one axis"""
    start = np.zeros(out_size + 1, dtype=np.int32)
    cap = 16 * max(in_size, out_size) + 64
    while True:
        src, w = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.float32)
        n = lib().mars_b200_resize_taps(in_size, out_size, start.ctypes.data, src.ctypes.data, w.ctypes.data, cap)
        if n < 0:
            raise ValueError("resize_taps(%d, %d)" % (in_size, out_size))
        if n <= cap:
            return start, src[:n].copy(), w[:n].copy()
        cap = n


# ---- post-process on host arrays (runs on the device) ---------------------------------
def parse_output(out_i8, npred, scale, maxd=1000):
    dets = np.zeros(max(maxd, 1), dtype=DET_DTYPE)
    buf = np.ascontiguousarray(out_i8).view(np.int8)
    n = lib().mars_yolo_parse_output(buf.ctypes.data, npred, scale, dets.ctypes.data, maxd)
    return dets[:n].copy()


def nms(dets, thresh=0.45):
    d = np.ascontiguousarray(dets.copy())
    n = lib().mars_yolo_nms(d.ctypes.data, len(d), thresh)
    return d[:max(n, 0)].copy()


def nms_boxes(boxes, thresh=0.45):
    d = np.ascontiguousarray(boxes.copy())
    n = lib().mars_yolo_nms_boxes(d.ctypes.data, len(d), thresh)
    return d[:max(n, 0)].copy()
