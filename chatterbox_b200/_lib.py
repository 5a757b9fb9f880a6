"""ctypes binding of libcbx.so (include/cbx.h).  Fails loudly when the CUDA library is missing:
there is no CPU / PyTorch fallback for the hot path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CBX_LIB") or os.path.join(_HERE, "libcbx.so")   # CBX_LIB: kernel A/B experiments


class CbxError(RuntimeError):
    pass


class Layout(C.Structure):
    _fields_ = [("n_seq", C.c_int), ("rows", C.c_int), ("max_len", C.c_int),
                ("tile_seq", C.c_void_p), ("start", C.c_void_p), ("len", C.c_void_p),
                ("h_start", C.c_void_p), ("h_len", C.c_void_p)]


class T3State(C.Structure):
    _fields_ = [("n_utts", C.c_int), ("n_rows", C.c_int), ("cfg", C.c_int),
                ("kv_pages", C.c_void_p), ("kv_dtype", C.c_int), ("page_tokens", C.c_int),
                ("page_table", C.c_void_p), ("max_pages_per_row", C.c_int), ("n_pages", C.c_int),
                ("positions", C.c_void_p), ("base_pos", C.c_void_p),
                ("tokens", C.c_void_p), ("max_tokens", C.c_int),
                ("n_gen", C.c_void_p), ("max_new", C.c_void_p), ("done", C.c_void_p), ("seen", C.c_void_p),
                ("x", C.c_void_p), ("logits", C.c_void_p), ("ldl", C.c_int),
                ("cfg_weight", C.c_float), ("rep_penalty", C.c_float), ("temperature", C.c_float),
                ("min_p", C.c_float), ("top_p", C.c_float),
                ("q_noise", C.c_void_p), ("seed", C.c_ulonglong), ("sampler", C.c_int), ("top_k", C.c_int),
                ("act_utt", C.c_void_p), ("n_act", C.c_void_p), ("src_slot", C.c_void_p), ("slot_row", C.c_void_p),
                ("m_live", C.c_void_p), ("force_tokens", C.c_void_p), ("sampled_out", C.c_void_p), ("act_fp16", C.c_int)]


class HiftGeom(C.Structure):
    _fields_ = [("LT", Layout), ("L8", Layout), ("L40", Layout), ("L120", Layout),
                ("sample_start", C.c_void_p), ("total_samples", C.c_longlong)]


# every symbol include/cbx.h declares: (restype, argtypes)
_P, _I, _F, _S = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SYMBOLS = {
    "cbx_create": (_I, [_I, C.POINTER(_P)]),
    "cbx_destroy": (None, [_P]),
    "cbx_last_error": (C.c_char_p, [_P]),
    "cbx_version": (_I, []),
    "cbx_set_option": (_I, [_P, C.c_char_p, C.c_char_p]),
    "cbx_launch_count": (C.c_longlong, [_P]),
    "cbx_timer_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "cbx_timer_read_bytes": (_I, [_P, C.POINTER(C.c_double)]),
    "cbx_timer_read_class": (_I, [_P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cbx_load_tensor": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(C.c_int64)]),
    "cbx_finalize_weights": (_I, [_P, C.c_char_p]),
    "cbx_t3_cond_encode": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _S, _P]),
    "cbx_t3_prefill": (_I, [_P, C.POINTER(T3State), _I, _P, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _S, _P]),
    "cbx_t3_decode": (_I, [_P, C.POINTER(T3State), _I, _I, _P, _S, _P]),
    "cbx_t3_workspace_bytes": (_S, [_P, _I, _I]),
    "cbx_flow_encode": (_I, [_P, _P, C.POINTER(Layout), C.POINTER(Layout), _P, _P, _P, _P, _S, _P]),
    "cbx_cfm_solve": (_I, [_P, _P, _P, _P, _P, C.POINTER(Layout), C.POINTER(Layout), _I, _F, _I, _P, _S, _P]),
    "cbx_flow_workspace_bytes": (_S, [_P, C.POINTER(Layout), C.POINTER(Layout), C.POINTER(Layout)]),
    "cbx_hift_source": (_I, [_P, _P, C.POINTER(HiftGeom), _P, _P, C.c_ulonglong, _P, _P, _P, _P, _S, _P]),
    "cbx_hift_decode": (_I, [_P, _P, _P, C.POINTER(HiftGeom), _P, _I, _P, _S, _P]),
    "cbx_hift_workspace_bytes": (_S, [_P, C.POINTER(HiftGeom)]),
    "cbx_test_gemm": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, C.POINTER(Layout), C.POINTER(Layout),
                           _I, _F, _P, _I, _I, _P, _I, _P]),
    "cbx_test_attention": (_I, [_P, _P, _P, _P, _I, _P, _I, _I, C.POINTER(Layout), _F, _I, _P, C.c_longlong, _I, _I, _I, _P]),
    "cbx_test_attention_tc": (_I, [_P, _P, _P, _I, C.POINTER(Layout), _F, _P, _S, _P]),
    "cbx_test_paged_decode": (_I, [_P, _P, _P, _I, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _S, _P]),
    "cbx_test_gemm_splitk": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _S, _P]),
    "cbx_bench_gemm_f16": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _S, _P]),
    "cbx_test_umma_rowshift": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "cbx_test_gemm_f16": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _S, _P]),
}

_lib = None


def load():
    """dlopen libcbx.so and type every declared symbol (raises CbxError if the library was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CbxError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(make -C chatterbox_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError -> missing export
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Handle:
    """RAII wrapper of cbx_handle; `call` turns non-zero status codes into CbxError."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.cbx_create(int(device), C.byref(h))
        if st != 0 or not h.value:
            msg = self.lib.cbx_last_error(h).decode() if h.value else "no CUDA device"
            raise CbxError(f"cbx_create failed ({st}): {msg} -- the engine needs a B200 (sm_100a); no CPU fallback")
        self.h = h

    def call(self, name, *args):
        st = getattr(self.lib, name)(self.h, *args)
        if st != 0:
            raise CbxError(f"{name} failed ({st}): {self.lib.cbx_last_error(self.h).decode()}")

    def set_option(self, key, value):
        self.call("cbx_set_option", key.encode(), value.encode())

    def launch_count(self):
        return int(self.lib.cbx_launch_count(self.h))

    def timer_read(self):
        ms, n, w = C.c_double(0), C.c_longlong(0), C.c_double(0)
        self.call("cbx_timer_read", C.byref(ms), C.byref(n), C.byref(w))
        return float(ms.value), int(n.value), float(w.value)

    def timer_read_class(self, cls):
        """-> dict(ms, n, work, bytes) of one kernel class after set_option('time_kernel', 'all' | cls)."""
        ms, n, w, b = C.c_double(0), C.c_longlong(0), C.c_double(0), C.c_double(0)
        self.call("cbx_timer_read_class", cls.encode(), C.byref(ms), C.byref(n), C.byref(w), C.byref(b))
        return dict(ms=float(ms.value), n=int(n.value), work=float(w.value), bytes=float(b.value))

    def timer_read_bytes(self):
        b = C.c_double(0)
        self.call("cbx_timer_read_bytes", C.byref(b))
        return float(b.value)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.cbx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
