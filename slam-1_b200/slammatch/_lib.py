"""ctypes binding of libslammatch.so (the C-ABI in include/slammatch.h).

The library is the product; this module only declares prototypes.  There is no CPU fallback:
if the shared object cannot be loaded, or no B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libslammatch.so")

SLM_OK = 0
VARIANTS = {"auto": 0, "popc": 1, "tensor": 2, "bmma": 3, "tensor4": 4}

# every symbol include/slammatch.h declares (tests check the library exports all of them)
SYMBOLS = (
    "slm_last_error", "slm_version", "slm_create", "slm_destroy", "slm_set_variant", "slm_last_variant",
    "slm_launch_count", "slm_last_kernel",
    "slm_profile_enable", "slm_profile_read",
    "slm_knn2", "slm_knn2_keys", "slm_knn2_filter", "slm_knn2_masked", "slm_knn2_batched", "slm_merge_top2",
    "slm_exchange_merge", "slm_knn2_exchange", "slm_exchange_status",
    "slm_compact_matches", "slm_gather_rows", "slm_filter_points3d", "slm_bow_hist", "slm_chi2_scan", "slm_vocab_update",
    "slm_knn2_host", "slm_probe_tensor_peak", "slm_probe_popc_peak",
)

_lib = None
_lock = threading.Lock()


class SlamMatchError(RuntimeError):
    """A libslammatch call returned a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libslammatch error {code}: {msg}")
        self.code = code


def load():
    """Load libslammatch.so (building it first if only the sources are present)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # fail loudly: no fallback path exists
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}. Build it with `python -m slammatch.build`.") from e
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        lib.slm_last_error.restype = ctypes.c_char_p
        lib.slm_last_error.argtypes = []
        lib.slm_version.restype = ctypes.c_int
        lib.slm_version.argtypes = []
        lib.slm_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        lib.slm_destroy.argtypes = [vp]
        lib.slm_set_variant.argtypes = [vp, ctypes.c_int]
        lib.slm_last_variant.argtypes = [vp]
        lib.slm_launch_count.argtypes = [vp]
        lib.slm_last_kernel.argtypes = [vp]
        lib.slm_launch_count.restype = i64
        lib.slm_profile_enable.argtypes = [vp, ctypes.c_int]
        lib.slm_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
        lib.slm_knn2.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp, vp]
        lib.slm_knn2_keys.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp]
        lib.slm_knn2_filter.argtypes = [vp, vp, i64, vp, i64, i64, i32, i32, i32, vp, vp, vp, vp]
        lib.slm_knn2_masked.argtypes = [vp, vp, i64, vp, i64, i64, vp, i64, i32, i32, vp, vp, vp, vp]
        lib.slm_knn2_batched.argtypes = [vp, vp, i64, i64, vp, i64, i32, i32, vp, vp, vp, vp]
        lib.slm_merge_top2.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, vp, vp]
        lib.slm_exchange_merge.argtypes = [vp, vp, i64, i64, i64, vp, vp, i32, i32, ctypes.c_uint32, i32, i32, vp, vp, vp, vp]
        lib.slm_knn2_exchange.argtypes = [vp, vp, i64, vp, i64, i64, i64, i64, vp, vp, i32, i32, ctypes.c_uint32, i32, i32,
                                          vp, vp, vp, vp]
        lib.slm_exchange_status.argtypes = [vp]
        lib.slm_filter_points3d.argtypes = [vp, vp, i64, ctypes.c_double, vp, vp]
        lib.slm_probe_tensor_peak.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        lib.slm_probe_popc_peak.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        lib.slm_compact_matches.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
        lib.slm_gather_rows.argtypes = [vp, vp, i32, vp, vp, i64, i32, vp, vp]
        lib.slm_bow_hist.argtypes = [vp, vp, i64, i32, i32, vp, vp]
        lib.slm_chi2_scan.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp]
        lib.slm_vocab_update.argtypes = [vp, vp, i64, vp, i32, vp, i32, vp, vp, vp]
        lib.slm_knn2_host.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, vp, vp, vp]
        for name in SYMBOLS:
            if name not in ("slm_last_error", "slm_launch_count", "slm_last_kernel"):
                getattr(lib, name).restype = ctypes.c_int
        lib.slm_last_kernel.restype = ctypes.c_char_p
        _lib = lib
        return lib


def check(status: int):
    if status != SLM_OK:
        raise SlamMatchError(status, load().slm_last_error().decode("utf-8", "replace"))


class Context:
    """Owns one slm_ctx (replaces the matcher object built at tracking.py:17).  One per (thread, device)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = ctypes.c_void_p()
        check(self.lib.slm_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)
        self.variant = 0

    def set_variant(self, variant):
        v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        check(self.lib.slm_set_variant(self.handle, v))
        self.variant = v

    def using(self, variant):
        """Context manager: run with ``variant`` (None = leave as is) and restore the previous setting afterwards, so a
        per-call ``variant=`` argument never leaks into later calls on the cached per-thread ctx."""
        return _VariantScope(self, variant)

    def probe_tensor_peak(self, kind: str = "f8f6f4", loops: int = 2048, reps: int = 5):
        """(dense TFLOP/s, MAC/clk/SM) of back-to-back tcgen05.mma on every SM, measured now, in this process."""
        tf, mac = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(self.lib.slm_probe_tensor_peak(self.handle, {"f8f6f4": 0, "mxf4": 1}[kind], int(loops), int(reps),
                                             ctypes.byref(tf), ctypes.byref(mac)))
        return tf.value, mac.value

    def probe_popc_peak(self, reps: int = 5):
        """(Tcmp/s, POPC32 lanes/clk/SM) of the integer-pipe comparison loop on every SM, measured now."""
        tc, lanes = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(self.lib.slm_probe_popc_peak(self.handle, int(reps), ctypes.byref(tc), ctypes.byref(lanes)))
        return tc.value, lanes.value

    def last_variant(self) -> str:
        v = int(self.lib.slm_last_variant(self.handle))
        return {n: k for k, n in VARIANTS.items()}.get(v, "auto")

    def last_kernel(self) -> str:
        """Name of the distance kernel the last search launched."""
        return (self.lib.slm_last_kernel(self.handle) or b"").decode()

    def launch_count(self) -> int:
        return int(self.lib.slm_launch_count(self.handle))

    def profile(self, enable: bool):
        check(self.lib.slm_profile_enable(self.handle, int(bool(enable))))

    def profile_read(self):
        """(summed duration in ms of the dominant kernel, number of launches) since the last read."""
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        check(self.lib.slm_profile_read(self.handle, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.slm_destroy(self.handle)
            self.handle = None

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass


class _VariantScope:
    def __init__(self, ctx, variant):
        self.ctx, self.variant, self.prev = ctx, variant, None

    def __enter__(self):
        if self.variant is not None:
            self.prev = self.ctx.variant
            self.ctx.set_variant(self.variant)
        return self.ctx

    def __exit__(self, *exc):
        if self.prev is not None:
            self.ctx.set_variant(self.prev)
        return False


_ctx_cache: dict = {}


def context(device: int = 0) -> Context:
    """Cached ctx per (thread, device): the reference re-constructs its matcher on every call
    (keypoint.py:43, Point3D.py:39), so construction must be free."""
    key = (threading.get_ident(), int(device))
    ctx = _ctx_cache.get(key)
    if ctx is None or ctx.handle is None:
        ctx = Context(device)
        _ctx_cache[key] = ctx
    return ctx
