"""ctypes binding of libnqmc_b200.so (include/nqmc_b200.h), device buffers, and the device-side Context.

No PyTorch, no JAX: device memory, pinned host memory and streams come from the library itself
(``nqmc_dev_alloc`` ...), wrapped by ``DeviceArray``.  Anything that exposes ``data_ptr()`` (a torch CUDA tensor in
the tests, an XLA buffer behind the FFI shim) is accepted wherever a device buffer is expected.  Every computation
on walkers goes through the C ABI.  There is no CPU fallback: if the shared library or a CUDA device is missing,
construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np

_LIB_NAME = "libnqmc_b200.so"
_lib = None

KIND_SIMPLE = 0
KIND_ANTISYMMETRIC = 1
ACTIVATIONS = {"tanh": 0, "silu": 1}
HDR_N, HDR_SUM_E, HDR_SUM_E2, HDR_N_ACCEPT, HDR_N_BAD, HDR_SIZE = 0, 1, 2, 3, 4, 8

EXPORTS = [
    "nqmc_create", "nqmc_destroy", "nqmc_last_error", "nqmc_abi_version", "nqmc_reserve", "nqmc_param_count",
    "nqmc_leaf_info", "nqmc_set_params", "nqmc_set_params_host", "nqmc_aos_to_soa", "nqmc_soa_to_aos",
    "nqmc_log_psi", "nqmc_mh_step", "nqmc_local_energy", "nqmc_grad_accum", "nqmc_energy_stats",
    "nqmc_walker_step_a", "nqmc_walker_step_b", "nqmc_walker_step", "nqmc_walker_step_host",
    "nqmc_walker_step_host_async", "nqmc_walker_step_host_wait", "nqmc_walker_step_host_upload",
    "nqmc_launch_count", "nqmc_rng_fill", "nqmc_profile", "nqmc_profile_read", "nqmc_ffma_peak", "nqmc_use_tensor_cores", "nqmc_get_params", "nqmc_get_params_host", "nqmc_adam_step",
    "nqmc_comm_create", "nqmc_comm_attach", "nqmc_comm_allreduce", "nqmc_comm_active", "nqmc_sample", "nqmc_blocked_stats",
    "nqmc_device_count", "nqmc_dev_alloc", "nqmc_dev_free", "nqmc_host_alloc", "nqmc_host_free", "nqmc_dev_memcpy",
    "nqmc_dev_copy_rows", "nqmc_dev_memset", "nqmc_stream_create", "nqmc_stream_destroy", "nqmc_stream_sync", "nqmc_stream_wait",
    "nqmc_mem_last_error", "nqmc_xla_ffi_available", "nqmc_gram", "nqmc_sr_workspace_bytes", "nqmc_sr_update", "nqmc_log_psi_grads", "nqmc_apply_update", "nqmc_debug_check_guards", "nqmc_debug_scratch_ptr",
    "nqmc_debug_item_iter",
]


class NativeError(RuntimeError):
    pass


class _System(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("n_elec", C.c_int32), ("n_up", C.c_int32), ("_pad", C.c_int32),
                ("R", C.POINTER(C.c_double)), ("Z", C.POINTER(C.c_double)), ("v_nn", C.c_double)]


class _Net(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hidden0", C.c_int32), ("hidden1", C.c_int32), ("use_cusp", C.c_int32),
                ("use_backflow", C.c_int32), ("_pad", C.c_int32), ("envelope_decay", C.c_double),
                ("cusp_same", C.c_double), ("cusp_diff", C.c_double),
                ("n_hidden", C.c_int32), ("activation", C.c_int32), ("hidden", C.c_int32 * 6)]


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """Load libnqmc_b200.so; raises NativeError (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(f"{path} not found -- build it with `python __graft_entry__.py` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i64, u64, f32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_float
    lib.nqmc_create.argtypes = [C.POINTER(_System), C.POINTER(_Net), C.c_int, C.POINTER(vp)]
    lib.nqmc_destroy.argtypes = [vp]
    lib.nqmc_last_error.argtypes = [vp]
    lib.nqmc_last_error.restype = C.c_char_p
    lib.nqmc_reserve.argtypes = [vp, i64]
    lib.nqmc_param_count.argtypes = [vp]
    lib.nqmc_param_count.restype = i64
    lib.nqmc_leaf_info.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.nqmc_leaf_info.restype = i64
    lib.nqmc_set_params.argtypes = [vp, vp, i64, vp]
    lib.nqmc_set_params_host.argtypes = [vp, vp, i64, vp]
    lib.nqmc_aos_to_soa.argtypes = [vp, vp, vp, i64, vp]
    lib.nqmc_soa_to_aos.argtypes = [vp, vp, vp, i64, vp]
    lib.nqmc_log_psi.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.nqmc_mh_step.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, vp, vp, vp]
    lib.nqmc_local_energy.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.nqmc_grad_accum.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.nqmc_energy_stats.argtypes = [vp, vp, i64, vp, vp]
    lib.nqmc_walker_step_a.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, vp, vp, vp, vp]
    lib.nqmc_walker_step_b.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    lib.nqmc_walker_step.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, vp, vp, vp, vp, vp]
    lib.nqmc_walker_step_host.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, vp, vp, vp, vp]
    lib.nqmc_walker_step_host_async.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, vp, vp, vp, vp]
    lib.nqmc_walker_step_host_wait.argtypes = [vp, C.c_int]
    lib.nqmc_walker_step_host_upload.argtypes = [vp, vp, vp, vp, vp, i64]
    lib.nqmc_launch_count.argtypes = [vp]
    lib.nqmc_launch_count.restype = i64
    lib.nqmc_rng_fill.argtypes = [vp, u64, u64, i64, i64, vp, vp, vp]
    lib.nqmc_profile.argtypes = [vp, C.c_int]
    lib.nqmc_get_params.argtypes = [vp, vp, i64, vp]
    lib.nqmc_get_params_host.argtypes = [vp, vp, i64, vp]
    lib.nqmc_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, f32, f32, vp, vp]
    lib.nqmc_use_tensor_cores.argtypes = [vp, C.c_int]
    lib.nqmc_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.nqmc_ffma_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), vp]
    lib.nqmc_comm_create.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.nqmc_comm_attach.argtypes = [vp, vp]
    lib.nqmc_comm_allreduce.argtypes = [vp, vp, i64, vp]
    lib.nqmc_comm_active.argtypes = [vp]
    lib.nqmc_sample.argtypes = [vp, vp, vp, vp, vp, u64, u64, i64, f32, i64, i64, i64, i64, vp, vp, vp, f32, i64, vp]
    lib.nqmc_blocked_stats.argtypes = [vp, vp, i64, i64, C.c_int32, vp, vp]
    sz = C.c_size_t
    lib.nqmc_gram.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, f32, f32, C.c_int, vp, i64, vp]
    lib.nqmc_sr_workspace_bytes.argtypes = [i64]
    lib.nqmc_sr_workspace_bytes.restype = i64
    lib.nqmc_sr_update.argtypes = [vp, vp, vp, i64, i64, i64, f32, f32, C.c_int32, C.c_double, vp, vp, vp, vp, vp]
    lib.nqmc_log_psi_grads.argtypes = [vp, vp, i64, vp, i64, vp]
    lib.nqmc_apply_update.argtypes = [vp, vp, i64, vp]
    lib.nqmc_debug_check_guards.argtypes = [vp]
    lib.nqmc_debug_scratch_ptr.argtypes = [vp, C.c_char_p]
    lib.nqmc_debug_scratch_ptr.restype = vp
    lib.nqmc_debug_item_iter.argtypes = [C.c_int32, C.c_int32, C.c_int32, i64, vp, vp]
    lib.nqmc_device_count.argtypes = []
    lib.nqmc_dev_alloc.argtypes = [C.c_int, sz, C.POINTER(vp)]
    lib.nqmc_dev_free.argtypes = [C.c_int, vp]
    lib.nqmc_host_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.nqmc_host_free.argtypes = [vp]
    lib.nqmc_dev_memcpy.argtypes = [C.c_int, vp, vp, sz, C.c_int, vp]
    lib.nqmc_dev_copy_rows.argtypes = [C.c_int, vp, sz, vp, sz, sz, sz, vp]
    lib.nqmc_dev_memset.argtypes = [C.c_int, vp, C.c_int, sz, vp]
    lib.nqmc_stream_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.nqmc_stream_destroy.argtypes = [C.c_int, vp]
    lib.nqmc_stream_wait.argtypes = [C.c_int, vp, vp]
    lib.nqmc_stream_sync.argtypes = [C.c_int, vp]
    lib.nqmc_mem_last_error.argtypes = []
    lib.nqmc_mem_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _mem_check(rc: int, what: str):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {load_library().nqmc_mem_last_error().decode()}")


def device_count() -> int:
    return int(load_library().nqmc_device_count())


def default_device() -> int:
    """The device a new Context / DeviceArray lands on when none is given: NQMC_DEVICE, else LOCAL_RANK (one
    process per GPU under torchrun / mpirun), else the current torch device if the CALLER has already imported
    torch and initialised CUDA (never imported from here), else 0."""
    for var in ("NQMC_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    t = sys.modules.get("torch")
    if t is not None:
        try:
            if t.cuda.is_initialized():
                return int(t.cuda.current_device())
        except Exception:
            pass
    return 0


# Free device blocks kept for reuse, keyed by (device, nbytes).  cudaMalloc / cudaFree cost ~0.1 ms each and cudaFree
# synchronises the whole device, which would serialise the walker loop on every temporary.  Reuse is safe for work
# issued on ONE stream per device (every Context call is, in order); callers that spread one device over several
# streams should keep their own buffers alive instead of relying on temporaries.
_POOL: Dict[Tuple[int, int], List[int]] = {}
_POOL_BYTES = [0]
_POOL_CAP = 8 << 30


def empty_cache() -> None:
    """Return every pooled block to the driver."""
    lib = load_library()
    for (dev, _), ptrs in _POOL.items():
        for p in ptrs:
            lib.nqmc_dev_free(dev, p)
    _POOL.clear()
    _POOL_BYTES[0] = 0


class _HostView(np.ndarray):
    """ndarray that also answers ``.numpy()`` / ``.cpu()`` so results read the same whether a caller holds a
    DeviceArray or (in the tests) a torch tensor."""

    def numpy(self):
        return np.asarray(self)

    def cpu(self):
        return self

    def item(self, *a):
        return np.asarray(self).item(*a)


class DeviceArray:
    """A caller-owned buffer in device memory (nqmc_dev_alloc).  Row-major, fixed dtype/shape; views share the
    allocation.  ``numpy()`` copies back (synchronising the stream); ``data_ptr()`` is what the C ABI takes."""

    __slots__ = ("ptr", "shape", "dtype", "device", "_base", "_owns", "__weakref__")

    def __init__(self, shape, dtype=np.float32, device: Optional[int] = None, _ptr: Optional[int] = None, _base=None):
        self.shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        self.device = default_device() if device is None else int(device)
        self._base = _base
        if _ptr is None:
            free = _POOL.get((self.device, self.nbytes))
            if free:
                self.ptr, self._owns = free.pop(), True
                _POOL_BYTES[0] -= self.nbytes
            else:
                p = C.c_void_p()
                _mem_check(load_library().nqmc_dev_alloc(self.device, self.nbytes, C.byref(p)), "nqmc_dev_alloc")
                self.ptr, self._owns = int(p.value), True
        else:
            self.ptr, self._owns = int(_ptr), False

    def __del__(self):
        try:
            if self._owns and self.ptr and _lib is not None:
                if _POOL_BYTES[0] + self.nbytes <= _POOL_CAP:
                    _POOL.setdefault((self.device, self.nbytes), []).append(self.ptr)
                    _POOL_BYTES[0] += self.nbytes
                else:
                    _lib.nqmc_dev_free(self.device, self.ptr)
                self.ptr = 0
        except Exception:
            pass

    # -- geometry
    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def numel(self) -> int:
        return self.size

    def data_ptr(self) -> int:
        return self.ptr

    def __len__(self):
        return self.shape[0]

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def view(self, *shape) -> "DeviceArray":
        shape = shape[0] if len(shape) == 1 and not np.isscalar(shape[0]) else shape
        shape = list(shape)
        if -1 in shape:
            k = shape.index(-1)
            shape[k] = self.size // max(1, -int(np.prod(shape)))
        if int(np.prod(shape)) != self.size:
            raise ValueError(f"cannot view {self.shape} as {tuple(shape)}")
        return DeviceArray(shape, self.dtype, self.device, _ptr=self.ptr, _base=self)

    reshape = view

    def row(self, i: int) -> "DeviceArray":
        """View of ``self[i]`` (leading index)."""
        if not 0 <= i < self.shape[0]:
            raise IndexError(i)
        inner = self.shape[1:]
        step = (int(np.prod(inner)) if inner else 1) * self.dtype.itemsize
        return DeviceArray(inner, self.dtype, self.device, _ptr=self.ptr + i * step, _base=self)

    # -- transfers
    def copy_from_host(self, a, stream: int = 0) -> "DeviceArray":
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.size != self.size:
            raise ValueError(f"host array has {a.size} elements, device buffer {self.size}")
        lib = load_library()
        _mem_check(lib.nqmc_dev_memcpy(self.device, self.ptr, a.ctypes.data, self.nbytes, 1, stream), "nqmc_dev_memcpy")
        _mem_check(lib.nqmc_stream_sync(self.device, stream), "nqmc_stream_sync")      # `a` may be a temporary
        return self

    def copy_from(self, other, stream: int = 0) -> "DeviceArray":
        """device -> device from anything with data_ptr() holding at least nbytes."""
        _mem_check(load_library().nqmc_dev_memcpy(self.device, self.ptr, _ptr(other), self.nbytes, 3, stream),
                   "nqmc_dev_memcpy")
        return self

    def zero_(self, stream: int = 0) -> "DeviceArray":
        _mem_check(load_library().nqmc_dev_memset(self.device, self.ptr, 0, self.nbytes, stream), "nqmc_dev_memset")
        return self

    def numpy(self, stream: int = 0) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        lib = load_library()
        _mem_check(lib.nqmc_dev_memcpy(self.device, out.ctypes.data, self.ptr, self.nbytes, 2, stream), "nqmc_dev_memcpy")
        _mem_check(lib.nqmc_stream_sync(self.device, stream), "nqmc_stream_sync")
        return out

    def cpu(self) -> _HostView:
        return self.numpy().view(_HostView)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def clone(self, stream: int = 0) -> "DeviceArray":
        return DeviceArray(self.shape, self.dtype, self.device).copy_from(self, stream)

    def contiguous(self):
        return self

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.device}, ptr=0x{self.ptr:x})"


def to_device(a, dtype=np.float32, device: Optional[int] = None, stream: int = 0):
    """numpy -> new DeviceArray; device buffers (anything with data_ptr()) pass through untouched."""
    if a is None or hasattr(a, "data_ptr"):
        return a
    a = np.ascontiguousarray(a, dtype=dtype)
    return DeviceArray(a.shape, dtype, device).copy_from_host(a, stream)


def _ptr(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _shape(t):
    return tuple(t.shape)


class Context:
    """One (geometry, network shape) on one GPU.  Thin, stateful wrapper over nqmc_ctx.
    ``stream``: a cudaStream_t handle (int) every call is issued on; 0 = the legacy default stream, which is also
    what torch uses unless told otherwise, so tests that mix torch tensors with this class stay ordered."""

    def __init__(self, R: np.ndarray, Z: np.ndarray, n_up: int, n_dn: int, v_nn: float, kind: int,
                 hidden: Tuple[int, ...], use_cusp: bool = False, use_backflow: bool = False,
                 envelope_decay: float = 1.0, cusp_same: float = 0.5, cusp_diff: float = 0.25,
                 device: Optional[int] = None, activation: str = "tanh", stream: int = 0):
        self.lib = load_library()
        if device_count() <= 0:
            raise NativeError("no CUDA device is visible; the nqmc_b200 walker loop has no CPU fallback")
        self.device = default_device() if device is None else int(device)
        self.stream = int(stream)
        self._R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
        self._Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1)
        self.n_atoms, self.n_up, self.n_dn = self._R.shape[0], int(n_up), int(n_dn)
        self.n_elec = self.n_up + self.n_dn
        self.kind = kind
        sysm = _System(self.n_atoms, self.n_elec, self.n_up, 0,
                       self._R.ctypes.data_as(C.POINTER(C.c_double)), self._Z.ctypes.data_as(C.POINTER(C.c_double)),
                       float(v_nn))
        if activation not in ACTIVATIONS:
            raise ValueError(f"Unknown activation: {activation}")          # simple.py:60
        hidden = tuple(int(h) for h in hidden)
        if kind == KIND_SIMPLE:
            if not 1 <= len(hidden) <= 6:
                raise NativeError("SimpleWavefunction: 1..6 hidden layers")
            net = _Net(kind, hidden[0], hidden[1] if len(hidden) > 1 else 0, 0, 0, 0,
                       float(envelope_decay), float(cusp_same), float(cusp_diff), len(hidden), ACTIVATIONS[activation],
                       (C.c_int32 * 6)(*hidden))
        else:
            net = _Net(kind, hidden[0], hidden[1], int(bool(use_cusp)), int(bool(use_backflow)), 0,
                       float(envelope_decay), float(cusp_same), float(cusp_diff), 0, 0, (C.c_int32 * 6)())
        h = C.c_void_p()
        rc = self.lib.nqmc_create(C.byref(sysm), C.byref(net), self.device, C.byref(h))
        if rc != 0:
            raise NativeError(f"nqmc_create failed ({rc}): {self.lib.nqmc_last_error(None).decode()}")
        self.h = h
        self.P = int(self.lib.nqmc_param_count(self.h))
        self.cap = 0
        self._theta_host = None      # host copy of the resident theta (set_params skips identical uploads)

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.nqmc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise NativeError(f"{what} failed ({rc}): {self.lib.nqmc_last_error(self.h).decode()}")

    def dev(self):
        """torch.device of this context -- a convenience for callers that use torch (the tests); imports torch."""
        import torch
        return torch.device("cuda", self.device)

    def empty(self, shape, dtype=np.float32) -> DeviceArray:
        return DeviceArray(shape, dtype, self.device)

    def zeros(self, shape, dtype=np.float32) -> DeviceArray:
        return DeviceArray(shape, dtype, self.device).zero_(self.stream)

    def upload(self, a, dtype=np.float32):
        return to_device(a, dtype, self.device, self.stream)

    def synchronize(self):
        _mem_check(self.lib.nqmc_stream_sync(self.device, self.stream), "nqmc_stream_sync")

    def reserve(self, W: int):
        if W > self.cap:
            self._check(self.lib.nqmc_reserve(self.h, int(W)), "nqmc_reserve")
            self.cap = int(W)

    def leaves(self) -> List[Tuple[str, Tuple[int, int], int]]:
        n = int(self.lib.nqmc_leaf_info(self.h, -1, None, 0, None, None, None))
        out = []
        buf = C.create_string_buffer(256)
        for i in range(n):
            r, c, o = C.c_int64(), C.c_int64(), C.c_int64()
            self._check(int(self.lib.nqmc_leaf_info(self.h, i, buf, 256, C.byref(r), C.byref(c), C.byref(o))), "leaf_info")
            out.append((buf.value.decode(), (r.value, c.value), o.value))
        return out

    def launch_count(self) -> int:
        return int(self.lib.nqmc_launch_count(self.h))

    # ------------------------------------------------------------------ parameters
    def set_params(self, theta):
        """theta: numpy float32 (P,) on host, or a device buffer (data_ptr()) of P floats."""
        if isinstance(theta, np.ndarray) and self._theta_host is not None and theta.shape == self._theta_host.shape \
                and np.array_equal(theta, self._theta_host):
            return                                   # already resident (VMCOptimizer binds theta once per iteration)
        self._theta_host = None
        if hasattr(theta, "data_ptr"):
            if int(np.prod(_shape(theta))) != self.P:
                raise ValueError(f"theta must have shape ({self.P},)")
            self._check(self.lib.nqmc_set_params(self.h, _ptr(theta), self.P, self.stream), "nqmc_set_params")
        else:
            th = np.ascontiguousarray(theta, dtype=np.float32)
            if th.shape != (self.P,):
                raise ValueError(f"theta must have shape ({self.P},), got {th.shape}")
            self._check(self.lib.nqmc_set_params_host(self.h, th.ctypes.data, self.P, self.stream), "nqmc_set_params_host")
            self._theta_host = th.copy()

    def get_params(self) -> np.ndarray:
        th = np.empty(self.P, dtype=np.float32)
        self._check(self.lib.nqmc_get_params_host(self.h, th.ctypes.data, self.P, self.stream), "nqmc_get_params_host")
        self._theta_host = th.copy()
        return th

    def adam_step(self, gsum, hdr, m, v, count: int, lr: float, clip: float, grad_norm=None):
        """Fused clip + Adam on the device; theta inside the ctx is updated in place."""
        self.reserve(max(self.cap, 1))
        self._theta_host = None                      # theta changes on the device
        self._check(self.lib.nqmc_adam_step(self.h, _ptr(gsum), _ptr(hdr), _ptr(m), _ptr(v), int(count),
                                            float(lr), float(clip), _ptr(grad_norm), self.stream), "nqmc_adam_step")

    # ------------------------------------------------------------------ layout
    def to_soa(self, r_aos) -> DeviceArray:
        """(W,N,3) numpy / device buffer -> device SoA [3N, W]."""
        a = self.upload(r_aos)
        shp = _shape(a)
        W = shp[0]
        if tuple(shp[1:]) != (self.n_elec, 3):
            raise ValueError(f"positions must have shape (W,{self.n_elec},3), got {tuple(shp)}")
        self.reserve(W)
        soa = self.empty((3 * self.n_elec, W))
        self._check(self.lib.nqmc_aos_to_soa(self.h, _ptr(a), soa.ptr, W, self.stream), "nqmc_aos_to_soa")
        if not isinstance(r_aos, DeviceArray) and not hasattr(r_aos, "data_ptr"):
            self.synchronize()          # `a` is a temporary upload; keep it alive until the transpose has run
        return soa

    def to_aos(self, soa) -> DeviceArray:
        W = _shape(soa)[1]
        a = self.empty((W, self.n_elec, 3))
        self._check(self.lib.nqmc_soa_to_aos(self.h, _ptr(soa), a.ptr, W, self.stream), "nqmc_soa_to_aos")
        return a

    # ------------------------------------------------------------------ hot path (device buffers, SoA)
    def log_psi(self, r_soa):
        W = _shape(r_soa)[1]
        self.reserve(W)
        la, sg = self.empty(W), self.empty(W)
        self._check(self.lib.nqmc_log_psi(self.h, _ptr(r_soa), W, la.ptr, sg.ptr, self.stream), "nqmc_log_psi")
        return sg, la

    def mh_step(self, r_soa, logp, sigma: float, normals=None, uniforms=None, seed: int = 0, step: int = 0,
                walker_id0: int = 0, accept=None, n_accepted=None):
        W = _shape(r_soa)[1]
        self.reserve(W)
        self._check(self.lib.nqmc_mh_step(self.h, _ptr(r_soa), _ptr(logp), _ptr(normals), _ptr(uniforms),
                                          int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                          _ptr(accept), _ptr(n_accepted), self.stream), "nqmc_mh_step")

    def sample(self, r_soa, logp, sigma: float, n_burn: int, n_samples: int, n_thin: int, samples, n_accepted=None,
               seed_burn: int = 0, seed_sample: int = 0, walker_id0: int = 0, normals=None, uniforms=None,
               sigma_dev=None, target_acceptance: float = 0.5, adapt_every: int = 0):
        """MetropolisSampler.sample on the device (nqmc_sample): walkers stay resident, kept samples land in
        ``samples`` [3N][n_samples][W]."""
        W = _shape(r_soa)[1]
        self.reserve(W)
        self._check(self.lib.nqmc_sample(self.h, _ptr(r_soa), _ptr(logp), _ptr(normals), _ptr(uniforms),
                                         int(seed_burn) & (2**64 - 1), int(seed_sample) & (2**64 - 1), int(walker_id0),
                                         float(sigma), W, int(n_burn), int(n_samples), int(n_thin), _ptr(samples),
                                         _ptr(n_accepted), _ptr(sigma_dev), float(target_acceptance), int(adapt_every),
                                         self.stream), "nqmc_sample")

    def local_energy(self, r_soa, want_parts: bool = False):
        W = _shape(r_soa)[1]
        self.reserve(W)
        e = self.empty(W)
        parts = self.empty((4, W)) if want_parts else None
        self._check(self.lib.nqmc_local_energy(self.h, _ptr(r_soa), W, e.ptr, _ptr(parts), self.stream), "nqmc_local_energy")
        return (e, parts) if want_parts else e

    def grad_accum(self, r_soa, weight=None):
        W = _shape(r_soa)[1]
        self.reserve(W)
        out = self.empty(self.P, np.float64)
        self._check(self.lib.nqmc_grad_accum(self.h, _ptr(r_soa), _ptr(weight), W, out.ptr, self.stream), "nqmc_grad_accum")
        return out

    def energy_stats(self, e_l):
        W = _shape(e_l)[0]
        self.reserve(W)
        hdr = self.empty(HDR_SIZE, np.float64)
        self._check(self.lib.nqmc_energy_stats(self.h, _ptr(e_l), W, hdr.ptr, self.stream), "nqmc_energy_stats")
        return hdr

    def blocked_stats(self, e_l, n_samples: int, W: int, n_blocks: int) -> np.ndarray:
        """(count, sum, sum of squares) of the per-chain block means -- see nqmc_blocked_stats."""
        out = self.empty(3, np.float64)
        self._check(self.lib.nqmc_blocked_stats(self.h, _ptr(e_l), int(n_samples), int(W), int(n_blocks), out.ptr,
                                                self.stream), "nqmc_blocked_stats")
        return out.numpy(self.stream)

    def walker_step_a(self, r_soa, logp, sigma, e_l, hdr, normals=None, uniforms=None, seed=0, step=0, walker_id0=0,
                      n_accepted=None):
        W = _shape(r_soa)[1]
        self.reserve(W)
        self._check(self.lib.nqmc_walker_step_a(self.h, _ptr(r_soa), _ptr(logp), _ptr(normals), _ptr(uniforms),
                                                int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                                _ptr(e_l), _ptr(n_accepted), _ptr(hdr), self.stream), "nqmc_walker_step_a")

    def walker_step_b(self, r_soa, e_l, hdr, gsum):
        W = _shape(r_soa)[1]
        self._check(self.lib.nqmc_walker_step_b(self.h, _ptr(r_soa), _ptr(e_l), W, _ptr(hdr), _ptr(gsum),
                                                self.stream), "nqmc_walker_step_b")

    def walker_step(self, r_soa, logp, sigma, e_l, hdr, gsum, normals=None, uniforms=None, seed=0, step=0,
                    walker_id0=0, n_accepted=None):
        W = _shape(r_soa)[1]
        self.reserve(W)
        self._check(self.lib.nqmc_walker_step(self.h, _ptr(r_soa), _ptr(logp), _ptr(normals), _ptr(uniforms),
                                              int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                              _ptr(e_l), _ptr(n_accepted), _ptr(hdr), _ptr(gsum),
                                              self.stream), "nqmc_walker_step")

    def walker_step_host(self, r_aos: np.ndarray, logp: np.ndarray, sigma: float, e_l: np.ndarray, hdr: np.ndarray,
                         gsum: np.ndarray, normals_aos: Optional[np.ndarray] = None,
                         uniforms: Optional[np.ndarray] = None, seed=0, step=0, walker_id0=0):
        """Host-buffer entry (the e2e path): all arrays are C-contiguous host memory, updated in place."""
        W = r_aos.shape[0]
        self.reserve(W)
        np_ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self.lib.nqmc_walker_step_host(self.h, r_aos.ctypes.data, logp.ctypes.data, np_ptr(normals_aos),
                                                   np_ptr(uniforms), int(seed) & (2**64 - 1), int(step), int(walker_id0),
                                                   float(sigma), W, e_l.ctypes.data, hdr.ctypes.data, gsum.ctypes.data,
                                                   self.stream), "nqmc_walker_step_host")

    WAIT_PHASE_A, WAIT_ALL = 0, 1

    def walker_step_host_async(self, r_aos: np.ndarray, logp: np.ndarray, sigma: float, e_l: np.ndarray, hdr: np.ndarray,
                               gsum: np.ndarray, normals_aos: Optional[np.ndarray] = None,
                               uniforms: Optional[np.ndarray] = None, seed=0, step=0, walker_id0=0):
        """``walker_step_host`` without the final synchronisation; pair with ``walker_step_host_wait``.  The host arrays
        must stay alive (and should be pinned) until the matching wait returns."""
        W = r_aos.shape[0]
        self.reserve(W)
        np_ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self.lib.nqmc_walker_step_host_async(self.h, r_aos.ctypes.data, logp.ctypes.data, np_ptr(normals_aos),
                                                         np_ptr(uniforms), int(seed) & (2**64 - 1), int(step),
                                                         int(walker_id0), float(sigma), W, e_l.ctypes.data,
                                                         hdr.ctypes.data, gsum.ctypes.data, self.stream),
                    "nqmc_walker_step_host_async")

    def walker_step_host_wait(self, what: int = 1):
        """Block until the phase-A outputs (``WAIT_PHASE_A``: walkers, log|psi|^2, E_L, header) or everything
        (``WAIT_ALL``: the gradient sums too) of the last ``walker_step_host_async`` is in the host arrays."""
        self._check(self.lib.nqmc_walker_step_host_wait(self.h, int(what)), "nqmc_walker_step_host_wait")

    def walker_step_host_upload(self, r_aos: np.ndarray, logp: np.ndarray, normals_aos: Optional[np.ndarray] = None,
                                uniforms: Optional[np.ndarray] = None):
        """Start copying the NEXT step's inputs to the device while the reverse pass of the running step executes
        (call after ``walker_step_host_wait(WAIT_PHASE_A)``); the next host step that names the same arrays skips its copy."""
        W = r_aos.shape[0]
        self.reserve(W)
        np_ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self.lib.nqmc_walker_step_host_upload(self.h, r_aos.ctypes.data, logp.ctypes.data, np_ptr(normals_aos),
                                                          np_ptr(uniforms), W), "nqmc_walker_step_host_upload")

    # ---- multi-GPU: peer-memory communicator (include/nqmc_b200.h, "multi-GPU") ----
    COMM_HANDLE_BYTES = 64

    def comm_create(self, rank: int, world: int) -> bytes:
        """Allocate this rank's inbox; returns the opaque handle to hand to every peer."""
        buf = C.create_string_buffer(self.COMM_HANDLE_BYTES)
        self._check(self.lib.nqmc_comm_create(self.h, int(rank), int(world), C.cast(buf, C.c_void_p)), "nqmc_comm_create")
        return buf.raw

    def comm_attach(self, handles) -> None:
        """Map the peers' inboxes; ``handles`` = one ``comm_create`` result per rank, in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self.lib.nqmc_comm_attach(self.h, C.cast(buf, C.c_void_p)), "nqmc_comm_attach")

    def comm_allreduce(self, vec) -> None:
        """In-place sum over ranks of a float64 device buffer (len <= max(param_count, 8))."""
        self._check(self.lib.nqmc_comm_allreduce(self.h, _ptr(vec), int(np.prod(_shape(vec))), self.stream),
                    "nqmc_comm_allreduce")

    def comm_active(self) -> bool:
        return bool(self.lib.nqmc_comm_active(self.h))

    def use_tensor_cores(self, enable: bool):
        self._check(self.lib.nqmc_use_tensor_cores(self.h, int(enable)), "nqmc_use_tensor_cores")

    def profile(self, enable: bool):
        self._check(self.lib.nqmc_profile(self.h, int(enable)), "nqmc_profile")

    def profile_read(self, kind: int) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.nqmc_profile_read(self.h, kind, C.byref(ms), C.byref(n)), "nqmc_profile_read")
        return ms.value, n.value

    def ffma_peak(self, reps: int = 5) -> float:
        self.reserve(max(self.cap, 1024))
        tf = C.c_double()
        self._check(self.lib.nqmc_ffma_peak(self.h, reps, C.byref(tf), self.stream), "nqmc_ffma_peak")
        return tf.value

    def gram(self, A, B=None, alpha: float = 1.0, diag: float = 0.0) -> DeviceArray:
        """alpha * A B^T + diag * I for device arrays A [M][K], B [N][K] (B None: symmetric A A^T) -- nqmc_gram."""
        M, K = _shape(A)
        sym = B is None
        Bm = A if sym else B
        N = _shape(Bm)[0]
        if _shape(Bm)[1] != K:
            raise ValueError("gram: contraction lengths differ")
        out = self.empty((M, N))
        self._check(self.lib.nqmc_gram(self.h, _ptr(A), _ptr(Bm), M, N, K, K, K, float(alpha), float(diag), int(sym),
                                       out.ptr, N, self.stream), "nqmc_gram")
        return out

    def log_psi_grads(self, r_soa) -> DeviceArray:
        """Per-walker log-derivatives, parameter-major [P][W] (nqmc_log_psi_grads; every ansatz)."""
        W = _shape(r_soa)[1]
        if W % 4:
            raise ValueError("log_psi_grads: the walker count must be a multiple of 4 (TMA row pitch of the SR kernels)")
        self.reserve(W)
        Ot = self.empty((self.P, W))
        self._check(self.lib.nqmc_log_psi_grads(self.h, _ptr(r_soa), W, Ot.ptr, W, self.stream), "nqmc_log_psi_grads")
        return Ot

    def apply_update(self, delta) -> None:
        """theta += delta (device fp64 vector) in place in the context."""
        self._check(self.lib.nqmc_apply_update(self.h, _ptr(delta), self.P, self.stream), "nqmc_apply_update")

    def sr_update(self, Ot, e_l, reg: float, lr: float, max_iter: int = 200, tol: float = 1e-6):
        """One Stochastic Reconfiguration update (nqmc_sr_update).  Ot: device [P][n] (centred in place), e_l: device [n].
        Returns (delta [P] fp64 device array, info [4] fp64 device array, S [P][P])."""
        P, n = _shape(Ot)
        S = self.empty((P, P))
        work = self.empty(int(self.lib.nqmc_sr_workspace_bytes(P)) // 8, np.float64)
        delta, info = self.empty(P, np.float64), self.empty(4, np.float64)
        self._check(self.lib.nqmc_sr_update(self.h, _ptr(Ot), _ptr(e_l), P, n, n, float(reg), float(lr), int(max_iter),
                                            float(tol), S.ptr, work.ptr, delta.ptr, info.ptr, self.stream), "nqmc_sr_update")
        return delta, info, S

    def debug_scratch_ptr(self, name: str) -> int:
        p = self.lib.nqmc_debug_scratch_ptr(self.h, name.encode())
        if not p:
            raise NativeError(f"no guarded scratch buffer named {name!r} (NQMC_GUARD=1 before the first call?)")
        return int(p)

    def check_guards(self) -> int:
        """NQMC_GUARD=1 only: number of scratch buffers whose guard bands were overwritten (raises with their names)."""
        n = int(self.lib.nqmc_debug_check_guards(self.h))
        if n != 0:
            self._check(-4 if n > 0 else n, "nqmc_debug_check_guards")
        return n

    def rng_fill(self, W: int, seed: int, step: int, walker_id0: int = 0):
        n, u = self.empty((3 * self.n_elec, W)), self.empty(W)
        self._check(self.lib.nqmc_rng_fill(self.h, int(seed) & (2**64 - 1), int(step), int(walker_id0), W, n.ptr,
                                           u.ptr, self.stream), "nqmc_rng_fill")
        return n, u
