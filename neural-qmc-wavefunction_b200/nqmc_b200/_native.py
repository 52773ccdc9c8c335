"""ctypes binding of libnqmc_b200.so (include/nqmc_b200.h) and the device-side Context.

PyTorch is used here for plumbing only: device buffers (``torch.empty(..., device='cuda')``),
the current CUDA stream handle and, in nqmc_b200.parallel, ``torch.distributed``.  Every
computation on walkers goes through the C ABI.  There is no CPU fallback: if the shared library
or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

_LIB_NAME = "libnqmc_b200.so"
_lib = None

KIND_SIMPLE = 0
KIND_ANTISYMMETRIC = 1
HDR_N, HDR_SUM_E, HDR_SUM_E2, HDR_N_ACCEPT, HDR_N_BAD, HDR_SIZE = 0, 1, 2, 3, 4, 8

EXPORTS = [
    "nqmc_create", "nqmc_destroy", "nqmc_last_error", "nqmc_abi_version", "nqmc_reserve", "nqmc_param_count",
    "nqmc_leaf_info", "nqmc_set_params", "nqmc_set_params_host", "nqmc_aos_to_soa", "nqmc_soa_to_aos",
    "nqmc_log_psi", "nqmc_mh_step", "nqmc_local_energy", "nqmc_grad_accum", "nqmc_energy_stats",
    "nqmc_walker_step_a", "nqmc_walker_step_b", "nqmc_walker_step", "nqmc_walker_step_host",
    "nqmc_launch_count", "nqmc_rng_fill", "nqmc_profile", "nqmc_profile_read", "nqmc_ffma_peak", "nqmc_use_tensor_cores", "nqmc_get_params", "nqmc_get_params_host", "nqmc_adam_step",
    "nqmc_comm_create", "nqmc_comm_attach", "nqmc_comm_allreduce",
]


class NativeError(RuntimeError):
    pass


class _System(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("n_elec", C.c_int32), ("n_up", C.c_int32), ("_pad", C.c_int32),
                ("R", C.POINTER(C.c_double)), ("Z", C.POINTER(C.c_double)), ("v_nn", C.c_double)]


class _Net(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hidden0", C.c_int32), ("hidden1", C.c_int32), ("use_cusp", C.c_int32),
                ("use_backflow", C.c_int32), ("_pad", C.c_int32), ("envelope_decay", C.c_double),
                ("cusp_same", C.c_double), ("cusp_diff", C.c_double)]


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
    _lib = lib
    return lib


def _torch():
    import torch
    return torch


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Context:
    """One (geometry, network shape) on one GPU.  Thin, stateful wrapper over nqmc_ctx."""

    def __init__(self, R: np.ndarray, Z: np.ndarray, n_up: int, n_dn: int, v_nn: float, kind: int,
                 hidden: Tuple[int, int], use_cusp: bool = False, use_backflow: bool = False,
                 envelope_decay: float = 1.0, cusp_same: float = 0.5, cusp_diff: float = 0.25,
                 device: Optional[int] = None):
        torch = _torch()
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device is visible; the nqmc_b200 walker loop has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
        self._Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1)
        self.n_atoms, self.n_up, self.n_dn = self._R.shape[0], int(n_up), int(n_dn)
        self.n_elec = self.n_up + self.n_dn
        self.kind = kind
        sysm = _System(self.n_atoms, self.n_elec, self.n_up, 0,
                       self._R.ctypes.data_as(C.POINTER(C.c_double)), self._Z.ctypes.data_as(C.POINTER(C.c_double)),
                       float(v_nn))
        net = _Net(kind, int(hidden[0]), int(hidden[1]), int(bool(use_cusp)), int(bool(use_backflow)), 0,
                   float(envelope_decay), float(cusp_same), float(cusp_diff))
        h = C.c_void_p()
        rc = self.lib.nqmc_create(C.byref(sysm), C.byref(net), self.device, C.byref(h))
        if rc != 0:
            raise NativeError(f"nqmc_create failed ({rc}): {self.lib.nqmc_last_error(None).decode()}")
        self.h = h
        self.P = int(self.lib.nqmc_param_count(self.h))
        self.cap = 0
        self._theta_version = None

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

    @property
    def stream(self) -> int:
        return _torch().cuda.current_stream(self.device).cuda_stream

    def dev(self):
        return _torch().device("cuda", self.device)

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
        """theta: numpy float32 (P,) on host, or a torch CUDA float32 tensor."""
        torch = _torch()
        if isinstance(theta, np.ndarray):
            th = np.ascontiguousarray(theta, dtype=np.float32)
            if th.shape != (self.P,):
                raise ValueError(f"theta must have shape ({self.P},), got {th.shape}")
            self._check(self.lib.nqmc_set_params_host(self.h, th.ctypes.data, self.P, self.stream), "nqmc_set_params_host")
        else:
            th = theta.to(self.dev(), torch.float32).contiguous()
            if tuple(th.shape) != (self.P,):
                raise ValueError(f"theta must have shape ({self.P},)")
            self._check(self.lib.nqmc_set_params(self.h, th.data_ptr(), self.P, self.stream), "nqmc_set_params")

    def get_params(self) -> np.ndarray:
        th = np.empty(self.P, dtype=np.float32)
        self._check(self.lib.nqmc_get_params_host(self.h, th.ctypes.data, self.P, self.stream), "nqmc_get_params_host")
        return th

    def adam_step(self, gsum, hdr, m, v, count: int, lr: float, clip: float, grad_norm=None):
        """Fused clip + Adam on the device; theta inside the ctx is updated in place."""
        self.reserve(max(self.cap, 1))
        self._check(self.lib.nqmc_adam_step(self.h, gsum.data_ptr(), hdr.data_ptr(), m.data_ptr(), v.data_ptr(), int(count),
                                            float(lr), float(clip), _ptr(grad_norm), self.stream), "nqmc_adam_step")

    # ------------------------------------------------------------------ layout
    def to_soa(self, r_aos):
        """(W,N,3) numpy / torch -> device SoA tensor [3N, W]."""
        torch = _torch()
        a = torch.as_tensor(np.ascontiguousarray(r_aos, dtype=np.float32) if isinstance(r_aos, np.ndarray) else r_aos)
        a = a.to(self.dev(), torch.float32).contiguous()
        W = a.shape[0]
        if tuple(a.shape[1:]) != (self.n_elec, 3):
            raise ValueError(f"positions must have shape (W,{self.n_elec},3), got {tuple(a.shape)}")
        self.reserve(W)
        soa = torch.empty((3 * self.n_elec, W), dtype=torch.float32, device=self.dev())
        self._check(self.lib.nqmc_aos_to_soa(self.h, a.data_ptr(), soa.data_ptr(), W, self.stream), "nqmc_aos_to_soa")
        return soa

    def to_aos(self, soa):
        torch = _torch()
        W = soa.shape[1]
        a = torch.empty((W, self.n_elec, 3), dtype=torch.float32, device=self.dev())
        self._check(self.lib.nqmc_soa_to_aos(self.h, soa.data_ptr(), a.data_ptr(), W, self.stream), "nqmc_soa_to_aos")
        return a

    # ------------------------------------------------------------------ hot path (device tensors, SoA)
    def log_psi(self, r_soa):
        torch = _torch()
        W = r_soa.shape[1]
        self.reserve(W)
        la = torch.empty(W, dtype=torch.float32, device=self.dev())
        sg = torch.empty(W, dtype=torch.float32, device=self.dev())
        self._check(self.lib.nqmc_log_psi(self.h, r_soa.data_ptr(), W, la.data_ptr(), sg.data_ptr(), self.stream), "nqmc_log_psi")
        return sg, la

    def mh_step(self, r_soa, logp, sigma: float, normals=None, uniforms=None, seed: int = 0, step: int = 0,
                walker_id0: int = 0, accept=None, n_accepted=None):
        W = r_soa.shape[1]
        self.reserve(W)
        self._check(self.lib.nqmc_mh_step(self.h, r_soa.data_ptr(), logp.data_ptr(), _ptr(normals), _ptr(uniforms),
                                          int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                          _ptr(accept), _ptr(n_accepted), self.stream), "nqmc_mh_step")

    def local_energy(self, r_soa, want_parts: bool = False):
        torch = _torch()
        W = r_soa.shape[1]
        self.reserve(W)
        e = torch.empty(W, dtype=torch.float32, device=self.dev())
        parts = torch.empty((4, W), dtype=torch.float32, device=self.dev()) if want_parts else None
        self._check(self.lib.nqmc_local_energy(self.h, r_soa.data_ptr(), W, e.data_ptr(), _ptr(parts), self.stream), "nqmc_local_energy")
        return (e, parts) if want_parts else e

    def grad_accum(self, r_soa, weight=None):
        torch = _torch()
        W = r_soa.shape[1]
        self.reserve(W)
        out = torch.empty(self.P, dtype=torch.float64, device=self.dev())
        self._check(self.lib.nqmc_grad_accum(self.h, r_soa.data_ptr(), _ptr(weight), W, out.data_ptr(), self.stream), "nqmc_grad_accum")
        return out

    def energy_stats(self, e_l):
        torch = _torch()
        W = e_l.shape[0]
        self.reserve(W)
        hdr = torch.empty(HDR_SIZE, dtype=torch.float64, device=self.dev())
        self._check(self.lib.nqmc_energy_stats(self.h, e_l.data_ptr(), W, hdr.data_ptr(), self.stream), "nqmc_energy_stats")
        return hdr

    def walker_step_a(self, r_soa, logp, sigma, e_l, hdr, normals=None, uniforms=None, seed=0, step=0, walker_id0=0,
                      n_accepted=None):
        W = r_soa.shape[1]
        self.reserve(W)
        self._check(self.lib.nqmc_walker_step_a(self.h, r_soa.data_ptr(), logp.data_ptr(), _ptr(normals), _ptr(uniforms),
                                                int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                                e_l.data_ptr(), _ptr(n_accepted), hdr.data_ptr(), self.stream), "nqmc_walker_step_a")

    def walker_step_b(self, r_soa, e_l, hdr, gsum):
        W = r_soa.shape[1]
        self._check(self.lib.nqmc_walker_step_b(self.h, r_soa.data_ptr(), e_l.data_ptr(), W, hdr.data_ptr(), gsum.data_ptr(),
                                                self.stream), "nqmc_walker_step_b")

    def walker_step(self, r_soa, logp, sigma, e_l, hdr, gsum, normals=None, uniforms=None, seed=0, step=0,
                    walker_id0=0, n_accepted=None):
        W = r_soa.shape[1]
        self.reserve(W)
        self._check(self.lib.nqmc_walker_step(self.h, r_soa.data_ptr(), logp.data_ptr(), _ptr(normals), _ptr(uniforms),
                                              int(seed) & (2**64 - 1), int(step), int(walker_id0), float(sigma), W,
                                              e_l.data_ptr(), _ptr(n_accepted), hdr.data_ptr(), gsum.data_ptr(),
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
        """In-place sum over ranks of a float64 device tensor (len <= max(param_count, 8))."""
        self._check(self.lib.nqmc_comm_allreduce(self.h, vec.data_ptr(), vec.numel(), self.stream), "nqmc_comm_allreduce")

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

    def rng_fill(self, W: int, seed: int, step: int, walker_id0: int = 0):
        torch = _torch()
        n = torch.empty((3 * self.n_elec, W), dtype=torch.float32, device=self.dev())
        u = torch.empty(W, dtype=torch.float32, device=self.dev())
        self._check(self.lib.nqmc_rng_fill(self.h, int(seed) & (2**64 - 1), int(step), int(walker_id0), W, n.data_ptr(),
                                           u.data_ptr(), self.stream), "nqmc_rng_fill")
        return n, u
