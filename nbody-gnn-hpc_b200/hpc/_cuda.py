"""ctypes binding of libnbody_b200.so plus the torch plumbing around it (device memory, streams).

This is the only module that touches the CUDA library.  There is no CPU path: if the library is
missing, or no CUDA device is visible, every entry point raises ``EngineUnavailable``.

PyTorch's role here is plumbing only -- allocating device/pinned tensors, host<->device copies and
the current stream.  All arithmetic of the hot path runs in the hand-written sm_100a kernels
behind the C ABI declared in ``include/nbody_b200.h``.
"""
from __future__ import annotations

import ctypes
import os
import sys
import threading
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent
LIB_PATH = _PKG / "lib" / "libnbody_b200.so"

NB_STEP_CONTINUE = 1
NB_STEP_SNAPSHOT = 2
NB_STEP_PEER_SYNC = 4

# N at or below which a single system runs through the one-launch ensemble kernel (K3, B = 1)
# instead of one force launch per step (K1/K2).
# Measured on B200 (tools/probe_small.py, fp64 / fp32 us per step): K3 on a cluster of 8 CTAs 3.0 / 2.1 at N = 200,
# 11.0 / 5.6 at N = 512, 29.5 / 13.9 at N = 768; K2 18.6 / 10.8 at N = 512, 18.9 / 11.1 at N = 768.
SMALL_SYSTEM_MAX_BODIES = 640

# Host results up to this size are ordinary pageable arrays (copied out of a reusable pinned staging block); larger
# ones are views of their own pinned block (see Engine.to_host).
PAGEABLE_RESULT_MAX_BYTES = 64 << 20
HOST_POOL_MAX_BYTES = 256 << 20


def _pool_refs(entry) -> int:
    """References to a pooled host array as _host_block sees them."""
    return sys.getrefcount(entry[1])


# what _pool_refs reports when nobody outside the pool refers to the array or a view of it (measured, not assumed)
_POOL_IDLE_REFS = _pool_refs((None, np.zeros(1)))


class EngineUnavailable(RuntimeError):
    """The CUDA engine cannot run here (library not built, or no GPU).  There is no fallback."""


_vp, _ci, _cd, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
_ip = ctypes.POINTER(ctypes.c_int)

_SIGNATURES = {
    "nb_abi_version": (_ci, []),
    "nb_last_error": (ctypes.c_char_p, []),
    "nb_device_info": (_ci, [_ci, _ip, _ip, _ip, _ip]),
    "nb_probe_fma_peak": (_ci, [_ci, ctypes.POINTER(ctypes.c_double), _vp, _vp]),
    "nb_probe_occupy": (_ci, [_ci, _sz, _cd, _vp]),
    "nb_padded_bodies": (_ci, [_ci]),
    "nb_segment_plan": (_ci, [_ci, _ip, _ip]),
    "nb_workspace_bytes": (_sz, [_ci, _ci, _ci]),
    "nb_pack_f64": (_ci, [_vp, _vp, _ci, _ci, _vp, _vp]),
    "nb_pack_f32": (_ci, [_vp, _vp, _ci, _ci, _vp, _vp]),
    "nb_unpack_f64": (_ci, [_vp, _ci, _vp, _vp]),
    "nb_unpack_f32": (_ci, [_vp, _ci, _vp, _vp]),
    "nb_accel_f64": (_ci, [_vp, _ci, _ci, _ci, _cd, _vp, _vp, _sz, _vp]),
    "nb_accel_f32": (_ci, [_vp, _ci, _ci, _ci, _cd, _vp, _vp, _sz, _vp]),
    "nb_kick_drift_f64": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _cd, _vp]),
    "nb_kick_drift_f32": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _cd, _vp]),
    "nb_step_f64": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _cd, _cd, _ci, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nb_step_f32": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _cd, _cd, _ci, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nb_step_peer_f64": (_ci, [_vp, _vp, _vp, _ci, _ci, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _ci, _ci, _ci, _cd, _cd, _ci,
                               _vp, _vp, _vp, _vp, _sz, _vp]),
    "nb_step_peer_f32": (_ci, [_vp, _vp, _vp, _ci, _ci, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _ci, _ci, _ci, _cd, _cd, _ci,
                               _vp, _vp, _vp, _vp, _sz, _vp]),
    "nb_step_status": (_ci, [_vp, _ci, _vp]),
    "nb_persist_max_bodies": (_ci, []),
    "nb_run_f64": (_ci, [_vp, _vp, _vp, _vp, _ci, _cd, _cd, _ci, _ci, _vp, _vp, _vp, _vp, _sz, _ip, _vp]),
    "nb_run_f32": (_ci, [_vp, _vp, _vp, _vp, _ci, _cd, _cd, _ci, _ci, _vp, _vp, _vp, _vp, _sz, _ip, _vp]),
    "nb_ensemble_max_bodies": (_ci, []),
    "nb_ensemble_workspace_bytes": (_sz, [_ci]),
    "nb_ensemble_worker_plan": (_ci, [_ci, _ci, _ci, _ci, _ip, _ci, _ip]),
    "nb_ensemble_f64": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _ci, _cd, _cd, _ci, _ci, _ci, _ci,
                              _vp, _vp, _vp, _ci, _ci, _vp, _sz, _vp]),
    "nb_ensemble_f32": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _ci, _cd, _cd, _ci, _ci, _ci, _ci,
                              _vp, _vp, _vp, _ci, _ci, _vp, _sz, _vp]),
    "nb_batched_max_bodies": (_ci, []),
    "nb_accel_batched_f64": (_ci, [_vp, _ci, _ci, _cd, _vp, _vp]),
    "nb_accel_batched_f32": (_ci, [_vp, _ci, _ci, _cd, _vp, _vp]),
    "nb_run_batched_f64": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _cd, _cd, _ci, _ci, _vp, _vp, _vp, _ip, _vp]),
    "nb_run_batched_f32": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _cd, _cd, _ci, _ci, _vp, _vp, _vp, _ip, _vp]),
    "nb_copy_rows_d2h_async": (_ci, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "nb_energy_workspace_bytes": (_sz, [_ci, _ci]),
    "nb_energy_f64": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _ci, _cd, _vp, _vp, _sz, _vp]),
    "nb_snapshot_energy_max_bodies": (_ci, []),
    "nb_snapshot_energy_f64": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _ci, _ci, _cd, _cd, _vp, _vp]),
    "nb_window_count": (_ci, [_ci, _ci, _ci]),
    "nb_window_gather_f32": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _vp]),
    "nbh_accel_direct": (_ci, [_vp, _vp, _ci, _ci, _cd, _ci, _vp]),
    "nbh_run": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _cd, _cd, _ci, _ci, _ci, _vp, _vp, _vp]),
    "nbh_ensemble_run": (_ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci, _ci, _cd, _cd, _ci, _ci, _ci, _vp, _vp, _vp]),
    "nbh_total_energy": (_ci, [_vp, _vp, _vp, _ci, _ci, _cd, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library() -> ctypes.CDLL:
    """dlopen the C-ABI library and declare every signature.  Works without a GPU."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = Path(os.environ.get("NBODY_B200_LIB", LIB_PATH))
        if not path.exists():
            raise EngineUnavailable(
                f"{path} not found: build it with `python nbody-gnn-hpc_b200/build.py` "
                "(nvcc, sm_100a).  This engine has no CPU fallback.")
        lib = ctypes.CDLL(str(path))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.nb_abi_version() != 1:
            raise EngineUnavailable(f"{path}: unexpected ABI version {lib.nb_abi_version()}")
        _lib = lib
        return lib


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def _torch():
    import torch
    return torch


def _current_device(torch) -> int:
    get = getattr(torch._C, "_cuda_getDevice", None)
    return int(get()) if get is not None else int(torch.cuda.current_device())


class DeferredDict(dict):
    """A result dict some of whose entries are still on the device: a deferred entry is fetched (one device -> host
    copy) the first time it is read, and is an ordinary entry from then on."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._deferred = {}

    def defer(self, key, fetch) -> None:
        self._deferred[key] = fetch

    def deferred(self):
        return tuple(self._deferred)

    def _fetch_all(self) -> None:
        for key in list(self._deferred):
            self[key]

    def __getitem__(self, key):
        fetch = self._deferred.pop(key, None)
        if fetch is not None:
            super().__setitem__(key, fetch())
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __contains__(self, key):
        return key in self._deferred or super().__contains__(key)

    def __iter__(self):
        yield from super().__iter__()
        yield from self._deferred

    def __len__(self):
        return super().__len__() + len(self._deferred)

    def keys(self):
        return list(iter(self))

    def values(self):
        self._fetch_all()
        return super().values()

    def items(self):
        self._fetch_all()
        return super().items()


SNAPSHOT_FIELDS = ("positions", "velocities", "accelerations")


class Engine:
    """One CUDA device's view of the library.  Mirrors the reference operations one to one."""

    def __init__(self, device=None):
        self.lib = load_library()
        torch = _torch()
        if not torch.cuda.is_available():
            raise EngineUnavailable("no CUDA device visible: the B200 engine has no CPU fallback")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        sm, maj, minr, clk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        with torch.cuda.device(self.device):
            self._check(self.lib.nb_device_info(self.device.index or 0, sm, maj, minr, clk))
        self.sm_count, self.cc, self.sm_clock_khz = sm.value, (maj.value, minr.value), clk.value
        if maj.value != 10:
            raise EngineUnavailable(
                f"device compute capability {maj.value}.{minr.value}: this library holds sm_100a code only")
        self.launches = 0  # kernels of this library enqueued through this object (for bench.py)

    # ---- helpers -----------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"libnbody_b200: {self.lib.nb_last_error().decode()} (code {rc})")

    def _raw_stream(self) -> int:
        """cudaStream_t of torch's current stream on this engine's device, as an int (the fast getter when this
        torch build has it: the public wrappers cost several microseconds per call)."""
        torch = _torch()
        get = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if get is not None:
            return int(get(self.device.index or 0))
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _stream(self):
        return ctypes.c_void_p(self._raw_stream())

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    def _suffix(self, dtype) -> str:
        d = np.dtype(dtype)
        if d == np.float64:
            return "f64"
        if d == np.float32:
            return "f32"
        raise ValueError(f"engine dtype must be float64 or float32, got {d}")

    def _tdtype(self, dtype):
        torch = _torch()
        return torch.float64 if np.dtype(dtype) == np.float64 else torch.float32

    def to_device(self, arr, dtype=None, pinned: bool = True):
        """Host ndarray -> device tensor through pinned staging (dtype conversion on device)."""
        torch = _torch()
        a = np.ascontiguousarray(arr)
        src = torch.from_numpy(a)
        if pinned and a.nbytes >= (1 << 16):
            stage = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            stage.copy_(src)
            dev = stage.to(self.device, non_blocking=True)
        else:
            dev = src.to(self.device)
        if dtype is not None and dev.dtype != dtype:
            dev = dev.to(dtype)
        return dev

    def _staging(self, nbytes: int, which: str = "sync"):
        """The engine's reusable pinned staging blocks (grown geometrically, never handed out): one for the
        synchronous to_host, one for the single to_host_async that may be in flight at a time."""
        torch = _torch()
        bufs = self.__dict__.setdefault("_stage_bufs", {})
        buf = bufs.get(which)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 20, 2 * (buf.numel() if buf is not None else 0)),
                              dtype=torch.uint8, pin_memory=True)
            bufs[which] = buf
        return buf

    def _host_block(self, nbytes: int) -> np.ndarray:
        """A host block of nbytes for a result that is handed out: page-locked, from a small pool.

        A fresh pageable array costs a page fault per 4 KB on first touch (0.4 ms for the 1.15 MB of a 400-step
        N = 200 run -- half the kernel time), a fresh pinned block a cudaHostAlloc.  So blocks are pooled and a block
        is reused once NOBODY outside the pool refers to it or to a view of it any more (reference count).  The pool
        is bounded (HOST_POOL_MAX_BYTES); when it is full of blocks still in use, the result is an ordinary array."""
        torch = _torch()
        pool = self.__dict__.setdefault("_host_pool", [])
        best = None
        for ent in pool:
            size = ent[1].nbytes
            if nbytes <= size <= max(4 * nbytes, 1 << 16) and _pool_refs(ent) <= _POOL_IDLE_REFS:
                if best is None or size < best[1].nbytes:
                    best = ent
        if best is None:
            # drop idle blocks that are too small, then add one if the budget allows
            pool[:] = [e for e in pool if _pool_refs(e) > _POOL_IDLE_REFS or e[1].nbytes >= nbytes]
            used = sum(e[1].nbytes for e in pool)
            if used + nbytes > HOST_POOL_MAX_BYTES:
                return np.empty(nbytes, dtype=np.uint8)
            tens = torch.empty(max(nbytes, 1 << 12), dtype=torch.uint8, pin_memory=True)
            best = (tens, tens.numpy())
            pool.append(best)
        return best[1][:nbytes]

    def to_host_async(self, t):
        """to_host in two halves: the device -> pinned host copy is enqueued now; the returned callable waits for
        it and returns the host array (same policy as to_host)."""
        torch = _torch()
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        t = t.contiguous()
        nbytes = t.numel() * 8
        stream = torch.cuda.current_stream(self.device)
        if nbytes <= PAGEABLE_RESULT_MAX_BYTES:
            block = self._host_block(nbytes)                      # pooled: no page faults, no second copy
            if block.base is not None:                            # page-locked: copy straight into it
                out = block.view(np.float64).reshape(tuple(t.shape))
                torch.from_numpy(out).copy_(t, non_blocking=True)
                done = torch.cuda.Event()
                done.record(stream)
                del block

                def finish_pooled():
                    done.synchronize()
                    return out
                return finish_pooled
            stage = self._staging(nbytes, "async")[:nbytes].view(torch.float64).view(t.shape)
            stage.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)

            def finish():
                done.synchronize()
                out = block.view(np.float64).reshape(tuple(t.shape))
                out[...] = stage.numpy()
                return out
            return finish
        host = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(t, non_blocking=True)
        done = torch.cuda.Event()
        done.record(stream)

        def finish_pinned():
            done.synchronize()
            return host.numpy()
        return finish_pinned

    def to_host(self, t, pinned: bool = True) -> np.ndarray:
        """Device tensor -> fresh host float64 ndarray.

        Up to PAGEABLE_RESULT_MAX_BYTES the result lives in a block of the engine's bounded pool of page-locked
        host blocks (_host_block: reused when the caller has dropped the previous result, at most
        HOST_POOL_MAX_BYTES in total, ordinary arrays beyond that).  Larger results (the 1.7 GB snapshot stacks
        of an ensemble) are views of their own pinned block: a second pass over them on the host would cost several
        times the whole GPU run; such an array keeps its block page-locked for as long as it is alive (np.array(x)
        makes a pageable copy)."""
        torch = _torch()
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        t = t.contiguous()
        nbytes = t.numel() * 8
        if not pinned or nbytes < (1 << 12):
            return t.cpu().numpy()
        if nbytes <= PAGEABLE_RESULT_MAX_BYTES:
            return self.to_host_async(t)()
        host = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def _masses_dev(self, masses):
        m = np.asarray(masses)
        if m.dtype == np.float32:
            return self.to_device(m, pinned=False), 1
        return self.to_device(np.asarray(m, dtype=np.float64), pinned=False), 0

    def fma_peak_tflops(self, mode: str) -> float:
        """Measured FMA-pipe peak: mode 'ffma', 'ffma2' or 'dfma' (TFLOP/s, 2 flops per FMA)."""
        torch = _torch()
        out = ctypes.c_double()
        scratch = torch.zeros(1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.nb_probe_fma_peak({"ffma": 0, "ffma2": 1, "dfma": 2}[mode], out,
                                                   self._p(scratch), self._stream()))
        return out.value

    def padded_bodies(self, n: int) -> int:
        return int(self.lib.nb_padded_bodies(int(n)))

    def segment_plan(self, n: int):
        a, b = ctypes.c_int(), ctypes.c_int()
        self.lib.nb_segment_plan(int(n), a, b)
        return a.value, b.value

    def workspace(self, n: int, n_i: int, dtype):
        """Scratch for nb_accel_* / nb_step_*: zeroed once (the i-tile arrival counters behind the partials must
        start at zero; every launch leaves them zero)."""
        torch = _torch()
        nbytes = int(self.lib.nb_workspace_bytes(int(n), int(n_i), int(np.dtype(dtype) == np.float64)))
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device), nbytes

    # ---- device-level operations (tensors in, tensors out; used by nbody.py and sharded.py) ----
    def pack(self, pos_dev, masses_dev, masses_f32: int, n: int, dtype, out=None):
        """API-layout positions (n,3) float64 + masses -> stream-layout tensor of `dtype`.

        `out`: optional preallocated stream (at least padded_bodies(n)*4 elements; the kernel writes
        exactly that many)."""
        torch = _torch()
        sfx = self._suffix(dtype)
        stream = out if out is not None else torch.empty(self.padded_bodies(n) * 4, dtype=self._tdtype(dtype),
                                                         device=self.device)
        self._check(getattr(self.lib, f"nb_pack_{sfx}")(self._p(pos_dev), self._p(masses_dev), masses_f32, n,
                                                         self._p(stream), self._stream()))
        self.launches += 1
        return stream

    def unpack(self, stream, n: int):
        torch = _torch()
        sfx = "f64" if stream.dtype == torch.float64 else "f32"
        pos = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        self._check(getattr(self.lib, f"nb_unpack_{sfx}")(self._p(stream), n, self._p(pos), self._stream()))
        self.launches += 1
        return pos

    def accel_slab(self, stream, n: int, i0: int, n_i: int, softening: float, ws=None):
        """K1 on rows [i0, i0+n_i): returns (n_i,3) tensor in the stream's dtype."""
        torch = _torch()
        dtype = np.float64 if stream.dtype == torch.float64 else np.float32
        sfx = self._suffix(dtype)
        if ws is None:
            ws = self.workspace(n, n_i, dtype)
        acc = torch.empty((n_i, 3), dtype=stream.dtype, device=self.device)
        self._check(getattr(self.lib, f"nb_accel_{sfx}")(self._p(stream), n, i0, n_i, float(softening), self._p(acc),
                                                          self._p(ws[0]), ws[1], self._stream()))
        self.launches += 1  # the force kernel reduces its own segment partials
        return acc

    def kick_drift_slab(self, cur, nxt, vel, acc, n: int, i0: int, n_i: int, dt: float):
        torch = _torch()
        sfx = "f64" if cur.dtype == torch.float64 else "f32"
        self._check(getattr(self.lib, f"nb_kick_drift_{sfx}")(self._p(cur), self._p(nxt), self._p(vel), self._p(acc),
                                                               n, i0, n_i, float(dt), self._stream()))
        self.launches += 1

    def step_slab(self, cur, nxt, vel, acc, n: int, i0: int, n_i: int, dt: float, softening: float, flags: int,
                  snap_pos, snap_vel, snap_acc, ws):
        torch = _torch()
        sfx = "f64" if cur.dtype == torch.float64 else "f32"
        self._check(getattr(self.lib, f"nb_step_{sfx}")(
            self._p(cur), self._p(nxt), self._p(vel), self._p(acc), n, i0, n_i, float(dt), float(softening),
            int(flags), self._p(snap_pos), self._p(snap_vel), self._p(snap_acc), self._p(ws[0]), ws[1],
            self._stream()))
        self.launches += 1  # force, segment reduction and leapfrog are one kernel

    def step_peer_slab(self, cur, next_ptrs, flag_ptrs, my_rank: int, wait_seq: int, signal_seq: int, vel, acc,
                       n: int, i0: int, n_i: int, dt: float, softening: float, flags: int, snap_pos, snap_vel,
                       snap_acc, ws):
        """K2 fused with its collective: the slab's new positions are stored into every rank's next stream
        (next_ptrs: device pointers as ints, own rank included) and ordering is by arrival words (flag_ptrs)."""
        torch = _torch()
        sfx = "f64" if cur.dtype == torch.float64 else "f32"
        P = len(next_ptrs)
        nxt = (ctypes.c_void_p * P)(*[int(q) for q in next_ptrs])
        flg = (ctypes.c_void_p * P)(*[int(q) for q in flag_ptrs])
        self._check(getattr(self.lib, f"nb_step_peer_{sfx}")(
            self._p(cur), nxt, flg, P, int(my_rank), int(wait_seq) & 0xFFFFFFFF, int(signal_seq) & 0xFFFFFFFF,
            self._p(vel), self._p(acc), n, i0, n_i, float(dt), float(softening), int(flags), self._p(snap_pos),
            self._p(snap_vel), self._p(snap_acc), self._p(ws[0]), ws[1], self._stream()))
        self.launches += 1  # force + leapfrog + peer stores + arrival words: one kernel

    def step_status(self, ws, n: int) -> None:
        """Synchronise and raise if a nb_step_peer_* launch on this workspace lost a peer (NB_ERR_PEER)."""
        self._check(self.lib.nb_step_status(self._p(ws[0]), int(n), self._stream()))

    # ---- host-level operations (ndarrays in, ndarrays out; the reference's call shapes) ----------
    def accelerations(self, positions, masses, softening: float, dtype=np.float64) -> np.ndarray:
        """compute_accelerations_direct (reference nbody.py:22-66) -> new (N,3) float64 ndarray."""
        torch = _torch()
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        n = pos.shape[0]
        with torch.cuda.device(self.device):
            pos_d = self.to_device(pos)
            m_d, f32 = self._masses_dev(masses)
            stream = self.pack(pos_d, m_d, f32, n, dtype)
            acc = self.accel_slab(stream, n, 0, n, softening)
            return self.to_host(acc)

    def resident(self, positions, velocities, accelerations, masses, dt: float, softening: float,
                 dtype=np.float64) -> "ResidentSystem":
        """Upload one system's (x, v, a, m) and keep it in device memory: what NBodySimulator holds between
        step()/run() calls."""
        return ResidentSystem(self, positions, velocities, accelerations, masses, dt, softening, dtype)

    def run(self, positions, velocities, accelerations, masses, dt: float, softening: float, n_steps: int,
            save_interval: int = 1, dtype=np.float64, snapshots: bool = True) -> dict:
        """The loop of NBodySimulator.run (reference nbody.py:232-248) from an explicit (x, v, a), one shot.

        Returns snapshot stacks (n_snap, N, 3) float64 -- row 0 is the entry state -- and the final
        synchronised state.
        """
        rs = self.resident(positions, velocities, accelerations, masses, dt, softening, dtype)
        out = rs.advance(n_steps, save_interval, snapshots=snapshots)
        fp, fv, fa = rs.download()
        res = {"final_positions": fp, "final_velocities": fv, "final_accelerations": fa}
        if snapshots:
            res.update(out)
        return res

    def run_device(self, stream_a, stream_b, vel, acc, n: int, dt: float, softening: float, n_steps: int,
                   save_interval: int, snap_pos, snap_vel, snap_acc, ws) -> bool:
        """The whole step loop of one system on device tensors, enqueued by ONE C call (nb_run_*): opening
        kick + drift, then one fused force + leapfrog launch per step.  Returns True when the final positions
        are in stream_a."""
        torch = _torch()
        sfx = "f64" if stream_a.dtype == torch.float64 else "f32"
        fin = ctypes.c_int(1)
        self._check(getattr(self.lib, f"nb_run_{sfx}")(
            self._p(stream_a), self._p(stream_b), self._p(vel), self._p(acc), n, float(dt), float(softening),
            int(n_steps), int(save_interval), self._p(snap_pos), self._p(snap_vel), self._p(snap_acc),
            self._p(ws[0]), ws[1], fin, self._stream()))
        self.launches += (1 if snap_pos is not None else 0) + (1 if n_steps else 0) + n_steps
        return bool(fin.value)

    def ensemble_device(self, x, v, a, m_d, masses_f32: int, mass_stride: int, B: int, N: int, dt: float,
                        softening: float, n_steps: int, save_interval: int, dtype, compute_a0: bool,
                        write_initial: bool, out_x, out_v, out_a, n_snap_total: int, snap_offset: int, ws=None):
        """K3 on device tensors (state in/out, snapshot stacks out).  One launch."""
        torch = _torch()
        sfx = self._suffix(dtype)
        if ws is None:
            nbytes = int(self.lib.nb_ensemble_workspace_bytes(B))
            ws = (torch.empty(nbytes, dtype=torch.uint8, device=self.device), nbytes)
        self._check(getattr(self.lib, f"nb_ensemble_{sfx}")(
            self._p(x), self._p(v), self._p(a), self._p(m_d), masses_f32, mass_stride, B, N, float(dt),
            float(softening), int(n_steps), int(save_interval), int(bool(compute_a0)), int(bool(write_initial)),
            self._p(out_x), self._p(out_v), self._p(out_a), int(n_snap_total), int(snap_offset),
            self._p(ws[0]), ws[1], self._stream()))
        self.launches += 1

    def ensemble(self, x0, v0, masses, dt: float, softening: float, n_steps: int, save_interval: int = 1,
                 dtype=np.float64, a0=None, snapshots: bool = True, outputs: str = "host", fields=None) -> dict:
        """B independent systems (reference generate_data.py:32-58,142-149): host arrays in and out.

        x0, v0: (B,N,3); masses: (N,) shared or (B,N).  a0 None -> evaluated from x0 (what the
        scripts do after assigning the shared masses, generate_data.py:47).
        fields: which snapshot stacks cross PCIe with the run (default: all three).  A stack that is left out stays in
        HBM and is fetched when the result's entry is first read (DeferredDict) -- the accelerations are a third of
        the 1.74 GB of a 300 x 200 x 400 run and create_training_dataset (reference checkpoint.py:362-384) never
        reads them.
        """
        eager = SNAPSHOT_FIELDS if fields is None else tuple(fields)
        for f in eager:
            if f not in SNAPSHOT_FIELDS:
                raise ValueError(f"fields must be a subset of {SNAPSHOT_FIELDS}, got {f!r}")
        torch = _torch()
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        B, N = x0.shape[0], x0.shape[1]
        m = np.asarray(masses)
        if m.ndim == 1:
            mass_stride = 0
        elif m.shape == (B, N):
            mass_stride = N
        else:
            raise ValueError(f"masses must be (N,) or (B,N); got {m.shape}")
        if outputs not in ("host", "device"):
            raise ValueError("outputs must be 'host' or 'device'")
        if int(self.lib.nb_ensemble_max_bodies()) < N <= int(self.lib.nb_batched_max_bodies()):
            # systems too large for one CTA's shared memory: all B side by side in every launch (K2s, grid = groups x B)
            return self._ensemble_batched(x0, v0, m, mass_stride, dt, softening, n_steps, save_interval, dtype, a0,
                                          snapshots, outputs)
        if N > int(self.lib.nb_ensemble_max_bodies()):
            if outputs == "device":
                raise NotImplementedError("device-resident outputs need systems that fit one CTA's shared memory "
                                          f"(N <= {int(self.lib.nb_ensemble_max_bodies())})")
            # systems too large for one CTA's shared memory: each fills the GPU on its own (K1/K2), one after the other
            outs = []
            for b in range(B):
                mb = m if mass_stride == 0 else m[b]
                ab = self.accelerations(x0[b], mb, softening, dtype) if a0 is None else np.asarray(a0)[b]
                outs.append(self.run(x0[b], v0[b], ab, mb, dt, softening, n_steps, save_interval, dtype=dtype,
                                     snapshots=snapshots))
            return {k: np.stack([o[k] for o in outs]) for k in outs[0]}
        n_snap = 1 + n_steps // save_interval
        with torch.cuda.device(self.device):
            x = self.to_device(x0)
            v = self.to_device(v0)
            if a0 is None:
                a = torch.zeros_like(x)
            else:
                a = self.to_device(np.ascontiguousarray(a0, dtype=np.float64))
            m_d, f32 = self._masses_dev(np.ascontiguousarray(m))
            if not snapshots:
                self.ensemble_device(x, v, a, m_d, f32, mass_stride, B, N, dt, softening, n_steps, save_interval,
                                     dtype, compute_a0=a0 is None, write_initial=False, out_x=None, out_v=None,
                                     out_a=None, n_snap_total=0, snap_offset=0)
                return {"final_positions": self.to_host(x), "final_velocities": self.to_host(v),
                        "final_accelerations": self.to_host(a)}
            if outputs == "device":
                # the stacks stay in HBM: one launch, nothing crosses PCIe but the final state
                ox = torch.empty((B, n_snap, N, 3), dtype=torch.float64, device=self.device)
                ov, oa = torch.empty_like(ox), torch.empty_like(ox)
                self.ensemble_device(x, v, a, m_d, f32, mass_stride, B, N, dt, softening, n_steps, save_interval,
                                     dtype, compute_a0=a0 is None, write_initial=True, out_x=ox, out_v=ov, out_a=oa,
                                     n_snap_total=n_snap, snap_offset=0)
                return {"positions": ox, "velocities": ov, "accelerations": oa, "final_positions": self.to_host(x),
                        "final_velocities": self.to_host(v), "final_accelerations": self.to_host(a)}
            host = [torch.empty((B, n_snap, N, 3), dtype=torch.float64, pin_memory=True) if name in eager else None
                    for name in SNAPSHOT_FIELDS]
            # (Letting the kernel store its rows straight into the pinned host arrays through UVA was measured
            # too: 35.4 ms per 300x200x400 ensemble against 34.6 ms for the staged, chunked copy below.)
            ox = torch.empty((B, n_snap, N, 3), dtype=torch.float64, device=self.device)
            ov = torch.empty_like(ox)
            oa = torch.empty_like(ox)
            # Snapshot volume (72*N bytes per system-step) drains over PCIe several times slower than the
            # kernel produces it, so the run is cut into step chunks: the rows of chunk c go to pinned host
            # memory on a copy stream while chunk c+1 computes.  Chunk edges are multiples of save_interval.
            row_bytes = N * 3 * 8
            total_bytes = 3 * B * n_snap * row_bytes
            want_chunks = int(os.environ.get("NBODY_D2H_CHUNKS", "8"))
            n_chunks = 1 if total_bytes < (32 << 20) else min(want_chunks, max(1, n_steps // save_interval))
            saves = n_steps // save_interval
            edges = sorted({(saves * c // n_chunks) * save_interval for c in range(n_chunks)} | {n_steps})
            if edges[0] != 0:
                edges.insert(0, 0)
            if len(edges) == 1:          # n_steps == 0: only the entry state is recorded
                edges.append(edges[0])
            compute = torch.cuda.current_stream(self.device)
            copier = self._copy_stream()
            copier.wait_stream(compute)
            row0 = 0
            for c in range(len(edges) - 1):
                k0, k1 = edges[c], edges[c + 1]
                first = c == 0
                rows = (k1 - k0) // save_interval + (1 if first else 0)
                self.ensemble_device(x, v, a, m_d, f32, mass_stride, B, N, dt, softening, k1 - k0, save_interval,
                                     dtype, compute_a0=first and a0 is None, write_initial=first, out_x=ox, out_v=ov,
                                     out_a=oa, n_snap_total=n_snap, snap_offset=row0)
                if rows:
                    done = torch.cuda.Event()
                    done.record(compute)
                    copier.wait_event(done)
                    pitch = n_snap * row_bytes
                    for dev_t, host_t in zip((ox, ov, oa), host):
                        if host_t is None:
                            continue
                        self._check(self.lib.nb_copy_rows_d2h_async(
                            ctypes.c_void_p(host_t.data_ptr() + row0 * row_bytes), pitch,
                            ctypes.c_void_p(dev_t.data_ptr() + row0 * row_bytes), pitch, rows * row_bytes, B,
                            ctypes.c_void_p(copier.cuda_stream)))
                row0 += rows
            res = DeferredDict({"final_positions": self.to_host(x), "final_velocities": self.to_host(v),
                                "final_accelerations": self.to_host(a)})
            copier.synchronize()
            for name, dev_t, host_t in zip(SNAPSHOT_FIELDS, (ox, ov, oa), host):
                if host_t is not None:
                    res[name] = host_t.numpy()
                else:
                    res.defer(name, lambda t=dev_t: self.to_host(t))     # keeps the device stack alive until read
            return res

    def _ensemble_batched(self, x0, v0, m, mass_stride, dt, softening, n_steps, save_interval, dtype, a0, snapshots,
                          outputs) -> dict:
        """Engine.ensemble for nb_ensemble_max_bodies() < N <= nb_batched_max_bodies(): one batched launch per step."""
        torch = _torch()
        B, N = x0.shape[0], x0.shape[1]
        sfx, td = self._suffix(dtype), self._tdtype(dtype)
        n_snap = 1 + n_steps // save_interval
        with torch.cuda.device(self.device):
            x_d = self.to_device(x0)                                           # (B, N, 3) float64
            m_d, f32 = self._masses_dev(np.ascontiguousarray(m))
            elems = self.stream_elems(N, dtype) if hasattr(self, "stream_elems") else self.padded_bodies(N) * 4
            sa = torch.empty((B, elems), dtype=td, device=self.device)
            for b in range(B):                                                  # layout change, once per run
                self.pack(x_d[b], m_d if mass_stride == 0 else m_d[b], f32, N, dtype, out=sa[b])
            sb = sa.clone()
            vel = self.to_device(v0, td).contiguous()
            if a0 is None:
                acc = torch.empty((B, N, 3), dtype=td, device=self.device)
                self._check(getattr(self.lib, f"nb_accel_batched_{sfx}")(self._p(sa), B, N, float(softening),
                                                                        self._p(acc), self._stream()))
                self.launches += 1
            else:
                acc = self.to_device(np.ascontiguousarray(a0, dtype=np.float64), td).contiguous()
            snaps = (torch.empty((3, B, n_snap, N, 3), dtype=torch.float64, device=self.device) if snapshots else None)
            fin = ctypes.c_int(1)
            self._check(getattr(self.lib, f"nb_run_batched_{sfx}")(
                self._p(sa), self._p(sb), self._p(vel), self._p(acc), B, N, float(dt), float(softening), int(n_steps),
                int(save_interval), self._p(snaps[0] if snapshots else None), self._p(snaps[1] if snapshots else None),
                self._p(snaps[2] if snapshots else None), fin, self._stream()))
            self.launches += (2 if snapshots else 1) + n_steps
            cur = sa if fin.value else sb
            final_pos = torch.stack([self.unpack(cur[b], N) for b in range(B)])
            res = {"final_positions": self.to_host(final_pos), "final_velocities": self.to_host(vel),
                   "final_accelerations": self.to_host(acc)}
            if snapshots:
                if outputs == "device":
                    res.update(positions=snaps[0], velocities=snaps[1], accelerations=snaps[2])
                else:
                    res.update(positions=self.to_host(snaps[0]), velocities=self.to_host(snaps[1]),
                               accelerations=self.to_host(snaps[2]))
            return res

    def _copy_stream(self):
        if getattr(self, "_copier", None) is None:
            self._copier = _torch().cuda.Stream(device=self.device)
        return self._copier

    def energy(self, positions, velocities, masses, softening: float):
        """compute_total_energy (reference nbody.py:101-130) -> (K, U, K+U) Python floats."""
        torch = _torch()
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        n = pos.shape[0]
        with torch.cuda.device(self.device):
            pos_d = self.to_device(pos)
            vel_d = self.to_device(np.ascontiguousarray(velocities, dtype=np.float64))
            m_d, f32 = self._masses_dev(masses)
            ku = self.energy_slab(pos_d, vel_d, m_d, f32, n, 0, n, softening)
            k, u = (float(t) for t in ku.cpu())
        return k, u, k + u

    def window_gather(self, pos_d, vel_d, n_states: int, sequence_length: int, stride: int = 1):
        """K5 on device snapshot stacks (B, rows, N, 3) float64: (inputs (B*S, L, N, 6), targets (B*S, N, 6)) float32
        device tensors, trajectory-major -- the sample loop of reference checkpoint.py:362-384 for B trajectories."""
        torch = _torch()
        B, rows, N = int(pos_d.shape[0]), int(pos_d.shape[1]), int(pos_d.shape[2])
        if not (pos_d.is_contiguous() and vel_d.is_contiguous() and pos_d.dtype == torch.float64
                and vel_d.dtype == torch.float64 and tuple(vel_d.shape) == tuple(pos_d.shape) and pos_d.shape[3] == 3):
            raise ValueError("window_gather needs contiguous float64 (B, rows, N, 3) position and velocity stacks")
        S = int(self.lib.nb_window_count(int(n_states), int(sequence_length), int(stride)))
        inputs = torch.empty((B * S, sequence_length, N, 6), dtype=torch.float32, device=self.device)
        targets = torch.empty((B * S, N, 6), dtype=torch.float32, device=self.device)
        if S:
            self._check(self.lib.nb_window_gather_f32(self._p(pos_d), self._p(vel_d), B, rows, N, int(n_states),
                                                      int(sequence_length), int(stride), self._p(inputs),
                                                      self._p(targets), self._stream()))
            self.launches += 1
        return inputs, targets

    def snapshot_energies(self, pos_d, vel_d, m_d, masses_f32: int, mass_stride: int, softening: float,
                          G: float = 6.67430e-11):
        """K4b on device snapshot stacks (B, S, N, 3) float64 -> (B, S, 5) device tensor: K, U, px, py, pz of every
        snapshot (reference src/utils/metrics.py:62-137 for B trajectories at once).  One launch."""
        torch = _torch()
        ok = vel_d.dim() == 4 and vel_d.shape[3] == 3 and vel_d.dtype == torch.float64 and vel_d.is_contiguous()
        if pos_d is not None:
            ok = ok and tuple(pos_d.shape) == tuple(vel_d.shape) and pos_d.dtype == torch.float64 and pos_d.is_contiguous()
        if not ok:
            raise ValueError("snapshot_energies needs contiguous float64 (B, S, N, 3) position and velocity stacks")
        B, S, N = int(vel_d.shape[0]), int(vel_d.shape[1]), int(vel_d.shape[2])
        out = torch.empty((B, S, 5), dtype=torch.float64, device=self.device)
        if N > int(self.lib.nb_snapshot_energy_max_bodies()) and pos_d is not None:
            # large systems: each state fills the GPU on its own (K4), one after the other
            for b in range(B):
                mb = m_d if mass_stride == 0 else m_d[b]
                for s_ in range(S):
                    out[b, s_, :2] = self.energy_slab(pos_d[b, s_], vel_d[b, s_], mb, masses_f32, N, 0, N, softening)
                    out[b, s_, 2:] = (mb.to(torch.float64)[:, None] * vel_d[b, s_]).sum(dim=0)
            if G != 6.67430e-11:
                out[:, :, 1] *= G / 6.67430e-11
            return out
        self._check(self.lib.nb_snapshot_energy_f64(self._p(pos_d), self._p(vel_d), self._p(m_d), masses_f32,
                                                    mass_stride, B, S, N, float(G), float(softening), self._p(out),
                                                    self._stream()))
        self.launches += 1
        return out

    def energy_slab(self, pos_d, vel_d, m_d, masses_f32: int, n: int, i0: int, n_i: int, softening: float):
        torch = _torch()
        nbytes = int(self.lib.nb_energy_workspace_bytes(n, n_i))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        ku = torch.empty(2, dtype=torch.float64, device=self.device)
        self._check(self.lib.nb_energy_f64(self._p(pos_d), self._p(vel_d), self._p(m_d), masses_f32, n, i0, n_i,
                                           float(softening), self._p(ku), self._p(ws), nbytes, self._stream()))
        self.launches += 2
        return ku


class ResidentSystem:
    """One system's state in device memory between calls: NBodySimulator.step()/run() (reference nbody.py:202-248)
    advance it where it lies; nothing crosses PCIe until the caller looks at the state.

    N <= SMALL_SYSTEM_MAX_BODIES: (x, v, a) float64 in API layout, advanced by the one-launch kernel K3 (a cluster of
    8 CTAs per system); larger: the two position streams plus (v, a) in the kernel dtype, advanced by one fused
    force + leapfrog launch per step (K2), all steps of a call enqueued by one C call."""

    def __init__(self, eng: Engine, positions, velocities, accelerations, masses, dt, softening, dtype):
        torch = _torch()
        self.eng, self.dt, self.softening, self.dtype = eng, float(dt), float(softening), np.dtype(dtype)
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        self.n = n = pos.shape[0]
        self.small = n <= SMALL_SYSTEM_MAX_BODIES
        self._step_call = None
        self._index = eng.device.index or 0
        xva = np.stack([pos, np.asarray(velocities, dtype=np.float64), np.asarray(accelerations, dtype=np.float64)])
        with torch.cuda.device(eng.device):
            self.m_d, self.m_f32 = eng._masses_dev(masses)
            xva_d = eng.to_device(xva)                                  # ONE host -> device copy of the state
            if self.small:
                self.xva = xva_d.view(3, 1, n, 3)
                nbytes = int(eng.lib.nb_ensemble_workspace_bytes(1))
                self.ws = (torch.empty(nbytes, dtype=torch.uint8, device=eng.device), nbytes)
            else:
                td = eng._tdtype(self.dtype)
                self.cur = eng.pack(xva_d[0], self.m_d, self.m_f32, n, self.dtype)
                self.nxt = self.cur.clone()
                self.vel = xva_d[1].to(td) if td != torch.float64 else xva_d[1].clone()
                self.acc = xva_d[2].to(td) if td != torch.float64 else xva_d[2].clone()
                self.ws = eng.workspace(n, n, self.dtype)

    def step(self) -> None:
        """One kick-drift-kick step, nothing returned: the per-call path of `for ...: sim.step()`.  The C call and
        its arguments are prepared once; per step only dt, softening and the current stream are refreshed."""
        torch = _torch()
        eng = self.eng
        if not self.small:
            self.advance(1)
            return
        call = self._step_call
        if call is None:
            fn = getattr(eng.lib, "nb_ensemble_" + eng._suffix(self.dtype))
            P = eng._p
            call = self._step_call = [fn, [P(self.xva[0]), P(self.xva[1]), P(self.xva[2]), P(self.m_d), self.m_f32, 0, 1,
                                           self.n, 0.0, 0.0, 1, 1, 0, 0, None, None, None, 0, 0, P(self.ws[0]),
                                           self.ws[1], None]]
        fn, args = call
        args[8], args[9] = self.dt, self.softening
        if _current_device(torch) == self._index:
            args[21] = eng._raw_stream()
            rc = fn(*args)
        else:
            with torch.cuda.device(eng.device):
                args[21] = eng._raw_stream()
                rc = fn(*args)
        if rc:
            eng._check(rc)
        eng.launches += 1

    def advance(self, n_steps: int, save_interval: int = 1, snapshots: bool = False):
        """n_steps kick-drift-kick steps on the device.  snapshots=True: returns {'positions','velocities',
        'accelerations'}: host stacks (1 + n_steps // save_interval, N, 3) float64, row 0 = the state on entry
        (one device -> host copy for all three)."""
        pending = self.advance_async(n_steps, save_interval, snapshots)
        return pending() if pending is not None else None

    def advance_async(self, n_steps: int, save_interval: int = 1, snapshots: bool = False):
        """advance() in two halves: everything is enqueued here (kernels and, with snapshots, the device -> pinned
        host copy); the returned callable waits for it and hands out the host stacks.  Host-side bookkeeping done
        between the two overlaps the GPU."""
        torch = _torch()
        eng, n = self.eng, self.n
        n_snap = 1 + n_steps // save_interval
        with torch.cuda.device(eng.device):
            snaps = torch.empty((3, n_snap, n, 3), dtype=torch.float64, device=eng.device) if snapshots else None
            if self.small:
                ox, ov, oa = ((snaps[0].view(1, n_snap, n, 3), snaps[1].view(1, n_snap, n, 3),
                               snaps[2].view(1, n_snap, n, 3)) if snapshots else (None, None, None))
                eng.ensemble_device(self.xva[0], self.xva[1], self.xva[2], self.m_d, self.m_f32, 0, 1, n, self.dt,
                                    self.softening, n_steps, save_interval, self.dtype, compute_a0=False,
                                    write_initial=snapshots, out_x=ox, out_v=ov, out_a=oa,
                                    n_snap_total=n_snap if snapshots else 0, snap_offset=0, ws=self.ws)
            else:
                sp, sv, sa = (snaps[0], snaps[1], snaps[2]) if snapshots else (None, None, None)
                in_a = eng.run_device(self.cur, self.nxt, self.vel, self.acc, n, self.dt, self.softening, n_steps,
                                      save_interval, sp, sv, sa, self.ws)
                if not in_a:
                    self.cur, self.nxt = self.nxt, self.cur
            if not snapshots:
                return None
            finish = eng.to_host_async(snaps)

        def wait():
            host = finish()
            return {"positions": host[0], "velocities": host[1], "accelerations": host[2]}
        return wait

    def download(self):
        """(positions, velocities, accelerations): fresh host float64 (N,3) arrays of the current state."""
        torch = _torch()
        eng = self.eng
        with torch.cuda.device(eng.device):
            if self.small:
                h = eng.to_host(self.xva.view(3, self.n, 3))
            else:
                h = eng.to_host(torch.stack([eng.unpack(self.cur, self.n), self.vel.to(torch.float64),
                                             self.acc.to(torch.float64)]))
        return h[0], h[1], h[2]

    def energy(self):
        """(K, U, K+U) of the resident state (K4, float64)."""
        torch = _torch()
        eng, n = self.eng, self.n
        with torch.cuda.device(eng.device):
            if self.small:
                pos_d, vel_d = self.xva[0, 0], self.xva[1, 0]
            else:
                pos_d, vel_d = eng.unpack(self.cur, n), self.vel.to(torch.float64)
            ku = eng.energy_slab(pos_d, vel_d, self.m_d, self.m_f32, n, 0, n, self.softening)
            k, u = (float(t) for t in ku.cpu())
        return k, u, k + u


_engines: dict = {}


def get_engine(device=None) -> Engine:
    """The process-wide Engine for `device` (default: torch's current CUDA device).  CUDA is
    initialised on first use, never at import time, so fork-based worker pools keep working."""
    torch = _torch()
    if device is None:
        if not torch.cuda.is_available():
            load_library()  # report a missing library first: it is the more actionable error
            raise EngineUnavailable("no CUDA device visible: the B200 engine has no CPU fallback")
        key = (os.getpid(), torch.cuda.current_device())
    else:
        key = (os.getpid(), torch.device(device).index or 0)
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(torch.device("cuda", key[1]))
        _engines[key] = eng
    return eng
