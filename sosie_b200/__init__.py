"""sosie_b200 -- B200-native (sm_100a) replacement for SOSIE's horizontal-interpolation hot path.

The product is the C-ABI shared library ``libsosie_b200.so`` (include/sosie_b200.h): hand-written CUDA
kernels behind the entry points the reference's Fortran modules would bind through ISO_C_BINDING
(sosie_b200/fortran/).  This Python package is only a thin ctypes mirror of that ABI, used by the
tests and by bench.py; names and argument meaning follow the reference routines:

    BILIN_2D / MAPPING_BL / INTERP_BL   src/mod_bilin_2d.f90:44,366,275
    AKIMA_2D                            src/mod_akima_2d.f90:59
    BDROWN / SMOOTHER                   src/mod_bdrown.f90:18,444
    DROWN                               src/mod_drown.f90:18
    vector rotation                     src/corr_vect.f90:622-624

There is NO CPU fallback: if the library is missing or no CUDA device is usable, every call raises.
Array convention: Fortran (ni,nj) column-major == numpy shape (nj,ni) C-contiguous.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["Sosie", "SosieError", "lib_path", "load_library", "EXPORTS", "mapping_write_bin", "mapping_read_bin", "host_alloc", "host_free", "RawDevice"]

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    # SOSIE_B200_LIB: an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("SOSIE_B200_LIB") or os.path.join(_HERE, "_build", "libsosie_b200.so")


# every symbol include/sosie_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "sosie_create", "sosie_destroy", "sosie_last_error", "sosie_set_stream", "sosie_sync", "sosie_launch_count",
    "sosie_version", "sosie_profile", "sosie_get_timing", "sosie_bilin_set_shard", "sosie_bilin_set_grids", "sosie_bilin_set_grids_dev", "sosie_bilin_map", "sosie_bilin_get_map",
    "sosie_bilin_load_map", "sosie_bilin_map_info", "sosie_bilin_apply", "sosie_bilin_apply_dev", "sosie_bilin_apply_ex_dev",
    "sosie_bilin_apply_uv_dev", "sosie_bilin_2d", "sosie_interp_bl", "sosie_akima_set_grids", "sosie_akima_apply",
    "sosie_akima_apply_dev", "sosie_akima_get_ixy", "sosie_bdrown", "sosie_bdrown_dev", "sosie_bdrown_force_general", "sosie_smoother", "sosie_drown",
    "sosie_rotate_uv", "sosie_rotate_uv_dev", "sosie_bilin_2d_first", "sosie_set_pipeline_min_targets",
    "sosie_host_register", "sosie_host_unregister", "sosie_bilin_src_rows",
    "sosie_akima_1d_3d", "sosie_akima_1d_3d_dev", "sosie_lbc_lnk", "sosie_lbc_lnk_dev", "sosie_interp3d_post", "sosie_interp3d_post_dev",
    "sosie_mapping_write_bin", "sosie_mapping_read_header_bin", "sosie_mapping_read_bin", "sosie_akima_untiled_slopes",
    "sosie_extrp_hl", "sosie_extrp_hl_dev", "sosie_angle2", "sosie_angle2_dev", "sosie_bilin_get_map_dev", "sosie_bilin_load_map_dev",
    "sosie_akima_1d_1d", "sosie_akima_1d_1d_dev", "sosie_interp3d_sigma", "sosie_interp3d_sigma_dev",
    "sosie_track_interp", "sosie_track_interp_dev", "sosie_set_exchange", "sosie_exchange_used", "sosie_find_nearest_point",
    "sosie_bilin_apply_enqueue", "sosie_akima_apply_enqueue", "sosie_bdrown_enqueue", "sosie_flush", "sosie_set_queue_depth",
    "sosie_queue_stats", "sosie_host_alloc", "sosie_host_free", "sosie_apply_tma", "sosie_apply_tile_classes", "sosie_selftest_div",
    "sosie_bilin_map_async", "sosie_set_exchange_dev", "sosie_bilin_src_axes_hint", "sosie_p2p_init", "sosie_p2p_connect",
]

_lib = None


class SosieError(RuntimeError):
    pass


def load_library():
    """dlopen libsosie_b200.so.  Raises if it has not been built (python -m sosie_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise SosieError(f"{p} is missing: build it with `python -m sosie_b200.build` (nvcc, sm_100a). "
                         "There is no CPU fallback.")
    lib = C.CDLL(p)
    lib.sosie_last_error.restype = C.c_char_p
    lib.sosie_version.restype = C.c_char_p
    lib.sosie_launch_count.restype = C.c_int64
    lib.sosie_launch_count.argtypes = [C.c_void_p]
    lib.sosie_last_error.argtypes = [C.c_void_p]
    lib.sosie_host_alloc.restype = C.c_void_p
    lib.sosie_host_alloc.argtypes = [C.c_size_t]
    lib.sosie_host_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _hp(a, dtype):
    """host address of a caller-owned array for the enqueue entries: the library writes to / reads from it AFTER the call
    returns, so no silent copy may be made here (wrong dtype or a non-contiguous view raise)"""
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):      # torch CPU tensor (pinned or not)
        if a.is_cuda or not a.is_contiguous():
            raise SosieError("enqueue: host tensors must be contiguous CPU tensors")
        return C.c_void_p(a.data_ptr())
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise SosieError(f"enqueue: arrays must be C-contiguous numpy arrays of {np.dtype(dtype).name} (no hidden copies: the call is asynchronous)")
    return C.c_void_p(a.ctypes.data)


def host_alloc(shape, dtype=np.float32):
    """page-locked numpy array (sosie_host_alloc); release with host_free(arr)"""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = lib.sosie_host_alloc(n * dtype.itemsize)
    if not p:
        raise SosieError("sosie_host_alloc failed")
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _pinned[arr.ctypes.data] = p
    return arr


def host_free(arr):
    p = _pinned.pop(arr.ctypes.data, None)
    if p:
        load_library().sosie_host_free(C.c_void_p(p))


_pinned = {}


class RawDevice:
    """a raw device address as a __cuda_array_interface__ object: torch.as_tensor(RawDevice(ptr, nbytes), device=...) gives a
    uint8 tensor VIEW of library-owned device memory (used by the exchange callbacks)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _dp(t):
    """raw device pointer of a torch CUDA tensor (or an int already)"""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


def mapping_write_bin(path, lon, lat, mapping, vflag=-9999.0):
    """binary twin of P2D_MAPPING_AB (src/io_ezcdf.f90:2195); mapping = dict(metrics, alphabeta, ipb[, dist_np])"""
    lib = load_library()
    lon, lat = _f64(lon), _f64(lat)
    ny, nx = lon.shape
    met, ab = _i32(mapping["metrics"]), _f64(mapping["alphabeta"])
    ipb = np.ascontiguousarray(mapping["ipb"], np.int16)
    dnp = mapping.get("dist_np")
    dnp = None if dnp is None else _f32(dnp)
    rc = lib.sosie_mapping_write_bin(os.fsencode(path), nx, ny, _p(lon), _p(lat), _p(met), _p(ab), _p(ipb), _p(dnp), C.c_double(vflag))
    if rc != 0:
        raise SosieError(f"sosie_mapping_write_bin({path}) failed with code {rc}")


def mapping_read_bin(path):
    """binary twin of RD_MAPPING_AB (src/io_ezcdf.f90:2280): returns dict(lon, lat, metrics, alphabeta, ipb, dist_np)"""
    lib = load_library()
    nx, ny, has = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.sosie_mapping_read_header_bin(os.fsencode(path), C.byref(nx), C.byref(ny), C.byref(has))
    if rc != 0:
        raise SosieError(f"{path}: not a sosie_b200 mapping file")
    nx, ny = nx.value, ny.value
    out = dict(lon=np.empty((ny, nx)), lat=np.empty((ny, nx)), metrics=np.empty((3, ny, nx), np.int32),
               alphabeta=np.empty((2, ny, nx)), ipb=np.empty((ny, nx), np.int16), dist_np=np.empty((ny, nx), np.float32))
    rc = lib.sosie_mapping_read_bin(os.fsencode(path), nx, ny, _p(out["lon"]), _p(out["lat"]), _p(out["metrics"]), _p(out["alphabeta"]),
                                    _p(out["ipb"]), _p(out["dist_np"]))
    if rc != 0:
        raise SosieError(f"sosie_mapping_read_bin({path}) failed with code {rc}")
    return out


class Sosie:
    """One GPU context (opaque ``sosie_ctx``): owns the device copies of the grids, the cached mapping
    (the reference's SAVE'd RAB/IMETRICS/IPB, src/mod_bilin_2d.f90:22-24) and the work buffers."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.sosie_create(C.byref(h), int(device))
        if rc != 0:
            raise SosieError(f"sosie_create(device={device}) failed with code {rc}: no usable CUDA device "
                             "(this library has no CPU fallback)")
        self._h = h
        self.device = device
        self.l_first_call_interp_routine = True   # the mod_conf global (src/mod_conf.f90:13)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sosie_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.sosie_last_error(self._h)
            raise SosieError(f"sosie_b200 error {rc}: {msg.decode() if msg else ''}")

    def set_stream(self, cuda_stream_handle):
        """run on the caller's stream (e.g. torch.cuda.Stream().cuda_stream); None = the context's own stream.
        The legacy default stream (handle 0) cannot be named through the ABI: use a real stream."""
        if cuda_stream_handle is not None and int(cuda_stream_handle) == 0:
            raise SosieError("set_stream: handle 0 is the legacy default stream; pass a non-default stream or None")
        self._ck(self._lib.sosie_set_stream(self._h, C.c_void_p(cuda_stream_handle) if cuda_stream_handle else None))

    def sync(self):
        self._ck(self._lib.sosie_sync(self._h))

    @property
    def launches(self) -> int:
        return int(self._lib.sosie_launch_count(self._h))

    T_MAP, T_PACK, T_APPLY, T_BDROWN, T_SMOOTH, T_AKIMA_PREP, T_AKIMA_EVAL, T_DROWN = range(8)

    def profile(self, enable=True):
        """bracket the dominant kernels with CUDA events on the context stream"""
        self._ck(self._lib.sosie_profile(self._h, int(enable)))

    def timing_ms(self, which: int) -> float:
        ms = C.c_float(-1.0)
        self._ck(self._lib.sosie_get_timing(self._h, int(which), C.byref(ms)))
        return float(ms.value)

    def bilin_set_shard(self, j0: int, j1: int):
        """this context handles target rows j0..j1 (1-based inclusive); (0,0) = all"""
        self._ck(self._lib.sosie_bilin_set_shard(self._h, int(j0), int(j1)))

    # ---------------------------------------------------------------- bilinear
    ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)

    def set_exchange(self, nranks, allgather):
        """prelude exchange of row-sharded first calls (sosie_set_exchange).  allgather(send: bytes) -> bytes of every rank's
        record concatenated (nranks * len(send)); None switches it off."""
        if allgather is None:
            self._xchg_cb = None
            self._ck(self._lib.sosie_set_exchange(self._h, 0, None, None))
            return

        def _cb(user, send, recv, nbytes):
            try:
                out = allgather(C.string_at(send, nbytes))
                if len(out) != nbytes * nranks:
                    return 2
                C.memmove(recv, out, len(out))
                return 0
            except Exception:  # noqa: BLE001 - reported through the C return code
                import traceback
                traceback.print_exc()
                return 1
        self._xchg_cb = self.ALLGATHER_FN(_cb)   # keep the trampoline alive
        self._ck(self._lib.sosie_set_exchange(self._h, int(nranks), self._xchg_cb, None))

    def exchange_used(self):
        """0: no exchange; 1: callback exchange (host or device); 2: peer-memory exchange"""
        return int(self._lib.sosie_exchange_used(self._h))

    def bilin_set_grids(self, k_ew_per, X10, Y10, X20, Y20, mask_domain_trg=None, l_reg_src=False, i_orca_trg=0):
        X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
        src_1d, trg_1d = X10.ndim == 1, X20.ndim == 1
        nx1, ny1 = (X10.size, Y10.size) if src_1d else (X10.shape[1], X10.shape[0])
        nx2, ny2 = (X20.size, Y20.size) if trg_1d else (X20.shape[1], X20.shape[0])
        m = None if mask_domain_trg is None else _i32(mask_domain_trg)
        self._shape = (nx1, ny1, nx2, ny2)
        self._ck(self._lib.sosie_bilin_set_grids(self._h, int(k_ew_per), nx1, ny1, _p(X10), _p(Y10), int(src_1d), int(l_reg_src),
                                                 nx2, ny2, _p(X20), _p(Y20), int(trg_1d), _p(m), int(i_orca_trg)))

    def bilin_set_grids_dev(self, k_ew_per, nx1, ny1, dX10, dY10, src_1d, nx2, ny2, dX20, dY20, trg_1d, dmask=None,
                            l_reg_src=False, i_orca_trg=0):
        self._shape = (nx1, ny1, nx2, ny2)
        self._keep = (dX10, dY10, dX20, dY20, dmask)   # borrowed by the library: keep alive
        self._ck(self._lib.sosie_bilin_set_grids_dev(self._h, int(k_ew_per), nx1, ny1, _dp(dX10), _dp(dY10), int(src_1d),
                                                     int(l_reg_src), nx2, ny2, _dp(dX20), _dp(dY20), int(trg_1d), _dp(dmask),
                                                     int(i_orca_trg)))

    def bilin_src_axes_hint(self, X10=None, Y10=None):
        """declare the host copy of device-resident 1-D source axes (sosie_bilin_src_axes_hint); no arguments clear it"""
        if X10 is None or Y10 is None:
            self._ck(self._lib.sosie_bilin_src_axes_hint(self._h, 0, 0, None, None))
            return
        X10, Y10 = _f64(X10), _f64(Y10)
        self._ck(self._lib.sosie_bilin_src_axes_hint(self._h, X10.size, Y10.size, _p(X10), _p(Y10)))

    def find_nearest_point(self, Xtrg, Ytrg, Xsrc, Ysrc, mask_domain_trg=None):
        """FIND_NEAREST_POINT (src/mod_manip.f90:834): 2-D coordinate arrays in, (JIp, JJp, mask) out; mask is None when
        mask_domain_trg is not given (the OPTIONAL dummy absent)."""
        Xtrg, Ytrg, Xsrc, Ysrc = _f64(Xtrg), _f64(Ytrg), _f64(Xsrc), _f64(Ysrc)
        assert Xtrg.ndim == 2 and Xsrc.ndim == 2 and Ytrg.shape == Xtrg.shape and Ysrc.shape == Xsrc.shape
        nyt, nxt = Xtrg.shape
        nys, nxs = Xsrc.shape
        JI = np.empty((nyt, nxt), np.int32)
        JJ = np.empty((nyt, nxt), np.int32)
        m = None if mask_domain_trg is None else _i32(mask_domain_trg).copy()
        self._ck(self._lib.sosie_find_nearest_point(self._h, nxt, nyt, _p(Xtrg), _p(Ytrg), nxs, nys, _p(Xsrc), _p(Ysrc),
                                                    _p(JI), _p(JJ), _p(m)))
        return JI, JJ, m

    def bilin_map(self):
        """MAPPING_BL on device."""
        self._ck(self._lib.sosie_bilin_map(self._h))

    def bilin_map_async(self):
        """MAPPING_BL enqueued on the context stream; with a regular 1-D source nothing is read back (status at the next sync())"""
        self._ck(self._lib.sosie_bilin_map_async(self._h))

    ALLGATHER_DEV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

    def set_exchange_dev(self, nranks, allgather):
        """device-side prelude exchange of row-sharded mappings (sosie_set_exchange_dev).  allgather(dsend, drecv, nbytes, stream):
        raw device addresses / cudaStream_t as ints; must enqueue an all-gather of nbytes bytes per rank on that stream."""
        if allgather is None:
            self._xchg_dev_cb = None
            self._ck(self._lib.sosie_set_exchange_dev(self._h, 0, None, None))
            return

        def _cb(user, dsend, drecv, nbytes, stream):
            try:
                allgather(int(dsend), int(drecv), int(nbytes), int(stream or 0))
                return 0
            except Exception:  # noqa: BLE001 - reported through the C return code
                import traceback
                traceback.print_exc()
                return 1
        self._xchg_dev_cb = self.ALLGATHER_DEV_FN(_cb)   # keep the trampoline alive
        self._ck(self._lib.sosie_set_exchange_dev(self._h, int(nranks), self._xchg_dev_cb, None))

    def p2p_init(self, nranks, rank, slot_bytes=0):
        """mailbox of the peer-memory prelude exchange (sosie_p2p_init): returns (64-byte CUDA IPC handle or None when IPC is
        not available, device address of the mailbox)"""
        h = (C.c_ubyte * 64)()
        box = C.c_void_p()
        lib = self._lib
        lib.sosie_p2p_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
        rc = lib.sosie_p2p_init(self._h, int(nranks), int(rank), int(slot_bytes), h, C.byref(box))
        if rc == 1 and box.value:      # SOSIE_ERR_CUDA from cudaIpcGetMemHandle: the mailbox exists, same-process peers can use it
            return None, int(box.value)
        self._ck(rc)
        return bytes(h), int(box.value)

    def p2p_connect(self, handles=None, mailboxes=None):
        """sosie_p2p_connect: `handles` = the ranks' 64-byte IPC handles in rank order (processes), or `mailboxes` = their device
        addresses (contexts of one process)"""
        lib = self._lib
        lib.sosie_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        if mailboxes is not None:
            arr = (C.c_void_p * len(mailboxes))(*[int(m) for m in mailboxes])
            self._ck(lib.sosie_p2p_connect(self._h, None, arr))
        else:
            blob = b"".join(handles)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            self._ck(lib.sosie_p2p_connect(self._h, buf, None))

    def bilin_get_map(self):
        nx1, ny1, nx2, ny2 = self._shape
        met = np.empty((3, ny2, nx2), np.int32)
        ab = np.empty((2, ny2, nx2), np.float64)
        ipb = np.empty((ny2, nx2), np.int16)
        dnp = np.empty((ny2, nx2), np.float32)
        msk = np.empty((ny2, nx2), np.int32)
        self._ck(self._lib.sosie_bilin_get_map(self._h, _p(met), _p(ab), _p(ipb), _p(dnp), _p(msk)))
        info = np.zeros(8, np.int32)
        self._ck(self._lib.sosie_bilin_map_info(self._h, _p(info)))
        return dict(metrics=met, alphabeta=ab, ipb=ipb, dist_np=dnp, mask=msk, info=info)

    def bilin_get_map_info(self):
        info = np.zeros(8, np.int32)
        self._ck(self._lib.sosie_bilin_map_info(self._h, _p(info)))
        return info

    def src_rows(self):
        """(lo, hi), 0-based inclusive: physical source rows the packed mapping records of this shard read"""
        r = np.zeros(2, np.int32)
        self._ck(self._lib.sosie_bilin_src_rows(self._h, _p(r)))
        return int(r[0]), int(r[1])

    def bilin_get_map_dev(self, dmetrics, dalphabeta, dipb=None, ddnp=None):
        """copy the cached mapping into the caller's device tensors (metrics i32 (3,ny,nx), alphabeta f64 (2,ny,nx), ...)"""
        self._ck(self._lib.sosie_bilin_get_map_dev(self._h, _dp(dmetrics), _dp(dalphabeta), _dp(dipb), _dp(ddnp)))

    def bilin_load_map_dev(self, nx2, ny2, dmetrics, dalphabeta, dipb=None):
        self._ck(self._lib.sosie_bilin_load_map_dev(self._h, int(nx2), int(ny2), _dp(dmetrics), _dp(dalphabeta), _dp(dipb)))

    def bilin_load_map(self, metrics, alphabeta, ipb=None):
        met, ab = _i32(metrics), _f64(alphabeta)
        ny2, nx2 = met.shape[1], met.shape[2]
        ip = None if ipb is None else np.ascontiguousarray(ipb, np.int16)
        self._ck(self._lib.sosie_bilin_load_map(self._h, nx2, ny2, _p(met), _p(ab), _p(ip)))

    def bilin_apply(self, Z1, first_call=False):
        Z1 = _f32(Z1)
        if Z1.ndim == 2:
            Z1 = Z1[None]
        nx1, ny1, nx2, ny2 = self._shape
        assert Z1.shape[1:] == (ny1, nx1), (Z1.shape, self._shape)
        Z2 = np.empty((Z1.shape[0], ny2, nx2), np.float32)
        self._ck(self._lib.sosie_bilin_apply(self._h, Z1.shape[0], _p(Z1), _p(Z2), int(first_call)))
        return Z2

    def bilin_apply_dev(self, nslab, dZ1, dZ2, first_call=False):
        self._ck(self._lib.sosie_bilin_apply_dev(self._h, int(nslab), _dp(dZ1), _dp(dZ2), int(first_call)))

    def bilin_apply_ex_dev(self, nslab, dZ1, dZ2, first_call=False, clamp=None, dmask_trg=None, rmiss_val=-9999.0):
        use = clamp is not None
        vmin, vmax = clamp if use else (0.0, 0.0)
        self._ck(self._lib.sosie_bilin_apply_ex_dev(self._h, int(nslab), _dp(dZ1), _dp(dZ2), int(first_call), int(use),
                                                    C.c_float(vmin), C.c_float(vmax), _dp(dmask_trg), C.c_float(rmiss_val)))

    def bilin_apply_uv_dev(self, nslab, dU1, dV1, dUo, dVo, dcos, dsin):
        self._ck(self._lib.sosie_bilin_apply_uv_dev(self._h, int(nslab), _dp(dU1), _dp(dV1), _dp(dUo), _dp(dVo), _dp(dcos), _dp(dsin)))

    def bilin_2d(self, k_ew_per, X10, Y10, Z1, X20, Y20, mask_domain_trg=None, l_reg_src=False, i_orca_trg=0, have_mapping=False):
        """BILIN_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2, cnpat, mask_domain_trg): builds the mapping on the
        first call (l_first_call_interp_routine), applies it to every slab of Z1."""
        X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
        Z1 = _f32(Z1)
        if Z1.ndim == 2:
            Z1 = Z1[None]
        src_1d, trg_1d = X10.ndim == 1, X20.ndim == 1
        nx1, ny1 = (X10.size, Y10.size) if src_1d else (X10.shape[1], X10.shape[0])
        nx2, ny2 = (X20.size, Y20.size) if trg_1d else (X20.shape[1], X20.shape[0])
        self._shape = (nx1, ny1, nx2, ny2)
        m = None if mask_domain_trg is None else _i32(mask_domain_trg)
        Z2 = np.empty((Z1.shape[0], ny2, nx2), np.float32)
        first = C.c_int32(1 if self.l_first_call_interp_routine else 0)
        self._ck(self._lib.sosie_bilin_2d(self._h, int(k_ew_per), nx1, ny1, _p(X10), _p(Y10), int(src_1d), _p(Z1), nx2, ny2,
                                          _p(X20), _p(Y20), int(trg_1d), _p(Z2), _p(m), int(l_reg_src), int(i_orca_trg),
                                          C.byref(first), int(have_mapping), Z1.shape[0]))
        self.l_first_call_interp_routine = bool(first.value)
        return Z2

    def bilin_2d_first(self, k_ew_per, X10, Y10, Z1, X20, Y20, mask_domain_trg=None, l_reg_src=False, i_orca_trg=0):
        """first call of BILIN_2D with MAPPING_BL's products returned in the same call (pipelined for a regular
        1-D source and a 2-D target): returns (Z2, mapping dict).  Rows outside this context's shard are zero."""
        X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
        Z1 = _f32(Z1)
        if Z1.ndim == 2:
            Z1 = Z1[None]
        src_1d, trg_1d = X10.ndim == 1, X20.ndim == 1
        nx1, ny1 = (X10.size, Y10.size) if src_1d else (X10.shape[1], X10.shape[0])
        nx2, ny2 = (X20.size, Y20.size) if trg_1d else (X20.shape[1], X20.shape[0])
        self._shape = (nx1, ny1, nx2, ny2)
        m = None if mask_domain_trg is None else _i32(mask_domain_trg)
        Z2 = np.zeros((Z1.shape[0], ny2, nx2), np.float32)
        met = np.zeros((3, ny2, nx2), np.int32)
        ab = np.zeros((2, ny2, nx2), np.float64)
        ipb = np.zeros((ny2, nx2), np.int16)
        dnp = np.zeros((ny2, nx2), np.float32)
        self._ck(self._lib.sosie_bilin_2d_first(self._h, int(k_ew_per), nx1, ny1, _p(X10), _p(Y10), int(src_1d), _p(Z1), nx2, ny2,
                                                _p(X20), _p(Y20), int(trg_1d), _p(Z2), _p(m), int(l_reg_src), int(i_orca_trg),
                                                Z1.shape[0], _p(met), _p(ab), _p(ipb), _p(dnp)))
        self.l_first_call_interp_routine = False
        info = np.zeros(8, np.int32)
        self._ck(self._lib.sosie_bilin_map_info(self._h, _p(info)))
        return Z2, dict(metrics=met, alphabeta=ab, ipb=ipb, dist_np=dnp, info=info)

    # ---------------------------------------------------------------- enqueue / flush (csrc/queue.cu)
    def bilin_apply_enqueue(self, Z1, Z2):
        """queue one slab of the BILIN_2D apply loop: Z1 (ny1, nx1) f32 -> Z2 (ny2, nx2) f32, both caller-owned numpy arrays
        (C-contiguous) or raw addresses; Z2 is complete after flush().  Page-locked inputs are borrowed until then."""
        self._ck(self._lib.sosie_bilin_apply_enqueue(self._h, _hp(Z1, np.float32), _hp(Z2, np.float32)))

    def akima_apply_enqueue(self, Z1, Z2):
        self._ck(self._lib.sosie_akima_apply_enqueue(self._h, _hp(Z1, np.float32), _hp(Z2, np.float32)))

    def bdrown_enqueue(self, k_ew, X, mask, nb_inc=200, nb_smooth=0, pmin=-1.0e24, pmax=1.0e24, pval_land=0.0, nsweeps=None):
        """queue BDROWN on one slab X (nj, ni) f32 IN PLACE with its mask (nj, ni) i32; nsweeps: optional 1-element i32 array"""
        nj, ni = X.shape
        self._ck(self._lib.sosie_bdrown_enqueue(self._h, int(k_ew), ni, nj, _hp(X, np.float32), _hp(mask, np.int32), int(nb_inc), int(nb_smooth),
                                                C.c_float(pmin), C.c_float(pmax), C.c_float(pval_land),
                                                None if nsweeps is None else _hp(nsweeps, np.int32)))

    def flush(self):
        """wait for every queued slab; the output arrays handed to the *_enqueue calls are complete afterwards"""
        self._ck(self._lib.sosie_flush(self._h))

    def set_queue_depth(self, depth: int):
        self._ck(self._lib.sosie_set_queue_depth(self._h, int(depth)))

    def queue_stats(self):
        st = np.zeros(4, np.int64)
        self._ck(self._lib.sosie_queue_stats(self._h, _p(st)))
        return dict(batches=int(st[0]), slabs=int(st[1]), staged_in=int(st[2]), staged_out=int(st[3]))

    def apply_tma(self, on=0):
        """test hook of the batched apply: bit 0 = TMA-staged kernel on (off by default), bit 2 = group-staged kernel off
        (only the 4-byte cp.async kernel runs); results are identical in every combination"""
        self._ck(self._lib.sosie_apply_tma(self._h, int(on)))

    def apply_tile_classes(self):
        """(tiles that fit the TMA box, tiles that fit the group-staged kernel, neither)"""
        c = np.zeros(3, np.int32)
        self._ck(self._lib.sosie_apply_tile_classes(self._h, _p(c)))
        return int(c[0]), int(c[1]), int(c[2])

    def selftest_div(self, n, seed=0):
        m = C.c_uint64(0)
        self._ck(self._lib.sosie_selftest_div(self._h, C.c_uint64(int(n)), C.c_uint64(int(seed)), C.byref(m)))
        return int(m.value)

    def set_pipeline_min_targets(self, n: int):
        """first calls with fewer targets take the sequential path (tests pass 0 to force the pipeline)"""
        self._ck(self._lib.sosie_set_pipeline_min_targets(self._h, C.c_int64(int(n))))

    def interp_bl(self, k_ew_per, jiP, jjP, iqd, xa, xb, Z_in):
        Z = _f32(Z_in)
        ji, jj, iq = _i32(np.atleast_1d(jiP)), _i32(np.atleast_1d(jjP)), _i32(np.atleast_1d(iqd))
        a, b = _f64(np.atleast_1d(xa)), _f64(np.atleast_1d(xb))
        out = np.empty(ji.size, np.float32)
        self._ck(self._lib.sosie_interp_bl(self._h, int(k_ew_per), Z.shape[1], Z.shape[0], _p(Z), ji.size, _p(ji), _p(jj), _p(iq),
                                           _p(a), _p(b), _p(out)))
        return out

    # ---------------------------------------------------------------- akima
    def akima_set_grids(self, k_ew_per, X10, Y10, X20, Y20, i_orca_src=0):
        X10, Y10, X20, Y20 = _f64(X10), _f64(Y10), _f64(X20), _f64(Y20)
        src_1d, trg_1d = X10.ndim == 1, X20.ndim == 1
        nx1, ny1 = (X10.size, Y10.size) if src_1d else (X10.shape[1], X10.shape[0])
        nx2, ny2 = (X20.size, Y20.size) if trg_1d else (X20.shape[1], X20.shape[0])
        self._ashape = (nx1, ny1, nx2, ny2)
        self._ck(self._lib.sosie_akima_set_grids(self._h, int(k_ew_per), nx1, ny1, _p(X10), _p(Y10), int(src_1d), int(i_orca_src),
                                                 nx2, ny2, _p(X20), _p(Y20), int(trg_1d)))

    def akima_apply(self, Z1):
        Z1 = _f32(Z1)
        if Z1.ndim == 2:
            Z1 = Z1[None]
        nx1, ny1, nx2, ny2 = self._ashape
        assert Z1.shape[1:] == (ny1, nx1)
        Z2 = np.empty((Z1.shape[0], ny2, nx2), np.float32)
        self._ck(self._lib.sosie_akima_apply(self._h, Z1.shape[0], _p(Z1), _p(Z2)))
        return Z2

    def akima_apply_dev(self, nslab, dZ1, dZ2):
        self._ck(self._lib.sosie_akima_apply_dev(self._h, int(nslab), _dp(dZ1), _dp(dZ2)))

    def akima_untiled_slopes(self, on=True):
        """test hook: 0 default (slope arrays, tiled slopes kernel), 1 per-point slopes kernel, 3 regular sources through the
        fused slopes + evaluation kernel (no slope arrays)"""
        self._ck(self._lib.sosie_akima_untiled_slopes(self._h, int(on)))

    def akima_get_ixy(self):
        nx1, ny1, nx2, ny2 = self._ashape
        ixy = np.empty((2, ny2, nx2), np.int32)
        self._ck(self._lib.sosie_akima_get_ixy(self._h, _p(ixy)))
        return ixy

    def akima_2d(self, k_ew_per, X10, Y10, Z1, X20, Y20, i_orca_src=0):
        """AKIMA_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2)"""
        self.akima_set_grids(k_ew_per, X10, Y10, X20, Y20, i_orca_src)
        return self.akima_apply(Z1)

    # ---------------------------------------------------------------- drown / smoother / rotation
    def bdrown(self, k_ew, X, mask, nb_inc=200, nb_smooth=0, pmin=-1.0e24, pmax=1.0e24, pval_land=0.0):
        """BDROWN(k_ew, X, mask, nb_inc, nb_smooth, pmin, pmax, pval_land); returns (X_drowned, nsweeps)."""
        X = _f32(X).copy()
        sq = X.ndim == 2
        if sq:
            X = X[None]
        nslab, nj, ni = X.shape
        mask = _i32(mask)
        per = mask.ndim == 3
        nsw = np.zeros(nslab, np.int32)
        self._ck(self._lib.sosie_bdrown(self._h, int(k_ew), ni, nj, nslab, _p(X), _p(mask), int(per), int(nb_inc), int(nb_smooth),
                                        C.c_float(pmin), C.c_float(pmax), C.c_float(pval_land), _p(nsw)))
        return (X[0] if sq else X), nsw

    def bdrown_force_general(self, on=True):
        """test hook: 0/False automatic, 1/True level-list general path for every slab size, 2 legacy dense path"""
        self._ck(self._lib.sosie_bdrown_force_general(self._h, int(on)))

    def bdrown_dev(self, k_ew, ni, nj, nslab, dX, dmask, mask_per_slab, nb_inc, nb_smooth, pmin, pmax, pval_land, want_nsweeps=False):
        nsw = np.zeros(nslab, np.int32) if want_nsweeps else None
        self._ck(self._lib.sosie_bdrown_dev(self._h, int(k_ew), ni, nj, nslab, _dp(dX), _dp(dmask), int(mask_per_slab), int(nb_inc),
                                            int(nb_smooth), C.c_float(pmin), C.c_float(pmax), C.c_float(pval_land), _p(nsw)))
        return nsw

    def smoother(self, k_ew, X, nb_smooth=10, msk=None, l_exclude_mask_points=False):
        X = _f32(X).copy()
        sq = X.ndim == 2
        if sq:
            X = X[None]
        nslab, nj, ni = X.shape
        m = None if msk is None else _i32(msk)
        self._ck(self._lib.sosie_smoother(self._h, int(k_ew), ni, nj, nslab, _p(X), int(nb_smooth), _p(m), int(l_exclude_mask_points)))
        return X[0] if sq else X

    def drown(self, k_ew, X, mask, nb_inc=200):
        X = _f32(X).copy()
        sq = X.ndim == 2
        if sq:
            X = X[None]
        nslab, nj, ni = X.shape
        mask = np.ascontiguousarray(mask, np.int8)
        nsw = np.zeros(nslab, np.int32)
        self._ck(self._lib.sosie_drown(self._h, int(k_ew), ni, nj, nslab, _p(X), _p(mask), int(nb_inc), _p(nsw)))
        return (X[0] if sq else X), nsw

    def track_interp(self, F1, F2, vt1, vt2, imask, rt, jiP, jjP, iqd, xa, xb, r_obs, obs_lo=-20.0, obs_hi=20.0, clip_missing=True,
                     rmissval=-9999.0):
        """per-point block of interp_to_ground_track.f90:718-779 / interp_to_hydro_section.f90:567-607 for the points between two
        model records; F1 (and F2 or None) (nk, nj, ni) or (nj, ni) f32, imask i8.  Returns (f_bl, f_np, fmask), each (n, nk)."""
        F1 = _f32(F1)
        if F1.ndim == 2:
            F1 = F1[None]
        nk, nj, ni = F1.shape
        F2 = None if F2 is None else _f32(F2).reshape(F1.shape)
        imask = np.ascontiguousarray(imask, np.int8).reshape(F1.shape)
        jiP, jjP, iqd = _i32(jiP), _i32(jjP), _i32(iqd)
        xa, xb, r_obs = _f64(xa), _f64(xb), _f64(r_obs)
        n = jiP.size
        rt = None if rt is None else _f64(rt)
        f_bl = np.full((n, nk), rmissval, np.float64); f_np = np.full((n, nk), rmissval, np.float64); fmask = np.zeros((n, nk), np.int8)
        self._ck(self._lib.sosie_track_interp(self._h, ni, nj, nk, _p(F1), _p(F2), C.c_double(vt1), C.c_double(vt2), _p(imask), C.c_int64(n),
                                              _p(rt), _p(jiP), _p(jjP), _p(iqd), _p(xa), _p(xb), _p(r_obs), C.c_double(obs_lo),
                                              C.c_double(obs_hi), int(clip_missing), C.c_double(rmissval), _p(f_bl), _p(f_np), _p(fmask)))
        return f_bl, f_np, fmask

    def track_interp_dev(self, ni, nj, nk, dF1, dF2, vt1, vt2, dimask, n, drt, djiP, djjP, diqd, dxa, dxb, dr_obs, obs_lo, obs_hi,
                         clip_missing, rmissval, df_bl, df_np, dfmask):
        self._ck(self._lib.sosie_track_interp_dev(self._h, int(ni), int(nj), int(nk), _dp(dF1), _dp(dF2), C.c_double(vt1), C.c_double(vt2),
                                                  _dp(dimask), C.c_int64(n), _dp(drt), _dp(djiP), _dp(djjP), _dp(diqd), _dp(dxa), _dp(dxb),
                                                  _dp(dr_obs), C.c_double(obs_lo), C.c_double(obs_hi), int(clip_missing),
                                                  C.c_double(rmissval), _dp(df_bl), _dp(df_np), _dp(dfmask)))

    # ---------------------------------------------------------------- vertical stage of INTERP_3D
    def akima_1d_3d(self, vz1, xf1, vz2, rfill_val=-9999.0):
        """AKIMA_1D_3D(vz1, xf1, vz2, xf2, rfill_val), src/mod_akima_1d.f90:243: xf1 (nk1, ny, nx) -> (nk2, ny, nx)"""
        vz1, vz2, xf1 = _f32(vz1), _f32(vz2), _f32(xf1)
        nk1, ny, nx = xf1.shape
        xf2 = np.empty((vz2.size, ny, nx), np.float32)
        self._ck(self._lib.sosie_akima_1d_3d(self._h, nx, ny, nk1, _p(vz1), _p(xf1), vz2.size, _p(vz2), _p(xf2), C.c_float(rfill_val)))
        return xf2

    def akima_1d_3d_dev(self, nx, ny, vz1, dxf1, vz2, dxf2, rfill_val=-9999.0):
        vz1, vz2 = _f32(vz1), _f32(vz2)
        self._ck(self._lib.sosie_akima_1d_3d_dev(self._h, nx, ny, vz1.size, _p(vz1), _dp(dxf1), vz2.size, _p(vz2), _dp(dxf2), C.c_float(rfill_val)))

    def akima_1d_1d(self, vz1, vf1, vz2, rmask_val=None, l_surf_persist=False, l_botm_persist=False):
        """AKIMA_1D_1D(vz1, vf1, vz2, vf2, rmask_val, l_surf_persist, l_botm_persist), src/mod_akima_1d.f90:26-232, over columns:
        vz1 (nk1,) shared or (ncol, nk1); vf1 (ncol, nk1); vz2 (nk2,) shared or (ncol, nk2) -> (vf2 (ncol, nk2) f32, status (ncol,)).
        Raises SosieError on columns for which the reference is undefined (status still reports which)."""
        vf1 = np.atleast_2d(_f64(vf1))
        ncol, nk1 = vf1.shape
        vz1, vz2 = _f64(vz1), _f64(vz2)
        nk2 = vz2.shape[-1]
        vf2 = np.empty((ncol, nk2), np.float32)
        status = np.zeros(ncol, np.int32)
        i64 = C.c_int64
        rc = self._lib.sosie_akima_1d_1d(self._h, i64(ncol), nk1, _p(vz1), i64(nk1 if vz1.ndim == 2 else 0), i64(1), _p(vf1), i64(nk1), i64(1),
                                         nk2, _p(vz2), i64(nk2 if vz2.ndim == 2 else 0), i64(1), _p(vf2), i64(nk2), i64(1),
                                         int(rmask_val is not None), C.c_double(0.0 if rmask_val is None else rmask_val),
                                         int(l_surf_persist), int(l_botm_persist), _p(status))
        self.last_status = status
        self.last_vf2 = vf2
        self._ck(rc)
        return vf2, status

    def akima_1d_1d_dev(self, ncol, nk1, dvz1, z1s, dvf1, f1s, nk2, dvz2, z2s, dvf2, f2s, rmask_val=None, l_surf_persist=False,
                        l_botm_persist=False, dstatus=None):
        """device arrays (f64 in, f32 out); each *s = (column stride, level stride) in elements"""
        i64 = C.c_int64
        self._ck(self._lib.sosie_akima_1d_1d_dev(self._h, i64(ncol), int(nk1), _dp(dvz1), i64(z1s[0]), i64(z1s[1]), _dp(dvf1), i64(f1s[0]), i64(f1s[1]),
                                                 int(nk2), _dp(dvz2), i64(z2s[0]), i64(z2s[1]), _dp(dvf2), i64(f2s[0]), i64(f2s[1]),
                                                 int(rmask_val is not None), C.c_double(0.0 if rmask_val is None else rmask_val),
                                                 int(l_surf_persist), int(l_botm_persist), _dp(dstatus)))

    def interp3d_sigma(self, depth_src, data3d_tmp, depth_trg, data3d_trg, mask_trg=None, lmout=False, src_sigma=False, trg_sigma=False):
        """generic-coordinate branch of INTERP_3D (src/mod_interp.f90:374-438): cubes (nk, nj, ni); returns the new data3d_trg"""
        depth_src, data3d_tmp, depth_trg = _f32(depth_src), _f32(data3d_tmp), _f32(depth_trg)
        out = _f32(data3d_trg).copy()
        nk_src, nj, ni = data3d_tmp.shape
        nk_trg = depth_trg.shape[0]
        m = None if mask_trg is None else _i32(mask_trg)
        self._ck(self._lib.sosie_interp3d_sigma(self._h, ni, nj, nk_src, _p(depth_src), _p(data3d_tmp), nk_trg, _p(depth_trg), _p(m), int(lmout),
                                                int(src_sigma), int(trg_sigma), _p(out)))
        return out

    def interp3d_sigma_dev(self, ni, nj, nk_src, ddepth_src, ddata3d_tmp, nk_trg, ddepth_trg, dmask_trg, lmout, src_sigma, trg_sigma, ddata3d_trg):
        self._ck(self._lib.sosie_interp3d_sigma_dev(self._h, int(ni), int(nj), int(nk_src), _dp(ddepth_src), _dp(ddata3d_tmp), int(nk_trg),
                                                    _dp(ddepth_trg), _dp(dmask_trg), int(lmout), int(src_sigma), int(trg_sigma), _dp(ddata3d_trg)))

    def lbc_lnk(self, nperio, X, cd_type="T", psgn=1.0, pval=None):
        """lbc_lnk(nperio, pt2d, cd_type, psgn [, pval]) on slabs X (nslab, jpj, jpi), src/mod_nemotools.f90:65"""
        X = _f32(X).copy()
        sq = X.ndim == 2
        if sq:
            X = X[None]
        nslab, jpj, jpi = X.shape
        self._ck(self._lib.sosie_lbc_lnk(self._h, int(nperio), jpi, jpj, nslab, _p(X), ord(cd_type), C.c_double(psgn),
                                         int(pval is not None), C.c_double(0.0 if pval is None else pval)))
        return X[0] if sq else X

    def interp3d_post(self, X, mask_trg=None, ixtrpl_bot=0, vmin=-1.0e30, vmax=1.0e30, rmiss_val=-9999.0, i_orca_trg=0, c_orca_trg="T", lmout=False):
        """post-interpolation block of INTERP_3D (src/mod_interp.f90:447-511) on X (nk, nj, ni)"""
        X = _f32(X).copy()
        nk, nj, ni = X.shape
        m = None if mask_trg is None else _i32(mask_trg)
        self._ck(self._lib.sosie_interp3d_post(self._h, ni, nj, nk, _p(X), _p(m), int(ixtrpl_bot), C.c_float(vmin), C.c_float(vmax),
                                               C.c_float(rmiss_val), int(i_orca_trg), ord(c_orca_trg), int(lmout)))
        return X

    def angle2(self, nperio, lamt, phit, lamu, phiu, lamv, phiv, lamf, phif):
        """ANGLE2, src/mod_nemotools.f90:607-823: (cost, sint, cosu, sinu, cosv, sinv, cosf, sinf), each (ny, nx) f64"""
        ins = [_f64(a) for a in (lamt, phit, lamu, phiu, lamv, phiv, lamf, phif)]
        ny, nx = ins[0].shape
        outs = [np.empty((ny, nx), np.float64) for _ in range(8)]
        self._ck(self._lib.sosie_angle2(self._h, int(nperio), nx, ny, *[_p(a) for a in ins], *[_p(a) for a in outs]))
        return tuple(outs)

    def extrp_hl(self, X, jj_ex_top, jj_ex_btm, nlat_icr_trg=1):
        """extrp_hl, src/mod_interp.f90:521-542, on slabs X (nslab, nj, ni) or (nj, ni)"""
        X = _f32(X).copy()
        sq = X.ndim == 2
        if sq:
            X = X[None]
        nslab, nj, ni = X.shape
        self._ck(self._lib.sosie_extrp_hl(self._h, ni, nj, nslab, _p(X), int(jj_ex_top), int(jj_ex_btm), int(nlat_icr_trg)))
        return X[0] if sq else X

    def interp3d_post_dev(self, ni, nj, nk, dX, dmask_trg, ixtrpl_bot, vmin, vmax, rmiss_val, i_orca_trg, c_orca_trg, lmout):
        self._ck(self._lib.sosie_interp3d_post_dev(self._h, int(ni), int(nj), int(nk), _dp(dX), _dp(dmask_trg), int(ixtrpl_bot), C.c_float(vmin),
                                                   C.c_float(vmax), C.c_float(rmiss_val), int(i_orca_trg), ord(c_orca_trg), int(lmout)))

    def rotate_uv(self, U, V, xcos, xsin, inverse=False):
        U, V = _f32(U), _f32(V)
        xcos, xsin = _f64(xcos), _f64(xsin)
        n = xcos.size
        nslab = U.size // n
        Uo, Vo = np.empty_like(U), np.empty_like(V)
        self._ck(self._lib.sosie_rotate_uv(self._h, C.c_int64(n), nslab, _p(U), _p(V), _p(xcos), _p(xsin), _p(Uo), _p(Vo), int(inverse)))
        return Uo, Vo

    def rotate_uv_dev(self, n, nslab, dU, dV, dcos, dsin, dUo, dVo, inverse=False):
        self._ck(self._lib.sosie_rotate_uv_dev(self._h, C.c_int64(n), int(nslab), _dp(dU), _dp(dV), _dp(dcos), _dp(dsin), _dp(dUo),
                                               _dp(dVo), int(inverse)))
