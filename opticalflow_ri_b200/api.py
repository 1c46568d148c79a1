"""Host-side Python API over libofri.so: a `Handle` (one GPU, one stream) with one method per C entry point.
All arithmetic happens in the CUDA library; this file only marshals numpy arrays and parameters."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (ALGO_EXTERNAL, ALGO_FB, ALGO_HS, ALGO_LK, ALGO_LS, ALGO_NONE, Algo, Band, FarnebackParams, LkParams,
                   OfriError, Params)

_EXC = {_lib.ERR_INVALID: ValueError, _lib.ERR_ALPHAS: IndexError, _lib.ERR_FILTER_OPT: TypeError,
        _lib.ERR_TOO_SMALL: ValueError, _lib.ERR_UNSUPPORTED: NotImplementedError, _lib.ERR_OOM: MemoryError,
        _lib.ERR_INDEX: IndexError}


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _batched(a):
    a = _f32(a)
    if a.ndim == 2:
        return a[None], True
    if a.ndim != 3:
        raise ValueError("expected a (H, W) or (batch, H, W) array, got shape %r" % (a.shape,))
    return a, False


def _same_shape(first, **others):
    """Every companion array of a call must have the (batch, H, W) of the first one: the C entry points take raw
    pointers + one set of sizes, so a mismatch would read past a buffer (the reference raises a broadcast error)."""
    for name, a in others.items():
        if a is not None and a.shape != first.shape:
            raise ValueError("operands could not be broadcast together: %s has shape %r, expected %r"
                             % (name, a.shape, first.shape))


def hs_algo(alphas_in_order, niter):
    a = Algo()
    a.kind = ALGO_HS
    a.hs_niter = int(niter)
    if len(alphas_in_order) > _lib.OFRI_MAX_ALPHAS:
        raise ValueError("more than %d HS alphas" % _lib.OFRI_MAX_ALPHAS)
    a.n_alphas = len(alphas_in_order)
    for i, v in enumerate(alphas_in_order):
        a.alphas[i] = float(np.float32(v))
    return a


def ls_algo(h, maxiter=60, tol=1e-8):
    a = Algo()
    a.kind = ALGO_LS
    a.ls_h = float(np.float32(h))
    a.ls_maxiter = int(maxiter)
    a.ls_tol = float(tol)
    return a


def external_algo():
    """A foreign adapter (the reference's duck-typed compute() protocol) driven through Handle.pyramidal_flow_external."""
    a = Algo()
    a.kind = ALGO_EXTERNAL
    return a


def fb_algo():
    """The Farneback adapter as main / optional adapter of make_params; its parameters go to the handle separately
    (Handle.set_farneback), they do not fit the fixed-size ofri_algo."""
    a = Algo()
    a.kind = ALGO_FB
    return a


def lk_algo():
    """The dense Lucas-Kanade adapter as main / optional adapter of make_params (parameters: Handle.set_lk)."""
    a = Algo()
    a.kind = ALGO_LK
    return a


def lk_params(n_iters=5, half_window=13, asym=(0, 0, 0, 0)):
    """ofri_lk_params: Niter, halfWindow (LK:34) and the asymmetric-window switches [left, right, top, bottom]."""
    p = LkParams()
    p.size = C.sizeof(LkParams)
    p.n_iters, p.half_window = int(n_iters), int(half_window)
    for i in range(4):
        p.asym[i] = int(asym[i])
    return p


def farneback_params(window_size, n_iters, poly_n, use_gaussian, extra_levels, pyr_scale, g_half, xg_half, xxg_half, ig,
                     win_half, blur_halves):
    """ofri_farneback_params from the adapter's constructor values and its host-side coefficient tables."""
    p = FarnebackParams()
    p.size = C.sizeof(FarnebackParams)
    p.window_size, p.n_iters, p.poly_n = int(window_size), int(n_iters), int(poly_n)
    p.use_gaussian, p.extra_levels, p.pyr_scale = int(bool(use_gaussian)), int(extra_levels), float(pyr_scale)
    if int(extra_levels) >= _lib.OFRI_FB_MAX_LEVELS or int(window_size) // 2 > _lib.OFRI_FB_MAX_HALF:
        raise ValueError("Farneback: too many internal levels / window too large")
    for dst, src in ((p.g, g_half), (p.xg, xg_half), (p.xxg, xxg_half), (p.ig, ig), (p.win_kernel, win_half)):
        for i, v in enumerate(np.asarray(src, np.float32).ravel()):
            dst[i] = float(v)
    for k, bk in enumerate(blur_halves):
        bk = np.asarray(bk, np.float32).ravel()
        if len(bk) - 1 > _lib.OFRI_FB_MAX_HALF:
            raise ValueError("Farneback: pre-blur kernel too long")
        p.n_blur[k] = len(bk) - 1
        for i, v in enumerate(bk):
            p.blur_kernel[k][i] = float(v)
    return p


def no_algo():
    a = Algo()
    a.kind = ALGO_NONE
    return a


def gaussian_taps(sigma, ksize):
    """prepareGaussianKernel (gaussian_filter.py:47-52), generated with numpy exactly as the reference does so the
    coefficients handed to the kernels are bit-identical."""
    k = np.zeros(ksize, dtype=np.float32)
    xs = np.arange(-ksize / 2, ksize / 2, 1, dtype=int)
    k[:] = 1.0 / np.sqrt(2.0 * np.pi * sigma ** 2) * np.exp(-xs ** 2 / (2.0 * sigma ** 2))
    k /= np.sum(k)
    return k


def make_params(main, optional=None, filter_sigma=0.0, filter_opt_sigma=None, pyramid_levels=1, k_levels=1,
                warping=True, bilinear=True, intermediate_scaling=True, final_scaling=False,
                main_taps=3, opt_taps=5):
    p = Params()
    p.size = C.sizeof(Params)
    p.pyramid_levels = int(pyramid_levels)
    p.k_levels = int(k_levels)
    p.warping = int(bool(warping))
    p.bilinear = int(bool(bilinear))
    p.intermediate_scaling = int(bool(intermediate_scaling))
    p.final_scaling = int(bool(final_scaling))
    p.main_algo = main
    p.opt_algo = optional if optional is not None else no_algo()
    if filter_sigma > 1e-3:                                         # GPOF:368
        t = gaussian_taps(filter_sigma, main_taps)
        p.n_taps_main = len(t)
        for i, v in enumerate(t):
            p.taps_main[i] = v
    p.refilter_k = int(filter_sigma > 1)                            # GPOF:396
    if warping and not bilinear:                                    # GPOF:210-212: gaussian_filter(x, 0.6*3, truncate=4/0.6*3)
        mask_size = 3
        sg, tr = 0.6 * mask_size, 4.0 / 0.6 * mask_size
        t = gaussian_taps(sg, 2 * int(tr * sg + 0.5) + 1)
        p.n_taps_lsw = len(t)
        for i, v in enumerate(t):
            p.taps_lsw[i] = v
    if optional is not None and optional.kind != ALGO_NONE:
        if filter_opt_sigma is None:                                # GPOF:380 compares None > 1e-3
            raise TypeError("'>' not supported between instances of 'NoneType' and 'float'")
        if filter_opt_sigma > 1e-3:
            t = gaussian_taps(filter_opt_sigma, opt_taps)
            p.n_taps_opt = len(t)
            for i, v in enumerate(t):
                p.taps_opt[i] = v
    return p


class Handle:
    """One GPU context of libofri (ofri_create / ofri_destroy)."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        rc = self._L.ofri_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = self._L.ofri_last_error(None).decode()
            self._h = None
            raise OfriError(rc, msg)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.ofri_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -----------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._L.ofri_last_error(self._h).decode()
            raise _EXC.get(rc, OfriError)(msg) if rc in _EXC else OfriError(rc, msg)

    def set_option(self, key, value):
        self._check(self._L.ofri_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int()
        self._check(self._L.ofri_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.ofri_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._check(self._L.ofri_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._L.ofri_launch_count(self._h))

    def stage_timings(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = self._L.ofri_stage_timings(self._h, names, ms, 64)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def phase_cycles(self, family):
        """ofri_debug_phase_read: per-phase SM cycles of the persistent kernels (0 = HS, 1 = LS) since the last read;
        all zero unless the loaded library is the -DOFRI_PHASE_TIMING build."""
        out = (C.c_ulonglong * 8)()
        self._check(self._L.ofri_debug_phase_read(self._h, int(family), out))
        return [int(x) for x in out]

    # -- whole path ---------------------------------------------------------------------------------------------
    def pyramidal_flow(self, im1, im2, params, want_errors=False):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        if a.shape != b.shape:
            raise ValueError("im1 and im2 must have the same shape")
        B, H, W = a.shape
        U = np.empty((B, H, W), np.float32)
        V = np.empty((B, H, W), np.float32)
        ncall = max(params.pyramid_levels * params.k_levels, 1)
        err = np.zeros((B, ncall, 2), np.float32) if want_errors else None
        self._check(self._L.ofri_pyramidal_flow(self._h, a.ctypes.data, b.ctypes.data, B, H, W, C.byref(params),
                                                U.ctypes.data, V.ctypes.data, err.ctypes.data if want_errors else None))
        if single:
            U, V = U[0], V[0]
            err = err[0] if want_errors else None
        return (U, V, err) if want_errors else (U, V)

    def pinned_empty(self, shape, dtype=np.float32):
        """A numpy array in page-locked host memory (ofri_host_alloc); freed when the array is garbage collected.  The
        host-pointer calls are fully asynchronous on such buffers (no bounce through the library's own pinned ring)."""
        shape = tuple(int(x) for x in np.atleast_1d(shape))
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        self._check(self._L.ofri_host_alloc(self._h, C.c_size_t(max(nbytes, 1)), C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        L, h, addr = self._L, self, p.value
        import weakref
        weakref.finalize(buf, lambda: h._h and L.ofri_host_free(h._h, C.c_void_p(addr)))
        return arr

    def pyramidal_flow_external(self, im1, im2, params, compute_main=None, compute_optional=None, want_errors=False):
        """One pair with FOREIGN adapters (ofri_pyramidal_flow_external): compute_main / compute_optional are Python
        callables compute(im1, im2, U, V) -> (U, V, error) for the adapters whose kind in `params` is ALGO_EXTERNAL.
        The level stages and any HS / Liu-Shen adapter stay on the GPU; per external compute() only the level's frames
        and the current (U, V) come down and the adapter's (U, V) goes back up."""
        a = _f32(im1)
        b = _f32(im2)
        if a.ndim != 2 or a.shape != b.shape:
            raise ValueError("foreign adapters take one (H, W) pair")
        H, W = a.shape
        U = np.empty((H, W), np.float32)
        V = np.empty((H, W), np.float32)
        ncall = max(params.pyramid_levels * params.k_levels, 1)
        err = np.zeros((ncall, 2), np.float32)
        raised = []

        def cb(user, which, call_index, p1, p2, pu, pv, h, w, perr):
            try:
                fn = compute_optional if which else compute_main
                shp = (h, w)
                i1 = np.ctypeslib.as_array(p1, shape=shp)
                i2 = np.ctypeslib.as_array(p2, shape=shp)
                u = np.ctypeslib.as_array(pu, shape=shp)
                v = np.ctypeslib.as_array(pv, shape=shp)
                res = fn(i1.copy(), i2.copy(), u.copy(), v.copy())      # the adapter may keep or modify what it gets
                u[...] = np.asarray(res[0], dtype=np.float32)
                v[...] = np.asarray(res[1], dtype=np.float32)
                try:
                    perr[0] = float(res[2])
                except Exception:
                    perr[0] = 0.0
                return 0
            except BaseException as e:          # noqa: BLE001 -- re-raised below, outside the C frame
                raised.append(e)
                return 1

        rc = self._L.ofri_pyramidal_flow_external(self._h, a.ctypes.data, b.ctypes.data, H, W, C.byref(params),
                                                  _lib.ADAPTER_FN(cb), None, U.ctypes.data, V.ctypes.data, err.ctypes.data)
        if raised:
            raise raised[0]
        self._check(rc)
        return (U, V, err) if want_errors else (U, V)

    def pyramidal_flow_ptr(self, im1_ptr, im2_ptr, batch, H, W, params, u_ptr, v_ptr, err_ptr=None, device=False):
        """Raw-pointer form: HOST pointers (device=False; e.g. pinned torch tensors) or DEVICE pointers."""
        fn = self._L.ofri_pyramidal_flow_dev if device else self._L.ofri_pyramidal_flow
        self._check(fn(self._h, im1_ptr, im2_ptr, int(batch), int(H), int(W), C.byref(params), u_ptr, v_ptr, err_ptr))

    # -- row-band mode: one very large pair over several handles / GPUs ----------------------------------------------
    def comm_init_nccl(self, rank, nranks, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._check(self._L.ofri_comm_init_nccl(self._h, int(rank), int(nranks), buf))

    def comm_init_local(self, group, rank):
        self._check(self._L.ofri_comm_init_local(self._h, group, int(rank)))

    def comm_destroy(self):
        self._check(self._L.ofri_comm_destroy(self._h))

    def band_plan(self, H, W, params, rank, nranks):
        b = Band()
        self._check(self._L.ofri_band_plan(self._h, int(H), int(W), C.byref(params), int(rank), int(nranks), C.byref(b)))
        return b

    def pyramidal_flow_banded_ptr(self, im1_rows_ptr, im2_rows_ptr, H, W, params, u_rows_ptr, v_rows_ptr, err_ptr=None):
        """DEVICE pointers: rows [in0, in1) of the frames in, rows [own0, own1) of the flow out (see band_plan).
        Collective over the handle's communicator."""
        self._check(self._L.ofri_pyramidal_flow_banded_dev(self._h, im1_rows_ptr, im2_rows_ptr, int(H), int(W),
                                                           C.byref(params), u_rows_ptr, v_rows_ptr, err_ptr))

    # -- adapters ------------------------------------------------------------------------------------------------
    def hs_compute(self, im1, im2, U0, V0, alpha, niter):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        B, H, W = a.shape
        u0 = _batched(U0)[0] if U0 is not None else None
        v0 = _batched(V0)[0] if V0 is not None else None
        _same_shape(a, im2=b, U0=u0, V0=v0)
        U = np.empty((B, H, W), np.float32)
        V = np.empty((B, H, W), np.float32)
        err = np.empty(B, np.float32)
        self._check(self._L.ofri_hs_compute(self._h, _ptr(a), _ptr(b), _ptr(u0) if u0 is not None else None,
                                            _ptr(v0) if v0 is not None else None, B, H, W, float(np.float32(alpha)),
                                            int(niter), _ptr(U), _ptr(V), _ptr(err)))
        return (U[0], V[0], float(err[0])) if single else (U, V, err)

    def ls_compute(self, im1, im2, U0, V0, h, maxiter=60, tol=1e-8):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        B, H, W = a.shape
        u0 = _batched(U0)[0] if U0 is not None else None
        v0 = _batched(V0)[0] if V0 is not None else None
        _same_shape(a, im2=b, U0=u0, V0=v0)
        U = np.empty((B, H, W), np.float32)
        V = np.empty((B, H, W), np.float32)
        err = np.empty(B, np.float32)
        it = np.empty(B, np.int32)
        self._check(self._L.ofri_ls_compute(self._h, _ptr(a), _ptr(b), _ptr(u0) if u0 is not None else None,
                                            _ptr(v0) if v0 is not None else None, B, H, W, float(np.float32(h)),
                                            int(maxiter), float(tol), _ptr(U), _ptr(V), _ptr(err),
                                            it.ctypes.data_as(C.POINTER(C.c_int32))))
        return (U[0], V[0], float(err[0]), int(it[0])) if single else (U, V, err, it)

    # -- stages ---------------------------------------------------------------------------------------------------
    def gauss_px(self, img, taps):
        a, single = _batched(img)
        B, H, W = a.shape
        t = _f32(taps)
        out = np.empty_like(a)
        self._check(self._L.ofri_gauss_px(self._h, _ptr(a), B, H, W, _ptr(t), len(t), _ptr(out)))
        return out[0] if single else out

    def resize_bicubic(self, img, out_h, out_w):
        a, single = _batched(img)
        B, H, W = a.shape
        out = np.empty((B, out_h, out_w), np.float32)
        self._check(self._L.ofri_resize_bicubic(self._h, _ptr(a), B, H, W, int(out_h), int(out_w), _ptr(out)))
        return out[0] if single else out

    def resize_bilinear(self, img, out_h, out_w):
        """Pillow-BILINEAR resample (Farneback_PyCL.py:61-62), both directions."""
        a, single = _batched(img)
        B, H, W = a.shape
        out = np.empty((B, out_h, out_w), np.float32)
        self._check(self._L.ofri_resize_bilinear(self._h, _ptr(a), B, H, W, int(out_h), int(out_w), _ptr(out)))
        return out[0] if single else out

    def set_lk(self, lk_params_):
        self._check(self._L.ofri_set_lk(self._h, C.byref(lk_params_)))

    def lk_compute(self, im1, im2, U0, V0, lk_params_):
        """denseLucasKanade_PyCl.compute (LK:113-169): (U, V) initial flow in, refined flow out."""
        a, single = _batched(im1)
        b, _ = _batched(im2)
        u0 = _batched(U0)[0] if U0 is not None else None
        v0 = _batched(V0)[0] if V0 is not None else None
        _same_shape(a, im2=b, U0=u0, V0=v0)
        B, H, W = a.shape
        U = np.empty((B, H, W), np.float32)
        V = np.empty((B, H, W), np.float32)
        self._check(self._L.ofri_lk_compute(self._h, _ptr(a), _ptr(b), _ptr(u0) if u0 is not None else None,
                                            _ptr(v0) if v0 is not None else None, B, H, W, C.byref(lk_params_),
                                            _ptr(U), _ptr(V)))
        return (U[0], V[0]) if single else (U, V)

    def set_farneback(self, fb_params):
        self._check(self._L.ofri_set_farneback(self._h, C.byref(fb_params)))

    def farneback_compute(self, im1, im2, U0, V0, fb_params):
        """Farneback_PyCL.compute (FB:462-604): (U, V) initial flow in, refined flow out."""
        a, single = _batched(im1)
        b, _ = _batched(im2)
        u0 = _batched(U0)[0] if U0 is not None else None
        v0 = _batched(V0)[0] if V0 is not None else None
        _same_shape(a, im2=b, U0=u0, V0=v0)
        B, H, W = a.shape
        U = np.empty((B, H, W), np.float32)
        V = np.empty((B, H, W), np.float32)
        self._check(self._L.ofri_farneback_compute(self._h, _ptr(a), _ptr(b), _ptr(u0) if u0 is not None else None,
                                                   _ptr(v0) if v0 is not None else None, B, H, W, C.byref(fb_params),
                                                   _ptr(U), _ptr(V)))
        return (U[0], V[0]) if single else (U, V)

    def level_size(self, n, scale):
        return int(self._L.ofri_level_size(int(n), float(scale)))

    def spline_upsample(self, a, out_h, out_w, mul=1.0):
        a, single = _batched(a)
        B, h, w = a.shape
        out = np.empty((B, out_h, out_w), np.float32)
        self._check(self._L.ofri_spline_upsample(self._h, _ptr(a), B, h, w, int(out_h), int(out_w),
                                                 float(np.float32(mul)), _ptr(out)))
        return out[0] if single else out

    def warp_bilinear(self, img, cy, cx):
        a, single = _batched(img)
        y, _ = _batched(cy)
        x, _ = _batched(cx)
        _same_shape(a, cy=y, cx=x)
        B, H, W = a.shape
        out = np.empty_like(a)
        self._check(self._L.ofri_warp_bilinear(self._h, _ptr(a), _ptr(y), _ptr(x), B, H, W, _ptr(out)))
        return out[0] if single else out

    def warp_pair(self, im1, im2, us, vs):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        u, _ = _batched(us)
        v, _ = _batched(vs)
        _same_shape(a, im2=b, us=u, vs=v)
        B, H, W = a.shape
        o1 = np.empty_like(a)
        o2 = np.empty_like(a)
        self._check(self._L.ofri_warp_pair(self._h, _ptr(a), _ptr(b), _ptr(u), _ptr(v), B, H, W, _ptr(o1), _ptr(o2)))
        return (o1[0], o2[0]) if single else (o1, o2)

    def liu_shen_warp(self, im1, us, vs):
        """The biLinear=False warp of frame 1 (reference GPOF:190-196, 204-221); IndexError like the reference."""
        a, single = _batched(im1)
        u, _ = _batched(us)
        v, _ = _batched(vs)
        _same_shape(a, us=u, vs=v)
        B, H, W = a.shape
        mask_size = 3
        sg, tr = 0.6 * mask_size, 4.0 / 0.6 * mask_size
        t = gaussian_taps(sg, 2 * int(tr * sg + 0.5) + 1)
        out = np.empty_like(a)
        self._check(self._L.ofri_liu_shen_warp(self._h, _ptr(a), _ptr(u), _ptr(v), B, H, W, _ptr(t), len(t), _ptr(out)))
        return out[0] if single else out

    def hs_derivatives(self, im1, im2):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        _same_shape(a, im2=b)
        B, H, W = a.shape
        fx, fy, ft = np.empty_like(a), np.empty_like(a), np.empty_like(a)
        self._check(self._L.ofri_hs_derivatives(self._h, _ptr(a), _ptr(b), B, H, W, _ptr(fx), _ptr(fy), _ptr(ft)))
        return (fx[0], fy[0], ft[0]) if single else (fx, fy, ft)

    def hs_iterate(self, U0, V0, fx, fy, ft, alpha, niter):
        u, single = _batched(U0)
        v, _ = _batched(V0)
        dx, _ = _batched(fx)
        dy, _ = _batched(fy)
        dt, _ = _batched(ft)
        _same_shape(u, V0=v, fx=dx, fy=dy, ft=dt)
        B, H, W = u.shape
        U, V = np.empty_like(u), np.empty_like(u)
        self._check(self._L.ofri_hs_iterate(self._h, _ptr(u), _ptr(v), _ptr(dx), _ptr(dy), _ptr(dt), B, H, W,
                                            float(np.float32(alpha)), int(niter), _ptr(U), _ptr(V)))
        return (U[0], V[0]) if single else (U, V)

    def ls_coefficients(self, im1, im2, h):
        a, single = _batched(im1)
        b, _ = _batched(im2)
        _same_shape(a, im2=b)
        B, H, W = a.shape
        coef = np.empty((8, B, H, W), np.float32)
        self._check(self._L.ofri_ls_coefficients(self._h, _ptr(a), _ptr(b), B, H, W, float(np.float32(h)), _ptr(coef)))
        return coef[:, 0] if single else coef


def band_plan_host(H, W, params, rank, nranks, hs_fuse=4, band_exchange=0, band_reach=0):
    """Row-band plan without a GPU (ofri_band_plan_host): which rows a rank owns and which input rows it needs."""
    b = Band()
    L = _lib.lib()
    rc = L.ofri_band_plan_host(int(H), int(W), C.byref(params), int(rank), int(nranks), int(hs_fuse), int(band_exchange),
                               int(band_reach), C.byref(b))
    if rc != 0:
        msg = L.ofri_last_error(None).decode()
        raise _EXC.get(rc, OfriError)(msg) if rc in _EXC else OfriError(rc, msg)
    return b


def nccl_unique_id():
    """128 opaque bytes from ncclGetUniqueId (rank 0 creates them and ships them to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = _lib.lib().ofri_nccl_unique_id(buf)
    if rc != 0:
        raise OfriError(rc, _lib.lib().ofri_last_error(None).decode())
    return buf.raw


class LocalGroup:
    """Rendezvous object of the single-process 'local' communicator: N bands driven by N host threads."""

    def __init__(self, nranks):
        self._L = _lib.lib()
        self.ptr = C.c_void_p()
        rc = self._L.ofri_local_group_create(int(nranks), C.byref(self.ptr))
        if rc != 0:
            raise OfriError(rc, self._L.ofri_last_error(None).decode())
        self.nranks = int(nranks)

    def abort(self):
        """A rank failed: wake the ranks blocked in a collective; their calls return an error instead of hanging."""
        if self.ptr:
            self._L.ofri_local_group_abort(self.ptr)

    def close(self):
        if self.ptr:
            self._L.ofri_local_group_destroy(self.ptr)
            self.ptr = None


_default = {}


def default_handle(device=0):
    """Process-wide handle per device (created on first use; raises without a B200)."""
    h = _default.get(device)
    if h is None:
        h = _default[device] = Handle(device)
    return h
