"""Drop-in for the reference's src/GenericPyramidalOpticalFlow.py: same public functions and argument order, the
whole coarse-to-fine loop running on the B200.

  genericPyramidalOpticalFlow(im1, im2, FILTER, mainOFlowAlgoAdapter, pyramidalLevels=1, kLevels=1, FILTER_OPT=None,
                              optionalOFlowAlgoAdapter=None, warping=True, biLinear=True,
                              pyramidalIntermediateScaling=True, pyramidalScaling=False) -> (U, V)

When the adapters are this package's HSOpticalFlowAlgoAdapter / LiuShenOpticalFlowAlgoAdapter the entire pyramid
(down-sample, spline up-sample, warp, pre-filters, HS sweeps, Liu-Shen sweeps, accumulation) is ONE native call
(ofri_pyramidal_flow) and the data never leaves the GPU between stages.  Any other duck-typed adapter
(compute / getAlgoName / hasGenericPyramidalDefaults / getGenericPyramidalDefaults, reference :256-290) still works:
the level stages run as GPU kernels and the adapter's own compute() is called with numpy arrays.

Behaviour mirrored from the reference (line numbers of its GenericPyramidalOpticalFlow.py): adapter-default override
304-327; level sizes int32(round(n*scale)) from the ORIGINAL frames 336-343; 'Invalid scale level' 345; level
transition 118-235 (bilinear branch); pre-filters 368-386 (the optional adapter sees the UNWARPED level images);
k-loop with re-warp 389-404; accumulation 413-414; biLinear=False "Liu-Shen warp" 190-196 / 204-221 (frame 1 only;
IndexError when the rounded flow moves a pixel past the right / bottom border, wrap-around at the left / top border).
One deliberate difference: with this package's adapters the caller's im1 array is never modified (the reference's
Liu-Shen warp overwrites it in place at the last pyramid level)."""
import numpy as np

import _native
from _native import log, ofri


# ---- stage-level functions with the reference's names --------------------------------------------------------------
def imresize(im, res):
    """Pillow-BICUBIC resample to res = (width, height) (reference :67-68); down-sampling only."""
    return _native.handle().resize_bicubic(im, int(res[1]), int(res[0]))


def doBiLinearWarping(img, coordsY, coordsX, order=1, mode='nearest'):
    """Bilinear gather at (coordsY, coordsX) with clamp-to-edge (reference :70-116; order / mode are ignored there too)."""
    return _native.handle().warp_bilinear(img, np.float32(coordsY), np.float32(coordsX))


def updateNextPyramidalLevel(im1IterNext, im1IterPrev, im2IterNext, Uaccum, Vaccum, U, V, warping=True, biLinear=True,
                             scale=False):
    """Level transition (reference :118-235): spline up-sample of the accumulated flow to the new level's size,
    optional scaling, then either the symmetric half-flow bilinear warp of both frames or the no-warp hand-over."""
    h = _native.handle()
    Hn, Wn = im1IterNext.shape
    Hp, Wp = im1IterPrev.shape
    if (Hp, Wp) != (Hn, Wn):
        mx = np.float32(Wn) / np.float32(Wp) if scale else np.float32(1)
        my = np.float32(Hn) / np.float32(Hp) if scale else np.float32(1)
        usNew = h.spline_upsample(Uaccum, Hn, Wn, mx)
        vsNew = h.spline_upsample(Vaccum, Hn, Wn, my)
    else:
        usNew, vsNew = np.float32(Uaccum), np.float32(Vaccum)
    zeros = np.zeros((Hn, Wn), dtype=np.float32)
    if not warping:
        return im1IterNext, im2IterNext, zeros, zeros.copy(), usNew, vsNew
    if not biLinear:
        log('Warping: Liu-Shen')
        im1IterNext[...] = h.liu_shen_warp(im1IterNext, usNew, vsNew)      # in place, like the reference (:207-221)
        return im1IterNext, im2IterNext, usNew, vsNew, zeros, zeros.copy()
    log('Warping: BiLinear')
    w1, w2 = h.warp_pair(im1IterNext, im2IterNext, usNew, vsNew)
    return w1, w2, usNew, vsNew, zeros, zeros.copy()


# ---- parameter handling ---------------------------------------------------------------------------------------------
def _apply_adapter_defaults(main, warping, biLinear, interScaling, finalScaling):
    if main.hasGenericPyramidalDefaults():
        d = main.getGenericPyramidalDefaults()
        if d is not None:
            if d.get('warping') is not None:
                warping = d.get('warping')
            if d.get('biLinear') is not None:
                biLinear = d.get('biLinear')
            if d.get('intermediateScaling') is not None:
                interScaling = d.get('intermediateScaling')
            if d.get('scaling') is not None:
                finalScaling = d.get('scaling')
    return warping, biLinear, interScaling, finalScaling


def _native_algo(adapter, ncalls):
    """ofri_algo for one of this package's adapters; HS consumes `ncalls` alphas from the END of the caller's list
    exactly like the per-call pop of the reference (HornSchunck.py:36)."""
    kind = getattr(adapter, '_ofri_native_kind', None)
    if kind == 'HS':
        used = []
        for _ in range(ncalls):
            used.append(adapter.alphas.pop())       # IndexError here == the reference's failure mode
        return ofri.hs_algo(used, adapter.Niter)
    if kind == 'LS':
        return ofri.ls_algo(adapter.alpha)
    return None


def _is_native(adapter):
    return getattr(adapter, '_ofri_native_kind', None) in ('HS', 'LS')


def _run(im1, im2, FILTER, main, pyramidalLevels, kLevels, FILTER_OPT, optional, warping, biLinear, interScaling,
         finalScaling):
    warping, biLinear, interScaling, finalScaling = _apply_adapter_defaults(main, warping, biLinear, interScaling,
                                                                            finalScaling)
    pyramidalLevels = int(pyramidalLevels)
    kLevels = int(kLevels)
    if pyramidalLevels < 1:
        raise Exception('Invalid scale level: ' + str(1.0 / (2.0 ** (pyramidalLevels - 1))))
    if optional is not None and FILTER_OPT is None:
        raise TypeError("'>' not supported between instances of 'NoneType' and 'float'")
    if _is_native(main) and (optional is None or _is_native(optional)):
        ncalls = pyramidalLevels * kLevels
        params = ofri.make_params(_native_algo(main, ncalls),
                                  _native_algo(optional, ncalls) if optional is not None else None,
                                  filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT, pyramid_levels=pyramidalLevels,
                                  k_levels=kLevels, warping=warping, bilinear=biLinear,
                                  intermediate_scaling=interScaling, final_scaling=finalScaling)
        return _native.handle().pyramidal_flow(im1, im2, params)
    return _generic_adapters(im1, im2, FILTER, main, pyramidalLevels, kLevels, FILTER_OPT, optional, warping, biLinear,
                             interScaling, finalScaling)


def _generic_adapters(im1, im2, FILTER, main, L, KL, FILTER_OPT, optional, warping, biLinear, interScaling,
                      finalScaling):
    """Foreign adapters: GPU stages + the adapter's own compute() on numpy arrays (one pair at a time)."""
    if np.ndim(im1) != 2:
        raise ValueError("batched input needs this package's HS / Liu-Shen adapters")
    h = _native.handle()
    im1 = np.ascontiguousarray(im1, dtype=np.float32)
    im2 = np.ascontiguousarray(im2, dtype=np.float32)
    H, W = im1.shape
    taps_main = ofri.gaussian_taps(FILTER, 3) if FILTER > 1e-3 else None
    taps_opt = ofri.gaussian_taps(FILTER_OPT, 5) if (optional is not None and FILTER_OPT > 1e-3) else None
    scale = 1.0 / (2.0 ** (L - 1))
    Uacc = Vacc = U = V = None
    prev = None
    for level in range(1, L + 1):
        last = level == L
        local_scaling = finalScaling if last else interScaling
        if scale < 1.0 and not last:
            hl, wl = h.level_size(H, scale), h.level_size(W, scale)
            n1, n2 = h.resize_bicubic(im1, hl, wl), h.resize_bicubic(im2, hl, wl)
        elif scale > 1.0:
            raise Exception('Invalid scale level: ' + str(scale))
        else:
            n1, n2 = im1, im2
        if level > 1:
            w1, w2, Uacc, Vacc, U, V = updateNextPyramidalLevel(n1, prev, n2, Uacc, Vacc, U, V, warping, biLinear,
                                                                local_scaling)
        else:
            w1, w2 = n1, n2
            U, V, Uacc, Vacc = (np.zeros(n1.shape, dtype=np.float32) for _ in range(4))
        work1 = h.gauss_px(w1, taps_main) if taps_main is not None else w1.copy()
        work2 = h.gauss_px(w2, taps_main) if taps_main is not None else w2
        if optional is not None:
            opt1 = h.gauss_px(n1, taps_opt) if taps_opt is not None else n1.copy()
            opt2 = h.gauss_px(n2, taps_opt) if taps_opt is not None else n2
        for k in range(KL):
            log('Level=', level, ' kIter=', k)
            if k > 0:
                if warping:
                    w1, w2, Uacc, Vacc, U, V = updateNextPyramidalLevel(n1.copy(), n1, n2, Uacc, Vacc, U, V, warping,
                                                                        biLinear, False)
                    if FILTER > 1:
                        work1, work2 = h.gauss_px(w1, taps_main), h.gauss_px(w2, taps_main)
                    else:
                        work1, work2 = w1.copy(), w2
                else:
                    work1, work2, Uacc, Vacc, U, V = updateNextPyramidalLevel(work1, work1, work2, Uacc, Vacc, U, V,
                                                                              warping, biLinear, False)
            U, V, err = main.compute(work1, work2, U, V)
            log(main.getAlgoName() + ' estimated error for image registration: ' + str(err))
            if optional is not None:
                U, V, err2 = optional.compute(np.copy(opt1), np.copy(opt2), U, V)
                log(optional.getAlgoName() + ' estimated error for image registration: ' + str(err2))
            Uacc = np.float32(Uacc) + np.float32(U)
            Vacc = np.float32(Vacc) + np.float32(V)
        prev = work1
        scale *= 2
    return Uacc, Vacc


# ---- public entry points ----------------------------------------------------------------------------------------------
def genericPyramidalOpticalFlow(im1, im2, FILTER, mainOFlowAlgoAdapter, pyramidalLevels=1, kLevels=1, FILTER_OPT=None,
                                optionalOFlowAlgoAdapter=None, warping=True, biLinear=True,
                                pyramidalIntermediateScaling=True, pyramidalScaling=False):
    """Coarse-to-fine optical flow of one frame pair; returns the accumulated (U, V) as float32 (H, W) arrays."""
    if np.ndim(im1) != 2 or np.shape(im1) != np.shape(im2):
        raise ValueError("im1 and im2 must be 2-D arrays of the same shape")
    return _run(im1, im2, FILTER, mainOFlowAlgoAdapter, pyramidalLevels, kLevels, FILTER_OPT, optionalOFlowAlgoAdapter,
                warping, biLinear, pyramidalIntermediateScaling, pyramidalScaling)


def genericPyramidalOpticalFlowBatch(im1s, im2s, FILTER, mainOFlowAlgoAdapter, pyramidalLevels=1, kLevels=1,
                                     FILTER_OPT=None, optionalOFlowAlgoAdapter=None, warping=True, biLinear=True,
                                     pyramidalIntermediateScaling=True, pyramidalScaling=False):
    """Same for a (batch, H, W) stack of independent pairs in one native call (new; the reference has no batch API).
    The HS adapter's alpha list is consumed ONCE for the whole batch (one value per level x k)."""
    if np.ndim(im1s) != 3 or np.shape(im1s) != np.shape(im2s):
        raise ValueError("im1s and im2s must be (batch, H, W) arrays of the same shape")
    return _run(im1s, im2s, FILTER, mainOFlowAlgoAdapter, pyramidalLevels, kLevels, FILTER_OPT,
                optionalOFlowAlgoAdapter, warping, biLinear, pyramidalIntermediateScaling, pyramidalScaling)
