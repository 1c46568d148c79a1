"""Drop-in for the reference's src/GenericPyramidalOpticalFlow.py: same public functions and argument order, the
whole coarse-to-fine loop running on the B200.

  genericPyramidalOpticalFlow(im1, im2, FILTER, mainOFlowAlgoAdapter, pyramidalLevels=1, kLevels=1, FILTER_OPT=None,
                              optionalOFlowAlgoAdapter=None, warping=True, biLinear=True,
                              pyramidalIntermediateScaling=True, pyramidalScaling=False) -> (U, V)

When the adapters are this package's HSOpticalFlowAlgoAdapter / LiuShenOpticalFlowAlgoAdapter the entire pyramid
(down-sample, spline up-sample, warp, pre-filters, HS sweeps, Liu-Shen sweeps, accumulation) is ONE native call
(ofri_pyramidal_flow) and the data never leaves the GPU between stages.  Any other duck-typed adapter
(compute / getAlgoName / hasGenericPyramidalDefaults / getGenericPyramidalDefaults, reference :256-290) still works:
the driver, the level stages and the native adapter of the pair (e.g. the Liu-Shen refinement behind a foreign main
adapter) stay on the GPU, and only the foreign adapter's own compute() is called on the host with numpy arrays
(ofri_pyramidal_flow_external: one D2H of the level's frames, one H2D of the adapter's result per call).

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
    """ofri_algo for one of this package's adapters.  HS takes `ncalls` alphas from the END of the caller's list in the
    order the reference's per-call pop would (HornSchunck.py:36) -- WITHOUT consuming them: _consume_alphas() removes
    them only after the native call has succeeded (the reference consumes one alpha per compute() that actually runs)."""
    kind = getattr(adapter, '_ofri_native_kind', None)
    if kind == 'HS':
        if len(adapter.alphas) < ncalls:
            raise IndexError('pop from empty list')  # the reference's failure mode, raised before anything ran
        used = list(adapter.alphas[len(adapter.alphas) - ncalls:])[::-1]
        return ofri.hs_algo(used, adapter.Niter)
    if kind == 'LS':
        return ofri.ls_algo(adapter.alpha)
    if kind == 'FB':
        return ofri.fb_algo()          # its parameters travel separately (Handle.set_farneback in _run)
    if kind == 'LK':
        return ofri.lk_algo()          # likewise (Handle.set_lk)
    return None


def _consume_alphas(adapter, ncalls):
    if getattr(adapter, '_ofri_native_kind', None) == 'HS':
        del adapter.alphas[len(adapter.alphas) - ncalls:]


def _is_native(adapter):
    return getattr(adapter, '_ofri_native_kind', None) in ('HS', 'LS', 'FB', 'LK')


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
    ncalls = pyramidalLevels * kLevels
    native_all = _is_native(main) and (optional is None or _is_native(optional))
    if not native_all and np.ndim(im1) != 2:
        raise ValueError("batched input needs this package's HS / Liu-Shen adapters")

    fb = [a for a in (main, optional) if getattr(a, '_ofri_native_kind', None) == 'FB']
    if len(fb) == 2 and fb[0] is not fb[1]:
        raise NotImplementedError('two different Farneback adapters in one pair: the handle holds one parameter set')
    if fb:
        _native.handle().set_farneback(fb[0].native_params())

    lk = [a for a in (main, optional) if getattr(a, '_ofri_native_kind', None) == 'LK']
    if len(lk) == 2 and lk[0] is not lk[1]:
        raise NotImplementedError('two different Lucas-Kanade adapters in one pair: the handle holds one parameter set')
    if lk:
        _native.handle().set_lk(lk[0].native_params())

    def algo_of(adapter):
        return _native_algo(adapter, ncalls) if _is_native(adapter) else ofri.external_algo()

    params = ofri.make_params(algo_of(main), algo_of(optional) if optional is not None else None,
                              filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT, pyramid_levels=pyramidalLevels,
                              k_levels=kLevels, warping=warping, bilinear=biLinear,
                              intermediate_scaling=interScaling, final_scaling=finalScaling)
    if native_all:
        out = _native.handle().pyramidal_flow(im1, im2, params)
    else:
        # Foreign adapters (any object with the reference's duck-typed protocol, e.g. a dense Lucas-Kanade or Farneback
        # main adapter refined by this package's Liu-Shen: examples/LiuSE_denseLK_Fs2_0_PyrLvls2.py:68-74): the whole
        # driver still runs natively -- level stages, the native adapter of the pair (if any) and the accumulation stay
        # on the GPU; only the foreign adapter's own compute() runs on the host, fed with numpy copies of the level's
        # frames and the current flow (ofri_pyramidal_flow_external).
        def wrap(adapter):
            def compute(i1, i2, U, V):
                res = adapter.compute(i1, i2, U, V)
                log(adapter.getAlgoName() + ' estimated error for image registration: ' + str(res[2]))
                return res
            return compute
        out = _native.handle().pyramidal_flow_external(
            im1, im2, params, compute_main=None if _is_native(main) else wrap(main),
            compute_optional=None if (optional is None or _is_native(optional)) else wrap(optional))
    _consume_alphas(main, ncalls)
    if optional is not None:
        _consume_alphas(optional, ncalls)
    return out


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
