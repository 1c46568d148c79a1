"""CPU-side checks of the product boundary: libofri.so builds, loads, exports every symbol include/ofri.h declares,
the ctypes structures match the C layout, and -- with no GPU in this container -- the library refuses to run instead
of falling back to the CPU.  No compute calls here."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ofri():
    import opticalflow_ri_b200 as o
    from opticalflow_ri_b200 import build
    build.build()
    return o


def test_exports_every_declared_symbol(ofri):
    L = ofri.lib()
    names = ofri.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", ofri._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert set(names) == set(ofri._lib._SIGNATURES)


def test_struct_layout_matches_c(ofri, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ofri.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu",'
                   'sizeof(ofri_params),sizeof(ofri_algo),offsetof(ofri_params,taps_opt),offsetof(ofri_params,main_algo),'
                   'offsetof(ofri_algo,ls_tol),sizeof(ofri_farneback_params),offsetof(ofri_farneback_params,win_kernel),'
                   'offsetof(ofri_farneback_params,blur_kernel),sizeof(ofri_lk_params),offsetof(ofri_lk_params,asym));'
                   'return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    P, A, F, K = ofri.Params, ofri.Algo, ofri._lib.FarnebackParams, ofri._lib.LkParams
    assert got == [C.sizeof(P), C.sizeof(A), P.taps_opt.offset, P.main_algo.offset, A.ls_tol.offset, C.sizeof(F),
                   F.win_kernel.offset, F.blur_kernel.offset, C.sizeof(K), K.asym.offset]


def test_host_helpers_without_gpu(ofri):
    L = ofri.lib()
    assert L.ofri_abi_version() == 1
    assert L.ofri_level_size(61, 0.5) == 30 and L.ofri_level_size(47, 0.5) == 24 and L.ofri_level_size(151, 0.5) == 76
    k = np.zeros(3, np.float32)
    assert L.ofri_gaussian_taps(3.4, 3, k.ctypes.data_as(C.POINTER(C.c_float))) == 0
    assert [hex(x) for x in k.view(np.uint32)] == ["0x3ea83048", "0x3eaf9f71", "0x3ea83048"]
    k5 = ofri.gaussian_taps(0.48, 5)
    assert [hex(x) for x in k5.view(np.uint32)] == ["0x3910f5e4", "0x3dbe4a71", "0x3f505b46", "0x3dbe4a71", "0x3910f5e4"]


def test_no_cpu_fallback(ofri):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert ofri.lib().ofri_device_count() < 0
    with pytest.raises(ofri.OfriError) as e:
        ofri.Handle(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_dropin_modules_import_and_mirror_reference_api(ofri):
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import GenericPyramidalOpticalFlow as G
        import GenericPyramidalOpticalFlowWrapper as Wr
        import HornSchunck as HS
        import PhysicsBasedOpticalFlowLiuShen as LS
        import gaussian_filter as GF
        import GaussianKernelBitExact as GK
    finally:
        sys.path.remove(ofri.SRC_DIR)
    import inspect
    sig = inspect.signature(G.genericPyramidalOpticalFlow)
    assert list(sig.parameters) == ["im1", "im2", "FILTER", "mainOFlowAlgoAdapter", "pyramidalLevels", "kLevels",
                                    "FILTER_OPT", "optionalOFlowAlgoAdapter", "warping", "biLinear",
                                    "pyramidalIntermediateScaling", "pyramidalScaling"]
    assert sig.parameters["pyramidalScaling"].default is False and sig.parameters["FILTER_OPT"].default is None
    assert list(inspect.signature(Wr.GenericPyramidalOpticalFlowWrapper.__init__).parameters)[1:] == [
        "algo_adapter", "filter_sigma", "pyr_levels", "k_levels", "filter_opt", "optional_algo_adapter", "warping",
        "bi_linear", "pyramidal_intermediate_scaling", "pyramidal_scaling"]
    hs = HS.HSOpticalFlowAlgoAdapter([21, 45], 600)
    assert hs.getAlgoName() == "Horn-Schunck" and hs.hasGenericPyramidalDefaults()
    assert hs.getGenericPyramidalDefaults() == {"warping": True, "biLinear": True, "scaling": True}
    ls = LS.LiuShenOpticalFlowAlgoAdapter(5)
    assert ls.getAlgoName() == "Liu-Shen Physics based OF" and not ls.hasGenericPyramidalDefaults()
    assert np.array_equal(GF.prepareGaussianKernel(3.4, 3).view(np.uint32), [0x3ea83048, 0x3eaf9f71, 0x3ea83048])
    s, k = GK.getGaussianKernelBitExact(5, 0.48)
    assert k.astype(">f8").tobytes()[:8].hex() == "3fb21dbeb2868cad"
    # argument errors surface before any GPU work, with the reference's exception types
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(TypeError):
        G.genericPyramidalOpticalFlow(z, z, 3.4, hs, 2, 1, None, ls)
    with pytest.raises(IndexError):
        G.genericPyramidalOpticalFlow(z, z, 3.4, HS.HSOpticalFlowAlgoAdapter([21], 10), 2, 1)
    with pytest.raises(Exception, match="Invalid scale level"):
        G.genericPyramidalOpticalFlow(z, z, 3.4, LS.LiuShenOpticalFlowAlgoAdapter(5), 0, 1)


def test_gkbe_dropin_matches_golden(ofri, stages):
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import GaussianKernelBitExact as GK
    finally:
        sys.path.remove(ofri.SRC_DIR)
    for tag, (n, sg) in {"3_0": (3, 0.0), "5_0": (5, 0.0), "3_34": (3, 3.4), "5_048": (5, 0.48), "7_15": (7, 1.5),
                         "33_495": (33, 4.95), "4_1": (4, 1.0), "5_m12": (5, -1.2), "11_0": (11, 0.0)}.items():
        s, k = GK.getGaussianKernelBitExact(n, sg)
        assert np.array_equal(np.asarray(k, dtype=np.float64), stages["gkbe_k_" + tag])
        assert float(s) == float(stages["gkbe_sum_" + tag])
