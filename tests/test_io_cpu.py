"""Host-side file I/O helpers: frames in (TIFF incl. packbits, 16-bit rescale, .npy), flow out in the
reference's .mat layout (examples/PyHSchunck_Fs3_4.py:35-51).  CPU only."""
import os

import numpy as np
import pytest

from opticalflow_ri_b200.io import read_frame, read_pairs, save_flow


def test_read_frame_formats(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    a = (np.arange(64 * 48) % 251).astype(np.uint8).reshape(64, 48)
    Image.fromarray(a).save(tmp_path / "a.tif", compression="packbits")        # the bundled pair is packbits TIFF
    b = (np.arange(64 * 48) * 7 % 65535).astype(np.uint16).reshape(64, 48)
    Image.fromarray(b).save(tmp_path / "b.tif")
    np.save(tmp_path / "c.npy", a.astype(np.float64))
    A, B, Cc = (read_frame(str(tmp_path / n)) for n in ("a.tif", "b.tif", "c.npy"))
    assert A.dtype == B.dtype == Cc.dtype == np.float32 and A.flags.c_contiguous
    assert np.array_equal(A, a.astype(np.float32)) and np.array_equal(Cc, A)
    assert np.array_equal(B, b.astype(np.float32) / np.float32(65535.0) * np.float32(255.0))   # BOM:134-137
    p1, p2 = read_pairs([str(tmp_path / "a.tif"), str(tmp_path / "c.npy"), str(tmp_path / "a.tif")])
    assert p1.shape == p2.shape == (2, 64, 48)
    with pytest.raises(ValueError):
        read_pairs([str(tmp_path / "a.tif")])


def test_save_flow_layout(tmp_path):
    sio = pytest.importorskip("scipy.io")
    U = np.linspace(-4, 0, 30 * 20, dtype=np.float32).reshape(30, 20)
    save_flow(U, U * 2, str(tmp_path / "f.mat"))
    m = sio.loadmat(str(tmp_path / "f.mat"), squeeze_me=True, struct_as_record=False)
    v, p = m["velocities"], m["parameters"]
    assert np.array_equal(v.u, U) and np.array_equal(v.v, U * 2) and v.iaWidth == 1 and v.iaHeight == 1
    assert (v.margins.top, v.margins.left, v.margins.bottom, v.margins.right) == (0, 0, 0, 0)
    assert (p.imageHeight, p.imageWidth, p.overlapFactor) == (30, 20, 1.0)
    assert os.path.getsize(str(tmp_path / "f.mat")) > 0
