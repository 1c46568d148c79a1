"""File I/O either side of the path: PIV frames in, flow fields out -- host-side conveniences, nothing here
computes flow.

  read_frame(path)            8/16-bit TIFF / PNG (Pillow) or .npy  ->  float32 (H, W); 16-bit frames are rescaled to
                              0..255 the way benchmark_of_methods.py:134-137 does
  read_pairs(paths)           consecutive frames -> (batch, H, W) stacks of (frame i, frame i+1) for calculateFlowBatch
  save_flow(U, V, filename)   MATLAB .mat with the layout the reference's examples write (examples/PyHSchunck_Fs3_4.py:
                              35-51): velocities{u, v, iaWidth, iaHeight, margins{top,left,bottom,right}},
                              parameters{overlapFactor, imageHeight, imageWidth}
"""
import os

import numpy as np


def read_frame(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path)
    else:
        from PIL import Image
        a = np.array(Image.open(path))
    if a.ndim == 3:                         # colour: luminance like skimage's rgb2gray weights
        a = a[..., :3].astype(np.float64) @ np.array([0.2125, 0.7154, 0.0721])
    if a.dtype == np.uint16:
        a = a.astype(np.float32) / np.float32(65535.0) * np.float32(255.0)
    return np.ascontiguousarray(a, dtype=np.float32)


def read_pairs(paths):
    frames = [read_frame(p) for p in paths]
    if len(frames) < 2:
        raise ValueError("need at least two frames")
    a = np.stack(frames[:-1])
    b = np.stack(frames[1:])
    return a, b


def save_flow(U, V, filename):
    import scipy.io
    U = np.asarray(U)
    V = np.asarray(V)
    velocities = {"u": U, "v": V, "iaWidth": 1, "iaHeight": 1,
                  "margins": {"top": 0, "left": 0, "bottom": 0, "right": 0}}
    parameters = {"overlapFactor": 1.0, "imageHeight": U.shape[-2], "imageWidth": U.shape[-1]}
    scipy.io.savemat(filename, mdict={"velocities": velocities, "parameters": parameters})
