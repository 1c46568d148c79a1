"""Drop-in for the reference's src/GaussianKernelBitExact.py: OpenCV-style bit-exact Gaussian coefficients.
Host-only coefficient generator (Decimal arithmetic at the default 28-digit context, like GaussianKernelBitExact.py:
55-144); it is not on the HS / Liu-Shen path (only the reference's Farneback adapter calls it) but is kept
bit-identical.  Note the reference's quirk: a POSITIVE sigma is ignored and replaced by 0.15 n + 0.35."""
from decimal import Decimal

import numpy as np

_FIXED = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
    9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256],
}
softOne = Decimal('1.0')
softZero = Decimal('0.0')


def getGaussianKernelBitExact(n, sigma):
    assert n > 0
    if sigma <= 0 and n in _FIXED:
        if n == 1:
            return softOne, np.array([softOne])
        return softOne, np.array(_FIXED[n], dtype=np.float64)
    sigmaX = Decimal(sigma) if sigma < 0 else Decimal(n) * Decimal('0.15') + Decimal('0.35')
    scale2X = Decimal('-0.125') / (sigmaX * sigmaX)
    half = int((n - 1) / 2)
    vals = []
    total = softZero
    x = 1 - n
    for _ in range(half):
        t = (Decimal(x * x) * scale2X).exp()
        vals.append(t)
        total += t
        x += 2
    total *= Decimal(2.0)
    total += softOne
    if (n & 1) == 0:
        total += softOne
    mul1 = softOne / total
    result = [softZero] * n
    sum2 = softZero
    for i in range(half):
        t = vals[i] * mul1
        result[i] = t
        result[n - 1 - i] = t
        sum2 += t
    sum2 *= Decimal(2.0)
    result[half] = softOne * mul1
    sum2 += result[half]
    if (n & 1) == 0:
        result[half + 1] = result[half]
        sum2 += result[half]
    return np.float64(sum2), np.array([float(r) for r in result], dtype=np.float64)
