"""Drop-in for the reference's src/PhysicsBasedOpticalFlowLiuShen.py, computing on the B200 through libofri.so.
Mirrors PhysicsBasedOpticalFlowLiuShen.py:33-45 (adapter) and 82-158 (solver: max-normalisation, coefficient
planes, up to 60 sweeps with the 1e-8 stopping rule)."""
import _native

MAXNUM = 60      # reference :88
TOL = 1e-8       # reference :89


class LiuShenOpticalFlowAlgoAdapter(object):
    _ofri_native_kind = "LS"

    def __init__(self, alpha):
        self.alpha = alpha

    def compute(self, im1, im2, U, V):
        resU, resV, error, _ = _native.handle().ls_compute(im1, im2, U, V, self.alpha, MAXNUM, TOL)
        return [resU, resV, error]

    def getAlgoName(self):
        return 'Liu-Shen Physics based OF'

    def hasGenericPyramidalDefaults(self):
        return False


def physicsBasedOpticalFlowLiuShen(im1, im2, h, U, V):
    """Reference signature: here U is the ROW component and V the COLUMN component (the adapter swaps them)."""
    resCol, resRow, error, _ = _native.handle().ls_compute(im1, im2, V, U, h, MAXNUM, TOL)
    return resRow, resCol, error
