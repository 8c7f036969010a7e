import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype); self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape; self.dtype = np.dtype(dtype)
