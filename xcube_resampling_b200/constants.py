"""Constants and option vocab of the resampling path.

Mirrors ``xcube_resampling/constants.py`` of the reference (values and names
only; the aggregation table maps to device kernel ids instead of numpy
callables).
"""

import logging

import numpy as np

# constants.py:79-82 of the reference
SCALE_LIMIT = 0.95
UV_DELTA = 1e-3
LOG = logging.getLogger("xcube.resampling")

# constants.py:74-77
FILLVALUE_UINT8 = 255
FILLVALUE_UINT16 = 65535
FILLVALUE_INT = -1
FILLVALUE_FLOAT = np.nan

# constants.py:66-70
INTERP_METHOD_MAPPING = {0: "nearest", 1: "bilinear", "nearest": 0, "bilinear": 1}

# device ids, include/xrs.h enum xrs_interp
INTERP_CODES = {"nearest": 0, "bilinear": 1, "triangular": 2}

# constants.py:34-65 -> include/xrs.h enum xrs_agg
AGG_CODES = {
    "center": 0, "count": 1, "first": 2, "last": 3, "max": 4, "mean": 5, "median": 6,
    "mode": 7, "min": 8, "prod": 9, "std": 10, "sum": 11, "var": 12,
}
AGG_METHODS = tuple(AGG_CODES)

# include/xrs.h enum xrs_dtype
DTYPE_CODES = {
    np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.uint8): 2, np.dtype(np.int8): 3,
    np.dtype(np.uint16): 4, np.dtype(np.int16): 5, np.dtype(np.int32): 6, np.dtype(np.uint32): 7,
    np.dtype(np.int64): 8,
}
