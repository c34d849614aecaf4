"""Structured live-point dtype surface (kept identical to the reference).

Mirror of /root/reference/src/nessai/livepoint.py:74-157 and the defaults in
/root/reference/src/nessai/config.py:21-114: a live point is
``[(name, 'f8')..., ('logP', 'f8'), ('logL', 'f8'), ('it', 'i4')]``; parameters
default to NaN, ``it`` to 0.  When nessai is importable its own (possibly
user-extended) configuration is used instead, so extra non-sampling parameters
are honoured.
"""

from __future__ import annotations

import numpy as np

NON_SAMPLING_PARAMETERS = ["logP", "logL", "it"]
NON_SAMPLING_DTYPE = ["f8", "f8", "i4"]
NON_SAMPLING_DEFAULTS = (np.nan, np.nan, 0)
DEFAULT_FLOAT_DTYPE = "f8"
DEFAULT_FLOAT_VALUE = np.nan


def get_dtype(names, array_dtype=None, non_sampling_parameters=True):
    if array_dtype is None:
        array_dtype = DEFAULT_FLOAT_DTYPE
    dtype = [(n, array_dtype) for n in names]
    if non_sampling_parameters:
        dtype += list(zip(NON_SAMPLING_PARAMETERS, NON_SAMPLING_DTYPE))
    return np.dtype(dtype)


def empty_structured_array(n, names=None, dtype=None, non_sampling_parameters=True):
    if dtype is None:
        dtype = get_dtype(names, non_sampling_parameters=non_sampling_parameters)
    else:
        dtype = np.dtype(dtype)
        names = [nm for nm in dtype.names if nm not in NON_SAMPLING_PARAMETERS]
    arr = np.empty((n), dtype=dtype)
    if n == 0:
        return arr
    for nm in names:
        if dtype.fields[nm][0].kind == "f":
            arr[nm] = DEFAULT_FLOAT_VALUE
        else:
            arr[nm] = 0
    if non_sampling_parameters:
        for nm, v in zip(NON_SAMPLING_PARAMETERS, NON_SAMPLING_DEFAULTS):
            if nm in dtype.names:
                arr[nm] = v
    return arr


def live_points_to_array(live_points, names=None, copy=False):
    """/root/reference/src/nessai/livepoint.py:158-188."""
    if names is None:
        names = [n for n in live_points.dtype.names if n not in NON_SAMPLING_PARAMETERS]
    return np.stack([np.asarray(live_points[n], dtype=np.float64) for n in names], axis=-1)


def numpy_array_to_live_points(array, names):
    """/root/reference/src/nessai/livepoint.py:227-257."""
    array = np.asarray(array)
    if array.size == 0:
        return empty_structured_array(0, names)
    if array.ndim == 1:
        array = array[np.newaxis, :]
    out = empty_structured_array(array.shape[0], names)
    for i, n in enumerate(names):
        out[n] = array[..., i]
    return out
