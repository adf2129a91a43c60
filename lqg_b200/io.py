"""Tracking-data loader with the contract of ``lqg/io.py:45-98`` (``load_tracking_data``): the Bonnen et al. (2015) target /
response traces stored in the reference repository's ``data/data.mat``, grouped by blob width.

Returns ``(data, sigmas)`` with ``data[condition, trial, time, (target, response)]`` float32 -- the layout
``System.log_likelihood`` expects per condition -- after the reference's preprocessing: response shifted by ``delay`` samples
against the target, the first ``clip`` samples dropped, per-trial means removed, and every trial re-referenced to its first
target sample.  (``tests/golden/ref_c2r_bounded_realdata_T1067.npz`` holds the first condition as produced by the reference's
own loader; ``tests/test_io.py`` checks this function against it when the .mat file is available.)"""
import os

import numpy as np

ARCMIN_PER_PIXEL = 1.32   # lqg/io.py:58


def load_tracking_data(delay=12, clip=120, subtract_mean=True, data_path="data/"):
    import scipy.io as spio
    mat = spio.loadmat(os.path.join(data_path, "data.mat"), struct_as_record=False, squeeze_me=True)
    width = np.round(np.asarray(mat["sigma"], dtype=np.float64) * ARCMIN_PER_PIXEL)
    target = np.asarray(mat["target"], dtype=np.float32)
    response = np.asarray(mat["response"], dtype=np.float32)
    stop = -delay if delay else None
    target, response = target[:, clip:stop], response[:, clip + delay:]
    if subtract_mean:
        target = target - target.mean(axis=1, keepdims=True)
        response = response - response.mean(axis=1, keepdims=True)
    sigmas = np.unique(width)
    per_condition = []
    for w in sigmas:
        rows = np.flatnonzero(width == w)
        per_condition.append(np.stack([target[rows], response[rows]], axis=-1))      # [trial, time, 2]
    data = np.stack(per_condition)                                                    # [condition, trial, time, 2]
    data = data - data[:, :, :1, :1]                                                  # first target sample of every trial -> 0
    return data, sigmas
