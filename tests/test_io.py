"""lqg_b200.io.load_tracking_data against the fixture made by the reference's own loader (lqg/io.py) -- runs only where the
reference's data file is present (this container); skipped on the GPU box."""
import os

import numpy as np
import pytest

from lqg_b200 import io
from tests import helpers as H

DATA_DIR = "/root/reference/data"


@pytest.mark.skipif(not os.path.exists(os.path.join(DATA_DIR, "data.mat")), reason="reference data file not available")
def test_loader_matches_reference_loader_output():
    data, sigmas = io.load_tracking_data(delay=12, clip=120, data_path=DATA_DIR)
    assert data.shape == (6, 20, 1068, 2) and data.dtype == np.float32 and len(sigmas) == 6
    z = np.load(os.path.join(H.ROOT, "tests", "golden", "ref_c2r_bounded_realdata_T1067.npz"))
    assert np.array_equal(data[0], z["X"])                       # bit-identical to lqg/io.py's first condition
    assert np.all(data[:, :, 0, 0] == 0.0)
