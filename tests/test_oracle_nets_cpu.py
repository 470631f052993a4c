"""CPU suite: network oracles against golden outputs of the UNMODIFIED reference modules (tests/golden/make_*_golden.py, generated in the
build container from /root/reference with the seeded synthetic weights of the package)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["leres_ref_96x128.npz"])
def test_leres_oracle_matches_reference_golden(name):
    from cartoonsegmentation_b200.depth_modules import leres as L
    from oracle import leres_oracle as O
    gold = np.load(os.path.join(GOLD, name))
    m = O.RelDepthModel().eval()
    missing, unexpected = m.load_state_dict(L.synthetic_state_dict(0), strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing)
    with torch.no_grad():
        out = m.depth_model(O.preprocess(gold['image']))[0, 0].numpy()
    np.testing.assert_allclose(out, gold['depth'], rtol=1e-4, atol=1e-3)
