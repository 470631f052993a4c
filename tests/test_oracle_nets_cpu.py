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


def test_zoe_dpt_oracle_vs_transformers_golden():
    """oracle/zoe_dpt_oracle.py (the restated MiDaS DPT_BEiT_L_384 the reference pulls from torch.hub) against the golden produced by transformers'
    independent port, with the same seeded parameters, on the 16 x 20-token case (exercises the bias-table interpolation)."""
    import os
    import sys
    pytest.importorskip("transformers")
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_zoe_dpt_golden as mk
    from oracle import zoe_dpt_oracle as zo
    sd = mk.hf_to_midas(mk.build_hf_model().state_dict())
    Hn, Wn = 256, 320
    with torch.no_grad():
        o = zo.forward(sd, mk.net_input(Hn, Wn))
    g = np.load(os.path.join(gold, f"zoe_dpt_ref_{Hn}x{Wn}.npz"))
    rr = lambda a, b: float(np.sqrt(((np.asarray(a, np.float64) - b) ** 2).mean() / ((np.asarray(b, np.float64) ** 2).mean() + 1e-30)))
    assert rr(o['rel'][0].numpy(), g['rel']) < 1e-5                                             # fp32 both sides
    for k in range(4):                                                                          # goldens stored as fp16
        assert rr(o['tokens'][k][0, ::8].numpy(), g[f'tok{k}'].astype(np.float64)) < 5e-4
        t = o['fused'][k][0].permute(1, 2, 0)
        st = max(1, t.shape[0] // 24)
        assert rr(t[::st, ::st].numpy(), g[f'fused{k}'].astype(np.float64)) < 5e-4
    assert rr(o['btl'][0].permute(1, 2, 0).numpy(), g['btl'].astype(np.float64)) < 5e-4
