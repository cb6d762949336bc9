"""network.GomokuNetEZ must be state_dict-compatible with the reference's (network.py:109-152) and
compute the same forward: checked against a KAT recorded from the reference (float32, CPU)."""
import os

import numpy as np
import torch

from _golden_util import GOLDEN_DIR


def test_network_matches_reference_kat():
    from datou_gomoku_muzero_b200.config import Config
    from datou_gomoku_muzero_b200.network import GomokuNetEZ
    z = np.load(os.path.join(GOLDEN_DIR, "network_kat.npz"))
    n, blocks, ch, hid = (int(x) for x in z["cfg"])
    cfg = Config(BOARD_SIZE=n, ACTION_SPACE_SIZE=n * n, NUM_RES_BLOCKS=blocks, NUM_FILTERS=ch, HEAD_HIDDEN_DIM=hid)
    net = GomokuNetEZ(cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")}
    net.load_state_dict(sd, strict=True)                      # same parameter / buffer names
    p, v, h = net.initial_inference(torch.from_numpy(z["obs"]))
    np.testing.assert_allclose(p.numpy(), z["p"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v.numpy(), z["v"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(h.numpy(), z["h"], rtol=1e-5, atol=1e-6)
    p2, v2, h2, r2 = net.recurrent_inference(torch.from_numpy(z["h"]), torch.from_numpy(z["act"]))
    np.testing.assert_allclose(p2.numpy(), z["p2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v2.numpy(), z["v2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(h2.numpy(), z["h2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r2.numpy(), z["r2"], rtol=1e-5, atol=1e-6)
    assert v.shape == (5, 1) and r2.shape == (5, 1) and h.shape == (5, ch, n, n)
