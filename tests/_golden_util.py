"""Helpers shared by the golden-vector tests (CPU oracle and CUDA engine)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_search_cases(mode, N):
    path = os.path.join(GOLDEN_DIR, f"search_{mode}_{N}.npz")
    z = np.load(path)
    cases = []
    for i in range(int(z["n_cases"])):
        c = {k[len(f"c{i}_"):]: z[k] for k in z.files if k.startswith(f"c{i}_")}
        p, f = c["params"], c["fparams"]
        c.update(N=int(p[0]), n_in_row=int(p[1]), S=int(p[2]), K=int(p[3]), kind=int(p[4]), logit_div=int(p[5]),
                 seed=int(p[6]), vdtype=int(p[7]) if len(p) > 7 else 0, c_visit=float(f[0]), c_scale=float(f[1]), delta=float(f[2]), discount=float(f[3]),
                 const_value=float(f[4]), const_reward=float(f[5]), idx=i, mode=mode)
        for k in ("player", "last_move", "move_count", "action", "root_n", "n_initial", "n_recurrent"):
            c[k] = int(c[k])
        c["value"] = float(c["value"]); c["root_w"] = float(c["root_w"])
        cases.append(c)
    return cases
