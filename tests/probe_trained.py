import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from oracle import omok_oracle as O, pvnet_ref
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trained_9x9_180927.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
rs = np.random.RandomState(0)
ids = [(0,) + tuple(int(a) for a in rs.permutation(81)[:rs.randint(0, 50)]) for _ in range(128)]
states = np.stack([O.get_state_pt(i, 9, 5) for i in ids]).astype(np.float32)
pr, vr = pvnet_ref.pvnet_forward(sd, torch.from_numpy(states))
for mode in (0, 1):
    eng = _cabi.Engine(board_size=9, num_mcts=8, max_games=128, nn_precision=mode)
    eng.load_state_dict(sd)
    p, v = eng.nn_forward(states)
    dp, dv = np.abs(p - pr.numpy()), np.abs(v - vr.numpy())
    print("mode", mode, "dp max %.3e mean %.3e  dv max %.3e mean %.3e  worst idx %d len %d" % (dp.max(), dp.mean(), dv.max(), dv.mean(), dp.max(1).argmax(), len(ids[dp.max(1).argmax()])))
    eng.close()
