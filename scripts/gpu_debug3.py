import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
from tests import util
from tests.test_gpu_step import _env_err
torch.set_printoptions(precision=2, linewidth=250, sci_mode=True)
dtype = torch.float64
mj = util.load_model("humanoid")
qpos, qvel, ctrl = util.seeded_state(mj, 64, drop=0.4)
mo, do = util.run_oracle(mj, qpos, qvel, ctrl, dtype)
mx, dc = util.run_cuda(mj, qpos, qvel, ctrl, dtype)
rng = np.random.RandomState(7)
spread = torch.zeros(64, dtype=torch.float64)
for _ in range(6):
    pq = qpos * (1 + 16 * 2e-16 * rng.randn(*qpos.shape)); pv = qvel * (1 + 16 * 2e-16 * rng.randn(*qvel.shape))
    dv = util.run_oracle(mj, pq, pv, ctrl, dtype)[1]
    spread = torch.maximum(spread, _env_err(dv.qacc, do.qacc))
err = _env_err(dc.qacc, do.qacc)
print("err   ", err)
print("spread", spread)
