"""GPU debug: per-env step errors against the oracle replaying the CUDA contacts, with the oracle's own perturbation spread."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from oracle import forward as oforward
name, drop = sys.argv[1], float(sys.argv[2])
opt = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
mj = util.load_model(name, **opt)
B = 64
qpos, qvel, ctrl = util.pile_state(mj, B) if drop > -1.5 else util.stack_state(mj, B)
mx, dc = util.run_cuda(mj, qpos, qvel, ctrl, torch.float64)
cd, cp, cf = dc.contact.dist.cpu(), dc.contact.pos.cpu(), dc.contact.frame.cpu()
def replay(m, d):
    d.contact_dist, d.contact_pos, d.contact_frame = cd.clone(), cp.clone(), cf.clone()
    return d
oforward.CONTACT_HOOK = replay
mo, do = util.run_oracle(mj, qpos, qvel, ctrl)
rng = np.random.RandomState(7)
spread = torch.zeros(B, dtype=torch.float64)
for _ in range(6):
    pq = qpos * (1 + 16 * 2e-16 * rng.randn(*qpos.shape)); pv = qvel * (1 + 16 * 2e-16 * rng.randn(*qvel.shape))
    dv = util.run_oracle(mj, pq, pv, ctrl)[1]
    spread = torch.maximum(spread, (dv.qacc - do.qacc).abs().amax(1) / do.qacc.abs().max())
err = (dc.qacc.cpu() - do.qacc).abs().amax(1) / do.qacc.abs().max()
for k in ("efc_J", "efc_D", "efc_aref", "qfrc_smooth", "qM"):
    a, b = getattr(dc, k).cpu(), getattr(do, k)
    print(k, float((a - b).abs().max() / b.abs().max()))
order = err.argsort(descending=True)[:6]
for e in order:
    nact = int((cd[e] < 0).sum())
    print(f"env {int(e)} qacc err {float(err[e]):.3e} oracle spread {float(spread[e]):.3e} active {nact} "
          f"force err {float((dc.efc_force.cpu()[e] - do.efc_force[e]).abs().max() / do.efc_force.abs().max()):.3e} "
          f"niter? qacc_o {do.qacc[e][:3].numpy()} qacc_c {dc.qacc.cpu()[e][:3].numpy()}")
