import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import util
torch.set_printoptions(precision=10, linewidth=200)
dtype = torch.float64
def run(name, opt, drop, fields=("qacc","efc_force","qfrc_constraint","qvel")):
    mj = util.load_model(name, **opt)
    qpos, qvel, ctrl = util.seeded_state(mj, 64, drop=drop)
    mo, do = util.run_oracle(mj, qpos, qvel, ctrl, dtype)
    mx, dc = util.run_cuda(mj, qpos, qvel, ctrl, dtype)
    print(name, opt, {k: f"{util.rel_err(getattr(dc,k), getattr(do,k)):.2e}" for k in fields}, 'niter', do.solver_niter_oracle.tolist()[:8] if hasattr(do,'solver_niter_oracle') else None)
    return mj, mo, do, mx, dc
run("halfcheetah", dict(iterations=1, ls_iterations=4), 0.6)
run("halfcheetah", dict(iterations=2, ls_iterations=50), 0.6)
run("humanoid", dict(iterations=100, ls_iterations=50), 0.4)
run("humanoid", dict(iterations=1, ls_iterations=50), 0.4)
run("humanoid", dict(iterations=1, ls_iterations=0), 0.4)
mj, mo, do, mx, dc = run("humanoid", {}, 0.4)
e = (dc.qacc.cpu()-do.qacc).abs().max(-1).values
print('per-env qacc err', e[:16])
# ant frames
mj, mo, do, mx, dc = run("ant", {}, 0.0)
fo, fc = do.contact_frame, dc.contact.frame.cpu()
diff = (fo-fc).abs().amax((-1,-2))
idx = (diff > 1e-6).nonzero()
print('n frame mismatches', idx.shape[0], 'of', diff.numel())
for b, c in idx[:6].tolist():
    print('env', b, 'contact', c, 'dist', do.contact_dist[b,c].item(), dc.contact.dist[b,c].item(), 'g', int(mo.contact['geom1'][c]), int(mo.contact['geom2'][c]))
    print(' oracle', fo[b,c].flatten()); print(' cuda  ', fc[b,c].flatten())
