import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
import mujoco_torch_b200 as mt
mj = util.load_model("humanoid")
qpos, qvel, ctrl = util.seeded_state(mj, 4, drop=0.4)
mx = mt.device_put(mj, dtype=torch.float64).to("cuda:0")
T = lambda a: torch.as_tensor(a, device="cuda:0")
d = mt.make_data(mx).expand(4).clone().replace(qpos=T(qpos), qvel=T(qvel), ctrl=T(ctrl))
md = types.SimpleNamespace(qpos=d.qpos[1].cpu().numpy(), qvel=d.qvel[1].cpu().numpy(), act=np.zeros(0), time=0.0, ctrl=d.ctrl[1].cpu().numpy(),
                           qacc_warmstart=np.zeros(mj.nv), qfrc_applied=np.zeros(mj.nv), xfrc_applied=np.zeros((mj.nbody, 6)),
                           mocap_pos=np.zeros((0, 3)), mocap_quat=np.zeros((0, 4)), eq_active=np.zeros(0, np.uint8), sensordata=np.zeros(0))
dx = mt.device_put(md).to("cuda:0")
ref = mt.step(mx, d)
out = mt.step(mx, dx)
one = mt.step(mx, d[1])
for k in ("qpos", "qvel", "qacc"):
    print(k, float((getattr(out, k) - getattr(ref, k)[1]).abs().max()), float((getattr(one, k) - getattr(ref, k)[1]).abs().max()))
for name in ("qpos","qvel","ctrl","qacc_warmstart","qfrc_applied","xfrc_applied","time","act","eq_active","mocap_pos","sensordata","qfrc_constraint"):
    a, b = getattr(dx, name), getattr(d[1], name)
    print(name, a.shape, b.shape, a.dtype, b.dtype, bool(torch.equal(a.double().reshape(-1), b.double().reshape(-1))) if a.numel()==b.numel() else 'SIZE')
