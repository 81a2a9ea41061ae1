"""Print per-field CUDA-vs-oracle errors for a list of models (debug helper for gpurun)."""
import sys, os, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import util
from tests.test_gpu_step import FIELDS_SMOOTH, FIELDS_EFC, FIELDS_STATE, CASES

names = sys.argv[1:] 
for name, opt, drop in CASES:
    if names and name not in names:
        continue
    for dtype in (torch.float64, torch.float32):
        try:
            mj = util.load_model(name, **opt)
            qpos, qvel, ctrl = util.seeded_state(mj, 64, drop=drop)
            mo, do = util.run_oracle(mj, qpos, qvel, ctrl, dtype)
            mx, dc = util.run_cuda(mj, qpos, qvel, ctrl, dtype)
            errs = {}
            if mo.ncon:
                act_o = (do.contact_dist - mo.contact["includemargin"]) < 0
                act_c = (dc.contact.dist.cpu() - dc.contact.includemargin.cpu()) < 0
                errs["active_mismatch"] = int((act_o != act_c).sum()); errs["n_active"] = int(act_o.sum())
                for k, ko in (("contact.dist", "contact_dist"), ("contact.pos", "contact_pos"), ("contact.frame", "contact_frame")):
                    errs[k] = util.rel_err(util.get(dc, k), getattr(do, ko))
            for k in FIELDS_SMOOTH + (FIELDS_EFC if mo.nefc else []) + FIELDS_STATE:
                if hasattr(do, k):
                    errs[k] = util.rel_err(getattr(dc, k), getattr(do, k))
            worst = sorted(((v, k) for k, v in errs.items() if isinstance(v, float)), reverse=True)[:8]
            print(f"== {name} {opt} {dtype}: active_mismatch={errs.get('active_mismatch')} n_active={errs.get('n_active')}")
            print("   worst:", ", ".join(f"{k}={v:.2e}" for v, k in worst))
        except Exception:
            print(f"== {name} {opt} {dtype}: EXCEPTION")
            print('   ', traceback.format_exc().strip().splitlines()[-1][:300])
