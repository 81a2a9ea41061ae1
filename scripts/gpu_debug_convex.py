"""GPU debug: per-contact-slot comparison of the CUDA narrowphase against a golden fixture."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
for name in sys.argv[1:]:
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    opt = {k: (int(v) if float(v).is_integer() else float(v)) for k, v in zip(z["opt_keys"], z["opt_vals"])}
    mj = util.load_model(str(z["model"]), **opt)
    mx, dc = util.run_cuda(mj, z["in_qpos"], z["in_qvel"], z["in_ctrl"], torch.float64, nsteps=int(z["nsteps"]))
    dist = dc.contact.dist.cpu().numpy(); pos = dc.contact.pos.cpu().numpy(); fr = dc.contact.frame.cpu().numpy()
    g1, g2 = z["contact_geom1"][0], z["contact_geom2"][0]
    gt = np.asarray(mj.geom_type)
    print(name, "ncon", dist.shape)
    for b in range(dist.shape[0]):
        for c in range(dist.shape[1]):
            ed = abs(dist[b, c] - z["contact_dist"][b, c]); ep = np.abs(pos[b, c] - z["contact_pos"][b, c]).max()
            ef = np.abs(fr[b, c] - z["contact_frame"][b, c]).max()
            act = z["contact_dist"][b, c] < 0
            if ed > 1e-9 or (act and (ep > 1e-9 or ef > 1e-9)):
                print(f" env {b} slot {c} geoms {g1[c]}({gt[g1[c]]}) {g2[c]}({gt[g2[c]]}) dist ref {z['contact_dist'][b, c]:.6f} cuda {dist[b, c]:.6f} "
                      f"pos_err {ep:.2e} frame_err {ef:.2e} n_ref {z['contact_frame'][b, c][0]} n_cuda {fr[b, c][0]}")
