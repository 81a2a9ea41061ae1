"""Compile the reference's bundled MJCF models with our mini-compiler and store the compiled arrays as .npz under
assets/models/ (the GPU box has no /root/reference; tests, smoke() and bench.py load these).

    python scripts/compile_models.py [/root/reference/mujoco_torch/test_data]
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mujoco_torch_b200 import mjcf  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/mujoco_torch/test_data"
OUT = os.path.join(ROOT, "assets", "models")
os.makedirs(OUT, exist_ok=True)

MODELS = ["cartpole", "humanoid", "halfcheetah", "ant", "hopper", "walker2d", "swimmer", "convex", "pendula", "equality",
          "frictionloss_dof", "satellite_small", "satellite_large"]
for name in MODELS:
    path = os.path.join(SRC, name + ".xml")
    try:
        m = mjcf.from_xml_path(path)
    except Exception as e:  # noqa: BLE001
        print(f"{name}: SKIPPED ({type(e).__name__}: {e})")
        continue
    m.save_npz(os.path.join(OUT, name + ".npz"))
    print(f"{name}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom}")

# zoo-patched Ant (config 3): free root + floor plane, dt = 0.01
with open(os.path.join(SRC, "ant.xml")) as fh:
    xml = mjcf.patch_zoo_ant(fh.read())
m = mjcf.from_xml_string(xml, SRC)
m.save_npz(os.path.join(OUT, "ant_zoo.npz"))
print(f"ant_zoo: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom}")

# scenes authored for this repo (assets/xml)
for name in ("shapes", "stack", "equality_bodies", "sink", "muscle", "capsules_topk", "terrain", "sensors_extra"):
    m = mjcf.from_xml_path(os.path.join(ROOT, "assets", "xml", name + ".xml"))
    m.save_npz(os.path.join(OUT, name + ".npz"))
    print(f"{name}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom}")
