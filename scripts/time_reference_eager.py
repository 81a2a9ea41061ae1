"""Time the UNMODIFIED reference's eager CPU step (build container only: it imports /root/reference through oracle/refshim).

    python scripts/time_reference_eager.py [model ...]      -> profiles/reference_eager_cpu.json

Protocol of the reference's benchmarks/bench_loop.py: one environment, a Python loop of `mujoco_torch.step(mx, d)` calls, float64,
eager (no torch.compile, no vmap).  The reference's Python sources cannot travel to the GPU box (they are not part of this repo
and /root/reference does not exist there), so bench.py quotes this file as `cpu_baseline_ref` -- context, measured here, with the
host's core count -- next to the CPU arm it does time on the box (the oracle port).
"""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MJT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, ROOT)
pkg = types.ModuleType("mujoco_torch")
pkg.__path__ = [os.path.join(REF, "mujoco_torch")]
sys.modules["mujoco_torch"] = pkg

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

import mujoco  # noqa: E402  (the shim)
from mujoco_torch._src import device as ref_device  # noqa: E402
from mujoco_torch._src import forward as ref_forward  # noqa: E402
from mujoco_torch._src import io as ref_io  # noqa: E402

from tests import util  # noqa: E402


def main():
    models = sys.argv[1:] or ["humanoid", "halfcheetah"]
    out = {}
    for name in models:
        lite = util.load_model(name)
        qpos, qvel, ctrl = util.seeded_state(lite, 1, drop=0.3)
        mx = ref_device.device_put(mujoco.MjModel(lite))
        d = ref_io.make_data(mx).replace(qpos=torch.as_tensor(qpos[0]), qvel=torch.as_tensor(qvel[0]), ctrl=torch.as_tensor(ctrl[0]))
        for _ in range(3):
            d = ref_forward.step(mx, d)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            d = ref_forward.step(mx, d)
            n += 1
        dt = time.perf_counter() - t0
        out[name] = {"value": n / dt, "unit": "env-steps/s", "cores": os.cpu_count(), "threads": torch.get_num_threads(), "kind": "_ref",
                     "sample": f"unmodified reference mujoco_torch._src.forward.step, eager float64, B=1 Python loop (benchmarks/bench_loop.py protocol), {n} steps in {dt:.1f} s",
                     "where": "build container (no GPU), through oracle/refshim; NOT measured in the bench run that quotes it",
                     "torch": torch.__version__}
        print(name, out[name]["value"], "steps/s", flush=True)
    path = os.path.join(ROOT, "profiles", "reference_eager_cpu.json")
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    old.update(out)
    json.dump(old, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
