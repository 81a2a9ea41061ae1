"""ORACLE: CPU restatement of the reference's batched physics step (test infrastructure only).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline/reference legs may
import this package; the product (``mujoco_torch_b200``) never does.
"""
