"""B200-native batched physics step with the mujoco_torch public API."""
