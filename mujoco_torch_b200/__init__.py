"""B200-native batched physics step with the ``mujoco_torch`` public API.

    import mujoco_torch_b200 as mujoco_torch
    mx = mujoco_torch.device_put(mujoco_torch.mjcf.load("humanoid.xml"), dtype=torch.float32).to("cuda")
    d = mujoco_torch.make_data(mx).expand(4096).clone()
    d = mujoco_torch.step(mx, d)
"""

from . import mjcf  # noqa: F401
from . import zoo  # noqa: F401
from .api import (  # noqa: F401
    collision, com_pos, com_vel, constraint_sizes, crb, device_put, factor_m, forward, fwd_acceleration, fwd_actuation,
    fwd_position, fwd_velocity, kinematics, make_constraint, make_data, passive, rne, sensor_acc, sensor_pos, sensor_vel,
    solve, step, tendon, transmission,
)
from .data import Contact, Data  # noqa: F401
from .devget import device_get_into  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .hostio import HostPipeline  # noqa: F401
from .mjmodel import (  # noqa: F401
    BiasType, ConeType, DisableBit, DynType, EnableBit, EqType, GainType, GeomType, IntegratorType, JacobianType, JointType,
    MjModelLite, SensorType, SolverType, TrnType,
)
from .model import Model, Option, Statistic  # noqa: F401

__version__ = "0.1.0"
