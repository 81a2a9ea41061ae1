from . import _TensorClassMeta  # noqa: F401
