from .hrnet import HighResolutionNet, build as build_hrnet  # noqa: F401
from .build import build_backbone  # noqa: F401
