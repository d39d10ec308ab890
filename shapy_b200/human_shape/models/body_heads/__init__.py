from .registry import BODY_HEAD_REGISTRY  # noqa: F401
from .body_heads import SMPLXRegressor  # noqa: F401
from .build import build as build_body_head  # noqa: F401
