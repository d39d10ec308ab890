from .build import build_model  # noqa: F401
from .body_heads import BODY_HEAD_REGISTRY, SMPLXRegressor  # noqa: F401
from .body_models import KeypointTensor, build_body_model  # noqa: F401
