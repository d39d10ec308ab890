from .body_models import SMPLX, build_body_model  # noqa: F401
from .utils import KeypointTensor  # noqa: F401
from .lbs import lbs  # noqa: F401
