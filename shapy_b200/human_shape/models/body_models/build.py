from .body_models import build_body_model  # noqa: F401
