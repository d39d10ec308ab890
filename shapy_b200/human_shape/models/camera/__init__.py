from .camera_projection import build_cam_proj, CameraParams, WeakPerspectiveCamera  # noqa: F401
