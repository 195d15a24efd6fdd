from .frame_cache import FrameCache  # noqa: F401
