from . import conv, models  # noqa: F401
