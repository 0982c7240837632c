from . import models, datasets  # noqa: F401
