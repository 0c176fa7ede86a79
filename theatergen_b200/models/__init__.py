from . import pipelines  # noqa: F401
