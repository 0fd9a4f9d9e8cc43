from . import controller_manager, message  # noqa: F401
