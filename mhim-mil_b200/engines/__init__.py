from .common_mil import CommonMIL  # noqa: F401
