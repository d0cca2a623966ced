"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
from . import Batch  # noqa: F401
