# Mirrors the reference package layout (src/model/__init__.py): ``from model import SSD``.
from .ssd import SSD  # noqa: F401
