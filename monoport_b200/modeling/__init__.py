from .MonoPortNet import MonoPortNet, PIFuNetG, PIFuNetC  # noqa: F401
from . import geometry  # noqa: F401
