from monoport_b200.engine import Seg3dLossless, Seg3dTopk  # noqa: F401
