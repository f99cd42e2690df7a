from monoport_b200.modeling.MonoPortNet import MonoPortNet, PIFuNetG, PIFuNetC  # noqa: F401
