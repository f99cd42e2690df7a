from monoport_b200.modeling.geometry import index, orthogonal, perspective  # noqa: F401
