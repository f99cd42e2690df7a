from monoport_b200.engine import plot_mask3D  # noqa: F401
