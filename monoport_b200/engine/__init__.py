from .seg3d import Seg3dLossless, Seg3dTopk, make_query_func, plot_mask3D  # noqa: F401
