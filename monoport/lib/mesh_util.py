from monoport_b200.mesh_util import save_obj_mesh, save_obj_mesh_with_color  # noqa: F401
