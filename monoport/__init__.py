"""Drop-in import surface: `monoport.lib.*` of the reference, backed by monoport_b200."""
