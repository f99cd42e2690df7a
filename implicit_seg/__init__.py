"""Drop-in for the un-vendored `implicit-seg` dependency (requirements.txt:15), backed by monoport_b200.engine."""
