from monoport_b200.config import CfgNode, CN, get_cfg_defaults  # noqa: F401
