"""Config tree of the hot path (monoport/lib/common/config.py:4-100).  Uses yacs when installed, otherwise a
minimal attribute-dict with the same get/set/clone surface MonoPortNet needs."""
try:  # pragma: no cover
    from yacs.config import CfgNode
except Exception:
    import copy

    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            return copy.deepcopy(self)

        def merge_from_list(self, kv):
            for key, val in zip(kv[0::2], kv[1::2]):
                node = self
                parts = key.split(".")
                for p in parts[:-1]:
                    node = node[p]
                node[parts[-1]] = val

        def freeze(self):
            pass

CN = CfgNode


def _net(backbone, head, loss):
    n = CN()
    n.mean = (0.5, 0.5, 0.5)
    n.std = (0.5, 0.5, 0.5)
    n.ckpt_path = ''
    n.projection = 'orthogonal'
    n.backbone = CN(); n.backbone.IMF = backbone
    n.normalizer = CN(); n.normalizer.IMF = 'PIFuNomalizer'; n.normalizer.soft_onehot = False; n.normalizer.soft_dim = 64
    n.head = CN(); n.head.IMF = head
    n.loss = CN(); n.loss.IMF = loss
    return n


def get_cfg_defaults():
    """netG / netC sub-trees with the reference's defaults (common/config.py:29-72)."""
    c = CN()
    c.name = 'default'
    c.netG = _net('PIFuHGFilters', 'PIFuNetGMLP', 'MSE')
    c.netC = _net('PIFuResBlkFilters', 'PIFuNetCMLP', 'L1')
    return c
