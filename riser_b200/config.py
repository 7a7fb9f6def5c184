"""Model configuration holder: the attribute-style dict riser/riser.py:21-23 builds
from the YAML files with ``attridict`` (config.cnn.n_layers, ...)."""
import yaml

CNN_SHIPPED = {"n_layers": 12, "depth": 1,
               "channels": [20, 30, 45, 67, 100, 150, 225, 337, 505, 757, 1135, 1702],
               "kernels": [3] * 12, "n_classes": 2, "classifier": "gap_fc"}   # riser/model/*.yaml:6-12


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v


def get_config(filepath):
    """riser/riser.py:21-23."""
    with open(filepath) as config_file:
        return AttrDict(yaml.safe_load(config_file))


def shipped_config():
    """The cnn block every shipped config carries (riser/model/*_config_*.yaml:6-12)."""
    return AttrDict({"model": "cnn", "cnn": dict(CNN_SHIPPED)})
