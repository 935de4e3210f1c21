"""Path / config helpers — mirrors dial_mpc/utils/io_utils.py:1-24."""
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def get_model_path(robot_name, model_name):
    """Compiled-model blob for ``<robot>/<scene>.xml`` (the reference returns the MJCF path,
    io_utils.py:5-7; here MJCF files are pre-compiled by ``dial_mpc_b200.modelc``)."""
    stem = os.path.splitext(os.path.basename(str(model_name)))[0]
    return os.path.join(_PKG, "models", f"{robot_name}_{stem}.json")


def get_example_path(example_name):
    return os.path.join(_PKG, "examples", example_name)


def load_dataclass_from_dict(dataclass, data_dict, convert_list_to_array=False):
    """Build ``dataclass`` from the intersection of its fields with ``data_dict``
    (unknown keys ignored) — io_utils.py:15-24.  Lists become numpy arrays when
    ``convert_list_to_array`` (the reference makes jnp arrays)."""
    keys = dataclass.__dataclass_fields__.keys() & data_dict.keys()
    kwargs = {key: data_dict[key] for key in keys}
    if convert_list_to_array:
        import numpy as np

        for key, value in kwargs.items():
            if isinstance(value, list):
                kwargs[key] = np.array(value, dtype=np.float64)
    return dataclass(**kwargs)
