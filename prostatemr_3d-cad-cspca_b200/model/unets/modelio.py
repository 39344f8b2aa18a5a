"""Mirror of tf2.5/scripts/model/unets/modelio.py: constructor-argument capture and model loading.

  store_config_args              R:modelio.py:20-55  (re-stated with inspect.signature; the
                                 reference's inspect.getargspec no longer exists in Python 3.11+, Q11)
  LoadableModel.get_config       R:modelio.py:81-88
  LoadableModel.from_config      R:modelio.py:90-95
  LoadableModel.load             R:modelio.py:98-117 (config JSON + weights; here a .npz archive
                                 written by save(): h5py is not available in this environment)
  LoadableModel.ReferenceContainer R:modelio.py:119-135
"""
import functools
import inspect
import json

import numpy as np


def store_config_args(func):
    """Class-method decorator that saves every argument provided to the function as a dictionary
    in 'self.config' (defaults first, then positionals, then keywords)."""
    sig = inspect.signature(func)
    names = list(sig.parameters)[1:]

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        retval = func(self, *args, **kwargs)
        params = {}
        for n in names:
            d = sig.parameters[n].default
            if d is not inspect.Parameter.empty:
                params[n] = d
        for attr, val in zip(names, args):
            params[attr] = val
        params.update(kwargs)
        self.config = ModelConfig(params)
        return retval
    return wrapper


class ModelConfig:
    """A separate class to contain the model config (R:modelio.py:58-65)."""

    def __init__(self, params):
        self.params = params


def _npz_path(path):
    """np.savez appends '.npz' to a path that lacks it: save() and load() normalise the same way, so that
    save('w.h5') / load('w.h5') round-trip (the archive is then 'w.h5.npz')."""
    path = str(path)
    return path if path.endswith(".npz") else path + ".npz"


def _jsonable(v):
    if hasattr(v, "get_config"):
        return {"class": type(v).__name__, "config": v.get_config()}
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if v is None or isinstance(v, (bool, int, float, str)):
        return v
    return str(v)                      # e.g. a torch.device passed as `device`


def _from_jsonable(v):
    from .. import initializers, regularizers
    if isinstance(v, dict) and set(v) == {"class", "config"}:
        cls = getattr(initializers, v["class"], None) or getattr(regularizers, v["class"])
        return cls(**v["config"])
    if isinstance(v, list):
        return tuple(_from_jsonable(x) for x in v)
    return v


class LoadableModel:
    """Base class for model loading without having to specify the architecture at load time."""

    def get_config(self):
        if not hasattr(self, "config"):
            raise RuntimeError("models that inherit from LoadableModel must decorate the constructor "
                               "with @store_config_args")
        return self.config.params

    @classmethod
    def from_config(cls, config, custom_objects=None):
        return cls(**config)

    def save(self, path):
        """Weights + optimizer state (Adam m/v/v-hat and the step, which the reference loses on
        resume) + the constructor config, as one .npz archive."""
        if getattr(self, "eng", None) is None:
            raise RuntimeError("save(): the model holds no weights yet (built with build=False / no GPU)")
        arrays = {"w/" + k: v for k, v in self.get_weights().items()}
        arrays.update({"opt/" + k: v for k, v in self.get_optimizer_state().items()})
        # `device` and `build` describe the process that saved, not the model: load() decides them afresh
        cfg = {k: _jsonable(v) for k, v in self.get_config().items() if k not in ("device", "build")}
        arrays["model_config"] = np.frombuffer(json.dumps({"config": cfg}).encode("utf-8"), dtype=np.uint8)
        np.savez(_npz_path(path), **arrays)

    @classmethod
    def load(cls, path, by_name=False):
        with np.load(_npz_path(path), allow_pickle=False) as f:
            config = json.loads(bytes(f["model_config"]).decode("utf-8"))["config"]
            config = {k: _from_jsonable(v) for k, v in config.items()}
            weights = {k[2:]: f[k] for k in f.files if k.startswith("w/")}
            opt = {k[4:]: f[k] for k in f.files if k.startswith("opt/")}
        model = cls(**config)
        model.set_weights(weights, strict=not by_name)
        if opt:
            model.set_optimizer_state(opt)
        return model

    class ReferenceContainer:
        """Attribute bag of layer/tensor references (R:modelio.py:119-135)."""

        def __init__(self):
            pass
