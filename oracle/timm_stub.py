"""TEST INFRASTRUCTURE ONLY -- six-name stand-in for the `timm` package.

The reference encoder (llava/model/multimodal_encoder/mobileclip/mci.py:15-17 and
mobileclip/__init__.py:10) imports `register_model`, `create_model`,
`IMAGENET_DEFAULT_MEAN/STD`, `DropPath` and `SqueezeExcite` from timm.  timm is not
installed in this image and there is no network.  None of these names does arithmetic
on the FastViTHD path (DropPath / SqueezeExcite are never constructed for `fastvithd`,
mci.py:1097,1173,414-417), so a stub lets the *unmodified* reference files execute.

Install with `install()` AFTER `import transformers` (transformers probes
importlib.util.find_spec("timm") at import time and a spec-less module raises).
"""
import sys
import types
import importlib.machinery

_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def create_model(model_name, **kwargs):
    if model_name not in _REGISTRY:
        raise RuntimeError(f"timm stub: unknown model {model_name!r}")
    return _REGISTRY[model_name](**kwargs)


class _NeverBuilt:
    """DropPath / SqueezeExcite must never be instantiated on the fastvithd path."""

    def __init__(self, *a, **k):
        raise RuntimeError("timm stub: %s constructed -- not on the FastViTHD path" % type(self).__name__)


class DropPath(_NeverBuilt):
    pass


class SqueezeExcite(_NeverBuilt):
    pass


def install():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "__fvhd_stub__", False):
        return
    import transformers  # noqa: F401  (must be imported before the stub exists)

    def mod(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        m.__fvhd_stub__ = True
        sys.modules[name] = m
        return m

    timm = mod("timm")
    models = mod("timm.models")
    data = mod("timm.data")
    layers = mod("timm.layers")
    timm.models, timm.data, timm.layers = models, data, layers
    timm.__path__ = []
    models.register_model = register_model
    models.create_model = create_model
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    layers.DropPath = DropPath
    layers.SqueezeExcite = SqueezeExcite
