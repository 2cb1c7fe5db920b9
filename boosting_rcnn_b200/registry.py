"""The registry boundary.

The reference builds every component from config dicts through
``MODELS = Registry('models', parent=MMCV_MODELS)`` with ``HEADS``,
``ROI_EXTRACTORS`` ... all aliasing it (mmdet/models/builder.py:7-15,38-59).
mmcv is not installable in the build image, so this module carries a minimal
work-alike (``Registry``, ``build_from_cfg``, ``ConfigDict``) exposing the
same names; when real mmdet IS importable, :func:`register_into_mmdet` drops
the B200 classes into its registries (``force=True``), which is what a
config's ``custom_imports = dict(imports=['boosting_rcnn_b200.mmdet_plugin'])``
triggers.
"""
import inspect


class ConfigDict(dict):
    """dict with attribute access (stand-in for mmcv.ConfigDict / addict)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return ConfigDict(self)

    def __deepcopy__(self, memo):
        import copy
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def build_from_cfg(cfg, registry, default_args=None):
    """mmcv.utils.build_from_cfg: pop ``type``, look it up, call ``cls(**cfg)``."""
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = dict(cfg)
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    try:
        return obj_cls(**args)
    except Exception as e:
        raise type(e)(f'{obj_cls.__name__}: {e}')


class Registry:

    def __init__(self, name, build_func=None, parent=None):
        self._name = name
        self._module_dict = {}
        self.build_func = build_func or build_from_cfg
        self.parent = parent

    @property
    def name(self):
        return self._name

    def __contains__(self, key):
        return self.get(key) is not None

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register(self, cls, name=None, force=False):
        names = [name or cls.__name__] if not isinstance(name, (list, tuple)) else list(name)
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self.name}')
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _decorator(cls):
            self._register(cls, name, force)
            return cls

        return _decorator


MODELS = Registry('models')
BACKBONES = NECKS = ROI_EXTRACTORS = SHARED_HEADS = HEADS = LOSSES = DETECTORS = MODELS
BBOX_CODERS = Registry('bbox_coder')
ANCHOR_GENERATORS = Registry('Anchor generator')
BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')


def build_head(cfg):
    return HEADS.build(cfg)


def build_roi_extractor(cfg):
    return ROI_EXTRACTORS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_bbox_coder(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_CODERS, default_args)


def build_anchor_generator(cfg, default_args=None):
    return build_from_cfg(cfg, ANCHOR_GENERATORS, default_args)


def build_assigner(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_ASSIGNERS, default_args)


def build_sampler(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_SAMPLERS, default_args)


HOT_PATH_CLASSES = ('ATSSRPNHead', 'SingleRoIExtractor', 'ProbRoIHead',
                    'ProbConvFCBBoxHead')


def register_into_mmdet(force=True, names=HOT_PATH_CLASSES):
    """Register the B200 classes under the reference's names in a real mmdet
    install (the drop-in step).  Raises ImportError when
    mmdet/mmcv are absent."""
    from mmdet.models.builder import HEADS as MM_HEADS  # noqa: N811
    from mmdet.models.builder import ROI_EXTRACTORS as MM_EXTRACTORS  # noqa: N811
    from . import bbox_head, roi_extractor, roi_head, rpn_head
    table = {
        'ATSSRPNHead': (MM_HEADS, rpn_head.ATSSRPNHead),
        'ProbRoIHead': (MM_HEADS, roi_head.ProbRoIHead),
        'ProbConvFCBBoxHead': (MM_HEADS, bbox_head.ProbConvFCBBoxHead),
        'SingleRoIExtractor': (MM_EXTRACTORS, roi_extractor.SingleRoIExtractor),
    }
    for n in names:
        reg, cls = table[n]
        reg.register_module(name=n, force=force, module=cls)
    return tuple(names)
