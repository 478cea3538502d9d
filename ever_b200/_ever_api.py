"""Binding to the reference's plugin surface.

When the reference package is importable (``import ever``) the real ``ever.ERModule`` base class and
``ever.registry.MODEL`` registry are used, so ``ever.core.builder.make_model`` (ever/core/builder.py:47-62),
``Launcher`` (ever/core/launcher.py) and ``infer_tool`` find the B200 models under their registered names.
On a box without the reference (the GPU runner) a minimal stand-in with the same contract is used:
``ERModule(config)`` -> ``set_default_config()`` then recursive ``config.update`` (ever/interface/module.py:12-29,
ever/interface/configurable.py:19-24, ever/core/config.py:57-89), ``Registry.register`` (ever/core/registry.py:46-85).
"""
from collections import OrderedDict

import torch.nn as nn

try:  # pragma: no cover - depends on the environment
    import ever as _er
    ERModule = _er.ERModule
    MODEL = _er.registry.MODEL
    AttrDict = _er.core.config.AttrDict
    HAVE_EVER = True
except Exception:  # reference not installed: same-contract stand-in
    HAVE_EVER = False

    class AttrDict(OrderedDict):
        def __init__(self, **kwargs):
            super().__init__()
            self.update(kwargs)

        def __setitem__(self, key, value):
            super().__setitem__(key, value)
            super().__setattr__(key, value)

        def __setattr__(self, key, value):
            super().__setitem__(key, value)
            super().__setattr__(key, value)

        def update(self, config):
            for k, v in config.items():
                if k not in self:
                    self[k] = AttrDict()
                if isinstance(v, dict):
                    if not isinstance(self[k], dict):
                        self[k] = AttrDict()
                    self[k].update(v)
                else:
                    self[k] = v

    class _Registry(dict):
        def register(self, module_name=None, module=None, override=False, verbose=True):
            def _do(name, obj):
                self[name if name is not None else obj.__name__] = obj
                return obj
            if module is not None:
                _do(module_name, module)
                return None
            return lambda fn: _do(module_name, fn)

    MODEL = _Registry()

    class ERModule(nn.Module):
        def __init__(self, config=None):
            super().__init__()
            self._cfg = AttrDict()
            self.set_default_config()
            self._cfg.update(config or {})
            if 'GLOBAL' not in self._cfg:
                self._cfg['GLOBAL'] = AttrDict()

        @property
        def config(self):
            return self._cfg

        def set_default_config(self):
            raise NotImplementedError('The default config should be overridden.')

        def init_from_weight_file(self):
            return None

        def log_info(self):
            return dict()

        def custom_param_groups(self):
            return [{'params': self.parameters()}]
