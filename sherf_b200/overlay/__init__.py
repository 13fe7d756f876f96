"""Drop-in overlay for the reference's `train.py`.

The reference builds its generator by dotted name (`train.py:310`: 'training.triplane.TriPlaneGenerator', constructed by
`dnnlib.util.construct_class_by_name` in `training_loop.py:193`; resumed by name at `:207-208`), and `triplane.py:19` imports
`ImportanceRenderer, read_pickle, SMPL_to_tensor` from `training.volumetric_rendering.renderer`.  `install()` puts a finder in
front of `sys.meta_path` that resolves exactly those two module names to this package's files; every other `training.*` /
`torch_utils.*` / `dnnlib.*` import keeps resolving to the reference tree, which is left untouched:

    PYTHONPATH=<repo>/sherf_b200/overlay/_site:<repo> SHERF_B200_OVERLAY=1 python train.py ...      # sitecustomize installs it
    python -m sherf_b200.overlay train.py ...                                                        # or: launcher (same effect)

(`train.py`'s own directory is `sys.path[0]`, ahead of PYTHONPATH, so plain path shadowing cannot work; a meta-path finder
can, and `sitecustomize` carries it into the `torch.multiprocessing.spawn` workers of `train.py:98-103`.)

Where the reference tree is absent (the GPU test box), `training` / `training.volumetric_rendering` resolve to empty namespace
packages so the dotted names still import, and the backbone / encoders / super-resolution modules -- reference code that is out
of this repo's scope -- must be supplied through `set_factories()`.  If spconv is not installed, `spconv.*` resolves to a stub
whose classes are plain parameter containers, which is what unpickling a reference snapshot (`legacy.load_network_pkl`) needs
before `copy_params_and_buffers` moves the tensors into this package's modules by name.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SHADOWED = {
    'training.triplane': os.path.join(HERE, 'triplane.py'),
    'training.volumetric_rendering.renderer': os.path.join(HERE, 'renderer.py'),
}
_PACKAGES = ('training', 'training.volumetric_rendering')

# reference-side modules the generator shell owns but this repo does not rebuild (OUT OF SCOPE)
FACTORIES = {'backbone': None, 'encoder_2d': None, 'superresolution': None}


def set_factories(backbone=None, encoder_2d=None, superresolution=None):
    """Callables that build the generator's out-of-scope sub-modules when the reference classes cannot be imported:
    backbone(z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs, **synthesis_kwargs) -> module with
    .mapping(...) / .synthesis(ws, ...) -> [B,96,256,256]; encoder_2d() -> module(x, extract_feature=False);
    superresolution(**kwargs) -> module.  Pass None to clear."""
    FACTORIES.update(backbone=backbone, encoder_2d=encoder_2d, superresolution=superresolution)


class _StubModule(types.ModuleType):
    """`spconv.*` stand-in for unpickling: any attribute is an nn.Module subclass that accepts any constructor arguments."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        import torch.nn as nn
        base = nn.Sequential if 'Sequential' in name else nn.Module

        def __init__(self, *a, **k):
            base.__init__(self)
        cls = type(name, (base,), {'__init__': __init__, '__module__': self.__name__})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in SHADOWED:
            return importlib.util.spec_from_file_location(fullname, SHADOWED[fullname])
        if fullname in _PACKAGES:
            for f in sys.meta_path:                                   # the reference's own package, when it is importable
                if f is self or not hasattr(f, 'find_spec'):
                    continue
                spec = f.find_spec(fullname, path, target)
                if spec is not None:
                    return spec
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)      # namespace package stand-in
            spec.submodule_search_locations = []
            return spec
        if fullname == 'imageio':                                     # imported, unused, by the reference's triplane.py:27
            for f in sys.meta_path:
                if f is self or not hasattr(f, 'find_spec'):
                    continue
                spec = f.find_spec(fullname, path, target)
                if spec is not None:
                    return spec
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        if fullname == 'spconv' or fullname.startswith('spconv.'):
            for f in sys.meta_path:
                if f is self or not hasattr(f, 'find_spec'):
                    continue
                try:
                    spec = f.find_spec(fullname, path, target)
                except (ImportError, AttributeError, ValueError):
                    spec = None
                if spec is not None:
                    return spec
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_finder = None


def install():
    """Idempotent.  Must run before the first `import training.triplane` (sitecustomize / the launcher do that)."""
    global _finder
    if _finder is None:
        _finder = _Finder()
        sys.meta_path.insert(0, _finder)
        for name in SHADOWED:                                         # a copy imported earlier (from the reference) is dropped
            sys.modules.pop(name, None)
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in list(SHADOWED) + [n for n in list(sys.modules) if n == 'imageio' or n == 'spconv' or n.startswith('spconv.')]:
        m = sys.modules.get(name)
        if m is not None and (name in SHADOWED or isinstance(m, _StubModule)):
            del sys.modules[name]
    for name in _PACKAGES[::-1]:
        m = sys.modules.get(name)
        if m is not None and getattr(m, '__file__', None) is None and not list(getattr(m, '__path__', [])):
            del sys.modules[name]


def construct_class_by_name(*args, class_name: str, **kwargs):
    """`dnnlib.util.construct_class_by_name` (dnnlib/util.py:303-305) for environments without the reference tree: import the
    module part of the dotted name, fetch the attribute, call it."""
    mod, _, attr = class_name.rpartition('.')
    return getattr(importlib.import_module(mod), attr)(*args, **kwargs)
