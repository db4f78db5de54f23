"""Import the UNMODIFIED reference (/root/reference/src/python) in this container.

Only `tests/golden/make_golden.py` uses this file, and only here (the GPU box has
no /root/reference).  Four third-party packages the reference imports at module
scope are absent from this image; they are replaced by inert stand-ins *before*
import.  Nothing from the reference is copied.
"""
import sys
import types

REF_SRC = "/root/reference/src/python"


class _AttrDict(dict):
    """Minimal stand-in for omegaconf.DictConfig (attribute + item access)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return _AttrDict(dict.copy(self))


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "skimage" not in sys.modules:
        sk = mod("skimage")
        sk.util = mod("skimage.util")
        sk.filters = mod("skimage.filters", threshold_multiotsu=lambda *a, **k: None)

    class _Vis:
        def __init__(self, *a, **k):
            pass

        def check_connection(self):
            return True

        def __getattr__(self, name):
            return lambda *a, **k: None

    if "visdom" not in sys.modules:
        mod("visdom", Visdom=_Vis)
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib", use=lambda *a, **k: None, rcParams={})
        mpl.pyplot = mod("matplotlib.pyplot")
    if "omegaconf" not in sys.modules:
        mod("omegaconf", DictConfig=_AttrDict, OmegaConf=types.SimpleNamespace(load=lambda f: _AttrDict()))


def load_reference():
    """Returns a namespace of the reference's hot-path modules."""
    _install_stubs()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import models, pytorch_tps, pytorch_ssim  # noqa: E401
    import img_proc, classifier, train_network, projector_based_attack  # noqa: E401
    import perc_al
    from perc_al import differential_color_functions as dcf

    return types.SimpleNamespace(
        models=models, pytorch_tps=pytorch_tps, pytorch_ssim=pytorch_ssim, img_proc=img_proc,
        classifier=classifier, train_network=train_network, pba=projector_based_attack,
        perc_al=perc_al, dcf=dcf, DictConfig=_AttrDict)
