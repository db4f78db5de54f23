"""Import the UNMODIFIED reference (its src/python directory of scripts) and run its hot-path functions.

Users: `tests/golden/make_golden.py` (fixtures, build container only) and `bench.py`'s reference legs (the reference's own `spaa` /
`perc_al_compennet_pp` / `train_pcnet` timed on the box's host cores and, as a side leg, on the B200 through stock PyTorch-CUDA).
Source location: `/root/reference/src/python` where it exists (the build container), else `baseline/_ref/src/python` -- the git-ignored
copy that `__graft_entry__.build()` makes so that the reference travels to the GPU box (the reference is a directory of scripts with no
setup.py / pyproject.toml: "installing" it is copying it).  Four third-party packages the reference imports at module scope are absent
from this image; they are replaced by inert stand-ins *before* import.  Nothing from the reference is
copied into the tracked tree, and nothing under spaa_b200/ imports this file.
"""
import contextlib
import io
import os
import sys
import time
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_CANDIDATES = ("/root/reference/src/python", os.path.join(_ROOT, "baseline", "_ref", "src", "python"))
REF_SRC = next((p for p in _CANDIDATES if os.path.isdir(p)), _CANDIDATES[0])


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "projector_based_attack.py"))


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


# ----------------------------------------------------------------------------------------------------------------
# helpers to run the reference's own functions on synthetic inputs (no pretrained weights / dataset offline, SURVEY.md 8c "Gaps")
# ----------------------------------------------------------------------------------------------------------------

def patched(fn_module, fn_name, old, new):
    """Re-execute a reference function from its source with one literal replaced (the text on disk is untouched): `iters = 50` of
    `spaa` (projector_based_attack.py:258) and `iters = 0` of `train_pcnet` (train_network.py:292) are hard-coded locals."""
    import inspect
    import textwrap
    src = textwrap.dedent(inspect.getsource(getattr(fn_module, fn_name)))
    assert old in src, (fn_name, old)
    ns = dict(fn_module.__dict__)
    exec(compile(src.replace(old, new), f"<patched {fn_name}>", "exec"), ns)
    return ns[fn_name]


def ref_classifier(R, model, input_sz, device="cpu", name="synthetic"):
    """The reference's Classifier (classifier.py:12-75) around a given torchvision module: built through __new__ because __init__ downloads
    the pretrained weights (classifier.py:36); fields as classifier.py:15-18,22-33,47-51."""
    import torch
    from torchvision import transforms as T
    c = R.classifier.Classifier.__new__(R.classifier.Classifier)
    c.name, c.fix_params, c.device, c.sort_results = name, True, torch.device(device), True
    c.input_sz, c.model = tuple(input_sz), model
    nz = T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
    c.normalize = T.Lambda(lambda x: torch.stack([nz(x[i]) for i in range(x.shape[0])], 0))
    return c


def ref_pcnet(R, P, cam_hw, device="cpu", use_rough=True):
    """The reference's PCNet built the way train_eval_pcnet does (train_network.py:536-552: DataParallel-wrapped parts), loaded with `P`."""
    import torch.nn as nn
    wn = R.models.WarpingNet(out_size=tuple(cam_hw))
    sn = R.models.ShadingNetSPAA(use_rough=use_rough)
    m = R.models.PCNet(P["mask"], nn.DataParallel(wn), nn.DataParallel(sn), use_rough=use_rough)
    m.load_state_dict(P, strict=True)
    return m.to(device)


def ref_compennet_pp(R, P, prj_hw, device="cpu"):
    import torch.nn as nn
    m = R.models.CompenNetPlusplus(nn.DataParallel(R.models.WarpingNet(out_size=tuple(prj_hw))), nn.DataParallel(R.models.CompenNet()))
    m.load_state_dict(P, strict=True)
    return m.to(device)


class _Stop(Exception):
    pass


def time_reference_spaa(R, pcnet, classifier, targets, scene, d_thr, stealth_loss, device, setup, *, warmup, steps, budget_s=None,
                        labels=None):
    """Runs the reference's `spaa` (projector_based_attack.py:212-339, `iters` literal patched to warmup + steps) and timestamps the start
    of every iteration (the PCNet forward call, :265).  Returns (seconds per timed iteration, iterations timed).  `budget_s`: stop early
    (after at least 2 timed iterations) once that much wall time has passed -- a bounded sample for the CPU arm."""
    import torch
    dev = torch.device(device)
    stamps = []
    t_begin = time.perf_counter()

    def pc(x, s):
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        now = time.perf_counter()
        if budget_s is not None and len(stamps) >= warmup + 2 and now - t_begin > budget_s:
            raise _Stop()
        stamps.append(now)
        return pcnet(x, s)
    fn = patched(R.pba, "spaa", "iters = 50", f"iters = {int(warmup + steps)}")
    labels = labels if labels is not None else {i: f"class{i}" for i in range(1000)}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            fn(pc, classifier, labels, list(targets), True, scene, d_thr, stealth_loss, dev, setup)
    except _Stop:
        pass
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    stamps.append(time.perf_counter())
    per = [stamps[i + 1] - stamps[i] for i in range(warmup, len(stamps) - 1)]
    return sum(per) / len(per), len(per)


def time_reference_train_step(R, model, train_data, cfg, *, warmup, steps, iter_offset=401):
    """Runs the reference's `train_pcnet` (train_network.py:235-363; its `iters = 0` start patched to `iter_offset`, 401 = the L1+SSIM phase,
    :300-303) for warmup + steps iterations and timestamps every step's forward call.  `model` must be DataParallel-wrapped (the optimiser
    groups are selected by the 'module.' prefix, :248-250).  Returns seconds per timed step."""
    import torch
    dev = torch.device(cfg.device)
    stamps = []

    def hook(_m, _inp):
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        stamps.append(time.perf_counter())
    h = model.register_forward_pre_hook(hook)
    cfg.max_iters = iter_offset + warmup + steps
    fn = patched(R.train_network, "train_pcnet", "iters = 0", f"iters = {int(iter_offset)}")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            fn(model, train_data, None, cfg)
    finally:
        h.remove()
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    stamps.append(time.perf_counter())              # (includes the final checkpoint write: excluded by dropping the last interval below)
    per = [stamps[i + 1] - stamps[i] for i in range(warmup, warmup + steps - 1)]
    return sum(per) / len(per)
