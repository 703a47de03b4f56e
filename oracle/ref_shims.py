"""Import shims for the *reference* modules (test infrastructure only).

Only ``oracle/make_golden.py`` uses this file, and only inside the build container
where ``/root/reference`` is mounted.  Nothing on the product path (``worddiffusion_b200``)
and nothing that runs on the GPU box imports it.

The three shims are the ones SURVEY.md section 8c lists:
  1. a stub ``omegaconf.listconfig.ListConfig`` (reference unet.py:1168, unetPhosc.py:818),
  2. ``open()`` of the hard-coded ``cropStyleDict_Numpy.pkl`` returns a pickled ``{}``
     (reference unet.py:1159-1161; the dict is never read by forward),
  3. ``logging.FileHandler`` on the ``/cluster`` path is neutralised (unetPhosc2.py:18-26).
"""
import builtins
import contextlib
import io
import logging
import os
import pickle
import sys
import types
from types import SimpleNamespace

REF_DIR = os.environ.get("WD_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "unet.py"))


def _install_omegaconf_stub():
    if "omegaconf" in sys.modules:
        return
    om = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass

    lc.ListConfig = ListConfig
    om.listconfig = lc
    sys.modules["omegaconf"] = om
    sys.modules["omegaconf.listconfig"] = lc


@contextlib.contextmanager
def _patched_open():
    real_open = builtins.open

    def fake_open(path, *a, **k):
        if isinstance(path, str) and path.endswith("cropStyleDict_Numpy.pkl"):
            return io.BytesIO(pickle.dumps({}))
        return real_open(path, *a, **k)

    builtins.open = fake_open
    try:
        yield
    finally:
        builtins.open = real_open


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def default_args(device="cpu", **over):
    ns = SimpleNamespace(device=device, interpolation=False, charLevelEmb=0, charImages=0,
                         attentionMaps=0, ocrTraining=0, imgConditioned=0, wrdChrWrStyl=0,
                         phosc=1, phos=0)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def import_reference(name: str):
    """name in {'unet', 'unetPhosc', 'unetPhosc2'}."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_DIR}")
    sys.dont_write_bytecode = True
    _install_omegaconf_stub()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if name == "unetPhosc2":
        real_fh = logging.FileHandler
        logging.FileHandler = lambda *a, **k: logging.NullHandler()
        try:
            with _quiet():
                mod = __import__(name)
        finally:
            logging.FileHandler = real_fh
        return mod
    with _quiet():
        return __import__(name)


MODEL_KW = dict(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4,
                num_res_blocks=1, attention_resolutions=(1, 1), channel_mult=(1, 1),
                num_heads=4, num_classes=339, context_dim=320, vocab_size=53, max_seq_len=10)


def build_reference_model(variant: str, args=None, **kw):
    """variant: 'unet' -> unet.UNetModel ; 'unetPhosc'/'unetPhosc2' -> UNetModelPhosc."""
    mod = import_reference(variant)
    args = args or default_args()
    k = dict(MODEL_KW)
    k.update(kw)
    with _patched_open(), _quiet():
        if variant == "unet":
            m = mod.UNetModel(args=args, **k)
        else:
            m = mod.UNetModelPhosc(args=args, **k)
    return m.eval()
