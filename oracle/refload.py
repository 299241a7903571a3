"""Loader for the UNMODIFIED reference env classes (jviquerat/beacon).

TEST INFRASTRUCTURE ONLY. Used here (build container) to pin the C/numpy oracle
against the real reference and to generate tests/golden/*.npz. `/root/reference`
does not exist on the GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py`
may depend on this module at run time.

Recipe: stub `gymnasium`/`matplotlib` first on sys.path,
chdir into beacon/<env>/ because `init_field.dat` is CWD-relative
(shkadov.py:50,93; rayleigh.py:41,72; sloshing.py:34,70).
"""
import contextlib
import importlib
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
# the checkout in the build container, else the copy staged by oracle/make_ref.py (travels to the GPU box for
# the CPU-baseline arm of bench.py only; tests never read it)
REF_ROOT = os.environ.get("BEACON_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/beacon") else os.path.join(_HERE, "_ref"))
_SHIM = os.path.join(_HERE, "refshim")

_MODULE_OF = {"shkadov_separable": "shkadov"}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "beacon"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_module(env_name):
    """Import the reference module that defines `env_name` (cached by Python)."""
    mod_name = _MODULE_OF.get(env_name, env_name)
    env_dir = os.path.join(REF_ROOT, "beacon", mod_name)
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if env_dir not in sys.path:
        sys.path.insert(1, env_dir)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(mod_name)


def make(env_name, **kwargs):
    """Instantiate a reference env (constructor runs inside its own directory)."""
    mod = load_module(env_name)
    mod_name = _MODULE_OF.get(env_name, env_name)
    with _cwd(os.path.join(REF_ROOT, "beacon", mod_name)):
        env = getattr(mod, env_name)(**kwargs)
    return env, mod


class NoiseFeeder:
    """Replacement for `np.random.uniform` inside a reference module: pops
    pre-generated numbers so reference, oracle and CUDA consume identical noise
    (shkadov.py:204 draws one per sub-step, burgers.py:127 one per action)."""

    def __init__(self, values):
        self.values = list(values)
        self.pos = 0

    def __call__(self, low, high, size=None):
        import numpy as np
        v = self.values[self.pos]
        self.pos += 1
        return np.array([v]) if size is not None else v


@contextlib.contextmanager
def patched_noise(mod, values):
    """Temporarily route `mod.np.random.uniform` through a NoiseFeeder."""
    import numpy as np
    feeder = NoiseFeeder(values)
    orig = np.random.uniform
    np.random.uniform = feeder
    try:
        yield feeder
    finally:
        np.random.uniform = orig
