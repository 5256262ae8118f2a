"""Import the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Source of the modules, first that exists: $BSIG_REFERENCE_ROOT, /root/reference (build
container only), oracle/_ref (the verbatim staged copy made by oracle/stage_reference.py,
git-ignored, shipped to the GPU box with the snapshot).  Used by
tests/golden/make_golden.py to produce the committed golden vectors, by the optional
live-reference tests (skipped when no source is present) and by bench.py's CPU legs
(``--impl reference`` / ``cpu_baseline``, kind "reference").

The reference package is called ``bayes_sim_ig`` -- the same name as this
repository's drop-in alias package -- so it is loaded under the private name
``_bsig_reference`` (its modules use relative imports only, so the alias is
transparent to it).  ``signatory`` and ``ghalton`` are supplied by
oracle/refshim (see the notes there).
"""
import importlib
import importlib.machinery
import os
import sys
import types

def _pick_root():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get('BSIG_REFERENCE_ROOT'), '/root/reference',
                 os.path.join(here, '_ref')):
        if cand and os.path.isfile(os.path.join(cand, 'bayes_sim_ig', 'bayes_sim.py')):
            return cand
    return os.environ.get('BSIG_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _pick_root()
_ALIAS = '_bsig_reference'


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'bayes_sim_ig', 'bayes_sim.py'))


def _ensure_alias():
    if _ALIAS in sys.modules:
        return
    shim_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'refshim')
    if shim_dir not in sys.path:
        sys.path.append(shim_dir)   # appended: never shadows a real install
    for sub in ('', '.models', '.utils'):
        name = _ALIAS + sub
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REFERENCE_ROOT, 'bayes_sim_ig',
                                     *sub.strip('.').split('.'))] if sub else \
            [os.path.join(REFERENCE_ROOT, 'bayes_sim_ig')]
        mod.__package__ = name
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        mod.__spec__.submodule_search_locations = mod.__path__
        sys.modules[name] = mod


def load(submodule):
    """load('bayes_sim') / load('models.mdnn') / load('utils.pdf') ..."""
    if not reference_available():
        raise RuntimeError('reference not present at ' + REFERENCE_ROOT)
    _ensure_alias()
    return importlib.import_module(_ALIAS + '.' + submodule)
