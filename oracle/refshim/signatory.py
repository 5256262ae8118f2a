"""Stand-in for the third-party ``signatory`` module (absent on this image).

Only used when importing the *reference* inside this container
(oracle/load_reference.py); the reference's ``import signatory`` at
utils/summarizers.py:14-17 is not guarded against ImportError.  ``signature``
delegates to the float64 restatement in oracle/signature_np.py and returns a
tensor of the input's dtype/device, which is what signatory does.
"""
import importlib.util
import os

import torch

_here = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location(
    '_oracle_signature_np', os.path.join(_here, '..', 'signature_np.py'))
_sig = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_sig)


def signature(path, depth, **kwargs):
    if kwargs:
        raise NotImplementedError('shim supports signatory defaults only')
    out = _sig.signature(path.detach().cpu().double().numpy(), depth)
    return torch.from_numpy(out).to(dtype=path.dtype, device=path.device)
