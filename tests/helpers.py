"""Shared helpers for the GPU parity tests."""
import contextlib

import numpy as np
import torch


@contextlib.contextmanager
def injected_rand_like(draws, device):
    """Make torch.rand_like return the recorded uniforms, in order."""
    queue = [torch.as_tensor(np.asarray(d), dtype=torch.float32).to(device) for d in draws]
    orig = torch.rand_like

    def fake(t, *a, **k):
        out = queue.pop(0)
        assert out.numel() == t.numel(), (out.shape, t.shape)
        return out.reshape(t.shape)
    torch.rand_like = fake
    try:
        yield queue
    finally:
        torch.rand_like = orig


def load_state(model, arrays):
    sd = {k: torch.from_numpy(np.asarray(v)).float() for k, v in arrays.items()}
    missing = model.load_state_dict(sd, strict=True)
    return missing


def mdn_meta(g, case):
    meta = g[case + '.meta']
    din, p, k, full, b = [int(v) for v in meta[:5]]
    hidden = tuple(int(v) for v in meta[5:])
    return din, p, k, bool(full), b, hidden


def rel_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30)


def synth_rollouts(seed, n, t1, d, a, device=None):
    """The seeded synthetic rollout generator shared by tests and bench
    (SURVEY 8.d): states ~ N(0,1) clamped to +-100, actions ~ U[0,1)."""
    g = torch.Generator('cpu').manual_seed(seed)
    states = torch.randn(n, t1, d, generator=g).clamp_(-100, 100)
    actions = torch.rand(n, t1, a, generator=g)
    if device is not None:
        states, actions = states.to(device), actions.to(device)
    return states, actions


def seeded_uniforms(call_index, shape):
    """The i-th torch.rand_like draw of a replayed run: same generator as
    tests/golden/make_golden.py::seeded_uniforms (golden_predict_multi)."""
    g = torch.Generator('cpu').manual_seed(424242 + int(call_index))
    return torch.rand(tuple(shape), generator=g)
