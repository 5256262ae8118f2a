"""Debug aid: tcgen05 engine on the dgrad / wgrad forms (MN-major operands)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from bayes_sim_ig_b200 import _lib
DEV = 'cuda:0'
lib = _lib
def ws_for(m, n, k):
    return torch.empty(max(_lib.load().bsig_linear_ws_bytes(m, n, k), 16), dtype=torch.uint8, device=DEV)
def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())
for (m, n, k) in [(256, 128, 128), (4096, 128, 302), (300, 270, 128)]:
    g = torch.Generator('cpu').manual_seed(1)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(DEV)
    dy = torch.randn(m, n, generator=g).to(DEV)
    h = torch.tanh(torch.randn(m, k, generator=g)).to(DEV)
    ws = ws_for(m, n, k)
    st = _lib.stream_ptr(DEV)
    for eng in (1, 2):
        for act in (0, 1):
            dx = torch.full((m, k), 7.0, device=DEV)
            _lib.call('bsig_linear_dgrad', dy.data_ptr(), w.data_ptr(), h.data_ptr(), dx.data_ptr(), m, n, k, act, eng, ws.data_ptr(), ws.numel(), st)
            torch.cuda.synchronize()
            ref = dy.double() @ w.double()
            if act: ref = ref * (1 - h.double() ** 2)
            print('dgrad', (m, n, k), 'eng', eng, 'act', act, 'rel', rel(dx, ref), 'dx[0,:4]', dx[0, :4].tolist(), 'ref', ref[0, :4].tolist())
        dw = torch.full((n, k), 7.0, device=DEV); db = torch.empty(n, device=DEV)
        _lib.call('bsig_linear_wgrad', dy.data_ptr(), x.data_ptr(), k, None, dw.data_ptr(), db.data_ptr(), m, n, k, eng, ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        ref = dy.double().T @ x.double()
        print('wgrad', (m, n, k), 'eng', eng, 'rel', rel(dw, ref), 'dw[0,:4]', dw[0, :4].tolist(), 'ref', ref[0, :4].tolist())
