import os, sys, torch
sys.path.insert(0, '/root/repo')
from bayes_sim_ig.utils import summarizers as S
dev = torch.device('cuda', 0)
n = 1 << 20
g = torch.Generator('cpu').manual_seed(0)
s = (torch.randn(n, 21, 4, generator=g) * 0.3).to(dev)
a = torch.rand(n, 21, 1, generator=g).to(dev)
for _ in range(3):
    out = S.summary_signatory(s, a)
torch.cuda.synchronize()
print(out.shape)
