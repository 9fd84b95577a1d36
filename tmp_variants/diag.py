import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, matgen
import modcholesky_b200 as m
from oracle import oracle
for n in (8, 16, 32, 64):
    for fam in ("bench", "uniform", "wishart", "pd"):
        a = matgen.make(fam, 3000 + n, 512 if n <= 32 else 64, n)
        ref = oracle.factor_batched("gmw81", a)
        d = m.factor_batched("gmw81", torch.from_numpy(a).cuda(), mode="fast")
        e = d.e.cpu().numpy()
        anorm = np.abs(a).max(axis=(1, 2))
        de = np.abs(e - ref.e).max(axis=1)
        same = (d.p.cpu().numpy() == ref.p).all(axis=1)
        print(n, fam, "max de/(eps*anorm) = %.2f" % (de[same] / (2.2e-16 * anorm[same])).max(), "escale/anorm min = %.2e" % (np.abs(ref.e).max(axis=1) / anorm).min(), "same_p", same.mean())
