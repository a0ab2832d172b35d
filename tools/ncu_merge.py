"""K4 / bandwidth kernels alone (for ncu): python tools/ncu_merge.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from art_sbir_b200 import ops  # noqa: E402

lists, nq, k = 8, 100_000, 10
d = torch.sort(torch.rand(lists, nq, k, device="cuda"), dim=2).values.contiguous()
i = torch.randint(0, 10_000_000, (lists, nq, k), device="cuda")
for _ in range(3):
    out = ops.topk_merge(d, i)
x = torch.randn(2_000_000, 1024, device="cuda")
buf = ops.GalleryBuffer(2_000_000, 1024, torch.bfloat16)
buf.append(x)
y = ops.l2_normalize(x)
torch.cuda.synchronize()
print("ok", out[0].shape, buf.filled, y.shape)
