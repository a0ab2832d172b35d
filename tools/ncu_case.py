"""One retrieval pass of a given shape (for ncu): python tools/ncu_case.py NQ NG D DTYPE K [rank] [iters] [option=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from art_sbir_b200 import _binding, ops  # noqa: E402

_binding.set_debug_option("k1_pair_coop", 0)   # Nsight Compute cannot replay cooperative cluster launches
from tools.gpu_probe import _clustered  # noqa: E402

for a in [a for a in sys.argv if "=" in a]:   # library tuning switches for A/B captures, e.g. k1_bands=2
    sys.argv.remove(a)
    _binding.set_debug_option(a.split("=")[0], int(a.split("=")[1]))
nq, ng, d, dtype, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), getattr(torch, sys.argv[4]), int(sys.argv[5])
rank = len(sys.argv) > 6 and sys.argv[6] == "rank"
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 2
q, g, pos = _clustered(nq, ng, d, dtype)
for _ in range(iters):
    out = ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos if rank else None)
torch.cuda.synchronize()
print("ok", [t.shape for t in out])
