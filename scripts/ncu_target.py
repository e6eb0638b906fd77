# coding: utf-8
"""One synthesis launch of BASELINE config 2 for ncu:  python scripts/ncu_target.py [T] [B]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 22050
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = bench.build_model().cuda()
eng = m._get_engine()
c = torch.randn(B, T, 80, device="cuda")
eng.generate(B=B, T=T, c=c, seed=0)
torch.cuda.synchronize()
print("done", T, B, eng.plan(B))
