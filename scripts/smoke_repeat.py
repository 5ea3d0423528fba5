"""The fp32 smoke case repeated in one process: is the parameter-gradient error against the CPU oracle stable?
(run plain, and once under `compute-sanitizer --tool initcheck`)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
from gan_control_b200 import kernels, modules as M  # noqa: E402
from oracle import params as P, stylegan2_oracle as O  # noqa: E402

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for i in range(n):
    e_img, e_pred, e_grad, g, d = G._smoke_case(torch, kernels, M, P, O, dev, 16, 64, 3, 2, 4, torch.float32, 11)
    # which parameters differ, and by how much
    print(f'run {i}: img {e_img:.2e} pred {e_pred:.2e} dG/dw(conv1) {e_grad:.2e}', flush=True)
