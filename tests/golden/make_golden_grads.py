"""Generate tests/golden/mask3d_tiny2_grads.npz: gradient signatures of the reference's UNMODIFIED Mask3D + Res16UNet34C
(imported from /root/reference/mask3d) on the CPU oracle, for the scalar `grad_loss` of tests/boundary_check.py on the scenes of
mask3d_tiny2.npz.  Per parameter: (l2 norm, sum, dot with a fixed cosine pattern) in float64.  Build container only:

    python tests/golden/make_golden_grads.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_golden as MG  # noqa: E402  (installs the oracle shims + the reference path, builds the reference model)

ME, measure_runtime = MG.ME, MG.measure_runtime


def main():
    fix = np.load(os.path.join(ROOT, "tests", "golden", "mask3d_tiny2.npz"))
    m = MG.build_reference_model()
    x = ME.SparseTensor(coordinates=torch.from_numpy(fix["coords"]), features=torch.from_numpy(fix["feats"]))
    measure_runtime.notify_start_item()
    out = m(x, is_eval=True)
    loss = 0.0
    for pm, pc in zip(out.pred_mask_all_decoders, out.pred_class_all_decoders):
        loss = loss + torch.cat(pm).square().mean() + pc.square().mean()
    loss.backward()
    names, sig = [], []
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().double().reshape(-1)
        r = torch.cos(torch.arange(g.numel(), dtype=torch.float64) * 0.37 + 0.11)
        names.append(n)
        sig.append([float(g.norm()), float(g.sum()), float((g * r).sum())])
    path = os.path.join(ROOT, "tests", "golden", "mask3d_tiny2_grads.npz")
    np.savez_compressed(path, names=np.array(names), sig=np.array(sig, dtype=np.float64), loss=float(loss))
    print("wrote", path, len(names), "parameters, loss", float(loss))


if __name__ == "__main__":
    main()
