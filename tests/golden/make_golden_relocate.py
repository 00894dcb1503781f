"""Generates tests/golden/relocate_D{6,7}.npz by running the REFERENCE's own MCMC relocation code on CPU tensors.

scene/beta_model.py cannot be imported in the build container (plyfile, fused_ssim and the CUDA extension are
missing), but the methods on this path -- relocate_gs, add_new_gs, _update_params, _sample_alives,
replace_tensors_to_optimizer, cat_tensors_to_optimizer, densification_postfix (scene/beta_model.py:446-657) -- are pure
torch.  This script takes their SOURCE TEXT out of /root/reference/scene/beta_model.py with `ast`, unmodified,
compiles them into a stub class that only supplies what BetaModel.__init__ / setup_functions would (the seven
parameter tensors, the optimizer with its seven named groups, get_opacity = sigmoid, inverse_opacity_activation =
utils/general_utils.py:21-22), and records inputs, the indices torch.multinomial drew, and outputs.  Nothing of the
reference is copied into the repository: the fixture holds tensors only.

    python tests/golden/make_golden_relocate.py
"""
import ast
import os

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scene/beta_model.py"
METHODS = ("replace_tensors_to_optimizer", "cat_tensors_to_optimizer", "densification_postfix", "_update_params",
           "_sample_alives", "relocate_gs", "add_new_gs")
GROUPS = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")


def reference_methods():
    text = open(SRC).read()
    tree = ast.parse(text)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BetaModel")
    lines = text.splitlines()
    body = []
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in METHODS:
            body.append("\n".join(lines[fn.lineno - 1:fn.end_lineno]))
    assert len(body) == len(METHODS), "reference methods not found"
    ns = {"torch": torch, "nn": nn}
    exec("class Ref:\n" + "\n\n".join(body), ns)
    return ns["Ref"]


def make_model(Ref, D, N, seed):
    g = torch.Generator().manual_seed(seed)
    widths = (3, D - 3, 3, 1, D - 2, D, D * (D - 1) // 2)
    m = Ref()
    tensors = [nn.Parameter(0.5 * torch.randn(N, w, generator=g)) for w in widths]
    with torch.no_grad():
        tensors[3][torch.rand(N, generator=g) < 0.12] = -8.0  # dead primitives: sigmoid(-8) <= 0.005
    m._xyz, m._mean, m._rgb, m._opacity, m._beta, m._scale, m._l_triangle = tensors
    m.optimizer = torch.optim.Adam([{"params": [p], "lr": 1e-3, "name": n} for n, p in zip(GROUPS, tensors)],
                                   lr=0.0, eps=1e-15)
    for p in tensors:  # one step so that every group has Adam state
        p.grad = torch.randn(p.shape, generator=g)
    m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    type(m).get_opacity = property(lambda self: torch.sigmoid(self._opacity))           # scene/beta_model.py:46,117
    m.inverse_opacity_activation = lambda x: torch.log(x / (1 - x))                     # utils/general_utils.py:21-22
    m.sampled = []
    orig = type(m)._sample_alives

    def recording(self, probs, num, alive_indices=None):
        idx, ratio = orig(self, probs, num, alive_indices)
        self.sampled.append(idx.clone())
        return idx, ratio

    type(m)._sample_alives = recording
    return m


def snapshot(m, tag, out):
    for n, p in zip(GROUPS, (m._xyz, m._mean, m._rgb, m._opacity, m._beta, m._scale, m._l_triangle)):
        st = m.optimizer.state[p]
        out[f"{tag}_{n}"] = p.detach().numpy().copy()
        out[f"{tag}_{n}_m"] = st["exp_avg"].numpy().copy()
        out[f"{tag}_{n}_v"] = st["exp_avg_sq"].numpy().copy()


if __name__ == "__main__":
    for D in (6, 7):
        Ref = reference_methods()
        torch.manual_seed(100 + D)  # torch.multinomial inside the reference code draws from the global stream
        m = make_model(Ref, D, 400, 7 + D)
        out = {}
        snapshot(m, "in", out)
        with torch.no_grad():
            dead = (torch.sigmoid(m._opacity) <= 0.005).squeeze(-1)  # train.py:155
            out["dead_mask"] = dead.numpy()
            m.relocate_gs(dead_mask=dead)
            out["reinit_idx"] = m.sampled[-1].numpy()
            snapshot(m, "relocated", out)
            added = m.add_new_gs(cap_max=100000)
            out["add_idx"] = m.sampled[-1].numpy()
            out["n_added"] = np.int64(added)
            snapshot(m, "grown", out)
        np.savez_compressed(os.path.join(HERE, f"relocate_D{D}.npz"), **out)
        print(D, "dead", int(dead.sum()), "max multiplicity", int(torch.bincount(m.sampled[0]).max()), "added", added)
