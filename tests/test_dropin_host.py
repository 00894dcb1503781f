"""CPU tests of the zero-edit drop-in route's host logic (ubs_b200/dropin.py, ubs_b200/rendering.py):

  * the caller harness tests/ref_caller.py really is the reference's caller: every restated method has the same AST as
    the method `ast` extracts from /root/reference/scene/beta_model.py (skipped where the reference tree is absent);
  * deferred tensors: metadata without computation, `[mask]` / `.squeeze()` stay deferred, anything else materialises;
  * the lazily-sized `meta` entries;
  * depth_to_normal against a fixture made with the reference's own function (tests/golden/make_golden_normals.py).
"""
import ast
import os
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODEL = "/root/reference/scene/beta_model.py"


def _methods(path, cls_name):
    tree = ast.parse(open(path).read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls_name)
    return {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}


def _norm(fn):
    fn = ast.parse(textwrap.dedent(ast.unparse(fn))).body[0]
    for n in ast.walk(fn):  # docstrings and comments carry no statements
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.body and isinstance(n.body[0], ast.Expr) and isinstance(
                getattr(n.body[0], "value", None), ast.Constant) and isinstance(n.body[0].value.value, str):
            n.body = n.body[1:] or [ast.Pass()]
    return ast.dump(fn, include_attributes=False)


@pytest.mark.skipif(not os.path.exists(REF_MODEL), reason="reference tree not present (GPU box)")
def test_caller_harness_has_the_reference_statements():
    ref = _methods(REF_MODEL, "BetaModel")
    mine = _methods(os.path.join(ROOT, "tests", "ref_caller.py"), "BetaModelCaller")
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ref_caller import BetaModelCaller

    assert "render" in BetaModelCaller.RESTATED and "get_cond_mean_convariance_opacity" in BetaModelCaller.RESTATED
    for name in BetaModelCaller.RESTATED:
        assert _norm(mine[name]) == _norm(ref[name]), "restated %s differs from scene/beta_model.py" % name
    # the rasterization(...) call of BetaModel.view (beta_model.py:797-814): same keyword set and value expressions
    def raster_call(fn):
        call = next(n for n in ast.walk(fn) if isinstance(n, ast.Call) and getattr(n.func, "id", "") == "rasterization")
        return {k.arg: ast.dump(k.value) for k in call.keywords}

    rv, mv = raster_call(ref["view"]), raster_call(mine["view_call"])
    renames = {"render_tab_state.near_plane": "near_plane", "render_tab_state.far_plane": "far_plane",
               "render_tab_state.radius_clip": "radius_clip"}
    assert set(rv) == set(mv)
    for k in rv:
        if k in ("near_plane", "far_plane", "radius_clip"):
            continue  # GUI state fields are plain arguments in the harness
        assert rv[k] == mv[k], k
    assert renames


def _fake_cond(N=7):
    from ubs_b200 import dropin

    calls = []

    def compute():
        calls.append(1)
        g = torch.Generator().manual_seed(0)
        return torch.randn(N, 3, generator=g), torch.randn(N, 3, 3, generator=g), torch.rand(N, 1, generator=g)

    node = dropin._Node("cond", (), compute)
    outs = (dropin.Deferred(node, 0, (N, 3), "cpu", True), dropin.Deferred(node, 1, (N, 3, 3), "cpu", True),
            dropin.Deferred(node, 2, (N, 1), "cpu", False))
    return node, outs, calls


def test_deferred_metadata_costs_nothing():
    node, (m, v, o), calls = _fake_cond()
    assert m.shape == (7, 3) and v.shape == (7, 3, 3) and o.shape == (7, 1)
    assert m.dim() == 2 and v.size(1) == 3 and len(o) == 7 and m.numel() == 21
    assert m.dtype == torch.float32 and m.device.type == "cpu" and not m.is_cuda
    assert m.requires_grad and not o.requires_grad
    assert "Deferred" in repr(m)
    assert calls == []


def test_mask_and_squeeze_stay_deferred_everything_else_materialises():
    from ubs_b200 import dropin

    node, (m, v, o), calls = _fake_cond()
    mask = torch.tensor([1, 0, 1, 1, 0, 1, 1], dtype=torch.bool)
    mm, oo, vv = m[mask], o.squeeze()[mask], v[mask]
    assert all(isinstance(t, dropin.Deferred) for t in (mm, oo, vv)) and calls == []
    assert [s[0] for s in oo._views] == ["squeeze", "mask"] and oo._views[1][1] is mask
    real_m, real_v, real_o = node.real()
    assert calls == [1]
    assert torch.equal(mm.materialize(), real_m[mask]) and torch.equal(oo.materialize(), real_o.squeeze()[mask])
    # arithmetic, reductions, torch.* functions, integer / slice indexing: all compute through the real tensors
    assert torch.equal(m + 1, real_m + 1) and torch.equal(torch.cat([o, o]), torch.cat([real_o, real_o]))
    assert torch.equal(m[2], real_m[2]) and torch.equal(v[:, 0], real_v[:, 0]) and float(o.sum()) == float(real_o.sum())
    assert torch.equal(m[mask].contiguous(), real_m[mask])
    assert calls == [1]  # computed once
    assert torch.equal(dropin.materialize((m, [o], {"k": v}))[2]["k"], real_v)


def test_rasterization_routing_rejects_foreign_tensors():
    from ubs_b200 import dropin

    node, (m, v, o), calls = _fake_cond()
    node2, (m2, v2, o2), _ = _fake_cond()
    vm, K = torch.eye(4)[None], torch.eye(3)[None]
    a = (torch.rand(7), torch.rand(7, 3), vm, K, 64, 48, 0.01, 1e10, 0.0, 0.3, 16, None, "RGB", "classic")
    # real tensors, outputs of two different calls, a non-squeezed opacity, different masks: no fused route
    assert dropin.try_fused_rasterization(torch.rand(7, 3), o.squeeze(), *a, v) is None
    assert dropin.try_fused_rasterization(m, o2.squeeze(), *a, v) is None
    assert dropin.try_fused_rasterization(m, o, *a, v) is None
    k1, k2 = torch.ones(7, dtype=torch.bool), torch.ones(7, dtype=torch.bool)
    assert dropin.try_fused_rasterization(m[k1], o.squeeze()[k2], *a, v[k1]) is None
    assert calls == []


def test_lazy_meta_sizes_on_first_access():
    from ubs_b200.dropin import LazyMeta

    hits = []
    meta = LazyMeta({"radii": 1, "n_cameras": 1})
    meta.defer("isect_ids", lambda: hits.append("i") or "IDS")
    meta.defer("flatten_ids", lambda: hits.append("f") or "FL")
    assert meta["radii"] == 1 and "isect_ids" in meta and hits == []
    assert meta["isect_ids"] == "IDS" and meta["isect_ids"] == "IDS" and hits == ["i"]
    assert dict(meta.items())["flatten_ids"] == "FL" and hits == ["i", "f"]
    assert set(meta) == {"radii", "n_cameras", "isect_ids", "flatten_ids"}


def test_depth_to_normal_matches_reference_fixture():
    from ubs_b200.rendering import depth_to_normal

    z = np.load(os.path.join(ROOT, "tests", "golden", "depth_to_normal.npz"))
    got = depth_to_normal(torch.from_numpy(z["depths"]), torch.from_numpy(z["c2w"]), torch.from_numpy(z["Ks"]))
    want = torch.from_numpy(z["normals"])
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=0, atol=1e-6)
    assert float(want[:, 1:-1, 1:-1].norm(dim=-1).min()) > 0.99 and float(want[:, 0].abs().max()) == 0.0
