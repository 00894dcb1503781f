"""PLY reader/writer of the packed record against the byte layout BetaModel.save_ply produces through plyfile
(scene/beta_model.py:286-321): header text + N rows of little-endian floats in construct_list_of_attributes order."""
import numpy as np
import pytest
import torch


def _reference_bytes(D, tensors):
    """What plyfile's PlyData([PlyElement.describe(elements, 'vertex')]).write() emits for the reference's dtype_full."""
    xyz, mean, rgb, opacity, beta, scale, l_tri = [t.numpy() for t in tensors]
    names = ["x", "y", "z", "red", "green", "blue", "opacity"]
    names += [f"beta_{i}" for i in range(beta.shape[1])] + [f"mean_{i}" for i in range(D - 3)]
    names += [f"scale_{i}" for i in range(scale.shape[1])] + [f"l_triangle_{i}" for i in range(l_tri.shape[1])]
    attributes = np.concatenate((xyz, rgb, opacity, beta, mean, scale, l_tri), axis=1).astype("<f4")
    head = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % xyz.shape[0]
    head += "".join("property float %s\n" % n for n in names) + "end_header\n"
    return head.encode("ascii") + attributes.tobytes()


@pytest.mark.parametrize("D", [6, 7])
def test_save_ply_is_byte_identical_to_reference_layout_and_round_trips(tmp_path, D):
    from ubs_b200 import fused, ply_io, synth

    scene = synth.make_scene(257, D, seed=3)
    tensors = [t.reshape(t.shape[0], -1) for t in scene.tensors()]
    rec = fused.pack_records(D, *tensors)
    p = tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply"
    ply_io.save_ply(str(p), rec, D)
    assert p.read_bytes() == _reference_bytes(D, tensors)
    back, D2 = ply_io.load_ply(str(p), device="cpu")
    assert D2 == D and torch.equal(back, rec)
    for a, b in zip(ply_io.records_to_tensors(back, D), tensors):
        assert torch.equal(a, b)


def test_load_ply_finds_properties_by_name_in_any_order_and_ascii(tmp_path):
    from ubs_b200 import fused, ply_io, synth

    D = 6
    scene = synth.make_scene(11, D, seed=4)
    tensors = [t.reshape(t.shape[0], -1) for t in scene.tensors()]
    rec = fused.pack_records(D, *tensors)
    names = ply_io.attribute_names(D)
    cols = ply_io._file_columns(D)
    order = list(reversed(range(len(names))))  # shuffled property order + one foreign property
    p = tmp_path / "a.ply"
    with open(p, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 11\nproperty float nx\n")
        for k in order:
            f.write("property float %s\n" % names[k])
        f.write("end_header\n")
        for r in range(11):
            f.write("0 " + " ".join(repr(float(rec[r, cols[k]])) for k in order) + "\n")
    back, D2 = ply_io.load_ply(str(p), device="cpu")
    assert D2 == D and torch.equal(back, rec)


def test_load_ply_rejects_inconsistent_files(tmp_path):
    from ubs_b200 import ply_io

    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float x\nproperty float y\n"
                  b"property float z\nproperty float red\nproperty float green\nproperty float blue\n"
                  b"property float opacity\nproperty float mean_0\nend_header\n")
    with pytest.raises(ValueError):
        ply_io.load_ply(str(p), device="cpu")
    q = tmp_path / "nope.ply"
    q.write_bytes(b"plx\n")
    with pytest.raises(ValueError):
        ply_io.load_ply(str(q), device="cpu")
