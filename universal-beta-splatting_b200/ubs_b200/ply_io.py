"""On-disk PLY format of a trained model <-> the packed record buffer.

Mirrors BetaModel.save_ply / load_ply (scene/beta_model.py:286-404): one `vertex` element, every property a
little-endian `float`, in the order  x y z | red green blue | opacity | beta_* | mean_* | scale_* | l_triangle_*
(construct_list_of_attributes, :286-296) holding the RAW (pre-activation) parameters.  The reference goes through
the third-party `plyfile` package (absent here); the binary layout it produces for an all-f4 element is a fixed
header followed by N rows of floats, which is what is written and parsed below.  Like load_ply, the reader finds
properties by NAME (any order, extra properties ignored) and infers input_dim from the number of `mean_*` columns.
"""
import os
from typing import Tuple

import numpy as np
import torch

from .fused import pack_records, record_slices, record_stride, unpack_records


def attribute_names(D: int):
    names = ["x", "y", "z", "red", "green", "blue", "opacity"]
    names += ["beta_%d" % i for i in range(D - 2)]
    names += ["mean_%d" % i for i in range(D - 3)]
    names += ["scale_%d" % i for i in range(D)]
    names += ["l_triangle_%d" % i for i in range(D * (D - 1) // 2)]
    return names


def _file_columns(D: int):
    """record column of every PLY property, in file order."""
    sl = record_slices(D)
    cols = []
    for group in ("xyz", "rgb", "opacity", "beta", "mean", "scale", "l_triangle"):
        cols += list(range(sl[group].start, sl[group].stop))
    return cols


def save_ply(path: str, records: torch.Tensor, D: int) -> None:
    """records [N, stride] (any device) -> binary little-endian PLY, byte-compatible with BetaModel.save_ply."""
    assert records.dim() == 2 and records.shape[1] == record_stride(D)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    cols = torch.tensor(_file_columns(D), device=records.device)
    rows = records.detach().index_select(1, cols).to("cpu", torch.float32).contiguous().numpy().astype("<f4")
    header = ["ply", "format binary_little_endian 1.0", "element vertex %d" % rows.shape[0]]
    header += ["property float %s" % n for n in attribute_names(D)]
    header += ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rows.tobytes())


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def _read_vertex_table(path: str):
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, props, n_vertex, in_vertex, seen_vertex = None, [], None, False, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: truncated PLY header" % path)
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if seen_vertex and tok[1] != "vertex":
                    in_vertex = False
                    continue
                if tok[1] != "vertex":
                    raise ValueError("%s: first element must be `vertex` (got %s)" % (path, tok[1]))
                in_vertex, seen_vertex, n_vertex = True, True, int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("%s: list properties are not supported in the vertex element" % path)
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if n_vertex is None:
            raise ValueError("%s: no vertex element" % path)
        if fmt in ("binary_little_endian", "binary_big_endian"):
            e = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(n, e + t) for n, t in props])
            table = np.fromfile(f, dtype=dt, count=n_vertex)
            if table.shape[0] != n_vertex:
                raise ValueError("%s: expected %d vertices, file holds %d" % (path, n_vertex, table.shape[0]))
        elif fmt == "ascii":
            raw = np.loadtxt(f, dtype=np.float64, max_rows=n_vertex, ndmin=2)
            dt = np.dtype([(n, "f8") for n, _ in props])
            table = np.zeros(n_vertex, dtype=dt)
            for k, (n, _) in enumerate(props):
                table[n] = raw[:, k]
        else:
            raise ValueError("%s: unknown PLY format %r" % (path, fmt))
    return table


def load_ply(path: str, device="cuda") -> Tuple[torch.Tensor, int]:
    """PLY written by the reference (or by save_ply) -> (records [N, stride] on `device`, D)."""
    table = _read_vertex_table(path)
    names = table.dtype.names

    def family(prefix):
        cols = sorted((n for n in names if n.startswith(prefix)), key=lambda s: int(s.split("_")[-1]))
        return np.stack([np.asarray(table[n], dtype=np.float32) for n in cols], axis=1) if cols else \
            np.zeros((table.shape[0], 0), np.float32)

    xyz = np.stack([np.asarray(table[n], dtype=np.float32) for n in ("x", "y", "z")], axis=1)
    rgb = np.stack([np.asarray(table[n], dtype=np.float32) for n in ("red", "green", "blue")], axis=1)
    opacity = np.asarray(table["opacity"], dtype=np.float32)[:, None]
    mean, beta, scale, l_tri = family("mean_"), family("beta_"), family("scale_"), family("l_triangle")
    D = mean.shape[1] + 3
    if beta.shape[1] != D - 2 or scale.shape[1] != D or l_tri.shape[1] != D * (D - 1) // 2:
        raise ValueError("%s: inconsistent attribute counts for input_dim %d: beta %d, scale %d, l_triangle %d" % (
            path, D, beta.shape[1], scale.shape[1], l_tri.shape[1]))
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in (xyz, mean, rgb, opacity, beta, scale, l_tri)]
    return pack_records(D, *t), D


def records_to_tensors(records: torch.Tensor, D: int):
    """Copies of the seven BetaModel tensors (xyz, mean, rgb, opacity [N,1], beta, scale, l_triangle)."""
    return tuple(t.clone() for t in unpack_records(D, records))
