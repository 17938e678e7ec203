"""On-disk formats of street_gaussians, without the `plyfile` dependency:

  * the multi-element PLY of StreetGaussianModel.save_ply / load_ply (lib/models/street_gaussian_model.py:94-117): one element
    `vertex_<model_name>` per sub-model, written by plyfile as binary_little_endian float32 properties in the order of
    GaussianModel.construct_list_of_attributes (lib/models/gaussian_model.py:327-342):
        x y z nx ny nz  f_dc_0..  f_rest_0..  opacity  scale_0..2  rot_0..3  semantic_0..
    with f_dc / f_rest stored channel-major (`_features_dc.transpose(1, 2).flatten(1)`, gaussian_model.py:83-84) and read back
    as reshape(n, 3, -1).transpose(1, 2) (:121-122, 148-149); the single-model file of GaussianModel.save_ply uses element `vertex`;
  * the `.pth` checkpoints of train.py:218-223: torch.save of {model_name: GaussianModel.state_dict(), ..., 'iter': n} with the
    keys of gaussian_model.py:180-205.

The reference does this in Python over numpy; so does this module (host-side IO, not a kernel).  Tensors are returned on the
CPU; callers move them to the device and wrap them in nn.Parameter as load_ply does.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping

import numpy as np
import torch

RAW = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation", "semantic")


def _get(model, key):
    if isinstance(model, Mapping):
        return model[key] if key in model else model["_" + key]
    return getattr(model, "_" + key)


def attribute_names(model) -> list:
    """GaussianModel.construct_list_of_attributes (gaussian_model.py:327-342)."""
    fdc, frest = _get(model, "features_dc"), _get(model, "features_rest")
    sem = _get(model, "semantic") if _has(model, "semantic") else None
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(fdc.shape[1] * fdc.shape[2])]
    names += [f"f_rest_{i}" for i in range(frest.shape[1] * frest.shape[2])]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(_get(model, "scaling").shape[1])]
    names += [f"rot_{i}" for i in range(_get(model, "rotation").shape[1])]
    if sem is not None:
        names += [f"semantic_{i}" for i in range(sem.shape[1])]
    return names


def _has(model, key) -> bool:
    if isinstance(model, Mapping):
        return key in model or ("_" + key) in model
    return hasattr(model, "_" + key)


def make_ply(model) -> np.ndarray:
    """[n, n_attributes] float32 rows in the reference's attribute order (GaussianModel.make_ply, gaussian_model.py:80-96)."""
    c = lambda t: t.detach().cpu().float()
    xyz = c(_get(model, "xyz"))
    n = xyz.shape[0]
    f_dc = c(_get(model, "features_dc")).transpose(1, 2).flatten(start_dim=1)
    f_rest = c(_get(model, "features_rest")).transpose(1, 2).flatten(start_dim=1)
    sem = c(_get(model, "semantic")) if _has(model, "semantic") else torch.zeros(n, 0)
    rows = torch.cat((xyz, torch.zeros_like(xyz), f_dc, f_rest, c(_get(model, "opacity")), c(_get(model, "scaling")),
                      c(_get(model, "rotation")), sem.reshape(n, -1)), dim=1)
    return np.ascontiguousarray(rows.numpy(), dtype="<f4")


def save_ply(path: str, models: Mapping[str, object], single_element: bool = False) -> None:
    """models: {model_name: model}.  Elements are named `vertex_<model_name>` (street model) or `vertex` (single_element=True,
    GaussianModel.save_ply, gaussian_model.py:98-102)."""
    if single_element and len(models) != 1:
        raise ValueError("single_element=True writes exactly one model")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    header, blobs = ["ply", "format binary_little_endian 1.0"], []
    for name, m in models.items():
        rows = make_ply(m)
        names = attribute_names(m)
        assert rows.shape[1] == len(names)
        header.append(f"element {'vertex' if single_element else 'vertex_' + name} {rows.shape[0]}")
        header += [f"property float {a}" for a in names]
        blobs.append(rows.tobytes())
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        for b in blobs:
            f.write(b)


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
              "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_ply_elements(path: str) -> Dict[str, np.ndarray]:
    """Minimal PLY reader (binary_little_endian and ascii, scalar properties): {element name: structured array}."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0].strip() != "ply":
        raise ValueError(f"{path} is not a PLY file")
    fmt, elements = None, []
    for ln in lines[1:]:
        tok = ln.split()
        if not tok or tok[0] == "comment" or tok[0] == "obj_info":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError("list properties are not used by street_gaussians files")
            elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
    out, off = {}, end
    if fmt == "binary_little_endian":
        for name, count, props in elements:
            dt = np.dtype(props)
            out[name] = np.frombuffer(data, dtype=dt, count=count, offset=off)
            off += dt.itemsize * count
    elif fmt == "ascii":
        rows = data[end:].decode("ascii").split("\n")
        at = 0
        for name, count, props in elements:
            arr = np.zeros(count, dtype=np.dtype([(p, t.lstrip("<")) for p, t in props]))
            for i in range(count):
                vals = rows[at + i].split()
                for (p, _), v in zip(props, vals):
                    arr[p][i] = float(v)
            at += count
            out[name] = arr
    else:
        raise ValueError(f"unsupported PLY format {fmt}")
    return out


def _cols(el: np.ndarray, prefix: str) -> np.ndarray:
    names = sorted((n for n in el.dtype.names if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    return np.stack([np.asarray(el[n], dtype=np.float64) for n in names], axis=1) if names else np.zeros((el.shape[0], 0))


def element_to_model(el: np.ndarray) -> Dict[str, torch.Tensor]:
    """GaussianModel.load_ply (gaussian_model.py:104-155): one PLY element -> raw parameter tensors (CPU, float32)."""
    n = el.shape[0]
    xyz = np.stack((np.asarray(el["x"]), np.asarray(el["y"]), np.asarray(el["z"])), axis=1)
    f = lambda a: torch.tensor(np.asarray(a), dtype=torch.float)
    fdc = f(_cols(el, "f_dc_").reshape(n, 3, -1)).transpose(1, 2).contiguous()
    frest = f(_cols(el, "f_rest_").reshape(n, 3, -1)).transpose(1, 2).contiguous()
    return dict(xyz=f(xyz), features_dc=fdc, features_rest=frest, opacity=f(np.asarray(el["opacity"])[..., None]), scaling=f(_cols(el, "scale_")),
                rotation=f(_cols(el, "rot_")), semantic=f(_cols(el, "semantic_")))


def load_ply(path: str) -> Dict[str, Dict[str, torch.Tensor]]:
    """{model_name: raw parameters}.  Element `vertex_<name>` -> key `<name>` (street_gaussian_model.py:107-115: name[7:]);
    a plain `vertex` element (single-model file) -> key ''."""
    out = {}
    for name, el in read_ply_elements(path).items():
        if name == "vertex" or name.startswith("vertex_"):
            out[name[7:]] = element_to_model(el)
    return out


# ---- .pth checkpoints (train.py:218-223; gaussian_model.py:157-205) ----
_STATE_KEYS = dict(xyz="xyz", feature_dc="features_dc", feature_rest="features_rest", scaling="scaling", rotation="rotation", opacity="opacity",
                   semantic="semantic")


def model_state_dict(model, extra: Mapping = None) -> dict:
    """GaussianModel.state_dict(is_final=True) keys; `extra` carries the training-time entries (spatial_lr_scale, denom, max_radii2D,
    xyz_gradient_accum, active_sh_degree, optimizer) when the checkpoint is not final."""
    sd = {k: _get(model, v) for k, v in _STATE_KEYS.items() if _has(model, v)}
    if extra:
        sd.update(extra)
    return sd


def save_checkpoint(path: str, models: Mapping[str, object], iteration: int, extras: Mapping[str, Mapping] = None, **other) -> None:
    sd = {name: model_state_dict(m, (extras or {}).get(name)) for name, m in models.items()}
    sd.update(other)
    sd["iter"] = int(iteration)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(sd, path)


def load_checkpoint(path: str, map_location="cpu") -> dict:
    """-> {'iter': n, model_name: {raw parameter name: tensor, ...training extras...}, ...} with the reference's key names mapped back
    (feature_dc -> features_dc, feature_rest -> features_rest)."""
    sd = torch.load(path, map_location=map_location, weights_only=False)
    out = {}
    for name, v in sd.items():
        if isinstance(v, Mapping) and "xyz" in v:
            out[name] = {_STATE_KEYS.get(k, k): t for k, t in v.items()}
        else:
            out[name] = v
    return out
