"""Mesh input for the surface examples: same entry point as the reference's ``RVGP/utils.py`` (``load_mesh``,
utils.py:7-47) -- ``load_mesh(name, folder)`` returns ``(vertices float64 (nv, 3), faces int (nf, 3), 0-based)`` of
``<folder>/<name>.obj``; the arrays feed ``RVGP.create_data_object``.

Written for this repo (not derived from the reference's line-by-line loop): the file is tokenised in one pass and the
numeric columns are converted array-at-a-time.  Behaviour kept: records are selected by their first token (``v`` = vertex,
``f`` = face), ``#`` comment lines and every other record type (``vn``, ``vt``, ``g``, ``s`` ...) are ignored, only the
first three numbers of a record are used, face indices are shifted from 1-based to 0-based.  Accepted superset: blank
lines (the reference raises IndexError) and ``f a/b/c`` vertex/texture/normal triplets (the reference raises ValueError).
"""
import os

import numpy as np

_DATA_ENV = "RVGP_DATA_DIR"


def _default_data_dir():
    """Where ``load_mesh(name)`` looks when no folder is given: $RVGP_DATA_DIR, else ``examples/data`` next to the
    package (the reference ships its meshes there, utils.py:13-14; this repo does not ship meshes)."""
    env = os.environ.get(_DATA_ENV)
    if env:
        return env
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "data")


def _records(path):
    """Token lists of the ``v`` and ``f`` records of an OBJ file, in file order."""
    verts, faces = [], []
    with open(path, "r") as fh:
        for raw in fh:
            tok = raw.split()
            if not tok or raw[0] == "#":
                continue
            if tok[0] == "v":
                verts.append(tok[1:4])
            elif tok[0] == "f":
                faces.append(tok[1:4])
    return verts, faces


def load_mesh(data='bunny', folder=None):
    path = os.path.join(_default_data_dir() if folder is None else folder, data) + ".obj"
    if not os.path.exists(path):
        raise FileNotFoundError("mesh %r not found (%s); pass folder=... or set $%s" % (data, path, _DATA_ENV))
    verts, faces = _records(path)
    vertices = np.asarray(verts, dtype=np.float64) if verts else np.array([])
    if faces:
        # 'a/b/c' -> 'a': only the vertex index of a face corner is used
        faces = np.asarray([[c.split("/", 1)[0] for c in f] for f in faces]).astype(np.int64) - 1
    else:
        faces = np.array([]) - 1
    return vertices, faces
