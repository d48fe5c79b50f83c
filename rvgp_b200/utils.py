"""Mirror of the reference's ``RVGP/utils.py``: ``load_mesh`` (utils.py:7-47), the Wavefront-OBJ reader used by the
surface-interpolation examples (vertices ``v x y z`` and triangular faces ``f a b c``, 1-based -> 0-based).  Host-side
file IO, exactly as in the reference; the arrays feed ``RVGP.create_data_object``."""
import os

import numpy as np


def load_mesh(data='bunny', folder=None):
    vertices = []
    faces = []
    dirname = os.path.dirname(os.path.realpath(__file__))
    if folder is None:
        file = os.path.join(dirname, '..', 'examples/data', data)
    else:
        file = os.path.join(folder, data)
    with open('{}.obj'.format(file), 'r') as fh:
        for line in fh:
            if line.startswith('#'):
                continue
            words = line.split()
            if not words:                       # the reference raises IndexError on blank lines; skipping is a superset
                continue
            if words[0] == 'v':
                vertices.append([float(words[1]), float(words[2]), float(words[3])])
            elif words[0] == 'f':
                faces.append([int(words[1]), int(words[2]), int(words[3])])
    vertices = np.array(vertices)
    faces = np.array(faces) - 1
    return vertices, faces
