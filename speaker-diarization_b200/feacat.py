"""feacat ``-H --raw-output`` feature files.

Little-endian ``int32 dim`` followed by ``float32[N][dim]`` frame-major; there
is no frame count in the header (spk-change-detection.py:37-41).
"""

import os.path as op

import numpy as np


def feature_file_name(audio, feapath, ext, concat=False):
    """Name of the feature file of a recipe line's ``audio=`` entry.

    ``concat=False`` joins with ``os.path.join`` (spk-change-detection.py:33-35,
    spk-clustering2.py:34-37); ``concat=True`` is the plain string
    concatenation of spk-clustering.py:33-35 (its caller appends the ``/``)."""
    base = op.splitext(op.basename(audio))[0] + ext
    return feapath + base if concat else op.join(feapath, base)


def read_features(path):
    """-> (dim, float32 array of shape (N, dim)); a ragged tail that does not
    fill a frame is an error, as the reference's ``reshape`` makes it."""
    with open(path, 'rb') as f:
        head = np.fromfile(f, dtype='<i4', count=1)
        if head.size != 1:
            raise ValueError('feature file %s has no dimension header' % path)
        dim = int(head[0])
        data = np.fromfile(f, dtype='<f4')
    if dim <= 0:
        raise ValueError('feature file %s: bad dimension %d' % (path, dim))
    return dim, data.reshape(data.size // dim, dim)


def write_features(path, frames):
    frames = np.ascontiguousarray(frames, dtype='<f4')
    with open(path, 'wb') as f:
        np.array([frames.shape[1]], dtype='<i4').tofile(f)
        frames.tofile(f)
