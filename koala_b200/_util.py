"""Library / model path resolution (mirror of /root/reference/binding/python/_util.py:59-84 for this package)."""
import os

_PKG = os.path.dirname(os.path.abspath(__file__))


def default_library_path(relative: str = '') -> str:
    """Path of the in-tree CUDA engine.  It is never silently substituted: a missing file is an error at load time."""
    return os.path.join(_PKG, relative, 'lib', 'libpv_koala_b200.so')


def default_model_path(relative: str = '') -> str:
    return os.path.join(_PKG, relative, 'lib', 'koala_b200_params.kpv')


__all__ = ['default_library_path', 'default_model_path']
