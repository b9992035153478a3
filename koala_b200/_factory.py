"""Factory functions with the reference's names and defaults (/root/reference/binding/python/_factory.py:27-76)."""
from typing import Optional, Sequence

from ._koala import Koala, list_hardware_devices
from ._util import default_library_path, default_model_path


def create(access_key: str, model_path: Optional[str] = None, device: Optional[str] = None,
           library_path: Optional[str] = None) -> Koala:
    """Factory for the single-stream engine; `device` defaults to `best` (first B200)."""
    return Koala(
        access_key=access_key,
        model_path=default_model_path() if model_path is None else model_path,
        device="best" if device is None else device,
        library_path=default_library_path() if library_path is None else library_path)


def available_devices(library_path: Optional[str] = None) -> Sequence[str]:
    """Every entry can be passed as `device` to `create`."""
    return list_hardware_devices(library_path=default_library_path() if library_path is None else library_path)


__all__ = ['available_devices', 'create']
