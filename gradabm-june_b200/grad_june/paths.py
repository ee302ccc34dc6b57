"""Package locations (mirrors the names exported by the reference's grad_june/paths.py:4-8)."""
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent

#: what the ``@grad_june/`` alias in ``data_path`` resolves against (reference: repository root)
grad_june_path = _pkg_dir.parent

#: default YAML; written on first use from :mod:`grad_june.default_config`
default_config_path = _pkg_dir / "configs" / "default.yaml"


def ensure_default_config():
    """Materialise configs/default.yaml from the in-code parameter tables if it is missing."""
    if not default_config_path.exists():
        from .default_config import write_default_config

        write_default_config(default_config_path)
    return default_config_path
