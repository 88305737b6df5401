"""smart_crossover -- B200-native network-crossover hot path behind the reference's API.

Drop-in for the `smart_crossover.network_methods` call surface and the `formats`
problem classes of wcwj0147/smart-crossover.  The hot path (flow scoring + argsort,
spanning-tree basis identification, tree potentials, pricing) runs as hand-written
sm_100a CUDA in libsxcross.so (C ABI: include/sxcross.h); see DESIGN.md.
"""
from pathlib import Path

__all__ = ["get_project_root", "get_data_dir_path"]


def get_project_root() -> Path:
    """Repository root (the reference walks up to a directory literally named
    `smart-crossover`, `__init__.py:4-12`; here it is the parent of the package tree)."""
    return Path(__file__).resolve().parents[2]


def get_data_dir_path() -> Path:
    return get_project_root() / "data"
