"""YAML configuration loader with the reference's interface (src/config_manager.py there):

    load_conf(startswith='parameters') -> dict

merges every ``conf/<startswith>*.yml|yaml`` file of the project into one dict (later files
override earlier top-level keys, as the reference's ``dict.update`` merge does, :28-38).
The project root is found by walking up from the working directory to the first directory
that contains ``.git`` (the reference checks cwd, then its parent, then the path prefix before
``src``, :15-26 -- all three are covered by the upward walk), or given explicitly.
"""
from __future__ import annotations

from pathlib import Path

import yaml

_SUFFIXES = (".yml", ".yaml")


def get_project_dir(start: str | Path | None = None) -> str:
    here = Path(start) if start is not None else Path.cwd()
    for cand in (here, *here.resolve().parents):
        if (cand / ".git").exists():
            return str(cand)
    raise AssertionError(f"no project directory (a parent holding .git) above {here}")


def load_all_ymls(config_folder, startswith: str = "parameters") -> dict:
    merged: dict = {}
    for path in sorted(Path(config_folder).iterdir()):
        if path.name.startswith(startswith) and path.suffix in _SUFFIXES:
            data = yaml.safe_load(path.read_text())
            if data:
                merged.update(data)
    return merged


def load_conf(startswith: str = "parameters", project_dir: str | Path | None = None) -> dict:
    root = Path(project_dir) if project_dir is not None else Path(get_project_dir())
    return load_all_ymls(root / "conf", startswith)
