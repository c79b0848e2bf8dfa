"""GridPatch / GridData containers -- mirrors src/GridTypes.jl:4-16 of the reference."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class GridPatch:
    LE: list          # LeftEdge
    RE: list          # RightEdge
    id: int           # identifier
    level: int        # For AMR someday ...
    dim: tuple        # dimensions of grid patch


@dataclass
class GridData:
    d: dict = field(default_factory=dict)
