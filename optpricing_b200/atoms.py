"""
Stand-alone mirrors of the reference's input containers
(/root/reference/src/optpricing/atoms/{option,stock,rate}.py): same field names,
defaults and validation, so the techniques run without the reference installed
(the GPU box has no /root/reference).  The techniques read these by attribute
only, so the reference's own atoms work equally well (drop-in mode).
"""
from __future__ import annotations

import dataclasses
import enum
from typing import Callable


class OptionType(enum.Enum):
    CALL = "Call"
    PUT = "Put"


class ExerciseStyle(enum.Enum):
    EUROPEAN = "European"
    AMERICAN = "American"
    BERMUDAN = "Bermudan"


@dataclasses.dataclass(frozen=True, slots=True)
class Option:
    strike: float
    maturity: float
    option_type: OptionType
    exercise_style: ExerciseStyle = ExerciseStyle.EUROPEAN

    def __post_init__(self):
        for label, value in (("Strike", self.strike), ("Maturity", self.maturity)):
            if value <= 0:
                raise ValueError(f"{label} must be positive, got {value}")


@dataclasses.dataclass(frozen=True, slots=True)
class Stock:
    spot: float
    volatility: float | None = None
    dividend: float = 0.0
    discrete_dividends: list[float] | None = None
    ex_div_times: list[float] | None = None

    def __post_init__(self):
        if self.spot <= 0:
            raise ValueError(f"Spot price must be positive, got {self.spot}")


@dataclasses.dataclass(frozen=True, slots=True)
class Rate:
    rate: float | Callable[[float], float]

    def get_rate(self, t: float = 0) -> float:
        return self.rate(t) if callable(self.rate) else self.rate


def is_call(option) -> bool:
    """True for a call; works for this module's enum and the reference's."""
    return getattr(option.option_type, "value", option.option_type) == "Call"
