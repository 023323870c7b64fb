"""Minimal duck-typed stand-ins for the blackboxopt / parameterspace types that appear in the
signature of `ScaMLGPBO` and `metadata_to_numerical` (reference scamlgp/optimizer.py:28-47,
scamlgp/utils.py:72-109).  Neither package is installable here; only what the hot
path's callers touch is provided: evaluations/objectives, a search space that maps configurations
to the unit cube and back (NaN for inactive conditional parameters), and the small helpers
`to_numerical`, `sort_evaluations`, `impute_nans_with_constant`, `filter_y_nans`.
"""
from __future__ import annotations

import json
import math
import operator
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


class OptimizerNotReady(Exception):
    """Raised when max_pending_evaluations specifications are outstanding."""


class OptimizationComplete(Exception):
    pass


class EvaluationsError(ValueError):
    """An evaluation does not match the optimizer's objective / search space."""


@dataclass
class Objective:
    name: str
    greater_is_better: bool


@dataclass
class EvaluationSpecification:
    configuration: Dict[str, Any]
    settings: Dict[str, Any] = field(default_factory=dict)
    optimizer_info: Dict[str, Any] = field(default_factory=dict)
    context: Optional[Dict[str, Any]] = None
    created_unixtime: float = field(default_factory=time.time)

    def create_evaluation(self, objectives: Dict[str, Optional[float]], **kwargs) -> "Evaluation":
        return Evaluation(objectives=objectives, configuration=dict(self.configuration), settings=dict(self.settings),
                          optimizer_info=dict(self.optimizer_info), context=self.context,
                          created_unixtime=self.created_unixtime, **kwargs)


@dataclass
class Evaluation:
    objectives: Dict[str, Optional[float]]
    configuration: Dict[str, Any]
    settings: Dict[str, Any] = field(default_factory=dict)
    optimizer_info: Dict[str, Any] = field(default_factory=dict)
    context: Optional[Dict[str, Any]] = None
    user_info: Optional[Dict[str, Any]] = None
    created_unixtime: float = field(default_factory=time.time)
    reported_unixtime: float = field(default_factory=time.time)

    def get_specification(self) -> EvaluationSpecification:
        return EvaluationSpecification(dict(self.configuration), dict(self.settings), dict(self.optimizer_info),
                                       self.context, self.created_unixtime)


# ---- parameters ------------------------------------------------------------------------------- #
class ContinuousParameter:
    is_continuous = True

    def __init__(self, name: str, bounds: Tuple[float, float]):
        self.name, self.bounds = name, (float(bounds[0]), float(bounds[1]))

    def to_num(self, v) -> float:
        lo, hi = self.bounds
        return (float(v) - lo) / (hi - lo)

    def to_num_array(self, v: np.ndarray) -> np.ndarray:
        """`to_num` on a float array (same operations in the same order: bit-identical to the scalar path)."""
        lo, hi = self.bounds
        return (v - lo) / (hi - lo)

    def from_num(self, u: float):
        lo, hi = self.bounds
        return lo + min(max(float(u), 0.0), 1.0) * (hi - lo)

    def sample(self, rng: np.random.Generator):
        return self.from_num(rng.random())


class IntegerParameter(ContinuousParameter):
    is_continuous = False

    def to_num(self, v) -> float:
        lo, hi = self.bounds
        return (float(v) - lo + 0.5) / (hi - lo + 1.0)

    def to_num_array(self, v: np.ndarray) -> np.ndarray:
        lo, hi = self.bounds
        return (v - lo + 0.5) / (hi - lo + 1.0)

    def from_num(self, u: float):
        lo, hi = self.bounds
        return int(min(max(math.floor(lo + float(u) * (hi - lo + 1.0)), lo), hi))


class CategoricalParameter:
    is_continuous = False

    def __init__(self, name: str, values: Sequence[Any]):
        self.name, self.values = name, list(values)

    def to_num(self, v) -> float:
        return (self.values.index(v) + 0.5) / len(self.values)

    def from_num(self, u: float):
        k = len(self.values)
        return self.values[int(min(max(math.floor(float(u) * k), 0), k - 1))]

    def sample(self, rng: np.random.Generator):
        return self.values[int(rng.integers(len(self.values)))]


OrdinalParameter = CategoricalParameter


class ParameterSpace:
    """Ordered collection of parameters, optionally conditional (active only if `condition(config)`)
    or fixed.  The numerical representation is the unit cube, NaN for inactive parameters."""

    def __init__(self):
        self._params: List[Any] = []
        self._conditions: Dict[str, Optional[Callable[[Dict[str, Any]], bool]]] = {}
        self._fixed: Dict[str, Any] = {}
        self._rng = np.random.default_rng()

    def add(self, parameter, condition: Optional[Callable[[Dict[str, Any]], bool]] = None) -> None:
        self._params.append(parameter)
        self._conditions[parameter.name] = condition

    def fix(self, **kwargs) -> None:
        self._fixed.update(kwargs)

    def seed(self, seed: Optional[int]) -> None:
        self._rng = np.random.default_rng(seed)

    def __len__(self) -> int:
        return len(self._params)

    @property
    def parameter_names(self) -> List[str]:
        return [p.name for p in self._params]

    @property
    def is_all_continuous(self) -> bool:
        return all(p.is_continuous for p in self._params) and not any(self._conditions.values())

    def _active(self, p, config) -> bool:
        cond = self._conditions[p.name]
        return True if cond is None else bool(cond(config))

    def sample(self) -> Dict[str, Any]:
        config: Dict[str, Any] = {}
        for p in self._params:
            if not self._active(p, config):
                continue
            config[p.name] = self._fixed[p.name] if p.name in self._fixed else p.sample(self._rng)
        return config

    def to_numerical(self, config: Dict[str, Any]) -> np.ndarray:
        out = np.full(len(self._params), np.nan)
        for i, p in enumerate(self._params):
            if p.name in config and config[p.name] is not None:
                out[i] = p.to_num(config[p.name])
        return out

    def to_numerical_batch(self, configs: Sequence[Dict[str, Any]]) -> np.ndarray:
        """[n, d] numerical representation of many configurations (column-wise: one pass per parameter)."""
        out = np.full((len(configs), len(self._params)), np.nan)
        if len(configs) and all(type(p) in (ContinuousParameter, IntegerParameter) for p in self._params):
            # all-numeric space, every parameter present in every configuration: ONE array conversion
            try:
                names = [p.name for p in self._params]
                get = operator.itemgetter(*names)
                rows = list(map(get, configs)) if len(names) > 1 else [(get(c),) for c in configs]
                raw = np.array(rows, dtype=float)
                for i, p in enumerate(self._params):
                    out[:, i] = p.to_num_array(raw[:, i])
                return out
            except (KeyError, TypeError, ValueError):
                pass  # conditional / inactive parameters: column by column below
        for i, p in enumerate(self._params):
            name, conv = p.name, p.to_num
            col = [c.get(name) for c in configs]
            if type(p) in (ContinuousParameter, IntegerParameter):
                # numeric parameters: one vectorised transform per column (the meta-data of 4096 tasks x 256 points
                # are 6 million values: a Python call per value was 3 of the 5 s of ScaMLGPBO's construction)
                try:
                    out[:, i] = p.to_num_array(np.array(col, dtype=float))
                    continue
                except (TypeError, ValueError):
                    pass  # inactive (None) or non-numeric entries: value by value below
            out[:, i] = [conv(v) if v is not None else np.nan for v in col]
        return out

    def from_numerical(self, vec) -> Dict[str, Any]:
        config: Dict[str, Any] = {}
        for p, u in zip(self._params, np.asarray(vec, dtype=float)):
            if not self._active(p, config):
                continue
            config[p.name] = self._fixed[p.name] if p.name in self._fixed else p.from_num(u)
        return config

    def numerical_bounds(self) -> np.ndarray:
        """[d, 2] box of the numerical representation (fixed parameters collapse to a point)."""
        b = np.tile(np.array([0.0, 1.0]), (len(self._params), 1))
        for i, p in enumerate(self._params):
            if p.name in self._fixed:
                b[i] = p.to_num(self._fixed[p.name])
        return b


# ---- helpers imported by the reference (utils.py:10-12,98-106; optimizer.py:10-13) -------------- #
def sort_evaluations(evaluations: Iterable[Evaluation]) -> List[Evaluation]:
    """Deterministic order independent of the reporting order."""
    def key(e: Evaluation):
        return json.dumps({"c": e.configuration, "o": e.objectives, "s": e.settings, "x": e.context}, sort_keys=True,
                          default=str)

    return sorted(evaluations, key=key)


def to_numerical(evaluations: Iterable[Evaluation], search_space: ParameterSpace, objectives: Sequence[Objective],
                 batch_shape: torch.Size = torch.Size(), torch_dtype: torch.dtype = torch.float64):
    """X [n, d] in the unit cube (NaN = inactive), Y [n, 1] as LOSSES (sign flipped for greater_is_better,
    NaN for missing objective values)."""
    evaluations = list(evaluations)
    obj = objectives[0]
    for e in evaluations:
        if obj.name not in e.objectives:
            raise EvaluationsError(f"Evaluation does not report the objective '{obj.name}'")
    X = search_space.to_numerical_batch([e.configuration for e in evaluations])
    sign = -1.0 if obj.greater_is_better else 1.0
    Y = np.array([np.nan if e.objectives[obj.name] is None else sign * float(e.objectives[obj.name])
                  for e in evaluations], dtype=float).reshape(-1, 1)
    return torch.tensor(X, dtype=torch_dtype), torch.tensor(Y, dtype=torch_dtype)


def sort_numerical(X: torch.Tensor, Y: torch.Tensor):
    """Rows of [X | Y] in lexicographic order (NaN last): the same deterministic, report-order independent
    arrangement `sort_evaluations` gives, obtained on the numerical representation (one numpy lexsort instead of
    one JSON key per evaluation -- the meta-data of 4096 tasks x 256 points are a million evaluations)."""
    if X.shape[0] <= 1:
        return X, Y
    A = np.concatenate([X.numpy().astype(float), Y.numpy().astype(float)], axis=1)
    if not np.isnan(A).any():  # the usual case (no inactive parameters, no missing objectives): the columns are the keys
        idx = torch.as_tensor(np.lexsort(A.T[::-1]).copy())
        return X[idx], Y[idx]
    keys = []
    for j in range(A.shape[1] - 1, -1, -1):  # lexsort: last key is the primary one
        col = A[:, j]
        keys.append(np.where(np.isnan(col), np.inf, col))
        keys.append(np.isnan(col))
    # primary: column 0 value (NaN flag first so that NaN rows sort last within a column)
    order = np.lexsort(tuple(keys))
    idx = torch.as_tensor(order.copy())
    return X[idx], Y[idx]


def impute_nans_with_constant(X: torch.Tensor, c: float = -1.0) -> torch.Tensor:
    return torch.where(torch.isnan(X), torch.full_like(X, c), X)


def filter_y_nans(X: torch.Tensor, Y: torch.Tensor):
    keep = ~torch.isnan(Y.reshape(Y.shape[0], -1)).any(dim=1)
    return X[keep], Y[keep]
