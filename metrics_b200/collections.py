"""``MetricCollection``: one ``update``/``compute``/``reset`` for many metrics, with shared state ("compute groups").

Keeps the reference API (src/torchmetrics/collections.py:59-616): constructor ``(metrics, *more, prefix, postfix,
compute_groups)``, dict-style access with prefix/postfix renaming, kwargs filtering per member, automatic detection
of members whose states are identical after the first ``update`` so that only one member per group runs ``update``
afterwards (collections.py:231-343).  Group detection here compares *every* state of two members (the reference
returns after the first state it looks at); groups are therefore never coarser than the reference's.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from collections.abc import Hashable, Iterable, Iterator, Mapping, Sequence
from copy import deepcopy
from typing import Any, Dict, List, Optional, Union

import torch
from torch import Tensor
from torch.nn import ModuleDict

from metrics_b200.metric import Metric
from metrics_b200.utilities.data import allclose
from metrics_b200.utilities.prints import rank_zero_warn


def _strip_prefix(text: str, prefix: str) -> str:
    return text[len(prefix):] if text.startswith(prefix) else text


def _strip_suffix(text: str, suffix: str) -> str:
    return text[: -len(suffix)] if suffix and text.endswith(suffix) else text


def _has_duplicate_inner_keys(results: Dict[str, Any]) -> bool:
    """Do two dict-valued results share a key (or clash with a scalar result's name)?"""
    seen = set()
    for name, value in results.items():
        keys = list(value.keys()) if isinstance(value, dict) else [name]
        for k in keys:
            if k in seen:
                return True
            seen.add(k)
    return False


def _states_match(a: Metric, b: Metric) -> bool:
    if not a._defaults or not b._defaults or a._defaults.keys() != b._defaults.keys():
        return False
    for name in a._defaults:
        sa, sb = getattr(a, name), getattr(b, name)
        if type(sa) != type(sb):  # noqa: E721
            return False
        if isinstance(sa, Tensor):
            if sa.shape != sb.shape or not allclose(sa, sb):
                return False
        elif isinstance(sa, list):
            if len(sa) != len(sb):
                return False
            for xa, xb in zip(sa, sb):
                if xa.shape != xb.shape or not allclose(xa, xb):
                    return False
    return True


class MetricCollection(ModuleDict):
    """A dict of metrics driven with a single call; see module docstring."""

    _modules: Dict[str, Metric]  # type: ignore[assignment]
    __jit_unused_properties__ = ["metric_state"]

    def __init__(
        self,
        metrics: Union[Metric, Sequence[Metric], Dict[str, Metric]],
        *additional_metrics: Metric,
        prefix: Optional[str] = None,
        postfix: Optional[str] = None,
        compute_groups: Union[bool, List[List[str]]] = True,
    ) -> None:
        super().__init__()
        self.prefix = self._check_arg(prefix, "prefix")
        self.postfix = self._check_arg(postfix, "postfix")
        self._enable_compute_groups = compute_groups
        self._groups_checked = False
        self._state_is_copy = False
        self.add_metrics(metrics, *additional_metrics)

    # ------------------------------------------------------------------------------------------------
    # driving the members
    # ------------------------------------------------------------------------------------------------
    @property
    def metric_state(self) -> Dict[str, Dict[str, Any]]:
        return {k: m.metric_state for k, m in self.items(keep_base=False, copy_state=False)}

    @torch.jit.unused
    def forward(self, *args: Any, **kwargs: Any) -> Dict[str, Any]:
        return self._compute_and_reduce("forward", *args, **kwargs)

    def update(self, *args: Any, **kwargs: Any) -> None:
        """Positional args go to every member, kwargs are filtered by each member's ``update`` signature."""
        if self._groups_checked:
            for name in self.keys(keep_base=True):
                self._modules[str(name)]._computed = None  # invalidate every member's cached result
            fused = self._fused_update(args, kwargs)
            for members in self._groups.values():
                if members[0] in fused:
                    continue
                leader = self._modules[members[0]]
                leader.update(*args, **leader._filter_kwargs(**kwargs))
            if self._state_is_copy:
                # someone read a member (copying its state) since the last update: re-link the group
                self._compute_groups_create_state_ref()
                self._state_is_copy = False
            return
        for member in self.values(copy_state=False):
            member.update(*args, **member._filter_kwargs(**kwargs))
        if self._enable_compute_groups:
            self._merge_compute_groups()
            self._compute_groups_create_state_ref()
            self._groups_checked = True

    def _fused_update(self, args: tuple, kwargs: dict) -> tuple:
        """Collection-level fusion (csrc/fused.cu, K11): when one group leader is a multiclass stat-scores metric and another
        an exact-mode multiclass curve metric over the same classes, ONE kernel reads the shared batch once and feeds both —
        the argmax counts of the first and the `normalize_logits_if_needed` probabilities the second keeps as its list state
        (the reference hands the batch to every member: collections.py:231-262).  Returns the names of the leaders served."""
        if len(args) != 2 or kwargs or os.environ.get("MB200_COLLECTION_FUSION", "1") == "0":
            return ()
        preds, target = args
        if not (isinstance(preds, Tensor) and isinstance(target, Tensor) and preds.is_cuda and preds.ndim == 2
                and target.ndim == 1 and preds.dtype in (torch.float32, torch.float16, torch.bfloat16)
                and not target.is_floating_point() and target.shape[0] == preds.shape[0] and preds.shape[0] > 0):
            return ()
        plan = self._fusion_plan()
        if plan is None or preds.shape[1] != plan[2]:
            return ()
        from metrics_b200 import _native

        stats, curve = self._modules[plan[0]], self._modules[plan[1]]
        for member in (stats, curve):  # what Metric._wrap_update does around a member's own update()
            member._computed = None
            member._update_count += 1
        curve._group_cache.clear()
        probs = _native.multiclass_stats_softmax_update_(
            stats.tp, stats.fp, stats.tn, stats.fn, stats._workspace(plan[2], stats.tp.device), preds, target, plan[2],
            stats.average == "micro")
        curve.preds.append(probs)
        curve.target.append(target)
        if curve.compute_on_cpu:
            curve._move_list_states_to_cpu()
        return plan[:2]

    def _fusion_plan(self) -> Optional[tuple]:
        """(stat-scores leader, curve leader, num_classes) or None; decided once per grouping."""
        cached = self.__dict__.get("_fusion_plan_cache")
        if cached is not None and cached[0] == id(self._groups):
            return cached[1]
        from metrics_b200.classification.precision_recall_curve import MulticlassPrecisionRecallCurve
        from metrics_b200.classification.stat_scores import MulticlassStatScores

        stats = curve = None
        for members in self._groups.values():
            leader = self._modules[members[0]]
            if getattr(leader, "validate_args", True) or getattr(leader, "ignore_index", 0) is not None:
                continue
            if (stats is None and isinstance(leader, MulticlassStatScores) and leader.top_k == 1 and leader.num_classes
                    and leader.multidim_average == "global" and leader.num_classes <= 1024):
                stats = members[0]
            elif (curve is None and isinstance(leader, MulticlassPrecisionRecallCurve) and leader.thresholds is None
                  and leader.average != "micro"):
                curve = members[0]
        plan = None
        if stats is not None and curve is not None:
            c = self._modules[stats].num_classes
            if self._modules[curve].num_classes == c:
                plan = (stats, curve, c)
        self.__dict__["_fusion_plan_cache"] = (id(self._groups), plan)
        return plan

    def _merge_compute_groups(self) -> None:
        """Partition members by equal states after the first update (first member of a group is its leader)."""
        merged: List[List[str]] = []
        for members in self._groups.values():
            probe = self._modules[members[0]]
            for group in merged:
                if self._equal_metric_states(self._modules[group[0]], probe):
                    group.extend(members)
                    break
            else:
                merged.append(list(members))
        self._groups = dict(enumerate(merged))

    @staticmethod
    def _equal_metric_states(metric1: Metric, metric2: Metric) -> bool:
        return _states_match(metric1, metric2)

    def _compute_groups_create_state_ref(self, copy: bool = False) -> None:
        """Point every follower's states at its leader's (or deep-copy them when a member is handed out)."""
        if not self._state_is_copy:
            for members in self._groups.values():
                leader = self._modules[members[0]]
                for name in members[1:]:
                    follower = self._modules[name]
                    for state in leader._defaults:
                        value = getattr(leader, state)
                        setattr(follower, state, deepcopy(value) if copy else value)
                    follower._update_count = leader._update_count
                    if hasattr(leader, "_group_cache"):  # memoised evaluations shared by the group (curve metrics)
                        follower._group_cache = {} if copy else leader._group_cache
        self._state_is_copy = copy

    def compute(self) -> Dict[str, Any]:
        return self._compute_and_reduce("compute")

    def _compute_and_reduce(self, method_name: str, *args: Any, **kwargs: Any) -> Dict[str, Any]:
        if method_name not in ("compute", "forward"):
            raise ValueError(f"method_name should be either 'compute' or 'forward', but got {method_name}")
        raw: Dict[str, Any] = {}
        for name, member in self.items(keep_base=True, copy_state=False):
            raw[name] = member.compute() if method_name == "compute" else member(*args, **member._filter_kwargs(**kwargs))

        clash = _has_duplicate_inner_keys(raw)
        flat: Dict[str, Any] = {}
        for name, member in self.items(keep_base=True, copy_state=False):
            value = raw[name]
            if not isinstance(value, dict):
                flat[name] = value
                continue
            from_nested = bool(getattr(member, "_from_collection", None))
            for key, item in value.items():
                if clash:
                    base = name.replace(getattr(member, "prefix", "") or "", "").replace(
                        getattr(member, "postfix", "") or "", ""
                    )
                    key = f"{base}_{key}"
                if from_nested and member.prefix is not None:
                    key = f"{member.prefix}{key}"
                if from_nested and member.postfix is not None:
                    key = f"{key}{member.postfix}"
                flat[key] = item
        return {self._set_name(k): v for k, v in flat.items()}

    def reset(self) -> None:
        for member in self.values(copy_state=False):
            member.reset()
        if self._enable_compute_groups and self._groups_checked:
            self._compute_groups_create_state_ref()

    def clone(self, prefix: Optional[str] = None, postfix: Optional[str] = None) -> "MetricCollection":
        other = deepcopy(self)
        if prefix:
            other.prefix = self._check_arg(prefix, "prefix")
        if postfix:
            other.postfix = self._check_arg(postfix, "postfix")
        return other

    def persistent(self, mode: bool = True) -> None:
        for member in self.values(copy_state=False):
            member.persistent(mode)

    def set_dtype(self, dst_type: Union[str, torch.dtype]) -> "MetricCollection":
        for member in self.values(copy_state=False):
            member.set_dtype(dst_type)
        return self

    # ------------------------------------------------------------------------------------------------
    # membership
    # ------------------------------------------------------------------------------------------------
    def _adopt_nested(self, nested: "MetricCollection", name_prefix: str = "") -> None:
        for key, member in nested.items(keep_base=False):
            member.postfix = nested.postfix
            member.prefix = nested.prefix
            member._from_collection = True
            self[f"{name_prefix}{key}"] = member

    def add_metrics(self, metrics: Union[Metric, Sequence[Metric], Dict[str, Metric]], *additional_metrics: Metric) -> None:
        if isinstance(metrics, Metric):
            metrics = [metrics]
        if isinstance(metrics, Sequence):
            metrics = list(metrics)
            ignored = [m for m in additional_metrics if not isinstance(m, Metric)]
            metrics.extend(m for m in additional_metrics if isinstance(m, Metric))
            if ignored:
                rank_zero_warn(
                    f"You have passes extra arguments {ignored} which are not `Metric` so they will be ignored."
                )
        elif additional_metrics:
            raise ValueError(
                f"You have passes extra arguments {additional_metrics} which are not compatible"
                f" with first passed dictionary {metrics} so they will be ignored."
            )

        if isinstance(metrics, dict):
            for name in sorted(metrics.keys()):  # deterministic member order
                member = metrics[name]
                if isinstance(member, Metric):
                    self[name] = member
                elif isinstance(member, MetricCollection):
                    self._adopt_nested(member, f"{name}_")
                else:
                    raise ValueError(
                        f"Value {member} belonging to key {name} is not an instance of"
                        " `torchmetrics.Metric` or `torchmetrics.MetricCollection`"
                    )
        elif isinstance(metrics, Sequence):
            for member in metrics:
                if isinstance(member, Metric):
                    name = member.__class__.__name__
                    if name in self:
                        raise ValueError(f"Encountered two metrics both named {name}")
                    self[name] = member
                elif isinstance(member, MetricCollection):
                    self._adopt_nested(member)
                else:
                    raise ValueError(
                        f"Input {member} to `MetricCollection` is not a instance of"
                        " `torchmetrics.Metric` or `torchmetrics.MetricCollection`"
                    )
        else:
            raise ValueError(
                "Unknown input to MetricCollection. Expected, `Metric`, `MetricCollection` or `dict`/`sequence` of the"
                f" previous, but got {metrics}"
            )

        self._groups_checked = False
        if self._enable_compute_groups:
            self._init_compute_groups()
        else:
            self._groups = {}

    def _init_compute_groups(self) -> None:
        if isinstance(self._enable_compute_groups, list):
            self._groups = dict(enumerate(self._enable_compute_groups))
            for members in self._groups.values():
                for name in members:
                    if name not in self:
                        raise ValueError(
                            f"Input {name} in `compute_groups` argument does not match a metric in the collection."
                            f" Please make sure that {self._enable_compute_groups} matches {self.keys(keep_base=True)}"
                        )
            self._groups_checked = True  # user-specified groups are trusted as-is
        else:
            self._groups = {i: [str(k)] for i, k in enumerate(self.keys(keep_base=True))}

    @property
    def compute_groups(self) -> Dict[int, List[str]]:
        return self._groups

    # ------------------------------------------------------------------------------------------------
    # dict protocol with prefix / postfix renaming
    # ------------------------------------------------------------------------------------------------
    def _set_name(self, base: str) -> str:
        name = base if self.prefix is None else self.prefix + base
        return name if self.postfix is None else name + self.postfix

    def _to_renamed_dict(self) -> Mapping[str, Metric]:
        renamed: Dict[str, Metric] = OrderedDict() if isinstance(self._modules, OrderedDict) else {}
        for key, member in self._modules.items():
            renamed[self._set_name(key)] = member
        return renamed

    def __iter__(self) -> Iterator[Hashable]:
        return iter(self.keys())

    def keys(self, keep_base: bool = False) -> Iterable[Hashable]:
        return self._modules.keys() if keep_base else self._to_renamed_dict().keys()

    def items(self, keep_base: bool = False, copy_state: bool = True) -> Iterable[tuple]:
        self._compute_groups_create_state_ref(copy_state)
        return self._modules.items() if keep_base else self._to_renamed_dict().items()

    def values(self, copy_state: bool = True) -> Iterable[Metric]:
        self._compute_groups_create_state_ref(copy_state)
        return self._modules.values()

    def __getitem__(self, key: str, copy_state: bool = True) -> Metric:
        self._compute_groups_create_state_ref(copy_state)
        if self.prefix:
            key = _strip_prefix(key, self.prefix)
        if self.postfix:
            key = _strip_suffix(key, self.postfix)
        return self._modules[key]

    @staticmethod
    def _check_arg(arg: Optional[str], name: str) -> Optional[str]:
        if arg is None or isinstance(arg, str):
            return arg
        raise ValueError(f"Expected input `{name}` to be a string, but got {type(arg)}")

    def __repr__(self) -> str:
        text = super().__repr__()[:-2]
        if self.prefix:
            text += f",\n  prefix={self.prefix}{',' if self.postfix else ''}"
        if self.postfix:
            text += f"{',' if not self.prefix else ''}\n  postfix={self.postfix}"
        return text + "\n)"
