"""String enums of the public API (reference: utilities/enums.py, built there on `lightning_utilities.StrEnum`).

Members compare case-insensitively with plain strings (and `AverageMethod.NONE` with ``None``), and `from_str` raises the
reference's ``Invalid <kind>: expected one of [...], but got <value>.`` message.
"""
from __future__ import annotations

from enum import Enum
from typing import List

from typing_extensions import Literal


class EnumStr(str, Enum):
    """Case-insensitive string enum (reference :20-52)."""

    @staticmethod
    def _name() -> str:
        return "Task"

    @classmethod
    def _allowed_matches(cls, source: str) -> List[str]:
        keys = list(cls._member_names_)
        values = [member.value for member in cls]
        return keys if source == "key" else values if source == "value" else keys + values

    @classmethod
    def from_str(cls, value: str, source: Literal["key", "value", "any"] = "key") -> "EnumStr":
        """The member whose name (``source="key"``), value (``"value"``) or either (``"any"``) equals ``value``, ignoring case
        and treating ``-`` as ``_``."""
        wanted = str(value).replace("-", "_").lower()
        for member in cls:
            by_key = source in ("key", "any") and member.name.lower() == wanted
            by_value = source in ("value", "any") and str(member.value).lower() == wanted
            if by_key or by_value:
                return member
        raise ValueError(f"Invalid {cls._name()}: expected one of {cls._allowed_matches(source)}, but got {value}.")

    def __eq__(self, other: object) -> bool:
        if isinstance(other, Enum):
            other = other.value
        return str(self.value).lower() == str(other).lower()

    def __hash__(self) -> int:
        return hash(str(self.value).lower())


class DataType(EnumStr):
    """Kinds of classification input of the legacy API (reference :55-69)."""

    @staticmethod
    def _name() -> str:
        return "Data type"

    BINARY = "binary"
    MULTILABEL = "multi-label"
    MULTICLASS = "multi-class"
    MULTIDIM_MULTICLASS = "multi-dim multi-class"


class AverageMethod(EnumStr):
    """Averaging over classes; ``AverageMethod.NONE == None`` and ``== "none"`` (reference :72-93)."""

    @staticmethod
    def _name() -> str:
        return "Average method"

    MICRO = "micro"
    MACRO = "macro"
    WEIGHTED = "weighted"
    NONE = None
    SAMPLES = "samples"


class MDMCAverageMethod(EnumStr):
    """Averaging over the extra dimensions of multi-dim multi-class input (reference :96-105)."""

    @staticmethod
    def _name() -> str:
        return "MDMC Average method"

    GLOBAL = "global"
    SAMPLEWISE = "samplewise"


class ClassificationTask(EnumStr):
    """Tasks of the task-dispatching wrappers (reference :108-122)."""

    @staticmethod
    def _name() -> str:
        return "Classification"

    BINARY = "binary"
    MULTICLASS = "multiclass"
    MULTILABEL = "multilabel"


class ClassificationTaskNoBinary(EnumStr):
    """Reference :125-138."""

    @staticmethod
    def _name() -> str:
        return "Classification"

    MULTILABEL = "multilabel"
    MULTICLASS = "multiclass"


class ClassificationTaskNoMultilabel(EnumStr):
    """Reference :141-154."""

    @staticmethod
    def _name() -> str:
        return "Classification"

    BINARY = "binary"
    MULTICLASS = "multiclass"
