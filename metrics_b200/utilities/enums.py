"""String enums of the public API (reference: utilities/enums.py, built there on `lightning_utilities.StrEnum`).

Members compare case-insensitively with plain strings (and `AverageMethod.NONE` with ``None``), and `from_str` raises the
reference's ``Invalid <kind>: expected one of [...], but got <value>.`` message.
"""
from __future__ import annotations

from enum import Enum
from typing import List

from typing_extensions import Literal


# what `from_str` calls each enum in its error message (the reference overrides a `_name()` staticmethod per class)
_KIND_OF = {"DataType": "Data type", "AverageMethod": "Average method", "MDMCAverageMethod": "MDMC Average method",
            "ClassificationTask": "Classification", "ClassificationTaskNoBinary": "Classification",
            "ClassificationTaskNoMultilabel": "Classification"}


class EnumStr(str, Enum):
    """Case-insensitive string enum (reference :20-52)."""

    @classmethod
    def _name(cls) -> str:
        return _KIND_OF.get(cls.__name__, "Task")

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

    BINARY = "binary"
    MULTILABEL = "multi-label"
    MULTICLASS = "multi-class"
    MULTIDIM_MULTICLASS = "multi-dim multi-class"


class AverageMethod(EnumStr):
    """Averaging over classes; ``AverageMethod.NONE == None`` and ``== "none"`` (reference :72-93)."""

    MICRO = "micro"
    MACRO = "macro"
    WEIGHTED = "weighted"
    NONE = None
    SAMPLES = "samples"


class MDMCAverageMethod(EnumStr):
    """Averaging over the extra dimensions of multi-dim multi-class input (reference :96-105)."""

    GLOBAL = "global"
    SAMPLEWISE = "samplewise"


class ClassificationTask(EnumStr):
    """Tasks of the task-dispatching wrappers (reference :108-122)."""

    BINARY = "binary"
    MULTICLASS = "multiclass"
    MULTILABEL = "multilabel"


class ClassificationTaskNoBinary(EnumStr):
    """Reference :125-138."""

    MULTILABEL = "multilabel"
    MULTICLASS = "multiclass"


class ClassificationTaskNoMultilabel(EnumStr):
    """Reference :141-154."""

    BINARY = "binary"
    MULTICLASS = "multiclass"
