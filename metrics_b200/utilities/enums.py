"""Task names accepted by the task-dispatching wrappers (reference: utilities/enums.py)."""
from enum import Enum


class _CaseInsensitiveStrEnum(str, Enum):
    @classmethod
    def from_str(cls, value: str) -> "_CaseInsensitiveStrEnum":
        key = str(value).lower().replace("-", "_")
        for member in cls:
            if member.value == key or member.name.lower() == key:
                return member
        raise ValueError(f"Invalid {cls._name()}: expected one of {[m.value for m in cls]}, but got {value}.")

    @staticmethod
    def _name() -> str:
        return "Task"

    def __eq__(self, other: object) -> bool:
        if isinstance(other, Enum):
            other = other.value
        return self.value == str(other).lower()

    def __hash__(self) -> int:
        return hash(self.value)


class ClassificationTask(_CaseInsensitiveStrEnum):
    BINARY = "binary"
    MULTICLASS = "multiclass"
    MULTILABEL = "multilabel"


class ClassificationTaskNoMultilabel(_CaseInsensitiveStrEnum):
    BINARY = "binary"
    MULTICLASS = "multiclass"


class ClassificationTaskNoBinary(_CaseInsensitiveStrEnum):
    MULTILABEL = "multilabel"
    MULTICLASS = "multiclass"
