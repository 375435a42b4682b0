"""Stand-in for lightning_utilities.core.enums (see package docstring)."""
from enum import Enum


class StrEnum(str, Enum):
    @classmethod
    def from_str(cls, value, source="key"):
        if source in ("key", "any"):
            for name in cls.__members__:
                if name.lower() == value.lower():
                    return cls[name]
        if source in ("value", "any"):
            for name, member in cls.__members__.items():
                if str(member.value).lower() == value.lower():
                    return cls[name]
        raise ValueError(f"Invalid match: expected one of {cls._allowed_matches(source)}, but got {value}.")

    @classmethod
    def try_from_str(cls, value, source="key"):
        try:
            return cls.from_str(value, source)
        except ValueError:
            return None

    @classmethod
    def _allowed_matches(cls, source):
        keys, vals = list(cls.__members__), [m.value for m in cls]
        return keys if source == "key" else vals if source == "value" else keys + vals

    def __eq__(self, other):
        if isinstance(other, Enum):
            other = other.value
        return self.value.lower() == str(other).lower()

    def __hash__(self):
        return hash(self.value.lower())
