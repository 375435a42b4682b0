"""Stand-in for lightning_utilities.core.imports (see package docstring)."""
import importlib.util
import re
from importlib import metadata


def package_available(name):
    try:
        return importlib.util.find_spec(name) is not None
    except Exception:
        return False


def _vtuple(v):
    v = v.split("+")[0]
    parts = []
    for p in re.split(r"[.\-]", v):
        m = re.match(r"\d+", p)
        if not m:
            break
        parts.append(int(m.group()))
    return tuple(parts)


class RequirementCache:
    def __init__(self, requirement, module=None):
        self.requirement = requirement
        self.module = module

    def _check(self):
        if hasattr(self, "available"):
            return
        m = re.match(r"^\s*([A-Za-z0-9_.\-]+)\s*(.*)$", self.requirement)
        name, spec = m.group(1), m.group(2)
        try:
            ver = _vtuple(metadata.version(name))
        except Exception:
            mod = self.module or name.replace("-", "_")
            self.available = package_available(mod) and not spec
            self.message = f"Requirement {self.requirement!r} " + ("met" if self.available else "not met")
            return
        ok = True
        for clause in [c.strip() for c in spec.split(",") if c.strip()]:
            mm = re.match(r"(>=|<=|==|!=|>|<)\s*(.+)", clause)
            if not mm:
                continue
            op, tgt = mm.group(1), _vtuple(mm.group(2))
            n = max(len(ver), len(tgt))
            a = ver + (0,) * (n - len(ver))
            b = tgt + (0,) * (n - len(tgt))
            ok &= {">=": a >= b, "<=": a <= b, "==": a == b, "!=": a != b, ">": a > b, "<": a < b}[op]
        self.available = ok
        self.message = f"Requirement {self.requirement!r} " + ("met" if ok else "not met")

    def __bool__(self):
        self._check()
        return self.available

    def __str__(self):
        self._check()
        return self.message

    __repr__ = __str__
