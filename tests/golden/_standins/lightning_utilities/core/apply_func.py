"""Stand-in for lightning_utilities.core.apply_func (see package docstring)."""
from collections import OrderedDict, defaultdict
from collections.abc import Mapping, Sequence


def _is_namedtuple(obj):
    return isinstance(obj, tuple) and hasattr(obj, "_asdict") and hasattr(obj, "_fields")


def apply_to_collection(data, dtype, function, *args, wrong_dtype=None, include_none=True, **kwargs):
    if isinstance(data, dtype) and (wrong_dtype is None or not isinstance(data, wrong_dtype)):
        return function(data, *args, **kwargs)
    elem_type = type(data)
    if isinstance(data, Mapping):
        out = []
        for k, v in data.items():
            v = apply_to_collection(v, dtype, function, *args, wrong_dtype=wrong_dtype,
                                    include_none=include_none, **kwargs)
            if include_none or v is not None:
                out.append((k, v))
        if isinstance(data, defaultdict):
            return elem_type(data.default_factory, OrderedDict(out))
        return elem_type(OrderedDict(out))
    is_nt = _is_namedtuple(data)
    if is_nt or (isinstance(data, Sequence) and not isinstance(data, str)):
        out = []
        for d in data:
            v = apply_to_collection(d, dtype, function, *args, wrong_dtype=wrong_dtype,
                                    include_none=include_none, **kwargs)
            if include_none or v is not None:
                out.append(v)
        return elem_type(*out) if is_nt else elem_type(out)
    return data
