"""Host-side helpers of the metric runtime (reference: src/torchmetrics/utilities/)."""
from metrics_b200.utilities.checks import check_forward_full_state_property  # noqa: F401
from metrics_b200.utilities.data import apply_to_collection, dim_zero_cat, dim_zero_max, dim_zero_mean, dim_zero_min, dim_zero_sum  # noqa: F401
from metrics_b200.utilities.distributed import class_reduce, gather_all_tensors, reduce  # noqa: F401
from metrics_b200.utilities.exceptions import TorchMetricsUserError, TorchMetricsUserWarning  # noqa: F401
from metrics_b200.utilities.prints import rank_zero_debug, rank_zero_info, rank_zero_warn  # noqa: F401
