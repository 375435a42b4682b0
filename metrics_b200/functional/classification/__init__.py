"""Functional classification metrics (reference: src/torchmetrics/functional/classification/)."""
from metrics_b200.functional.classification.accuracy import multiclass_accuracy  # noqa: F401
from metrics_b200.functional.classification.confusion_matrix import multiclass_confusion_matrix  # noqa: F401
from metrics_b200.functional.classification.f_beta import multiclass_f1_score, multiclass_fbeta_score  # noqa: F401
from metrics_b200.functional.classification.stat_scores import multiclass_stat_scores  # noqa: F401
from metrics_b200.functional.classification.auroc import binary_auroc, multiclass_auroc  # noqa: F401,E402
from metrics_b200.functional.classification.average_precision import (  # noqa: F401,E402
    binary_average_precision,
    multiclass_average_precision,
)
from metrics_b200.functional.classification.precision_recall_curve import (  # noqa: F401,E402
    binary_precision_recall_curve,
    multiclass_precision_recall_curve,
)
from metrics_b200.functional.classification.roc import binary_roc, multiclass_roc  # noqa: F401,E402
from metrics_b200.functional.classification.accuracy import binary_accuracy, multilabel_accuracy  # noqa: F401,E402
from metrics_b200.functional.classification.confusion_matrix import (  # noqa: F401,E402
    binary_confusion_matrix,
    confusion_matrix,
    multilabel_confusion_matrix,
)
from metrics_b200.functional.classification.f_beta import (  # noqa: F401,E402
    binary_f1_score,
    binary_fbeta_score,
    multilabel_f1_score,
    multilabel_fbeta_score,
)
from metrics_b200.functional.classification.stat_scores import (  # noqa: F401,E402
    binary_stat_scores,
    multilabel_stat_scores,
    stat_scores,
)
from metrics_b200.functional.classification.auroc import multilabel_auroc  # noqa: F401,E402
from metrics_b200.functional.classification.average_precision import multilabel_average_precision  # noqa: F401,E402
from metrics_b200.functional.classification.precision_recall_curve import multilabel_precision_recall_curve  # noqa: F401,E402
from metrics_b200.functional.classification.roc import multilabel_roc  # noqa: F401,E402
