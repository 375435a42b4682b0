"""Generate golden vectors from the UNMODIFIED reference (TorchMetrics under /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference is imported from /root/reference/src through the documented stand-in for its missing
`lightning_utilities` dependency (tests/golden/_standins/).  No reference source is copied
or modified.  Inputs that are too large to commit are regenerated from their seed by the tests and pinned by a
sha256 of their bytes stored here.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_standins"))
sys.path.insert(0, "/root/reference/src")

import torchmetrics  # noqa: E402
from torchmetrics.classification import (  # noqa: E402
    MulticlassAccuracy,
    MulticlassConfusionMatrix,
    MulticlassF1Score,
    MulticlassStatScores,
)
from torchmetrics.functional.classification import (  # noqa: E402
    multiclass_accuracy,
    multiclass_confusion_matrix,
    multiclass_f1_score,
    multiclass_fbeta_score,
    multiclass_stat_scores,
)


def sha(t: torch.Tensor) -> str:
    t = t.contiguous()
    if t.dtype == torch.bfloat16:
        t = t.view(torch.int16)
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()


def np_of(t: torch.Tensor) -> np.ndarray:
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.float().numpy()
    return t.numpy()


def edge_rows(C: int) -> torch.Tensor:
    """Adversarial rows for argmax: ties, NaN, +-inf, signed zeros, max in the last / first column."""
    g = torch.Generator().manual_seed(100 + C)
    rows = [torch.randn(C, generator=g) for _ in range(24)]
    nan, inf = float("nan"), float("inf")
    r = rows
    r[0][:] = 1.0  # all equal -> 0
    r[1][C // 2] = 50.0
    r[1][C - 1] = 50.0  # tie -> first
    r[2][C - 1] = 99.0  # max in last column
    r[3][0] = 99.0
    r[4][C // 3] = nan  # NaN is max
    r[5][C // 3] = nan
    r[5][C - 1] = nan  # first NaN wins
    r[6][1] = inf
    r[6][C - 2] = inf
    r[7][:] = -inf  # all -inf -> 0
    r[8][:] = -1.0
    r[8][2] = -0.0
    r[8][C - 1] = 0.0  # -0.0 == 0.0, first wins
    r[9][:] = -1.0
    r[9][C - 1] = -0.0
    r[10][2] = inf
    r[10][1] = nan  # NaN beats inf
    r[11][:] = nan
    r[12][C - 1] = nan
    r[13][:] = -inf
    r[13][C - 1] = -1e30
    r[14][:] = 0.0
    r[14][C // 2] = 1e-30  # f32 denormal-ish / underflows to 0 in f16
    return torch.stack(rows)


def classification_golden() -> dict:
    out: dict = {}

    # ---- A. argmax edge semantics per dtype and row width -------------------------------------------------------
    for C in (4, 37, 64, 1000, 1024, 2500):
        base = edge_rows(C)
        for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16), ("f16", torch.float16), ("f64", torch.float64)):
            x = base.to(dt)
            out[f"argmax/{name}/C{C}/x"] = np_of(x) if dt != torch.float64 else x.numpy()
            out[f"argmax/{name}/C{C}/y"] = x.argmax(dim=1).numpy()

    # ---- B. confusion matrix cases ----------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    for C, N in ((5, 128), (37, 512), (130, 2048)):
        logits = torch.randn(N, C, generator=g)
        target = torch.randint(0, C, (N,), generator=g)
        labels = torch.randint(0, C, (N,), generator=g)
        out[f"confmat/C{C}/logits"] = logits.numpy()
        out[f"confmat/C{C}/target"] = target.numpy()
        out[f"confmat/C{C}/labels"] = labels.numpy()
        for ign in (None, -1, 0):
            t = target.clone()
            if ign == -1:
                t[::7] = -1
            tag = "none" if ign is None else str(ign)
            out[f"confmat/C{C}/ign{tag}/target"] = t.numpy()
            out[f"confmat/C{C}/ign{tag}/from_logits"] = multiclass_confusion_matrix(logits, t, C, ignore_index=ign).numpy()
            out[f"confmat/C{C}/ign{tag}/from_labels"] = multiclass_confusion_matrix(labels, t, C, ignore_index=ign).numpy()
        for norm in ("true", "pred", "all"):
            out[f"confmat/C{C}/norm_{norm}"] = multiclass_confusion_matrix(logits, target, C, normalize=norm).numpy()
    # multidim (N, C, d1, d2) and small integer dtypes
    md_logits = torch.randn(16, 6, 5, 3, generator=g)
    md_target = torch.randint(0, 6, (16, 5, 3), generator=g)
    out["confmat/multidim/logits"] = md_logits.numpy()
    out["confmat/multidim/target"] = md_target.numpy()
    out["confmat/multidim/confmat"] = multiclass_confusion_matrix(md_logits, md_target, 6).numpy()
    u8p = torch.randint(0, 200, (300,), generator=g).to(torch.uint8)
    u8t = torch.randint(0, 200, (300,), generator=g).to(torch.uint8)
    out["confmat/uint8/preds"] = u8p.numpy()
    out["confmat/uint8/target"] = u8t.numpy()
    out["confmat/uint8/confmat"] = multiclass_confusion_matrix(u8p, u8t, 200).numpy()

    # ---- C. cfg1: MulticlassAccuracy(5), 100 updates of [1024, 5] -------------------------------------------------
    g = torch.Generator().manual_seed(0)
    preds = torch.randn(100, 1024, 5, generator=g)
    target = torch.randint(0, 5, (100, 1024), generator=g)
    m = MulticlassAccuracy(num_classes=5)
    for i in range(100):
        m.update(preds[i], target[i])
    out["cfg1/value"] = m.compute().numpy()
    for s in ("tp", "fp", "tn", "fn"):
        out[f"cfg1/{s}"] = getattr(m, s).numpy()
    out["cfg1/preds_sha256"] = np.array(sha(preds))
    out["cfg1/target_sha256"] = np.array(sha(target))

    # ---- D. cfg2: MulticlassConfusionMatrix(1000), [65536, 1000] bf16 -----------------------------------------------
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(65536, 1000, generator=g).bfloat16()
    target = torch.randint(0, 1000, (65536,), generator=g)
    cm = MulticlassConfusionMatrix(num_classes=1000)
    cm.update(logits, target)
    confmat = cm.compute()
    out["cfg2/argmax_i16"] = logits.argmax(dim=1).to(torch.int16).numpy()
    out["cfg2/confmat_sha256"] = np.array(sha(confmat))
    out["cfg2/confmat_rowsum"] = confmat.sum(1).numpy()
    out["cfg2/confmat_colsum"] = confmat.sum(0).numpy()
    out["cfg2/confmat_diag"] = confmat.diag().numpy()
    out["cfg2/logits_sha256"] = np.array(sha(logits))
    out["cfg2/target_sha256"] = np.array(sha(target))
    out["cfg2/tied_rows"] = np.array(int((logits == logits.max(dim=1, keepdim=True).values).sum(1).gt(1).sum()))

    # ---- E. stat scores / accuracy / f-beta ----------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    for C, N in ((5, 300), (1000, 4096)):
        logits = torch.randn(N, C, generator=g)
        target = torch.randint(0, C, (N,), generator=g)
        if C == 5:
            target[target == 3] = 1  # class 3 never occurs in target
            out[f"stats/C{C}/logits"] = logits.numpy()
            out[f"stats/C{C}/target"] = target.numpy()
        else:
            out[f"stats/C{C}/logits_sha256"] = np.array(sha(logits))
            out[f"stats/C{C}/target_sha256"] = np.array(sha(target))
        for avg in ("micro", "macro", "weighted", "none"):
            for ign in (None, -1, 1):
                t = target.clone()
                if ign == -1:
                    t[::5] = -1
                tag = f"stats/C{C}/{avg}/ign{'none' if ign is None else ign}"
                out[f"{tag}/stat_scores"] = multiclass_stat_scores(logits, t, C, average=avg, ignore_index=ign).numpy()
                out[f"{tag}/accuracy"] = multiclass_accuracy(logits, t, C, average=avg, ignore_index=ign).numpy()
                out[f"{tag}/f1"] = multiclass_f1_score(logits, t, C, average=avg, ignore_index=ign).numpy()
                out[f"{tag}/fbeta2"] = multiclass_fbeta_score(logits, t, 2.0, C, average=avg, ignore_index=ign).numpy()
        # modular class over 4 batches (states)
        for avg in ("micro", "macro"):
            mm = MulticlassStatScores(num_classes=C, average=avg)
            f1 = MulticlassF1Score(num_classes=C, average=avg)
            for chunk_l, chunk_t in zip(logits.chunk(4), target.chunk(4)):
                mm.update(chunk_l, chunk_t)
                f1.update(chunk_l, chunk_t)
            for s in ("tp", "fp", "tn", "fn"):
                out[f"stats/C{C}/{avg}/class_state_{s}"] = getattr(mm, s).numpy()
            out[f"stats/C{C}/{avg}/class_f1"] = f1.compute().numpy()
    # ---- F. binary / multilabel stat scores, confusion matrices, accuracy, f1 -------------------------------------------
    from torchmetrics.functional.classification import (
        binary_accuracy,
        binary_confusion_matrix,
        binary_f1_score,
        binary_stat_scores,
        multilabel_accuracy,
        multilabel_confusion_matrix,
        multilabel_f1_score,
        multilabel_stat_scores,
    )

    g = torch.Generator().manual_seed(17)
    bp = {"probs": torch.rand(64, 7, generator=g), "logits": torch.randn(64, 7, generator=g) * 2,
          "labels": torch.randint(0, 2, (64, 7), generator=g)}
    bt = torch.randint(0, 2, (64, 7), generator=g)
    bt_ign = bt.clone()
    bt_ign[::5] = -1
    out["bin2/target"], out["bin2/target_ign"] = bt.numpy(), bt_ign.numpy()
    for kind, p in bp.items():
        out[f"bin2/{kind}/preds"] = p.numpy()
        for ign, t in ((None, bt), (-1, bt_ign)):
            for mda in ("global", "samplewise"):
                tag = f"bin2/{kind}/ign{'none' if ign is None else ign}/{mda}"
                out[f"{tag}/stat_scores"] = binary_stat_scores(p, t, multidim_average=mda, ignore_index=ign).numpy()
                out[f"{tag}/accuracy"] = binary_accuracy(p, t, multidim_average=mda, ignore_index=ign).numpy()
                out[f"{tag}/f1"] = binary_f1_score(p, t, multidim_average=mda, ignore_index=ign).numpy()
            out[f"bin2/{kind}/ign{'none' if ign is None else ign}/confmat"] = binary_confusion_matrix(p, t, ignore_index=ign).numpy()
    out["bin2/probs/thr0.3/stat_scores"] = binary_stat_scores(bp["probs"], bt, threshold=0.3).numpy()
    L = 6
    mp = {"probs": torch.rand(40, L, 5, generator=g), "logits": torch.randn(40, L, 5, generator=g) * 2,
          "labels": torch.randint(0, 2, (40, L, 5), generator=g)}
    mt = torch.randint(0, 2, (40, L, 5), generator=g)
    mt_ign = mt.clone()
    mt_ign[::3] = -1
    out["ml/target"], out["ml/target_ign"] = mt.numpy(), mt_ign.numpy()
    for kind, p in mp.items():
        out[f"ml/{kind}/preds"] = p.numpy()
        for ign, t in ((None, mt), (-1, mt_ign)):
            for mda in ("global", "samplewise"):
                for avg in ("micro", "macro", "weighted", "none"):
                    tag = f"ml/{kind}/ign{'none' if ign is None else ign}/{mda}/{avg}"
                    out[f"{tag}/stat_scores"] = multilabel_stat_scores(p, t, L, average=avg, multidim_average=mda, ignore_index=ign).numpy()
                    out[f"{tag}/accuracy"] = multilabel_accuracy(p, t, L, average=avg, multidim_average=mda, ignore_index=ign).numpy()
                    out[f"{tag}/f1"] = multilabel_f1_score(p, t, L, average=avg, multidim_average=mda, ignore_index=ign).numpy()
            out[f"ml/{kind}/ign{'none' if ign is None else ign}/confmat"] = multilabel_confusion_matrix(p, t, L, ignore_index=ign).numpy()
    # ---- G. multiclass top-k and samplewise stat scores --------------------------------------------------------------
    g = torch.Generator().manual_seed(23)
    tk_logits = torch.randn(300, 7, generator=g)
    tk_target = torch.randint(0, 7, (300,), generator=g)
    tk_ign = tk_target.clone()
    tk_ign[::6] = -1
    out["topk/logits"], out["topk/target"], out["topk/target_ign"] = tk_logits.numpy(), tk_target.numpy(), tk_ign.numpy()
    for k in (2, 3):
        for avg in ("micro", "macro", "none"):
            for ign, t in ((None, tk_target), (-1, tk_ign)):
                tag = f"topk/k{k}/{avg}/ign{'none' if ign is None else ign}"
                out[f"{tag}/stat_scores"] = multiclass_stat_scores(tk_logits, t, 7, average=avg, top_k=k, ignore_index=ign).numpy()
                out[f"{tag}/accuracy"] = multiclass_accuracy(tk_logits, t, 7, average=avg, top_k=k, ignore_index=ign).numpy()
                out[f"{tag}/f1"] = multiclass_f1_score(tk_logits, t, 7, average=avg, top_k=k, ignore_index=ign).numpy()
    mk = MulticlassAccuracy(num_classes=7, top_k=2, average="micro")
    for a, b in zip(tk_logits.chunk(3), tk_target.chunk(3)):
        mk.update(a, b)
    out["topk/class_acc_k2_micro"] = mk.compute().numpy()
    sw_logits = torch.randn(20, 5, 11, generator=g)
    sw_labels = torch.randint(0, 5, (20, 11), generator=g)
    sw_target = torch.randint(0, 5, (20, 11), generator=g)
    out["sw/logits"], out["sw/labels"], out["sw/target"] = sw_logits.numpy(), sw_labels.numpy(), sw_target.numpy()
    for kind, p in (("logits", sw_logits), ("labels", sw_labels)):
        for avg in ("micro", "macro", "none"):
            for ign in (None, -1, 1):
                t = sw_target.clone()
                if ign == -1:
                    t[:, ::4] = -1
                tag = f"sw/{kind}/{avg}/ign{'none' if ign is None else ign}"
                out[f"{tag}/stat_scores"] = multiclass_stat_scores(p, t, 5, average=avg, multidim_average="samplewise", ignore_index=ign).numpy()
                out[f"{tag}/accuracy"] = multiclass_accuracy(p, t, 5, average=avg, multidim_average="samplewise", ignore_index=ign).numpy()
    msw = MulticlassStatScores(num_classes=5, average="none", multidim_average="samplewise")
    msw.update(sw_logits[:8], sw_target[:8])
    msw.update(sw_logits[8:], sw_target[8:])
    out["sw/class_none"] = msw.compute().numpy()
    out["meta/torchmetrics_version"] = np.array(torchmetrics.__version__)
    out["meta/torch_version"] = np.array(torch.__version__)
    return out


def curves_golden() -> dict:
    from torchmetrics import MetricCollection
    from torchmetrics.classification import (
        BinaryAUROC,
        BinaryAveragePrecision,
        MulticlassAUROC,
        MulticlassAveragePrecision,
    )
    from torchmetrics.functional.classification import (
        binary_auroc,
        binary_average_precision,
        binary_precision_recall_curve,
        binary_roc,
        multiclass_auroc,
        multiclass_average_precision,
        multiclass_precision_recall_curve,
        multiclass_roc,
    )
    from torchmetrics.functional.classification.precision_recall_curve import _binary_clf_curve

    import warnings

    warnings.simplefilter("ignore")
    out: dict = {}
    # ---- binary cases --------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(21)
    cases = {}
    cases["doc"] = (torch.tensor([0.1, 0.4, 0.35, 0.8]), torch.tensor([0, 0, 1, 1]))
    cases["rand"] = (torch.rand(3000, generator=g), torch.randint(0, 2, (3000,), generator=g))
    cases["ties"] = ((torch.rand(5000, generator=g) * 50).floor() / 50, torch.randint(0, 2, (5000,), generator=g))
    cases["logits"] = (torch.randn(2000, generator=g) * 3, torch.randint(0, 2, (2000,), generator=g))
    cases["allpos"] = (torch.rand(257, generator=g), torch.ones(257, dtype=torch.long))
    cases["allneg"] = (torch.rand(257, generator=g), torch.zeros(257, dtype=torch.long))
    cases["alltied"] = (torch.full((100,), 0.5), torch.randint(0, 2, (100,), generator=g))
    cases["one"] = (torch.tensor([0.3]), torch.tensor([1]))
    cases["odd"] = (torch.rand(4097, generator=g), torch.randint(0, 2, (4097,), generator=g))
    cases["skew"] = (torch.rand(20000, generator=g), (torch.rand(20000, generator=g) < 0.01).long())
    for name, (p, t) in cases.items():
        out[f"bin/{name}/preds"] = p.numpy()
        out[f"bin/{name}/target"] = t.numpy()
        out[f"bin/{name}/auroc"] = binary_auroc(p, t).numpy()
        out[f"bin/{name}/ap"] = binary_average_precision(p, t).numpy()
        pf = torch.sigmoid(p) if name == "logits" else p
        fps, tps, thr = _binary_clf_curve(pf, t)
        out[f"bin/{name}/clf_fps"], out[f"bin/{name}/clf_tps"], out[f"bin/{name}/clf_thr"] = fps.numpy(), tps.numpy(), thr.numpy()
        fpr, tpr, th = binary_roc(p, t)
        out[f"bin/{name}/roc_fpr"], out[f"bin/{name}/roc_tpr"], out[f"bin/{name}/roc_thr"] = fpr.numpy(), tpr.numpy(), th.numpy()
        pr, rc, th = binary_precision_recall_curve(p, t)
        out[f"bin/{name}/prc_p"], out[f"bin/{name}/prc_r"], out[f"bin/{name}/prc_thr"] = pr.numpy(), rc.numpy(), th.numpy()
        for mf in (0.5, 0.8):
            out[f"bin/{name}/auroc_maxfpr{mf}"] = binary_auroc(p, t, max_fpr=mf).numpy()
    p, t = cases["rand"]
    t2 = t.clone()
    t2[::9] = -1
    out["bin/ignore/target"] = t2.numpy()
    out["bin/ignore/auroc"] = binary_auroc(p, t2, ignore_index=-1).numpy()
    out["bin/ignore/ap"] = binary_average_precision(p, t2, ignore_index=-1).numpy()
    pb = cases["logits"][0].bfloat16()
    out["bin/bf16/auroc"] = binary_auroc(pb, cases["logits"][1]).numpy()
    out["bin/bf16/ap"] = binary_average_precision(pb, cases["logits"][1]).numpy()
    out["bin/bf16/roc_thr"] = binary_roc(pb, cases["logits"][1])[2].float().numpy()

    # ---- cfg3: 1000 x 10000 samples through a MetricCollection ----------------------------------------------------------
    g = torch.Generator().manual_seed(0)
    preds = torch.rand(1000, 10000, generator=g)
    target = torch.randint(0, 2, (1000, 10000), generator=g)
    mc = MetricCollection([BinaryAUROC(), BinaryAveragePrecision()])
    for i in range(1000):
        mc.update(preds[i], target[i])
    res = mc.compute()
    out["cfg3/auroc"] = res["BinaryAUROC"].numpy()
    out["cfg3/ap"] = res["BinaryAveragePrecision"].numpy()
    out["cfg3/preds_sha256"] = np.array(sha(preds))
    out["cfg3/target_sha256"] = np.array(sha(target))
    out["cfg3/compute_groups"] = np.array(str(mc.compute_groups))

    # ---- multiclass ------------------------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(33)
    for C, N, kind in ((5, 400, "probs"), (5, 400, "logits"), (37, 1500, "logits"), (1000, 2048, "logits")):
        logits = torch.randn(N, C, generator=g)
        tgt = torch.randint(0, C, (N,), generator=g)
        if C == 5:
            tgt[tgt == 4] = 2  # class 4 has no positive sample
        p = torch.softmax(logits, 1) if kind == "probs" else logits
        key = f"mc/C{C}_{kind}"
        if C <= 37:
            out[f"{key}/preds"], out[f"{key}/target"] = p.numpy(), tgt.numpy()
        else:
            out[f"{key}/preds_sha256"], out[f"{key}/target_sha256"] = np.array(sha(p)), np.array(sha(tgt))
        for avg in ("macro", "weighted", "none"):
            out[f"{key}/auroc_{avg}"] = multiclass_auroc(p, tgt, C, average=avg).numpy()
            out[f"{key}/ap_{avg}"] = multiclass_average_precision(p, tgt, C, average=avg).numpy()
        if C == 5:
            fpr, tpr, thr = multiclass_roc(p, tgt, C)
            pr, rc, th2 = multiclass_precision_recall_curve(p, tgt, C)
            for c in range(C):
                out[f"{key}/roc_fpr{c}"], out[f"{key}/roc_tpr{c}"], out[f"{key}/roc_thr{c}"] = fpr[c].numpy(), tpr[c].numpy(), thr[c].numpy()
                out[f"{key}/prc_p{c}"], out[f"{key}/prc_r{c}"], out[f"{key}/prc_thr{c}"] = pr[c].numpy(), rc[c].numpy(), th2[c].numpy()
            t3 = tgt.clone()
            t3[::6] = -1
            out[f"{key}/ignore_target"] = t3.numpy()
            out[f"{key}/ignore_auroc"] = multiclass_auroc(p, t3, C, average="none", ignore_index=-1).numpy()
    # modular: cfg5-like single rank, 4 batches of [4096, 1000] fp32 logits (rank 0 stream: torch.manual_seed(0))
    torch.manual_seed(0)
    m_auc = MulticlassAUROC(num_classes=1000)
    m_ap = MulticlassAveragePrecision(num_classes=1000)
    for _ in range(2):
        lg = torch.randn(4096, 1000)
        tg = torch.randint(0, 1000, (4096,))
        m_auc.update(lg, tg)
        m_ap.update(lg, tg)
    out["mc/cfg5_rank0_2batches/auroc"] = m_auc.compute().numpy()
    out["mc/cfg5_rank0_2batches/ap"] = m_ap.compute().numpy()
    return out


def synth_detection(seed: int, n_img: int, n_gt: int, n_det: int, n_cls: int, crowd_frac: float = 0.0, dup_scores: bool = False):
    """Scaled-down BASELINE cfg4 recipe: 640x480 images, jittered-gt detections + random ones."""
    g = torch.Generator().manual_seed(seed)
    preds, target = [], []
    for _ in range(n_img):
        x1 = torch.rand(n_gt, generator=g) * 540
        y1 = torch.rand(n_gt, generator=g) * 380
        w = 8 + torch.rand(n_gt, generator=g) * 192
        h = 8 + torch.rand(n_gt, generator=g) * 192
        gt = torch.stack([x1, y1, (x1 + w).clamp(max=640), (y1 + h).clamp(max=480)], 1)
        gl = torch.randint(0, n_cls, (n_gt,), generator=g)
        crowd = (torch.rand(n_gt, generator=g) < crowd_frac).long()
        n_jit = min(n_gt, n_det)
        jit = gt[:n_jit] + torch.randn(n_jit, 4, generator=g) * 0.1 * torch.stack([w, h, w, h], 1)[:n_jit]
        jl = torch.where(torch.rand(n_jit, generator=g) < 0.9, gl[:n_jit], torch.randint(0, n_cls, (n_jit,), generator=g))
        n_rand = n_det - n_jit
        rx = torch.rand(n_rand, generator=g) * 540
        ry = torch.rand(n_rand, generator=g) * 380
        rnd = torch.stack([rx, ry, rx + 8 + torch.rand(n_rand, generator=g) * 192, ry + 8 + torch.rand(n_rand, generator=g) * 192], 1)
        boxes = torch.cat([jit, rnd])
        boxes = torch.stack([boxes[:, 0].clamp(0, 639), boxes[:, 1].clamp(0, 479), boxes[:, 2], boxes[:, 3]], 1)
        boxes[:, 2] = torch.maximum(boxes[:, 2], boxes[:, 0] + 1)
        boxes[:, 3] = torch.maximum(boxes[:, 3], boxes[:, 1] + 1)
        labels = torch.cat([jl, torch.randint(0, n_cls, (n_rand,), generator=g)])
        scores = torch.rand(n_det, generator=g)
        if dup_scores:
            scores = (scores * 20).floor() / 20
        preds.append({"boxes": boxes, "scores": scores, "labels": labels})
        t = {"boxes": gt, "labels": gl}
        if crowd_frac > 0:
            t["iscrowd"] = crowd
        target.append(t)
    return preds, target


def map_golden() -> dict:
    """Known answers for detection mAP.  pycocotools is unavailable, so the only machine-generated vectors come from the
    reference's LEGACY in-tree evaluator (detection/_mean_ap.py, pure torch), which agrees with COCOeval only for the
    'all'-area statistics on crowd-free data; its availability guard is bypassed with a dummy
    `pycocotools.mask` module, exactly as the survey did."""
    import types

    sys.modules.setdefault("pycocotools", types.ModuleType("pycocotools"))
    sys.modules.setdefault("pycocotools.mask", types.ModuleType("pycocotools.mask"))
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    import torchmetrics.detection._mean_ap as legacy

    legacy._PYCOCOTOOLS_AVAILABLE = True
    out: dict = {}
    for name, kw in (("small", dict(seed=5, n_img=12, n_gt=6, n_det=20, n_cls=4)),
                     ("mid", dict(seed=6, n_img=60, n_gt=10, n_det=40, n_cls=8)),
                     ("dup", dict(seed=7, n_img=30, n_gt=8, n_det=30, n_cls=5, dup_scores=True))):
        preds, target = synth_detection(**kw)
        m = legacy.MeanAveragePrecision(class_metrics=True)
        m.update(preds, target)
        res = m.compute()
        for k in ("map", "map_50", "map_75", "mar_1", "mar_10", "mar_100", "map_per_class", "mar_100_per_class"):
            out[f"legacy/{name}/{k}"] = res[k].numpy()
        out[f"legacy/{name}/kw"] = np.array(repr(kw))
    # the class docstring example (detection/mean_ap.py:250-283)
    out["doc/map"] = np.float32(0.6)
    return out


def binned_golden() -> dict:
    import warnings

    from torchmetrics.classification import BinaryAUROC, MulticlassAveragePrecision, MulticlassPrecisionRecallCurve
    from torchmetrics.functional.classification import (
        binary_auroc,
        binary_average_precision,
        binary_precision_recall_curve,
        binary_roc,
        multiclass_auroc,
        multiclass_average_precision,
        multiclass_precision_recall_curve,
        multiclass_roc,
    )
    from torchmetrics.functional.classification.precision_recall_curve import (
        _binary_precision_recall_curve_format,
        _binary_precision_recall_curve_update,
        _multiclass_precision_recall_curve_format,
        _multiclass_precision_recall_curve_update,
    )

    warnings.simplefilter("ignore")
    out: dict = {}
    g = torch.Generator().manual_seed(55)
    bp = torch.rand(5000, generator=g)
    bp[::13] = (bp[::13] * 10).round() / 10  # scores sitting exactly on thresholds
    bl = torch.randn(3000, generator=g) * 2
    bt = torch.randint(0, 2, (5000,), generator=g)
    out["b/preds"], out["b/logits"], out["b/target"] = bp.numpy(), bl.numpy(), bt.numpy()
    thr_sets = {"int11": 11, "int200": 200, "list": [0.9, 0.1, 0.5, 0.3], "tensor": torch.tensor([0.0, 0.2, 0.7, 1.0])}
    for name, thr in thr_sets.items():
        for kind, (p, t) in (("probs", (bp, bt)), ("logits", (bl, bt[:3000]))):
            tag = f"b/{name}/{kind}"
            pf, tf, th = _binary_precision_recall_curve_format(p, t, thr)
            out[f"{tag}/confmat"] = _binary_precision_recall_curve_update(pf, tf, th).numpy()
            out[f"{tag}/auroc"] = binary_auroc(p, t, thresholds=thr).numpy()
            out[f"{tag}/auroc_maxfpr"] = binary_auroc(p, t, thresholds=thr, max_fpr=0.6).numpy()
            out[f"{tag}/ap"] = binary_average_precision(p, t, thresholds=thr).numpy()
            f, tp_, h = binary_roc(p, t, thresholds=thr)
            out[f"{tag}/roc_fpr"], out[f"{tag}/roc_tpr"], out[f"{tag}/roc_thr"] = f.numpy(), tp_.numpy(), h.numpy()
            pr, rc, h = binary_precision_recall_curve(p, t, thresholds=thr)
            out[f"{tag}/prc_p"], out[f"{tag}/prc_r"], out[f"{tag}/prc_thr"] = pr.numpy(), rc.numpy(), h.numpy()
    C = 6
    ml = torch.randn(1200, C, generator=g)
    mt = torch.randint(0, C, (1200,), generator=g)
    mt[mt == 5] = 2
    out["m/logits"], out["m/target"] = ml.numpy(), mt.numpy()
    for name, thr in (("int7", 7), ("list", [0.05, 0.2, 0.6])):
        pf, tf, th = _multiclass_precision_recall_curve_format(ml, mt, C, thr)
        out[f"m/{name}/confmat"] = _multiclass_precision_recall_curve_update(pf, tf, C, th).numpy()
        for avg in ("macro", "weighted", "none"):
            out[f"m/{name}/auroc_{avg}"] = multiclass_auroc(ml, mt, C, average=avg, thresholds=thr).numpy()
            out[f"m/{name}/ap_{avg}"] = multiclass_average_precision(ml, mt, C, average=avg, thresholds=thr).numpy()
        f, tp_, h = multiclass_roc(ml, mt, C, thresholds=thr)
        out[f"m/{name}/roc_fpr"], out[f"m/{name}/roc_tpr"], out[f"m/{name}/roc_thr"] = f.numpy(), tp_.numpy(), h.numpy()
        pr, rc, h = multiclass_precision_recall_curve(ml, mt, C, thresholds=thr)
        out[f"m/{name}/prc_p"], out[f"m/{name}/prc_r"] = pr.numpy(), rc.numpy()
        for avg in ("micro", "macro"):
            pr, rc, h = multiclass_precision_recall_curve(ml, mt, C, thresholds=thr, average=avg)
            out[f"m/{name}/prc_{avg}_p"], out[f"m/{name}/prc_{avg}_r"], out[f"m/{name}/prc_{avg}_thr"] = pr.numpy(), rc.numpy(), h.numpy()
    # modular, several updates
    m = BinaryAUROC(thresholds=50)
    for a, b in zip(bp.chunk(5), bt.chunk(5)):
        m.update(a, b)
    out["class/binary_auroc_50"] = m.compute().numpy()
    out["class/binary_auroc_50_confmat"] = m.confmat.numpy()
    m2 = MulticlassAveragePrecision(num_classes=C, thresholds=20)
    for a, b in zip(ml.chunk(3), mt.chunk(3)):
        m2.update(a, b)
    out["class/mc_ap_20"] = m2.compute().numpy()
    return out


def multilabel_golden() -> dict:
    """Multilabel curve family (exact + binned) from the unmodified reference."""
    import warnings

    from torchmetrics.classification import MultilabelAUROC, MultilabelAveragePrecision
    from torchmetrics.functional.classification import (
        multilabel_auroc,
        multilabel_average_precision,
        multilabel_precision_recall_curve,
        multilabel_roc,
    )
    from torchmetrics.functional.classification.precision_recall_curve import (
        _multilabel_precision_recall_curve_format,
        _multilabel_precision_recall_curve_update,
    )

    warnings.simplefilter("ignore")
    out: dict = {}
    g = torch.Generator().manual_seed(77)
    cases = {}
    cases["L4_probs"] = (torch.rand(500, 4, generator=g), torch.randint(0, 2, (500, 4), generator=g))
    lg = torch.randn(1300, 6, generator=g) * 2
    tg = torch.randint(0, 2, (1300, 6), generator=g)
    tg[:, 5] = 0  # label 5 has no positive
    tg[:, 4] = 1  # label 4 has no negative
    cases["L6_logits"] = (lg, tg)
    cases["L3_ties"] = ((torch.rand(2100, 3, generator=g) * 20).floor() / 20, torch.randint(0, 2, (2100, 3), generator=g))
    cases["L5_extra"] = (torch.rand(40, 5, 7, generator=g), torch.randint(0, 2, (40, 5, 7), generator=g))  # extra dim
    for name, (p, t) in cases.items():
        L = p.shape[1]
        out[f"{name}/preds"], out[f"{name}/target"] = p.numpy(), t.numpy()
        ti = t.clone()
        ti.view(-1)[::7] = -1
        out[f"{name}/target_ignore"] = ti.numpy()
        for tag, tt, ig in (("", t, None), ("ign_", ti, -1)):
            for avg in ("micro", "macro", "weighted", "none"):
                out[f"{name}/{tag}auroc_{avg}"] = multilabel_auroc(p, tt, L, average=avg, ignore_index=ig).numpy()
                out[f"{name}/{tag}ap_{avg}"] = multilabel_average_precision(p, tt, L, average=avg, ignore_index=ig).numpy()
            fpr, tpr, thr = multilabel_roc(p, tt, L, ignore_index=ig)
            pr, rc, th2 = multilabel_precision_recall_curve(p, tt, L, ignore_index=ig)
            for l in range(L):
                out[f"{name}/{tag}roc_fpr{l}"], out[f"{name}/{tag}roc_tpr{l}"], out[f"{name}/{tag}roc_thr{l}"] = fpr[l].numpy(), tpr[l].numpy(), thr[l].numpy()
                out[f"{name}/{tag}prc_p{l}"], out[f"{name}/{tag}prc_r{l}"], out[f"{name}/{tag}prc_thr{l}"] = pr[l].numpy(), rc[l].numpy(), th2[l].numpy()
            for tname, thrs in (("int9", 9), ("list", [0.8, 0.15, 0.5])):
                pf, tf, th = _multilabel_precision_recall_curve_format(p, tt, L, thrs, ig)
                out[f"{name}/{tag}{tname}/confmat"] = _multilabel_precision_recall_curve_update(pf, tf, L, th).numpy()
                for avg in ("micro", "macro", "weighted", "none"):
                    out[f"{name}/{tag}{tname}/auroc_{avg}"] = multilabel_auroc(p, tt, L, average=avg, thresholds=thrs, ignore_index=ig).numpy()
                    out[f"{name}/{tag}{tname}/ap_{avg}"] = multilabel_average_precision(p, tt, L, average=avg, thresholds=thrs, ignore_index=ig).numpy()
                f, tp_, h = multilabel_roc(p, tt, L, thresholds=thrs, ignore_index=ig)
                out[f"{name}/{tag}{tname}/roc_fpr"], out[f"{name}/{tag}{tname}/roc_tpr"], out[f"{name}/{tag}{tname}/roc_thr"] = f.numpy(), tp_.numpy(), h.numpy()
                pr, rc, h = multilabel_precision_recall_curve(p, tt, L, thresholds=thrs, ignore_index=ig)
                out[f"{name}/{tag}{tname}/prc_p"], out[f"{name}/{tag}{tname}/prc_r"] = pr.numpy(), rc.numpy()
    # modular, 3 updates
    p, t = cases["L6_logits"]
    for avg in ("macro", "micro"):
        m1, m2, m3 = MultilabelAUROC(num_labels=6, average=avg), MultilabelAveragePrecision(num_labels=6, average=avg), MultilabelAUROC(num_labels=6, average=avg, thresholds=25)
        for a, b in zip(p.chunk(3), t.chunk(3)):
            m1.update(a, b), m2.update(a, b), m3.update(a, b)
        out[f"class/auroc_{avg}"], out[f"class/ap_{avg}"], out[f"class/auroc_binned25_{avg}"] = m1.compute().numpy(), m2.compute().numpy(), m3.compute().numpy()
    return out


def consumers_golden() -> dict:
    """Stat-score / confusion-matrix consumer metrics (Precision, Recall, Specificity, NPV, Hamming, Jaccard, CohenKappa,
    MatthewsCorrCoef) from the unmodified reference, on inputs shared with the product tests."""
    import warnings

    import torchmetrics.classification as TC
    import torchmetrics.functional.classification as F

    warnings.simplefilter("ignore")
    out: dict = {}
    g = torch.Generator().manual_seed(91)
    # binary
    bp = torch.rand(700, generator=g)
    bt = torch.randint(0, 2, (700,), generator=g)
    bt_good = ((bp + 0.35 * torch.randn(700, generator=g)) > 0.5).long()  # correlated with preds
    bti = bt_good.clone()
    bti[::11] = -1
    bp_multi = torch.rand(30, 4, 5, generator=g)
    bt_multi = torch.randint(0, 2, (30, 4, 5), generator=g)
    out["b/preds"], out["b/target"], out["b/target_good"], out["b/target_ign"] = bp.numpy(), bt.numpy(), bt_good.numpy(), bti.numpy()
    out["b/preds_multi"], out["b/target_multi"] = bp_multi.numpy(), bt_multi.numpy()
    # multiclass
    C = 7
    ml = torch.randn(900, C, generator=g)
    mt = torch.randint(0, C, (900,), generator=g)
    mt[mt == 6] = 1  # class 6 never a target
    ml[torch.arange(900), mt] += 1.2  # informative
    ml[:, 5] -= 50  # class 5 never predicted
    mti = mt.clone()
    mti[::9] = -1
    ml_multi = torch.randn(20, C, 6, generator=g)
    mt_multi = torch.randint(0, C, (20, 6), generator=g)
    out["mc/logits"], out["mc/target"], out["mc/target_ign"] = ml.numpy(), mt.numpy(), mti.numpy()
    out["mc/logits_multi"], out["mc/target_multi"] = ml_multi.numpy(), mt_multi.numpy()
    # multilabel
    L = 5
    lp = torch.rand(600, L, generator=g)
    lt = ((lp + 0.4 * torch.randn(600, L, generator=g)) > 0.5).long()
    lt[:, 4] = 0  # label 4 never positive
    lti = lt.clone()
    lti.view(-1)[::13] = -1
    lp_multi = torch.rand(25, L, 4, generator=g)
    lt_multi = torch.randint(0, 2, (25, L, 4), generator=g)
    out["ml/preds"], out["ml/target"], out["ml/target_ign"] = lp.numpy(), lt.numpy(), lti.numpy()
    out["ml/preds_multi"], out["ml/target_multi"] = lp_multi.numpy(), lt_multi.numpy()

    ratio = ("precision", "recall", "specificity", "negative_predictive_value", "hamming_distance")
    zd_kinds = ("precision", "recall", "negative_predictive_value")
    for kind in ratio:
        fb, fm, fl = getattr(F, f"binary_{kind}"), getattr(F, f"multiclass_{kind}"), getattr(F, f"multilabel_{kind}")
        out[f"b/{kind}"] = fb(bp, bt_good).numpy()
        out[f"b/{kind}_ign"] = fb(bp, bti, ignore_index=-1).numpy()
        out[f"b/{kind}_thr0.8"] = fb(bp, bt_good, threshold=0.8).numpy()
        out[f"b/{kind}_samplewise"] = fb(bp_multi, bt_multi, multidim_average="samplewise").numpy()
        for avg in ("micro", "macro", "weighted", "none"):
            out[f"mc/{kind}_{avg}"] = fm(ml, mt, C, average=avg).numpy()
            out[f"mc/{kind}_{avg}_ign"] = fm(ml, mti, C, average=avg, ignore_index=-1).numpy()
            out[f"mc/{kind}_{avg}_top2"] = fm(ml, mt, C, average=avg, top_k=2).numpy()
            out[f"mc/{kind}_{avg}_samplewise"] = fm(ml_multi, mt_multi, C, average=avg, multidim_average="samplewise").numpy()
            out[f"ml/{kind}_{avg}"] = fl(lp, lt, L, average=avg).numpy()
            out[f"ml/{kind}_{avg}_ign"] = fl(lp, lti, L, average=avg, ignore_index=-1).numpy()
            out[f"ml/{kind}_{avg}_samplewise"] = fl(lp_multi, lt_multi, L, average=avg, multidim_average="samplewise").numpy()
            if kind in zd_kinds:
                out[f"mc/{kind}_{avg}_zd1"] = fm(ml, mt, C, average=avg, zero_division=1).numpy()
                out[f"ml/{kind}_{avg}_zd1"] = fl(lp, lt, L, average=avg, zero_division=1).numpy()
    for avg in ("micro", "macro", "weighted", "none"):
        out[f"mc/jaccard_{avg}"] = F.multiclass_jaccard_index(ml, mt, C, average=avg).numpy()
        out[f"mc/jaccard_{avg}_ign"] = F.multiclass_jaccard_index(ml, mti, C, average=avg, ignore_index=-1).numpy()
        out[f"mc/jaccard_{avg}_ign2"] = F.multiclass_jaccard_index(ml, mt, C, average=avg, ignore_index=2).numpy()
        out[f"mc/jaccard_{avg}_zd1"] = F.multiclass_jaccard_index(ml, mt, C, average=avg, zero_division=1.0).numpy()
        out[f"ml/jaccard_{avg}"] = F.multilabel_jaccard_index(lp, lt, L, average=avg).numpy()
        out[f"ml/jaccard_{avg}_ign"] = F.multilabel_jaccard_index(lp, lti, L, average=avg, ignore_index=-1).numpy()
    out["b/jaccard"] = F.binary_jaccard_index(bp, bt_good).numpy()
    out["b/jaccard_ign"] = F.binary_jaccard_index(bp, bti, ignore_index=-1).numpy()
    for w in ("none", "linear", "quadratic"):
        out[f"b/kappa_{w}"] = F.binary_cohen_kappa(bp, bt_good, weights=w).numpy()
        out[f"mc/kappa_{w}"] = F.multiclass_cohen_kappa(ml, mt, C, weights=w).numpy()
        out[f"mc/kappa_{w}_ign"] = F.multiclass_cohen_kappa(ml, mti, C, weights=w, ignore_index=-1).numpy()
    out["b/mcc"] = F.binary_matthews_corrcoef(bp, bt_good).numpy()
    out["b/mcc_rand"] = F.binary_matthews_corrcoef(bp, bt).numpy()
    out["b/mcc_ign"] = F.binary_matthews_corrcoef(bp, bti, ignore_index=-1).numpy()
    out["b/mcc_perfect"] = F.binary_matthews_corrcoef((bt > 0).float(), bt).numpy()
    out["b/mcc_inverse"] = F.binary_matthews_corrcoef((bt == 0).float(), bt).numpy()
    out["b/mcc_allpos_pred"] = F.binary_matthews_corrcoef(torch.ones(700), bt).numpy()
    out["b/mcc_allneg_target"] = F.binary_matthews_corrcoef(bp, torch.zeros(700, dtype=torch.long)).numpy()
    out["mc/mcc"] = F.multiclass_matthews_corrcoef(ml, mt, C).numpy()
    out["mc/mcc_ign"] = F.multiclass_matthews_corrcoef(ml, mti, C, ignore_index=-1).numpy()
    out["mc/mcc_const"] = F.multiclass_matthews_corrcoef(torch.zeros(50, dtype=torch.long), torch.zeros(50, dtype=torch.long), C).numpy()
    out["ml/mcc"] = F.multilabel_matthews_corrcoef(lp, lt, L).numpy()
    out["ml/mcc_ign"] = F.multilabel_matthews_corrcoef(lp, lti, L, ignore_index=-1).numpy()
    # modular: 3 updates each
    mods = {
        "MulticlassPrecision": TC.MulticlassPrecision(num_classes=C, average="macro"),
        "MulticlassRecall_top2": TC.MulticlassRecall(num_classes=C, average="weighted", top_k=2),
        "MulticlassSpecificity": TC.MulticlassSpecificity(num_classes=C, average="none"),
        "MulticlassHammingDistance": TC.MulticlassHammingDistance(num_classes=C, average="micro"),
        "MulticlassJaccardIndex": TC.MulticlassJaccardIndex(num_classes=C),
        "MulticlassCohenKappa": TC.MulticlassCohenKappa(num_classes=C, weights="linear"),
        "MulticlassMatthewsCorrCoef": TC.MulticlassMatthewsCorrCoef(num_classes=C),
    }
    for name, m in mods.items():
        for a, b in zip(ml.chunk(3), mt.chunk(3)):
            m.update(a, b)
        out[f"class/{name}"] = m.compute().numpy()
    mods = {
        "MultilabelPrecision": TC.MultilabelPrecision(num_labels=L, average="macro"),
        "MultilabelNegativePredictiveValue": TC.MultilabelNegativePredictiveValue(num_labels=L, average="none"),
        "MultilabelJaccardIndex": TC.MultilabelJaccardIndex(num_labels=L, average="weighted"),
        "MultilabelMatthewsCorrCoef": TC.MultilabelMatthewsCorrCoef(num_labels=L),
    }
    for name, m in mods.items():
        for a, b in zip(lp.chunk(3), lt.chunk(3)):
            m.update(a, b)
        out[f"class/{name}"] = m.compute().numpy()
    mods = {"BinaryRecall": TC.BinaryRecall(), "BinaryCohenKappa": TC.BinaryCohenKappa(), "BinaryJaccardIndex": TC.BinaryJaccardIndex(),
            "BinaryMatthewsCorrCoef": TC.BinaryMatthewsCorrCoef(), "BinaryHammingDistance": TC.BinaryHammingDistance()}
    for name, m in mods.items():
        for a, b in zip(bp.chunk(3), bt_good.chunk(3)):
            m.update(a, b)
        out[f"class/{name}"] = m.compute().numpy()
    return out


def fuzz_golden(seed: int = 2026, mult: int = 1) -> dict:
    """Randomised differential cases: random shapes / dtypes / ignore_index / averages through the reference's
    functionals.  Each case k stores `k/spec` (JSON: functional name + kwargs), `k/preds`, `k/target` and `k/out*`."""
    import json
    import warnings

    import torchmetrics.functional.classification as F
    import torchmetrics.functional.regression as _FR

    FR_holder = [_FR]
    warnings.simplefilter("ignore")
    out: dict = {}
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    k = 0

    def emit(name, kwargs, preds, target):
        nonlocal k
        try:
            res = getattr(F if hasattr(F, name) else FR_holder[0], name)(preds, target, **kwargs)
        except Exception as err:  # the reference itself cannot run this combination: not a parity case
            print("skipped", name, kwargs, tuple(preds.shape), preds.dtype, type(err).__name__, str(err)[:60])
            return
        out[f"{k}/spec"] = np.array(json.dumps({"fn": name, "kwargs": kwargs, "preds_dtype": str(preds.dtype).replace("torch.", "")}))
        out[f"{k}/preds"] = preds.float().numpy() if preds.dtype in (torch.bfloat16, torch.float16) else preds.numpy()
        out[f"{k}/target"] = target.numpy()
        if isinstance(res, (tuple, list)):
            flat = []
            for part in res:
                flat.extend(part if isinstance(part, (tuple, list)) else [part])
            out[f"{k}/n_out"] = np.array(len(flat))
            for i, t in enumerate(flat):
                out[f"{k}/out{i}"] = t.float().numpy() if t.dtype in (torch.bfloat16, torch.float16) else t.numpy()
        else:
            out[f"{k}/n_out"] = np.array(1)
            out[f"{k}/out0"] = res.float().numpy() if res.dtype in (torch.bfloat16, torch.float16) else res.numpy()
        k += 1

    def pick(*xs):
        return xs[int(rng.integers(len(xs)))]

    # ---- multiclass: confusion matrix / stat scores / accuracy / f1 / precision / jaccard -------------------------------
    for _ in range(36 * mult):
        C = int(pick(2, 3, 5, 9, 17))
        N = int(pick(1, 7, 64, 257))
        extra = pick((), (), (3,), (2, 2))
        kind = pick("logits", "logits", "probs", "labels")
        dt = pick(torch.float32, torch.float32, torch.float16, torch.bfloat16, torch.float64)
        ig = pick(None, None, -1, C - 1, 0)
        # top_k > 1 cases get tie-free scores: which of two EQUAL maxima `torch.topk` lists first is implementation-defined
        # (CPU and CUDA differ), so the reference has no single answer there; the kernels pick the lowest index like argmax
        use_topk = kind != "labels" and C > 2 and not extra and rng.random() < 0.3
        if kind == "labels":
            preds = torch.randint(0, C, (N, *extra), generator=g)
        else:
            preds = torch.randn(N, C, *extra, generator=g)
            if rng.random() < 0.4 and not use_topk:  # ties between classes
                preds = (preds * 2).round() / 2
            if kind == "probs":
                preds = torch.softmax(preds, 1)
            preds = preds.to(torch.float32 if use_topk else dt)
        target = torch.randint(0, C, (N, *extra), generator=g)
        if ig is not None:
            mask = torch.rand(N, *extra, generator=g) < 0.2
            target = torch.where(mask, torch.full_like(target, ig), target)
        fn = pick("multiclass_confusion_matrix", "multiclass_stat_scores", "multiclass_accuracy", "multiclass_f1_score",
                  "multiclass_precision", "multiclass_recall", "multiclass_jaccard_index", "multiclass_specificity")
        kw = {"num_classes": C, "ignore_index": ig}
        if fn == "multiclass_confusion_matrix":
            kw["normalize"] = pick(None, None, "true", "pred", "all")
        elif fn == "multiclass_jaccard_index":
            kw["average"] = pick("micro", "macro", "weighted", "none")
        else:
            kw["average"] = pick("micro", "macro", "weighted", "none")
            if use_topk:
                kw["top_k"] = 2
            if extra and rng.random() < 0.4:
                kw["multidim_average"] = "samplewise"
        emit(fn, kw, preds, target)
    # ---- binary / multilabel counts ----------------------------------------------------------------------------------------
    for _ in range(24 * mult):
        ml = rng.random() < 0.5
        L = int(pick(2, 3, 6))
        N = int(pick(1, 9, 130))
        extra = pick((), (), (4,))
        shape = (N, L, *extra) if ml else (N, *extra)
        kind = pick("probs", "logits", "labels")
        ig = pick(None, None, -1)
        if kind == "labels":
            preds = torch.randint(0, 2, shape, generator=g)
        elif kind == "probs":
            preds = torch.rand(shape, generator=g).to(pick(torch.float32, torch.float16, torch.float64))
        else:
            preds = (torch.randn(shape, generator=g) * 3).to(pick(torch.float32, torch.bfloat16))
        target = torch.randint(0, 2, shape, generator=g)
        if ig is not None:
            target = torch.where(torch.rand(shape, generator=g) < 0.2, torch.full_like(target, ig), target)
        thr = float(pick(0.5, 0.5, 0.25, 0.8))
        if ml:
            fn = pick("multilabel_stat_scores", "multilabel_confusion_matrix", "multilabel_accuracy", "multilabel_f1_score",
                      "multilabel_recall", "multilabel_hamming_distance", "multilabel_jaccard_index")
            kw = {"num_labels": L, "threshold": thr, "ignore_index": ig}
            if fn not in ("multilabel_confusion_matrix",):
                kw["average"] = pick("micro", "macro", "weighted", "none")
            if fn not in ("multilabel_confusion_matrix", "multilabel_jaccard_index") and extra and rng.random() < 0.5:
                kw["multidim_average"] = "samplewise"
        else:
            fn = pick("binary_stat_scores", "binary_confusion_matrix", "binary_accuracy", "binary_f1_score", "binary_precision",
                      "binary_specificity", "binary_matthews_corrcoef", "binary_cohen_kappa")
            kw = {"threshold": thr, "ignore_index": ig}
            if fn in ("binary_stat_scores", "binary_accuracy", "binary_f1_score", "binary_precision", "binary_specificity") \
                    and extra and rng.random() < 0.5:
                kw["multidim_average"] = "samplewise"
        emit(fn, kw, preds, target)
    # ---- curves ---------------------------------------------------------------------------------------------------------------
    for _ in range(30 * mult):
        task = pick("binary", "multiclass", "multilabel")
        N = int(pick(2, 33, 400, 1500))
        ig = pick(None, None, -1)
        thresholds = pick(None, None, None, 7, [0.1, 0.5, 0.9])
        ties = rng.random() < 0.4
        logits = rng.random() < 0.4
        dt = pick(torch.float32, torch.float32, torch.float16, torch.bfloat16)
        if task == "binary":
            preds = torch.randn(N, generator=g) * 2 if logits else torch.rand(N, generator=g)
            target = torch.randint(0, 2, (N,), generator=g)
            fn = pick("binary_auroc", "binary_average_precision", "binary_roc", "binary_precision_recall_curve")
            kw = {}
            if fn == "binary_auroc" and thresholds is None and rng.random() < 0.3:
                kw["max_fpr"] = float(pick(0.3, 0.7))
        elif task == "multiclass":
            C = int(pick(3, 5, 11))
            preds = torch.randn(N, C, generator=g)
            if not logits:
                preds = torch.softmax(preds, 1)
            target = torch.randint(0, C, (N,), generator=g)
            fn = pick("multiclass_auroc", "multiclass_average_precision", "multiclass_roc", "multiclass_precision_recall_curve")
            kw = {"num_classes": C}
            if fn in ("multiclass_auroc", "multiclass_average_precision"):
                kw["average"] = pick("macro", "weighted", "none")
        else:
            L = int(pick(2, 4, 7))
            preds = torch.randn(N, L, generator=g) * 2 if logits else torch.rand(N, L, generator=g)
            target = torch.randint(0, 2, (N, L), generator=g)
            fn = pick("multilabel_auroc", "multilabel_average_precision", "multilabel_roc", "multilabel_precision_recall_curve")
            kw = {"num_labels": L}
            if fn in ("multilabel_auroc", "multilabel_average_precision"):
                kw["average"] = pick("micro", "macro", "weighted", "none")
        if ties:
            preds = (preds * 8).round() / 8
            if not logits:
                preds = preds.clamp(0, 1)
        preds = preds.to(dt)
        if ig is not None:
            target = torch.where(torch.rand(target.shape, generator=g) < 0.15, torch.full_like(target, ig), target)
        kw.update({"thresholds": thresholds, "ignore_index": ig})
        emit(fn, kw, preds, target)
    # ---- exact match (multiclass / multilabel, extra dims, samplewise, ignore_index) ---------------------------------------------
    for _ in range(16 * mult):
        N, P = int(pick(1, 5, 40)), int(pick(1, 3, 6))
        ig = pick(None, None, -1)
        mda = pick("global", "samplewise")
        if rng.random() < 0.5:
            C = int(pick(2, 4, 9))
            kind = pick("logits", "labels")
            target = torch.randint(0, C, (N, P), generator=g)
            if kind == "labels":
                preds = torch.where(torch.rand(N, P, generator=g) < 0.8, target, torch.randint(0, C, (N, P), generator=g))
            else:
                preds = torch.randn(N, C, P, generator=g)
                preds.scatter_(1, target.unsqueeze(1), 4.0 * (torch.rand(N, 1, P, generator=g) < 0.8).float(), reduce="add")
            if ig is not None:
                target = torch.where(torch.rand(N, P, generator=g) < 0.2, torch.full_like(target, ig), target)
            emit("multiclass_exact_match", {"num_classes": C, "multidim_average": mda, "ignore_index": ig}, preds, target)
        else:
            L = int(pick(2, 3, 5))
            target = torch.randint(0, 2, (N, L, P), generator=g)
            noise = torch.rand(N, L, P, generator=g)
            preds = torch.where(noise < 0.85, target.float() * 0.8 + 0.1, 1 - (target.float() * 0.8 + 0.1))
            if rng.random() < 0.4:
                preds = (preds - 0.5) * 6  # logits
            if ig is not None:
                target = torch.where(torch.rand(N, L, P, generator=g) < 0.1, torch.full_like(target, ig), target)
            emit("multilabel_exact_match", {"num_labels": L, "threshold": float(pick(0.5, 0.3)), "multidim_average": mda,
                                            "ignore_index": ig}, preds, target)
    # ---- regression ------------------------------------------------------------------------------------------------------------
    import torchmetrics.functional.regression as FR

    for _ in range(26 * mult):
        fn = pick("mean_squared_error", "mean_absolute_error", "mean_absolute_percentage_error",
                  "symmetric_mean_absolute_percentage_error", "weighted_mean_absolute_percentage_error",
                  "mean_squared_log_error", "log_cosh_error", "minkowski_distance", "r2_score", "relative_squared_error",
                  "explained_variance")
        N = int(pick(2, 17, 300, 4097))
        D = int(pick(1, 1, 3, 8))
        dt = pick(torch.float32, torch.float32, torch.float64)
        shape = (N,) if D == 1 else (N, D)
        target = torch.randn(shape, generator=g).to(dt) * float(pick(1.0, 10.0))
        preds = target + torch.randn(shape, generator=g).to(dt) * float(pick(0.1, 1.0))
        kw = {}
        if fn == "mean_squared_log_error":
            preds, target = preds.abs(), target.abs()
        if fn == "mean_squared_error":
            kw = {"squared": bool(pick(True, False)), "num_outputs": D}
        elif fn == "mean_absolute_error":
            kw = {"num_outputs": D}
        elif fn == "minkowski_distance":
            kw = {"p": float(pick(1.0, 2.0, 3.5))}
        elif fn == "r2_score":
            kw = {"multioutput": pick("uniform_average", "raw_values", "variance_weighted")}
            if N > 20 and rng.random() < 0.3:
                kw["adjusted"] = 2
        elif fn == "explained_variance":
            kw = {"multioutput": pick("uniform_average", "raw_values", "variance_weighted")}
        elif fn == "relative_squared_error":
            kw = {"squared": bool(pick(True, False))}
        kk = k
        emit(fn, kw, preds, target)
        if k > kk:
            spec = json.loads(str(out[f"{kk}/spec"]))
            spec["module"] = "regression"
            out[f"{kk}/spec"] = np.array(json.dumps(spec))
    out["n_cases"] = np.array(k)
    return out


def atfixed_golden() -> dict:
    """Operating-point metrics (recall@precision, precision@recall, sensitivity@specificity, specificity@sensitivity) from
    the unmodified reference, exact and binned, on the multilabel / curves inputs regenerated here."""
    import warnings

    import torchmetrics.classification as TC
    import torchmetrics.functional.classification as F

    warnings.simplefilter("ignore")
    out: dict = {}
    g = torch.Generator().manual_seed(123)
    bt = torch.randint(0, 2, (1500,), generator=g)
    bp = (0.5 * torch.rand(1500, generator=g) + 0.35 * bt.float() + 0.1 * torch.rand(1500, generator=g)).clamp(0, 1)
    bp = (bp * 64).round() / 64  # ties
    bl = torch.randn(900, generator=g) * 2
    C = 5
    mt = torch.randint(0, C, (1200,), generator=g)
    ml = torch.randn(1200, C, generator=g)
    ml[torch.arange(1200), mt] += 1.0
    L = 4
    lt = torch.randint(0, 2, (800, L), generator=g)
    lp = (torch.rand(800, L, generator=g) * 0.7 + 0.3 * lt.float() * torch.rand(800, L, generator=g)).clamp(0, 1)
    lti = lt.clone()
    lti.view(-1)[::9] = -1
    for key, val in (("b/preds", bp), ("b/target", bt), ("b/logits", bl), ("mc/logits", ml), ("mc/target", mt),
                     ("ml/preds", lp), ("ml/target", lt), ("ml/target_ign", lti)):
        out[key] = val.numpy()
    fams = (("recall_at_fixed_precision", "min_precision"), ("precision_at_fixed_recall", "min_recall"),
            ("sensitivity_at_specificity", "min_specificity"), ("specificity_at_sensitivity", "min_sensitivity"))
    for fam, arg in fams:
        for floor in (0.0, 0.35, 0.6, 0.9, 1.0):
            for tname, thr in (("exact", None), ("int21", 21), ("list", [0.2, 0.5, 0.8])):
                tag = f"{fam}/{floor}/{tname}"
                v, t = getattr(F, f"binary_{fam}")(bp, bt, **{arg: floor}, thresholds=thr)
                out[f"b/{tag}/value"], out[f"b/{tag}/thr"] = v.numpy(), t.numpy()
                v, t = getattr(F, f"binary_{fam}")(bl, bt[:900], **{arg: floor}, thresholds=thr)
                out[f"bl/{tag}/value"], out[f"bl/{tag}/thr"] = v.numpy(), t.numpy()
                v, t = getattr(F, f"multiclass_{fam}")(ml, mt, C, **{arg: floor}, thresholds=thr)
                out[f"mc/{tag}/value"], out[f"mc/{tag}/thr"] = v.numpy(), t.numpy()
                v, t = getattr(F, f"multilabel_{fam}")(lp, lt, L, **{arg: floor}, thresholds=thr)
                out[f"ml/{tag}/value"], out[f"ml/{tag}/thr"] = v.numpy(), t.numpy()
                v, t = getattr(F, f"multilabel_{fam}")(lp, lti, L, **{arg: floor}, thresholds=thr, ignore_index=-1)
                out[f"mli/{tag}/value"], out[f"mli/{tag}/thr"] = v.numpy(), t.numpy()
    # modular: three updates
    m = TC.BinaryRecallAtFixedPrecision(min_precision=0.6)
    m2 = TC.MulticlassSpecificityAtSensitivity(num_classes=C, min_sensitivity=0.5, thresholds=30)
    for a, b in zip(bp.chunk(3), bt.chunk(3)):
        m.update(a, b)
    for a, b in zip(ml.chunk(3), mt.chunk(3)):
        m2.update(a, b)
    out["class/b_recall_at_p/value"], out["class/b_recall_at_p/thr"] = (x.numpy() for x in m.compute())
    out["class/mc_spec_at_sens/value"], out["class/mc_spec_at_sens/thr"] = (x.numpy() for x in m2.compute())
    return out


def logauc_golden() -> dict:
    """LogAUC: (i) the reference's window integration on synthetic strictly increasing ROC curves (pins the reducer; with
    repeated fpr values the reference interpolates through an UNSTABLE argsort, utilities/data.py:259, so its output is
    implementation-defined there), (ii) end-to-end cases whose ROC curves have no repeated fpr next to the window ends."""
    import warnings

    import torchmetrics  # noqa: F401
    import torchmetrics.functional.classification as F

    L = sys.modules["torchmetrics.functional.classification.logauc"]
    warnings.simplefilter("ignore")
    out: dict = {}
    g = torch.Generator().manual_seed(321)
    for k in range(12):
        n = int(torch.randint(5, 400, (1,), generator=g))
        fpr = torch.cat([torch.zeros(1), torch.cumsum(torch.rand(n, generator=g) + 1e-3, 0)])
        fpr = fpr / fpr[-1]
        tpr = torch.cat([torch.zeros(1), torch.cumsum(torch.rand(n, generator=g), 0)])
        tpr = tpr / tpr[-1]
        out[f"curve/{k}/fpr"], out[f"curve/{k}/tpr"] = fpr.numpy(), tpr.numpy()
        for j, rng in enumerate(((0.001, 0.1), (0.01, 0.5), (0.0005, 1.0))):
            out[f"curve/{k}/logauc{j}"] = L._binary_logauc_compute(fpr, tpr, rng).numpy()
    # end to end, exact mode, continuous scores (distinct thresholds); negatives dominate so fpr moves at almost every step
    n = 3000
    t = (torch.rand(n, generator=g) < 0.05).long()
    p = (torch.rand(n, generator=g) * 0.7 + 0.3 * t.float() * torch.rand(n, generator=g)).clamp(0, 1)
    out["b/preds"], out["b/target"] = p.numpy(), t.numpy()
    for j, rng in enumerate(((0.001, 0.1), (0.01, 0.5))):
        out[f"b/logauc{j}"] = F.binary_logauc(p, t, fpr_range=rng).numpy()
    return out


def regression_golden() -> dict:
    import torchmetrics.functional as TF
    import torchmetrics.regression as TR

    out: dict = {}
    g = torch.Generator().manual_seed(41)
    p1 = torch.rand(1000, generator=g) * 4 + 0.1
    t1 = torch.rand(1000, generator=g) * 4 + 0.1
    p2 = torch.randn(600, 5, generator=g) * 2 + 1
    t2 = torch.randn(600, 5, generator=g) * 2 + 1
    for k, v in (("p1", p1), ("t1", t1), ("p2", p2), ("t2", t2)):
        out[f"reg/{k}"] = v.numpy()
    out["reg/mse"] = TF.mean_squared_error(p1, t1).numpy()
    out["reg/rmse"] = TF.mean_squared_error(p1, t1, squared=False).numpy()
    out["reg/mse_multi"] = TF.mean_squared_error(p2, t2, num_outputs=5).numpy()
    out["reg/mae"] = TF.mean_absolute_error(p1, t1).numpy()
    out["reg/mape"] = TF.mean_absolute_percentage_error(p1, t1).numpy()
    out["reg/smape"] = TF.symmetric_mean_absolute_percentage_error(p1, t1).numpy()
    out["reg/wmape"] = TF.weighted_mean_absolute_percentage_error(p1, t1).numpy()
    out["reg/msle"] = TF.mean_squared_log_error(p1, t1).numpy()
    out["reg/logcosh"] = TF.log_cosh_error(p1, t1).numpy()
    out["reg/logcosh_multi"] = TF.log_cosh_error(p2, t2).numpy()
    out["reg/minkowski3"] = TF.minkowski_distance(p1, t1, 3).numpy()
    out["reg/minkowski1.5"] = TF.minkowski_distance(p2, t2, 1.5).numpy()
    for mo in ("raw_values", "uniform_average", "variance_weighted"):
        out[f"reg/r2_{mo}"] = TF.r2_score(p2, t2, multioutput=mo).numpy()
        out[f"reg/ev_{mo}"] = TF.explained_variance(p2, t2, multioutput=mo).numpy()
    out["reg/r2_1d"] = TF.r2_score(p1, t1).numpy()
    out["reg/r2_adj"] = TF.r2_score(p2, t2, adjusted=3).numpy()
    out["reg/rse"] = TF.relative_squared_error(p2, t2).numpy()
    out["reg/rrse"] = TF.relative_squared_error(p2, t2, squared=False).numpy()
    out["reg/ev_1d"] = TF.explained_variance(p1, t1).numpy()
    # modular, 4 updates
    for name, m, (pp, tt) in (("MeanSquaredError", TR.MeanSquaredError(), (p1, t1)),
                              ("MeanAbsoluteError", TR.MeanAbsoluteError(), (p1, t1)),
                              ("R2Score", TR.R2Score(), (p1, t1)),
                              ("ExplainedVariance", TR.ExplainedVariance(multioutput="raw_values"), (p2, t2)),
                              ("MeanSquaredErrorMulti", TR.MeanSquaredError(num_outputs=5), (p2, t2))):
        for a, b in zip(pp.chunk(4), tt.chunk(4)):
            m.update(a, b)
        out[f"reg/class/{name}"] = m.compute().numpy()
    return out


def coco_format_golden() -> dict:
    """The reference's COCO-json formatter (detection/mean_ap.py:867-958, bbox) applied to per-image states.  The
    reference class cannot be instantiated here (no pycocotools), but `_get_coco_format` only needs `iou_type` and
    `_get_classes` from `self`, so it is called unbound on a stand-in object.  Inputs are stored already converted to
    xywh (what `update` caches); outputs are the two dataset dicts as json strings."""
    import json
    import types

    from torchmetrics.detection.mean_ap import MeanAveragePrecision as Ref

    g = torch.Generator().manual_seed(20240921)
    images = []
    for i in range(8):
        nd, ng = int(torch.randint(0, 6, (1,), generator=g)), int(torch.randint(0, 5, (1,), generator=g))
        img = {
            "d_box": torch.cat([torch.rand(nd, 2, generator=g) * 100, torch.rand(nd, 2, generator=g) * 50 + 1], 1),
            "d_score": torch.rand(nd, generator=g),
            "d_label": torch.randint(0, 5, (nd,), generator=g),
            "g_box": torch.cat([torch.rand(ng, 2, generator=g) * 100, torch.rand(ng, 2, generator=g) * 50 + 1], 1),
            "g_label": torch.randint(0, 5, (ng,), generator=g),
        }
        if i % 2:
            img["g_crowd"] = torch.randint(0, 2, (ng,), generator=g)
            img["g_area"] = torch.rand(ng, generator=g) * 100 * (torch.rand(ng, generator=g) > 0.3)
        else:  # what `update` stores when the user gives neither: zeros_like(labels)
            img["g_crowd"] = torch.zeros_like(img["g_label"])
            img["g_area"] = torch.zeros_like(img["g_label"])
        images.append(img)
    labels = torch.cat([im["d_label"] for im in images] + [im["g_label"] for im in images]).unique().tolist()
    fake = types.SimpleNamespace(iou_type=("bbox",), _get_classes=lambda: labels)
    target = Ref._get_coco_format(fake, labels=[im["g_label"] for im in images], boxes=[im["g_box"] for im in images],
                                  masks=None, crowds=[im["g_crowd"] for im in images], area=[im["g_area"] for im in images])
    preds = Ref._get_coco_format(fake, labels=[im["d_label"] for im in images], boxes=[im["d_box"] for im in images],
                                 masks=None, scores=[im["d_score"] for im in images])
    out = {"n_images": np.array(len(images)), "target_json": np.array(json.dumps(target)),
           "preds_json": np.array(json.dumps(preds["annotations"]))}
    for i, im in enumerate(images):
        for k, v in im.items():
            out[f"img{i}/{k}"] = np_of(v)
    return out


def fairness_golden() -> dict:
    """Group fairness (reference functional/classification/group_fairness.py): per-group counters from
    `_binary_groups_stat_scores`, the rates of `binary_groups_stat_rates`, and the DP / EO dictionaries of
    `binary_fairness` — keys included, they carry the arg-min / arg-max group."""
    import warnings

    from torchmetrics.classification.group_fairness import BinaryFairness
    from torchmetrics.functional.classification.group_fairness import (
        _binary_groups_stat_scores,
        binary_fairness,
        binary_groups_stat_rates,
    )

    g = torch.Generator().manual_seed(777)
    out = {}
    case = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for kind in ("probs", "logits", "labels"):
            for ign in (None, 0, -1):
                for num_groups, shape in ((2, (256,)), (5, (193,)), (3, (64, 7))):
                    n = shape[0]
                    if kind == "probs":
                        preds = torch.rand(shape, generator=g)
                    elif kind == "logits":
                        preds = torch.randn(shape, generator=g) * 3
                    else:
                        preds = torch.randint(0, 2, shape, generator=g)
                    target = torch.randint(0, 2, shape, generator=g)
                    if ign is not None:
                        target[torch.rand(shape, generator=g) < 0.1] = ign
                    groups = torch.randint(0, num_groups, (n,), generator=g)
                    groups[:num_groups] = torch.arange(num_groups)  # every group occurs in both halves of the batch
                    groups[n // 2: n // 2 + num_groups] = torch.arange(num_groups)
                    stats = _binary_groups_stat_scores(preds, target, groups, num_groups, 0.5, ign, True)
                    rates = binary_groups_stat_rates(preds, target, groups, num_groups, 0.5, ign)
                    fair = binary_fairness(preds, target, groups, "all", 0.5, ign)
                    metric = BinaryFairness(num_groups, ignore_index=ign)
                    metric.update(preds[: n // 2], target[: n // 2], groups[: n // 2])
                    metric.update(preds[n // 2:], target[n // 2:], groups[n // 2:])
                    fair_two_updates = metric.compute()
                    key = f"case{case}"
                    out[f"{key}/preds"] = np_of(preds)
                    out[f"{key}/target"] = np_of(target)
                    out[f"{key}/groups"] = np_of(groups)
                    out[f"{key}/meta"] = np.array([num_groups, -999 if ign is None else ign])
                    out[f"{key}/counts"] = np.stack([np.array([int(v) for v in st]) for st in stats])
                    out[f"{key}/rates"] = np.stack([np_of(rates[f"group_{i}"]) for i in range(num_groups)])
                    out[f"{key}/fair_keys"] = np.array(",".join(fair.keys()))
                    out[f"{key}/fair_values"] = np.array([float(v) for v in fair.values()], dtype=np.float32)
                    out[f"{key}/fair2_keys"] = np.array(",".join(fair_two_updates.keys()))
                    out[f"{key}/fair2_values"] = np.array([float(v) for v in fair_two_updates.values()], dtype=np.float32)
                    case += 1
    out["n_cases"] = np.array(case)
    return out


def tweedie_golden() -> dict:
    """Tweedie deviance (reference functional/regression/tweedie_deviance.py + regression/tweedie_deviance.py): the mean
    deviance, the `(sum, count)` of `_tweedie_deviance_score_update`, the class after two updates, and per-element
    deviances of a small batch (for the host harness of the kernel's term function)."""
    from torchmetrics.functional.regression.tweedie_deviance import _tweedie_deviance_score_update, tweedie_deviance_score
    from torchmetrics.regression.tweedie_deviance import TweedieDevianceScore

    g = torch.Generator().manual_seed(4242)
    out = {}
    powers = [-1.5, 0.0, 1.0, 1.5, 2.0, 3.0]
    out["powers"] = np.array(powers)
    case = 0
    for dtype in (torch.float32, torch.float64):
        for shape in ((257,), (64, 3)):
            preds = (torch.rand(shape, generator=g, dtype=torch.float64) * 4 + 0.05).to(dtype)
            targets = (torch.rand(shape, generator=g, dtype=torch.float64) * 4 + 0.05).to(dtype)
            targets_with_zeros = targets.clone()
            targets_with_zeros.view(-1)[::7] = 0  # legal for powers 1 and (1, 2) only
            for power in powers:
                tg = targets_with_zeros if power in (1.0, 1.5) else targets
                key = f"case{case}"
                out[f"{key}/preds"], out[f"{key}/targets"] = preds.numpy(), tg.numpy()
                out[f"{key}/power"] = np.array(power)
                out[f"{key}/value"] = tweedie_deviance_score(preds, tg, power).numpy()
                total, count = _tweedie_deviance_score_update(preds, tg, power)
                out[f"{key}/sum"], out[f"{key}/count"] = total.numpy(), count.numpy()
                metric = TweedieDevianceScore(power=power)
                half = shape[0] // 2
                metric.update(preds[:half], tg[:half])
                metric.update(preds[half:], tg[half:])
                out[f"{key}/class_value"] = metric.compute().numpy()
                case += 1
    out["n_cases"] = np.array(case)
    # per-element deviances (float64) for the term-function harness: x = preds, y = targets
    x = torch.tensor([0.3, 1.0, 2.5, 4.0, 0.75, 3.25], dtype=torch.float64)
    y = torch.tensor([2.0, 1.0, 0.5, 0.0, 3.5, 3.25], dtype=torch.float64)
    out["elem/preds"], out["elem/targets"] = x.numpy(), y.numpy()
    for power in powers:
        tg = y if power in (1.0, 1.5) else y.clamp(min=0.125)
        per = torch.stack([_tweedie_deviance_score_update(x[i:i + 1], tg[i:i + 1], power)[0] for i in range(len(x))])
        out[f"elem/p{power}/targets"], out[f"elem/p{power}/deviance"] = tg.numpy(), per.numpy()
    return out


def csi_golden() -> dict:
    """Critical success index (reference functional/regression/csi.py + regression/csi.py): counts, scores, kept sequence
    dimensions, the class after two updates."""
    from torchmetrics.functional.regression.csi import _critical_success_index_update, critical_success_index
    from torchmetrics.regression.csi import CriticalSuccessIndex

    g = torch.Generator().manual_seed(31337)
    out = {}
    case = 0
    for shape in ((513,), (40, 9), (6, 5, 4, 3), (2, 3, 8, 8, 2)):
        for dtype in (torch.float32, torch.float64, torch.float16):
            preds = torch.rand(shape, generator=g).to(dtype)
            target = torch.rand(shape, generator=g).to(dtype)
            target.view(-1)[::5] = preds.view(-1)[::5]  # ties with each other and (below) with the threshold
            preds.view(-1)[::11] = 0.5
            for threshold in (0.5, 0.25):
                for keep in (None, *range(len(shape))):
                    key = f"case{case}"
                    out[f"{key}/preds"], out[f"{key}/target"] = np_of(preds), np_of(target)
                    out[f"{key}/meta"] = np.array([threshold, -1 if keep is None else keep, {torch.float32: 0, torch.float64: 1, torch.float16: 2}[dtype]])
                    hits, misses, fa = _critical_success_index_update(preds, target, threshold, keep)
                    out[f"{key}/counts"] = np.stack([hits.numpy(), misses.numpy(), fa.numpy()])
                    out[f"{key}/value"] = critical_success_index(preds, target, threshold, keep).numpy()
                    metric = CriticalSuccessIndex(threshold, keep_sequence_dim=keep)
                    metric.update(preds, target)
                    metric.update(target, preds)
                    out[f"{key}/class_value"] = metric.compute().numpy()
                    case += 1
    out["n_cases"] = np.array(case)
    return out


def kld_golden() -> dict:
    """KL divergence (reference functional/regression/kl_divergence.py + regression/kl_divergence.py): per-row measures, the
    three reductions, probabilities (normalised by the reference) and log-probabilities, zeros in p (xlogy convention), the
    class after two updates."""
    from torchmetrics.functional.regression.kl_divergence import _kld_update, kl_divergence
    from torchmetrics.regression.kl_divergence import KLDivergence

    g = torch.Generator().manual_seed(4711)
    out = {}
    case = 0
    for n, d in ((1, 3), (7, 1), (33, 10), (24, 257), (5, 1000), (120, 40)):
        for dtype in (torch.float32, torch.float64):
            for log_prob in (False, True):
                p = torch.rand(n, d, generator=g, dtype=torch.float64) * 3
                q = torch.rand(n, d, generator=g, dtype=torch.float64) * 3 + 1e-3
                if log_prob:
                    p, q = torch.log_softmax(p, 1), torch.log_softmax(q, 1)
                else:
                    p.view(-1)[::7] = 0.0  # xlogy(0, .) = 0
                p, q = p.to(dtype), q.to(dtype)
                key = f"case{case}"
                out[f"{key}/p"], out[f"{key}/q"] = np_of(p), np_of(q)
                out[f"{key}/meta"] = np.array([int(log_prob), {torch.float32: 0, torch.float64: 1}[dtype]])
                out[f"{key}/measures"] = _kld_update(p, q, log_prob)[0].numpy()
                for red in ("mean", "sum", "none"):
                    out[f"{key}/{red}"] = kl_divergence(p, q, log_prob, red).numpy()
                    metric = KLDivergence(log_prob=log_prob, reduction=red)
                    metric.update(p, q)
                    metric.update(q.abs() if not log_prob else q, p.abs() + 1e-3 if not log_prob else p)
                    out[f"{key}/class_{red}"] = metric.compute().numpy()
                case += 1
    p, q = torch.tensor([[0.36, 0.48, 0.16]]), torch.tensor([[1 / 3, 1 / 3, 1 / 3]])
    out["doc/value"] = kl_divergence(p, q).numpy()  # docstring: 0.0853
    out["n_cases"] = np.array(case)
    return out


def curves64_golden() -> dict:
    """float64 scores through the exact curve functionals (the reference sorts them as doubles) and `_binary_clf_curve`
    with `sample_weights` (functional/classification/precision_recall_curve.py:30-82).  Scores are built so that pairs differ
    ONLY below float32 resolution: a float32 sort would merge their thresholds."""
    import torchmetrics.functional.classification as RF
    from torchmetrics.functional.classification.precision_recall_curve import _binary_clf_curve

    g = torch.Generator().manual_seed(6464)
    out = {}

    def flat(res):
        parts = []
        for part in (res if isinstance(res, (tuple, list)) else [res]):
            parts.extend(part if isinstance(part, (tuple, list)) else [part])
        return parts

    def put(key, preds, target, res, **meta):
        out[f"{key}/preds"], out[f"{key}/target"] = preds.numpy(), target.numpy()
        parts = flat(res)
        out[f"{key}/n_out"] = np.array(len(parts))
        for i, t in enumerate(parts):
            out[f"{key}/out{i}"] = t.numpy()
        for k, v in meta.items():
            out[f"{key}/{k}"] = np.array(v)

    case = 0
    # binary, probabilities and logits, with sub-float32 perturbations and exact ties
    for n, logits in ((257, False), (4099, True), (20000, False)):
        base = torch.rand(n, generator=g, dtype=torch.float64)
        base[::3] = base[::3].float().double()                      # exactly representable in float32
        base[1::3] = base[::3][: base[1::3].numel()] + 1e-12        # differs from its neighbour below float32 resolution
        base[5::7] = base[4::7][: base[5::7].numel()]               # exact ties
        preds = (base * 8 - 4) if logits else base.clamp(0, 1)
        target = torch.randint(0, 2, (n,), generator=g)
        for fn in ("binary_roc", "binary_precision_recall_curve", "binary_auroc", "binary_average_precision"):
            put(f"case{case}", preds, target, getattr(RF, fn)(preds, target, thresholds=None), fn=fn, num_classes=0)
            case += 1
    # multiclass one-vs-rest
    for n, c in ((300, 5), (1500, 37)):
        lg = torch.randn(n, c, generator=g, dtype=torch.float64)
        # odd rows repeat the even ones with ONE logit moved by 1e-10: every probability of the pair then differs by ~1e-11
        # relative — far above float64 rounding of the softmax (so the order is well defined), far below float32 resolution
        # (shifting the whole row would leave the softmax unchanged up to rounding noise)
        lg[1::2] = lg[::2][: lg[1::2].shape[0]]
        lg[1::2, 0] += 1e-10
        target = torch.randint(0, c, (n,), generator=g)
        for fn in ("multiclass_roc", "multiclass_precision_recall_curve", "multiclass_auroc", "multiclass_average_precision"):
            kw = dict(num_classes=c, thresholds=None)
            if "auroc" in fn or "average_precision" in fn:
                kw["average"] = None
            put(f"case{case}", lg, target, getattr(RF, fn)(lg, target, **kw), fn=fn, num_classes=c)
            case += 1
    # multilabel
    for n, l in ((400, 4),):
        pr = torch.rand(n, l, generator=g, dtype=torch.float64)
        pr[1::2] = (pr[::2][: pr[1::2].shape[0]] + 1e-13).clamp(0, 1)
        target = torch.randint(0, 2, (n, l), generator=g)
        for fn in ("multilabel_roc", "multilabel_precision_recall_curve", "multilabel_auroc", "multilabel_average_precision"):
            kw = dict(num_labels=l, thresholds=None)
            if "auroc" in fn or "average_precision" in fn:
                kw["average"] = None
            put(f"case{case}", pr, target, getattr(RF, fn)(pr, target, **kw), fn=fn, num_classes=l)
            case += 1
    out["n_cases"] = np.array(case)
    # sample weights through the private helper (no public functional forwards them)
    wcase = 0
    for n, pdt, wdt in ((64, torch.float32, torch.float32), (1000, torch.float32, torch.float64), (5000, torch.float64, torch.float32),
                        (333, torch.float16, torch.float32)):
        preds = (torch.rand(n, generator=g, dtype=torch.float64) * 16).round() / 16 if n < 2000 else torch.rand(n, generator=g, dtype=torch.float64)
        preds = preds.to(pdt)
        target = torch.randint(0, 3, (n,), generator=g)
        weights = (torch.rand(n, generator=g, dtype=torch.float64) * 3).to(wdt)
        for pos in (1, 2):
            fps, tps, thr = _binary_clf_curve(preds, target, sample_weights=weights, pos_label=pos)
            key = f"w{wcase}"
            out[f"{key}/preds"] = preds.float().numpy() if pdt == torch.float16 else preds.numpy()
            out[f"{key}/half"] = np.array(pdt == torch.float16)
            out[f"{key}/target"], out[f"{key}/weights"], out[f"{key}/pos"] = target.numpy(), weights.numpy(), np.array(pos)
            out[f"{key}/fps"], out[f"{key}/tps"] = fps.numpy(), tps.numpy()
            out[f"{key}/thr"] = thr.float().numpy() if pdt == torch.float16 else thr.numpy()
            wcase += 1
    # weights given as a python list (reference :45-46 converts to float32)
    preds = torch.tensor([0.1, 0.4, 0.35, 0.8, 0.4])
    target = torch.tensor([0, 0, 1, 1, 1])
    fps, tps, thr = _binary_clf_curve(preds, target, sample_weights=[1.0, 2.0, 0.5, 1.5, 1.0])
    out["wlist/fps"], out["wlist/tps"], out["wlist/thr"] = fps.numpy(), tps.numpy(), thr.numpy()
    out["n_weighted"] = np.array(wcase)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["classification"]
    if "classification" in which:
        data = classification_golden()
        path = os.path.join(HERE, "classification.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "kld" in which:
        data = kld_golden()
        path = os.path.join(HERE, "kld.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "curves64" in which:
        data = curves64_golden()
        path = os.path.join(HERE, "curves64.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "binned" in which:
        data = binned_golden()
        path = os.path.join(HERE, "binned.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "regression" in which:
        data = regression_golden()
        path = os.path.join(HERE, "regression.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "map" in which:
        data = map_golden()
        path = os.path.join(HERE, "detection.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "multilabel" in which:
        data = multilabel_golden()
        path = os.path.join(HERE, "multilabel.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "consumers" in which:
        data = consumers_golden()
        path = os.path.join(HERE, "consumers.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "fuzz" in which:
        data = fuzz_golden()
        path = os.path.join(HERE, "fuzz.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "atfixed" in which:
        data = atfixed_golden()
        path = os.path.join(HERE, "atfixed.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "logauc" in which:
        data = logauc_golden()
        path = os.path.join(HERE, "logauc.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "curves" in which:
        data = curves_golden()
        path = os.path.join(HERE, "curves.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "coco_format" in which:
        data = coco_format_golden()
        path = os.path.join(HERE, "coco_format.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "fairness" in which:
        data = fairness_golden()
        path = os.path.join(HERE, "fairness.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "tweedie" in which:
        data = tweedie_golden()
        path = os.path.join(HERE, "tweedie.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "fuzz2" in which:  # a second, larger draw (other seed): replayed by the CPU host twin and by the LAST GPU test file
        data = fuzz_golden(seed=77077, mult=3)
        path = os.path.join(HERE, "fuzz2.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
    if "csi" in which:
        data = csi_golden()
        path = os.path.join(HERE, "csi.npz")
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(data), "arrays")
