"""MeanAveragePrecision(iou_type="segm") at a COCO-like shape: `n_img` images of 480 x 640 with 100 predicted and 20 ground-truth
instance masks each, 80 classes.  Wall clock of the public API (update = bit packing on the device, compute = pair
intersections + matching + accumulation), device-resident inputs.  The reference needs pycocotools for this path (absent in
this project's containers), so there is no reference arm; prints ONE JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from metrics_b200.detection import MeanAveragePrecision  # noqa: E402

dev = torch.device("cuda", 0)
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 200
H, W, ND, NG, C = 480, 640, 100, 20, 80
g = torch.Generator(device=dev).manual_seed(0)


def boxes_to_masks(n):
    """n random rectangles as [n, H, W] bool (built on the device)."""
    y0 = torch.randint(0, H - 40, (n, 1, 1), generator=g, device=dev)
    x0 = torch.randint(0, W - 40, (n, 1, 1), generator=g, device=dev)
    hh = torch.randint(20, 200, (n, 1, 1), generator=g, device=dev)
    ww = torch.randint(20, 200, (n, 1, 1), generator=g, device=dev)
    yy = torch.arange(H, device=dev).view(1, H, 1)
    xx = torch.arange(W, device=dev).view(1, 1, W)
    return (yy >= y0) & (yy < y0 + hh) & (xx >= x0) & (xx < x0 + ww)


preds, target = [], []
for _ in range(n_img):
    preds.append(dict(masks=boxes_to_masks(ND), scores=torch.rand(ND, generator=g, device=dev),
                      labels=torch.randint(0, C, (ND,), generator=g, device=dev)))
    target.append(dict(masks=boxes_to_masks(NG), labels=torch.randint(0, C, (NG,), generator=g, device=dev)))
torch.cuda.synchronize()
m = MeanAveragePrecision(iou_type="segm").to(dev)
m.warn_on_many_detections = False


def updates():
    m.reset()
    for i in range(0, n_img, 20):
        m.update(preds[i:i + 20], target[i:i + 20])


updates()
torch.cuda.synchronize()
upd = []
for _ in range(3):
    t0 = time.perf_counter()
    updates()
    torch.cuda.synchronize()
    upd.append(time.perf_counter() - t0)
cmp_ = []
for rep in range(4):
    m._computed = None
    torch.cuda.synchronize()
    prof = None
    if rep == 0 and os.environ.get("MB200_PROFILE_FIRST"):
        import cProfile
        import pstats

        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    r = m.compute()
    torch.cuda.synchronize()
    cmp_.append(time.perf_counter() - t0)
    if prof is not None:
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(8)
first_compute = cmp_.pop(0)  # pays one-off costs (first use of a few torch kernels, allocator growth for the flat mask buffers)
mask_bytes = n_img * (ND + NG) * H * W
print(json.dumps({
    "workload": f"MeanAveragePrecision(iou_type='segm'), {n_img} images of {H}x{W}, {ND} predicted + {NG} ground-truth masks each, {C} classes",
    "update_phase_s_wall": min(upd), "update_phase_s_wall_all": upd, "compute_s_wall": min(cmp_), "compute_s_wall_all": cmp_, "first_compute_s_wall": first_compute,
    "mask_bytes_in": mask_bytes, "update_gb_per_s": mask_bytes / min(upd) / 1e9,
    "masks_per_s_end_to_end": n_img * (ND + NG) / (min(upd) + min(cmp_)), "images_per_s_end_to_end": n_img / (min(upd) + min(cmp_)),
    "map": float(r["map"]), "state_bytes": int(sum(t.numel() for t in m.detection_mask + m.groundtruth_mask) * 4),
}))
