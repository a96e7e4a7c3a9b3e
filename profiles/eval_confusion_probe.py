"""Evaluation counters: ConfusionMeter.update (one kernel) vs the reference's interp -> softmax -> argmax -> .cpu() ->
fast_hist per image (adaptation_model.py:143-160), Cityscapes shapes.  Run on the GPU box:
python profiles/eval_confusion_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from onda_b200 import ConfusionMeter

dev = torch.device("cuda:0")
C = 19
for (B, h, w, H, W) in [(1, 65, 129, 512, 1024), (1, 129, 257, 1024, 2048), (4, 129, 257, 1024, 2048)]:
    pred = torch.randn(B, C, h, w, device=dev) * 3
    labels_host = torch.randint(0, C, (B, H, W))
    labels_host[torch.rand(B, H, W) < 0.1] = 255
    labels_dev = labels_host.to(dev)
    meter = ConfusionMeter(C, dev)
    for _ in range(3):
        meter.update(pred, labels_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        meter.update(pred, labels_dev)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        meter.update(pred, labels_host)          # labels from the loader (host): includes the H2D copy and int64 cast
    torch.cuda.synchronize()
    t_host_labels = (time.perf_counter() - t0) / 5 * 1e3
    interp = torch.nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)

    def reference():
        counters = 0
        p = interp(pred).softmax(axis=1)
        for item_pred, label in zip(p, labels_host):
            a = label.numpy().flatten()
            b = item_pred.permute(1, 2, 0).argmax(dim=2).cpu().numpy().flatten()
            k = (a >= 0) & (a < C)
            counters = counters + np.bincount(C * a[k].astype(int) + b[k], minlength=C ** 2).reshape(C, C)
        return counters
    reference()
    t0 = time.perf_counter()
    for _ in range(3):
        ref = reference()
    t_ref = (time.perf_counter() - t0) / 3 * 1e3
    meter.reset()
    meter.update(pred, labels_dev)
    same = np.array_equal(meter.hist(), ref)
    px = B * H * W
    print(f"B={B} {h}x{w} -> {H}x{W}: kernel {t_dev * 1e3:7.1f} us ({px / t_dev / 1e6:6.2f} Gpx/s, {px * 9 / t_dev / 1e6:6.0f} GB/s of label+prediction traffic), "
          f"with host labels {t_host_labels:6.2f} ms, reference path {t_ref:7.2f} ms -> {t_ref / t_host_labels:5.1f}x; hist equal: {same}")
