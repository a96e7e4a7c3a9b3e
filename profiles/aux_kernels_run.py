"""Launches every small kernel of the path once or twice on bench-sized inputs (for `ncu --set full`):
table_kernel (EMA + distance table), reduce_partials_kernel, prior_mix_kernel (K4), step_log_kernel, target_loss_kernel
(f1), weight_ema_kernel (f2), confusion_kernel (f3), and split_finish_kernel (D = 2048).
    ncu --set full --clock-control none -k regex:'table_kernel|reduce_partials|prior_mix|step_log|target_loss|weight_ema|confusion|split_finish' \
        -o gpurun_out/r2_aux python profiles/aux_kernels_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onda_b200 import prototype_handler, WeightEma, ConfusionMeter, target_losses

dev = torch.device("cuda:0")
bench.B_PER_GPU = 32
protos, sq, cnt, sets = bench.gpu_inputs(torch, dev, 256, 1, 1234)
h = prototype_handler(**bench.PARAMS)
h.prototypes, h.squared_mean, h.counter = protos, sq, cnt
feat, prior, out = sets[0]
g = torch.Generator(device=dev).manual_seed(5)
static = torch.randn_like(out)
dyn = torch.randn_like(out)
for _ in range(2):
    pr, conf, pc = h.prior_mix([out, static, dyn], [0.0, 1.0, 1.0], scale01=0.4)          # prior_mix_kernel
    labels, soft = h.pseudo_labels_fused(feat, prior, out)                                  # fused + reduce_partials
    h.ma(feat, out)                                                                         # table_kernel
    h.step_log_stats(labels, static)                                                        # step_log_kernel
    student = static.clone().requires_grad_(True)
    target_losses(student, labels, n_valid=h.last_pixel_count())["Total target loss"].backward()   # target_loss_kernel
shapes = [(64, 3, 7, 7)] + [(256, 256, 3, 3)] * 60 + [(512,)] * 100 + [(2048, 512, 1, 1)] * 4
net = lambda: torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, generator=g, device=dev)) for s in shapes])
plan = WeightEma(net(), net())
meter = ConfusionMeter(19)
pred = torch.randn(1, 19, 129, 257, generator=g, device=dev)
lab = torch.randint(0, 19, (1, 1024, 2048), generator=g, device=dev)
for _ in range(2):
    plan.update(0.999)                                                                      # weight_ema_kernel
    meter.update(pred, lab)                                                                 # confusion_kernel
del feat, prior, out, sets
bench.B_PER_GPU = 4
protos, sq, cnt, sets = bench.gpu_inputs(torch, dev, 2048, 1, 99)
h2 = prototype_handler(**bench.PARAMS)
h2.prototypes, h2.squared_mean, h2.counter = protos, sq, cnt
for _ in range(2):
    h2.pseudo_labels_fused(*sets[0])                                                        # split_finish_kernel
    h2.ma(sets[0][0], sets[0][2])
torch.cuda.synchronize()
print("done")
