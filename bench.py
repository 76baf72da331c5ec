#!/usr/bin/env python
"""Benchmark of the output-space adaptation step (train_adapt.py:126-181 of the reference):
DeepLabV3+/MobileNetV2 + FCDiscriminator, source + target 512x1024 crops, batch 8 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (BASELINE config 2 / 3)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU
    python bench.py --workload feature ...                    # BASELINE config 4: train.py feature-adaptation step
    python bench.py --workload val ...                        # BASELINE config 5: val_adapt.py, 500 images 1024x2048
(each workload with its own --impl reference arm; the default line is config 2, as the driver expects).

Prints ONE JSON line (rank 0).  metric = train img-pairs/s (one pair = one source + one target
image through one full step), whole job.  `value` is measured with the step's inputs resident in
HBM; `e2e` runs the same step through the public module API from pinned HOST tensors, with the
host->device copies and a device->host read of the losses inside the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "synthetic-to-real-semantic-segmentation_b200"
METRIC = "train img-pairs/sec (DeepLabV3+ MNv2 512x1024 output-space adaptation step)"
UNIT = "img-pairs/s"
METRICS = {"adapt": (METRIC, UNIT),
           "feature": ("train img-pairs/sec (DeepLabV3+ MNv2 512x1024 feature-adaptation step, train.py)", "img-pairs/s"),
           "val": ("val img/s (DeepLabV3+ MNv2 1024x2048 batch 1, Evaluator confusion matrix, val_adapt.py)", "img/s")}


def sub(name=""):
    return importlib.import_module(PKG + ("." + name if name else ""))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 10; val: images, default 500)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="adapt", choices=["adapt", "feature", "val"],
                    help="adapt: train_adapt.py step (BASELINE config 2/3, the default line); feature: train.py step "
                         "(config 4); val: val_adapt.py over distinct 1024x2048 images (config 5)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="image pairs per GPU")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--cpu-batch", type=int, default=2, help="pairs per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 500 if a.workload == "val" else 10
    return a


# ----------------------------------------------------------------------------- CPU arm (oracle)
def synth(seed, n, h, w, pin=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, 3, h, w, generator=g)
    tgt = torch.randn(n, 3, h, w, generator=g)
    lab = torch.randint(0, 19, (n, h, w), generator=g).float()
    lab[torch.rand(n, h, w, generator=g) < 0.05] = 255        # ~5 % ignore
    if pin:
        src, tgt, lab = src.pin_memory(), tgt.pin_memory(), lab.pin_memory()
    return src, lab, tgt


def cpu_adapt_steps(batch, h, w, steps, warmup):
    """The reference algorithm (oracle port of train_adapt.py:137-181, fp32, stock torch CPU kernels)
    on the host cores.  Returns (pairs/s, seconds per step, threads)."""
    import torch
    from oracle import ref_port as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # nothing of this repo's package on this path: weights, forward, backward and optimizers are the oracle's + torch's
    g_sd, d_sd = O.init_deeplab(seed=1), O.init_discriminator(seed=2)
    for sd in (g_sd, d_sd):
        for v in O.leaf_params(sd).values():
            v.requires_grad_(True)
    one, ten = O.split_lr_groups(list(O.leaf_params(g_sd).keys()))
    opt = torch.optim.SGD([{'params': [g_sd[k] for k in one], 'lr': 5e-4},
                           {'params': [g_sd[k] for k in ten], 'lr': 5e-3}], momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(list(O.leaf_params(d_sd).values()), lr=1e-4, betas=(0.9, 0.99))
    src, lab, tgt = synth(1000, batch, h, w)
    cfg = O.BNCfg(True)
    for _ in range(warmup):
        O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.adapt_step(g_sd, d_sd, opt, opt_d, src, lab, tgt, cfg)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt, threads


def _oracle_dc_state(O, torch, seed=4):
    """DomainClassifer state dict (modeling/domian.py:15-23,35-44: kaiming-normal convs, default-initialised bias)
    from plain tensors, like O.init_deeplab."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    O._conv_w(sd, 'DC_adnn1.0', 1024, 256, 1, gen)
    O._bn(sd, 'DC_adnn1.1', 1024)
    O._conv_w(sd, 'DC_adnn2.0', 1024, 1024, 3, gen)
    O._bn(sd, 'DC_adnn2.1', 1024)
    O._conv_w(sd, 'DC_adnn3', 2, 1024, 3, gen, bias=True)
    return sd


def cpu_feature_steps(batch, h, w, steps, warmup):
    """Oracle port of train.py:173-204 (the script's default Adam optimizers) on the host cores."""
    import torch
    from oracle import ref_port as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g_sd = O.init_deeplab(seed=1)
    f_sd = {k[len('backbone.'):]: v for k, v in g_sd.items() if k.startswith('backbone.')}
    a_sd = {k[len('aspp.'):]: v for k, v in g_sd.items() if k.startswith('aspp.')}
    y_sd = {k[len('decoder.'):]: v for k, v in g_sd.items() if k.startswith('decoder.')}
    dc_sd = _oracle_dc_state(O, torch)
    for sd in (f_sd, a_sd, y_sd, dc_sd):
        for v in O.leaf_params(sd).values():
            v.requires_grad_(True)
    fp = list(O.leaf_params(f_sd).values()) + list(O.leaf_params(a_sd).values())
    opts = (torch.optim.Adam(fp + list(O.leaf_params(y_sd).values()), lr=5e-4),
            torch.optim.Adam(list(O.leaf_params(dc_sd).values()), lr=5e-4), torch.optim.Adam(fp, lr=5e-4))
    src, lab, tgt = synth(1000, batch, h, w)
    cfg = O.BNCfg(True)
    for _ in range(warmup):
        O.feature_step(f_sd, a_sd, y_sd, dc_sd, opts, src, lab, tgt, cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.feature_step(f_sd, a_sd, y_sd, dc_sd, opts, src, lab, tgt, cfg)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt, threads


def val_image(torch, k, h, w, device=None, generator=None):
    """Image k of the synthetic validation set: N(0,1) pixels, labels uniform over {0..18} with ~5 % = 255."""
    g = generator if generator is not None else torch.Generator(device=device or "cpu")
    g.manual_seed(5000 + k)
    img = torch.randn(1, 3, h, w, generator=g, device=device)
    lab = torch.randint(0, 19, (1, h, w), generator=g, device=device).float()
    lab[torch.rand(1, h, w, generator=g, device=device) < 0.05] = 255
    return img, lab


def cpu_val_images(n, h, w):
    """Oracle port of val_adapt.py:122-135 on the host cores: eval forward, np.argmax on the host copy of the logits,
    Evaluator.add_batch (numpy bincount).  Returns (img/s, seconds per image, threads)."""
    import numpy as np
    import torch
    from oracle import ref_port as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = O.init_deeplab(seed=1)
    cm = np.zeros((19, 19), dtype=np.float64)
    cfg = O.BNCfg(False)

    def one(k):
        img, lab = val_image(torch, k, h, w)
        with torch.no_grad():
            out = O.deeplab_forward(sd, img, cfg, 16)
        pred = np.argmax(out.numpy(), axis=1)
        return O.confusion_matrix(lab.numpy(), pred, 19)

    one(0)
    t0 = time.perf_counter()
    for k in range(n):
        cm += one(k)
    dt = (time.perf_counter() - t0) / max(1, n)
    return 1.0 / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric, unit = METRICS[args.workload]
    if args.workload == "val":
        n = max(1, min(args.steps, 4))
        val, dt, threads = cpu_val_images(n, 2 * args.height, 2 * args.width)
        steps, warmup = n, 1
        sample = "%d image(s) (after 1 warm-up) of the oracle port of val_adapt.py:122-135 at 1x3x%dx%d, fp32, %d torch CPU threads" % (
            n, 2 * args.height, 2 * args.width, threads)
        work = "val_adapt.py:122-135: eval forward + host argmax + Evaluator.add_batch, 1x3x%dx%d" % (2 * args.height, 2 * args.width)
        cfgx = {"images": n}
    else:
        steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        fn = cpu_adapt_steps if args.workload == "adapt" else cpu_feature_steps
        val, dt, threads = fn(args.cpu_batch, args.height, args.width, steps, warmup)
        where = "train_adapt.py:137-181" if args.workload == "adapt" else "train.py:173-204"
        sample = "%d step(s) of the oracle port of %s at batch %d, %dx%d, fp32, %d torch CPU threads" % (
            steps, where, args.cpu_batch, args.height, args.width, threads)
        work = ("AdaptSegNet output-space adaptation step, DeepLabV3+ MNv2 OS16 + FCDiscriminator, src+tgt %dx%d crops"
                if args.workload == "adapt" else
                "FCN-in-the-wild feature adaptation step, MNv2 + ASPP + decoder + DomainClassifer, src+tgt %dx%d crops") % (
            args.height, args.width)
        cfgx = {"pairs_per_step": args.cpu_batch}
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict({"workload": work}, **cfgx),
        "cpu_baseline": {"value": val, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons during the timed region, every 100 ms.  Sampled in-process through NVML
    (pynvml): forking `nvidia-smi` from a thread stalls the benchmark's own Python thread (and its CUDA calls) for
    milliseconds, which the end-to-end loop -- one host synchronisation per step -- would pay for directly.
    `nvidia-smi` remains the fallback when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML enumerates all of them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        act = lambda bit: "Active" if (r & bit) else "Not Active"   # noqa: E731
        return [str(sm), str(mx), act(nv.nvmlClocksThrottleReasonHwSlowdown), act(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                act(nv.nvmlClocksThrottleReasonSwThermalSlowdown), act(nv.nvmlClocksThrottleReasonSwPowerCap)]

    def _run(self):
        while not self.stop:
            try:
                if self.nv is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                                   "--format=csv,noheader,nounits"], timeout=5).decode()
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1 if self.nv is not None else 0.2)

    def sample_once(self):
        try:
            if self.nv is not None:
                self.rows.append(self._sample_nvml())
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nv is not None else "nvidia-smi",
                "sampled": "every 100 ms during the resident timed region; one sample before and after the e2e region"}


# ----------------------------------------------------------------------------- rooflines
def _timed_launches(torch, fn, flush, reps=4, warm=2):
    """Average device time (s) of one launch of fn, CUDA events on the launching stream, L2 flushed before each."""
    times = []
    for i in range(warm + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            times.append(e0.elapsed_time(e1) * 1e-3)
    return sum(times) / len(times)


def conv_roofline(torch, peaks, name, batch, H, W, Cin, Cout, k, pad, traffic=None):
    """One dense tap-GEMM convolution forward (with BN statistics) timed alone: the dominant kernel of a workload."""
    eng = sub("engine")
    dev = torch.device("cuda", torch.cuda.current_device())
    cx = eng.Ctx(dev, True)
    x = eng.Act(torch.randn(batch, H, W, Cin, device=dev).to(torch.bfloat16))
    wgt = torch.nn.Parameter(torch.randn(Cout, Cin, k, k, device=dev) * 0.02)
    out = cx.new(batch, H, W, Cout)
    stats = cx.f64(2 * Cout)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t = _timed_launches(torch, lambda: eng.conv_fwd(cx, x, wgt, out, 1, pad, 1, stats=stats), flush)
    flops = 2.0 * batch * H * W * Cout * Cin * k * k
    peak = peaks.get("bf16_tflops", 1590.0)
    ach = flops / t / 1e12
    return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peaks.get("_source", "fallback"), "launch_ms": t * 1e3}


def dominant_kernel_roofline(torch, batch, h, w, peaks):
    """decoder.last_conv.0: 3x3 304->256 on [B, H/4, W/4] (45.9 GF/img forward), the
    largest single kernel of the step; timed alone with CUDA events, L2 flushed between launches."""
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this launch at batch 8, 512x1024 from the committed
    # ncu --set full capture (profiles/r2_conv_tc_decoder_ncu.txt: 161.2 MB read + 94.3 MB written -- part of the
    # 134 MB output is still in the 126 MB L2 when the kernel ends; algorithmic 159 + 134 = 293 MB)
    traffic = 255.5e6 if (batch, h, w) == (8, 512, 1024) else None
    return conv_roofline(torch, peaks, "tap-GEMM conv fwd 3x3 304->256 (decoder.last_conv.0)", batch, h // 4, w // 4, 304,
                         256, 3, 1, traffic)


# depthwise launches of one generator pass at OS=16 (modeling/backbone/mobilenet.py:78-109 of the reference):
# (block(s), input H and W as a fraction of the image, hidden channels, stride, dilation, BN+ReLU6 halo, launches per pass)
DW_FAMILY = (("features.1", 2, 32, 1, 1, False, 1), ("features.2", 2, 96, 2, 1, True, 1), ("features.3", 4, 144, 1, 1, True, 1),
             ("features.4", 4, 144, 2, 1, True, 1), ("features.5-6", 8, 192, 1, 1, True, 2), ("features.7", 8, 192, 2, 1, True, 1),
             ("features.8-11", 16, 384, 1, 1, True, 4), ("features.12-14", 16, 576, 1, 1, True, 3),
             ("features.15-16", 16, 960, 1, 1, True, 2), ("features.17", 16, 960, 1, 2, True, 1))


def depthwise_family_roofline(torch, batch, h, w, peaks):
    """The WHOLE depthwise family of the step, not its best layer: every depthwise launch shape of a generator pass
    (17 per pass; forward = BN/ReLU6 prologue + 3x3 + statistics, backward = data + weight gradient + BN-backward sums
    in one kernel) timed alone with CUDA events and the L2 flushed; achieved = sum of algorithmic bytes / sum of
    launch times, weighted by the number of launches of each shape.  Algorithmic bytes: forward =
    input + output elements x 2 B; backward = dy + x read, g written."""
    import ctypes as C
    eng, L = sub("engine"), sub("_lib")
    dev = torch.device("cuda", torch.cuda.current_device())
    cx = eng.Ctx(dev, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
    tot = {"fwd": [0.0, 0.0], "bwd": [0.0, 0.0]}
    layers = []
    for name, div, Cc, s, d, halo, cnt in DW_FAMILY:
        H, W = h // div, w // div
        x = eng.Act((torch.randn(batch, H, W, Cc, device=dev) * 2).to(torch.bfloat16))
        ss = torch.cat([torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev)]).contiguous()
        mi = torch.cat([torch.randn(Cc, device=dev) * 0.1, torch.rand(Cc, device=dev) + 0.5]).contiguous()
        st = eng.BNState(ss, mi, 1.0, False)
        wgt = torch.randn(Cc, 1, 3, 3, device=dev) * 0.3
        Ho, Wo = eng.conv_out_hw(H, W, 3, 3, s, d, d)
        dy = eng.Act(torch.randn(batch, Ho, Wo, Cc, device=dev).to(torch.bfloat16))
        g = cx.new(batch, H, W, Cc)
        dwg = torch.zeros_like(wgt)

        def fwd():
            eng.dw_fwd(cx, x, st, L.ACT_RELU6, halo, wgt, s, d, d, cx.f64(2 * Cc))

        def bwd():
            L.call("s2r_dwconv3x3_bwd", dy.vp(), vp(wgt), x.vp(), vp(ss), vp(mi), L.ACT_RELU6, 1 if halo else 0, 1, g.vp(),
                   vp(cx.f64(2 * Cc)), vp(dwg), batch, H, W, Cc, s, d, d, cx.stream)

        tf, tb = _timed_launches(torch, fwd, flush, reps=3, warm=1), _timed_launches(torch, bwd, flush, reps=3, warm=1)
        bf_ = 2.0 * batch * Cc * (H * W + Ho * Wo)
        bb_ = 2.0 * batch * Cc * (Ho * Wo + 2 * H * W)
        tot["fwd"][0] += cnt * bf_
        tot["fwd"][1] += cnt * tf
        tot["bwd"][0] += cnt * bb_
        tot["bwd"][1] += cnt * tb
        layers.append({"layer": name, "launches_per_pass": cnt, "fwd_us": round(tf * 1e6, 1), "fwd_gbs": round(bf_ / tf / 1e9),
                       "bwd_us": round(tb * 1e6, 1), "bwd_gbs": round(bb_ / tb / 1e9)})
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {}
    for key, what in (("fwd", "forward (BN/ReLU6 prologue + depthwise 3x3 + statistics)"),
                      ("bwd", "backward (data + weight gradient + BN-backward sums, one kernel)")):
        ach = tot[key][0] / tot[key][1] / 1e9
        out[key] = {"kernel": "depthwise 3x3 family, %s: all 17 launches of a generator pass, sum of bytes / sum of time" % what,
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peaks.get("_source", "fallback"), "launch_ms": tot[key][1] * 1e3 / 17,
                    "sum_ms_per_pass": tot[key][1] * 1e3}
    out["fwd"]["per_layer"] = layers
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "_source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- B200 arms: shared plumbing
class Job(object):
    """torch / torch.distributed set-up of one rank and the timing helpers of the contract: W warm-up steps, K timed
    steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.L = sub("_lib")
        if not os.path.exists(self.L.LIB_PATH):
            import __graft_entry__ as ge
            if self.local == 0:
                ge.build()
            if self.world > 1:
                dist.barrier()
        assert self.L.lib().s2r_device_ok() == 1, "built for sm_100a only"
        self.graphs = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, n_steps, it0):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.L.launches
        e0.record()
        for k in range(n_steps):
            fn(it0 + k)
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()), self.L.launches - l0

    def finish(self, *holders):
        """Clean shutdown: drop the CUDA graphs (they hold captured NCCL work), close the peer exchange, destroy the
        process group.  A watchdog ends the process if a communicator teardown blocks (seen in round 1 with NCCL
        collectives inside a live graph) so that a hang can never outlive the printed result."""
        sys.stdout.flush()
        if self.world <= 1:
            return
        import gc
        threading.Timer(30.0, lambda: os._exit(0)).start()
        self.barrier()
        for h in holders:
            for name in ("_graph", "_static_out", "_lanes"):
                if hasattr(h, name):
                    setattr(h, name, None)
        gc.collect()
        self.torch.cuda.synchronize()
        try:
            self.L.lib().s2r_comm_destroy()
            self.dist.destroy_process_group()
        finally:
            sys.stdout.flush()
            os._exit(0)       # the watchdog thread must not keep the interpreter alive; everything is flushed


def pipeline_e2e(step, host, i, torch):
    """One end-to-end step: pinned host -> device staging (copy stream) -> static inputs -> replay; the copy of the NEXT
    step's inputs is started before this step's losses are read back, so it overlaps the replay."""
    if getattr(step, "_staged", None) is None:
        step.stage(*host)
    out = step.replay_staged(i=i, epoch=0)
    step.stage(*host)
    return out


# ----------------------------------------------------------------------------- multi-rank parity record (world > 1)
def multi_rank_parity(job, G, run_step, d_src, it):
    """Untimed, after the timed regions: evidence in the driver's own record that the N-rank step computes what the
    reference's DataParallel + SynchronizedBatchNorm computes on the gathered batch.
      sync_bn: one more training-mode forward of G; the raw (pre-BN) activations of the first block's BatchNorm
        (backbone.features.1.conv.1) and of the last one (decoder.last_conv.5) are gathered from all ranks; rank 0
        evaluates modeling/sync_batchnorm/batchnorm.py:113-125 of the reference on the GATHERED tensor in fp64 (mean =
        sum/n, unbiased running variance with the global n, momentum 0.1) and compares with the running statistics the
        product's exchange + finalize produced.  max |diff| / max |expected| per buffer.
      weights_bit_equal: after one more full step every rank holds the same parameters and buffers, bit for bit
        (int64 sum of the raw words and the abs-sum, compared across ranks).
      softmax_dim0 / ce: the stated deviation -- F.softmax(dim=0) runs over the rank-local batch of 8 where the
        reference's single-process DataParallel normalises over the gathered global batch (train_adapt.py:151); its
        size is measured on the gathered logits.  The cross-entropy mean IS the global-batch mean (valid-pixel counts
        all-reduced, functional.GLOBAL_BATCH_MEAN)."""
    torch, dist = job.torch, job.dist
    eng = sub("engine")
    world, rank, dev = job.world, job.rank, job.dev
    bns = {"first": G.backbone.features[1].conv[1], "last": G.decoder.last_conv[5]}
    before = {k: (bn.running_mean.detach().clone().double(), bn.running_var.detach().clone().double()) for k, bn in bns.items()}
    eng.BN_PROBE = {id(bn): [] for bn in bns.values()}
    try:
        with torch.no_grad():
            logits = G(d_src)
        probes = {k: eng.BN_PROBE[id(bn)][0] for k, bn in bns.items()}
    finally:
        eng.BN_PROBE = None
    rec = {}
    for k, bn in bns.items():
        z = probes[k]
        zt = z.t[..., z.off:z.off + z.C].contiguous()
        parts = [torch.empty_like(zt) for _ in range(world)] if rank == 0 else None
        dist.gather(zt, parts, dst=0)
        if rank == 0:
            s = torch.zeros(z.C, dtype=torch.float64, device=dev)
            ss = torch.zeros(z.C, dtype=torch.float64, device=dev)
            n = 0
            for p in parts:                     # fp64 sums rank by rank (a [8 x 256 x 512 x C] fp64 copy would be GBs)
                for img in p:
                    v = img.reshape(-1, z.C).double()
                    s += v.sum(0)
                    ss += (v * v).sum(0)
                    n += v.shape[0]
            mean = s / n
            unbiased = (ss - s * mean) / (n - 1)
            want_m = 0.9 * before[k][0] + 0.1 * mean
            want_v = 0.9 * before[k][1] + 0.1 * unbiased
            rec["sync_bn_%s_running_mean" % k] = float((bn.running_mean.double() - want_m).abs().max() / want_m.abs().max())
            rec["sync_bn_%s_running_var" % k] = float((bn.running_var.double() - want_v).abs().max() / want_v.abs().max())
            rec["sync_bn_%s_samples" % k] = n
        del parts
    # rank-local batch softmax against the global-batch one, on the gathered logits (rank 0's own slice)
    parts = [torch.empty_like(logits) for _ in range(world)] if rank == 0 else None
    dist.gather(logits.contiguous(), parts, dst=0)
    if rank == 0:
        allx = torch.cat(parts, 0)
        glob = torch.softmax(allx, 0)[:logits.shape[0]]
        loc = torch.softmax(logits, 0)
        rec["softmax_dim0_rank_local_vs_global_rel_l2"] = float((loc - glob).double().norm() / glob.double().norm())
        rec["softmax_dim0_rank_local_over_global_mean_ratio"] = float(loc.double().mean() / glob.double().mean())
        del allx, glob, loc
    del parts
    if sub("functional").GLOBAL_SOFTMAX0[0]:
        rec["softmax_dim0"] = ("global batch (S2R_GLOBAL_SOFTMAX0=1: batch maximum / sum of exponentials / sum of g*y all-reduced per "
                               "discriminator evaluation; two-rank parity: tests/test_gpu_multirank.py)")
    else:
        rec["softmax_dim0"] = "rank-local batch (stated deviation; north_star: only BN statistics and gradients are exchanged)"
    rec["cross_entropy_mean"] = "global batch (valid-pixel counts all-reduced)"
    # one more full step, then compare the replicas
    run_step(it)
    torch.cuda.synchronize()
    flat = torch.cat([v.detach().reshape(-1).float() for v in G.state_dict().values() if v.dtype.is_floating_point])
    mine = torch.tensor([float(flat.view(torch.int32).to(torch.int64).sum().item() % (1 << 52)), float(flat.double().abs().sum().item())],
                        dtype=torch.float64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    rec["weights_bit_equal"] = bool(all(torch.equal(a, allv[0]) for a in allv))
    rec["comm_error"] = int(job.L.lib().s2r_comm_error())
    return rec


# ----------------------------------------------------------------------------- workload: adaptation step (default)
def run_adapt(args):
    job = Job()
    torch, dev, world, rank = job.torch, job.dev, job.world, job.rank
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=world > 1)
    D = sub("modeling.discriminator").FCDiscriminator(num_classes=19)
    if args.no_dropout:
        G._s2r_no_dropout = True
    G.to(dev).train()
    D.to(dev).train()
    step = sub("steps").AdaptStep(G, D, lr=5e-4, epochs=1, iters_per_epoch=max(10, 2 * args.steps + 3 * args.warmup + 8))
    B, H, W = args.batch, args.height, args.width
    host = synth(1000 + rank, B, H, W, pin=True)
    d_src, d_lab, d_tgt = (t.to(dev) for t in host)
    use_graph = not args.no_graph
    launches_per_step = 0
    if use_graph:
        l0 = job.L.launches
        step.capture(d_src, d_lab, d_tgt, warmup=1)
        launches_per_step = (job.L.launches - l0) // 2          # one warm-up step + the captured step
        run = step.replay
    else:
        run = step

    def resident(i):
        return run(d_src, d_lab, d_tgt, i=i, epoch=0)

    def end_to_end(i):
        if use_graph:
            out = pipeline_e2e(step, host, i, torch)
        else:
            out = step(*(t.to(dev, non_blocking=True) for t in host), i=i, epoch=0)
        return torch.stack([out['loss_seg'], out['loss_adv'], out['loss_D_src'], out['loss_D_tgt']]).cpu()

    for k in range(args.warmup):
        resident(k)
    # Clocks are polled (every 100 ms) during the device-resident region only: every NVML / nvidia-smi query takes
    # the driver lock for tens of milliseconds, which the end-to-end loop -- a host synchronisation and fresh CUDA
    # calls every step -- pays for directly (measured: 24 ms -> 41-47 ms per step when polled).  The end-to-end
    # region, which follows immediately, is bracketed by one sample before and one after instead.
    with ClockSampler(job.local) as clk:
        ms, launches = job.timed(resident, args.steps, args.warmup)
    clk.sample_once()
    for k in range(args.warmup):
        end_to_end(args.warmup + args.steps + k)
    ms_e2e, _ = job.timed(end_to_end, args.steps, 2 * args.warmup + args.steps)
    clk.sample_once()
    parity = multi_rank_parity(job, G, resident, d_src, 3 * args.warmup + 2 * args.steps) if world > 1 else None
    pairs = B * world
    if rank != 0:
        job.finish(step)
        return
    peaks = load_peaks()
    dwf = depthwise_family_roofline(torch, B, H, W, peaks)
    line = {
        "metric": METRIC, "value": pairs * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "AdaptSegNet output-space adaptation step (train_adapt.py:126-181): DeepLabV3+ MNv2 OS16 "
                               "+ FCDiscriminator, src+tgt %dx%d crops, batch %d/GPU, SGD+Adam, random init" % (H, W, B),
                   "pairs_per_step_per_gpu": B, "parallelism": "dp%d" % world, "sync_bn": world > 1,
                   "dropout": not args.no_dropout, "cuda_graph": use_graph,
                   "streams": "two pass chains (G(src) fwd/bwd + D training | G(tgt) fwd + adversarial bwd) and one "
                              "weight-gradient side stream per chain, all inside the one graph"
                              if os.environ.get("S2R_OVERLAP", "1") != "0" else "one",
                   "discriminator_forward_on_target": ("evaluated once per step and shared by the adversarial pass "
                                                       "(train_adapt.py:151) and the discriminator's training pass (:174): same "
                                                       "tensor, same weights, identical values and gradients"
                                                       if os.environ.get("S2R_SHARE_D_FWD", "1") != "0" else "evaluated twice"),
                   "bn_exchange": ("nvlink peer memory (csrc/comm.cu)" if sub("engine").PEER["world"] == world else "nccl")
                   if world > 1 else "none",
                   "l2": "per-step working set (>4 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "e2e_input_pipeline": "H2D of step k+1 (pinned -> staging, copy stream) overlaps replay k; every step's "
                                         "inputs are copied inside the timed region"},
        "e2e": {"value": pairs * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host)), "d2h_bytes_per_step": 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps if use_graph else launches,
        "clocks": clk.summary(),
        "roofline": dominant_kernel_roofline(torch, B, H, W, peaks),
        "roofline_depthwise": dwf["fwd"],
        "roofline_depthwise_bwd": dwf["bwd"],
    }
    if parity is not None:
        line["multi_rank_parity"] = parity
    if world == 1:
        line["val"] = val_quick(torch, G)
    if not args.no_cpu_baseline and world == 1:
        val, dt, threads = cpu_adapt_steps(args.cpu_batch, H, W, 1, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "1 step (after 1 warm-up) of the oracle port of train_adapt.py:137-181 at batch %d, "
                                          "%dx%d, fp32, %d torch CPU threads (%.1f s/step)" % (args.cpu_batch, H, W, threads, dt)}
    print(json.dumps(line), flush=True)
    job.finish(step)


def val_quick(torch, G):
    """Extra key of the default line: BASELINE config 5 in brief (one resident 1x3x1024x2048 image replayed 100x over 3
    graph lanes).  The full workload -- 500 distinct images from pinned host memory, e2e, roofline, CPU arm -- is
    `bench.py --workload val`."""
    dev = torch.device("cuda", torch.cuda.current_device())
    was_training = G.training
    G.eval()
    vstep = sub("steps").ValStep(G, 19)
    img, tl = val_image(torch, 0, 1024, 2048)
    img, tl = img.to(dev), tl.to(dev)
    lanes = int(os.environ.get("S2R_VAL_LANES", "3"))
    vstep.capture(img, tl, lanes=lanes)
    for _ in range(4):
        vstep.replay(img, tl)
    vstep.finish()
    torch.cuda.synchronize()
    n_img = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_img):
        vstep.replay(img, tl)
    vstep.finish()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_img
    miou, _ = vstep.evaluator.Mean_Intersection_over_Union()
    total = int(vstep.evaluator.confusion_matrix.sum())
    assert total == (n_img + 4) * int((tl != 255).sum()), "confusion matrix lost counts"
    G.train(was_training)
    return {"value": 1e3 / ms, "unit": "img/s", "ms_per_image": ms,
            "config": "val_adapt.py:122-135 at 1x3x1024x2048, eval forward + fused argmax/confusion matrix, CUDA graph, "
                      "%d replays of one resident image over %d graph lane(s); full workload: --workload val" % (n_img, lanes),
            "miou_random_init": float(miou)}


# ----------------------------------------------------------------------------- workload: feature-adaptation step
def run_feature(args):
    """BASELINE config 4: train.py:163-216 -- backbone, ASPP, decoder and DomainClassifer as four separate modules
    (train.py:47-57), the script's default Adam optimizers (task, d, d_inv), batch 8/GPU, 512x1024."""
    job = Job()
    torch, dev, world, rank = job.torch, job.dev, job.world, job.rank
    nn = torch.nn
    BN = sub("modeling.sync_batchnorm").SynchronizedBatchNorm2d if world > 1 else nn.BatchNorm2d
    torch.manual_seed(1)
    bb = sub("modeling.backbone.mobilenet").MobileNetV2(output_stride=16, BatchNorm=BN).to(dev).train()
    aspp = sub("modeling.assp").ASPP('mobilenet', 16, BN).to(dev).train()
    dec = sub("modeling.decoder").Decoder(19, 'mobilenet', BN).to(dev).train()
    dc = sub("modeling.domian").DomainClassifer('mobilenet', BN).to(dev).train()
    if args.no_dropout:
        for m in (aspp, dec, dc):
            m._s2r_no_dropout = True
    step = sub("steps").FeatureStep(bb, aspp, dec, dc, lr=5e-4, optimizer='Adam', epochs=1,
                                    iters_per_epoch=max(10, 2 * args.steps + 3 * args.warmup + 8))
    B, H, W = args.batch, args.height, args.width
    host = synth(1000 + rank, B, H, W, pin=True)
    d_src, d_lab, d_tgt = (t.to(dev) for t in host)
    l0 = job.L.launches
    step.capture(d_src, d_lab, d_tgt, warmup=1)
    launches_per_step = (job.L.launches - l0) // 2

    def resident(i):
        return step.replay(d_src, d_lab, d_tgt, i=i, epoch=0)

    def end_to_end(i):
        out = pipeline_e2e(step, host, i, torch)
        return torch.stack([out['task_loss'], out['d_loss'], out['d_inv_loss'], out['d_acc'].float()]).cpu()

    for k in range(args.warmup):
        resident(k)
    with ClockSampler(job.local) as clk:
        ms, _ = job.timed(resident, args.steps, args.warmup)
    clk.sample_once()
    for k in range(args.warmup):
        end_to_end(args.warmup + args.steps + k)
    ms_e2e, _ = job.timed(end_to_end, args.steps, 2 * args.warmup + args.steps)
    clk.sample_once()
    pairs = B * world
    if rank != 0:
        job.finish(step)
        return
    peaks = load_peaks()
    metric, unit = METRICS["feature"]
    line = {
        "metric": metric, "value": pairs * args.steps / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "FCN-in-the-wild feature adaptation step (train.py:163-216): MobileNetV2 + ASPP + decoder + "
                               "DomainClassifer, src+tgt %dx%d crops, batch %d/GPU, three Adam optimizers, random init" % (H, W, B),
                   "pairs_per_step_per_gpu": B, "parallelism": "dp%d" % world, "sync_bn": world > 1,
                   "dropout": not args.no_dropout, "cuda_graph": True,
                   "l2": "per-step working set (>4 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "e2e_input_pipeline": "H2D of step k+1 (pinned -> staging, copy stream) overlaps replay k"},
        "e2e": {"value": pairs * args.steps / (ms_e2e * 1e-3), "unit": unit,
                "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host)), "d2h_bytes_per_step": 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk.summary(),
        # the largest kernel of this step: DomainClassifer's 3x3 1024->1024 on the 32x64 ASPP map (38.7 GF/img, §8a10)
        "roofline": conv_roofline(torch, peaks, "tap-GEMM conv fwd 3x3 1024->1024 (DC_adnn2.0)", B, H // 16, W // 16, 1024,
                                  1024, 3, 1),
    }
    if not args.no_cpu_baseline and world == 1:
        val, dt, threads = cpu_feature_steps(args.cpu_batch, H, W, 1, 1)
        line["cpu_baseline"] = {"value": val, "unit": unit, "cores": threads, "kind": "port",
                                "sample": "1 step (after 1 warm-up) of the oracle port of train.py:173-204 at batch %d, %dx%d, "
                                          "fp32, %d torch CPU threads (%.1f s/step)" % (args.cpu_batch, H, W, threads, dt)}
    print(json.dumps(line), flush=True)
    job.finish(step)


# ----------------------------------------------------------------------------- workload: validation
def run_val(args):
    """BASELINE config 5: val_adapt.py:117-141 over `steps` (default 500) DISTINCT synthetic 1x3x1024x2048 images.
    value: images resident in HBM; e2e: every image and label map copied from pinned host memory inside the timed
    region (copy stream, two images ahead of the replay), one [19,19] int64 device->host read at the end.  N ranks:
    the images are sharded round-robin (rank r takes r, r+N, ...), one all-reduce of the counts at the end; the total
    stays `steps` images, so the scaling is "strong"."""
    job = Job()
    torch, dev, world, rank = job.torch, job.dev, job.world, job.rank
    H, W = 2 * args.height, 2 * args.width
    torch.manual_seed(1)
    G = sub("modeling.deeplab").DeepLab(backbone='mobilenet', output_stride=16, num_classes=19, sync_bn=False).to(dev).eval()
    n_total = args.steps
    mine = list(range(rank, n_total, world))
    gen = torch.Generator(device=dev)
    d_img = torch.empty((len(mine), 3, H, W), dtype=torch.float32, device=dev)
    d_lab = torch.empty((len(mine), H, W), dtype=torch.float32, device=dev)
    for j, k in enumerate(mine):                               # generated on the device (16.8 GB for 500 images) ...
        im, lb = val_image(torch, k, H, W, device=dev, generator=gen)
        d_img[j], d_lab[j] = im[0], lb[0]
    h_img = torch.empty(d_img.shape, dtype=torch.float32, pin_memory=True)     # ... and kept in pinned host memory for e2e
    h_lab = torch.empty(d_lab.shape, dtype=torch.float32, pin_memory=True)
    h_img.copy_(d_img)
    h_lab.copy_(d_lab)
    valid = int((d_lab != 255).sum().item())
    vstep = sub("steps").ValStep(G, 19)
    lanes = int(os.environ.get("S2R_VAL_LANES", "3"))
    l0 = job.L.launches
    vstep.capture(d_img[0:1], d_lab[0:1], lanes=lanes)
    launches_per_image = (job.L.launches - l0) // (lanes + 1)
    for j in range(min(args.warmup + 1, len(mine))):
        vstep.replay(d_img[j:j + 1], d_lab[j:j + 1])
    vstep.finish()
    vstep._evaluator.reset()

    def resident_pass(_):
        for j in range(len(mine)):
            vstep.replay(d_img[j:j + 1], d_lab[j:j + 1])
        vstep.finish()

    copy = torch.cuda.Stream(device=dev)
    depth = 2 * lanes
    ring = [(torch.empty((1, 3, H, W), dtype=torch.float32, device=dev), torch.empty((1, H, W), dtype=torch.float32, device=dev))
            for _ in range(depth)]

    def e2e_pass(_):
        cur = torch.cuda.current_stream(dev)
        ready, freed = [None] * depth, [None] * depth

        def stage(j):
            s = j % depth
            if freed[s] is not None:
                copy.wait_event(freed[s])
            with torch.cuda.stream(copy):
                ring[s][0].copy_(h_img[j:j + 1], non_blocking=True)
                ring[s][1].copy_(h_lab[j:j + 1], non_blocking=True)
                ready[s] = torch.cuda.Event()
                ready[s].record(copy)

        for j in range(min(depth - 1, len(mine))):
            stage(j)
        for j in range(len(mine)):
            if j + depth - 1 < len(mine):
                stage(j + depth - 1)
            s = j % depth
            cur.wait_event(ready[s])
            # the lane copies the slot into its static inputs and replays; the slot is free once that copy is done
            freed[s] = vstep.replay(*ring[s])
        return vstep.evaluator.confusion_matrix      # orders every lane, then the one device->host read: [19,19] int64

    with ClockSampler(job.local) as clk:
        ms, _ = job.timed(resident_pass, 1, 0)
    clk.sample_once()
    cm_res = vstep.all_reduce().confusion_matrix.copy() if world > 1 else vstep.evaluator.confusion_matrix.copy()
    vstep._evaluator.reset()
    ms_e2e, _ = job.timed(e2e_pass, 1, 0)
    clk.sample_once()
    cm_e2e = vstep.all_reduce().confusion_matrix.copy() if world > 1 else vstep.evaluator.confusion_matrix.copy()
    tot = torch.tensor([valid], dtype=torch.int64, device=dev)
    if world > 1:
        job.dist.all_reduce(tot)
    if rank != 0:
        job.finish(vstep)
        return
    import numpy as np
    assert int(cm_res.sum()) == int(tot.item()) and np.array_equal(cm_res, cm_e2e), "confusion matrix lost counts"
    peaks = load_peaks()
    metric, unit = METRICS["val"]
    ev = sub("utils.metrics").Evaluator(19)
    ev.confusion_matrix = cm_res
    line = {
        "metric": metric, "value": n_total / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": n_total, "warmup": args.warmup,
        "ms_per_step": ms / n_total, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "val_adapt.py:117-141: eval forward at 1x3x%dx%d + fused argmax / Evaluator confusion matrix, "
                               "%d distinct synthetic images, random init" % (H, W, n_total),
                   "images": n_total, "parallelism": "dp%d (images sharded round-robin)" % world, "cuda_graph": True,
                   "graph_lanes": lanes,
                   "l2": "every image is a different 25 MB input and the per-image working set (~1.2 GB) exceeds the L2",
                   "e2e_input_pipeline": "pinned host -> device ring of %d slots on a copy stream, ahead of the replays" % depth},
        "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": int(4 * (3 * H * W + H * W)),
                "d2h_bytes_per_step": int(19 * 19 * 8 / max(1, n_total)), "d2h_bytes_total": 19 * 19 * 8,
                "ms_per_step": ms_e2e / n_total},
        "gpu_launches": launches_per_image * n_total,
        "clocks": clk.summary(),
        # the largest kernel of the eval forward: the decoder's 3x3 304->256 on the 256x512 map of ONE image
        "roofline": conv_roofline(torch, peaks, "tap-GEMM conv fwd 3x3 304->256 (decoder.last_conv.0), 1 image 256x512", 1, H // 4,
                                  W // 4, 304, 256, 3, 1),
        "miou_random_init": float(ev.Mean_Intersection_over_Union()[0]),
        "confusion_matrix_total": int(cm_res.sum()),
    }
    if not args.no_cpu_baseline and world == 1:
        val, dt, threads = cpu_val_images(3, H, W)
        line["cpu_baseline"] = {"value": val, "unit": unit, "cores": threads, "kind": "port",
                                "sample": "3 images (after 1 warm-up) of the oracle port of val_adapt.py:122-135 at 1x3x%dx%d, fp32, "
                                          "%d torch CPU threads (%.2f s/image)" % (H, W, threads, dt)}
    print(json.dumps(line), flush=True)
    job.finish(vstep)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        {"adapt": run_adapt, "feature": run_feature, "val": run_val}[a.workload](a)
