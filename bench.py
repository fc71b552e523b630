#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 A-KAZE engine (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload extract|match]

Default workload (BASELINE.json metric "AKAZE extract 1080p images/s", configs[2]): one STEP is one pass of
full A-KAZE extraction (Config::default(), 4 octaves x 4 sublevels) over a batch of synthetic textured
1920x1080 grayscale images. `value` is whole-job images/s with the u8 images already resident in HBM
(akz_extract_batch_u8_device; keypoints+descriptors stay on the device, only per-image counts come back);
`e2e` is the same metric through the reference-facing host-buffer call (akz_extract_batch_u8 on pinned host
images: H2D of the images and D2H of keypoints + descriptors inside the timed region).
Multi-GPU: one process per GPU (torchrun), images are independent -> sharded over ranks with no data-path
collective ("weak" scaling: every rank processes a full batch). `--workload match` measures the brute-force
Hamming matcher (configs[4]) instead, database sharded over ranks with an NCCL all-gather + merge kernel.

`--impl reference` times the reference's CPU path. The Rust crate cannot be built in this image (no cargo),
so that arm runs the line-faithful C restatement in oracle/ (cpu_baseline.kind = "port").
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W1080, H1080 = 1920, 1080
ALG_BYTES_PER_PX = 96.9375  # whole default-config extraction, per input pixel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="extract", choices=["extract", "match"])
    ap.add_argument("--images", type=int, default=1024, help="images per step and per GPU (configs[2]: 1024)")
    ap.add_argument("--unique", type=int, default=128, help="distinct synthetic images (cycled to fill a step)")
    ap.add_argument("--batch", type=int, default=1024, help="images per engine call")
    ap.add_argument("--sub-batch", type=int, default=0, help="images per pipeline sub-batch (0 = library default)")
    ap.add_argument("--match-n", type=int, default=1 << 20, help="queries = database size for --workload match")
    ap.add_argument("--match-path", default="auto", choices=["auto", "popc", "tensor"], help="matcher kernel (auto = tensor at these sizes)")
    ap.add_argument("--cpu-images", type=int, default=8, help="images in the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16 TFLOP/s of this pool's B200s (sustained), for the int8 matcher's tensor roofline."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured burst bf16 (MEASURED_PEAKS.json)"
    return 1400.0, "fallback sustained bf16 1.4 PFLOP/s (B200_PROFILING.md)"


# ---- synthetic inputs -----------------------------------------------------------------------------
def synth_images_torch(n, h, w, seed, device):
    """Textured, corner-rich u8 images with photo-like keypoint density (about 4-5 k keypoints and 18 k
    4-neighbour maxima per 1080p frame; the reference's own test photos give 5.1 k / 13.5 k per 2 Mpx):
    five octaves of Gaussian-blurred noise, amplitudes (4,8,16,32,48) at sigma (1.5,3,6,12,24), plus 48
    filled rectangles, then a sigma-1 blur."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device).manual_seed(seed)

    def blur(x, sigma):
        r = int(3 * sigma + 0.5)
        t = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
        k = torch.exp(-t * t / (2 * sigma * sigma))
        k = (k / k.sum()).view(1, 1, 1, -1)
        x = F.conv2d(F.pad(x, (r, r, 0, 0), mode="reflect"), k)
        return F.conv2d(F.pad(x, (0, 0, r, r), mode="reflect"), k.transpose(2, 3))

    def band(m, sigma, down):
        x = blur(torch.randn((m, 1, h // down, w // down), device=device, generator=g), sigma / down)
        if down > 1:
            x = F.interpolate(x, size=(h, w), mode="bicubic", align_corners=False)
        return x / x.std()

    out = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    chunk = 8
    for i0 in range(0, n, chunk):
        m = min(chunk, n - i0)
        img = 128.0 + 4.0 * band(m, 1.5, 1) + 8.0 * band(m, 3.0, 1) + 16.0 * band(m, 6.0, 2) + 32.0 * band(m, 12.0, 4) \
            + 48.0 * band(m, 24.0, 8)
        rects = torch.rand((m, 48, 5), device=device, generator=g).cpu().numpy()
        for j in range(m):
            for r in rects[j]:
                rw, rh = int(8 + r[0] * 112), int(8 + r[1] * 112)
                x0, y0 = int(r[2] * (w - 8)), int(r[3] * (h - 8))
                img[j, 0, y0:y0 + rh, x0:x0 + rw] = float(r[4] * 255.0)
        img = blur(img, 1.0)
        out[i0:i0 + m] = img[:, 0].clamp(0, 255).to(torch.uint8)
    return out


def synth_descriptors_torch(n, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    d = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device=device, generator=g)
    d[:, 60] &= 0x3F
    d[:, 61:] = 0
    return d


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU baseline (oracle port of the reference CPU path) ---------------------------------------------
def cpu_threads():
    return max(1, min(16, os.cpu_count() or 1))


def cpu_extract_sample(images_u8, threads):
    from oracle import akaze_oracle as O
    O.build()
    t0 = time.perf_counter()
    nk = 0
    for im in images_u8:
        r = O.extract(O.unit_float_from_u8(im), threads=threads)
        nk += len(r.keypoints)
        r.close()
    dt = time.perf_counter() - t0
    return len(images_u8) / dt, nk / max(1, len(images_u8))


def cpu_match_sample(q, db):
    from oracle import akaze_oracle as O
    O.build()
    t0 = time.perf_counter()
    O.match_top2(q, db, desc_len=61)
    dt = time.perf_counter() - t0
    return len(q) * len(db) / dt


def numpy_images(n, seed0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import np_restatement as R
    return [R.natural_image(H1080, W1080, seed0 + i) for i in range(n)]


# ---- reference arm -----------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    if args.workload == "extract":
        imgs = numpy_images(args.cpu_images, 1000)
        for _ in range(args.warmup):
            cpu_extract_sample(imgs[:1], threads)
        t0 = time.perf_counter()
        kp = 0.0
        for _ in range(args.steps):
            _, kp = cpu_extract_sample(imgs, threads)
        dt = time.perf_counter() - t0
        value = args.steps * len(imgs) / dt
        sample = "%d synthetic 1920x1080 images per step (numpy generator), default config; derivative stage on %d threads like the reference's scoped pool, everything else single-threaded like the reference" % (len(imgs), threads)
        line = {"metric": "extract_1080p_images_per_s", "value": value, "unit": "images/s", "config": {"workload": "synthetic 1920x1080 grayscale, full extraction, Config::default()", "images_per_step": len(imgs), "keypoints_per_image": kp}}
    else:
        rng = np.random.default_rng(42)
        q = rng.integers(0, 256, (256, 64), dtype=np.uint8)
        db = rng.integers(0, 256, (1 << 18, 64), dtype=np.uint8)
        for _ in range(min(1, args.warmup)):
            cpu_match_sample(q[:16], db)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_match_sample(q, db)
        dt = time.perf_counter() - t0
        value = args.steps * len(q) * len(db) / dt
        threads = 1
        sample = "256 queries x 262144 database descriptors per step (pairs/s is size independent), single thread like the reference"
        line = {"metric": "hamming_match_pairs_per_s", "value": value, "unit": "pairs/s", "config": {"workload": "brute-force Hamming top-2, 486-bit descriptors"}}
    line.update({"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                 "ms_per_step": 1000.0 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "dtype": "f32" if args.workload == "extract" else "u8", "data": "synthetic",
                 "cpu_baseline": {"value": value, "unit": line["unit"], "cores": threads, "kind": "port", "sample": sample},
                 "e2e": {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "note": "reference Rust crate cannot be built here (no cargo/rustc); this is the C restatement in oracle/"})
    print(json.dumps(line))


# ---- B200 arm ------------------------------------------------------------------------------------------
def dist_setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_extract(args):
    import torch
    import akaze_rust_b200 as A
    world, rank, local = dist_setup(args)
    dev = torch.device("cuda", local)
    peak, peak_src = measured_peaks()
    n_img, B = args.images, min(args.batch, args.images)
    uniq = min(args.unique, n_img)
    # inputs: `uniq` distinct images, cycled to n_img per step; 2 MB each, so every step streams far more than L2
    d_uniq = synth_images_torch(uniq, H1080, W1080, 1000 + 7919 * rank, dev)
    reps = (n_img + uniq - 1) // uniq
    d_imgs = d_uniq.repeat(reps, 1, 1)[:n_img].contiguous()
    h_imgs = torch.empty((n_img, H1080, W1080), dtype=torch.uint8, pin_memory=True)
    h_imgs.copy_(d_imgs)
    torch.cuda.synchronize()
    eng = A.Engine(local, W1080, H1080, B)
    if args.sub_batch:
        eng.set_sub_batch(args.sub_batch)
    cfg = A.Config.default()
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    img_bytes = H1080 * W1080

    def step_device():
        kp = 0
        for i0 in range(0, n_img, B):
            m = min(B, n_img - i0)
            counts = eng.extract_batch_u8_device(d_imgs.data_ptr() + i0 * img_bytes, m, W1080, H1080, W1080, cfg)
            kp += int(counts.sum())
        return kp

    def step_host():
        kp, d2h = 0, 0
        for i0 in range(0, n_img, B):
            m = min(B, n_img - i0)
            fs = eng.extract_batch_u8([h_imgs[i0 + j].numpy() for j in range(m)], cfg)
            for f in fs:  # keypoints + descriptors are in (pinned) host memory now; only the counts are read here
                kp += f.count
                d2h += f.count * (28 + 64)
                f.release()
        return kp, d2h

    # ---- device-resident leg (`value`)
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier(world)
    l0 = eng.launch_count
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kp_total = 0
    for _ in range(args.steps):
        kp_total += step_device()
    e1.record(stream)
    barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = eng.launch_count - l0
    total_images = sum_over_ranks(float(n_img * args.steps), world)
    value = total_images / (ms / 1000.0)
    kp_per_image = kp_total / float(n_img * args.steps)

    # ---- per-stage timing pass (CUDA events inside the library, one extra step) -> roofline of the dominant kernel
    eng.enable_timing(True)
    eng.stage_times(reset=True)
    step_device()
    st = eng.stage_times(reset=True)
    eng.enable_timing(False)
    tot_ms = sum(v[0] for v in st.values())
    px = W1080 * H1080
    sum_px = px * (1 + 0.25 + 0.0625 + 0.015625) * 4  # all 16 levels
    chunks_px = 0.0  # FED: pixels x launches (each launch reads Lt+Lflow, writes Lt = 12 B/px)
    for o, ns in enumerate(([3, 3, 4], [4, 5, 6, 7], [8, 10, 12, 14], [17, 20, 24, 29])):
        for nsteps in ns:
            chunks_px += (px / 4 ** o) * ((nsteps + 3) // 4)  # k_fed_pp runs up to 4 steps per launch
    alg = {  # algorithmic HBM bytes per image of each stage as implemented
        "fed": 12.0 * chunks_px,
        "detector": 16.0 * sum_px,
        "prep": 12.0 * (sum_px - px),
        "level0": 5.0 * px,
        "contrast": 8.0 * px,
        # gathers of the keypoint stages (mostly L2 hits; the planes they sample were just written):
        "descriptor": kp_per_image * (1241 * 12.0 + 28 + 64),   # 1241 samples x (Lt, Lx, Ly) + keypoint in, descriptor out
        "finalize": kp_per_image * (109 * 8.0 + 5 * 4.0 + 28),  # 109 orientation samples x (Lx, Ly), 5 Ldet reads, keypoint out
    }
    # the timing pass runs the two pipeline stages back to back (no overlap), so each stage's CUDA-event time is
    # that of its kernels alone; the dominant kernel is the stage with the largest share
    dom = max(st, key=lambda k: st[k][0])
    stages = {k: {"ms_per_image": v[0] / n_img, "share": v[0] / tot_ms if tot_ms else 0.0, "launches": int(v[1]),
                  "achieved_gbs": (alg[k] * n_img / (v[0] / 1000.0) / 1e9) if (k in alg and v[0] > 0) else None}
              for k, v in st.items()}
    if dom in alg and st[dom][0] > 0:
        ach = alg[dom] * n_img / (st[dom][0] / 1000.0) / 1e9
    else:
        ach = 0.0
    per_launch = st[dom][0] / max(1, st[dom][1])
    # measured DRAM traffic of the stage from the committed ncu capture, scaled to this run's average launch
    traffic, traffic_src = None, None
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*stage_traffic.json")))  # the newest capture (names sort by round tag)
    if tps:
        with open(tps[-1]) as fh:
            tj = json.load(fh)
        if dom in tj["stages"]:
            traffic = tj["stages"][dom]["dram_bytes_per_image"] * n_img / max(1, st[dom][1])
            traffic_src = "profiles/%s (ncu --set full, dram__bytes_read+write summed over the stage, %d-image capture)" % (
                os.path.basename(tps[-1]), tj.get("images_in_capture", 0))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "avg_launch_ms": per_launch,
                "algorithmic_bytes_per_launch": (alg[dom] * n_img / max(1, st[dom][1])) if dom in alg else None,
                "algorithmic_bytes_per_image": alg.get(dom),
                "note": "stage timed alone by CUDA events inside the library (serialised timing pass). The streaming detector is bound by "
                        "the shared-memory/LSU pipe (82 % of its peak at full load, profiles/r1z_ncu_fullload.txt), its DRAM traffic equals "
                        "the algorithmic bytes; fp32_ridge_frac is the fraction of the FP32 non-FMA peak (SURVEY 8d: 724 flop per input "
                        "pixel, 37.2 TFLOP/s) the stencil stages reach",
                "fp32_ridge_frac": (724.0 * px / 37.2e12) / (1e-3 * sum(st[k][0] for k in ("level0", "contrast", "prep", "fed", "detector") if k in st) / n_img)}
    pipe_ach = ALG_BYTES_PER_PX * px * (n_img * args.steps) / (ms / 1000.0) / 1e9
    roofline_pipeline = {"bound": "hbm", "achieved": pipe_ach, "peak": peak, "unit": "GB/s", "frac": pipe_ach / peak,
                         "algorithmic_bytes_per_image": ALG_BYTES_PER_PX * px}

    # ---- end-to-end leg through the host-buffer API
    e2e = None
    if not args.no_e2e:
        step_host()
        barrier(world)
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            _, b = step_host()
            d2h += b
        barrier(world)
        dt = max_over_ranks(time.perf_counter() - t0, world)
        e2e = {"value": total_images / dt, "unit": "images/s", "h2d_bytes_per_step": n_img * img_bytes,
               "d2h_bytes_per_step": d2h // max(1, args.steps), "ms_per_step": 1000.0 * dt / args.steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        th = cpu_threads()
        imgs = [h_imgs[i].numpy() for i in range(min(args.cpu_images, n_img))]
        v, kpc = cpu_extract_sample(imgs, th)
        cpu = {"value": v, "unit": "images/s", "cores": th, "kind": "port",
               "sample": "%d of this run's 1920x1080 images (oracle C restatement; derivative stage on %d threads like the reference's pool, rest single-threaded); %.0f keypoints/image" % (len(imgs), th, kpc)}
    if rank == 0:
        line = {"metric": "extract_1080p_images_per_s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[2]: synthetic 1920x1080 grayscale batch, full A-KAZE extraction, Config::default() (4 octaves x 4 sublevels)",
                           "images_per_step_per_gpu": n_img, "distinct_images": uniq, "engine_batch": B, "keypoints_per_image": kp_per_image,
                           "l2": "inputs larger than L2 (each step streams %.1f GB of u8 images and ~%.0f MB of intermediates per image)" % (n_img * img_bytes / 1e9, 4 * 4 * sum_px / 1e6)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_pipeline": roofline_pipeline,
                "stages": stages, "cpu_baseline": cpu}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_match(args):
    import torch
    import torch.distributed as dist
    import akaze_rust_b200 as A
    world, rank, local = dist_setup(args)
    dev = torch.device("cuda", local)
    n = args.match_n
    eng = A.Engine(local, 64, 64, 1)
    eng.set_match_path(args.match_path)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    q = synth_descriptors_torch(n, 42, dev)
    db_full = synth_descriptors_torch(n, 43, dev)
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, rank * per + per)
    db = db_full[lo:hi].contiguous()
    del db_full
    part = torch.zeros(n, dtype=torch.int64, device=dev)
    gathered = torch.zeros((world, n), dtype=torch.int64, device=dev)
    out = torch.zeros(n, dtype=torch.int64, device=dev)

    def step():
        eng.match_top2_device(q.data_ptr(), n, db.data_ptr(), hi - lo, part.data_ptr(), db_index_base=lo)
        if world > 1:
            torch.cuda.current_stream().wait_stream(stream)
            dist.all_gather_into_tensor(gathered.view(-1), part)
            stream.wait_stream(torch.cuda.current_stream())
            eng.merge_top2_device(gathered.data_ptr(), world, n, out.data_ptr())

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier(world)
    l0 = eng.launch_count
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    pairs = float(n) * float(n) * args.steps
    value = pairs / (ms / 1000.0)
    launches = eng.launch_count - l0
    # roofline. The matcher runs on the tcgen05 int8 path (matcher_tc.cu): 512 int8 MACs = 1024 ops per descriptor pair,
    # exact s32 accumulation. Peak = 2 x the measured dense bf16 cuBLAS rate (int8 is twice bf16 on B200), the
    # sustained figure because one step is a seconds-long tensor loop under the power cap. The integer-popc
    # roofline of the previous kernel (16 POPC32 per pair, 16 lanes/clk/SM) is kept beside it for comparison.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    popc_peak = 148 * 16 * sm_mhz * 1e6
    bf16, bf16_src = measured_tensor_peak()
    tops = 1024.0 * value / world / 1e12
    roofline = {"bound": "tensor", "kernel": "k_match_tc", "achieved": tops, "peak": 2.0 * bf16, "unit": "TOP/s (int8)", "frac": tops / (2.0 * bf16),
                "traffic": None, "peak_source": "2 x " + bf16_src, "ops_per_pair": 1024,
                "popc_roofline": {"achieved_tpopc": 16.0 * value / world / 1e12, "peak_tpopc": popc_peak / 1e12,
                                  "frac": 16.0 * value / world / popc_peak,
                                  "note": "what the integer-popc kernel (matcher.cu, 95% of this peak) is bounded by; > 1 means the tensor path beats that bound"}}
    if args.match_path == "popc":
        roofline = {"bound": "popc", "kernel": "k_match_top2", "achieved": 16.0 * value / world / 1e12, "peak": popc_peak / 1e12, "unit": "Tpopc/s",
                    "frac": 16.0 * value / world / popc_peak, "traffic": None, "peak_source": "148 SM x 16 POPC/clk x median SM clock under load"}
    e2e = None
    if not args.no_e2e and world == 1:
        m = min(n, 1 << 16)
        hq = q[:m].cpu().numpy()
        hdb = db.cpu().numpy()
        eng.match_top2(hq[:256], hdb)
        t1 = time.perf_counter()
        eng.match_top2(hq, hdb)
        dt = time.perf_counter() - t1
        e2e = {"value": m * float(len(hdb)) / dt, "unit": "pairs/s", "h2d_bytes_per_step": (m + len(hdb)) * 64, "d2h_bytes_per_step": m * 8,
               "note": "%d host queries x %d host database through akz_match_top2" % (m, len(hdb))}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v = cpu_match_sample(q[:256].cpu().numpy(), db[:1 << 18].cpu().numpy())
        cpu = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "port", "sample": "256 queries x 262144 database (oracle, single thread like the reference)"}
    if rank == 0:
        print(json.dumps({"metric": "hamming_match_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "config": {"workload": "configs[4]: brute-force Hamming top-2 of %d x %d 486-bit descriptors, database sharded over %d GPU(s)" % (n, n, world), "match_path": args.match_path},
                          "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import __graft_entry__ as G
    G.build()
    if args.workload == "match":
        return run_match(args)
    return run_extract(args)


if __name__ == "__main__":
    main()
