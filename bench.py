#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 A-KAZE engine (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--legs extract,extract_4k,match]

One run prints ONE JSON line. Its headline fields are BASELINE.json's first metric on configs[2] -- full A-KAZE
extraction (Config::default(), 4 octaves x 4 sublevels) of a batch of synthetic textured 1920x1080 grayscale images:
one STEP is one pass over the batch; `value` is whole-job images/s with the u8 images already resident in HBM
(akz_extract_batch_u8_device; keypoints + descriptors stay on the device, only per-image counts come back); `e2e` is
the same metric through the reference-facing host-buffer call (akz_extract_batch_u8 on pinned host images: H2D of the
images and D2H of keypoints + descriptors inside the timed region). Two more legs ride in the same line:
  "extract_4k": configs[3], 3840x2160, 32 images per GPU and step (256 images over 8 GPUs), same fields;
  "single_image": configs[0]/[1] analogue: one 1080p image per call through the public host API, median wall-clock ms;
  "match":      configs[4] and BASELINE.json's second metric: brute-force Hamming top-2 of 1M x 1M 486-bit descriptors,
                the database sharded by index over the ranks, per-shard top-2 records all-gathered with NCCL and merged
                INSIDE the library (akz_match_top2_sharded_device); strong scaling.
Multi-GPU: one process per GPU (torchrun). Extraction shards by image with no data-path collective ("weak": every
rank processes a full batch). `parity_check` compares the GPU's keypoints and descriptors of this run's first images
with the CPU oracle's for the same images; a red check fails the run.

`--impl reference` times the reference's CPU path on the same images. The Rust crate cannot be built in this image (no
cargo), so that arm runs the line-faithful C restatement in oracle/ (cpu_baseline.kind = "port").
"""
import argparse
import glob
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W1080, H1080 = 1920, 1080
W4K, H4K = 3840, 2160
ALG_BYTES_PER_PX = 96.9375  # whole default-config extraction, per input pixel
ANGLE_TOL = 5e-7            # 2 ulp at pi: f64 atan2 rounded once (device) vs glibc atan2f (oracle)
DESC_BIT_TOL = 1e-2         # north star: >= 99 % of the descriptor bits on matched keypoints


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--legs", default="extract,extract_4k,match", help="comma list of extract, extract_4k, match")
    ap.add_argument("--workload", default=None, choices=["extract", "match"], help="legacy: one leg only, printed as the headline")
    ap.add_argument("--images", type=int, default=1024, help="1080p images per step and per GPU (configs[2]: 1024)")
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic 1080p images per rank (cycled to fill a step)")
    ap.add_argument("--images-4k", type=int, default=32, help="3840x2160 images per step and per GPU (configs[3]: 256 over 8 GPUs)")
    ap.add_argument("--unique-4k", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1024, help="images per engine call")
    ap.add_argument("--sub-batch", type=int, default=0, help="images per pipeline sub-batch (0 = library default)")
    ap.add_argument("--match-n", type=int, default=1 << 20, help="queries = database size of the match leg")
    ap.add_argument("--match-path", default="auto", choices=["auto", "popc", "tensor"], help="matcher kernel (auto = tensor at these sizes)")
    ap.add_argument("--cpu-images", type=int, default=8, help="images in the bounded CPU sample / parity check")
    ap.add_argument("--single-calls", type=int, default=30, help="calls of the one-image latency measurement (0 = skip)")
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


# ---- synthetic inputs: ONE generator for both arms --------------------------------------------------------------------
# Image i of rank r is np_restatement.natural_image(h, w, seed0 + r * distinct + i): numpy PCG64, five octaves of
# Gaussian-blurred noise (amplitudes 4,8,16,32,48 at sigma 1.5,3,6,12,24) plus 48 filled rectangles per 2 Mpx and a
# sigma-1 blur -- photo-like keypoint density (~7 k keypoints per 1080p frame; the reference's own 3 Mpx test photos
# give 7.4 k / 5.6 k). SURVEY 8(d) seeds: 1000 + i for configs[2], 5000 + i for configs[3].
IMAGE_RECIPE = "tests/np_restatement.natural_image (numpy PCG64(seed0 + i)): 5 octaves of blurred noise + rectangles, photo-like density"


def _gen_one(a):
    import np_restatement as R
    h, w, seed = a
    return R.natural_image(h, w, seed)


def bench_images(n, h, w, seed0, world=1):
    """n distinct images, generated by forked workers BEFORE anything touches CUDA."""
    procs = max(1, min(16, n, (os.cpu_count() or 1) // max(1, world)))
    args = [(h, w, seed0 + i) for i in range(n)]
    if procs == 1:
        return [_gen_one(a) for a in args]
    with multiprocessing.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_one, args)


def synth_descriptors_torch(n, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    d = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device=device, generator=g)
    d[:, 60] &= 0x3F
    d[:, 61:] = 0
    return d


def extract_config(h, w):
    which = "configs[2]" if (h, w) == (H1080, W1080) else "configs[3]"
    return {"workload": "%s: synthetic %dx%d grayscale batch, full A-KAZE extraction, Config::default() (4 octaves x 4 sublevels)" % (which, w, h),
            "images": IMAGE_RECIPE, "seed0": 1000 if (h, w) == (H1080, W1080) else 5000}


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


def merge_clocks(a, b):
    """clocks of the whole run = the samples of all legs (headline key `clocks`)."""
    if not a:
        return b
    if not b:
        return a
    sm = [x for x in (a.get("sm_mhz"), b.get("sm_mhz")) if x is not None]
    return {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": max([x for x in (a.get("sm_max_mhz"), b.get("sm_max_mhz")) if x is not None] or [None]),
            "reasons": sorted(set(a.get("reasons", [])) | set(b.get("reasons", []))), "samples": a.get("samples", 0) + b.get("samples", 0)}


# ---- CPU baseline (oracle port of the reference CPU path) ---------------------------------------------
def cpu_threads():
    return max(1, min(16, os.cpu_count() or 1))


def cpu_extract_sample(images_u8, threads, keep=False):
    """Times the oracle on the images; with keep=True also returns (keypoints, descriptors) per image for the parity check."""
    from oracle import akaze_oracle as O
    O.build()
    t0 = time.perf_counter()
    nk, feats = 0, []
    for im in images_u8:
        r = O.extract(O.unit_float_from_u8(im), threads=threads)
        nk += len(r.keypoints)
        if keep:
            feats.append((r.keypoints.copy(), r.descriptors.copy()))
        r.close()
    dt = time.perf_counter() - t0
    return len(images_u8) / dt, nk / max(1, len(images_u8)), feats


def cpu_match_sample(q, db):
    from oracle import akaze_oracle as O
    O.build()
    t0 = time.perf_counter()
    O.match_top2(q, db, desc_len=61)
    dt = time.perf_counter() - t0
    return len(q) * len(db) / dt


def parity_report(gpu_feats, cpu_feats):
    """GPU vs oracle on the same images: north-star tolerances (>= 99 % keypoints within 0.5 px and the same octave,
    >= 99 % descriptor bits on matched keypoints); the measured state is identical keypoints in identical order."""
    rep = {"images": len(cpu_feats), "keypoints_gpu": 0, "keypoints_oracle": 0, "identical_keypoints": 0, "max_angle_diff": 0.0,
           "descriptor_bits": 0, "descriptor_bits_differing": 0, "ok": True}
    for (kg, dg), (kc, dc) in zip(gpu_feats, cpu_feats):
        rep["keypoints_gpu"] += len(kg)
        rep["keypoints_oracle"] += len(kc)
        if len(kg) != len(kc):
            rep["ok"] = False
            continue
        same = np.ones(len(kc), bool)
        for f in ("x", "y", "response", "size", "octave", "class_id"):
            same &= kg[f] == kc[f]
        rep["identical_keypoints"] += int(same.sum())
        if len(kc):
            rep["max_angle_diff"] = max(rep["max_angle_diff"], float(np.abs(kg["angle"] - kc["angle"]).max()))
            rep["descriptor_bits"] += int(dc.size * 8)
            rep["descriptor_bits_differing"] += int(np.unpackbits(dg[:, :dc.shape[1]] ^ dc).sum())
    n = max(1, rep["keypoints_oracle"])
    rep["keypoint_agreement"] = rep["identical_keypoints"] / n
    rep["descriptor_bit_agreement"] = 1.0 - rep["descriptor_bits_differing"] / max(1, rep["descriptor_bits"])
    rep["ok"] = bool(rep["ok"] and rep["keypoints_gpu"] == rep["keypoints_oracle"] and rep["keypoint_agreement"] >= 0.99
                     and rep["descriptor_bit_agreement"] >= 1.0 - DESC_BIT_TOL and rep["max_angle_diff"] <= 1e-3)
    rep["bit_exact_keypoints"] = rep["identical_keypoints"] == rep["keypoints_oracle"] == rep["keypoints_gpu"]
    rep["tolerance"] = "north star: >= 99 % keypoints (here: bit-identical x, y, response, size, octave, class_id counted), >= 99 % descriptor bits"
    return rep


# ---- reference arm -----------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    if args.workload == "match":
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
        line = {"metric": "hamming_match_pairs_per_s", "value": value, "unit": "pairs/s", "config": match_config(args.match_n, args.gpus, args.match_path)}
    else:
        imgs = bench_images(args.cpu_images, H1080, W1080, 1000)
        for _ in range(args.warmup):
            cpu_extract_sample(imgs[:1], threads)
        t0 = time.perf_counter()
        kp = 0.0
        for _ in range(args.steps):
            _, kp, _ = cpu_extract_sample(imgs, threads)
        dt = time.perf_counter() - t0
        value = args.steps * len(imgs) / dt
        sample = ("%d of the GPU arm's 1920x1080 images per step (seeds 1000..%d), default config, %.0f keypoints/image; derivative stage on %d threads like the "
                  "reference's scoped pool, everything else single-threaded like the reference" % (len(imgs), 1000 + len(imgs) - 1, kp, threads))
        line = {"metric": "extract_1080p_images_per_s", "value": value, "unit": "images/s", "config": extract_config(H1080, W1080)}
    line.update({"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                 "ms_per_step": 1000.0 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "dtype": "u8" if args.workload == "match" else "f32", "data": "synthetic",
                 "cpu_baseline": {"value": value, "unit": line["unit"], "cores": threads, "kind": "port", "sample": sample},
                 "e2e": {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "note": "reference Rust crate cannot be built here (no cargo/rustc); this is the C restatement in oracle/"})
    print(json.dumps(line))


# ---- B200 arm ------------------------------------------------------------------------------------------
def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
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


def fed_chunks_px(px):
    """FED: pixels x launches (each launch reads Lt + Lflow and writes Lt = 12 B/px; k_fed_pp runs up to 4 steps per launch)."""
    tot = 0.0
    for o, ns in enumerate(([3, 3, 4], [4, 5, 6, 7], [8, 10, 12, 14], [17, 20, 24, 29])):
        for nsteps in ns:
            tot += (px / 4 ** o) * ((nsteps + 3) // 4)
    return tot


def run_extract(args, images, h, w, n_img, full):
    """One extraction leg on images of h x w. `full`: the headline leg (stage roofline, CPU sample, parity check)."""
    import torch
    import akaze_rust_b200 as A
    world, rank, local = dist_setup()
    dev = torch.device("cuda", local)
    peak, peak_src = measured_peaks()
    B = min(args.batch, n_img)
    uniq = len(images)
    # inputs: `uniq` distinct images, cycled to n_img per step, in pinned host memory and (for `value`) in HBM
    h_uniq = torch.from_numpy(np.stack(images))
    reps = (n_img + uniq - 1) // uniq
    h_imgs = torch.empty((n_img, h, w), dtype=torch.uint8, pin_memory=True)
    h_imgs.copy_(h_uniq.repeat(reps, 1, 1)[:n_img])
    h_np = h_imgs.numpy()
    d_imgs = h_imgs.to(dev)
    torch.cuda.synchronize()
    eng = A.Engine(local, w, h, B)
    if args.sub_batch:
        eng.set_sub_batch(args.sub_batch)
    cfg = A.Config.default()
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    img_bytes = h * w
    n_par = min(args.cpu_images, uniq, n_img) if full else 0

    def step_device():
        kp = 0
        for i0 in range(0, n_img, B):
            m = min(B, n_img - i0)
            counts = eng.extract_batch_u8_device(d_imgs.data_ptr() + i0 * img_bytes, m, w, h, w, cfg)
            kp += int(counts.sum())
        return kp

    def step_host(keep=0):
        kp, d2h, kept = 0, 0, []
        for i0 in range(0, n_img, B):
            m = min(B, n_img - i0)
            fs = eng.extract_batch_u8(h_np[i0:i0 + m], cfg)  # (m, h, w) view of the pinned host images
            for j, f in enumerate(fs):  # keypoints + descriptors are in (pinned) host memory now; only the counts are read here
                kp += f.count
                d2h += f.count * (28 + 64)
                if i0 + j < keep:
                    kept.append((f.keypoints.copy(), f.descriptors.copy()))
                f.release()
        return kp, d2h, kept

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
    px = w * h
    pipe_ach = ALG_BYTES_PER_PX * px * (n_img * args.steps) / (ms / 1000.0) / 1e9
    out = {"value": value, "unit": "images/s", "ms_per_step": ms / args.steps, "gpu_launches": int(launches), "clocks": clocks,
           "roofline_pipeline": {"bound": "hbm", "achieved": pipe_ach, "peak": peak, "unit": "GB/s", "frac": pipe_ach / peak,
                                 "algorithmic_bytes_per_image": ALG_BYTES_PER_PX * px, "peak_source": peak_src,
                                 "note": "SURVEY 8(d): 96.9375 B per input pixel under perfect fusion x images / s of the timed region"},
           "config": extract_config(h, w),  # identical in the reference arm; the run's own parameters follow
           "run": dict(images_per_step_per_gpu=n_img, distinct_images_per_gpu=uniq, engine_batch=B, keypoints_per_image=kp_per_image,
                       l2="inputs larger than L2: %.0f MB of distinct u8 images, %.1f GB of images and ~%.0f MB of intermediates per image streamed per step"
                             % (uniq * img_bytes / 1e6, n_img * img_bytes / 1e9, 85.0 * px / 1e6))}

    # ---- per-stage timing pass (CUDA events inside the library, one extra step) -> roofline of the dominant kernel
    eng.enable_timing(True)
    eng.stage_times(reset=True)
    step_device()
    st = eng.stage_times(reset=True)
    eng.enable_timing(False)
    tot_ms = sum(v[0] for v in st.values())
    sum_px = px * (1 + 0.25 + 0.0625 + 0.015625) * 4  # all 16 levels
    alg = {  # algorithmic HBM bytes per image of each stage as implemented
        "fed": 12.0 * fed_chunks_px(px) + 4.0 * px,  # level 1 reads the stored gradients (8 B/px) instead of Lflow (4)
        "detector": 16.0 * sum_px,                   # read Lsmooth, write Lx, Ly, Ldet
        "prep": 12.0 * (sum_px - 2 * px) + 12.0 * (px / 4 + px / 16 + px / 64),  # levels 2..15: read Lt (16 B/px when halving), write Lsmooth + Lflow
        "level0": 5.0 * px,
        "contrast": 24.0 * px,                       # read Lt0, write Lsmooth1 + gx + gy, re-read gx + gy for the histogram
        # gathers of the keypoint stages (mostly L2 hits; the planes they sample were just written):
        "descriptor": kp_per_image * (1241 * 12.0 + 28 + 64),   # 1241 samples x (Lt, Lx, Ly) + keypoint in, descriptor out
        "finalize": kp_per_image * (109 * 8.0 + 5 * 4.0 + 28),  # 109 orientation samples x (Lx, Ly), 5 Ldet reads, keypoint out
    }
    # the timing pass runs the pipeline stages back to back (no overlap), so each stage's CUDA-event time is that of its
    # kernels alone; the dominant kernel is the stage with the largest share
    dom = max(st, key=lambda k: st[k][0])
    out["stages"] = {k: {"ms_per_image": v[0] / n_img, "share": v[0] / tot_ms if tot_ms else 0.0, "launches": int(v[1]),
                         "achieved_gbs": (alg[k] * n_img / (v[0] / 1000.0) / 1e9) if (k in alg and v[0] > 0) else None}
                     for k, v in st.items()}
    ach = alg[dom] * n_img / (st[dom][0] / 1000.0) / 1e9 if (dom in alg and st[dom][0] > 0) else 0.0
    per_launch = st[dom][0] / max(1, st[dom][1])
    # measured DRAM traffic of the stage from the committed ncu capture (1080p), scaled to this run's average launch
    traffic, traffic_src = None, None
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*stage_traffic.json")))  # the newest capture (names sort by round tag)
    if tps and (h, w) == (H1080, W1080):
        with open(tps[-1]) as fh:
            tj = json.load(fh)
        if dom in tj["stages"]:
            traffic = tj["stages"][dom]["dram_bytes_per_image"] * n_img / max(1, st[dom][1])
            traffic_src = "profiles/%s (ncu --set full, dram__bytes_read+write summed over the stage, %d-image capture)" % (
                os.path.basename(tps[-1]), tj.get("images_in_capture", 0))
    out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "avg_launch_ms": per_launch,
                       "algorithmic_bytes_per_launch": (alg[dom] * n_img / max(1, st[dom][1])) if dom in alg else None,
                       "algorithmic_bytes_per_image": alg.get(dom),
                       "note": "dominant stage, timed alone by CUDA events inside the library (serialised timing pass); achieved = the stage's "
                               "algorithmic bytes as implemented (DESIGN.md section 4) / its time. fp32_ridge_frac = fraction of the FP32 non-FMA "
                               "peak (SURVEY 8d: 724 flop per input pixel, 37.2 TFLOP/s) the stencil stages reach; roofline_pipeline is the "
                               "whole-pipeline fraction on SURVEY 8(d)'s 96.9375 B/px",
                       "fp32_ridge_frac": (724.0 * px / 37.2e12) / (1e-3 * sum(st[k][0] for k in ("level0", "contrast", "prep", "fed", "detector") if k in st) / n_img)}

    # ---- end-to-end leg through the host-buffer API
    gpu_feats = []
    if not args.no_e2e:
        step_host()
        barrier(world)
        t0 = time.perf_counter()
        d2h = 0
        for s in range(args.steps):
            _, b, kept = step_host(keep=n_par if s == args.steps - 1 else 0)
            d2h += b
            gpu_feats = kept or gpu_feats
        barrier(world)
        dt = max_over_ranks(time.perf_counter() - t0, world)
        out["e2e"] = {"value": total_images / dt, "unit": "images/s", "h2d_bytes_per_step": n_img * img_bytes,
                      "d2h_bytes_per_step": d2h // max(1, args.steps), "ms_per_step": 1000.0 * dt / args.steps}
    else:
        out["e2e"] = None

    # ---- CPU sample of the same images + parity check of the headline workload
    out["cpu_baseline"], out["parity_check"] = None, None
    if full and rank == 0 and world == 1 and not args.no_cpu:
        th = cpu_threads()
        v, kpc, cpu_feats = cpu_extract_sample(images[:n_par], th, keep=True)
        out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": th, "kind": "port",
                               "sample": "the first %d of this run's %dx%d images (oracle C restatement; derivative stage on %d threads like the reference's "
                                         "pool, rest single-threaded); %.0f keypoints/image" % (n_par, w, h, th, kpc)}
        if not gpu_feats:  # --no-e2e: fetch the same images through the host API once
            _, _, gpu_feats = step_host(keep=n_par)
        out["parity_check"] = parity_report(gpu_feats, cpu_feats)
        out["parity_check"]["what"] = ("keypoints + descriptors of images 0..%d of the LAST timed e2e step (batch of %d, default mode) vs the CPU oracle on the "
                                       "same images" % (n_par - 1, n_img))
    eng.close()
    del d_imgs, h_np, h_imgs
    torch.cuda.empty_cache()
    return out


def match_config(n, world, path):
    return {"workload": "configs[4]: brute-force Hamming top-2 of %d x %d 486-bit descriptors, database sharded by index over %d GPU(s), "
                        "NCCL all-gather of the per-shard top-2 records + merge inside the library" % (n, n, world), "match_path": path}


def run_match(args):
    import torch
    import torch.distributed as dist
    import akaze_rust_b200 as A
    world, rank, local = dist_setup()
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
    out = torch.zeros(n, dtype=torch.int64, device=dev)
    if world > 1:
        # the communicator lives in the library (ncclCommInitRank); torch.distributed only ships the 128-byte id
        uid = torch.zeros(A.COMM_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(A.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    torch.cuda.synchronize()

    def step():
        if world > 1:  # shard scan -> ncclAllGather -> merge, all on the engine's stream
            eng.match_top2_sharded_device(q.data_ptr(), n, db.data_ptr(), hi - lo, lo, out.data_ptr())
        else:
            eng.match_top2_device(q.data_ptr(), n, db.data_ptr(), hi - lo, out.data_ptr(), db_index_base=lo)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier(world)
    l0 = eng.launch_count
    if rank == 0:
        sampler.start()
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
    # all ranks must hold the same merged result (checksum of the top-2 records)
    chk = int((out & 0xffffffff).sum().item()) ^ int((out >> 32).sum().item())
    same = True
    if world > 1:
        t = torch.tensor([chk, -chk], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = int(t[0].item()) == chk and int(-t[1].item()) == chk
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
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        hq, hdb = q[:256].cpu().numpy(), db[:1 << 18].cpu().numpy()
        v = cpu_match_sample(hq, hdb)
        cpu = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "port", "sample": "256 queries x 262144 database (oracle, single thread like the reference)"}
        from oracle import akaze_oracle as O
        bi, b, s = O.match_top2(hq, hdb, desc_len=61)
        t = eng.match_top2(hq, hdb, desc_len=61)
        parity = {"ok": bool(np.array_equal(t["best_idx"], bi) and np.array_equal(t["best"], b) and np.array_equal(t["second"], s)),
                  "what": "top-2 records of 256 queries x 262144 database vs the CPU oracle, bit-exact"}
    res = {"metric": "hamming_match_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic", "config": match_config(n, world, args.match_path),
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "parity_check": parity, "ranks_agree": same,
           "exchange": None if world == 1 else "ncclAllGather of %d x 8-byte records per rank inside akz_match_top2_sharded_device, merged by k_merge_top2" % n}
    eng.close()
    del q, db, out
    torch.cuda.empty_cache()
    return res


def run_single_image(args, images, h, w):
    """configs[0]/[1] analogue: ONE image per call through the public API, host buffer in, keypoints and descriptors in host
    memory out (wall clock around the call; a one-sub-batch call replays a captured CUDA graph)."""
    import time
    import torch
    import akaze_rust_b200 as A
    world, rank, local = dist_setup()
    eng = A.Engine(local, w, h, 1)
    img = np.ascontiguousarray(images[0])
    n_kp = 0
    for _ in range(5):
        f = eng.extract_u8(img)
        n_kp = len(f.keypoints)
        f.release()
    l0 = eng.launch_count
    ts = []
    for _ in range(args.single_calls):
        t0 = time.perf_counter()
        f = eng.extract_u8(img)
        kp, de = f.keypoints, f.descriptors_padded
        ts.append((time.perf_counter() - t0) * 1e3)
        f.release()
    launches = (eng.launch_count - l0) // max(1, args.single_calls)
    eng.close()
    torch.cuda.empty_cache()
    ts.sort()
    return {"workload": "one %dx%d image per call (extract_features on an image in host memory): host buffer in, keypoints + descriptors in host memory out" % (w, h),
            "ms_median": ts[len(ts) // 2], "ms_min": ts[0], "calls": args.single_calls, "keypoints": n_kp, "gpu_launches_per_call": int(launches),
            "timing": "host wall clock around the call"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries the ONE JSON line and nothing else: whatever a library prints to file descriptor 1 on the way (NCCL's
    # "NCCL version ..." banner at the first communicator, for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line, failures = run_all(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line))
        sys.stdout.flush()
    if failures:
        sys.stderr.write("PARITY FAILED: %s\n" % ", ".join(failures))
        sys.exit(1)


def run_all(args):
    legs = [x for x in args.legs.split(",") if x]
    if args.workload == "match":
        legs = ["match"]
    elif args.workload == "extract":
        legs = ["extract"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # images first: forked generator processes must not inherit a CUDA context
    imgs, imgs4k = None, None
    if "extract" in legs:
        uniq = min(args.unique, args.images)
        imgs = bench_images(uniq, H1080, W1080, 1000 + rank * uniq, world)
    if "extract_4k" in legs:
        uniq4 = min(args.unique_4k, args.images_4k)
        imgs4k = bench_images(uniq4, H4K, W4K, 5000 + rank * uniq4, world)
    import __graft_entry__ as G
    G.build()
    res = {}
    if "extract" in legs:
        res["extract"] = run_extract(args, imgs, H1080, W1080, args.images, full=True)
    if "extract" in legs and rank == 0 and args.single_calls > 0:
        res["single_image"] = run_single_image(args, imgs, H1080, W1080)
    if "extract_4k" in legs:
        res["extract_4k"] = run_extract(args, imgs4k, H4K, W4K, args.images_4k, full=False)
    if "match" in legs:
        res["match"] = run_match(args)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank != 0:
        return None, []
    failures = []
    if "extract" in res:
        x = res["extract"]
        line = {"metric": "extract_1080p_images_per_s", "value": x["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": x["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": x["config"], "run": x["run"], "clocks": x["clocks"], "e2e": x["e2e"],
                "gpu_launches": x["gpu_launches"], "roofline": x["roofline"], "roofline_pipeline": x["roofline_pipeline"],
                "stages": x["stages"], "cpu_baseline": x["cpu_baseline"], "parity_check": x["parity_check"]}
        if "single_image" in res:
            line["single_image"] = res["single_image"]
        if x["parity_check"] and not x["parity_check"]["ok"]:
            failures.append("extract parity_check")
        if "extract_4k" in res:
            y = res["extract_4k"]
            line["extract_4k"] = {"metric": "extract_4k_images_per_s", "value": y["value"], "unit": "images/s", "ms_per_step": y["ms_per_step"],
                                  "scaling": "weak", "config": y["config"], "run": y["run"], "e2e": y["e2e"], "gpu_launches": y["gpu_launches"],
                                  "roofline": y["roofline"], "roofline_pipeline": y["roofline_pipeline"], "stages": y["stages"], "clocks": y["clocks"]}
            line["clocks"] = merge_clocks(line["clocks"], y["clocks"])
            line["gpu_launches"] += y["gpu_launches"]
        if "match" in res:
            m = res["match"]
            line["match"] = {k: m[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "roofline",
                                               "cpu_baseline", "parity_check", "ranks_agree", "exchange", "clocks")}
            line["clocks"] = merge_clocks(line["clocks"], m["clocks"])
            line["gpu_launches"] += m["gpu_launches"]
    else:
        line = res.get("match") or res.get("extract_4k")
    if "match" in res:
        if res["match"]["parity_check"] and not res["match"]["parity_check"]["ok"]:
            failures.append("match parity_check")
        if not res["match"]["ranks_agree"]:
            failures.append("match: ranks disagree on the merged result")
    return line, failures


if __name__ == "__main__":
    main()
