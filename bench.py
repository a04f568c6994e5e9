#!/usr/bin/env python3
"""Benchmark of the B200 AVIF encode path (contract: task statement + DESIGN.md section 'Measurement').

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload 1080p|4k|1080p-alpha|8k]

One "step" = one pass of the whole hot path over one batch of synthetic RGBA frames per GPU (default:
BASELINE.json configs[1], 64 x 1920x1080 RGBA, quality 80, speed 4, 8-bit), producing the AVIF files in host
memory.  `value` = Mpixels/s with the inputs already resident in HBM; `e2e` = the same through the public API
with pinned HOST buffers (H2D and D2H inside the timed region); `e2e_pageable` = the same with ordinary
(pageable) host memory, what a ravif caller passes.  The one JSON line also carries: per-kernel CUDA-event
times of the timed region, the roofline of the dominant kernel and of the whole path (SURVEY.md 8d bytes),
the other BASELINE workloads (4K 10-bit, 1080p + alpha, 8192^2) measured in the same process, a byte-exact
check of one encoded image per workload against the CPU oracle, the CPU baselines (oracle port on the usable
host cores, libaom proxy) and the SM clocks.
Reference arm (`--impl reference`): rav1e (crates.io) cannot be built here (no Rust toolchain, no network),
so the CPU arm is the oracle PORT of the same algorithm (oracle/ hot path + host packer), compiled on this
host with -O3 -march=native, one process per usable core, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# `threads`: ravif's with_num_threads.  It only sets the AV1 tile count, min(threads, area / min_tile_size^2)
# (av1encoder.rs:665-668), so it changes the bytes; None would mean "cores of this host" (16 on a 1-GPU lease, 128 on the
# full node).  The bench pins it to the node's 128 hardware threads = what the reference does on this machine.
THREADS = 128
WORKLOADS = {
    "1080p": dict(name="64x1920x1080 RGBA8 opaque, quality 80, speed 4, 8-bit (BASELINE configs[1])",
                  n=64, w=1920, h=1080, quality=80.0, speed=4, depth=8, survey_bytes_per_px=37.5),
    "4k": dict(name="16x3840x2160 RGBA8 opaque, quality 80, speed 4, 10-bit (BASELINE configs[2])",
               n=16, w=3840, h=2160, quality=80.0, speed=4, depth=10, survey_bytes_per_px=70.5),
    # configs[3]: 256 images over 8 GPUs = 32 per GPU, real alpha (clean-up kernels + alpha frames), depth Auto -> 10-bit
    "1080p-alpha": dict(name="32x1920x1080 RGBA8 with alpha per GPU, quality 60, speed 6, 10-bit, alpha quality 80 (BASELINE configs[3])",
                        n=32, w=1920, h=1080, quality=60.0, speed=6, depth=10, alpha=True, survey_bytes_per_px=92.5),
    # configs[4]: one large RGB image, tile-parallel on one GPU
    "8k": dict(name="1x8192x8192 RGB8, quality 90, speed 2, 10-bit (BASELINE configs[4])",
               n=1, w=8192, h=8192, quality=90.0, speed=2, depth=10, channels=3, survey_bytes_per_px=69.5),
}


def stage_bytes_per_px(stage, bps, planes):
    """ALGORITHMIC bytes one launch of a kernel moves per image pixel: compulsory reads + writes with
    samples at their native width (bps bytes), `planes` coded planes per pixel (3 colour, +1 with an alpha frame)."""
    lv = {"16x16": 256.0, "8x8": 64.0, "4x4": 16.0, "32x32": 1024.0}
    if stage.startswith("satd_"):          # luma source read + 14-word table written per block
        return bps * planes / 3.0 + 56.0 / lv[stage[5:]]
    if stage.startswith("analyze_"):       # all planes read + table read + 16-byte decision written per block
        return bps * planes + (56.0 + 16.0) / lv[stage[8:]]
    if stage == "decide":                  # decisions read, 8-byte mode info per 4x4 written
        return 16.0 * (1 / 16.0 + 1 / 64.0 + 1 / 256.0) + 0.5
    if stage == "recon_chain":             # source read, reconstruction written, mode info + eob records, levels (int16, ~0.25 / sample)
        return planes * (2 * bps + 0.125 + 0.5) + 0.5
    if stage == "deblock_cdef":            # reconstruction + source read, CDEF output written, 4 of 64 deblocked rows kept
        return planes * bps * (3 + 1 / 16.0)
    if stage == "lr_search":               # CDEF output + source read
        return planes * bps * 2
    if stage == "stream_gather":           # levels read and written once (int16, ~0.25 / sample)
        return planes * 1.0
    return None


def make_images(wl, seed0=0):
    """Deterministic synthetic frames (SURVEY.md 8d recipe).  A few base frames, the rest are cyclic shifts of them
    (distinct content, same statistics) to keep set-up time low; the 8192^2 frame is a 4096^2 frame mirrored 2 x 2."""
    import numpy as np
    import harness as H
    if wl["w"] >= 8192:
        q = H.synth_image(wl["w"] // 2, wl["h"] // 2, seed0, alpha=False)
        top = np.concatenate([q, q[:, ::-1]], axis=1)
        base = [np.ascontiguousarray(np.concatenate([top, top[::-1]], axis=0))]
    else:
        base = [H.synth_image(wl["w"], wl["h"], seed0 + i, alpha=bool(wl.get("alpha"))) for i in range(min(8 if wl["w"] < 3000 else 4, wl["n"]))]
    if wl.get("channels", 4) == 3:
        base = [np.ascontiguousarray(b[..., :3]) for b in base]
    imgs = []
    for i in range(wl["n"]):
        b = base[i % len(base)]
        k = i // len(base)
        imgs.append(np.ascontiguousarray(np.roll(b, (37 * k, 91 * k), axis=(0, 1))) if k else b)
    return imgs


def usable_cores():
    """Cores this process may use: affinity mask bounded by the cgroup CPU quota (what rayon / nproc would see)."""
    n = len(os.sched_getaffinity(0))
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(q) // int(p)))
    except Exception:
        pass
    return max(n, 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


# ---- CPU arms (worker processes; nothing here touches CUDA) ----------------------------------------------------------
_W = {}


def _cpu_init(w, h, alpha, channels):
    """Worker initialiser (untimed): load the libraries and synthesise this worker's sample frame."""
    os.environ["B200_ORACLE_NATIVE"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_encode as CE
    import harness as H
    H.oracle(); H.hostlib()
    img = H.synth_image(w, h, os.getpid() % 8, alpha=alpha)
    _W["img"] = img[..., :3].copy() if channels == 3 else img
    _W["CE"] = CE


def _cpu_wait(sec):
    time.sleep(sec)          # keeps a worker busy so that the executor starts (and initialises) every worker
    return os.getpid()


def _cpu_encode(args):
    """Timed: encode the worker's sample with the oracle CPU pipeline; returns pixels encoded."""
    quality, speed, depth, alpha = args
    img = _W["img"]
    _W["CE"].encode_rgba(img, quality, speed, depth, alpha_mode=1 if alpha else 0, threads=THREADS)
    return img.shape[0] * img.shape[1]


def _aom_encode(args):
    """Timed: the same sample through libaom (Pillow's AVIF plugin), single-threaded per worker.  8-bit 4:4:4 only."""
    import io
    from PIL import Image
    quality, speed = args
    img = _W["img"]
    buf = io.BytesIO()
    Image.fromarray(img[..., :3]).save(buf, "AVIF", quality=int(quality), speed=int(speed), subsampling="4:4:4", max_threads=1)
    return img.shape[0] * img.shape[1], buf.tell()


def _sample_size(wl):
    """Bounded CPU sample: a full-width strip of a workload frame, a quarter of its height (rounded up to 8 rows)."""
    return wl["w"], min(wl["h"], ((wl["h"] // 4 + 7) // 8) * 8)


def cpu_steps(wl, n_steps, with_libaom=False):
    """Oracle port (oracle/ hot path + host packer, -O3 -march=native build made on this host) on every usable core: one
    process per core, each encoding one sample strip per step; process start-up, library build and image synthesis are
    outside the timed regions.  Returns (per-step dicts, libaom proxy dict or None)."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = usable_cores()
    os.environ["B200_ORACLE_NATIVE"] = "1"
    import harness as H
    H.oracle(); H.hostlib()          # build once in the parent
    sw, sh = _sample_size(wl)
    out, proxy = [], None
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn"), initializer=_cpu_init,
                             initargs=(sw, sh, bool(wl.get("alpha")), wl.get("channels", 4))) as ex:
        pids = set(ex.map(_cpu_wait, [1.5] * cores))          # all workers up and initialised
        pids |= set(ex.map(_cpu_wait, [0.01] * cores))
        for _ in range(n_steps):
            t0 = time.perf_counter()
            px = sum(ex.map(_cpu_encode, [(wl["quality"], wl["speed"], wl["depth"], bool(wl.get("alpha")))] * cores))
            dt = time.perf_counter() - t0
            out.append({"value": px / 1e6 / dt, "unit": "Mpixels/s", "cores": cores, "kind": "port", "px": px,
                        "os_cpu_count": os.cpu_count(), "build": "gcc -O3 -march=native (oracle/_native, built on this host)",
                        "sample": "%d strips of %dx%d (of the %dx%d workload frames), one per worker process on %d usable cores "
                                  "(affinity + cgroup quota; os.cpu_count() = %d), threads=%d, encode-only wall %.1f s per step; "
                                  "reference rav1e CPU path: not runnable in this environment"
                                  % (cores, sw, sh, wl["w"], wl["h"], len(pids), os.cpu_count(), THREADS, dt)})
        if with_libaom and wl["depth"] == 8:
            proxy = {"label": "proxy CPU (libaom 3.13 via Pillow, 4:4:4 8-bit), NOT rav1e", "cores": cores, "sample": "%d strips of %dx%d, max_threads=1 per worker" % (cores, sw, sh)}
            for spd in (6, 4):
                t0 = time.perf_counter()
                res = list(ex.map(_aom_encode, [(wl["quality"], spd)] * cores))
                dt = time.perf_counter() - t0
                proxy["speed%d_mpx_s" % spd] = sum(r[0] for r in res) / 1e6 / dt
                proxy["speed%d_bytes_per_mpx" % spd] = sum(r[1] for r in res) / (sum(r[0] for r in res) / 1e6)
    return out, proxy


def cpu_baseline(wl):
    steps, proxy = cpu_steps(wl, 2, with_libaom=True)
    r = steps[-1]
    r.pop("px")
    if proxy:
        r["proxy_libaom"] = proxy
    return r


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, _ = cpu_steps(wl, args.warmup + args.steps)
    steps = steps[args.warmup:]
    v = sum(s["value"] for s in steps) / len(steps)
    cb = dict(steps[-1]); px = cb.pop("px")
    cb["value"] = v
    out = {"impl": "reference", "metric": "Mpixels/s encoded", "value": v, "unit": "Mpixels/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": px / 1e6 / v * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8" if wl["depth"] == 8 else "u16", "data": "synthetic",
           "config": {"workload": wl["name"], "sample": cb["sample"], "threads": THREADS}, "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---- checks of what was timed ----------------------------------------------------------------------------------------
def payloads(avif):
    i = avif.find(b"iloc")
    n = int.from_bytes(avif[i + 10:i + 12], "big")
    items, p = {}, i + 12
    for _ in range(n):
        iid = int.from_bytes(avif[p:p + 2], "big")
        off = int.from_bytes(avif[p + 6:p + 10], "big"); ln = int.from_bytes(avif[p + 10:p + 14], "big")
        items[iid] = avif[off:off + ln]
        p += 14
    return items.get(1), items.get(2)


def verify(wl, img, res):
    """One image of the timed batch against the CPU oracle pipeline, byte for byte (frames up to 4K); the 8192^2 frame
    (minutes of oracle time) is decoded instead: size + PSNR against the source."""
    import io
    import numpy as np
    if wl["w"] * wl["h"] > 3840 * 2160:
        from PIL import Image
        im = Image.open(io.BytesIO(res.avif_file))
        dec = np.asarray(im.convert("RGB"))[::8, ::8].astype(np.float64)
        mse = float(np.mean((dec - img[::8, ::8, :3]) ** 2))
        psnr = 10 * np.log10(255 ** 2 / max(mse, 1e-9))
        return {"how": "decoded by libavif/dav1d: size %dx%d, PSNR %.1f dB" % (im.size[0], im.size[1], psnr), "ok": bool(im.size == (wl["w"], wl["h"]) and psnr > 35)}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_encode as CE
    ref = CE.encode_rgba(img, wl["quality"], wl["speed"], wl["depth"], alpha_mode=1 if wl.get("alpha") else 0, threads=THREADS)
    col, al = payloads(res.avif_file)
    ok = col == ref["color_obu"] and (al == ref["alpha_obu"] if wl.get("alpha") else al is None)
    return {"how": "AV1 payload bytes == CPU oracle pipeline (colour%s)" % (" + alpha" if wl.get("alpha") else ""), "ok": bool(ok)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other workloads, the oracle check and the pageable pass (profiling runs)")
    ap.add_argument("--images", type=int, default=0, help="override the batch size (profiling only; not a bench value)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.images > 0:
        wl["n"] = args.images
        wl["name"] += " [batch overridden to %d images: profiling run]" % args.images
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    # CPU baselines first (rank 0, N == 1 only), before CUDA is initialised in this process
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(wl)

    import cavif_rs_b200 as C
    from cavif_rs_b200 import shard
    if C.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    collective = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout when the communicator is created: keep fd 1 pointed at stderr until the
        # first collective is through, so that rank 0's stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            # north_star: NCCL only to broadcast the quantiser / rate tables (rank 0 -> all), once per batch; every rank
            # installs the copy it RECEIVED into its encoder and never builds tables of its own
            qs = sorted({max(int(C.quality_to_quantizer(wl["quality"])), 1), max(int(C.quality_to_quantizer(80.0)), 1)})
            nbytes = shard.broadcast_and_install_rate_tables(qs, rank, dev)
            dist.barrier()
            torch.cuda.synchronize()
            collective = {"op": "ncclBroadcast of the rate tables, rank 0 -> all, installed from device memory", "bytes": int(nbytes), "base_q_idx": qs}
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def encoder_for(w):
        return (C.Encoder.new().with_quality(w["quality"]).with_speed(w["speed"]).with_bit_depth(w["depth"]).with_num_threads(THREADS)
                .with_alpha_color_mode(C.AlphaColorMode.UnassociatedClean if w.get("alpha") else C.AlphaColorMode.UnassociatedDirty)
                .with_device(local_rank))

    def timed(fn, steps, warmup, sample_clocks):
        last = None
        for _ in range(warmup):
            last = fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        s0, g0 = C.stats(), C.stage_times()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out_bytes = 0
        for _ in range(steps):
            last = fn()
            out_bytes += sum(len(r.avif_file) for r in last)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the step contains host work (symbol packing): the events bracket it on the launching stream, like the wall clock
        s1, g1 = C.stats(), C.stage_times()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        delta = {k: s1[k] - s0[k] for k in s0}
        stages = {k: (g1[k][0] - g0.get(k, (0, 0))[0], g1[k][1] - g0.get(k, (0, 0))[1]) for k in g1}
        return ms, wall, delta, clocks, out_bytes, stages, last

    C.set_profiling(True)
    imgs = make_images(wl, seed0=1000 * rank)     # every rank encodes its own batch (images are independent units)
    px_per_step = wl["n"] * wl["w"] * wl["h"]
    enc = encoder_for(wl)
    nch = wl.get("channels", 4)
    d_imgs = [torch.from_numpy(a).to(dev) for a in imgs]                       # resident in HBM
    h_imgs = [torch.from_numpy(a).pin_memory() for a in imgs]                  # pinned host buffers for e2e
    h_np = [t.numpy() for t in h_imgs]

    ms, wall, d, clocks, out_bytes, stages, _ = timed(lambda: enc.encode_batch(d_imgs, nch), args.steps, args.warmup, True)
    ms2, wall2, d2, _, _, _, last = timed(lambda: enc.encode_batch(h_np, nch), args.steps, max(1, args.warmup // 2), False)
    ms3 = None
    if not args.no_extras:
        ms3, _, _, _, _, _, _ = timed(lambda: enc.encode_batch(imgs, nch), max(2, args.steps // 2), 1, False)
        steps3 = max(2, args.steps // 2)

    if rank == 0:
        total_px = px_per_step * world * args.steps
        value = total_px / 1e6 / (ms / 1e3)
        e2e = total_px / 1e6 / (ms2 / 1e3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
        bps = 1 if wl["depth"] == 8 else 2
        planes = 4 if wl.get("alpha") else 3
        chunks = 4 if wl["n"] >= 16 else (2 if wl["n"] >= 4 else 1)
        px_per_launch = px_per_step / chunks
        table = {}
        for k, (sms, n) in sorted(stages.items(), key=lambda kv: -kv[1][0]):
            if n <= 0:
                continue
            b = stage_bytes_per_px(k, bps, planes)
            avg = sms / n
            table[k] = {"ms_per_step": sms / args.steps, "launches_per_step": n / args.steps, "avg_launch_ms": avg,
                        "algorithmic_bytes_per_px": b, "achieved_gbs": (b * px_per_launch / 1e9 / (avg / 1e3)) if b and avg > 0 else None}
        dom = next(iter(table)) if table else None
        roof = None
        if dom:
            t = table[dom]
            roof = {"bound": "hbm", "kernel": dom, "achieved": t["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": (t["achieved_gbs"] / peak) if t["achieved_gbs"] else None, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(args.workload, {}).get(dom),
                    "peak_source": peak_src, "avg_launch_ms": t["avg_launch_ms"],
                    "note": "algorithmic bytes (%.2f B/px x %.1f Mpx per launch) / CUDA-event time between the marks around the kernel on its launching "
                            "stream, inside the timed region; traffic = ncu dram bytes of one launch (profiles/), null if not captured for this workload. "
                            "Integer-ALU / latency bound by construction (SURVEY.md 8d), not DRAM-bound." % (t["algorithmic_bytes_per_px"] or 0, px_per_launch / 1e6)}
        path_gbs = wl["survey_bytes_per_px"] * px_per_step / 1e9 / (ms / args.steps / 1e3)
        out = {
            "metric": "Mpixels/s encoded", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if wl["depth"] == 8 else "u16", "data": "synthetic",
            "config": {"workload": wl["name"], "images_per_gpu": wl["n"], "quality": wl["quality"], "speed": wl["speed"],
                       "depth": wl["depth"], "threads": THREADS, "parallelism": "images sharded one batch per GPU, no data-path collective",
                       "l2": "inputs (%.0f MB/step) larger than L2, no flush needed" % (px_per_step * nch / 1e6),
                       "tools": "what ravif selects for this preset: square partitions (range of the preset, top-down early termination), 13 intra modes + angle deltas + CFL + intra edge filter, transform-type RDO (5-type reduced / 7-type set), Tune::Psychovisual distortion (luma >= 8x8 in the trials and the CDEF search), deblocking level search, CDEF per-SB strength search, loop restoration (sgrproj, 4 or 16 sets, per-unit search)",
                       "avif_bytes_per_step": out_bytes // max(args.steps, 1)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "Mpixels/s", "h2d_bytes_per_step": d2["h2d_bytes"] // args.steps,
                    "d2h_bytes_per_step": d2["d2h_bytes"] // args.steps, "ms_per_step": ms2 / args.steps, "host_memory": "pinned"},
            "gpu_launches": int(d["kernel_launches"]),
            "breakdown_ms_per_step": {"host_pack_mux_overlapped": d["host_ms_pack"] / args.steps, "wall": wall * 1e3 / args.steps},
            "kernels": table,
            "roofline": roof,
            "roofline_path": {"bound": "hbm", "achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
                              "note": "whole path on SURVEY.md 8d algorithmic bytes (%.1f B/px) over the step time" % wl["survey_bytes_per_px"]},
            "cpu_baseline": cb,
        }
        if collective:
            out["collective"] = collective
        if ms3 is not None:
            out["e2e_pageable"] = {"value": px_per_step * world * steps3 / 1e6 / (ms3 / 1e3), "unit": "Mpixels/s",
                                   "note": "ordinary (pageable) host buffers: staged through the pinned arena by a host memcpy"}
        if not args.no_extras:
            out["verified"] = {args.workload: verify(wl, imgs[-1], last[-1])}
        print_later = out
    del d_imgs, h_imgs, h_np
    torch.cuda.empty_cache()

    # ---- the other BASELINE workloads in the same process (N == 1 only): 3 steps each, host buffers in, files out ---
    if world == 1 and not args.no_extras and args.images == 0:
        others = {}
        for name in WORKLOADS:
            if name == args.workload:
                continue
            w2 = dict(WORKLOADS[name])
            im2 = make_images(w2, seed0=7)
            h2 = [torch.from_numpy(a).pin_memory() for a in im2]
            n2 = [t.numpy() for t in h2]
            e2 = encoder_for(w2)
            c2 = w2.get("channels", 4)
            m2, _, dd, _, ob, _, last2 = timed(lambda: e2.encode_batch(n2, c2), 3, 1, False)
            px2 = w2["n"] * w2["w"] * w2["h"]
            others[name] = {"workload": w2["name"], "e2e_mpx_s": px2 * 3 / 1e6 / (m2 / 1e3), "ms_per_step": m2 / 3, "steps": 3,
                            "avif_bytes_per_step": ob // 3, "h2d_bytes_per_step": dd["h2d_bytes"] // 3, "d2h_bytes_per_step": dd["d2h_bytes"] // 3,
                            "roofline_path_gbs": w2["survey_bytes_per_px"] * px2 / 1e9 / (m2 / 3 / 1e3)}
            print_later["verified"][name] = verify(w2, im2[-1], last2[-1])
            del h2, n2, im2
        print_later["other_workloads"] = others
    if rank == 0:
        print(json.dumps(print_later), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ncu dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of a kernel (profiles/, 16-image chunks of the 1080p workload)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {
    # profiles/r02_final_kernel_times_and_dram.txt (read + write per launch)
    "1080p": {"analyze_16x16": 111.2e6, "recon_chain": 215.3e6, "deblock_cdef": 378.6e6, "lr_search": 205.7e6, "analyze_8x8": 75.9e6,
              "satd_16x16": 38.9e6, "lf_search": 56.5e6},
}

if __name__ == "__main__":
    main()
