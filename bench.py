#!/usr/bin/env python3
"""Benchmark of the B200 AVIF encode path (contract: see the task statement / DESIGN.md section 'Measurement').

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload 1080p|4k]

One "step" = one pass of the whole hot path over one batch of synthetic RGBA frames per GPU
(BASELINE.json configs[1]: 64 x 1920x1080 RGBA, quality 80, speed 4, 8-bit), producing the AVIF files
in host memory.  `value` = Mpixels/s with the inputs already resident in HBM; `e2e` = the same through
the public API with pinned HOST input buffers (H2D and D2H inside the timed region).
Reference arm (`--impl reference`): the reference's codec (rav1e, crates.io) cannot be built here (no Rust
toolchain, no network - BASELINE.md section 2), so the CPU arm is the oracle PORT of the same algorithm
(oracle/ hot path + host packer) on all host cores, on a bounded sample of the same workload.
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

WORKLOADS = {
    "1080p": dict(name="64x1920x1080 RGBA8 opaque, quality 80, speed 4, 8-bit (BASELINE configs[1])",
                  n=64, w=1920, h=1080, quality=80.0, speed=4, depth=8),
    "4k": dict(name="16x3840x2160 RGBA8 opaque, quality 80, speed 4, 10-bit (BASELINE configs[2])",
               n=16, w=3840, h=2160, quality=80.0, speed=4, depth=10),
    # configs[3]: 256 images over 8 GPUs = 32 per GPU, real alpha (clean-up kernels + alpha frames), depth Auto -> 10-bit
    "1080p-alpha": dict(name="32x1920x1080 RGBA8 with alpha per GPU, quality 60, speed 6, 10-bit, alpha quality 80 (BASELINE configs[3])",
                        n=32, w=1920, h=1080, quality=60.0, speed=6, depth=10, alpha=True),
    # configs[4]: one large RGB image, tile-parallel on one GPU (partition range clamped to 4..16)
    "8k": dict(name="1x8192x8192 RGB8, quality 90, speed 2, 10-bit (BASELINE configs[4])",
               n=1, w=8192, h=8192, quality=90.0, speed=2, depth=10, channels=3),
}
# Algorithmic bytes per pixel of the RDO stage group (SATD + analysis per level, partition decision, reconstruction
# chain), 16-bit samples:
#   analysis: per level the SATD kernel reads the luma source plane and the trial kernel reads the 3 source planes   3 * (2 + 6) = 24
#             SATD tables written + read: 14 x 4 B per block at 3 levels                      2 * 56 * (1/16 + 1/64 + 1/256) = 9.2
#             decisions written + read: 16 B per block at 3 levels                            2 * 16 * (1/16 + 1/64 + 1/256) = 2.6
#   reconstruction: 3 source planes read, 3 reconstructed planes + 3 coefficient planes written, records      6 + 6 + 6 + 0.9 = 18.9
RDO_BYTES_PER_PX = 24 + 9.2 + 2.6 + 18.9
# ncu dram__bytes_read.sum + dram__bytes_write.sum of the RDO stage group over one 64-image step
# (profiles/r01_final_kernel_times_and_dram.txt: 14.05 GB read + 6.79 GB written over the 4 steps of that run); below the
# algorithmic figure because pruned levels are not read and the 126 MB L2 serves most re-reads of the source planes
TRAFFIC_BYTES_PER_STEP = {"1080p": 5.21e9}


def make_images(wl, seed0=0):
    """Deterministic synthetic frames (SURVEY.md 8d recipe). 8 base frames, the rest are cyclic shifts of them
    (distinct content, same statistics) to keep set-up time low."""
    import numpy as np
    import harness as H
    base = [H.synth_image(wl["w"], wl["h"], seed0 + i, alpha=bool(wl.get("alpha"))) for i in range(min(8, wl["n"]))]
    if wl.get("channels", 4) == 3:
        base = [np.ascontiguousarray(b[..., :3]) for b in base]
    imgs = []
    for i in range(wl["n"]):
        b = base[i % len(base)]
        k = i // len(base)
        imgs.append(np.ascontiguousarray(np.roll(b, (37 * k, 91 * k), axis=(0, 1))) if k else b)
    return imgs


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


_W = {}


def _cpu_init(w, h):
    """Worker initialiser (untimed): load the libraries and synthesise this worker's frame."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_encode as CE
    import harness as H
    H.oracle(); H.hostlib()
    _W["img"] = H.synth_image(w, h, os.getpid() % 8)
    _W["CE"] = CE


def _cpu_wait(sec):
    time.sleep(sec)          # keeps a worker busy so that the executor starts (and initialises) every worker
    return os.getpid()


def _cpu_encode(args):
    """Timed: encode the worker's frame with the oracle CPU pipeline; returns pixels encoded."""
    quality, speed, depth = args
    img = _W["img"]
    _W["CE"].encode_rgba(img, quality, speed, depth)
    return img.shape[0] * img.shape[1]


def _sample_size(wl):
    """Bounded CPU sample: the upper half of a workload frame (full width, height / 2 rounded up to 8 rows)."""
    return wl["w"], min(wl["h"], ((wl["h"] // 2 + 7) // 8) * 8)


def cpu_steps(wl, n_steps):
    """Oracle port (oracle/ hot path + host packer) on all host cores: one process per core, each encoding one sample
    frame per step; process start-up and image synthesis are outside the timed regions.  Returns per-step dicts."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    import harness as H
    H.oracle(); H.hostlib()          # build once in the parent
    sw, sh = _sample_size(wl)
    out = []
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn"), initializer=_cpu_init,
                             initargs=(sw, sh)) as ex:
        pids = set(ex.map(_cpu_wait, [1.5] * cores))          # all workers up and initialised
        pids |= set(ex.map(_cpu_wait, [0.01] * cores))
        for _ in range(n_steps):
            t0 = time.perf_counter()
            px = sum(ex.map(_cpu_encode, [(wl["quality"], wl["speed"], wl["depth"])] * cores))
            dt = time.perf_counter() - t0
            out.append({"value": px / 1e6 / dt, "unit": "Mpixels/s", "cores": cores, "kind": "port", "px": px,
                        "sample": "%d half frames of %dx%d (of the %dx%d workload frames) on %d worker processes, encode-only "
                                  "wall %.1f s per step; reference rav1e CPU path: not runnable in this environment"
                                  % (cores, sw, sh, wl["w"], wl["h"], len(pids), dt)})
    return out


def cpu_baseline(wl):
    r = cpu_steps(wl, 1)[0]
    r.pop("px")
    return r


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = cpu_steps(wl, args.warmup + args.steps)[args.warmup:]
    v = sum(s["value"] for s in steps) / len(steps)
    cb = dict(steps[-1]); px = cb.pop("px")
    cb["value"] = v
    out = {"impl": "reference", "metric": "Mpixels/s encoded", "value": v, "unit": "Mpixels/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": px / 1e6 / v * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
           "config": {"workload": wl["name"], "sample": cb["sample"]}, "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    # CPU baseline first (rank 0, N == 1 only), before CUDA is initialised in this process
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(wl)

    import cavif_rs_b200 as C
    if C.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout when the communicator is created: keep fd 1 pointed at stderr until the
        # first collective is through, so that rank 0's stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            # north_star: NCCL only to broadcast the quantiser/rate tables (rank 0 -> all), once per batch
            import harness as H
            q = C.quality_to_quantizer(wl["quality"])
            tbl = torch.from_numpy(H.rate_tables(max(q, 1)).astype(np.int32)).to(dev) if rank == 0 else \
                torch.zeros(H.RATE_TABLES_BYTES // 2, dtype=torch.int32, device=dev)
            dist.broadcast(tbl, src=0)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    imgs = make_images(wl, seed0=1000 * rank)     # every rank encodes its own batch (images are independent units)
    px_per_step = wl["n"] * wl["w"] * wl["h"]
    enc = (C.Encoder.new().with_quality(wl["quality"]).with_speed(wl["speed"]).with_bit_depth(wl["depth"])
           .with_alpha_color_mode(C.AlphaColorMode.UnassociatedClean if wl.get("alpha") else C.AlphaColorMode.UnassociatedDirty)
           .with_device(local_rank))
    nch = wl.get("channels", 4)
    d_imgs = [torch.from_numpy(a).to(dev) for a in imgs]                       # resident in HBM
    h_imgs = [torch.from_numpy(a).pin_memory() for a in imgs]                  # pinned host buffers for e2e
    h_np = [t.numpy() for t in h_imgs]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        s0 = C.stats()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out_bytes = 0
        for _ in range(steps):
            res = fn()
            out_bytes += sum(len(r.avif_file) for r in res)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the step contains host work (symbol packing): the events bracket it on the launching stream, like the wall clock
        s1 = C.stats()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        delta = {k: s1[k] - s0[k] for k in s0}
        return ms, wall, delta, clocks, out_bytes

    ms, wall, d, clocks, out_bytes = timed(lambda: enc.encode_batch(d_imgs, nch), args.steps, args.warmup, True)
    ms2, wall2, d2, _, _ = timed(lambda: enc.encode_batch(h_np, nch), args.steps, max(1, args.warmup // 2), False)

    if rank == 0:
        total_px = px_per_step * world * args.steps
        value = total_px / 1e6 / (ms / 1e3)
        e2e = total_px / 1e6 / (ms2 / 1e3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
        rdo_ms = d["gpu_ms_rdo"] / max(args.steps, 1)      # CUDA-event time of the RDO stage group, summed over the chunks of a step
        achieved = RDO_BYTES_PER_PX * px_per_step / 1e9 / (rdo_ms / 1e3) if rdo_ms > 0 else 0.0
        out = {
            "metric": "Mpixels/s encoded", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": wl["name"], "images_per_gpu": wl["n"], "quality": wl["quality"], "speed": wl["speed"],
                       "depth": wl["depth"], "parallelism": "images sharded one batch per GPU, no data-path collective",
                       "l2": "inputs (%.0f MB/step) larger than L2, no flush needed" % (px_per_step * nch / 1e6),
                       "tools": "partitions 4..16 (square, top-down early termination), 13 intra modes + angle deltas + CFL, 5 transform types, deblock + CDEF (per-SB strength search) + loop restoration (sgrproj, per-unit search)",
                       "avif_bytes_per_step": out_bytes // max(args.steps, 1)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "Mpixels/s", "h2d_bytes_per_step": d2["h2d_bytes"] // args.steps,
                    "d2h_bytes_per_step": d2["d2h_bytes"] // args.steps, "ms_per_step": ms2 / args.steps},
            "gpu_launches": int(d["kernel_launches"]),
            "breakdown_ms_per_step": {"gpu_rdo_stages": rdo_ms, "gpu_all_streams_summed": d["gpu_ms_total"] / args.steps,
                                      "host_pack_mux_overlapped": d["host_ms_pack"] / args.steps, "wall": wall * 1e3 / args.steps},
            "roofline": {"bound": "hbm", "kernel": "RDO stage group: b200_satd_kernel + b200_analyze_kernel<lv> x3 + b200_decide_kernel + b200_recon_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_STEP.get(args.workload) if args.images == 0 else None,
                         "peak_source": peak_src,
                         "note": "integer-ALU / latency bound by construction (small integer butterflies, SURVEY.md 8d), not DRAM-bound; "
                                 "algorithmic bytes = %.1f B/px; traffic = ncu dram bytes of the group per step (profiles/), null if not captured" % RDO_BYTES_PER_PX},
            "cpu_baseline": cb,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
