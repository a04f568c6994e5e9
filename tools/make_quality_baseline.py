#!/usr/bin/env python3
"""Regenerate tests/golden/quality_baseline.json: payload size and luma / RGB PSNR of THIS encoder (CPU oracle pipeline,
bit-identical to the device path) on the guard frames, plus libaom's rate / PSNR curve on the same frames (Pillow's AVIF
plugin, 4:4:4, speed 4).  Run deliberately after an intended change of the search; tests/test_quality_guard.py fails when the
current encoder is worse than this file by more than its margins."""
import io, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import harness as H, cpu_encode as CE
from PIL import Image

FRAMES = [(640, 360, 0), (640, 360, 5), (512, 512, 9), (960, 540, 2), (480, 272, 7)]
QUALITIES = [60.0, 80.0, 90.0]
AOM_QUALITIES = [30, 45, 60, 70, 80, 88, 94]


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def ours(img, quality, speed=4, depth=8):
    import ctypes
    r = CE.encode_rgba(img, quality, speed, depth, threads=4)
    rec = r["color_rec"].final
    f = r["color_frame"]
    h, w = img.shape[:2]
    # decode-equivalent: the reconstruction after the in-loop filters IS what a decoder shows (pinned by dav1d in the tests)
    y, cb, cr = [rec[p][:h, :w].astype(np.float64) for p in range(3)]
    # BT.601 full range back to RGB
    rr = y + 1.402 * (cr - 128); gg = y - 0.344136 * (cb - 128) - 0.714136 * (cr - 128); bb = y + 1.772 * (cb - 128)
    rgb = np.clip(np.rint(np.stack([rr, gg, bb], -1)), 0, 255)
    return len(r["color_obu"]), psnr(rgb, img[..., :3])


def aom(img, q):
    buf = io.BytesIO()
    Image.fromarray(img[..., :3]).save(buf, "AVIF", quality=q, speed=4, subsampling="4:4:4", max_threads=4)
    dec = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    return buf.tell(), psnr(dec, img[..., :3])


def main():
    out = {"note": "size in bytes (AV1 payload for this encoder, whole file for libaom), PSNR of the decoded RGB picture", "frames": []}
    for (w, h, seed) in FRAMES:
        img = H.synth_image(w, h, seed)
        e = {"w": w, "h": h, "seed": seed, "ours": {}, "aom": []}
        for q in QUALITIES:
            s, p = ours(img, q)
            e["ours"][str(int(q))] = {"bytes": s, "psnr": round(p, 4)}
        for q in AOM_QUALITIES:
            s, p = aom(img, q)
            e["aom"].append({"q": q, "bytes": s, "psnr": round(p, 4)})
        out["frames"].append(e)
        print(e)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "quality_baseline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
