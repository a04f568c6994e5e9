"""Small encodes that touch every kernel (alpha clean, convert, SATD / analysis / decide / recon at every level incl. 32x32,
deblock, CDEF, LR) - run under compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cavif_rs_b200 as C
import harness as H

n = 0
for (w, h, speed, depth, alpha) in ((200, 136, 4, 8, True), (161, 100, 2, 10, False), (96, 80, 6, 10, True), (130, 70, 10, 8, False)):
    img = H.synth_image(w, h, w + h, alpha=alpha)
    r = C.Encoder.new().with_quality(70).with_speed(speed).with_bit_depth(depth).encode_rgba(img)
    n += len(r.avif_file)
# speed 1: four fully evaluated luma modes (complex_prediction_modes), exhaustive partition search
r = C.Encoder.new().with_quality(60).with_speed(1).encode_rgba(H.synth_image(136, 104, 3, alpha=True))
n += len(r.avif_file)
# one tall tile (seven superblock rows): the wavefront of the reconstruction chain at its deepest row-slot count
r = C.Encoder.new().with_quality(80).with_speed(4).with_num_threads(1).encode_rgba(H.synth_image(128, 420, 9, alpha=False))
n += len(r.avif_file)
res = C.Encoder.new().with_quality(80).with_speed(4).encode_batch([H.synth_image(256, 192, i) for i in range(4)])
print("ok", n + sum(len(r.avif_file) for r in res), C.stats())
