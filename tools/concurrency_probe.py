"""Repeat the concurrent-encode scenario of tests/test_gpu_e2e.py and report WHAT differs (colour / alpha payload, sizes)."""
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cavif_rs_b200 as C
import harness as H
from test_gpu_e2e import payloads
import dav1d, numpy as np

enc = C.Encoder.new().with_quality(65).with_speed(int(os.environ.get('PROBE_SPEED', '6')))
if os.environ.get('PROBE_DIRTY'):
    enc = enc.with_alpha_color_mode(C.AlphaColorMode.UnassociatedDirty)
if os.environ.get('PROBE_DEPTH'):
    enc = enc.with_bit_depth(int(os.environ['PROBE_DEPTH']))
imgs = [H.synth_image(160 + 16 * i, 96 + 8 * i, 20 + i, alpha=(i % 3 == 0)) for i in range(8)]
serial = [enc.encode_rgba(im).avif_file for im in imgs]
again = [enc.encode_rgba(im).avif_file for im in imgs]
print("serial repeat identical:", again == serial)
for rnd in range(int(os.environ.get('PROBE_ROUNDS', '20'))):
    out = [None] * len(imgs)
    def work(i):
        out[i] = enc.encode_rgba(imgs[i]).avif_file
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    for t in th: t.start()
    for t in th: t.join()
    for i in range(len(imgs)):
        if out[i] != serial[i]:
            c0, a0 = payloads(serial[i]); c1, a1 = payloads(out[i])
            print("round", rnd, "image", i, "colour differs" if c0 != c1 else "colour same", len(c0), len(c1),
                  "alpha differs" if a0 != a1 else "alpha same", None if a0 is None else len(a0), None if a1 is None else len(a1))
            for name, x0, x1 in (("colour", c0, c1), ("alpha", a0, a1)):
                if x0 is None or x0 == x1:
                    continue
                for flt in (0, 7):
                    d0, _, _ = dav1d.decode(x0, flt); d1, _, _ = dav1d.decode(x1, flt)
                    for p in range(len(d0)):
                        diff = np.argwhere(d0[p] != d1[p])
                        if len(diff):
                            print("   ", name, "filters", flt, "plane", p, "differing px", len(diff), "rows", diff[:, 0].min(), diff[:, 0].max(),
                                  "cols", diff[:, 1].min(), diff[:, 1].max(), "of", d0[p].shape)
                        else:
                            print("   ", name, "filters", flt, "plane", p, "decoded identical")
print("done", {k: os.environ[k] for k in os.environ if k.startswith("PROBE") or k.startswith("B200AVIF")})
