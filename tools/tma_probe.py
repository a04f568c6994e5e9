"""GPU probe: encodes at the depths given on the command line (default 10 8); prints whether the fused filter kernel staged its
tiles by TMA and whether any tile failed to land (b200avif_stats.tma_launches / tma_timeouts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cavif_rs_b200 as C, harness as H
for depth in [int(a) for a in sys.argv[1:]] or [10, 8]:
    img = H.synth_image(640, 360, 1)
    s0 = C.stats()
    r = C.Encoder.new().with_quality(80).with_speed(4).with_bit_depth(depth).encode_rgba(img)
    s1 = C.stats()
    print("depth", depth, "bytes", len(r.avif_file), "tma launches", s1["tma_launches"] - s0["tma_launches"], "timeouts", s1["tma_timeouts"], flush=True)
