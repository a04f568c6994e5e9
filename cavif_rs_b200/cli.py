"""cavif-compatible command line driver over the B200 encode path.

Mirrors /root/reference/src/main.rs:44-252: same flags, same defaults (quality 80, speed 4, alpha quality derived
from the colour quality :113, depth auto), same output-path rules and the same per-file stdout line (:213)

    out.avif: 12KB (11000B color, 500B alpha, 300B HEIF)

Where the reference fans the files out over a rayon thread pool (one CPU encode per file, :223), this driver hands
ALL files to one `encode_batch` call: the batch is the unit the device pipeline chunks and overlaps.  Image loading
(PNG/JPEG -> RGBA8, :254-287) goes through Pillow.

    python -m cavif_rs_b200 [-Q n] [-s n] [-j n] [-f] [-o path] [-q] [--dirty-alpha] [--color ycbcr|rgb]
                            [--depth 8|10|auto] IMAGES...
"""
import argparse
import io
import os
import sys


def _quality(arg):                                   # main.rs:28-34
    q = float(arg)
    if q < 1.0 or q > 100.0:
        raise argparse.ArgumentTypeError("quality must be in 1-100 range")
    return q


def _speed(arg):                                     # main.rs:36-42 (the builder asserts 1..=10)
    s = int(arg)
    if s < 1 or s > 10:
        raise argparse.ArgumentTypeError("speed must be in 1-10 range")
    return s


def build_parser():
    ap = argparse.ArgumentParser(prog="cavif-b200", description="Convert JPEG/PNG images to AVIF image format (B200 encode path)")
    ap.add_argument("-Q", "--quality", type=_quality, default=80.0, metavar="n", help="Quality from 1 (worst) to 100 (best)")
    ap.add_argument("-s", "--speed", type=_speed, default=4, metavar="n", help="Encoding speed from 1 (best) to 10 (fast but ugly)")
    ap.add_argument("-j", "--threads", type=int, default=0, metavar="n",
                    help="Maximum threads the tile layout is planned for (0 = no limit); it changes the output bytes like in ravif")
    ap.add_argument("-f", "--overwrite", "--force", action="store_true", help="Replace files if there's .avif already")
    ap.add_argument("-o", "--output", metavar="path", help="Write output to this path instead of same_file.avif. It may be a file or a directory.")
    ap.add_argument("-q", "--quiet", action="store_true", help="Don't print anything")
    ap.add_argument("--dirty-alpha", action="store_true", help="Keep RGB data of fully-transparent pixels (makes larger, lower quality files)")
    ap.add_argument("--color", choices=["ycbcr", "rgb"], default="ycbcr", help="Internal AVIF color model. YCbCr works better for human eyes.")
    ap.add_argument("--depth", choices=["8", "10", "auto"], default="auto", help="Write 8-bit (more compatible) or 10-bit (better quality) images")
    ap.add_argument("images", nargs="+", metavar="IMAGES", help='One or more JPEG or PNG files to convert. "-" is interpreted as stdin/stdout.')
    return ap


def alpha_quality_for(quality):                      # main.rs:113
    return min((quality + 100.0) / 2.0, quality + quality / 4.0 + 2.0)


def filter_inputs(images, quiet, err=sys.stderr):
    """main.rs:136-163: .avif inputs are skipped (with a warning), everything else is kept in order."""
    keep = []
    for s in images:
        if s == "-":
            keep.append(s)
            continue
        ext = os.path.splitext(s)[1].lower()
        if ext == ".avif":
            # main.rs:144-156: with --quiet every .avif argument is dropped; otherwise an existing one is ignored with a
            # warning and a missing one is kept (with the "-o" hint)
            if quiet:
                continue
            if os.path.exists(s):
                print("warning: ignoring %s, because it's already an AVIF" % s, file=err)
                continue
            print("warning: Did you mean to use -o %s?" % s, file=err)
        keep.append(s)
    return keep


def output_path(inp, output, use_dir):
    """main.rs:179-193.  Returns "-" for stdout."""
    if output is None:
        return "-" if inp == "-" else os.path.splitext(inp)[0] + ".avif"
    if output == "-":
        return "-"
    if inp == "-":
        return output
    if use_dir:
        return os.path.join(output, os.path.splitext(os.path.basename(inp))[0] + ".avif")
    return output


def load_rgba(data):
    """main.rs:254-287 through Pillow: any PNG/JPEG -> HxWx4 uint8 (16-bit samples keep their high byte)."""
    import numpy as np
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    im.load()
    if im.mode in ("I;16", "I;16B", "I;16L", "I"):
        a = (np.asarray(im).astype(np.uint32) >> 8).clip(0, 255).astype(np.uint8)
        return np.ascontiguousarray(np.stack([a, a, a, np.full_like(a, 255)], -1))
    return np.ascontiguousarray(np.asarray(im.convert("RGBA"), dtype=np.uint8))


def summary_line(path, avif_len, color, alpha):      # main.rs:213
    return "%s: %dKB (%dB color, %dB alpha, %dB HEIF)" % (path, -(-avif_len // 1000), color, alpha, avif_len - color - alpha)


def main(argv=None, stdin=None, stdout=None, stderr=None):
    stdin = stdin or sys.stdin.buffer
    stdout = stdout or sys.stdout
    stderr = stderr or sys.stderr
    args = build_parser().parse_args(argv)
    files = filter_inputs(args.images, args.quiet, stderr)
    if not files:
        print("error: No PNG/JPEG files specified", file=stderr)
        return 1
    use_dir = False
    if args.output not in (None, "-"):
        if len(files) > 1:
            os.makedirs(args.output, exist_ok=True)
        use_dir = len(files) > 1 or os.path.isdir(args.output)

    from . import Encoder, ColorModel, AlphaColorMode, BitDepth, Error
    enc = (Encoder.new().with_quality(args.quality).with_speed(args.speed)
           .with_bit_depth({"8": BitDepth.Eight, "10": BitDepth.Ten, "auto": BitDepth.Auto}[args.depth])
           .with_alpha_quality(alpha_quality_for(args.quality))
           .with_internal_color_model(ColorModel.YCbCr if args.color == "ycbcr" else ColorModel.RGB)
           .with_alpha_color_mode(AlphaColorMode.UnassociatedDirty if args.dirty_alpha else AlphaColorMode.UnassociatedClean)
           .with_num_threads(args.threads if args.threads > 0 else None))

    failures, work = [], []          # work: (input name, output path, pixels)
    for inp in files:
        name = "stdin" if inp == "-" else inp
        try:
            data = stdin.read() if inp == "-" else open(inp, "rb").read()
        except OSError as e:
            failures.append("%s: error: Unable to read input image %s: %s" % (name, inp, e))
            continue
        out = output_path(inp, args.output, use_dir)
        if out != "-" and not args.overwrite and os.path.exists(out):
            failures.append("%s: error: %s already exists; skipping" % (name, out))
            continue
        try:
            work.append((name, out, load_rgba(data)))
        except Exception as e:      # undecodable input
            failures.append("%s: error: %s" % (name, e))

    results = []
    if work:
        try:
            results = enc.encode_batch([w[2] for w in work])            # one device pass for every file
        except (Error, ValueError, AssertionError):                     # status 4 maps to ValueError, bad buffers assert
            results = []
            for name, out, px in work:                                  # isolate the failing image(s)
                try:
                    results.append(enc.encode_rgba(px))
                except (Error, ValueError, AssertionError) as e:
                    results.append(None)
                    failures.append("%s: error: %s" % (name, e))
    for (name, out, _), res in zip(work, results):
        if res is None:
            continue
        try:
            if out == "-":
                (stdout.buffer if hasattr(stdout, "buffer") else stdout).write(res.avif_file)
            else:
                if not args.quiet:
                    print(summary_line(out, len(res.avif_file), res.color_byte_size, res.alpha_byte_size), file=stdout)
                with open(out, "wb") as f:
                    f.write(res.avif_file)
        except OSError as e:
            failures.append("%s: error: Unable to write output image: %s" % (name, e))
    if failures:
        if not args.quiet:
            for f in failures:
                print("error: %s" % f, file=stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
