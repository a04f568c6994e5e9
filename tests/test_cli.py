"""cavif-compatible CLI driver (cavif_rs_b200/cli.py) against the reference's behaviour (src/main.rs, tests/stdio.rs).
Host logic runs everywhere; the encoding cases need the GPU."""
import io
import os
import subprocess
import sys
import numpy as np
import pytest
import harness as H
from PIL import Image

sys.path.insert(0, H.ROOT)
from cavif_rs_b200 import cli


def test_flag_parsing_and_defaults():                     # main.rs:44-110
    a = cli.build_parser().parse_args(["x.png"])
    assert (a.quality, a.speed, a.threads, a.overwrite, a.output, a.quiet, a.dirty_alpha, a.color, a.depth) == \
        (80.0, 4, 0, False, None, False, False, "ycbcr", "auto")
    a = cli.build_parser().parse_args(["-Q", "55", "-s", "7", "-j", "3", "--force", "-o", "out", "-q", "--dirty-alpha",
                                       "--color", "rgb", "--depth", "10", "a.png", "b.jpg"])
    assert (a.quality, a.speed, a.threads, a.overwrite, a.output, a.quiet, a.dirty_alpha, a.color, a.depth, a.images) == \
        (55.0, 7, 3, True, "out", True, True, "rgb", "10", ["a.png", "b.jpg"])
    for bad in (["-Q", "0", "x.png"], ["-Q", "101", "x.png"], ["-s", "0", "x.png"], ["-s", "11", "x.png"], ["--depth", "12", "x.png"]):
        with pytest.raises(SystemExit):
            cli.build_parser().parse_args(bad)


def test_alpha_quality_rule():                            # main.rs:113, SURVEY App. A.1
    assert cli.alpha_quality_for(80.0) == 90.0 and cli.alpha_quality_for(60.0) == 77.0 and cli.alpha_quality_for(90.0) == 95.0


def test_output_path_rules(tmp_path):                     # main.rs:165-193
    assert cli.output_path("dir/a.png", None, False) == "dir/a.avif"
    assert cli.output_path("-", None, False) == "-"
    assert cli.output_path("a.png", "-", False) == "-"
    assert cli.output_path("-", "o.avif", False) == "o.avif"
    assert cli.output_path("x/a.jpeg", "outdir", True) == os.path.join("outdir", "a.avif")
    assert cli.output_path("x/a.jpeg", "single.avif", False) == "single.avif"


def test_avif_inputs_are_skipped(tmp_path):               # main.rs:136-156
    p = tmp_path / "already.avif"
    p.write_bytes(b"x")
    err = io.StringIO()
    assert cli.filter_inputs([str(p), "a.png", "-"], False, err) == ["a.png", "-"]
    assert "already an AVIF" in err.getvalue()
    err = io.StringIO()
    assert cli.filter_inputs([str(tmp_path / "typo.avif")], False, err) == [str(tmp_path / "typo.avif")]
    assert "Did you mean to use -o" in err.getvalue()


def test_summary_line_format():                           # main.rs:213
    assert cli.summary_line("o.avif", 12001, 11000, 500) == "o.avif: 13KB (11000B color, 500B alpha, 501B HEIF)"


def test_loader_matches_reference_conversions():          # main.rs:254-270
    rgb = Image.fromarray(np.full((3, 5, 3), 7, np.uint8)); b = io.BytesIO(); rgb.save(b, "PNG")
    a = cli.load_rgba(b.getvalue())
    assert a.shape == (3, 5, 4) and (a[..., 3] == 255).all() and (a[..., :3] == 7).all()
    g16 = Image.fromarray(np.full((2, 2), 0xABCD, np.uint16)); b = io.BytesIO(); g16.save(b, "PNG")
    a = cli.load_rgba(b.getvalue())
    assert (a[..., :3] == 0xAB).all() and (a[..., 3] == 255).all()
    la = Image.fromarray(np.dstack([np.full((2, 2), 9, np.uint8), np.full((2, 2), 100, np.uint8)]), "LA"); b = io.BytesIO(); la.save(b, "PNG")
    a = cli.load_rgba(b.getvalue())
    assert (a[..., :3] == 9).all() and (a[..., 3] == 100).all()


def test_existing_output_is_not_overwritten(tmp_path):    # main.rs:195-199 (no device work is reached)
    src = tmp_path / "a.png"
    Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(src)
    (tmp_path / "a.avif").write_bytes(b"keep")
    err = io.StringIO()
    assert cli.main([str(src)], stderr=err, stdout=io.StringIO()) == 1
    assert "already exists; skipping" in err.getvalue() and (tmp_path / "a.avif").read_bytes() == b"keep"


@pytest.mark.gpu
def test_cli_files_and_stdio(tmp_path):                   # tests/stdio.rs:6-44 + the per-file line of main.rs:213
    imgs = [H.synth_image(128, 85, 1), H.synth_image(96, 64, 2, alpha=True)]
    paths = []
    for i, im in enumerate(imgs):
        p = tmp_path / ("in%d.png" % i)
        Image.fromarray(im if i else im[..., :3]).save(p)
        paths.append(str(p))
    out = subprocess.run([sys.executable, "-m", "cavif_rs_b200", "--speed", "6", "-o", str(tmp_path / "out")] + paths,
                         cwd=H.ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 2 and lines[0].startswith(str(tmp_path / "out" / "in0.avif") + ": ") and "B color, 0B alpha" in lines[0]
    assert "B alpha" in lines[1] and "0B alpha" not in lines[1]
    for i in range(2):
        im = Image.open(tmp_path / "out" / ("in%d.avif" % i))
        assert im.size == (imgs[i].shape[1], imgs[i].shape[0])
    # stdin -> stdout
    r = subprocess.run([sys.executable, "-m", "cavif_rs_b200", "--speed", "10", "-"], cwd=H.ROOT, input=open(paths[0], "rb").read(),
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout[4:12] == b"ftypavif"
    # refusing to overwrite without -f, overwriting with it
    again = subprocess.run([sys.executable, "-m", "cavif_rs_b200", "-s", "10", "-o", str(tmp_path / "out")] + paths, cwd=H.ROOT,
                           capture_output=True, text=True)
    assert again.returncode == 1 and "already exists; skipping" in again.stderr
    forced = subprocess.run([sys.executable, "-m", "cavif_rs_b200", "-s", "10", "-f", "-q", "-o", str(tmp_path / "out")] + paths, cwd=H.ROOT,
                            capture_output=True, text=True)
    assert forced.returncode == 0 and forced.stdout == ""
