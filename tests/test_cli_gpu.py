"""vrender_b200 (the headless stand-in for the reference's `vrender`, built on the reference-named C++ host classes): the
flags and log lines scripts/benchmark.py of the reference depends on (scripts/benchmark.py:41-60), and the PNG screenshot."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "vkvolume_b200" / "lib" / "vrender_b200"


def run(*args):
    assert APP.exists(), "vrender_b200 is not built (python -m vkvolume_b200.build)"
    p = subprocess.run([str(APP), *args], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return p.stdout


@pytest.mark.parametrize("skipmode", [0, 1, 2, 3])
def test_benchmark_mode_log_lines_parse_with_the_reference_regexes(skipmode):
    out = run("--width=320", "--height=200", "--benchmark=5", "--imin=0.1", "--imax=1.0", "--gmin=0.0", "--gmax=0.2", "--blocksize=4",
              f"--skipmode={skipmode}", "synth:0:128x96x80")
    fps = re.search(r"ran [\d]+ frames, averaged ([\d\.]+) fps", out)
    upd = re.search(r"Updated occupancy/distance map in ([\d\.]+)ms", out)
    occ = re.search(r"Occupied voxels: ([\d\.]+)%", out)
    assert fps and upd and occ, out
    assert float(fps.group(1)) > 0 and 0.0 < float(occ.group(1)) < 100.0
    assert re.search(r"Updated gradient map in ([\d\.]+)ms", out)


def test_screenshot_png_and_gradient_test_flag(tmp_path):
    from PIL import Image
    png, ppm = tmp_path / "shot.png", tmp_path / "shot.ppm"
    common = ["--width=256", "--height=160", "--stop-after-frame=1", "--imin=0.1", "--gmin=0.0", "--gmax=0.2", "synth:0:96x96x64"]
    run(f"--screenshot-output={png}", *common)
    run(f"--screenshot-output={ppm}", *common)
    im = np.array(Image.open(png))
    assert im.shape == (160, 256, 4) and (im[..., 3] == 255).all()        # alpha forced to 255 (VS/framework/common/utils.cpp:141-175)
    raw = ppm.read_bytes()
    hdr = b"P6\n256 160\n255\n"
    assert raw.startswith(hdr)
    rgb = np.frombuffer(raw[len(hdr):], np.uint8).reshape(160, 256, 3)
    assert np.array_equal(im[..., :3], rgb) and rgb.max() > 0        # same frame through both writers, and something was drawn
    # --gradient_test: no precomputed gradient map; the frame agrees with the precomputed one to within the north-star bar
    png2 = tmp_path / "otf.png"
    run(f"--screenshot-output={png2}", "--gradient_test", *common)
    d = np.abs(np.array(Image.open(png2))[..., :3].astype(int) - im[..., :3].astype(int)).max(axis=2)
    assert (d <= 8).mean() > 0.97        # different estimator (filtered taps of V vs filtered 8-bit map): close, not identical
