"""Host-side image plumbing (no GPU): the asynchronous PNG writer of SURVEY 8f rank 4."""
import numpy as np
import pytest

from ceng795_b200 import image


def test_async_png_writer_round_trips_and_reports_errors(tmp_path):
    """Encode off the critical path: frames come back byte-identical, in order, and a failed
    encode is raised to the render loop instead of being lost."""
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (24, 32, 3), dtype=np.uint8) for _ in range(6)]
    with image.AsyncImageWriter(max_pending=2) as w:
        for i, f in enumerate(frames):
            w.submit(str(tmp_path / f"f{i}.png"), f)
    assert w.written == [str(tmp_path / f"f{i}.png") for i in range(6)]
    for i, f in enumerate(frames):
        assert np.array_equal(image.read_png(str(tmp_path / f"f{i}.png")), f)
    w = image.AsyncImageWriter()
    w.submit(str(tmp_path / "no_such_dir" / "x.png"), frames[0])
    with pytest.raises(Exception):
        w.close()


def test_cxx_async_writer_behind_the_dropin_mains(tmp_path):
    """integration/b200_host.h: the C++ twin of AsyncImageWriter that main_hw*_b200.cpp use — jobs in order on a worker
    thread, producer not blocked while the queue has room, bounded depth, first error code kept."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "async_writer_host")
    subprocess.run(["g++", "-std=c++14", "-O1", "-pthread", "-I", os.path.join(root, "include"), "-I",
                    os.path.join(root, "integration"), os.path.join(root, "tests", "async_writer_host.cpp"), "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.count("-> ok") == 3, r.stdout + r.stderr
