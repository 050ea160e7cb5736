"""Host-side logic of the library that needs no GPU: the copy pool behind the drop-in host path."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "copy_pool_stress.cpp"
BUILD = ROOT / "tests" / "_build"


def _build(flags, name, src=SRC):
    BUILD.mkdir(exist_ok=True)
    exe = BUILD / name
    res = subprocess.run(["g++", "-std=c++17", "-pthread", *flags, "-o", str(exe), str(src)], capture_output=True, text=True)
    return exe if res.returncode == 0 else None


def test_rgb_pack_rows():
    """Alpha-stripping staging copy (goofy_b200/csrc/rgb_pack.h): scalar and SSSE3 packers write exactly the R, G, B
    bytes of every pixel and nothing outside the packed row, at every alignment and width % 4 == 0."""
    exe = _build(["-O2"], "rgb_pack_host", ROOT / "tests" / "rgb_pack_host.cpp")
    assert exe is not None
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_hybrid_choice_policy():
    """When the host path packs large pinned images (goofy_b200/csrc/hybrid_choice.h): link-bound single process yes,
    shared host below link rate never, packing slower than plain only as a probe."""
    exe = _build(["-O2"], "hybrid_choice_host", ROOT / "tests" / "hybrid_choice_host.cpp")
    assert exe is not None
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_copy_pool_stress():
    """Spin-then-sleep wake-up handshake of CopyPool (goofy_b200/csrc/copy_pool.h): every copy (and every
    alpha-stripping pack) of 2000 jobs of varying size is checked, with pauses that let the workers fall asleep between jobs."""
    exe = _build(["-O2"], "copy_pool_stress")
    assert exe is not None
    out = subprocess.run([str(exe), "2000"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_copy_pool_under_thread_sanitizer():
    exe = _build(["-O1", "-g", "-fsanitize=thread"], "copy_pool_stress_tsan")
    if exe is None:
        pytest.skip("toolchain without ThreadSanitizer")
    out = subprocess.run([str(exe), "400"], capture_output=True, text=True, timeout=240)
    if "FATAL: ThreadSanitizer" in out.stderr and "unexpected memory mapping" in out.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (ASLR layout)")
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stderr[-2000:]
    assert out.stdout.startswith("ok")
