"""The CUDA block codec's arithmetic (goofy_b200/csrc/block_codec.cuh) compiled for the host with the
device intrinsics emulated, against the oracle.  This checks the kernel SOURCE on a machine without a
GPU; the GPU tests (-m gpu) check the compiled kernel itself."""
import numpy as np
import pytest

from oracle.oracle import DXT1, ETC1, splitmix_rgba, synth_family

CODECS = [DXT1, ETC1]
KEY = {DXT1: "dxt1", ETC1: "etc1"}


@pytest.mark.parametrize("codec", CODECS)
def test_kernel_math_on_fixtures(codec, kernel_math):
    fx = np.load("tests/golden/fixtures.npz")
    for n in sorted(k[:-5] for k in fx.files if k.endswith("_rgba")):
        img = fx[n + "_rgba"]
        h, w = img.shape[:2]
        rc, got = kernel_math(codec, img, w, h)
        assert rc == 0 and np.array_equal(got, fx[f"{n}_{KEY[codec]}"]), n
        rc, got = kernel_math(codec | 64, img, w, h)   # through the fused dual-output encoder
        assert rc == 0 and np.array_equal(got, fx[f"{n}_{KEY[codec]}"]), n


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("family", [0, 1, 2, 3])
def test_kernel_math_on_synthetic(codec, family, kernel_math, oracle):
    img = synth_family(family, 512, 512, seed=1234 + family)
    want = oracle.compress(codec, img, 512, 512)[1]
    rc, got = kernel_math(codec, img, 512, 512)
    assert rc == 0 and np.array_equal(got, want)
    rc, got = kernel_math(codec | 64, img, 512, 512)   # through the fused dual-output encoder
    assert rc == 0 and np.array_equal(got, want)
    # the packed-RGB kernels' way in (3 bytes per pixel, widened by widen_rgb24; byte 3 of every pixel word is junk)
    for flags in (128, 128 | 64):
        rc, got = kernel_math(codec | flags, img, 512, 512)
        assert rc == 0 and np.array_equal(got, want), flags
    rc, got = kernel_math(16 + codec + 128, img, 512, 512)   # float-reference flavour from packed RGB
    assert rc == 0 and np.array_equal(got, oracle.compress_float_reference(codec, img, 512, 512)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_kernel_math_low_contrast_sweep(codec, kernel_math, oracle):
    rng = np.random.default_rng(7)
    for spread in (0, 1, 2, 5, 8, 9, 21, 22, 44, 74, 106, 152, 182, 254, 255):
        base = rng.integers(0, 256 - spread, size=(64, 64, 1, 1, 3))
        blk = base + rng.integers(0, spread + 1, size=(64, 64, 4, 4, 3))
        img = np.zeros((256, 256, 4), dtype=np.uint8)
        img[..., :3] = blk.transpose(0, 2, 1, 3, 4).reshape(256, 256, 3)
        img[..., 3] = rng.integers(0, 256, size=(256, 256))
        rc, got = kernel_math(codec, img, 256, 256)
        assert rc == 0 and np.array_equal(got, oracle.compress(codec, img, 256, 256)[1]), spread
        rc2, got2 = kernel_math(codec | 64, img, 256, 256)
        assert rc2 == 0 and np.array_equal(got2, got), spread


@pytest.mark.parametrize("codec", CODECS)
def test_kernel_math_return_codes(codec, kernel_math):
    img = splitmix_rgba(64 * 64, seed=1)
    assert kernel_math(codec, img, 24, 32)[0] == -1
    assert kernel_math(codec, img, 32, 6)[0] == -2


@pytest.mark.parametrize("codec", CODECS)
def test_floatref_kernel_math(codec, kernel_math, oracle):
    """Integer re-derivation of the float reference (block_codec.cuh, float-reference flavour) vs the float oracle."""
    rng = np.random.default_rng(11)
    cases = [synth_family(f, 256, 256, seed=31 + f) for f in range(4)]
    for spread in (0, 1, 2, 7, 8, 9, 15, 16, 17, 20, 21, 22, 41, 42, 43, 72, 104, 150, 180, 252, 255):
        base = rng.integers(0, 256 - spread, size=(32, 32, 1, 1, 3))
        blk = base + rng.integers(0, spread + 1, size=(32, 32, 4, 4, 3))
        img = np.zeros((128, 128, 4), dtype=np.uint8)
        img[..., :3] = blk.transpose(0, 2, 1, 3, 4).reshape(128, 128, 3)
        cases.append(img)
    fx = np.load("tests/golden/fixtures.npz")
    cases += [fx[k] for k in fx.files if k.endswith("_rgba")]
    for img in cases:
        h, w = img.shape[:2]
        rc, got = kernel_math(16 + codec, img, w, h)
        assert rc == 0 and np.array_equal(got, oracle.compress_float_reference(codec, img, w, h)[1])
    assert kernel_math(16 + codec, splitmix_rgba(20 * 8, seed=1), 20, 8)[0] == 0
    assert kernel_math(16 + codec, splitmix_rgba(64 * 64, seed=1), 18, 8)[0] == -1


def test_kernel_math_property_based(kernel_math, oracle):
    """hypothesis-driven blocks: few distinct values, saturated channels, single-pixel outliers -- the shapes that
    hit ties (brightness == mid), the range clamp, the to5 edges and the ETC1 table steps."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    levels = st.sampled_from([0, 1, 7, 8, 9, 15, 16, 17, 127, 128, 129, 247, 248, 254, 255])
    pixel = st.tuples(st.one_of(levels, st.integers(0, 255)), st.one_of(levels, st.integers(0, 255)),
                      st.one_of(levels, st.integers(0, 255)), st.integers(0, 255))
    palette = st.lists(pixel, min_size=1, max_size=4)

    @settings(max_examples=300, deadline=None, derandomize=True)
    @given(pal=palette, picks=st.lists(st.integers(0, 3), min_size=64, max_size=64), jitter=st.integers(0, 3))
    def run(pal, picks, jitter):
        # a 16x4 tile = four blocks; every pixel is one of <= 4 palette colours, optionally jittered by +-jitter
        img = np.zeros((4, 16, 4), dtype=np.int64)
        for i, k in enumerate(picks):
            img[i // 16, i % 16] = pal[k % len(pal)]
        if jitter:
            img[..., :3] += (np.arange(4 * 16 * 3).reshape(4, 16, 3) * 7919 % (2 * jitter + 1)) - jitter
        img = np.clip(img, 0, 255).astype(np.uint8)
        for codec in CODECS:
            want = oracle.compress(codec, img, 16, 4)[1]
            assert np.array_equal(kernel_math(codec, img, 16, 4)[1], want)
            rc, got = kernel_math(codec | 64, img, 16, 4)
            assert rc == 0 and np.array_equal(got, want)
            assert np.array_equal(kernel_math(16 + codec, img, 16, 4)[1], oracle.compress_float_reference(codec, img, 16, 4)[1])

    run()


@pytest.mark.parametrize("codec", CODECS)
def test_decoder_math(codec, kernel_math, oracle):
    """block_decode.cuh (planar palettes + byte-permute lookups) against the oracle decoder: encoder output, and
    arbitrary blocks -- every BC1 mode, ETC1 individual / differential, both flips, wrapping differential deltas."""
    w, h = 256, 128
    for family in range(4):
        img = synth_family(family, w, h, seed=77 + family)
        blocks = oracle.compress(codec, img, w, h)[1]
        want = oracle.decode(codec, blocks, w, h)
        assert np.array_equal(kernel_math.decode(codec, blocks, w, h), want), family
        assert np.array_equal(kernel_math.sse(codec, blocks, img, w, h), oracle.sse_rgb(want, img)), family
    for seed in range(4):
        rnd = splitmix_rgba(w * h // 8, seed=500 + seed).reshape(-1)
        want = oracle.decode(codec, rnd, w, h)
        assert np.array_equal(kernel_math.decode(codec, rnd, w, h), want), seed
        img = synth_family(seed, w, h, seed=seed)
        assert np.array_equal(kernel_math.sse(codec, rnd, img, w, h), oracle.sse_rgb(want, img)), seed
    # BC1: equal endpoints and c0 < c1 (three-colour mode with transparent black) in every block
    if codec == DXT1:
        rnd = splitmix_rgba(w * h // 8, seed=9).reshape(-1, 8).copy()
        lo = np.minimum(rnd[:, 0:2].view(np.uint16), rnd[:, 2:4].view(np.uint16))
        hi = np.maximum(rnd[:, 0:2].view(np.uint16), rnd[:, 2:4].view(np.uint16))
        rnd[:, 0:2] = lo.view(np.uint8)
        rnd[:, 2:4] = hi.view(np.uint8)
        rnd[::7, 2:4] = rnd[::7, 0:2]
        assert np.array_equal(kernel_math.decode(codec, rnd.reshape(-1), w, h), oracle.decode(codec, rnd.reshape(-1), w, h))
