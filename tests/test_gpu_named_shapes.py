"""GPU parity at the shapes BASELINE.json names (configs[1]-[4]) and for the round-2 API additions, against the
unmodified reference run row-parallel on the host (oracle/_ref) -- or the scalar oracle where that was never built.
Multi-device cases are separate tests that SKIP on a single-GPU box instead of passing on N = 1.  Run with -m gpu."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import goofy_b200 as gb  # noqa: E402
from oracle.oracle import DXT1, ETC1, Reference, aligned_copy, aligned_empty, splitmix_rgba, synth_family  # noqa: E402

CODECS = [DXT1, ETC1]
LAYERS = ["LOAD_AUTO", "LOAD_DIRECT", "LOAD_ONESHOT", "LOAD_ASYNC", "LOAD_TMA"]
N_GPUS = gb.device_count()
need_two_gpus = pytest.mark.skipif(N_GPUS < 2, reason="needs at least two GPUs in this process")


@pytest.fixture(scope="module")
def checker(oracle):
    """want(codec, host_image, w, h, stride) -> blocks, from the reference (all host threads) or the oracle."""
    if Reference.available():
        ref = Reference()
        threads = ref.hardware_threads() or 1

        def want(codec, img, w, h, stride=None):
            stride = w * 4 if stride is None else stride
            rc, out = ref.compress_mt(codec, aligned_copy(img), w, h, stride, threads)
            assert rc == 0
            return out
    else:
        def want(codec, img, w, h, stride=None):
            rc, out = oracle.compress(codec, aligned_copy(img), w, h, stride)
            assert rc == 0
            return out
    return want


@pytest.fixture
def load_layer(request):
    prev = gb.set_load_path(getattr(gb, request.param))
    yield request.param
    gb.set_load_path(prev)


def dev(a, device=0):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda(device)


# ------------------------------------------------------------------------------------------- configs[4]
@pytest.mark.parametrize("load_layer", LAYERS, indirect=True)
def test_strip_16384x2048_stride_65792_every_load_layer(load_layer, checker):
    """One rank's share of BASELINE.json configs[4] at 8 GPUs: a 2048-row strip of a 16384-wide texture whose rows
    are 65 792 bytes apart, pad bytes 0xAB (they must be ignored).  DXT1 is what configs[4] names; ETC1s and the
    dual-output kernel ride along on the same strip."""
    w, h, stride = 16384, 2048, 16384 * 4 + 256
    host = aligned_empty(h * stride)
    host[:] = 0xAB
    rows = host.reshape(h, stride)
    tile = synth_family(1, w, 256, seed=41)           # photo-like, 256 rows at a time (bounded host memory)
    for y0 in range(0, h, 256):
        rows[y0:y0 + 256, : w * 4] = np.roll(tile.reshape(256, w * 4), 64 * (y0 // 256), axis=1)
    rows[512:768, : w * 4] = synth_family(0, w, 256, seed=42).reshape(256, w * 4)   # a band of uniform random
    rows[768:1024, : w * 4] = synth_family(2, w, 256, seed=43).reshape(256, w * 4)  # a band of 0 / 255
    d_src = dev(host)
    want = {c: checker(c, host, w, h, stride) for c in CODECS}
    for codec in CODECS:
        d_dst = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
        assert gb.encode_device(codec, d_dst, d_src, w, h, stride) == 0
        torch.cuda.synchronize()
        assert np.array_equal(d_dst.cpu().numpy(), want[codec]), (load_layer, codec)
    d_a = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
    assert gb.encode_dual_device(d_a, d_b, d_src, w, h, stride) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_a.cpu().numpy(), want[DXT1]) and np.array_equal(d_b.cpu().numpy(), want[ETC1]), load_layer
    # the scheduler's partition: eight strips of THIS strip, encoded separately, concatenate to the same bytes
    d_c = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
    for g in range(8):
        first, count = gb.strip_partition(h, 8, g)
        assert gb.encode_device(DXT1, d_c[first * (w // 4) * 8:], d_src[first * 4 * stride:], w, count * 4, stride) == 0
    torch.cuda.synchronize()
    assert torch.equal(d_c, d_a)


# ------------------------------------------------------------------------------------------- configs[3]
@pytest.mark.parametrize("pitch_pad", [0, 4096])
@pytest.mark.parametrize("load_layer", LAYERS, indirect=True)
def test_uniform_batch_64x1024_mixed_families(load_layer, pitch_pad, checker):
    """BASELINE.json configs[3] in small: 64 textures of 1024 x 1024 mixing the four synthetic families, as one
    uniform batch (back to back, and 4 KiB apart so that it is NOT one tall image), DXT1, ETC1s and dual-output."""
    w = h = 1024
    n = 64
    img_bytes, out_bytes = w * h * 4, w * h // 2
    pitch = img_bytes + pitch_pad
    host = np.full(n * pitch, 0xAB, dtype=np.uint8)
    distinct = [synth_family(f, w, h, seed=1000 + 7 * f + k) for f in range(4) for k in range(2)]
    for i in range(n):
        host[i * pitch: i * pitch + img_bytes] = np.roll(distinct[i % 8].reshape(-1), 16 * 4 * (i // 8))   # shifted by whole tiles
    d_src = dev(host)
    want = {c: [checker(c, host[i * pitch: i * pitch + img_bytes], w, h) for i in range(8 * 2)] for c in CODECS}
    for codec in CODECS:
        d_dst = torch.zeros(n * out_bytes, dtype=torch.uint8, device="cuda")
        assert gb.encode_batch_uniform_device(codec, d_dst, d_src, w, h, w * 4, pitch, out_bytes, n) == 0
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy().reshape(n, out_bytes)
        for i in range(16):
            assert np.array_equal(got[i], want[codec][i]), (load_layer, codec, i)
        # the rest of the batch against the GPU's own single-image path (already pinned above for 16 of them)
        d_one = torch.zeros(out_bytes, dtype=torch.uint8, device="cuda")
        for i in range(16, n, 5):
            assert gb.encode_device(codec, d_one, d_src[i * pitch:], w, h, w * 4) == 0
            torch.cuda.synchronize()
            assert np.array_equal(got[i], d_one.cpu().numpy()), (load_layer, codec, i)
    d_a = torch.zeros(n * out_bytes, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n * out_bytes, dtype=torch.uint8, device="cuda")
    assert gb.encode_dual_device(d_a, d_b, d_src, w, h, w * 4, pitch, out_bytes, n) == 0
    torch.cuda.synchronize()
    a, b = d_a.cpu().numpy().reshape(n, out_bytes), d_b.cpu().numpy().reshape(n, out_bytes)
    for i in range(16):
        assert np.array_equal(a[i], want[DXT1][i]) and np.array_equal(b[i], want[ETC1][i]), (load_layer, i)


def test_batch_pitches_must_cover_the_images():
    """A pitch of 0 (the default of the dual-output wrappers) or one smaller than an image is an error for n > 1,
    not a silent overlap."""
    w = h = 64
    d_src = torch.zeros(4 * w * h * 4, dtype=torch.uint8, device="cuda")
    d_dst = torch.zeros(4 * w * h // 2, dtype=torch.uint8, device="cuda")
    E_ARGS = -8
    assert gb.encode_batch_uniform_device(DXT1, d_dst, d_src, w, h, w * 4, 0, w * h // 2, 4) == E_ARGS
    assert gb.encode_batch_uniform_device(DXT1, d_dst, d_src, w, h, w * 4, w * h * 4, 0, 4) == E_ARGS
    assert gb.encode_batch_uniform_device(ETC1, d_dst, d_src, w, h, w * 4, w * h * 4 - 16, w * h // 2, 4) == E_ARGS
    assert gb.encode_dual_device(d_dst, d_dst, d_src, w, h, w * 4, 0, 0, 4) == E_ARGS
    assert gb.encode_batch_uniform_device(gb.DXT1_FLOATREF, d_dst, d_src, w, h, w * 4, 0, 0, 4) == E_ARGS
    assert gb.encode_dual_device(d_dst, d_dst[w * h // 2:], d_src, w, h, w * 4) == 0      # n = 1: pitches unused
    assert gb.encode_batch_uniform_device(DXT1, d_dst, d_src, w, h, w * 4, w * h * 4, w * h // 2, 4) == 0
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------- configs[1], [2]
@pytest.mark.parametrize("family", ["S0 uniform random", "S2 binary 0/255"])
def test_full_size_8192_worst_case_families(family, checker):
    """configs[1]/[2] at full size on the two stress families of SURVEY.md 8(d), generated on the device by torch's
    counter-based (Philox) generator: every block full-range (S0), every channel saturated (S2)."""
    size = 8192
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED + len(family))
    src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
    for y0 in range(0, size, 1024):
        if family.startswith("S0"):
            src[y0:y0 + 1024] = torch.randint(0, 256, (1024, size, 4), device="cuda", dtype=torch.int32, generator=g).to(torch.uint8)
        else:
            src[y0:y0 + 1024] = (torch.randint(0, 2, (1024, size, 4), device="cuda", dtype=torch.int32, generator=g) * 255).to(torch.uint8)
    host = aligned_copy(src.cpu().numpy())
    for codec in CODECS:
        dst = torch.zeros(size * size // 2, dtype=torch.uint8, device="cuda")
        assert gb.encode_device(codec, dst, src, size, size, size * 4) == 0
        torch.cuda.synchronize()
        assert np.array_equal(dst.cpu().numpy(), checker(codec, host, size, size)), (family, codec)


# ------------------------------------------------------------------------------------------- ragged batches
def ragged_inputs(seed0, floatref=False):
    shapes = [(16, 4), (64, 64), (272, 12), (1024, 8), (48, 100), (0, 0), (320, 36), (1040, 68)]
    if floatref:
        shapes += [(20, 8), (36, 36)]       # goofyRef:: accepts any width that is a multiple of 4
    imgs = [splitmix_rgba(max(w * h, 1), seed=seed0 + i)[: w * h * 4] for i, (w, h) in enumerate(shapes)]
    return shapes, imgs


def test_ragged_batch_both_codecs_in_one_launch(oracle):
    """GOOFY_B200_BOTH in goofy_b200_encode_batch_device: DXT1 to dst, ETC1s to dst2, inline and uploaded tables."""
    for repeat in (1, 8):     # 8 x 8 = 64 descriptors: more than travel as kernel parameters
        shapes, imgs = ragged_inputs(500)
        shapes, imgs = shapes * repeat, imgs * repeat
        keep, d1, d2, descs = [], [], [], []
        for (w, h), img in zip(shapes, imgs):
            keep.append(dev(img) if w else torch.zeros(16, dtype=torch.uint8, device="cuda"))
            d1.append(torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda"))
            d2.append(torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda"))
            descs.append((keep[-1], d1[-1], w, h, w * 4, -1, d2[-1]))
        assert gb.encode_batch_device(gb.BOTH, descs) == 0
        torch.cuda.synchronize()
        for (w, h), img, a, b in zip(shapes, imgs, d1, d2):
            if w:
                assert np.array_equal(a.cpu().numpy()[: w * h // 2], oracle.compress(DXT1, img, w, h)[1]), (w, h)
                assert np.array_equal(b.cpu().numpy()[: w * h // 2], oracle.compress(ETC1, img, w, h)[1]), (w, h)
    # a null dst2 is an error, before anything is launched
    w, h = 64, 64
    s, d = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda"), torch.full((w * h // 2,), 7, dtype=torch.uint8, device="cuda")
    assert gb.encode_batch_device(gb.BOTH, [(s, d, w, h, w * 4)]) == -3
    torch.cuda.synchronize()
    assert bool((d == 7).all())


@pytest.mark.parametrize("codec", CODECS)
def test_ragged_batch_float_reference_flavour(codec, oracle):
    flav = {DXT1: gb.DXT1_FLOATREF, ETC1: gb.ETC1_FLOATREF}[codec]
    shapes, imgs = ragged_inputs(700, floatref=True)
    keep, dsts, descs = [], [], []
    for (w, h), img in zip(shapes, imgs):
        stride = (w * 4 + 15) // 16 * 16      # rows stay 16-byte aligned for widths that are only multiples of 4
        padded = np.zeros((max(h, 1), max(stride, 16)), dtype=np.uint8)
        if w:
            padded[:, : w * 4] = img.reshape(h, w * 4)
        keep.append(dev(padded))
        dsts.append(torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda"))
        descs.append((keep[-1], dsts[-1], w, h, stride))
    assert gb.encode_batch_device(flav, descs) == 0
    torch.cuda.synchronize()
    for (w, h), img, d in zip(shapes, imgs, dsts):
        if w:
            assert np.array_equal(d.cpu().numpy()[: w * h // 2], oracle.compress_float_reference(codec, img, w, h)[1]), (w, h)
    assert gb.encode_batch_device(codec, [(keep[-1], dsts[-1], 20, 8, 80)]) == -1     # the SSE2-exact flavour still wants w % 16


def test_host_batch_both_codecs(oracle):
    shapes, imgs = ragged_inputs(900)
    pinned = torch.empty(1040 * 68 // 2, dtype=torch.uint8).pin_memory()
    items, outs = [], []
    for i, ((w, h), img) in enumerate(zip(shapes, imgs)):
        a = np.zeros(max(w * h // 2, 8), dtype=np.uint8)
        b = pinned if (w, h) == (1040, 68) else np.zeros(max(w * h // 2, 8), dtype=np.uint8)
        outs.append((a, b))
        items.append((aligned_copy(img) if w else np.zeros(16, dtype=np.uint8), a, w, h, w * 4, b))
    assert gb.encode_host_batch(gb.BOTH, items) == 0
    for (w, h), img, (a, b) in zip(shapes, imgs, outs):
        if w:
            b = b.numpy() if hasattr(b, "numpy") else b
            assert np.array_equal(a[: w * h // 2], oracle.compress(DXT1, img, w, h)[1]), (w, h)
            assert np.array_equal(b[: w * h // 2], oracle.compress(ETC1, img, w, h)[1]), (w, h)
    assert gb.encode_host_batch(gb.BOTH, [(aligned_copy(imgs[1]), outs[1][0], 64, 64, 256)]) == -3    # no second result


# ------------------------------------------------------------------------------------------- host scratch
def test_short_lived_threads_recycle_the_host_scratch(oracle):
    """A thread per texture (the natural way to replace the multi-threaded CPU reference) must not keep one set of
    streams + device strips + pinned strips per thread that ever called: sets are leased from a pool."""
    w, h = 256, 64
    img = synth_family(1, w, h)
    want = oracle.compress(DXT1, img, w, h)[1]
    src = aligned_copy(img)
    errors = []

    def call():
        out = np.zeros(w * h // 2, dtype=np.uint8)
        if gb.compressDXT1(out, src, w, h, w * 4) != 0 or not np.array_equal(out, want):
            errors.append("mismatch")

    call()
    before = gb.host_scratch_sets()
    for _ in range(40):                     # forty threads, one after the other
        t = threading.Thread(target=call)
        t.start()
        t.join()
    assert not errors
    assert gb.host_scratch_sets() - before <= 1      # the main thread holds its lease; the forty share one set
    ts = [threading.Thread(target=call) for _ in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors
    assert gb.host_scratch_sets() - before <= 6      # at most as many sets as threads inside the library at once


# ------------------------------------------------------------------------------------------- more than one GPU
@need_two_gpus
@pytest.mark.parametrize("codec", CODECS)
def test_multi_gpu_sharded_host_strips(codec, checker):
    w, h = 2048, 1000
    img = synth_family(1, w, h, seed=77)
    want = checker(codec, img, w, h)
    src = aligned_copy(img)
    for n in range(1, N_GPUS + 1):
        out = np.zeros(w * h // 2, dtype=np.uint8)
        assert gb.encode_sharded_host(codec, out, src, w, h, w * 4, n) == 0
        assert np.array_equal(out, want), n
    assert gb.encode_sharded_host(codec, np.zeros(w * h // 2, dtype=np.uint8), src, w, h, w * 4, N_GPUS + 1) == -7


@need_two_gpus
def test_multi_gpu_dual_sharded_host(checker):
    w, h = 2048, 1000
    img = synth_family(0, w, h, seed=78)
    src = aligned_copy(img)
    for n in (2, N_GPUS):
        a, b = np.zeros(w * h // 2, dtype=np.uint8), np.zeros(w * h // 2, dtype=np.uint8)
        assert gb.encode_dual_sharded_host(a, b, src, w, h, w * 4, n) == 0
        assert np.array_equal(a, checker(DXT1, img, w, h)) and np.array_equal(b, checker(ETC1, img, w, h)), n


@need_two_gpus
@pytest.mark.parametrize("codec", [DXT1, ETC1, gb.BOTH])
def test_multi_gpu_sharded_batch(codec, checker):
    """goofy_b200_encode_batch_sharded: textures resident on different devices, one host thread per device."""
    w, h = 512, 256
    imgs = [synth_family(i % 4, w, h, seed=90 + i) for i in range(3 * N_GPUS + 1)]
    descs, outs = [], []
    for i, img in enumerate(imgs):
        d = i % N_GPUS
        s = dev(img.reshape(-1), d)
        a = torch.zeros(w * h // 2, dtype=torch.uint8, device=f"cuda:{d}")
        b = torch.zeros(w * h // 2, dtype=torch.uint8, device=f"cuda:{d}")
        outs.append((s, a, b))
        descs.append((s, a, w, h, w * 4, d, b))
    assert gb.encode_batch_sharded(codec, descs) == 0
    for img, (_, a, b) in zip(imgs, outs):
        if codec == gb.BOTH:
            assert np.array_equal(a.cpu().numpy(), checker(DXT1, img, w, h)) and np.array_equal(b.cpu().numpy(), checker(ETC1, img, w, h))
        else:
            assert np.array_equal(a.cpu().numpy(), checker(codec, img, w, h))
    # a descriptor naming a device that does not exist is refused before anything starts
    bad = list(descs[0])
    bad[5] = N_GPUS
    assert gb.encode_batch_sharded(DXT1, [tuple(bad)]) == -7


@need_two_gpus
def test_large_ragged_batch_on_two_devices_from_one_thread(oracle):
    """More descriptors than travel as kernel parameters, first on GPU 0 and then on GPU 1, from the same host thread:
    the descriptor arena (and its event) must follow the device."""
    w, h = 64, 32
    n = 64
    imgs = [splitmix_rgba(w * h, seed=1200 + i) for i in range(n)]
    for d in (0, 1, 0, 1):
        with torch.cuda.device(d):
            keep = [dev(img, d) for img in imgs]
            dsts = [torch.zeros(w * h // 2, dtype=torch.uint8, device=f"cuda:{d}") for _ in imgs]
            assert gb.encode_batch_device(ETC1, [(s, t, w, h, w * 4) for s, t in zip(keep, dsts)]) == 0
            torch.cuda.synchronize(d)
            for img, t in zip(imgs, dsts):
                assert np.array_equal(t.cpu().numpy(), oracle.compress(ETC1, img, w, h)[1]), d
