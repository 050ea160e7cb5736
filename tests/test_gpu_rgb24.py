"""Packed-RGB input (goofy_b200_encode_rgb24_device) and the alpha-stripping host path built on it: the bytes must be
those the RGBA path produces, i.e. those of the oracle / the unmodified reference on the same pixels with any alpha
(the encoders ignore alpha, GoofyTC/goofy_tc.h:297)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import goofy_b200 as gb  # noqa: E402
from oracle.oracle import DXT1, ETC1, aligned_copy, image_names, load_test_image, splitmix_rgba, synth_family  # noqa: E402

CODECS = [DXT1, ETC1]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rgb_of(img, w, h, stride=None, pad=0xCD):
    """RGBA (h, w, 4) -> packed RGB rows `stride` bytes apart (default tight), padding filled with `pad`."""
    stride = w * 3 if stride is None else stride
    out = np.full((h, stride), pad, dtype=np.uint8)
    out[:, : w * 3] = np.ascontiguousarray(img).reshape(h, w, 4)[..., :3].reshape(h, w * 3)
    return out


def rgb24_device(codec, img, w, h, stride=None, base_offset=0):
    stride = w * 3 if stride is None else stride
    rows = rgb_of(img, w, h, stride).reshape(-1)
    d_all = torch.zeros(rows.size + base_offset + 64, dtype=torch.uint8, device="cuda")
    d_all[base_offset: base_offset + rows.size] = dev(rows)
    d_src = d_all[base_offset:]
    d_dst = torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda")
    d_dst2 = torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda")
    rc = gb.encode_rgb24_device(codec, d_dst, d_src, w, h, stride, d_result2=d_dst2 if codec == gb.BOTH else None)
    torch.cuda.synchronize()
    return rc, d_dst.cpu().numpy()[: w * h // 2], d_dst2.cpu().numpy()[: w * h // 2]


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("family", [0, 1, 2, 3])
def test_rgb24_synthetic_families_vs_oracle(codec, family, oracle):
    for (w, h, seed) in ((256, 256, 1), (1024, 512, 77), (2064, 36, 5)):
        img = synth_family(family, w, h, seed=seed)
        rc, got, _ = rgb24_device(codec, img, w, h)
        assert rc == 0 and "rgb24" in gb.last_launch_kernel()
        assert np.array_equal(got, oracle.compress(codec, img, w, h)[1]), (family, w, h)


@pytest.mark.parametrize("shape", [(16, 4), (16, 8), (32, 4), (48, 12), (112, 20), (144, 4), (528, 36), (1040, 8), (4112, 4), (8192, 16)])
def test_rgb24_ragged_shapes_all_outputs(shape, oracle):
    """Every width class: warps with fewer than 32 live blocks, CTAs stacking several block rows, both codecs from one read."""
    w, h = shape
    img = splitmix_rgba(w * h, seed=w * 31 + h)
    want = {c: oracle.compress(c, img, w, h)[1] for c in CODECS}
    for codec in CODECS:
        rc, got, _ = rgb24_device(codec, img, w, h)
        assert rc == 0 and np.array_equal(got, want[codec]), (codec, shape)
    rc, got, got2 = rgb24_device(gb.BOTH, img, w, h)
    assert rc == 0 and np.array_equal(got, want[DXT1]) and np.array_equal(got2, want[ETC1]), shape


@pytest.mark.parametrize("codec", CODECS + [gb.BOTH])
def test_rgb24_padded_and_unaligned_rows(codec, oracle):
    """Row stride beyond width*3 (padding never read into the result); strides / bases that are only 4-byte aligned."""
    w, h = 320, 64
    img = synth_family(1, w, h)
    want = oracle.compress(DXT1 if codec == gb.BOTH else codec, img, w, h)[1]
    want2 = oracle.compress(ETC1, img, w, h)[1]
    for stride, base in ((w * 3 + 256, 0), (w * 3 + 4, 0), (w * 3, 4), (w * 3 + 52, 8)):
        rc, got, got2 = rgb24_device(codec, img, w, h, stride, base)
        assert rc == 0 and np.array_equal(got, want), (stride, base)
        if codec == gb.BOTH:
            assert np.array_equal(got2, want2), (stride, base)


def test_rgb24_float_reference_flavour(reference):
    """goofyRef-exact flavour from packed RGB, including widths that are only multiples of 4."""
    for (w, h) in ((256, 64), (20, 8), (1028, 12)):
        img = synth_family(1, w, h, seed=w)
        for codec, ref_codec in ((gb.DXT1_FLOATREF, DXT1), (gb.ETC1_FLOATREF, ETC1)):
            rc, got, _ = rgb24_device(codec, img, w, h)
            assert rc == 0 and np.array_equal(got, reference.compress_float_reference(ref_codec, aligned_copy(img), w, h)[1]), (w, h, codec)


def test_rgb24_uniform_batch_and_argument_checks(oracle):
    w, h, n = 128, 64, 5
    imgs = [synth_family(i % 4, w, h, seed=i + 1) for i in range(n)]
    pitch = w * 3 * h + 48
    buf = np.zeros(n * pitch, dtype=np.uint8)
    for i, im in enumerate(imgs):
        buf[i * pitch: i * pitch + w * 3 * h] = rgb_of(im, w, h).reshape(-1)
    d_src = dev(buf)
    out_pitch = w * h // 2 + 64
    d_a = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
    gb.check(gb.encode_rgb24_device(gb.BOTH, d_a, d_src, w, h, w * 3, d_result2=d_b, input_image_pitch=pitch,
                                    result_image_pitch=out_pitch, n_images=n))
    torch.cuda.synchronize()
    a, b = d_a.cpu().numpy(), d_b.cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(a[i * out_pitch: i * out_pitch + w * h // 2], oracle.compress(DXT1, im, w, h)[1]), i
        assert np.array_equal(b[i * out_pitch: i * out_pitch + w * h // 2], oracle.compress(ETC1, im, w, h)[1]), i
    # the codec's own shape rules, then the packed-row rules
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, 24, 32, 72) == -1
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, 32, 30, 96) == -2
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, 0, 0, 0) == 0
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, 32, 32, 95) == -5      # stride < width*3
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, 32, 32, 98) == -4      # stride % 4
    assert gb.encode_rgb24_device(DXT1, d_a, 0, 32, 32, 96) == -3
    assert gb.encode_rgb24_device(gb.BOTH, d_a, d_src, 32, 32, 96) == -3   # no second result
    assert gb.encode_rgb24_device(7, d_a, d_src, 32, 32, 96) == -6
    assert gb.encode_rgb24_device(DXT1, d_a, d_src, w, h, w * 3, input_image_pitch=w * 3, result_image_pitch=out_pitch, n_images=2) == -8


def test_rgb24_all_test_images(oracle):
    names = image_names()
    if not names:
        pytest.skip("oracle/_ref/test-data not present")
    for n in names[::3]:
        img = load_test_image(n)
        h, w = img.shape[:2]
        for codec in CODECS:
            rc, got, _ = rgb24_device(codec, img, w, h)
            assert rc == 0 and np.array_equal(got, oracle.compress(codec, img, w, h)[1]), n


# ------------------------------------------------------------------ host path: alpha-stripped staging
@pytest.fixture()
def rgb_mode():
    before = gb.get_host_rgb_staging()
    yield
    gb.set_host_rgb_staging(before)


def _pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).pin_memory()
    return t


@pytest.mark.parametrize("mode", [gb.HOST_RGB_OFF, gb.HOST_RGB_AUTO, gb.HOST_RGB_ALWAYS, gb.HOST_RGB_PAGEABLE])
def test_host_path_same_bytes_in_every_staging_mode(mode, rgb_mode, reference):
    """Pageable and pinned buffers, small (zero-copy) and large (hybrid: raw strips from the front, alpha-stripped strips
    from the back) images, padded stride, random alpha: every mode returns the reference's bytes."""
    assert gb.set_host_rgb_staging(mode) in (0, 1, 2, 3)
    assert gb.set_host_rgb_staging(7) == -8 and gb.get_host_rgb_staging() == mode
    assert gb.get_host_rgb_staging() == mode
    cases = [(768, 512, 0), (2048, 2048, 0), (4096, 3072, 512), (8192, 2048, 0)]
    for (w, h, pad) in cases:
        stride = w * 4 + pad
        tight = splitmix_rgba(w * h, seed=w + h + mode).reshape(h, w * 4)
        padded = np.full((h, stride), 0xAB, dtype=np.uint8)
        padded[:, : w * 4] = tight
        want = {c: reference.compress_mt(c, aligned_copy(padded), w, h, stride, 8)[1] for c in CODECS}
        before = gb.host_link_stats()
        for codec, fn in ((DXT1, gb.compressDXT1), (ETC1, gb.compressETC1)):
            out = np.zeros(w * h // 2, dtype=np.uint8)
            assert fn(out, aligned_copy(padded), w, h, stride) == 0           # pageable
            assert np.array_equal(out, want[codec]), (mode, w, h, codec, "pageable")
            t_in, t_out = _pinned(padded), torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
            assert fn(t_out, t_in, w, h, stride) == 0                         # pinned
            assert np.array_equal(t_out.numpy(), want[codec]), (mode, w, h, codec, "pinned")
        # both codecs from one upload
        a, b = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory(), torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
        assert gb.encode_dual_host(a, b, _pinned(padded), w, h, stride) == 0
        assert np.array_equal(a.numpy(), want[DXT1]) and np.array_equal(b.numpy(), want[ETC1]), (mode, w, h, "dual pinned")
        a2, b2 = np.zeros(w * h // 2, dtype=np.uint8), np.zeros(w * h // 2, dtype=np.uint8)
        assert gb.encode_dual_host(a2, b2, aligned_copy(padded), w, h, stride) == 0
        assert np.array_equal(a2, want[DXT1]) and np.array_equal(b2, want[ETC1]), (mode, w, h, "dual pageable")
        after = gb.host_link_stats()
        sent = after["bytes_uploaded"] - before["bytes_uploaded"]
        full = 6 * w * h * 4       # six host calls
        if mode == gb.HOST_RGB_OFF:
            assert sent == full and after["packed_strips"] == before["packed_strips"]
        else:
            assert sent < full     # pageable calls always strip the alpha byte
        if mode == gb.HOST_RGB_PAGEABLE:
            assert after["packed_strips"] == before["packed_strips"] and after["packing_calls"] == before["packing_calls"]
            assert sent == 6 * w * h * 4 - 3 * w * h     # the three pageable calls strip the alpha byte, the pinned ones do not
        if mode == gb.HOST_RGB_ALWAYS and w * h * 4 > 32 << 20:
            assert after["raw_strips"] == before["raw_strips"] and after["packed_strips"] > before["packed_strips"]


def test_host_path_float_reference_and_batch_with_rgb_staging(rgb_mode, reference):
    gb.set_host_rgb_staging(gb.HOST_RGB_AUTO)
    for (w, h) in ((20, 8), (1028, 64), (4096, 2560)):
        img = aligned_copy(synth_family(1, w, h, seed=w + 3))
        for fn, codec in ((gb.goofyRef.compressDXT1, DXT1), (gb.goofyRef.compressETC1, ETC1)):
            want = reference.compress_float_reference(codec, img, w, h)[1]
            out = np.zeros(w * h // 2, dtype=np.uint8)
            assert fn(out, img, w, h, w * 4) == 0
            assert np.array_equal(out, want), (w, h, codec, "pageable")
            t_out = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
            assert fn(t_out, _pinned(img), w, h, w * 4) == 0
            assert np.array_equal(t_out.numpy(), want), (w, h, codec, "pinned")
    # a host batch mixing small pageable images and one large pinned image (which takes the hybrid scheduler mid-batch)
    shapes = [(256, 128), (4096, 2304), (768, 512), (64, 16)]
    imgs = [aligned_copy(synth_family(i % 4, w, h, seed=i + 9)) for i, (w, h) in enumerate(shapes)]
    srcs = [aligned_copy(im) if i != 1 else _pinned(im) for i, im in enumerate(imgs)]
    outs = [np.zeros(w * h // 2, dtype=np.uint8) if i != 1 else torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
            for i, (w, h) in enumerate(shapes)]
    outs2 = [np.zeros(w * h // 2, dtype=np.uint8) if i != 1 else torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
             for i, (w, h) in enumerate(shapes)]
    gb.check(gb.encode_host_batch(gb.BOTH, [(s, o, w, h, w * 4, o2) for s, o, o2, (w, h) in zip(srcs, outs, outs2, shapes)]))
    for im, o, o2, (w, h) in zip(imgs, outs, outs2, shapes):
        o = o.numpy() if hasattr(o, "numpy") else o
        o2 = o2.numpy() if hasattr(o2, "numpy") else o2
        assert np.array_equal(o, reference.compress(DXT1, im, w, h)[1]) and np.array_equal(o2, reference.compress(ETC1, im, w, h)[1]), (w, h)


def test_hybrid_scheduler_splits_the_image(rgb_mode, reference):
    """A large pinned image with packing forced (ALWAYS) and under AUTO (which packs only where its own measurements say
    it pays: the first two calls of a thread are plain DMA): the result is the reference's either way; the split itself
    depends on the host and is only reported."""
    w, h = 8192, 8192
    img = torch.from_numpy(synth_family(1, w, h, seed=11).reshape(-1)).pin_memory()
    want = reference.compress_mt(DXT1, aligned_copy(img.numpy()), w, h, w * 4, 16)[1]
    for mode, calls in ((gb.HOST_RGB_AUTO, 5), (gb.HOST_RGB_ALWAYS, 1)):
        gb.set_host_rgb_staging(mode)
        before = gb.host_link_stats()
        for _ in range(calls):
            out = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
            assert gb.compressDXT1(out, img, w, h, w * 4) == 0
            assert np.array_equal(out.numpy(), want), mode
        after = gb.host_link_stats()
        d = {k: after[k] - before[k] for k in after}
        assert d["packing_calls"] + d["plain_calls"] == calls
        assert d["bytes_uploaded"] <= calls * w * h * 4
        if mode == gb.HOST_RGB_ALWAYS:
            assert d["packing_calls"] == 1 and d["raw_strips"] == 0 and d["packed_strips"] >= 16
            assert d["bytes_uploaded"] == w * h * 3
        else:
            assert d["plain_calls"] >= 2
        print(f"mode {mode}: {d}, {gb.host_threads()} host threads")
    # the scheduler polls events: "not ready" must not be left behind as the calling thread's last CUDA error
    import ctypes as C
    from goofy_b200 import _lib
    _lib.load()
    rt = None
    for line in open("/proc/self/maps"):
        if "libcudart.so" in line and "torch" not in line:
            rt = C.CDLL(line.split()[-1])   # the runtime libgoofy_b200.so is linked against (already loaded)
            break
    if rt is not None:
        assert rt.cudaPeekAtLastError() == 0
    assert float((torch.zeros(4, device="cuda") + 1).sum()) == 4.0


def test_rgb24_host_call(reference):
    """goofy_b200_encode_rgb24_host: packed-RGB HOST images, pageable and pinned, small (zero-copy) and large (copy-engine
    strips), padded rows, one and both codecs, float-reference flavour: the reference's bytes for the same pixels."""
    for (w, h, pad) in ((768, 512, 0), (320, 64, 52), (2048, 2048, 0), (4096, 3072, 256)):
        img = splitmix_rgba(w * h, seed=3 * w + h).reshape(h, w, 4)
        stride = w * 3 + pad
        rows = rgb_of(img, w, h, stride)
        want = {c: reference.compress_mt(c, aligned_copy(img), w, h, w * 4, 8)[1] for c in CODECS}
        for codec in CODECS:
            out = np.zeros(w * h // 2, dtype=np.uint8)
            assert gb.encode_rgb24_host(codec, out, rows.reshape(-1), w, h, stride) == 0          # pageable
            assert np.array_equal(out, want[codec]), (w, h, codec, "pageable")
            t_out = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
            assert gb.encode_rgb24_host(codec, t_out, _pinned(rows), w, h, stride) == 0            # pinned
            assert np.array_equal(t_out.numpy(), want[codec]), (w, h, codec, "pinned")
        a, b = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory(), np.zeros(w * h // 2, dtype=np.uint8)
        before = gb.host_link_stats()["bytes_uploaded"]
        assert gb.encode_rgb24_host(gb.BOTH, a, _pinned(rows), w, h, stride, result2=b) == 0
        assert gb.host_link_stats()["bytes_uploaded"] - before == w * h * 3
        assert np.array_equal(a.numpy(), want[DXT1]) and np.array_equal(b, want[ETC1]), (w, h, "both")
    w, h = 1028, 64
    img = aligned_copy(synth_family(1, w, h, seed=5)).reshape(h, w, 4)
    out = np.zeros(w * h // 2, dtype=np.uint8)
    assert gb.encode_rgb24_host(gb.ETC1_FLOATREF, out, rgb_of(img, w, h).reshape(-1), w, h, w * 3) == 0
    assert np.array_equal(out, reference.compress_float_reference(ETC1, img, w, h)[1])
    # argument checks: the codec's shape rules first, then the packed-row rules
    buf = np.zeros(64 * 64 * 3, dtype=np.uint8)
    assert gb.encode_rgb24_host(DXT1, out, buf, 24, 32, 72) == -1
    assert gb.encode_rgb24_host(DXT1, out, buf, 32, 30, 96) == -2
    assert gb.encode_rgb24_host(DXT1, out, buf, 0, 0, 0) == 0
    assert gb.encode_rgb24_host(DXT1, out, buf, 32, 32, 92) == -5
    assert gb.encode_rgb24_host(DXT1, out, buf, 32, 32, 98) == -4
    assert gb.encode_rgb24_host(DXT1, out, None, 32, 32, 96) == -3
    assert gb.encode_rgb24_host(gb.BOTH, out, buf, 32, 32, 96) == -3
    assert gb.encode_rgb24_host(9, out, buf, 32, 32, 96) == -6


def test_host_calls_from_concurrent_threads_with_packing(rgb_mode, reference):
    """Four host threads inside the host path at once, every one on its own large pinned image with packing forced: the
    copy pool serialises their staging jobs, every thread has its own pack ring, events and device strips."""
    import threading
    gb.set_host_rgb_staging(gb.HOST_RGB_ALWAYS)
    w, h = 4096, 2304
    imgs = [aligned_copy(synth_family(i % 4, w, h, seed=40 + i)) for i in range(4)]
    pinned = [_pinned(im) for im in imgs]
    outs = [torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory() for _ in imgs]
    pageable_outs = [np.zeros(w * h // 2, dtype=np.uint8) for _ in imgs]
    want = [reference.compress_mt(DXT1 if i % 2 == 0 else ETC1, imgs[i], w, h, w * 4, 8)[1] for i in range(4)]
    errors = []

    def work(i):
        fn = gb.compressDXT1 if i % 2 == 0 else gb.compressETC1
        try:
            for _ in range(3):
                if fn(outs[i], pinned[i], w, h, w * 4) != 0:
                    errors.append((i, "pinned rc"))
                if fn(pageable_outs[i], imgs[i], w, h, w * 4) != 0:      # pageable input: packed while it is staged
                    errors.append((i, "pageable rc"))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        assert np.array_equal(outs[i].numpy(), want[i]), i
        assert np.array_equal(pageable_outs[i], want[i]), i


def test_host_neighbours_sees_another_process():
    """AUTO asks NVML whether some GPU of the box runs somebody else's compute process (this process's own contexts do
    not count): alone -> 0; with a second process holding a context -> at least 1, and AUTO then leaves pinned input to
    plain DMA; after it has gone -> 0 again (the answer is refreshed every few seconds)."""
    import subprocess
    import sys
    import time
    torch.zeros(1, device="cuda")
    n0 = gb.host_neighbours()
    if n0 < 0:
        pytest.skip("NVML not available: the host path falls back to its rate gate")
    if n0 != 0:
        pytest.skip(f"{n0} GPU(s) of this box already run somebody else's compute process: the test needs the box to itself")
    child = subprocess.Popen([sys.executable, "-c", "import torch, sys, time; torch.zeros(1, device='cuda'); print('up', flush=True); time.sleep(60)"],
                             stdout=subprocess.PIPE, text=True)
    try:
        assert child.stdout.readline().strip() == "up"
        deadline = time.time() + 10
        while gb.host_neighbours() == 0 and time.time() < deadline:
            time.sleep(0.5)
        assert gb.host_neighbours() >= 1
        before = gb.host_link_stats()
        gb.set_host_rgb_staging(gb.HOST_RGB_AUTO)
        w, h = 8192, 2048
        img = torch.from_numpy(synth_family(1, w, h, seed=3).reshape(-1)).pin_memory()
        out = torch.zeros(w * h // 2, dtype=torch.uint8).pin_memory()
        for _ in range(6):
            assert gb.compressDXT1(out, img, w, h, w * 4) == 0
        after = gb.host_link_stats()
        assert after["packing_calls"] == before["packing_calls"] and after["plain_calls"] - before["plain_calls"] == 6
    finally:
        child.kill()
        child.wait()
    deadline = time.time() + 10
    while gb.host_neighbours() != 0 and time.time() < deadline:
        time.sleep(0.5)
    assert gb.host_neighbours() == 0
