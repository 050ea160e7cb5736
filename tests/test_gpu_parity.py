"""GPU parity tests: the CUDA path, called through the C ABI, must be bit-exact with the oracle
(and with the unmodified reference where oracle/_ref is present).  Run with -m gpu on a B200."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import goofy_b200 as gb  # noqa: E402
from oracle.oracle import (DXT1, ETC1, aligned_copy, load_test_image, splitmix_rgba, synth_family,  # noqa: E402
                           image_names)

CODECS = [DXT1, ETC1]
HOST_FN = {DXT1: gb.compressDXT1, ETC1: gb.compressETC1}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_device(codec, img, w, h, stride=None):
    stride = w * 4 if stride is None else stride
    d_src = dev(np.ascontiguousarray(img).reshape(-1))
    d_dst = torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda")
    rc = gb.encode_device(codec, d_dst, d_src, w, h, stride)
    torch.cuda.synchronize()
    return rc, d_dst.cpu().numpy()[: w * h // 2]


def gpu_host(codec, img, w, h, stride=None):
    stride = w * 4 if stride is None else stride
    out = np.zeros(w * h // 2, dtype=np.uint8)
    rc = HOST_FN[codec](out, aligned_copy(img), w, h, stride)
    return rc, out


def test_library_sees_a_gpu():
    assert gb.device_count() >= 1
    assert "B200" in torch.cuda.get_device_name(0) or torch.cuda.get_device_capability(0)[0] >= 10


@pytest.mark.parametrize("codec", CODECS)
def test_golden_fixtures(codec):
    """Committed fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
    fx = np.load("tests/golden/fixtures.npz")
    key = {DXT1: "dxt1", ETC1: "etc1"}[codec]
    names = sorted(k[:-5] for k in fx.files if k.endswith("_rgba"))
    assert len(names) >= 39
    for n in names:
        img = fx[n + "_rgba"]
        h, w = img.shape[:2]
        rc, got = gpu_device(codec, img, w, h)
        assert rc == 0
        assert np.array_equal(got, fx[f"{n}_{key}"]), n
        rc, got = gpu_host(codec, img, w, h)
        assert rc == 0 and np.array_equal(got, fx[f"{n}_{key}"]), n


@pytest.mark.parametrize("codec", CODECS)
def test_all_test_images_match_golden_hashes_and_oracle(codec, oracle, golden):
    names = image_names()
    if not names:
        pytest.skip("oracle/_ref/test-data not present")
    key = {DXT1: "dxt1", ETC1: "etc1"}[codec]
    for n in names:
        img = load_test_image(n)
        h, w = img.shape[:2]
        rc, got = gpu_device(codec, img, w, h)
        assert rc == 0
        assert hashlib.sha256(got.tobytes()).hexdigest() == golden["images"][n][key]["sha256"], n
        assert np.array_equal(got, oracle.compress(codec, img, w, h)[1]), n


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("family", [0, 1, 2, 3])
def test_synthetic_families_vs_oracle(codec, family, oracle, golden):
    img = synth_family(family, 256, 256)
    rc, got = gpu_device(codec, img, 256, 256)
    assert rc == 0
    key = {DXT1: "dxt1", ETC1: "etc1"}[codec]
    assert hashlib.sha256(got.tobytes()).hexdigest() == golden["synthetic"][f"family{family}_256"][key]["sha256"]
    img = synth_family(family, 1024, 512, seed=77)
    rc, got = gpu_device(codec, img, 1024, 512)
    assert rc == 0 and np.array_equal(got, oracle.compress(codec, img, 1024, 512)[1])


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("shape", [(16, 4), (16, 8), (32, 4), (48, 12), (112, 20), (144, 4), (528, 36), (1040, 8), (4112, 4)])
def test_ragged_shapes(codec, shape, oracle):
    w, h = shape
    img = splitmix_rgba(w * h, seed=w * 131 + h)
    rc, got = gpu_device(codec, img, w, h)
    assert rc == 0 and np.array_equal(got, oracle.compress(codec, img, w, h)[1])
    rc, got = gpu_host(codec, img, w, h)
    assert rc == 0 and np.array_equal(got, oracle.compress(codec, img, w, h)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_padded_stride_and_pad_bytes_ignored(codec, oracle):
    w, h, pad = 320, 64, 256
    stride = w * 4 + pad
    tight = synth_family(1, w, h)
    padded = np.full((h, stride), 0xAB, dtype=np.uint8)
    padded[:, : w * 4] = tight.reshape(h, w * 4)
    want = oracle.compress(codec, tight, w, h)[1]
    for runner in (gpu_device, gpu_host):
        rc, got = runner(codec, padded, w, h, stride)
        assert rc == 0 and np.array_equal(got, want)
    padded[:, w * 4:] = 0x11
    rc, got = gpu_device(codec, padded, w, h, stride)
    assert rc == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("codec", CODECS)
def test_alpha_is_ignored(codec):
    img = synth_family(0, 128, 64)
    a = img.copy(); a[..., 3] = 255
    b = img.copy(); b[..., 3] = 0
    assert np.array_equal(gpu_device(codec, a, 128, 64)[1], gpu_device(codec, b, 128, 64)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_return_codes_match_reference(codec):
    """-1 width%16, -2 height%4 in that order, 0 for empty images (goofy_tc.h:1500-1508); output untouched on error."""
    buf = torch.full((4096,), 0x5A, dtype=torch.uint8, device="cuda")
    src = torch.zeros(64 * 64 * 4, dtype=torch.uint8, device="cuda")
    assert gb.encode_device(codec, buf, src, 24, 32, 96) == -1
    assert gb.encode_device(codec, buf, src, 32, 6, 128) == -2
    assert gb.encode_device(codec, buf, src, 24, 6, 96) == -1
    assert gb.encode_device(codec, buf, src, 0, 0, 0) == 0
    assert gb.encode_device(codec, buf, src, 0, 8, 0) == 0
    assert gb.encode_device(codec, buf, src, 32, 32, 64) == -5
    assert gb.encode_device(codec, buf, src, 32, 32, 136) == -4
    assert gb.encode_device(codec, buf, 0, 32, 32, 128) == -3
    assert gb.encode_device(codec, buf, src.data_ptr() + 4, 32, 32, 128) == -4
    assert gb.encode_device(7, buf, src, 32, 32, 128) == -6
    torch.cuda.synchronize()
    assert bool((buf == 0x5A).all())
    out = np.full(512, 0x5A, dtype=np.uint8)
    img = np.zeros(32 * 32 * 4, dtype=np.uint8)
    assert HOST_FN[codec](out, img, 24, 32, 96) == -1
    assert HOST_FN[codec](out, img, 32, 30, 128) == -2
    assert HOST_FN[codec](out, img, 0, 0, 0) == 0
    assert (out == 0x5A).all()


@pytest.mark.parametrize("codec", CODECS)
def test_uniform_batch_and_dual(codec, oracle):
    n, w, h = 5, 192, 64
    imgs = np.stack([synth_family(i % 4, w, h, seed=i) for i in range(n)])
    d_src = dev(imgs)
    d_dst = torch.zeros((n, w * h // 2), dtype=torch.uint8, device="cuda")
    rc = gb.encode_batch_uniform_device(codec, d_dst, d_src, w, h, w * 4, w * h * 4, w * h // 2, n)
    torch.cuda.synchronize()
    assert rc == 0
    for i in range(n):
        assert np.array_equal(d_dst[i].cpu().numpy(), oracle.compress(codec, imgs[i], w, h)[1])
    if codec == DXT1:
        d_a = torch.zeros((n, w * h // 2), dtype=torch.uint8, device="cuda")
        d_b = torch.zeros((n, w * h // 2), dtype=torch.uint8, device="cuda")
        rc = gb.encode_dual_device(d_a, d_b, d_src, w, h, w * 4, w * h * 4, w * h // 2, n)
        torch.cuda.synchronize()
        assert rc == 0
        for i in range(n):
            assert np.array_equal(d_a[i].cpu().numpy(), oracle.compress(DXT1, imgs[i], w, h)[1])
            assert np.array_equal(d_b[i].cpu().numpy(), oracle.compress(ETC1, imgs[i], w, h)[1])


@pytest.mark.parametrize("shape", [(16, 4), (272, 12), (1040, 68), (4112, 36), (64, 1028)])
def test_dual_output_shapes_strides_and_pitches(shape, oracle):
    """Both codecs from one read (AUTO: row-walking CTAs through the cp.async ring; pitched batches: one-shot CTAs):
    ragged widths and heights, a padded stride whose pad bytes must be ignored, and images that are not back to back."""
    w, h = shape
    stride = w * 4 + 48
    n = 3
    pitch = stride * h + 256           # not back to back -> the pitched one-shot launch
    out_pitch = w * h // 2 + 64
    rng = np.random.default_rng(w + h)
    host = rng.integers(0, 256, size=n * pitch, dtype=np.uint8)
    imgs = []
    for i in range(n):
        rows = host[i * pitch: i * pitch + stride * h].reshape(h, stride)
        imgs.append(np.ascontiguousarray(rows[:, : w * 4]).reshape(-1))
    d_src = dev(host)
    d_a = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
    assert gb.encode_dual_device(d_a, d_b, d_src, w, h, stride, pitch, out_pitch, n) == 0
    # the first image alone: a single (tall, if h is large) image through the AUTO row-walking launch
    d_a1 = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
    d_b1 = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
    assert gb.encode_dual_device(d_a1, d_b1, d_src, w, h, stride) == 0
    torch.cuda.synchronize()
    a, b = d_a.cpu().numpy(), d_b.cpu().numpy()
    for i in range(n):
        want_d = oracle.compress(DXT1, imgs[i], w, h)[1]
        want_e = oracle.compress(ETC1, imgs[i], w, h)[1]
        assert np.array_equal(a[i * out_pitch: i * out_pitch + w * h // 2], want_d), (shape, i)
        assert np.array_equal(b[i * out_pitch: i * out_pitch + w * h // 2], want_e), (shape, i)
        assert not a[i * out_pitch + w * h // 2: (i + 1) * out_pitch].any()      # gaps between results untouched
    assert np.array_equal(d_a1.cpu().numpy(), oracle.compress(DXT1, imgs[0], w, h)[1])
    assert np.array_equal(d_b1.cpu().numpy(), oracle.compress(ETC1, imgs[0], w, h)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_ragged_batch_descriptors(codec, oracle):
    shapes = [(16, 4), (64, 64), (272, 12), (1024, 8), (48, 100), (0, 0), (320, 36)]
    imgs, srcs, dsts, descs = [], [], [], []
    for i, (w, h) in enumerate(shapes):
        img = splitmix_rgba(max(w * h, 1), seed=100 + i)[: w * h * 4]
        imgs.append(img)
        srcs.append(dev(img) if w else torch.zeros(16, dtype=torch.uint8, device="cuda"))
        dsts.append(torch.zeros(max(w * h // 2, 8), dtype=torch.uint8, device="cuda"))
        descs.append((srcs[-1], dsts[-1], w, h, w * 4))
    rc = gb.encode_batch_device(codec, descs)
    torch.cuda.synchronize()
    assert rc == 0
    for (w, h), img, d in zip(shapes, imgs, dsts):
        if w:
            assert np.array_equal(d.cpu().numpy()[: w * h // 2], oracle.compress(codec, img, w, h)[1]), (w, h)


@pytest.mark.parametrize("codec", CODECS)
def test_ragged_batch_large_table_and_tall_tiles(codec, oracle):
    """More images than travel as kernel parameters (the descriptor table is uploaded instead), with heights on
    both sides of the 16-block-row tile, plus a mip chain whose levels are sub-rectangles of one texture."""
    rng = np.random.default_rng(5)
    shapes = [(16 * int(rng.integers(1, 9)), 4 * int(rng.integers(1, 40))) for _ in range(60)] + [(1040, 68), (16, 260)]
    imgs, dsts, descs, keep = [], [], [], []
    for i, (w, h) in enumerate(shapes):
        img = splitmix_rgba(w * h, seed=300 + i)
        imgs.append(img)
        keep.append(dev(img))
        dsts.append(torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda"))
        descs.append((keep[-1], dsts[-1], w, h, w * 4))
    assert gb.encode_batch_device(codec, descs) == 0
    torch.cuda.synchronize()
    for (w, h), img, d in zip(shapes, imgs, dsts):
        assert np.array_equal(d.cpu().numpy(), oracle.compress(codec, img, w, h)[1]), (w, h)
    # mip chain: level k is the top-left (size >> k)^2 corner of one 256x256 texture, pitch of the base level
    base = synth_family(1, 256, 256, seed=3)
    d_base = dev(base)
    outs, chain, s = [], [], 256
    while s >= 16:
        outs.append(torch.zeros(s * s // 2, dtype=torch.uint8, device="cuda"))
        chain.append((d_base, outs[-1], s, s, 256 * 4))
        s //= 2
    assert gb.encode_batch_device(codec, chain) == 0
    torch.cuda.synchronize()
    for (_, out, s, _, _) in chain:
        want = oracle.compress(codec, np.ascontiguousarray(base.reshape(256, 256, 4)[:s, :s]), s, s)[1]
        assert np.array_equal(out.cpu().numpy(), want), s


@pytest.mark.parametrize("codec", CODECS)
def test_sharded_entry_points_on_one_device(codec, oracle):
    """The shard scheduler with a single shard / a single device (the multi-device cases are separate tests in
    tests/test_gpu_named_shapes.py that skip on a one-GPU box)."""
    w, h = 512, 200
    img = synth_family(1, w, h)
    want = oracle.compress(codec, img, w, h)[1]
    out = np.zeros(w * h // 2, dtype=np.uint8)
    assert gb.encode_sharded_host(codec, out, aligned_copy(img), w, h, w * 4, 1) == 0
    assert np.array_equal(out, want)
    descs, dsts = [], []
    for i in range(6):
        s = torch.from_numpy(img.reshape(-1)).cuda(0)
        t = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda:0")
        dsts.append(t)
        descs.append((s, t, w, h, w * 4, 0))
    assert gb.encode_batch_sharded(codec, descs) == 0
    for t in dsts:
        assert np.array_equal(t.cpu().numpy(), want)


@pytest.mark.parametrize("codec", CODECS)
def test_full_size_8192_properties(codec, oracle, reference):
    """BASELINE.json configs[1]/[2] at full size: bit-exact vs the unmodified reference run row-parallel on the
    host, strip-concatenation equivalence, determinism, and a decode/PSNR sanity check."""
    size = 8192
    src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
    from bench import fill_texture_device
    fill_texture_device(torch, src, seed=5)
    dst = torch.zeros(size * size // 2, dtype=torch.uint8, device="cuda")
    assert gb.encode_device(codec, dst, src, size, size, size * 4) == 0
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    host = aligned_copy(src.cpu().numpy())
    rc, want = reference.compress_mt(codec, host, size, size, size * 4, reference.hardware_threads())
    assert rc == 0 and np.array_equal(got, want)
    # strips of whole block rows, encoded separately, concatenate to the same bytes (the multi-GPU partition)
    dst2 = torch.zeros_like(dst)
    for g in range(8):
        first, count = gb.strip_partition(size, 8, g)
        off_src = first * 4 * size * 4
        off_dst = first * (size // 4) * 8
        assert gb.encode_device(codec, dst2[off_dst:], src.view(-1)[off_src:], size, count * 4, size * 4) == 0
    torch.cuda.synchronize()
    assert torch.equal(dst, dst2)
    # decode a 512-row band and check quality is in the encoder's known range
    band = 512
    dec = oracle.decode(codec, got[: size * band // 2], size, band)
    p = __import__("oracle.oracle", fromlist=["psnr_from_sse"]).psnr_from_sse(oracle.sse_rgb(dec, host[: size * band * 4]), size * band)
    # (the synthetic texture carries independent per-channel noise, which ETC1s' single base colour cannot follow)
    assert (30.0 if codec == DXT1 else 24.0) < p["psnr_rgb768"] < 60.0


@pytest.fixture
def tma_path():
    prev = gb.set_load_path(gb.LOAD_TMA)
    yield
    gb.set_load_path(prev)


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("shape", [(16, 4), (64, 8), (1024, 4), (1040, 12), (2048, 64), (3088, 20), (4096, 516), (272, 1028)])
def test_tma_tile_path_shapes(codec, shape, oracle, tma_path):
    """The TMA 2D-tile load layer (persistent CTAs, mbarrier ring) produces the same bytes as the oracle,
    including widths that are not a multiple of the 1024-pixel tile (hardware zero-fill of the ragged edge)."""
    w, h = shape
    img = splitmix_rgba(w * h, seed=w * 7 + h)
    launches = gb.kernel_launches()
    rc, got = gpu_device(codec, img, w, h)
    assert rc == 0 and gb.kernel_launches() == launches + 1
    assert np.array_equal(got, oracle.compress(codec, img, w, h)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_tma_tile_path_padded_stride_batch_and_dual(codec, oracle, tma_path):
    n, w, h, pad = 7, 1296, 36, 256
    stride = w * 4 + pad
    pitch = stride * h + 4096
    buf = np.full(n * pitch, 0xAB, dtype=np.uint8)
    imgs = []
    for i in range(n):
        img = synth_family(i % 4, w, h, seed=40 + i)
        imgs.append(img)
        view = buf[i * pitch: i * pitch + stride * h].reshape(h, stride)
        view[:, : w * 4] = img.reshape(h, w * 4)
    d_src = dev(buf)
    out_pitch = w * h // 2 + 64
    d_dst = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
    assert gb.encode_batch_uniform_device(codec, d_dst, d_src, w, h, stride, pitch, out_pitch, n) == 0
    torch.cuda.synchronize()
    got = d_dst.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i * out_pitch: i * out_pitch + w * h // 2], oracle.compress(codec, imgs[i], w, h)[1]), i
        assert not got[i * out_pitch + w * h // 2: (i + 1) * out_pitch].any()
    if codec == DXT1:
        d_a = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
        d_b = torch.zeros(n * out_pitch, dtype=torch.uint8, device="cuda")
        assert gb.encode_dual_device(d_a, d_b, d_src, w, h, stride, pitch, out_pitch, n) == 0
        torch.cuda.synchronize()
        for i in range(n):
            assert np.array_equal(d_a.cpu().numpy()[i * out_pitch: i * out_pitch + w * h // 2], oracle.compress(DXT1, imgs[i], w, h)[1])
            assert np.array_equal(d_b.cpu().numpy()[i * out_pitch: i * out_pitch + w * h // 2], oracle.compress(ETC1, imgs[i], w, h)[1])


def test_tma_and_direct_agree_on_a_large_texture():
    size = 4096
    src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
    from bench import fill_texture_device
    fill_texture_device(torch, src, seed=11)
    outs = {}
    for path in (gb.LOAD_DIRECT, gb.LOAD_TMA):
        prev = gb.set_load_path(path)
        try:
            for codec in CODECS:
                d = torch.zeros(size * size // 2, dtype=torch.uint8, device="cuda")
                assert gb.encode_device(codec, d, src, size, size, size * 4) == 0
                torch.cuda.synchronize()
                outs[(path, codec)] = d
        finally:
            gb.set_load_path(prev)
    for codec in CODECS:
        assert torch.equal(outs[(gb.LOAD_DIRECT, codec)], outs[(gb.LOAD_TMA, codec)])


@pytest.mark.parametrize("codec", CODECS)
def test_gpu_decoder_and_sse_match_oracle(codec, oracle):
    """Row N2 of SURVEY.md 8(f): BC1 / ETC1 decode and the squared-error reduction on the device,
    against the oracle decoder (itself checked against Src/decoder.cpp in tests/test_oracle.py)."""
    w, h = 512, 256
    img = synth_family(1, w, h)
    blocks = oracle.compress(codec, img, w, h)[1]
    d_blocks, d_src = dev(blocks), dev(img)
    d_out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    assert gb.decode_device(codec, d_out, d_blocks, w, h, w * 4) == 0
    torch.cuda.synchronize()
    want = oracle.decode(codec, blocks, w, h)
    assert np.array_equal(d_out.cpu().numpy(), want)
    d_sse = torch.zeros(3, dtype=torch.int64, device="cuda")
    assert gb.block_sse_device(codec, d_blocks, d_src, w, h, w * 4, d_sse) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_sse.cpu().numpy().astype(np.float64), oracle.sse_rgb(want, img))
    # arbitrary blocks (all BC1 modes; ETC1 individual + differential without overflow, both flips)
    rnd = splitmix_rgba(w * h // 8, seed=99).reshape(-1, 8).copy()
    if codec == ETC1:
        diff = (rnd[:, 3] & 2) != 0
        rnd[diff, 0:3] &= 0xF8
    rnd = rnd.reshape(-1)
    assert gb.decode_device(codec, d_out, dev(rnd), w, h, w * 4) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), oracle.decode(codec, rnd, w, h))


def test_gpu_psnr_of_test_images_matches_reference_numbers(oracle):
    """Encode -> decode -> PSNR entirely on the device; Kodak means measured with the reference's own
    decoder are 37.218 (DXT1) / 36.519 (ETC1s) dB."""
    names = [n for n in image_names() if n.startswith("kodim")]
    if len(names) != 24:
        pytest.skip("Kodak images not present")
    for codec, want in ((DXT1, 37.218), (ETC1, 36.519)):
        vals = []
        for n in names:
            img = load_test_image(n)
            h, w = img.shape[:2]
            d_src = dev(img)
            d_blk = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
            d_sse = torch.zeros(3, dtype=torch.int64, device="cuda")
            assert gb.encode_device(codec, d_blk, d_src, w, h, w * 4) == 0
            assert gb.block_sse_device(codec, d_blk, d_src, w, h, w * 4, d_sse) == 0
            torch.cuda.synchronize()
            vals.append(gb.psnr_rgb768(d_sse.cpu().tolist(), w * h))
        assert abs(float(np.mean(vals)) - want) < 0.01, (codec, np.mean(vals))


@pytest.mark.parametrize("codec", CODECS)
def test_floatref_flavour_bit_exact_vs_goofyref(codec, oracle):
    """GOOFY_B200_*_FLOATREF: bit-exact with goofyRef:: (Src/goofy_tc_reference.cpp) -- checked against our float
    restatement everywhere and against the unmodified reference where oracle/_ref is present."""
    from oracle.oracle import Reference
    ref = Reference() if Reference.available() else None
    fcodec = {DXT1: gb.DXT1_FLOATREF, ETC1: gb.ETC1_FLOATREF}[codec]
    host_fn = {DXT1: gb.goofyRef.compressDXT1, ETC1: gb.goofyRef.compressETC1}[codec]
    cases = [synth_family(f, 512, 256, seed=3 + f) for f in range(4)]
    cases.append(splitmix_rgba(20 * 12, seed=8).reshape(12, 20, 4))     # width % 4 only
    cases.append(splitmix_rgba(1036 * 8, seed=9).reshape(8, 1036, 4))
    for n in image_names():
        cases.append(load_test_image(n))
    for img in cases:
        h, w = img.shape[:2]
        want = oracle.compress_float_reference(codec, img, w, h)[1]
        if ref is not None:
            assert np.array_equal(want, ref.compress_float_reference(codec, aligned_copy(img), w, h)[1])
        rc, got = gpu_device(fcodec, img, w, h)
        assert rc == 0 and np.array_equal(got, want), (w, h)
        out = np.zeros(w * h // 2, dtype=np.uint8)
        assert host_fn(out, aligned_copy(img), w, h, w * 4) == 0 and np.array_equal(out, want)
    # and it is a different result from the SSE2-exact flavour, as surveyed
    img = cases[1]
    assert not np.array_equal(gpu_device(fcodec, img, 512, 256)[1], gpu_device(codec, img, 512, 256)[1])
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    src = torch.zeros(64 * 64 * 4, dtype=torch.uint8, device="cuda")
    assert gb.encode_device(fcodec, buf, src, 18, 8, 80) == -1
    assert gb.encode_device(fcodec, buf, src, 20, 6, 80) == -2
    # uniform batch with free-form pitches
    n, w, h = 3, 64, 16
    imgs = np.stack([synth_family(i, w, h, seed=70 + i) for i in range(n)])
    d_dst = torch.zeros((n, w * h // 2 + 8), dtype=torch.uint8, device="cuda")
    assert gb.encode_batch_uniform_device(fcodec, d_dst, dev(imgs), w, h, w * 4, w * h * 4, w * h // 2 + 8, n) == 0
    torch.cuda.synchronize()
    for i in range(n):
        assert np.array_equal(d_dst[i].cpu().numpy()[: w * h // 2], oracle.compress_float_reference(codec, imgs[i], w, h)[1])


def test_images_of_4_gib_and_more_use_64_bit_offsets():
    """Maximum sizes: a 16384 x 65544 image (4.3 GB) takes the WIDE kernels; its output must equal the
    same image encoded as two strips (each below 4 GiB, 32-bit offsets) -- the strip-partition property."""
    free, _ = torch.cuda.mem_get_info()
    w, h = 16384, 65544
    need = w * h * 4 + 2 * (w * h // 2) + (1 << 30)
    if free < need:
        pytest.skip("not enough device memory")
    src = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
    gen = torch.Generator(device="cuda"); gen.manual_seed(17)
    for y0 in range(0, h, 4096):
        y1 = min(h, y0 + 4096)
        src[y0:y1] = torch.randint(0, 256, (y1 - y0, w, 4), device="cuda", dtype=torch.uint8, generator=gen)
    for codec in CODECS:
        whole = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
        parts = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
        assert gb.encode_device(codec, whole, src, w, h, w * 4) == 0
        h0 = 32768
        assert gb.encode_device(codec, parts, src, w, h0, w * 4) == 0
        assert gb.encode_device(codec, parts[(h0 // 4) * (w // 4) * 8:], src[h0:], w, h - h0, w * 4) == 0
        torch.cuda.synchronize()
        assert torch.equal(whole, parts)
        del whole, parts


def test_concurrent_host_calls_from_threads(oracle):
    """The drop-in entry points are re-entrant like the reference (pure function, no init call): four host
    threads encode different images at once through their own thread-local staging pipes."""
    import threading
    jobs = []
    for i in range(8):
        w, h = 256 + 64 * (i % 3), 128 + 32 * (i % 4)
        img = synth_family(i % 4, w, h, seed=500 + i)
        codec = CODECS[i % 2]
        jobs.append((codec, w, h, aligned_copy(img), oracle.compress(codec, img, w, h)[1]))
    results = [None] * len(jobs)

    def work(k):
        codec, w, h, img, _ = jobs[k]
        out = np.zeros(w * h // 2, dtype=np.uint8)
        for _ in range(5):
            rc = HOST_FN[codec](out, img, w, h, w * 4)
            if rc != 0:
                break
        results[k] = (rc, out)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for (codec, w, h, img, want), (rc, out) in zip(jobs, results):
        assert rc == 0 and np.array_equal(out, want)


@pytest.mark.parametrize("codec", CODECS)
def test_stream_order_is_respected_between_launches(codec, oracle):
    """Back-to-back launches on one stream behave exactly like serial execution (the kernels use programmatic
    dependent launch to hide launch latency, which must not relax ordering): a later encode into the same
    destination wins, and an encode sees pixels written by the kernel just before it."""
    w, h = 2048, 1024
    a, b = synth_family(0, w, h, seed=1), synth_family(1, w, h, seed=2)
    want_b = oracle.compress(codec, b, w, h)[1]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d_a, d_b = dev(a), dev(b)
        dst = torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda")
        scratch = torch.empty_like(d_a)
        for _ in range(10):
            assert gb.encode_device(codec, dst, d_a, w, h, w * 4) == 0      # WAW: overwritten below
            assert gb.encode_device(codec, dst, d_b, w, h, w * 4) == 0
        got1 = dst.clone()
        for _ in range(10):
            scratch.copy_(d_a)                                               # RAW: the encode must see the copy
            assert gb.encode_device(codec, dst, scratch, w, h, w * 4) == 0
            scratch.copy_(d_b)
            assert gb.encode_device(codec, dst, scratch, w, h, w * 4) == 0
        got2 = dst.clone()
    stream.synchronize()
    assert np.array_equal(got1.cpu().numpy(), want_b)
    assert np.array_equal(got2.cpu().numpy(), want_b)


@pytest.mark.parametrize("codec", CODECS)
def test_random_shapes_strides_and_alignments(codec, oracle):
    """Seeded sweep over widths (multiples of 16), heights (multiples of 4), paddings (multiples of 16) and
    16-byte-aligned base offsets, device and host entry points."""
    rng = np.random.default_rng(4242 + codec)
    for _ in range(40):
        w = 16 * int(rng.integers(1, 40))
        h = 4 * int(rng.integers(1, 40))
        pad = 16 * int(rng.integers(0, 9))
        off = 16 * int(rng.integers(0, 5))
        stride = w * 4 + pad
        tight = rng.integers(0, 256, size=(h, w * 4), dtype=np.uint8)
        buf = np.full(off + h * stride, 0xEE, dtype=np.uint8)
        view = buf[off:].reshape(h, stride)
        view[:, : w * 4] = tight
        want = oracle.compress(codec, tight, w, h)[1]
        d_buf = dev(buf)
        d_dst = torch.zeros(w * h // 2 + 16, dtype=torch.uint8, device="cuda")
        assert gb.encode_device(codec, d_dst[8:], d_buf[off:], w, h, stride) == 0   # 8-byte aligned output
        torch.cuda.synchronize()
        assert np.array_equal(d_dst.cpu().numpy()[8: 8 + w * h // 2], want), (w, h, pad, off)
        assert not d_dst.cpu().numpy()[:8].any() and not d_dst.cpu().numpy()[8 + w * h // 2:].any()
        host = aligned_copy(buf)
        out = np.zeros(w * h // 2, dtype=np.uint8)
        assert HOST_FN[codec](out, host[off:], w, h, stride) == 0 and np.array_equal(out, want)


def test_device_entry_points_are_cuda_graph_capturable(oracle):
    """Launch-bound loops over many small textures can be captured once and replayed: the device entry
    points only enqueue kernels on the given stream (no allocation, no synchronisation)."""
    w, h, n = 256, 256, 12
    imgs = [synth_family(i % 4, w, h, seed=900 + i) for i in range(n)]
    d_src = [dev(im) for im in imgs]
    d_dst = [torch.zeros(w * h // 2, dtype=torch.uint8, device="cuda") for _ in range(n)]
    assert gb.encode_device(DXT1, d_dst[0], d_src[0], w, h, w * 4) == 0   # one-time device set-up happens outside capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(n):
            assert gb.encode_device(CODECS[i % 2], d_dst[i], d_src[i], w, h, w * 4) == 0
    for t in d_dst:
        t.zero_()
    launches = gb.kernel_launches()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert gb.kernel_launches() == launches          # replays launch from the graph, not through the library
    for i in range(n):
        assert np.array_equal(d_dst[i].cpu().numpy(), oracle.compress(CODECS[i % 2], imgs[i], w, h)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_host_api_pinned_and_pageable_buffers(codec, oracle):
    """Pinned buffers are DMA'd in place; pageable ones are staged through pinned strips by the library's copy
    threads.  Large enough for several strips (48 MiB in), padded stride, all four pinned/pageable combinations."""
    w, h, pad = 4096, 3072, 512
    stride = w * 4 + pad
    tight = synth_family(1, w, h, seed=77)
    want = oracle.compress(codec, tight, w, h)[1]
    padded = np.full((h, stride), 0xAB, dtype=np.uint8)
    padded[:, : w * 4] = tight.reshape(h, w * 4)
    pageable_in = aligned_copy(padded)
    pinned_in = torch.empty(padded.size, dtype=torch.uint8).pin_memory()
    pinned_in.numpy()[:] = padded.reshape(-1)
    for src in (pageable_in, pinned_in):
        for pinned_out in (False, True):
            out = torch.zeros(w * h // 2, dtype=torch.uint8)
            out = out.pin_memory() if pinned_out else out
            assert HOST_FN[codec](out, src, w, h, stride) == 0
            assert np.array_equal(out.numpy(), want)


@pytest.mark.parametrize("shape", [(16, 4, 0), (768, 512, 0), (1040, 68, 32), (4096, 2052, 0)])
@pytest.mark.parametrize("pinned", [(False, False, False), (True, True, True), (False, True, False), (True, False, True)])
def test_dual_output_host_call(shape, pinned, oracle):
    """goofy_b200_encode_dual_host: both codecs from one upload, every mix of pageable / pinned buffers (input, DXT1
    result, ETC1s result), one block row up to several strips, padded stride."""
    from oracle.oracle import aligned_copy
    w, h, pad = shape
    stride = w * 4 + pad
    img = splitmix_rgba(w * h, seed=w + h)
    rows = np.full((h, stride), 0x5A, dtype=np.uint8)
    rows[:, : w * 4] = img.reshape(h, w * 4)

    def buf(n, pin, fill=None):
        if pin:
            t = torch.zeros(n, dtype=torch.uint8).pin_memory()
            if fill is not None:
                t.numpy()[:] = fill
            return t
        return aligned_copy(fill) if fill is not None else np.zeros(n, dtype=np.uint8)

    src = buf(rows.size, pinned[0], rows.reshape(-1))
    d1, d2 = buf(w * h // 2, pinned[1]), buf(w * h // 2, pinned[2])
    assert gb.encode_dual_host(d1, d2, src, w, h, stride) == 0
    g1 = d1.numpy() if hasattr(d1, "numpy") else d1
    g2 = d2.numpy() if hasattr(d2, "numpy") else d2
    assert np.array_equal(g1, oracle.compress(DXT1, img, w, h)[1])
    assert np.array_equal(g2, oracle.compress(ETC1, img, w, h)[1])


@pytest.mark.parametrize("codec", CODECS)
def test_host_batch_mixed_buffers_and_shapes(codec, oracle):
    """goofy_b200_encode_host_batch: many host images through one pipeline -- pageable and pinned buffers mixed,
    padded strides, images from one block row to several strips, an empty image in the middle."""
    shapes = [(16, 4, 0), (768, 512, 0), (272, 12, 64), (2048, 1024, 0), (64, 64, 0), (0, 0, 0), (1040, 68, 16), (4096, 1028, 0)]
    items, keep, wants = [], [], []
    for i, (w, h, pad) in enumerate(shapes):
        stride = w * 4 + pad
        img = splitmix_rgba(max(w * h, 1), seed=700 + i)[: w * h * 4]
        if w:
            rows = np.full((h, stride), 0xAB, dtype=np.uint8)
            rows[:, : w * 4] = img.reshape(h, w * 4)
        else:
            rows = np.zeros((1, 16), dtype=np.uint8)
        if i % 2:   # pinned input / pinned output for every other image
            src = torch.from_numpy(rows.reshape(-1).copy()).pin_memory()
            dst = torch.zeros(max(w * h // 2, 8), dtype=torch.uint8).pin_memory()
        else:
            from oracle.oracle import aligned_copy
            src = aligned_copy(rows.reshape(-1))
            dst = np.zeros(max(w * h // 2, 8), dtype=np.uint8)
        keep.append((src, dst))
        items.append((src, dst, w, h, stride))
        wants.append(oracle.compress(codec, img, w, h)[1] if w else None)
    assert gb.encode_host_batch(codec, items) == 0
    for (src, dst, w, h, stride), want in zip(items, wants):
        if want is not None:
            got = dst.numpy() if hasattr(dst, "numpy") else dst
            assert np.array_equal(got[: w * h // 2], want), (w, h)


@pytest.mark.parametrize("path", ["LOAD_DIRECT", "LOAD_ONESHOT", "LOAD_ASYNC", "LOAD_TMA"])
@pytest.mark.parametrize("codec", CODECS)
def test_every_load_layer_is_bit_exact(codec, path, oracle):
    """All selectable load layers (goofy_b200_set_load_path) run the same block arithmetic: identical bytes."""
    prev = gb.set_load_path(getattr(gb, path))
    try:
        for (w, h, pad) in [(16, 4, 0), (272, 36, 0), (1024, 64, 256), (2064, 1028, 64)]:
            stride = w * 4 + pad
            tight = splitmix_rgba(w * h, seed=w + h + pad).reshape(h, w * 4)
            buf = np.full((h, stride), 0xCD, dtype=np.uint8)
            buf[:, : w * 4] = tight
            want = oracle.compress(codec, tight, w, h)[1]
            rc, got = gpu_device(codec, buf, w, h, stride)
            assert rc == 0 and np.array_equal(got, want), (path, w, h, pad)
        if codec == DXT1:
            n, w, h = 3, 528, 40
            imgs = np.stack([synth_family(i, w, h, seed=60 + i) for i in range(n)])
            d_a = torch.zeros((n, w * h // 2), dtype=torch.uint8, device="cuda")
            d_b = torch.zeros((n, w * h // 2), dtype=torch.uint8, device="cuda")
            assert gb.encode_dual_device(d_a, d_b, dev(imgs), w, h, w * 4, w * h * 4, w * h // 2, n) == 0
            torch.cuda.synchronize()
            for i in range(n):
                assert np.array_equal(d_a[i].cpu().numpy(), oracle.compress(DXT1, imgs[i], w, h)[1])
                assert np.array_equal(d_b[i].cpu().numpy(), oracle.compress(ETC1, imgs[i], w, h)[1])
    finally:
        gb.set_load_path(prev)


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("flavour", ["sse2", "floatref"])
def test_relaxed_shapes_edge_replication(codec, flavour, oracle):
    """Row N4: any width / height.  Expected output = the strict encoder applied to the image padded by
    replicating its last column / row up to the next legal size, cropped to ceil(w/4) x ceil(h/4) blocks."""
    rng = np.random.default_rng(99)
    gcodec = codec if flavour == "sse2" else {DXT1: gb.DXT1_FLOATREF, ETC1: gb.ETC1_FLOATREF}[codec]
    for (w, h) in [(1, 1), (3, 5), (4, 4), (17, 9), (30, 31), (64, 64), (100, 37), (259, 6), (1025, 3)]:
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        stride = w * 4
        bw, bh = (w + 3) // 4, (h + 3) // 4
        pw, ph = (w + 15) // 16 * 16, bh * 4
        padded = np.pad(img, ((0, ph - h), (0, pw - w), (0, 0)), mode="edge")
        if flavour == "sse2":
            full = oracle.compress(codec, padded, pw, ph)[1]
        else:
            full = oracle.compress_float_reference(codec, padded, pw, ph)[1]
        want = full.reshape(ph // 4, pw // 4, 8)[:, :bw].reshape(-1)
        d_dst = torch.zeros(bw * bh * 8, dtype=torch.uint8, device="cuda")
        assert gb.encode_relaxed_device(gcodec, d_dst, dev(img), w, h, stride) == 0
        torch.cuda.synchronize()
        assert np.array_equal(d_dst.cpu().numpy(), want), (w, h)
    # on a strict shape the relaxed path gives the strict path's bytes
    img = synth_family(1, 256, 64)
    d_dst = torch.zeros(256 * 64 // 2, dtype=torch.uint8, device="cuda")
    assert gb.encode_relaxed_device(gcodec, d_dst, dev(img), 256, 64, 1024) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_dst.cpu().numpy(), gpu_device(gcodec, img, 256, 64)[1])
    assert gb.encode_relaxed_device(gcodec, d_dst, dev(img), 0, 0, 0) == 0
    assert gb.encode_relaxed_device(gcodec, d_dst, dev(img), 10, 10, 36) == -5


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("flavour", ["sse2", "floatref"])
def test_relaxed_shapes_aligned_rows_interior_and_edge_blocks(codec, flavour, oracle):
    """16-byte-aligned rows: interior blocks take the 128-bit loads, the last block column / row the clamped ones.  Same
    expectation as above, padded strides, every combination of partial right column / partial bottom row, widths on both
    sides of a warp and of a CTA, and nothing written past the result."""
    rng = np.random.default_rng(1234)
    gcodec = codec if flavour == "sse2" else {DXT1: gb.DXT1_FLOATREF, ETC1: gb.ETC1_FLOATREF}[codec]
    shapes = [(17, 9, 80), (30, 31, 128), (64, 61, 256), (61, 64, 256), (127, 5, 512), (129, 130, 528), (259, 6, 1040),
              (1025, 3, 4112), (1021, 1023, 4096), (1024, 1022, 4096), (4099, 17, 16400), (8189, 61, 32768)]
    for (w, h, stride) in shapes:
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        rows = np.full((h, stride), 0xAB, dtype=np.uint8)
        rows[:, : w * 4] = img.reshape(h, w * 4)
        bw, bh = (w + 3) // 4, (h + 3) // 4
        pw, ph = (w + 15) // 16 * 16, bh * 4
        padded = np.pad(img, ((0, ph - h), (0, pw - w), (0, 0)), mode="edge")
        full = (oracle.compress(codec, padded, pw, ph) if flavour == "sse2" else oracle.compress_float_reference(codec, padded, pw, ph))[1]
        want = full.reshape(ph // 4, pw // 4, 8)[:, :bw].reshape(-1)
        d_dst = torch.full((bw * bh * 8 + 64,), 0x5A, dtype=torch.uint8, device="cuda")
        assert gb.encode_relaxed_device(gcodec, d_dst, dev(rows), w, h, stride) == 0
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy()
        assert np.array_equal(got[: bw * bh * 8], want), (w, h, stride)
        assert (got[bw * bh * 8:] == 0x5A).all(), (w, h, stride)   # nothing written past the result
