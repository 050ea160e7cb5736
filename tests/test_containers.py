"""Row N3: DDS / KTX containers and PNG ingest (host-side glue around the encoders)."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from goofy_b200 import containers
from oracle.oracle import DXT1, ETC1, synth_family

ROOT = Path(__file__).resolve().parent.parent


def test_dds_roundtrip_and_header_fields(tmp_path, oracle):
    img = synth_family(1, 64, 32)
    blocks = oracle.compress(DXT1, img, 64, 32)[1]
    p = tmp_path / "a.dds"
    containers.write_dds(p, blocks, 64, 32)
    raw = p.read_bytes()
    assert len(raw) == 128 + 64 * 32 // 2
    assert raw[:4] == b"DDS " and raw[84:88] == b"DXT1"
    assert struct.unpack_from("<I", raw, 4)[0] == 124 and struct.unpack_from("<2I", raw, 12) == (32, 64)
    back = containers.read_dds(p)
    assert (back["width"], back["height"], back["fourcc"], back["mips"]) == (64, 32, containers.FOURCC_DXT1, 1)
    assert np.array_equal(back["blocks"], blocks)


def test_dds_is_readable_by_an_independent_decoder(tmp_path, oracle):
    """PIL's own BC1 decoder opens the file; endpoint colours (indices 0 and 1) must agree exactly with ours."""
    Image = pytest.importorskip("PIL.Image")
    img = synth_family(1, 64, 64)
    blocks = oracle.compress(DXT1, img, 64, 64)[1]
    p = tmp_path / "b.dds"
    containers.write_dds(p, blocks, 64, 64)
    with Image.open(p) as im:
        assert im.size == (64, 64)
        theirs = np.array(im.convert("RGBA"))
    ours = oracle.decode(DXT1, blocks, 64, 64)
    idx = np.zeros((64, 64), dtype=np.uint8)
    words = blocks.reshape(-1, 8)[:, 4:].copy().view("<u4").reshape(16, 16)
    for y in range(64):
        for x in range(64):
            idx[y, x] = (int(words[y // 4, x // 4]) >> (2 * (4 * (y % 4) + (x % 4)))) & 3
    ends = idx < 2
    assert np.array_equal(theirs[ends][:, :3], ours[ends][:, :3])
    assert np.abs(theirs[..., :3].astype(int) - ours[..., :3].astype(int)).max() <= 1   # interpolants: rounding conventions differ by <= 1


def test_ktx_roundtrip_and_header_fields(tmp_path, oracle):
    img = synth_family(1, 48, 16)
    blocks = oracle.compress(ETC1, img, 48, 16)[1]
    p = tmp_path / "a.ktx"
    containers.write_ktx(p, blocks, 48, 16)
    raw = p.read_bytes()
    assert raw[:12] == containers.KTX_IDENTIFIER and len(raw) == 64 + 4 + 48 * 16 // 2
    back = containers.read_ktx(p)
    assert back["gl_internal_format"] == 0x8D64 and back["gl_base_internal_format"] == 0x1907
    assert (back["width"], back["height"], back["mips"], back["faces"]) == (48, 16, 1, 1)
    assert np.array_equal(back["blocks"], blocks)


def test_cpp_writers_produce_the_same_files(tmp_path, oracle):
    """include/goofy_containers.h against the Python writers, byte for byte."""
    img = synth_family(2, 32, 16)
    d = oracle.compress(DXT1, img, 32, 16)[1]
    e = oracle.compress(ETC1, img, 32, 16)[1]
    (tmp_path / "d.bin").write_bytes(d.tobytes())
    (tmp_path / "e.bin").write_bytes(e.tobytes())
    src = tmp_path / "w.cpp"
    src.write_text('''
#include "goofy_containers.h"
#include <vector>
static std::vector<unsigned char> slurp(const char* p) { std::vector<unsigned char> v(256); FILE* f = fopen(p, "rb"); fread(v.data(), 1, 256, f); fclose(f); return v; }
int main(int, char** a) {
    auto d = slurp(a[1]); auto e = slurp(a[2]);
    return (goofy::containers::writeDdsDxt1(a[3], d.data(), 32, 16) && goofy::containers::writeKtxEtc1(a[4], e.data(), 32, 16)) ? 0 : 1;
}''')
    exe = tmp_path / "w"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True, capture_output=True)
    subprocess.run([str(exe), str(tmp_path / "d.bin"), str(tmp_path / "e.bin"), str(tmp_path / "c.dds"), str(tmp_path / "c.ktx")], check=True)
    assert (tmp_path / "c.dds").read_bytes() == containers.dds_bytes(d, 32, 16)
    assert (tmp_path / "c.ktx").read_bytes() == containers.ktx_bytes(e, 32, 16)


def test_png_ingest_contract(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rgb = (np.arange(32 * 16 * 3) % 251).astype(np.uint8).reshape(16, 32, 3)
    Image.fromarray(rgb, "RGB").save(tmp_path / "ok.png")
    a = containers.load_png_rgba(tmp_path / "ok.png")
    assert a.shape == (16, 32, 4) and a.ctypes.data % 64 == 0
    assert np.array_equal(a[..., :3], rgb) and (a[..., 3] == 255).all()
    Image.fromarray(rgb[:, :24], "RGB").save(tmp_path / "w.png")
    Image.fromarray(rgb[:14], "RGB").save(tmp_path / "h.png")
    with pytest.raises(ValueError):
        containers.load_png_rgba(tmp_path / "w.png")
    with pytest.raises(ValueError):
        containers.load_png_rgba(tmp_path / "h.png")


def test_ktx_base_internal_format_follows_the_reference_rule():
    """saveKtx (Src/main.cpp:199 of the reference): glBaseInternalFormat is GL_RGB for ETC1 and GL_RGBA otherwise."""
    import struct
    blocks = np.zeros(16 * 16 // 2, dtype=np.uint8)
    etc = containers.ktx_bytes(blocks, 16, 16)
    dxt = containers.ktx_bytes(blocks, 16, 16, containers.GL_COMPRESSED_RGB_S3TC_DXT1_EXT)
    assert struct.unpack_from("<I", etc, 12 + 5 * 4)[0] == containers.GL_RGB
    assert struct.unpack_from("<I", dxt, 12 + 5 * 4)[0] == containers.GL_RGBA
