"""Texture decoding (TracerBoy::InitializeTexture, TracerBoy.cpp:2186-2246): PNG and TGA files written by
this test with Python's zlib (an independent encoder) in every colour type / bit depth / interlace mode, decoded by the
library's self-contained decoder through the C ABI, and compared with what DirectXTex's loaders would hand to the GPU
(format table DirectXTexWIC.cpp:34-81, sRGB rule :582-645)."""
import os
import struct
import zlib

import numpy as np
import pytest


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)


def _pack_rows(img, depth):
    """img: [h, w, c] integer samples -> list of packed scanlines (bytes) at `depth` bits per sample."""
    h, w, c = img.shape
    rows = []
    for y in range(h):
        s = img[y].reshape(-1)
        if depth == 16:
            rows.append(s.astype(">u2").tobytes())
        elif depth == 8:
            rows.append(s.astype(np.uint8).tobytes())
        else:
            bits = np.zeros((len(s) * depth + 7) // 8 * 8, np.uint8)
            for k in range(depth):
                bits[np.arange(len(s)) * depth + k] = (s >> (depth - 1 - k)) & 1
            rows.append(np.packbits(bits).tobytes())
    return rows


def _filter_rows(rows, bpp, rng):
    """Apply a random PNG filter (0-4) to every scanline."""
    out = b""
    prev = bytes(len(rows[0])) if rows else b""
    for r in rows:
        ft = int(rng.integers(0, 5))
        cur = np.frombuffer(r, np.uint8).astype(np.int32)
        up = np.frombuffer(prev, np.uint8).astype(np.int32)
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        c = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
        if ft == 0: f = cur
        elif ft == 1: f = cur - a
        elif ft == 2: f = cur - up
        elif ft == 3: f = cur - ((a + up) >> 1)
        else:
            p = a + up - c
            pa, pb, pc = abs(p - a), abs(p - up), abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
            f = cur - pred
        out += bytes([ft]) + (f & 255).astype(np.uint8).tobytes()
        prev = r
    return out


def _write_png(path, img, ctype, depth, interlace=False, extra=(), rng=None):
    rng = rng or np.random.default_rng(0)
    h, w, c = img.shape
    bpp = max(1, c * depth // 8)
    if not interlace:
        data = _filter_rows(_pack_rows(img, depth), bpp, rng)
    else:
        data = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = img[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                data += _filter_rows(_pack_rows(sub, depth), bpp, rng)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    for t, d in extra:
        png += _chunk(t, d)
    comp = zlib.compress(data, 6)
    half = len(comp) // 2
    png += _chunk(b"IDAT", comp[:half]) + _chunk(b"IDAT", comp[half:]) + _chunk(b"IEND", b"")  # split IDAT: must be concatenated
    open(path, "wb").write(png)


@pytest.mark.parametrize("ctype,depth,interlace", [(2, 8, False), (6, 8, False), (2, 8, True), (6, 16, False), (2, 16, True), (0, 8, False),
                                                    (0, 1, False), (0, 4, True), (0, 16, False), (4, 8, False), (3, 8, False), (3, 2, False), (3, 4, True)])
def test_png_decoding_matches_the_wic_loader_rules(ctype, depth, interlace, tmp_path, built):
    import tracerboy_b200 as tb
    rng = np.random.default_rng(ctype * 100 + depth)
    w, h = 37, 23  # not a multiple of 8: partial Adam7 passes, partial bytes at low bit depths
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    maxv = (1 << depth) - 1
    img = rng.integers(0, maxv + 1, (h, w, ch))
    extra = []
    pal = None
    if ctype == 3:
        pal = rng.integers(0, 256, (maxv + 1, 3)).astype(np.uint8)
        extra.append((b"PLTE", pal.tobytes()))
        trns = rng.integers(0, 256, maxv // 2 + 1).astype(np.uint8)   # shorter than the palette: the rest is opaque
        extra.append((b"tRNS", trns.tobytes()))
    p = str(tmp_path / "t.png")
    _write_png(p, img, ctype, depth, interlace, extra, rng)
    px, fmt, alpha = tb.load_image_file(p)
    assert px.shape == (h, w, 4)
    if depth == 16:   # R16G16B16A16_UNORM (R16_UNORM for grey: the shader reads (v, 0, 0, 1)), stored as float = v / 65535
        assert fmt == 0 and px.dtype == np.float32
        f = (img / np.float32(65535.0)).astype(np.float32)
        if ctype == 0:
            want = np.concatenate([f, np.zeros((h, w, 2), np.float32), np.ones((h, w, 1), np.float32)], -1)
        elif ctype == 2:
            want = np.concatenate([f, np.ones((h, w, 1), np.float32)], -1)
        else:
            want = f
        assert np.array_equal(px, want)
        assert alpha == (ctype == 6 and bool((img[..., 3] != 65535).any()))
        return
    assert fmt == 1 and px.dtype == np.uint8
    if ctype == 0:    # 8bppGray (1/2/4-bit widened) -> R8_UNORM
        v = ((img[..., 0] * 255 + maxv // 2) // maxv).astype(np.uint8)
        want = np.stack([v, np.zeros_like(v), np.zeros_like(v), np.full_like(v, 255)], -1)
    elif ctype == 2:
        want = np.concatenate([img, np.full((h, w, 1), 255)], -1).astype(np.uint8)
    elif ctype == 3:
        a = np.full(maxv + 1, 255, np.uint8); a[:len(trns)] = trns
        want = np.concatenate([pal[img[..., 0]], a[img[..., 0]][..., None]], -1)
    elif ctype == 4:  # grey + alpha -> 32bppRGBA (v, v, v, a)
        want = np.stack([img[..., 0]] * 3 + [img[..., 1]], -1).astype(np.uint8)
    else:
        want = img.astype(np.uint8)
    assert np.array_equal(px, want)
    assert alpha == bool((want[..., 3] != 255).any())


def test_png_srgb_metadata_selects_the_srgb_format(tmp_path, built):
    """DirectXTexWIC.cpp:582-645: an sRGB chunk, or a gAMA chunk of exactly 45455, makes the texture R8G8B8A8_UNORM_SRGB
    (the sampler linearises); any other gamma, or no metadata with WIC_FLAGS_NONE, leaves it R8G8B8A8_UNORM."""
    import tracerboy_b200 as tb
    img = np.random.default_rng(1).integers(0, 256, (8, 8, 3))
    for extra, want in (((), 1), (((b"sRGB", b"\x00"),), 2), (((b"gAMA", struct.pack(">I", 45455)),), 2), (((b"gAMA", struct.pack(">I", 100000)),), 1)):
        p = str(tmp_path / "s.png")
        _write_png(p, img, 2, 8, False, extra)
        assert tb.load_image_file(p)[1] == want


def _write_tga(path, img_bgr, bits, rle, top_down, cmap=None):
    h, w = img_bgr.shape[:2]
    typ = (1 if cmap is not None else (3 if bits == 8 else 2)) + (8 if rle else 0)
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 1 if cmap is not None else 0, typ, 0, len(cmap) if cmap is not None else 0,
                      24 if cmap is not None else 0, 0, 0, w, h, bits, (0x20 if top_down else 0) | (8 if bits == 32 else 0))
    body = cmap.tobytes() if cmap is not None else b""
    rows = img_bgr if top_down else img_bgr[::-1]
    flat = rows.reshape(h * w, -1).astype(np.uint8)
    if not rle:
        body += flat.tobytes()
    else:
        i = 0
        n = flat.shape[0]
        while i < n:
            run = 1
            while i + run < n and run < 128 and np.array_equal(flat[i + run], flat[i]):
                run += 1
            if run > 1:
                body += bytes([0x80 | (run - 1)]) + flat[i].tobytes()
                i += run
            else:
                lit = 1
                while i + lit < n and lit < 128 and not np.array_equal(flat[i + lit], flat[i + lit - 1]):
                    lit += 1
                body += bytes([lit - 1]) + flat[i:i + lit].tobytes()
                i += lit
    open(path, "wb").write(hdr + body)


@pytest.mark.parametrize("bits,rle,top_down", [(24, False, False), (24, True, True), (32, False, True), (32, True, False), (8, False, False), (8, True, True)])
def test_tga_decoding(bits, rle, top_down, tmp_path, built):
    import tracerboy_b200 as tb
    rng = np.random.default_rng(bits)
    w, h = 19, 11
    img = rng.integers(0, 4, (h, w, bits // 8)) * 60 + 10  # few distinct values: real runs for the RLE encoder
    if bits == 32:
        img[..., 3] = rng.integers(0, 2, (h, w)) * 255
    p = str(tmp_path / "t.tga")
    _write_tga(p, img, bits, rle, top_down)
    px, fmt, alpha = tb.load_image_file(p)
    assert fmt == 1 and px.shape == (h, w, 4)
    if bits == 8:
        want = np.stack([img[..., 0], np.zeros((h, w)), np.zeros((h, w)), np.full((h, w), 255)], -1)
    else:
        a = img[..., 3] if bits == 32 else np.full((h, w), 255)
        want = np.stack([img[..., 2], img[..., 1], img[..., 0], a], -1)  # BGR(A) on disk
    assert np.array_equal(px, want.astype(np.uint8))
    assert alpha == (bits == 32 and bool((img[..., 3] != 255).any()))


def test_colour_mapped_tga_and_errors(tmp_path, built):
    import tracerboy_b200 as tb
    rng = np.random.default_rng(5)
    cmap = rng.integers(0, 256, (16, 3)).astype(np.uint8)  # BGR entries
    idx = rng.integers(0, 16, (9, 13, 1))
    p = str(tmp_path / "c.tga")
    _write_tga(p, idx, 8, True, False, cmap)
    px, fmt, alpha = tb.load_image_file(p)
    assert np.array_equal(px[..., :3], cmap[idx[..., 0]][..., ::-1]) and (px[..., 3] == 255).all() and not alpha
    open(str(tmp_path / "bad.png"), "wb").write(b"\x89PNG\r\n\x1a\n" + b"\x00" * 40)
    with pytest.raises(tb.TracerBoyError) as e:
        tb.load_image_file(str(tmp_path / "bad.png"))
    assert e.value.code == -3
    open(str(tmp_path / "x.jpg"), "wb").write(b"\xff\xd8\xff")
    with pytest.raises(tb.TracerBoyError) as e:
        tb.load_image_file(str(tmp_path / "x.jpg"))
    assert e.value.code == -2   # JPEG / BMP / DDS: not implemented, said so


def test_damaged_images_are_rejected_not_trusted(tmp_path, built):
    """Malformed input is an error code, never a crash, a hang or an allocation sized by an untrusted header: a PNG whose
    header announces 59 484 x 42 301 x RGBA16 (20 GB of samples) over a few hundred bytes of data, truncations at every
    tenth byte, and a few hundred random byte flips of a valid PNG and a valid RLE TGA (the same mutations ran clean
    under -fsanitize=address,undefined over 5 500 files when the decoders were written)."""
    import time
    import tracerboy_b200 as tb
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (13, 17, 4))
    p = str(tmp_path / "a.png")
    _write_png(p, img, 6, 8, True, rng=rng)
    good = open(p, "rb").read()
    ih = bytearray(good[16:29]); ih[0:8] = struct.pack(">II", 59484, 42301); ih[8] = 16
    lying = good[:16] + bytes(ih) + struct.pack(">I", zlib.crc32(b"IHDR" + bytes(ih)) & 0xffffffff) + good[33:]
    open(p, "wb").write(lying)
    t0 = time.time()
    with pytest.raises(tb.TracerBoyError):
        tb.load_image_file(p)
    assert time.time() - t0 < 2.0
    t = str(tmp_path / "a.tga")
    _write_tga(t, rng.integers(0, 256, (9, 13, 4)), 32, True, False)
    good_tga = open(t, "rb").read()
    decoded = rejected = 0
    for path, data in ((p, good), (t, good_tga)):
        cases = [data[:k] for k in range(0, len(data), 10)]
        for _ in range(300):
            m = bytearray(data)
            for _ in range(int(rng.integers(1, 5))):
                m[int(rng.integers(len(m)))] = int(rng.integers(256))
            cases.append(bytes(m))
        for c in cases:
            open(path, "wb").write(c)
            try:
                px, fmt, alpha = tb.load_image_file(path)
                assert px.shape[0] * px.shape[1] <= 1 << 16   # a flipped size byte stays within what the data can hold
                decoded += 1
            except tb.TracerBoyError as e:
                assert e.code in (-3, -2)
                rejected += 1
    assert decoded > 50 and rejected > 200


def test_bundled_png_decodes_like_an_independent_decoder(built):
    """The one PNG in the reference's tree (Scenes/Teapot/textures/envmap.png, 8-bit RGB, iCCP but no sRGB / gAMA chunk
    -> R8G8B8A8_UNORM) against zlib + numpy."""
    import tracerboy_b200 as tb
    path = "/root/reference/Scenes/Teapot/textures/envmap.png"
    if not os.path.exists(path):
        pytest.skip("reference mount not present")
    raw = open(path, "rb").read()
    pos, idat, hdr = 8, b"", None
    while pos < len(raw):
        n, = struct.unpack(">I", raw[pos:pos + 4])
        t, d = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + n]
        if t == b"IHDR": hdr = struct.unpack(">IIBBBBB", d)
        if t == b"IDAT": idat += d
        pos += 12 + n
    w, h, depth, ctype, _, _, inter = hdr
    assert (depth, ctype, inter) == (8, 2, 0)
    data = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    out = np.zeros((h, 3 * w), np.int32)
    for y in range(h):
        ft, cur = data[y, 0], data[y, 1:].astype(np.int32)
        up = out[y - 1] if y else np.zeros(3 * w, np.int32)
        if ft == 0: out[y] = cur
        elif ft == 2: out[y] = (cur + up) & 255
        else:
            row = np.zeros(3 * w, np.int32)
            for x in range(3 * w):
                a = row[x - 3] if x >= 3 else 0
                c = up[x - 3] if x >= 3 else 0
                b = up[x]
                if ft == 1: pr = a
                elif ft == 3: pr = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pr = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                row[x] = (cur[x] + pr) & 255
            out[y] = row
    px, fmt, alpha = tb.load_image_file(path)
    assert fmt == 1 and not alpha and np.array_equal(px[..., :3], out.reshape(h, w, 3).astype(np.uint8)) and (px[..., 3] == 255).all()


TEXTURED_PBRT = """LookAt 0 0 5  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
WorldBegin
LightSource "distant" "point from" [1 3 2] "point to" [0 0 0] "rgb L" [3 3 3]
Texture "kd_srgb" "spectrum" "imagemap" "string filename" "srgb.png"
Texture "kd_lin" "spectrum" "imagemap" "string filename" "linear.png"
Texture "kd_tga" "spectrum" "imagemap" "string filename" "wood.tga"
Material "uber" "texture Kd" "kd_srgb" "float roughness" 0.4
Shape "trianglemesh" "point P" [-2 -1 0  0 -1 0  0 1 0  -2 1 0] "integer indices" [0 1 2 0 2 3] "float uv" [0 0 1 0 1 1 0 1]
Material "matte" "texture Kd" "kd_lin"
Shape "trianglemesh" "point P" [0 -1 0  2 -1 0  2 1 0  0 1 0] "integer indices" [0 1 2 0 2 3] "float uv" [0 0 2 0 2 2 0 2]
Material "substrate" "texture Kd" "kd_tga" "rgb Ks" [0.04 0.04 0.04] "float uroughness" 0.1 "float vroughness" 0.1
Shape "trianglemesh" "point P" [-2 -1.2 -1  2 -1.2 -1  2 -1.2 2  -2 -1.2 2] "integer indices" [0 1 2 0 2 3] "float uv" [0 0 1 0 1 1 0 1]
WorldEnd
"""


def write_textured_scene(d):
    """A PBRT scene whose albedo maps are PNG (with and without sRGB metadata) and TGA files, written next to it."""
    rng = np.random.default_rng(9)
    _write_png(os.path.join(d, "srgb.png"), rng.integers(0, 256, (16, 16, 3)), 2, 8, False, ((b"sRGB", b"\x00"),))
    _write_png(os.path.join(d, "linear.png"), rng.integers(0, 256, (8, 12, 4)), 6, 8, True)
    _write_tga(os.path.join(d, "wood.tga"), rng.integers(0, 256, (10, 10, 3)), 24, True, False)
    src = os.path.join(d, "textured.pbrt")
    open(src, "w").write(TEXTURED_PBRT)
    return src


def test_pbrt_scene_with_png_and_tga_albedo_maps_flattens(tmp_path, built):
    """LoadScene with image textures other than .hdr (TracerBoy.cpp:177-251, 2186-2246): the .tbscene holds one image per
    map in the loader's format; NEEDS_GAMMA_CORRECTION only on the uber material's map (CreateTexture(map_kd, true),
    TracerBoy.cpp:340) because its format is "normalized"; an image with alpha clears NO_ALPHA on its material."""
    import tracerboy_b200 as tb
    from tracerboy_b200 import build
    if not os.path.exists(os.path.join(build.LIB, "libtb_pbrtimport.so")):
        pytest.skip("PBRT importer not built (needs the reference mount at build time)")
    src = write_textured_scene(str(tmp_path))
    dst = str(tmp_path / "t.tbscene")
    tb.convert_scene(src, dst)
    raw = open(dst, "rb").read()
    magic, version, flip, ng, nv, ni, nm, nl, nt, nimg, env = struct.unpack_from("<8sII7Ii", raw, 0)
    assert (ng, nt, nimg) == (3, 3, 3)
    off = 196 + ng * 32 + nv * 12 + nv * 32 + ni * 4
    mats = np.frombuffer(raw, np.uint32, nm * 21, off).reshape(nm, 21)
    off += nm * 84 + nl * 104
    tex = np.frombuffer(raw, np.uint32, nt * 20, off).reshape(nt, 20)
    off += nt * 80 + nm * 64
    fmts = []
    for i in range(nimg):
        w, h, fmt, nbytes = struct.unpack_from("<4I", raw, off)
        fmts.append((w, h, fmt))
        assert nbytes == w * h * 4
        off += 16 + nbytes
    assert fmts == [(16, 16, 2), (12, 8, 1), (10, 10, 1)]
    assert [int(t[0]) for t in tex] == [0, 0, 0] and [int(t[1]) for t in tex] == [0, 1, 2]   # image textures -> images 0..2
    assert [int(t[2]) & 1 for t in tex] == [1, 0, 0]                                            # gamma flag: uber map_kd only
    geoms = np.frombuffer(raw, np.uint32, ng * 8, 196).reshape(ng, 8)
    by_shape = [mats[g[0]] for g in geoms]
    assert [int(m[3]) for m in by_shape] == [0, 1, 2]                                           # albedoIndex
    NOALPHA = 0x20
    assert [bool(int(m[19]) & NOALPHA) for m in by_shape] == [True, False, True]                # the RGBA png has alpha
