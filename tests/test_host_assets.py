"""CPU tests of the real-asset path of the host library: image decoders behind Texture::GET
(host/image_decode.cpp, the stbi_load(..., 4) replacement) and the glTF 2.0 reader behind Prefab::GET (host/gltf.cpp,
vk_mesh.cpp:611-901).  Inputs are written by the tests themselves (Pillow / numpy / struct): the reference's assets are
git-ignored upstream."""
import base64
import io
import json
import struct
import zlib

import numpy as np
import pytest

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200.capi import decode_image

PIL = pytest.importorskip("PIL.Image")


def _picture(w, h, seed=0):
    """Smooth colour field with some structure (so that JPEG has AC energy in every band) plus mild noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    r = 128 + 100 * np.sin(x / 7.0) * np.cos(y / 11.0)
    g = 128 + 90 * np.sin((x + y) / 13.0)
    b = 128 + 110 * np.cos(x / 5.0 + y / 17.0)
    img = np.stack([r, g, b], -1) + rng.normal(0, 6, size=(h, w, 3))
    img[h // 3: h // 3 + 5, :, :] = 255          # hard edges
    img[:, w // 2: w // 2 + 3, :] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------ PNG
@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "P+trns", "1"])
def test_png_matches_pillow_bit_for_bit(tmp_path, mode):
    rgb = _picture(67, 45, seed=1)
    im = PIL.fromarray(rgb)
    kw = {}
    if mode == "RGBA":
        a = (np.arange(67 * 45).reshape(45, 67) % 256).astype(np.uint8)
        im = PIL.fromarray(np.dstack([rgb, a]))
    elif mode in ("L", "LA", "1"):
        im = im.convert(mode)
    elif mode.startswith("P"):
        im = im.quantize(37)
        if mode == "P+trns":
            kw["transparency"] = bytes((i * 7) % 256 for i in range(37))
    path = tmp_path / "t.png"
    im.save(path, **kw)
    want = np.asarray(PIL.open(path).convert("RGBA"))
    got = decode_image(path)
    assert got.shape == want.shape and (got == want).all()


def _png_chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def _write_png(path, w, h, depth, color, rows_fn, interlace=0, extra=b""):
    """Hand-written PNG: rows_fn(pass_w, pass_h, xs, ys) -> list of packed scanlines (filter byte 0 is added here)."""
    raw = b""
    if interlace:
        xo, yo, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
        for k in range(7):
            xs, ys = list(range(xo[k], w, dx[k])), list(range(yo[k], h, dy[k]))
            if xs and ys:
                raw += b"".join(b"\x00" + r for r in rows_fn(xs, ys))
    else:
        raw = b"".join(b"\x00" + r for r in rows_fn(list(range(w)), list(range(h))))
    z = zlib.compress(raw)
    data = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, interlace)) + extra
    data += _png_chunk(b"IDAT", z[: len(z) // 2]) + _png_chunk(b"IDAT", z[len(z) // 2:]) + _png_chunk(b"IEND", b"")   # split IDAT
    path.write_bytes(data)


def test_png_adam7_16bit_and_low_depth(tmp_path):
    w, h = 21, 13
    rgb = _picture(w, h, seed=2)
    # interlaced 8-bit RGB
    _write_png(tmp_path / "i.png", w, h, 8, 2, lambda xs, ys: [rgb[y][xs].tobytes() for y in ys], interlace=1)
    got = decode_image(tmp_path / "i.png")
    assert (got[..., :3] == rgb).all() and (got[..., 3] == 255).all()
    assert (np.asarray(PIL.open(tmp_path / "i.png").convert("RGBA")) == got).all()
    # 16-bit RGBA: the high byte is kept (stb's 16 -> 8 conversion)
    v16 = (np.arange(w * h * 4, dtype=np.uint32).reshape(h, w, 4) * 977 % 65536).astype(">u2")
    _write_png(tmp_path / "s.png", w, h, 16, 6, lambda xs, ys: [v16[y][xs].tobytes() for y in ys])
    assert (decode_image(tmp_path / "s.png") == (v16.astype(np.uint16) >> 8).astype(np.uint8)).all()
    # 2-bit grey, interlaced, with a tRNS colour key: samples scale by 0x55, key 2 becomes transparent
    g2 = (np.arange(w * h).reshape(h, w) * 5 // 3 % 4).astype(np.uint8)

    def pack2(xs, ys):
        rows = []
        for y in ys:
            bits = "".join(format(int(g2[y, x]), "02b") for x in xs)
            bits += "0" * (-len(bits) % 8)
            rows.append(int(bits, 2).to_bytes(len(bits) // 8, "big"))
        return rows
    _write_png(tmp_path / "g.png", w, h, 2, 0, pack2, interlace=1, extra=_png_chunk(b"tRNS", struct.pack(">H", 2)))
    got = decode_image(tmp_path / "g.png")
    assert (got[..., 0] == g2 * 0x55).all() and (got[..., 1] == got[..., 0]).all() and (got[..., 3] == np.where(g2 == 2, 0, 255)).all()


def test_image_errors_are_reported(tmp_path):
    (tmp_path / "x.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 10)
    with pytest.raises(vk.VkpError, match="png"):
        decode_image(tmp_path / "x.png")
    (tmp_path / "x.bmp").write_bytes(b"BM" + b"\0" * 64)
    with pytest.raises(vk.VkpError, match="unknown image type"):
        decode_image(tmp_path / "x.bmp")
    with pytest.raises(vk.VkpError, match="cannot open"):
        decode_image(tmp_path / "missing.png")
    buf = io.BytesIO()
    PIL.fromarray(_picture(32, 32)).save(buf, "JPEG")
    (tmp_path / "cut.jpg").write_bytes(buf.getvalue()[:150])            # headers cut short: no frame
    with pytest.raises(vk.VkpError, match="jpeg"):
        decode_image(tmp_path / "cut.jpg")


# ------------------------------------------------------------------------------------------------------------ JPEG
@pytest.mark.parametrize("subsampling,progressive,grey,restart", [
    ("4:4:4", False, False, 0), ("4:2:0", False, False, 0), ("4:2:2", False, False, 0), ("4:2:0", True, False, 0),
    ("4:4:4", True, False, 0), ("4:4:4", False, True, 0), ("4:2:0", False, False, 3), ("4:2:0", True, False, 2)])
def test_jpeg_agrees_with_pillow_within_decoder_tolerance(tmp_path, subsampling, progressive, grey, restart):
    """The entropy decode is exact or nothing (a wrong coefficient shows as a whole wrong block); what may differ between
    two conforming decoders is IDCT rounding (+-1), chroma up-sampling taps and YCbCr rounding.  Sizes are not multiples
    of the MCU on purpose."""
    rgb = _picture(83, 61, seed=3)
    im = PIL.fromarray(rgb)
    if grey:
        im = im.convert("L")
    path = tmp_path / "t.jpg"
    kw = dict(quality=92, progressive=progressive, optimize=progressive)
    if not grey:
        kw["subsampling"] = subsampling
    if restart:
        kw["restart_marker_rows"] = restart
    im.save(path, "JPEG", **kw)
    data = path.read_bytes()
    assert (b"\xff\xc2" in data) == progressive
    if restart:
        assert b"\xff\xdd" in data
    want = np.asarray(PIL.open(path).convert("RGBA")).astype(np.int32)
    got = decode_image(path).astype(np.int32)
    assert got.shape == want.shape and (got[..., 3] == 255).all()
    d = np.abs(got - want)[..., :3]
    if grey:
        assert (got[..., 0] == got[..., 1]).all() and (got[..., 1] == got[..., 2]).all()
    full_res_chroma = grey or subsampling == "4:4:4"
    if subsampling == "4:2:2" and not grey:
        d = d[:, :-1]           # stb's h2v1 row filter weights the last pair the other way round (3*in[w-2] + in[w-1]); kept as published
    assert d.max() <= (2 if full_res_chroma else 12), d.max()            # up-sampled chroma differs at hard edges only
    assert d.mean() <= (0.35 if full_res_chroma else 0.8), d.mean()
    # and it is a good reconstruction of the source picture in its own right
    src = np.asarray(im.convert("RGB")).astype(np.int32)
    assert np.abs(got[..., :3] - src).mean() <= np.abs(want[..., :3] - src).mean() + 0.25


def test_jpeg_flat_block_and_colour_conversion_are_the_stated_integer_pipeline(tmp_path):
    """A DC-only block must decode exactly (the IDCT's DC shortcut and its full path agree on d0 * 4), and a flat colour patch
    must follow the stated 20-bit fixed-point BT.601 conversion bit for bit, masking of the Cb term included."""
    flat = np.full((16, 16), 77, np.uint8)
    PIL.fromarray(flat).save(tmp_path / "flat.jpg", "JPEG", quality=100)
    assert (decode_image(tmp_path / "flat.jpg")[..., :3] == 77).all()
    # YCbCr -> RGB in 20-bit fixed point: check the stated formula on a flat colour patch (no AC, chroma exact)
    im = PIL.new("YCbCr", (16, 16), (90, 200, 60))
    im.save(tmp_path / "col.jpg", "JPEG", quality=100, subsampling="4:4:4")
    got = decode_image(tmp_path / "col.jpg")[0, 0, :3].astype(int).tolist()

    def f2f(x):
        return int(np.float32(x) * np.float32(4096.0) + np.float32(0.5)) << 8
    yf, cb, cr = (90 << 20) + (1 << 19), 200 - 128, 60 - 128
    r = (yf + cr * f2f(1.402)) >> 20
    # masking a negative product with 0xffff0000 in two's complement == floor to a multiple of 65536
    gm = (cb * -f2f(0.34414)) // 65536 * 65536
    g = (yf + cr * -f2f(0.71414) + gm) >> 20
    b = (yf + cb * f2f(1.772)) >> 20
    assert got == [min(max(v, 0), 255) for v in (r, g, b)]


# ------------------------------------------------------------------------------------------------------------ HDR
def _rgbe(lin):
    """float RGB -> RGBE bytes (Ward)."""
    m = lin.max(-1)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0.0)
    out = np.zeros(lin.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(lin * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    return out


def _hdr_expected(rgbe):
    f1 = np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
    lin = np.where(rgbe[..., 3:4] != 0, rgbe[..., :3].astype(np.float32) * f1[..., None], np.float32(0)).astype(np.float32)
    z = np.power(lin.astype(np.float64), np.float64(np.float32(1.0) / np.float32(2.2))).astype(np.float32) * np.float32(255) + np.float32(0.5)
    z = np.clip(z, 0, 255)
    out = np.full(rgbe.shape[:-1] + (4,), 255, np.uint8)
    out[..., :3] = z.astype(np.int32).astype(np.uint8)
    return out


@pytest.mark.parametrize("rle", [False, True])
def test_radiance_hdr_is_tonemapped_like_stbi_load(tmp_path, rle):
    w, h = 40, 9
    rng = np.random.default_rng(5)
    lin = rng.uniform(0, 1, size=(h, w, 3)) ** 3 * 4.0
    lin[0, :8] = 0.0
    lin[1] = lin[1, 0]                                   # a constant scanline: long runs
    px = _rgbe(lin)
    body = b""
    if rle:
        for y in range(h):
            body += bytes([2, 2, w >> 8, w & 255])
            for c in range(4):
                row = px[y, :, c].tolist()
                i = 0
                while i < w:
                    j = i
                    while j < w and row[j] == row[i] and j - i < 127:
                        j += 1
                    if j - i >= 3:
                        body += bytes([128 + (j - i), row[i]]); i = j
                    else:
                        j = i
                        while j < w and j - i < 128 and not (j + 2 < w and row[j] == row[j + 1] == row[j + 2]):
                            j += 1
                        j = max(j, i + 1)
                        body += bytes([j - i] + row[i:j]); i = j
    else:
        body = px.tobytes()
    (tmp_path / "e.hdr").write_bytes(b"#?RADIANCE\n# test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n" + f"-Y {h} +X {w}\n".encode() + body)
    got = decode_image(tmp_path / "e.hdr")
    assert (got == _hdr_expected(px)).all()


# ------------------------------------------------------------------------------------------------------------ glTF
def _quat(axis, deg):
    a = np.asarray(axis, np.float64); a /= np.linalg.norm(a)
    s = np.sin(np.radians(deg) / 2)
    return [float(a[0] * s), float(a[1] * s), float(a[2] * s), float(np.cos(np.radians(deg) / 2))]


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _make_gltf(tmp, glb=False, embed=False):
    """Two root nodes; the first has a child.  Meshes: a quad (u16 indices, interleaved position / normal with a
    byteStride, un-normalised normals) and two triangles as two primitives (u32 and u8 indices).  Three materials, two of
    them textured; textures[i].source == i."""
    quad_pn = np.array([[-1, -1, 0, 0, 0, 2], [1, -1, 0, 0, 0, 2], [1, 1, 0, 0, 0, 2], [-1, 1, 0, 0, 0, 2]], np.float32)     # interleaved, stride 24
    quad_uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    quad_idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tri_a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tri_a_n = np.array([[0, 0, 1]] * 3, np.float32)
    tri_a_idx = np.array([0, 1, 2], np.uint32)
    tri_b = np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float32)      # no normals, no uv
    tri_b_idx = np.array([2, 1, 0], np.uint8)
    chunks, views = [], []

    def add(arr, stride=None):
        off = sum(len(c) for c in chunks)
        b = arr.tobytes()
        chunks.append(b + b"\0" * (-len(b) % 4))
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(b)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1
    v_quad, v_uv, v_qi = add(quad_pn, 24), add(quad_uv), add(quad_idx)
    v_ta, v_tan, v_tai, v_tb, v_tbi = add(tri_a), add(tri_a_n), add(tri_a_idx), add(tri_b), add(tri_b_idx)
    images = []
    tex_pixels = []
    for k, name in enumerate(["base.png", "normal.png", "emissive.jpg", "mr.png"]):
        pix = _picture(16 + 4 * k, 8 + 2 * k, seed=10 + k)
        buf = io.BytesIO()
        PIL.fromarray(pix).save(buf, "JPEG" if name.endswith("jpg") else "PNG")
        if embed:
            images.append({"bufferView": add(np.frombuffer(buf.getvalue(), np.uint8)), "mimeType": "image/jpeg" if name.endswith("jpg") else "image/png"})
        else:
            (tmp / name).write_bytes(buf.getvalue())
            images.append({"uri": name})
        tex_pixels.append(np.asarray(PIL.open(io.BytesIO(buf.getvalue())).convert("RGBA")))
    blob = b"".join(chunks)
    q = _quat([0, 1, 0], 90)
    doc = {
        "asset": {"version": "2.0"},
        "scene": 0,
        "scenes": [{"nodes": [0, 2]}],
        "nodes": [
            {"name": "root", "translation": [1, 2, 3], "rotation": q, "scale": [2, 2, 2], "children": [1], "mesh": 0},
            {"name": "child", "matrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.5, 0, 0, 1], "mesh": 1},
            {"name": "second", "translation": [0, -1, 0], "mesh": 1},
        ],
        "meshes": [
            {"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0}]},
            {"primitives": [{"attributes": {"POSITION": 4, "NORMAL": 5}, "indices": 6, "material": 1},
                            {"attributes": {"POSITION": 7}, "indices": 8, "material": 2}]},
        ],
        "accessors": [
            {"bufferView": v_quad, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
            {"bufferView": v_quad, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC3"},
            {"bufferView": v_uv, "componentType": 5126, "count": 4, "type": "VEC2"},
            {"bufferView": v_qi, "componentType": 5123, "count": 6, "type": "SCALAR"},
            {"bufferView": v_ta, "componentType": 5126, "count": 3, "type": "VEC3"},
            {"bufferView": v_tan, "componentType": 5126, "count": 3, "type": "VEC3"},
            {"bufferView": v_tai, "componentType": 5125, "count": 3, "type": "SCALAR"},
            {"bufferView": v_tb, "componentType": 5126, "count": 3, "type": "VEC3"},
            {"bufferView": v_tbi, "componentType": 5121, "count": 3, "type": "SCALAR"},
        ],
        "bufferViews": views,
        "materials": [
            {"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.25, 1.0, 1.0], "metallicFactor": 0.25, "roughnessFactor": 0.75,
                                      "baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 3}},
             "normalTexture": {"index": 1}, "emissiveTexture": {"index": 2}, "name": "textured é\"quoted\""},
            {"pbrMetallicRoughness": {"roughnessFactor": 0.5}},
            {"pbrMetallicRoughness": {"roughnessFactor": 0.5}},      # equal by value to material 1: de-duplicated
        ],
        "textures": [{"source": 0}, {"source": 1}, {"source": 2}, {"source": 3}],
        "images": images,
        "buffers": [{"byteLength": len(blob)}],
    }
    if glb:
        js = json.dumps(doc).encode()
        js += b" " * (-len(js) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
        (tmp / "DamagedHelmet.gltf").unlink(missing_ok=True)
        return struct.pack("<III", 0x46546C67, 2, 12 + len(body)) + body, tex_pixels
    if embed:
        doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    else:
        (tmp / "helmet data.bin").write_bytes(blob)
        doc["buffers"][0]["uri"] = "helmet%20data.bin"
    return json.dumps(doc, indent=1).encode(), tex_pixels


@pytest.mark.parametrize("variant", ["files", "embedded"])
def test_gltf_prefab_follows_loadnode(tmp_path, monkeypatch, variant):
    data, tex = _make_gltf(tmp_path, embed=variant == "embedded")
    (tmp_path / "DamagedHelmet.gltf").write_bytes(data)
    monkeypatch.setenv("VKPINUT_ASSETS", str(tmp_path))
    sc = vk.HostScene(0, detail=(2, 2, 2, 2))
    # the helmet entity (TLAS position 5..) now has 5 primitives: root quad; child 2; second 2 -- in node_to_instance order
    inst = sc.instances[5:]
    assert len(sc.instances) == 10 and (inst["mesh_id"] == inst["mesh_id"][0]).all()
    verts, idx = sc.meshes[int(inst["mesh_id"][0])]
    # loadNode loads children before the node's own primitives (vk_mesh.cpp:652-659): child's two triangles come first in the mesh
    assert len(verts) == 3 + 3 + 4 + 3 + 3 and len(idx) == 3 + 3 + 6 + 3 + 3
    assert verts["position"][:3].tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0]] and verts["position"][6:10, :2].tolist() == [[-1, -1], [1, -1], [1, 1], [-1, 1]]
    # instances: root (quad) first, then its child's primitives, then the second root
    assert inst["first_index"].tolist() == [6, 0, 3, 12, 15] and inst["index_count"].tolist() == [6, 3, 3, 3, 3]
    # indices are rebased by firstVertex (vk_mesh.cpp:722-745); the u8 primitive is 2,1,0
    assert idx[6:12].tolist() == [6, 7, 8, 6, 8, 9] and idx[0:3].tolist() == [0, 1, 2] and idx[3:6].tolist() == [5, 4, 3] and idx[15:18].tolist() == [15, 14, 13]
    # normals are normalised (:702), missing normals are glm::normalize(0) = NaN, uv is NOT flipped (unlike OBJ), colour white
    assert (verts["normal"][6:10] == [0, 0, 1]).all() and (verts["normal"][0:3] == [0, 0, 1]).all() and np.isnan(verts["normal"][3:6]).all()
    assert verts["uv"][6:10].tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]] and (verts["uv"][:6] == 0).all() and (verts["color"] == 1).all()
    # transforms: entity T(-5,1,-5) * node global; root = T * R * S (get_local_matrix), child = root * matrix, second = T(0,-1,0)
    E = np.eye(4); E[:3, 3] = [-5, 1, -5]
    R = np.eye(4); R[:3, :3] = _rot(_quat([0, 1, 0], 90)) * 2.0; R[:3, 3] = [1, 2, 3]
    Cm = np.eye(4); Cm[0, 3] = 0.5
    S2 = np.eye(4); S2[1, 3] = -1
    want = [E @ R, E @ R @ Cm, E @ R @ Cm, E @ S2, E @ S2]
    for k in range(5):
        m = inst["model"][k].reshape(4, 4).T.astype(np.float64)
        assert np.allclose(m, want[k], atol=1e-5), k
        assert np.allclose(m @ inst["inv_model"][k].reshape(4, 4).T.astype(np.float64), np.eye(4), atol=1e-4)
    # materials (loadMaterial, vk_mesh.cpp:777-803): textured one, then the two equal-by-value ones collapse into one entry
    mats = inst["material_id"].tolist()
    assert mats[1] == mats[2] == mats[3] == mats[4] and mats[0] != mats[1] and len(sc.materials) == 7
    mt = sc.materials[mats[0]]
    assert np.allclose(mt["diffuse"], [0.5, 0.25, 1.0, 1.0]) and np.allclose(mt["shadingMetallicRoughness"][:3], [0, 0.25, 0.75])
    assert np.allclose(sc.materials[mats[1]]["shadingMetallicRoughness"][:3], [0, 1.0, 0.5]) and (sc.materials[mats[1]]["textures"] == -1).all()
    # textures register in the order diffuse, normal, emissive, metallic-roughness (loadTextures :805-829); slots = (diffuse, normal, emissive, mr)
    t = [int(x) for x in mt["textures"]]
    assert t == [0, 1, 2, 3]
    for slot, k in zip(t, range(4)):
        got = sc.textures[slot]
        if k == 2:      # the JPEG: decoder tolerance
            assert got.shape == tex[k].shape and np.abs(got.astype(int) - tex[k].astype(int)).max() <= 12
        else:
            assert (got == tex[k]).all()
    assert sc.sky_tex == 4 and sc.irradiance_tex == 5                  # environment textures come after the scene's
    sc.close()


def test_glb_container_and_errors(tmp_path, monkeypatch):
    data, tex = _make_gltf(tmp_path, glb=True, embed=True)
    (tmp_path / "DamagedHelmet.gltf").write_bytes(b"{ not json")
    monkeypatch.setenv("VKPINUT_ASSETS", str(tmp_path))
    with pytest.raises(vk.VkpError, match="json"):
        vk.HostScene(0, detail=(2, 2, 2, 2))
    # a .glb is read through Prefab::GET("x.glb"): exercised with the viewer scene of the C shim
    (tmp_path / "DamagedHelmet.gltf").unlink()
    (tmp_path / "model.glb").write_bytes(data)
    sc = vk.HostScene.from_prefab(str(tmp_path / "model.glb"), detail=(2, 2, 2, 2))
    assert len(sc.instances) == 1 + 5 and sc.instances["index_count"][1:].tolist() == [6, 3, 3, 3, 3]      # floor + the prefab
    assert (sc.textures[0] == tex[0]).all() and len(sc.lights) == 3
    sc.close()
    with pytest.raises(vk.VkpError, match="cannot|not found|No valid"):
        vk.HostScene.from_prefab(str(tmp_path / "absent.glb"))
    bad = bytearray(data); bad[0:4] = b"XXXX"
    (tmp_path / "bad.glb").write_bytes(bytes(bad))
    with pytest.raises(vk.VkpError, match="GLB"):
        vk.HostScene.from_prefab(str(tmp_path / "bad.glb"))


def test_environment_textures_come_from_files_when_present(tmp_path, monkeypatch):
    sky = _picture(64, 32, seed=20)
    PIL.fromarray(sky).save(tmp_path / "LA_Downtown_Helipad_GoldenHour_8k.jpg", "JPEG", quality=95, subsampling="4:4:4")
    px = _rgbe(np.full((4, 8, 3), 0.25))
    (tmp_path / "LA_Downtown_Helipad_GoldenHour_Env.hdr").write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 8\n" + px.tobytes())
    monkeypatch.setenv("VKPINUT_ASSETS", str(tmp_path))
    sc = vk.HostScene(0, detail=(2, 2, 2, 2))
    assert sc.textures[sc.sky_tex].shape == (32, 64, 4) and np.abs(sc.textures[sc.sky_tex][..., :3].astype(int) - sky).mean() < 3
    assert sc.textures[sc.irradiance_tex].shape == (4, 8, 4) and (sc.textures[sc.irradiance_tex] == _hdr_expected(px)).all()
    sc.close()
