"""Output quantisation of each reference tree (last row).

The reference does this on the host after the hot path (Pixel::get_color and
main.cpp); it stays on the host here too and is shared by the GPU arm and the
oracle arm, so image parity tests compare the float accumulators *through the
same quantiser* the reference binary applies before lodepng.
"""
from __future__ import annotations

import numpy as np

from ._lib import SEM_HW1, SEM_HW2, SEM_HW3, SEM_HW4, SEM_HW5, SEM_HW6, SEM_HW7, SEM_PPM


def quantise(rgbw: np.ndarray, semantics: int) -> np.ndarray:
    """rgbw[H,W,4] = Pixel{sum color*w, sum w} -> uint8[H,W,3] as the tree's main.cpp writes it."""
    rgbw = np.asarray(rgbw, np.float32)
    w = rgbw[..., 3:4]
    with np.errstate(divide="ignore", invalid="ignore"):
        col = np.where(w == 0, np.float32(0), rgbw[..., :3] / w).astype(np.float32)
    if semantics in (SEM_HW1, SEM_HW2):
        # HW1/Scene.cpp:20-22, HW2/Pixel.h:24-26: min(255, max(0, int(c))) — C truncation toward zero
        q = np.clip(np.trunc(np.nan_to_num(col, nan=0.0, posinf=3e9, neginf=-3e9)), 0, 255)
        return q.astype(np.uint8)
    if semantics in (SEM_HW3, SEM_HW4, SEM_HW5, SEM_HW6, SEM_HW7):
        # HW7 without a tonemapper: write_png clamps to [0,255] and stores into unsigned char (HW7/src/main.cpp:105-115)
        # HW3/include/Pixel.h:25-28 clamps in float, HW3/src/main.cpp then stores into unsigned char
        q = np.minimum(np.float32(255), np.maximum(np.float32(0), np.nan_to_num(col, nan=0.0, posinf=3e9, neginf=-3e9)))
        return np.trunc(q).astype(np.uint8)
    if semantics == SEM_PPM:
        # PPM/src/main.cpp:150-157: int(pow(1 - exp(-x), 1 / 2.2f) * 255 + 0.5f), then write_png's clamp
        x = np.nan_to_num(col.astype(np.float64), nan=0.0, posinf=1e30, neginf=-1e30)
        v = np.power(np.maximum(0.0, 1.0 - np.exp(-x)), float(np.float32(1) / np.float32(2.2))) * 255 + 0.5
        return np.clip(np.trunc(v), 0, 255).astype(np.uint8)
    raise NotImplementedError(f"quantise: semantics {semantics}")


def read_png(path: str) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)


def write_png(path: str, rgb8: np.ndarray) -> None:
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(rgb8, np.uint8), "RGB").save(path)


def compare_8bit(a: np.ndarray, b: np.ndarray) -> dict:
    """Pixel-level agreement: exact fraction, within-1-LSB fraction, max abs diff."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=-1)
    return {"exact": float((d == 0).mean()), "within1": float((d <= 1).mean()), "max": int(d.max()),
            "n": int(d.size)}


def tonemap_photographic(rgbw: np.ndarray, key: float, burn_percent: float, saturation: float) -> np.ndarray:
    """Host restatement (test / oracle side) of Photographic_tmo::apply_tmo (HW7/src/Photographic_tmo.cpp:10-49)
    followed by main's sRGB curve (HW7/src/main.cpp:59-91) and write_png's clamp + uchar store: uint8[H,W,3].
    The log-average is a sequential float sum in the reference; here it is summed in float32 in the same order."""
    rgbw = np.asarray(rgbw, np.float32)
    h, w = rgbw.shape[:2]
    wt = rgbw[..., 3:4]
    with np.errstate(divide="ignore", invalid="ignore"):
        col = np.where(wt == 0, np.float32(0), rgbw[..., :3] / wt).astype(np.float32).reshape(-1, 3)
    f = np.float32
    lum = (f(0.21) * col[:, 0] + f(0.72) * col[:, 1]).astype(f) + f(0.07) * col[:, 2]
    lum = lum.astype(f)
    logs = np.log((lum + f(0.001)).astype(f)).astype(f)
    acc = f(0)
    for chunk in np.array_split(logs, max(1, logs.size // 4096)):      # sequential float32 accumulation
        for v in chunk:
            acc = f(acc + v)
    log_avg = f(np.exp(f(acc / f(lum.size))))
    lums = (f(key) * lum / log_avg).astype(f)
    srt = np.sort(lums)
    white_count = int(f(lum.size) * f(burn_percent) / f(100.0))
    white_count = max(0, min(lum.size, white_count))
    white = srt[lum.size - 1] if white_count == 0 else srt[lum.size - white_count]
    ld = ((lums * (f(1) + lums / f(white * white))) / (lums + f(1))).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (col / lum[:, None]).astype(f)
        out = (np.power(ratio, f(saturation)).astype(f) * ld[:, None]).astype(f)
    inv = f(1.0) / f(2.4)
    c = np.minimum(f(1), np.maximum(f(0), np.nan_to_num(out, nan=0.0, posinf=3e38, neginf=-3e38))).astype(f)
    g = (f(1.055) * np.power(c, inv).astype(f) - f(0.055)).astype(f)
    g = np.minimum(f(1), np.maximum(f(0), g)) * f(255.0)
    q = np.minimum(f(255), np.maximum(f(0), g))
    return np.trunc(q).astype(np.uint8).reshape(h, w, 3)


class AsyncImageWriter:
    """PNG encode off the critical path.  The reference's main encodes each camera's image on the
    thread that then starts the next camera's render (HW7/src/main.cpp:93-199, lodepng); here frames are handed to a
    worker thread, so the encode of camera i overlaps the render of camera i+1.  `submit` copies nothing: pass an
    array the caller will not overwrite (b200rt_render / b200rt_tonemap_photographic return fresh host buffers).

        with AsyncImageWriter() as w:
            for cam in range(n):
                rgbw, _ = scene.render_image(cam)
                w.submit(f"cam{cam}.png", quantise(rgbw, scene.semantics))
    """

    def __init__(self, max_pending: int = 4):
        import queue
        import threading
        self._q = queue.Queue(maxsize=max_pending)   # bounded: a slow disk throttles the producer, not memory
        self._err = None
        self.written = []
        self._t = threading.Thread(target=self._run, name="b200rt-png", daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            path, rgb8 = item
            try:
                write_png(path, rgb8)
                self.written.append(path)
            except Exception as e:  # surfaced by close(): the render loop must not lose an encode error
                self._err = e

    def submit(self, path: str, rgb8: np.ndarray) -> None:
        if self._err is not None:
            raise self._err
        self._q.put((path, rgb8))

    def close(self) -> None:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ---- minimal OpenEXR I/O: uncompressed scanline float32 RGB, the one flavour the env-map tests need ------------------
# (the reference reads its environment maps with tinyexr, HW7/src/Spherical_directional_light.cpp:4-12; real EXR
# decoding is host I/O that stays with the reference — this is just enough to hand both sides the same file)
def write_exr_f32(path: str, rgb: np.ndarray) -> None:
    import struct
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w = rgb.shape[:2]

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(c.encode() + b"\0" + struct.pack("<iBxxxii", 2, 0, 1, 1) for c in ("B", "G", "R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = (attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0") + attr("dataWindow", "box2i", box)
              + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0")
              + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
              + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
              + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    head = struct.pack("<II", 20000630, 2) + header
    line_bytes = 3 * w * 4
    offset0 = len(head) + 8 * h
    with open(path, "wb") as f:
        f.write(head)
        f.write(b"".join(struct.pack("<Q", offset0 + y * (8 + line_bytes)) for y in range(h)))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            for c in (2, 1, 0):                       # channels are stored in alphabetical order: B, G, R
                f.write(rgb[y, :, c].tobytes())


def read_exr_f32(path: str) -> np.ndarray:
    """Reads back what write_exr_f32 writes (uncompressed scanline float32 B/G/R); anything else is refused."""
    import struct
    blob = open(path, "rb").read()
    if struct.unpack_from("<I", blob, 0)[0] != 20000630:
        raise ValueError(f"{path}: not an OpenEXR file")
    pos = 8
    attrs = {}
    while blob[pos] != 0:
        e = blob.index(b"\0", pos); name = blob[pos:e].decode(); pos = e + 1
        e = blob.index(b"\0", pos); pos = e + 1
        (n,) = struct.unpack_from("<i", blob, pos); pos += 4
        attrs[name] = blob[pos:pos + n]; pos += n
    pos += 1
    if attrs.get("compression") != b"\0" or attrs["channels"].count(b"\0") != 3 * 1 + 1 + 0 + 3 * 0 and False:
        raise ValueError(f"{path}: only uncompressed EXR files are supported by this mirror")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    out = np.zeros((h, w, 3), np.float32)
    offsets = struct.unpack_from(f"<{h}Q", blob, pos)
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", blob, off)
        if nbytes != 3 * w * 4:
            raise ValueError(f"{path}: only float32 B/G/R scanlines are supported by this mirror")
        row = np.frombuffer(blob, np.float32, 3 * w, off + 8).reshape(3, w)
        out[y - y0, :, 2], out[y - y0, :, 1], out[y - y0, :, 0] = row[0], row[1], row[2]
    return out
