"""Python handle on the CPU oracle (oracle/oracle.c -> liboracle.so) and on the
compiled, untouched reference binaries in oracle/_ref/ (built by build_ref.sh).

TEST INFRASTRUCTURE ONLY: the checker, never the thing measured or shipped.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_TREE_BIN = {"hw1": "hw1_raytracer", "hw2": "hw2_raytracer", "hw3": "hw3_raytracer",
             "hw4": "hw4_raytracer", "hw5": "hw5_raytracer", "hw6": "hw6_raytracer", "hw7": "hw7_raytracer",
             "ppm": "ppm_raytracer",
             # float-dump twins (oracle/ref_driver_hdr.cpp instead of the tree's main.cpp)
             "hw2f": "hw2f_raytracer", "hw3f": "hw3f_raytracer", "hw4f": "hw4f_raytracer", "hw7f": "hw7f_raytracer"}


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_shadow", "rays_secondary", "nodes_visited",
                                          "tris_tested", "spheres_tested", "instances_entered")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    @property
    def rays_total(self):
        return int(self.rays_primary + self.rays_shadow + self.rays_secondary)


_lib = None


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        try:  # rebuild when oracle.c / b200rt.h changed (struct layout must match)
            from ceng795_b200 import _build
            _build.build_oracle()
        except Exception:
            if not os.path.exists(path):
                raise
        _lib = C.CDLL(path)
        _lib.oracle_render.restype = C.c_int
        _lib.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(Counters)]
        _lib.oracle_trace_rays.restype = C.c_int
        _lib.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.POINTER(Counters)]
        _lib.oracle_box_intersect.restype = C.c_float
        _lib.oracle_box_intersect.argtypes = [C.c_void_p] * 4 + [C.c_float]
        _lib.oracle_tri_intersect.restype = C.c_int
        _lib.oracle_tri_intersect.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_void_p]
        _lib.oracle_brdf_eval.restype = None
        _lib.oracle_brdf_eval.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float] + [C.c_void_p] * 6
        _lib.oracle_tex_color.restype = None
        _lib.oracle_tex_color.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        _lib.oracle_perlin_value.restype = C.c_float
        _lib.oracle_perlin_value.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def _desc_addr(host_scene):
    return C.cast(host_scene.desc_ptr, C.c_void_p)


def render(host_scene, camera_index=0, rows=(0, 0), seed=0, want_image=True, want_ids=False, threads=None,
           lit_counts=None):
    """Returns (rgbw[H,W,4] or None, ids[H,W,2] or None, t[H,W] or None, Counters).
    lit_counts: int32[H, W, n_lights] array that receives, per pixel and light, the number of shadow tests that came
    out lit (HW5-HW7 shading) — the device's b200rt_debug_lit_counts twin."""
    L = load()
    cam = host_scene.spec.cameras[camera_index]
    w, h = cam["width"], cam["height"]
    if lit_counts is not None:
        assert lit_counts.dtype == np.int32 and lit_counts.flags["C_CONTIGUOUS"]
        L.oracle_set_lit_counts.restype = None
        L.oracle_set_lit_counts.argtypes = [C.c_void_p]
        L.oracle_set_lit_counts(lit_counts.ctypes.data)
    rgbw = np.zeros((h, w, 4), np.float32) if want_image else None
    ids = np.full((h, w, 2), -1, np.int32) if want_ids else None
    t = np.full((h, w), np.inf, np.float32) if want_ids else None
    cnt = Counters()
    old = os.environ.get("OMP_NUM_THREADS")
    rc = L.oracle_render(_desc_addr(host_scene), camera_index, rows[0], rows[1], seed,
                         rgbw.ctypes.data if want_image else None,
                         ids.ctypes.data if want_ids else None, t.ctypes.data if want_ids else None,
                         C.byref(cnt))
    if lit_counts is not None:
        L.oracle_set_lit_counts(None)
    if rc != 0:
        raise RuntimeError(f"oracle_render rc={rc}")
    return rgbw, ids, t, cnt


def trace_rays(host_scene, rays):
    L = load()
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = rays.shape[0]
    ids = np.zeros((n, 2), np.int32)
    t = np.zeros(n, np.float32)
    cnt = Counters()
    rc = L.oracle_trace_rays(_desc_addr(host_scene), rays.ctypes.data, n, ids.ctypes.data, t.ctypes.data,
                             C.byref(cnt))
    if rc != 0:
        raise RuntimeError(f"oracle_trace_rays rc={rc}")
    return ids, t, cnt


def ppm(host_scene, camera_index=0, seed=0, n_photons=0, total_num_of_photons=None, want_hit_points=True,
        reverse_order=False):
    """The PPM frame through the oracle (single sequential photon stream).  Returns
    (rgbw[H,W,4], hit_points[n,16], initial_radius, Counters).  reverse_order traces the same photons last to first:
    the reference's threads leave that order free, so two orders are the yardstick for device-vs-oracle bounds."""
    L = load()
    L.oracle_ppm_photon_order.restype = None
    L.oracle_ppm_photon_order.argtypes = [C.c_int]
    L.oracle_ppm_photon_order(1 if reverse_order else 0)
    L.oracle_ppm.restype = C.c_int
    L.oracle_ppm.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                             C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(Counters)]
    cam = host_scene.spec.cameras[camera_index]
    w, h = cam["width"], cam["height"]
    rgbw = np.zeros((h, w, 4), np.float32)
    cap = w * h * 4 * 64
    hp = np.zeros((cap, 16), np.float32) if want_hit_points else None
    n_hp = C.c_int64()
    r0 = C.c_float()
    cnt = Counters()
    total = float(n_photons if total_num_of_photons is None else total_num_of_photons)
    rc = L.oracle_ppm(_desc_addr(host_scene), camera_index, seed, int(n_photons), total, rgbw.ctypes.data,
                      hp.ctypes.data if hp is not None else None, cap, C.byref(n_hp), C.byref(r0), C.byref(cnt))
    L.oracle_ppm_photon_order(0)
    if rc != 0:
        raise RuntimeError(f"oracle_ppm rc={rc}")
    return rgbw, (hp[: n_hp.value] if hp is not None else None), float(r0.value), cnt


def have_ref(tree: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, _TREE_BIN[tree]))


_TIME_UNITS = {"day": 86400e3, "hour": 3600e3, "minute": 60e3, "second": 1e3, "millisecond": 1.0,
               "microsecond": 1e-3}


def parse_time_ms(text: str) -> float:
    """Inverse of print_time_diff (HW3/include/timeutil.h:3-44)."""
    ms = 0.0
    for val, unit in re.findall(r"(\d+) (day|hour|minute|second|millisecond|microsecond)s?", text):
        ms += int(val) * _TIME_UNITS[unit]
    return ms


def run_reference(tree: str, xml_text: str, workdir: str | None = None, timeout: float = 3600.0):
    """Runs the untouched reference binary on the XML.  Returns dict(images={name: uint8[H,W,3]},
    ms=<the reference's own "is saved in" timer>, threads=N, stdout=...)."""
    from ceng795_b200.image import read_png
    exe = os.path.join(REF_DIR, _TREE_BIN[tree])
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    own = workdir is None
    tmp = tempfile.mkdtemp(prefix="b200rt_ref_") if own else workdir
    xml_path = os.path.join(tmp, "scene.xml")
    with open(xml_path, "w") as f:
        f.write(xml_text)
    r = subprocess.run([exe, xml_path], cwd=tmp, capture_output=True, text=True, timeout=timeout,
                       stdin=subprocess.DEVNULL)
    out = r.stdout
    if r.returncode != 0:
        raise RuntimeError(f"reference {tree} failed rc={r.returncode}: {out[-500:]} {r.stderr[-500:]}")
    res = {"stdout": out, "images": {}, "hdr": {}, "ms": {}, "threads": 1}
    m = re.search(r"Starting rendering on #(\d+) thread", out)
    if m:
        res["threads"] = int(m.group(1))
    for name, tail in re.findall(r"(\S+)\(\d+x\d+\) is saved in: ([^\n]*)", out):
        res["ms"][name] = parse_time_ms(tail)
        p = os.path.join(tmp, name)
        if not os.path.exists(p) and os.path.exists(p + ".png"):
            p = p + ".png"   # HW7/PPM print the name without its extension (HW7/src/main.cpp:56-57,97)
        if os.path.exists(p) and p.lower().endswith(".png"):
            res["images"][name] = read_png(p)
        if os.path.exists(p) and p.endswith(".f32"):
            # HW5/HW6 through oracle/ref_driver_hdr.cpp: int32 W, H, then W*H*3 float32 of Pixel::get_color()
            wh = np.fromfile(p, np.int32, 2)
            res.setdefault("hdr", {})[name[:-4]] = np.fromfile(p, np.float32, offset=8).reshape(int(wh[1]), int(wh[0]), 3)
    # PPM prints per-phase timers and "Saved png file. [ name ]" instead (PPM/src/main.cpp:56-102, write_png)
    for label, key in (("Eye pass", "eye"), ("Building hash grid", "grid"), ("Tracing photon rays", "photon"),
                       ("Density estimation", "density")):
        m = re.search(label + r" is completed in: ([^\n]*)", out)
        if m:
            res["ms"][key] = parse_time_ms(m.group(1))
    for name in re.findall(r"Saved png file\. \[ (\S+) \]", out):
        p = os.path.join(tmp, name)
        if os.path.exists(p) and name not in res["images"]:
            res["images"][name] = read_png(p)
    m = re.search(r"Starting photon_trace on #(\d+) thread", out)
    if m:
        res["threads"] = int(m.group(1))
    if own:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return res


# ---- libstdc++ restated, for the replay tests -------------------------------------------------------------------
class _Mt19937:
    def __init__(self, seed):
        x = [0] * 624
        x[0] = seed & 0xFFFFFFFF
        for k in range(1, 624):
            x[k] = (1812433253 * (x[k - 1] ^ (x[k - 1] >> 30)) + k) & 0xFFFFFFFF
        self.x, self.i = x, 624

    def next(self):
        if self.i >= 624:
            x = self.x
            for k in range(624):
                y = (x[k] & 0x80000000) | (x[(k + 1) % 624] & 0x7FFFFFFF)
                x[k] = x[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.i = 0
        y = self.x[self.i]
        self.i += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def _uniform_int(g, a, b):
    """std::uniform_int_distribution on a 32-bit engine, libstdc++ 11+: Lemire's nearly-divisionless method."""
    rng = (b - a) + 1
    prod = g.next() * rng
    low = prod & 0xFFFFFFFF
    if low < rng:
        thr = ((1 << 32) - rng) % rng
        while low < thr:
            prod = g.next() * rng
            low = prod & 0xFFFFFFFF
    return a + (prod >> 32)


def perlin_table_for_clock(seed: int) -> list:
    """The permutation table Perlin_noise::Perlin_noise builds (HW7/src/Perlin_noise.cpp:16-20) when the clock it seeds
    its mt19937 from reads `seed`: std::shuffle of 0..15 as libstdc++ does it (one swap for the even count, then two
    positions per draw).  Checked against the real std::shuffle for several seeds while writing it."""
    p = list(range(16))
    g = _Mt19937(seed)
    i = 1
    j = _uniform_int(g, 0, 1)
    p[i], p[j] = p[j], p[i]
    i += 1
    while i < 16:
        r = i + 1
        x = _uniform_int(g, 0, r * (r + 1) - 1)
        a, b = x // (r + 1), x % (r + 1)
        p[i], p[a] = p[a], p[i]
        i += 1
        p[i], p[b] = p[b], p[i]
        i += 1
    return p
