"""ctypes binding of include/b200rt.h (the C ABI) — mirrors the header field for field.

The product path has NO CPU fallback: if ``libb200rt.so`` is missing or the box
has no CUDA device, rendering calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

c_f = C.c_float
c_i = C.c_int32
c_u = C.c_uint32

SEM_HW1, SEM_HW2, SEM_HW3, SEM_HW4, SEM_HW5, SEM_HW6, SEM_HW7, SEM_PPM = range(1, 9)
SEM_BY_NAME = {"hw1": SEM_HW1, "hw2": SEM_HW2, "hw3": SEM_HW3, "hw4": SEM_HW4, "hw5": SEM_HW5,
               "hw6": SEM_HW6, "hw7": SEM_HW7, "ppm": SEM_PPM}

KIND_INNER, KIND_TRI, KIND_SPHERE, KIND_TRIPRIM, KIND_INSTANCE, KIND_NONE = 0, 1, 2, 3, 4, 15
KIND_LEAF1, KIND_LEAF4, KIND_FNODE, KIND_LIGHTINST = 5, 8, 9, 10

LIGHT_POINT, LIGHT_AREA, LIGHT_SPOT, LIGHT_DIRECTIONAL, LIGHT_ENV, LIGHT_MESH, LIGHT_SPHERE = range(7)
TEX_IMAGE, TEX_PERLIN = 0, 1
TEX_NEAREST, TEX_BILINEAR = 0, 1
TEX_REPLACE_KD, TEX_BLEND_KD, TEX_REPLACE_ALL = 0, 1, 2
TEX_REPEAT, TEX_CLAMP, TEX_VEIN, TEX_PATCH = 0, 1, 2, 3
BRDF_PHONG, BRDF_MODIFIED_PHONG, BRDF_BLINN_PHONG, BRDF_MODIFIED_BLINN_PHONG, BRDF_TORRANCE_SPARROW = range(5)


class BvhNode(C.Structure):
    _fields_ = [("lmin", c_f * 3), ("lmax", c_f * 3), ("rmin", c_f * 3), ("rmax", c_f * 3),
                ("left", c_u), ("right", c_u), ("pad", c_u * 2)]


class Mesh(C.Structure):
    _fields_ = [("root", c_u), ("root_box", c_f * 6), ("tri_begin", c_i), ("tri_count", c_i),
                ("node_begin", c_i), ("node_count", c_i), ("fast_root", c_u), ("pad", c_i)]


class Instance(C.Structure):
    _fields_ = [("inv", c_f * 12), ("nrm", c_f * 9), ("fwd", c_f * 16), ("box", c_f * 6),
                ("velocity", c_f * 3), ("mesh", c_i), ("material", c_i), ("texture", c_i),
                ("flags", c_u), ("rank", c_i), ("light", c_i), ("pad", c_i * 2)]


class Sphere(C.Structure):
    _fields_ = [("center", c_f * 3), ("radius", c_f), ("inv", c_f * 12), ("nrm", c_f * 9),
                ("fwd", c_f * 16), ("velocity", c_f * 3), ("material", c_i), ("texture", c_i),
                ("flags", c_u), ("rank", c_i), ("light", c_i), ("pad", c_i)]


class TriPrim(C.Structure):
    _fields_ = [("v0", c_f * 3), ("a1", c_f * 3), ("a2", c_f * 3), ("normal", c_f * 3),
                ("inv", c_f * 12), ("nrm", c_f * 9), ("material", c_i), ("flags", c_u), ("rank", c_i)]


class Material(C.Structure):
    _fields_ = [("ambient", c_f * 3), ("diffuse", c_f * 3), ("specular", c_f * 3), ("mirror", c_f * 3),
                ("transparency", c_f * 3), ("roughness", c_f), ("refraction_index", c_f),
                ("phong_exponent", c_f), ("brdf", c_i), ("degamma", c_i), ("pad", c_f * 4)]


class Light(C.Structure):
    _fields_ = [("kind", c_i), ("position", c_f * 3), ("intensity", c_f * 3), ("edge1", c_f * 3),
                ("edge2", c_f * 3), ("normal", c_f * 3), ("object", c_i), ("cdf_begin", c_i),
                ("cdf_count", c_i), ("total_area", c_f), ("pad", c_f)]


class Texture(C.Structure):
    _fields_ = [("kind", c_i), ("interp", c_i), ("decal", c_i), ("appearance", c_i), ("normalizer", c_f),
                ("scaling", c_f), ("bump_multiplier", c_f), ("is_bump", c_i), ("width", c_i), ("height", c_i),
                ("texel_offset", C.c_int64), ("perm", c_i * 16)]


class Brdf(C.Structure):
    _fields_ = [("kind", c_i), ("exponent", c_f), ("normalized", c_i), ("refractive_index", c_f)]


class Camera(C.Structure):
    _fields_ = [("e", c_f * 3), ("u", c_f * 3), ("v", c_f * 3), ("w", c_f * 3),
                ("top_left", c_f * 3), ("su", c_f * 3), ("sv", c_f * 3),
                ("left", c_f), ("right", c_f), ("bottom", c_f), ("top", c_f), ("distance", c_f),
                ("width", c_i), ("height", c_i), ("samples_sqrt", c_i), ("aperture", c_f),
                ("pad", c_i * 3)]


class CameraIn(C.Structure):
    _fields_ = [("position", c_f * 3), ("gaze", c_f * 3), ("up", c_f * 3), ("near_plane", c_f * 4),
                ("near_distance", c_f), ("width", c_i), ("height", c_i), ("num_samples", c_i),
                ("focus_distance", c_f), ("aperture_size", c_f), ("left_handed", c_i), ("simple", c_i),
                ("fov_y", c_f)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("semantics", c_i), ("max_depth", c_i), ("background", c_f * 3), ("ambient", c_f * 3),
        ("shadow_eps", c_f), ("isect_eps", c_f), ("root", c_u), ("root_box", c_f * 6),
        ("n_nodes", c_i), ("nodes", C.POINTER(BvhNode)),
        ("n_tris", c_i), ("tri_geom", C.POINTER(c_f)), ("tri_vidx", C.POINTER(c_i)),
        ("tri_info", C.POINTER(c_i)),
        ("n_fast_nodes", c_i), ("fast_nodes", C.POINTER(BvhNode)), ("n_fast_tris", c_i),
        ("fast_tri_geom", C.POINTER(c_f)),
        ("fast_tri_ref", C.POINTER(c_i)), ("tri_gate", C.POINTER(c_f)),
        ("n_verts", c_i), ("vert_pos", C.POINTER(c_f)), ("vert_normal", C.POINTER(c_f)),
        ("n_meshes", c_i), ("meshes", C.POINTER(Mesh)),
        ("n_instances", c_i), ("instances", C.POINTER(Instance)),
        ("n_spheres", c_i), ("spheres", C.POINTER(Sphere)),
        ("n_triprims", c_i), ("triprims", C.POINTER(TriPrim)),
        ("n_materials", c_i), ("materials", C.POINTER(Material)),
        ("n_lights", c_i), ("lights", C.POINTER(Light)),
        ("n_cameras", c_i), ("cameras", C.POINTER(Camera)),
        ("tri_uv", C.POINTER(c_f)), ("tri_tex", C.POINTER(c_i)),
        ("n_textures", c_i), ("textures", C.POINTER(Texture)),
        ("n_texels", C.c_int64), ("texels", C.POINTER(c_f)),
        ("n_brdfs", c_i), ("brdfs", C.POINTER(Brdf)),
        ("n_light_cdf", c_i), ("light_cdf", C.POINTER(c_f)), ("light_tri", C.POINTER(c_i)),
        ("background_texture", c_i), ("env_light", c_i), ("integrator", c_i), ("importance_sampling", c_i),
    ]


class RenderOpts(C.Structure):
    _fields_ = [("row_begin", c_i), ("row_end", c_i), ("row_stride", c_i), ("seed", C.c_uint64),
                ("sample_begin", c_i), ("sample_end", c_i), ("exact_traversal", c_i), ("contexts", c_i), ("row_group", c_i), ("pad", c_i * 3)]


class SharedFrame(C.Structure):
    """include/b200rt.h: B200SharedFrame — plain bytes, `bytes(frame)` travels through any channel."""
    _fields_ = [("ipc_handle", C.c_uint8 * 64), ("device_ptr", C.c_uint64), ("pid", C.c_int64),
                ("tail_offset", C.c_uint64), ("device", c_i), ("width", c_i), ("height", c_i), ("direct", c_i)]


class RenderStats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_secondary", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("spheres_tested", C.c_uint64),
                ("instances_entered", C.c_uint64),
                ("sh_nodes_visited", C.c_uint64), ("sh_tris_tested", C.c_uint64), ("sh_spheres_tested", C.c_uint64),
                ("sh_instances_entered", C.c_uint64),
                ("ms_generate", c_f), ("ms_extend", c_f), ("ms_shade", c_f), ("ms_shadow", c_f),
                ("ms_total", c_f), ("waves", c_i), ("launches", c_i), ("node_bytes", c_i), ("top_objects", c_i), ("contexts", c_i), ("pad", c_i)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}

    @property
    def rays_total(self) -> int:
        return self.rays_primary + self.rays_shadow + self.rays_secondary


class PpmOpts(C.Structure):
    _fields_ = [("photons", C.c_int64), ("first_photon", C.c_uint64), ("seed", C.c_uint64), ("pad", c_i * 4)]


class PpmStats(C.Structure):
    _fields_ = [("hit_points", C.c_uint64), ("grid_entries", C.c_uint64), ("photons", C.c_uint64),
                ("photon_rays", C.c_uint64), ("cell_scans", C.c_uint64), ("deposits", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64),
                ("ms_eye", c_f), ("ms_grid", c_f), ("ms_photon", c_f), ("ms_extend", c_f), ("ms_deposit", c_f),
                ("ms_density", c_f), ("initial_radius", c_f), ("launches", c_i), ("node_bytes", c_i), ("top_objects", c_i),
                ("entries_scanned", C.c_uint64), ("scan_cell", c_f), ("grid_rebuilds", c_i)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class B200RTError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libb200rt.so (building it in-tree first when nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("B200RT_LIB"):   # tuning aid: a variant built by tools/build_variant.sh (same ABI, other -D knobs)
        path = os.environ["B200RT_LIB"]
        build_if_missing = False
    if build_if_missing:
        try:
            _build.build_lib()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise B200RTError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "There is no CPU fallback.")
    L = C.CDLL(path)
    P = C.POINTER
    fp = P(c_f)
    sig = {
        "b200rt_xf_identity": (None, [fp]),
        "b200rt_xf_scaling": (None, [c_f, c_f, c_f, fp]),
        "b200rt_xf_translation": (None, [c_f, c_f, c_f, fp]),
        "b200rt_xf_rotation": (None, [c_f, c_f, c_f, c_f, fp]),
        "b200rt_xf_mul": (None, [fp, fp, fp]),
        "b200rt_xf_invert": (C.c_int, [fp, fp]),
        "b200rt_builder_create": (C.c_void_p, [C.c_int]),
        "b200rt_builder_destroy": (None, [C.c_void_p]),
        "b200rt_builder_set_globals": (C.c_int, [C.c_void_p, fp, fp, c_f, C.c_int]),
        "b200rt_builder_add_camera": (C.c_int, [C.c_void_p, P(CameraIn)]),
        "b200rt_builder_add_material": (C.c_int, [C.c_void_p, P(Material)]),
        "b200rt_builder_add_light": (C.c_int, [C.c_void_p, P(Light)]),
        "b200rt_builder_set_vertices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
        "b200rt_builder_add_mesh": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, fp, fp]),
        "b200rt_builder_add_instance": (C.c_int, [C.c_void_p, C.c_int, C.c_int, fp, fp]),
        "b200rt_builder_add_triangle": (C.c_int, [C.c_void_p, C.c_int, P(c_i), fp]),
        "b200rt_builder_add_sphere": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f, fp, fp]),
        "b200rt_builder_set_texcoords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
        "b200rt_builder_add_texture": (C.c_int, [C.c_void_p, P(Texture), C.c_void_p]),
        "b200rt_builder_add_brdf": (C.c_int, [C.c_void_p, P(Brdf)]),
        "b200rt_builder_set_mesh_texture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
        "b200rt_builder_set_sphere_texture": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        "b200rt_builder_add_light_mesh": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, fp, fp]),
        "b200rt_builder_add_light_sphere": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f, fp, fp]),
        "b200rt_builder_set_integrator": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        "b200rt_builder_set_background_texture": (C.c_int, [C.c_void_p, C.c_int]),
        "b200rt_builder_finish": (C.c_int, [C.c_void_p, P(P(SceneDesc))]),
        "b200rt_device_count": (C.c_int, []),
        "b200rt_scene_create": (C.c_int, [P(SceneDesc), C.c_int, P(C.c_void_p)]),
        "b200rt_scene_destroy": (None, [C.c_void_p]),
        "b200rt_render": (C.c_int, [C.c_void_p, C.c_int, P(RenderOpts), C.c_void_p, P(RenderStats)]),
        "b200rt_render_async": (C.c_int, [C.c_void_p, C.c_int, P(RenderOpts), C.c_void_p, P(RenderStats)]),
        "b200rt_render_wait": (C.c_int, [C.c_void_p]),
        "b200rt_frame_collect_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
        "b200rt_render_device": (C.c_int, [C.c_void_p, C.c_int, P(RenderOpts), P(C.c_void_p), P(RenderStats)]),
        "b200rt_primary_ids": (C.c_int, [C.c_void_p, C.c_int, P(RenderOpts), C.c_void_p, C.c_void_p]),
        "b200rt_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_void_p, P(RenderStats)]),
        "b200rt_debug_pool": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
        "b200rt_debug_lit_counts": (C.c_int, [C.c_void_p, C.c_int]),
        "b200rt_debug_get_lit_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
        "b200rt_tonemap_photographic": (C.c_int, [C.c_void_p, C.c_int, c_f, c_f, c_f, C.c_void_p]),
        "b200rt_ppm_eye_pass": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, P(PpmStats)]),
        "b200rt_ppm_trace_photons": (C.c_int, [C.c_void_p, P(PpmOpts), P(PpmStats)]),
        "b200rt_ppm_density": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
        "b200rt_ppm_export": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_int64)]),
        "b200rt_ppm_density_from": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
        "b200rt_ppm_get_hit_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, P(C.c_int64)]),
        "b200rt_frame_export": (C.c_int, [C.c_void_p, C.c_int, P(SharedFrame)]),
        "b200rt_frame_attach": (C.c_int, [C.c_void_p, P(SharedFrame)]),
        "b200rt_frame_detach": (C.c_int, [C.c_void_p]),
        "b200rt_frame_begin": (C.c_int, [C.c_void_p]),
        "b200rt_render_shared": (C.c_int, [C.c_void_p, C.c_int, P(RenderOpts), P(RenderStats)]),
        "b200rt_frame_collect": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, P(C.c_void_p)]),
        "b200rt_multi_share": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, P(RenderOpts)]),
        "b200rt_render_multi": (C.c_int, [P(C.c_void_p), C.c_int, C.c_int, P(RenderOpts), C.c_void_p, P(RenderStats)]),
        "b200rt_ppm_set_share": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        "b200rt_ppm_counts_delta": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_int64)]),
        "b200rt_ppm_counts_apply": (C.c_int, [C.c_void_p]),
        "b200rt_ppm_eye_pass_multi": (C.c_int, [P(C.c_void_p), C.c_int, C.c_int, C.c_uint64, P(PpmStats)]),
        "b200rt_ppm_trace_photons_multi": (C.c_int, [P(C.c_void_p), C.c_int, P(PpmOpts), P(PpmStats)]),
        "b200rt_ppm_density_multi": (C.c_int, [P(C.c_void_p), C.c_int, C.c_double, C.c_void_p]),
        "b200rt_scene_stream": (C.c_void_p, [C.c_void_p]),
        "b200rt_scene_bytes": (C.c_uint64, [C.c_void_p]),
        "b200rt_builder_set_instance_transform": (C.c_int, [C.c_void_p, C.c_int, fp, fp]),
        "b200rt_builder_set_sphere_transform": (C.c_int, [C.c_void_p, C.c_int, fp, fp]),
        "b200rt_builder_refit": (C.c_int, [C.c_void_p, P(P(SceneDesc))]),
        "b200rt_scene_refit": (C.c_int, [C.c_void_p, P(SceneDesc), C.POINTER(C.c_float)]),
        "b200rt_scene_build_bvh_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
        "b200rt_lbvh_ranges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
        "b200rt_lbvh_morton63": (C.c_uint64, [C.c_float, C.c_float, C.c_float]),
        "b200rt_debug_leaf": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_void_p, C.c_int]),
        "b200rt_scene_desc_save": (C.c_int, [C.POINTER(SceneDesc), C.c_char_p]),
        "b200rt_scene_desc_load": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(SceneDesc))]),
        "b200rt_scene_desc_free": (None, [C.POINTER(SceneDesc)]),
        "b200rt_measure_read_bandwidth": (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
        "b200rt_last_error": (C.c_char_p, []),
        "b200rt_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError = header/lib drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


#: every symbol include/b200rt.h declares (tests/test_abi.py checks the .so exports them all)
def declared_symbols() -> list[str]:
    import re
    hdr = os.path.join(_build.ROOT, "include", "b200rt.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"\b(b200rt_[a-z0-9_]+)\s*\(", text)))


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = load()
        msg = (L.b200rt_last_error() or b"").decode()
        raise B200RTError(f"{what} failed (rc={rc}): {msg}")


def fptr(a):
    return a.ctypes.data_as(C.POINTER(c_f))
