"""Host-side mirror of the reference's ``Scene`` for the hot path.

``Scene(xml_path, semantics)`` plays the role of ``Scene::Scene(const std::string&)``
(HW3/src/Scene.cpp:314-878, HW1/Scene.cpp:125-348) for the XML elements the hot
path consumes, and ``Scene.render_image(camera_index)`` the role of the thread
fan-out over ``Scene::render_image`` (HW7/src/main.cpp:35-46).  XML text is
parsed here; every float that feeds the renderer is then produced by the C++
host builder (csrc/host_build.cpp) in the reference's arithmetic.

There is no CPU fallback: rendering requires libb200rt.so and a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import B200RTError, SEM_BY_NAME, SEM_HW1, SEM_HW2, SEM_HW3, SEM_HW4, SEM_HW5, SEM_HW6, SEM_HW7, SEM_PPM


def _floats(text, n=None):
    a = np.array((text or "").split(), dtype=np.float64).astype(np.float32)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} numbers, got {a.size}: {text!r}")
    return a


def _big_floats(text):
    return np.fromstring(text or "", sep=" ", dtype=np.float64).astype(np.float32)


def _big_ints(text):
    return np.fromstring(text or "", sep=" ", dtype=np.float64).astype(np.int32)


def _child_text(el, name, default=None):
    c = el.find(name)
    if c is None or c.text is None:
        return default
    return c.text


@dataclass
class MeshSpec:
    material: int
    smooth: bool
    faces: np.ndarray          # (F,3) int32, 0-based, vertex offset applied
    transforms: list = field(default_factory=list)   # [(type, index0), ...]
    velocity: np.ndarray = None
    texture: int = -1
    texture_offset: int = 0        # relative to the (offset-applied) face indices
    radiance: np.ndarray = None    # LightMesh only


@dataclass
class InstanceSpec:
    base_mesh: int
    material: int
    reset_transform: bool
    transforms: list
    velocity: np.ndarray


@dataclass
class SceneSpec:
    semantics: int
    background: np.ndarray
    ambient: np.ndarray
    shadow_eps: float
    max_depth: int
    cameras: list
    point_lights: list
    area_lights: list
    materials: list
    scalings: list
    translations: list
    rotations: list
    vertices: np.ndarray
    meshes: list
    instances: list
    triangles: list
    spheres: list
    texcoords: np.ndarray = None
    textures: list = field(default_factory=list)
    brdfs: list = field(default_factory=list)
    other_lights: list = field(default_factory=list)   # spot / directional / env, in the reference's push order
    light_meshes: list = field(default_factory=list)
    light_spheres: list = field(default_factory=list)
    background_texture: int = -1
    path_tracing: bool = False
    importance_sampling: bool = False
    ppm: dict = None


def _parse_transform_list(text):
    toks = (text or "").split()
    out = []
    i = 0
    # the reference reads `char type; int index;` pairs, so "t1" and "t 1" both parse
    while i < len(toks):
        tok = toks[i]
        if len(tok) > 1:
            out.append((tok[0], int(tok[1:]) - 1))
            i += 1
        else:
            out.append((tok, int(toks[i + 1]) - 1))
            i += 2
    return out


def _load_image_texels(path: str, degamma: bool) -> tuple:
    """stbi_load(..., 3) + the float conversion of Texture::Texture (HW7/src/Texture.cpp:13-24)."""
    from PIL import Image
    img = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
    h, w = img.shape[:2]
    if degamma:
        # std::pow(texture_image[i] / 255.0f, 2.2) * 255.0f : float ratio, DOUBLE pow and product, stored as float
        r = (img.astype(np.float32) / np.float32(255.0)).astype(np.float64)
        tex = (np.power(r, 2.2) * 255.0).astype(np.float32)
    else:
        tex = img.astype(np.float32)
    return np.ascontiguousarray(tex.reshape(-1)), w, h


def _default_perm(index: int) -> list:
    """The reference shuffles 0..15 with a clock-seeded mt19937 (HW7/src/Perlin_noise.cpp:17-20):
    not reproducible, so the host supplies the table."""
    import random
    p = list(range(16))
    random.Random(795 + index).shuffle(p)
    return p


def parse_xml(path_or_text: str, semantics, base_dir: str | None = None) -> SceneSpec:
    """Parse the ceng795 XML dialect (elements on the hot path only); mirrors the section
    order of Scene::Scene (HW7/src/Scene.cpp:449-1591, HW3/src/Scene.cpp:314-878)."""
    import os
    sem = SEM_BY_NAME[semantics] if isinstance(semantics, str) else int(semantics)
    if path_or_text.lstrip().startswith("<"):
        root = ET.fromstring(path_or_text)
        base_dir = base_dir or os.getcwd()
    else:
        root = ET.parse(path_or_text).getroot()
        base_dir = base_dir or os.path.dirname(os.path.abspath(path_or_text))
    bg = _floats(_child_text(root, "BackgroundColor", "0 0 0"), 3)
    seps = float(_floats(_child_text(root, "ShadowRayEpsilon", "0.001"), 1)[0])
    depth = int(float(_child_text(root, "MaxRecursionDepth", "20" if sem == SEM_PPM else "0")))
    ppm = None
    if sem == SEM_PPM:
        # PPM/src/Scene.cpp:398-415
        ppm = dict(photon_count_per_iteration=int(float(_child_text(root, "PhotonCountPerIteration", "8000"))),
                   number_of_iterations=int(float(_child_text(root, "NumberOfIterations", "1000"))))
    cams = []
    for cel in root.find("Cameras").findall("Camera"):
        simple = cel.get("type") == "simple" and sem >= SEM_HW4
        res = _floats(_child_text(cel, "ImageResolution"), 2)
        gaze_txt = _child_text(cel, "Gaze") or _child_text(cel, "GazePoint")
        cam = dict(
            position=_floats(_child_text(cel, "Position"), 3), gaze=_floats(gaze_txt, 3),
            up=_floats(_child_text(cel, "Up"), 3),
            near_plane=np.zeros(4, np.float32) if simple else _floats(_child_text(cel, "NearPlane"), 4),
            near_distance=float(_floats(_child_text(cel, "NearDistance"), 1)[0]),
            width=int(res[0]), height=int(res[1]),
            num_samples=int(float(_child_text(cel, "NumSamples", "1"))),
            focus_distance=float(_floats(_child_text(cel, "FocusDistance", "0"), 1)[0]),
            aperture_size=float(_floats(_child_text(cel, "ApertureSize", "0"), 1)[0]),
            image_name=(_child_text(cel, "ImageName", "out.png") or "out.png").strip(),
            left_handed=(cel.get("handedness") == "left") and sem >= SEM_HW7,
            simple=simple, fov_y=float(_floats(_child_text(cel, "FovY", "0"), 1)[0]) if simple else 0.0,
            tonemap=None)
        tm = cel.find("Tonemap")
        if tm is not None and sem >= SEM_HW5 and (_child_text(tm, "TMO", "") or "").strip() == "Photographic":
            o = _floats(_child_text(tm, "TMOOptions"), 2)
            cam["tonemap"] = dict(key=float(o[0]), burn=float(o[1]),
                                  saturation=float(_floats(_child_text(tm, "Saturation", "1"), 1)[0]))
        cams.append(cam)
    brdfs = []
    bel = root.find("BRDFs")
    if bel is not None and sem >= SEM_HW6:
        brdfs = [None] * len(list(bel))
        kinds = {"OriginalPhong": _lib.BRDF_PHONG, "ModifiedPhong": _lib.BRDF_MODIFIED_PHONG,
                 "OriginalBlinnPhong": _lib.BRDF_BLINN_PHONG, "ModifiedBlinnPhong": _lib.BRDF_MODIFIED_BLINN_PHONG,
                 "TorranceSparrow": _lib.BRDF_TORRANCE_SPARROW}
        for e in bel:
            if e.tag not in kinds:
                continue
            brdfs[int(e.get("id")) - 1] = dict(
                kind=kinds[e.tag], exponent=float(_floats(_child_text(e, "Exponent", "1"), 1)[0]),
                normalized=(e.get("normalized") == "true"),
                refractive_index=float(_floats(_child_text(e, "RefractiveIndex", "1"), 1)[0]))
    lights = root.find("Lights")
    ambient = _floats(_child_text(lights, "AmbientLight", "0 0 0"), 3) if lights is not None else np.zeros(3, np.float32)
    pls, als, others = [], [], []
    if lights is not None:
        pls = [dict(position=_floats(_child_text(l, "Position"), 3), intensity=_floats(_child_text(l, "Intensity"), 3))
               for l in lights.findall("PointLight")]
        if sem >= SEM_HW3:
            for l in lights.findall("AreaLight"):
                als.append(dict(position=_floats(_child_text(l, "Position"), 3),
                                intensity=_floats(_child_text(l, "Intensity"), 3),
                                edge1=_floats(_child_text(l, "EdgeVector1"), 3),
                                edge2=_floats(_child_text(l, "EdgeVector2"), 3)))
        if sem >= SEM_HW5 and sem != SEM_PPM:
            # HW5/HW6 advance these two loops with FirstChildElement instead of NextSiblingElement, so they stop
            # after the first SpotLight / DirectionalLight (HW6/src/Scene.cpp:661,678); HW7 fixed it (:385,404)
            first_only = sem < SEM_HW7
            for l in lights.findall("SpotLight")[:1 if first_only else None]:
                others.append(dict(kind=_lib.LIGHT_SPOT, position=_floats(_child_text(l, "Position"), 3),
                                   intensity=_floats(_child_text(l, "Intensity"), 3),
                                   direction=_floats(_child_text(l, "Direction"), 3),
                                   coverage=float(_floats(_child_text(l, "CoverageAngle"), 1)[0]),
                                   falloff=float(_floats(_child_text(l, "FalloffAngle"), 1)[0])))
            for l in lights.findall("DirectionalLight")[:1 if first_only else None]:
                others.append(dict(kind=_lib.LIGHT_DIRECTIONAL, direction=_floats(_child_text(l, "Direction"), 3),
                                   radiance=_floats(_child_text(l, "Radiance"), 3)))
            l = lights.find("SphericalDirectionalLight")
            if l is not None and sem >= SEM_HW7:
                others.append(dict(kind=_lib.LIGHT_ENV, envmap=(_child_text(l, "EnvMapName") or "").strip()))
    mats = []
    for m in root.find("Materials").findall("Material"):
        mats.append(dict(
            ambient=_floats(_child_text(m, "AmbientReflectance", "0 0 0"), 3),
            diffuse=_floats(_child_text(m, "DiffuseReflectance", "0 0 0"), 3),
            specular=_floats(_child_text(m, "SpecularReflectance", "0 0 0"), 3),
            mirror=_floats(_child_text(m, "MirrorReflectance", "0 0 0"), 3),
            roughness=float(_floats(_child_text(m, "Roughness", "0"), 1)[0]) if sem >= SEM_HW3 else 0.0,
            phong_exponent=float(_floats(_child_text(m, "PhongExponent", "1"), 1)[0]),
            transparency=_floats(_child_text(m, "Transparency", "0 0 0"), 3) if sem >= SEM_HW2 else np.zeros(3, np.float32),
            refraction_index=float(_floats(_child_text(m, "RefractionIndex", "1.0"), 1)[0]) if sem >= SEM_HW2 else 1.0,
            brdf=(int(m.get("BRDF", "0")) - 1) if sem >= SEM_HW6 else -1,
            degamma=(m.get("degamma") == "true") and sem >= SEM_HW5))
    scal, tran, rot = [], [], []
    tel = root.find("Transformations")
    if tel is not None and sem >= SEM_HW3:
        scal = [_floats(e.text, 3) for e in tel.findall("Scaling")]
        tran = [_floats(e.text, 3) for e in tel.findall("Translation")]
        rot = [_floats(e.text, 4) for e in tel.findall("Rotation")]
    vel_ = root.find("VertexData")
    if vel_.get("binaryFile"):
        raise B200RTError("binaryFile vertex data is host I/O (out of scope, SURVEY §2b)")
    verts = _big_floats(vel_.text).reshape(-1, 3)
    texcoords = None
    tcel = root.find("TexCoordData")
    if tcel is not None and sem >= SEM_HW4:
        texcoords = np.ascontiguousarray(_big_floats(tcel.text).reshape(-1, 2), np.float32)
    objs = root.find("Objects")
    meshes, insts, tris, sphs, lmeshes, lsphs = [], [], [], [], [], []

    def xf_list(el):
        return _parse_transform_list(_child_text(el, "Transformations", "")) if sem >= SEM_HW3 else []

    def tex_id(el):
        t = _child_text(el, "Texture")
        return (int(float(t)) - 1) if (t is not None and sem >= SEM_HW4) else -1

    for mel in objs.findall("Mesh"):
        fel = mel.find("Faces")
        if fel.get("plyFile") or fel.get("binaryFile"):
            raise B200RTError("plyFile / binaryFile meshes are host I/O (out of scope, SURVEY §2b)")
        off = int(fel.get("vertexOffset", "0")) if sem >= SEM_HW3 else 0
        toff = int(fel.get("textureOffset", "0")) if sem >= SEM_HW4 else 0
        faces = _big_ints(fel.text).reshape(-1, 3) - 1 + off
        vel = _floats(_child_text(mel, "MotionBlur", "0 0 0"), 3) if sem >= SEM_HW3 else np.zeros(3, np.float32)
        meshes.append(MeshSpec(
            material=int(float(_child_text(mel, "Material"))) - 1,
            smooth=(mel.get("shadingMode") == "smooth") and sem >= SEM_HW3,
            faces=np.ascontiguousarray(faces, dtype=np.int32),
            transforms=xf_list(mel), velocity=vel, texture=tex_id(mel),
            # the reference indexes texcoords with (face index + textureOffset), faces BEFORE the vertex offset
            texture_offset=toff - off))
    if sem >= SEM_HW3:
        for iel in objs.findall("MeshInstance"):
            insts.append(InstanceSpec(
                base_mesh=int(iel.get("baseMeshId")) - 1,
                material=int(float(_child_text(iel, "Material"))) - 1,
                reset_transform=(iel.get("resetTransform") == "true"),
                transforms=_parse_transform_list(_child_text(iel, "Transformations", "")),
                velocity=_floats(_child_text(iel, "MotionBlur", "0 0 0"), 3)))
    if sem == SEM_HW7:
        for lel in objs.findall("LightMesh"):
            fel = lel.find("Faces")
            off = int(fel.get("vertexOffset", "0"))
            faces = _big_ints(fel.text).reshape(-1, 3) - 1 + off
            lmeshes.append(MeshSpec(material=int(float(_child_text(lel, "Material"))) - 1, smooth=False,
                                    faces=np.ascontiguousarray(faces, dtype=np.int32), transforms=xf_list(lel),
                                    velocity=np.zeros(3, np.float32),
                                    radiance=_floats(_child_text(lel, "Radiance"), 3)))
    for tel_ in objs.findall("Triangle"):
        idx = _floats(_child_text(tel_, "Indices"), 3).astype(np.int32) - 1
        tris.append(dict(material=int(float(_child_text(tel_, "Material"))) - 1, idx=idx, transforms=xf_list(tel_)))
    if sem == SEM_HW7:
        for sel in objs.findall("LightSphere"):
            lsphs.append(dict(material=int(float(_child_text(sel, "Material"))) - 1,
                              center=int(float(_child_text(sel, "Center"))) - 1,
                              radius=float(_floats(_child_text(sel, "Radius"), 1)[0]),
                              transforms=xf_list(sel), radiance=_floats(_child_text(sel, "Radiance"), 3)))
    for sel in objs.findall("Sphere"):
        sphs.append(dict(material=int(float(_child_text(sel, "Material"))) - 1,
                         center=int(float(_child_text(sel, "Center"))) - 1,
                         radius=float(_floats(_child_text(sel, "Radius"), 1)[0]),
                         transforms=xf_list(sel), texture=tex_id(sel),
                         velocity=_floats(_child_text(sel, "MotionBlur", "0 0 0"), 3) if sem >= SEM_HW3 else np.zeros(3, np.float32)))
    textures = []
    txs = root.find("Textures")
    names = {"nearest": _lib.TEX_NEAREST, "bilinear": _lib.TEX_BILINEAR}
    if txs is not None and sem >= SEM_HW4:
        for k, tel_ in enumerate(txs.findall("Texture")):
            name = (_child_text(tel_, "ImageName") or "").strip()
            interp = (_child_text(tel_, "Interpolation", "bilinear") or "").strip()
            decal = (_child_text(tel_, "DecalMode", "blend_kd") or "").strip()
            app = (_child_text(tel_, "Appearance", "repeat") or "").strip()
            t = dict(name=name,
                     interp=_lib.TEX_NEAREST if interp == "nearest" else _lib.TEX_BILINEAR,
                     decal=_lib.TEX_REPLACE_KD if decal == "replace_kd" else (_lib.TEX_BLEND_KD if decal == "blend_kd" else _lib.TEX_REPLACE_ALL),
                     normalizer=float(_floats(_child_text(tel_, "Normalizer", "255.0"), 1)[0]),
                     scaling=float(_floats(_child_text(tel_, "ScalingFactor", "1.0"), 1)[0]),
                     is_bump=(tel_.get("bumpmap") == "true"),
                     bump_multiplier=float(tel_.get("bumpmapMultiplier", "1.0")),
                     degamma=(tel_.get("degamma") == "true") and sem >= SEM_HW7)
            if name == "perlin":
                t["kind"] = _lib.TEX_PERLIN
                t["appearance"] = _lib.TEX_VEIN if app == "vein" else _lib.TEX_PATCH
                ptxt = _child_text(tel_, "Permutation")   # our extension: the host object's table
                t["perm"] = [int(x) for x in ptxt.split()] if ptxt else _default_perm(k)
            else:
                t["kind"] = _lib.TEX_IMAGE
                t["appearance"] = _lib.TEX_REPEAT if app == "repeat" else _lib.TEX_CLAMP
                t["texels"], t["width"], t["height"] = _load_image_texels(os.path.join(base_dir, name), t["degamma"])
            textures.append(t)
    bg_tex = -1
    bgt = _child_text(root, "BackgroundTexture")
    if bgt and sem >= SEM_HW4:
        # Texture(bg, "bilinear", "replace_all", "clamp", 255, 1.0f, false, 1.0f, false) — HW7/src/Scene.cpp:479-480
        t = dict(name=bgt.strip(), kind=_lib.TEX_IMAGE, interp=_lib.TEX_BILINEAR, decal=_lib.TEX_REPLACE_ALL,
                 appearance=_lib.TEX_CLAMP, normalizer=255.0, scaling=1.0, is_bump=False, bump_multiplier=1.0,
                 degamma=False)
        t["texels"], t["width"], t["height"] = _load_image_texels(os.path.join(base_dir, t["name"]), False)
        bg_tex = len(textures)
        textures.append(t)
    for o in others:
        if o["kind"] == _lib.LIGHT_ENV:
            o["texels"], o["width"], o["height"] = _load_envmap(os.path.join(base_dir, o["envmap"]))
    integ = (_child_text(root, "Integrator", "") or "").strip()
    path_tracing = (integ == "PathTracing") and sem == SEM_HW7
    importance = path_tracing and (_child_text(root, "IntegratorParams", "") or "").strip() == "ImportanceSampling"
    if sem == SEM_HW1 and tris:
        # HW1/Scene.cpp:320 passes `mesh.material_id` (the last mesh's) for every stand-alone
        # triangle; without any mesh the reference reads an uninitialised int.
        if meshes:
            for t in tris:
                t["material"] = meshes[-1].material
    return SceneSpec(sem, bg, ambient, seps, depth, cams, pls, als, mats, scal, tran, rot,
                     np.ascontiguousarray(verts, dtype=np.float32), meshes, insts, tris, sphs,
                     texcoords=texcoords, textures=textures, brdfs=brdfs, other_lights=others,
                     light_meshes=lmeshes, light_spheres=lsphs, background_texture=bg_tex,
                     path_tracing=path_tracing, importance_sampling=importance, ppm=ppm)


def _load_envmap(path: str):
    """The reference loads an EXR (tinyexr, host I/O).  This mirror accepts a .npy float array
    [H, W, 3] (or an 8-bit image scaled to 0..1) — the device side only sees RGB floats."""
    if path.endswith(".npy"):
        a = np.load(path).astype(np.float32)
    elif path.endswith(".exr"):
        from .image import read_exr_f32      # uncompressed float32 scanline files only (tests); real decoding is host I/O
        a = read_exr_f32(path)
    else:
        from PIL import Image
        a = np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / np.float32(255.0)
    h, w = a.shape[:2]
    return np.ascontiguousarray(a[..., :3].reshape(-1)), w, h


def _arr3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class HostScene:
    """Owns a b200rt_builder and the flattened B200SceneDesc it produced."""

    def __init__(self, spec: SceneSpec):
        L = _lib.load()
        self._L = L
        self.spec = spec
        self._b = L.b200rt_builder_create(spec.semantics)
        if not self._b:
            raise B200RTError("b200rt_builder_create failed")
        self._keep = []
        fp = _lib.fptr
        bg = np.ascontiguousarray(spec.background, np.float32)
        am = np.ascontiguousarray(spec.ambient, np.float32)
        _lib.check(L.b200rt_builder_set_globals(self._b, fp(bg), fp(am), spec.shadow_eps, spec.max_depth), "set_globals")
        for c in spec.cameras:
            ci = _lib.CameraIn()
            ci.position = _arr3(c["position"]); ci.gaze = _arr3(c["gaze"]); ci.up = _arr3(c["up"])
            ci.near_plane = (C.c_float * 4)(*[float(x) for x in c["near_plane"]])
            ci.near_distance = c["near_distance"]; ci.width = c["width"]; ci.height = c["height"]
            ci.num_samples = c["num_samples"]; ci.focus_distance = c["focus_distance"]
            ci.aperture_size = c["aperture_size"]
            ci.left_handed = int(bool(c.get("left_handed", False)))
            ci.simple = int(bool(c.get("simple", False)))
            ci.fov_y = float(c.get("fov_y", 0.0))
            rc = L.b200rt_builder_add_camera(self._b, C.byref(ci))
            if rc < 0:
                _lib.check(rc, "add_camera")
        for m in spec.materials:
            mm = _lib.Material()
            mm.ambient = _arr3(m["ambient"]); mm.diffuse = _arr3(m["diffuse"]); mm.specular = _arr3(m["specular"])
            mm.mirror = _arr3(m["mirror"]); mm.transparency = _arr3(m["transparency"])
            mm.roughness = m["roughness"]; mm.refraction_index = m["refraction_index"]
            mm.phong_exponent = m["phong_exponent"]; mm.brdf = int(m.get("brdf", -1))
            mm.degamma = int(bool(m.get("degamma", False)))
            L.b200rt_builder_add_material(self._b, C.byref(mm))
        # textures / BRDFs / texcoords (HW4+) — before the lights so that the env map can refer to a texture
        n_tex = 0
        for t in spec.textures:
            n_tex = self._add_texture(t) + 1
        for r in spec.brdfs:
            rr = _lib.Brdf()
            if r is not None:
                rr.kind = r["kind"]; rr.exponent = r["exponent"]; rr.normalized = int(r["normalized"])
                rr.refractive_index = r["refractive_index"]
            L.b200rt_builder_add_brdf(self._b, C.byref(rr))
        if spec.texcoords is not None and len(spec.texcoords):
            tc = np.ascontiguousarray(spec.texcoords, np.float32)
            self._keep.append(tc)
            _lib.check(L.b200rt_builder_set_texcoords(self._b, tc.ctypes.data, tc.shape[0]), "set_texcoords")
        if spec.background_texture >= 0:
            L.b200rt_builder_set_background_texture(self._b, spec.background_texture)
        if spec.semantics == SEM_HW7:
            L.b200rt_builder_set_integrator(self._b, int(spec.path_tracing), int(spec.importance_sampling))
        for l in spec.point_lights:
            ll = _lib.Light(); ll.kind = _lib.LIGHT_POINT
            ll.position = _arr3(l["position"]); ll.intensity = _arr3(l["intensity"]); ll.object = -1
            L.b200rt_builder_add_light(self._b, C.byref(ll))
        for l in spec.area_lights:
            ll = _lib.Light(); ll.kind = _lib.LIGHT_AREA
            ll.position = _arr3(l["position"]); ll.intensity = _arr3(l["intensity"])
            ll.edge1 = _arr3(l["edge1"]); ll.edge2 = _arr3(l["edge2"]); ll.object = -1
            L.b200rt_builder_add_light(self._b, C.byref(ll))
        for l in spec.other_lights:
            ll = _lib.Light(); ll.kind = l["kind"]; ll.object = -1
            if l["kind"] == _lib.LIGHT_SPOT:
                ll.position = _arr3(l["position"]); ll.intensity = _arr3(l["intensity"])
                ll.edge1 = _arr3(l["direction"]); ll.edge2 = _arr3([l["coverage"], l["falloff"], 0.0])
            elif l["kind"] == _lib.LIGHT_DIRECTIONAL:
                ll.edge1 = _arr3(l["direction"]); ll.intensity = _arr3(l["radiance"])
            elif l["kind"] == _lib.LIGHT_ENV:
                t = dict(kind=_lib.TEX_IMAGE, interp=_lib.TEX_NEAREST, decal=_lib.TEX_REPLACE_ALL,
                         appearance=_lib.TEX_CLAMP, normalizer=1.0, scaling=1.0, is_bump=False, bump_multiplier=1.0,
                         texels=l["texels"], width=l["width"], height=l["height"])
                ll.object = self._add_texture(t)
            L.b200rt_builder_add_light(self._b, C.byref(ll))
        v = np.ascontiguousarray(spec.vertices, np.float32)
        self._keep.append(v)
        _lib.check(L.b200rt_builder_set_vertices(self._b, v.ctypes.data, v.shape[0]), "set_vertices")
        for m in spec.meshes:
            xf = self._compose(m.transforms, None)
            vel = np.ascontiguousarray(m.velocity, np.float32)
            rc = L.b200rt_builder_add_mesh(self._b, m.material, int(m.smooth), m.faces.ctypes.data,
                                           m.faces.shape[0], fp(xf), fp(vel))
            if rc < 0:
                _lib.check(rc, "add_mesh")
            if m.texture >= 0:
                _lib.check(L.b200rt_builder_set_mesh_texture(self._b, rc, m.texture, m.texture_offset), "set_mesh_texture")
        for m in spec.light_meshes:
            xf = self._compose(m.transforms, None)
            rad = np.ascontiguousarray(m.radiance, np.float32)
            rc = L.b200rt_builder_add_light_mesh(self._b, m.material, m.faces.ctypes.data, m.faces.shape[0], fp(xf), fp(rad))
            if rc < 0:
                _lib.check(rc, "add_light_mesh")
        for s_ in spec.light_spheres:
            xf = self._compose(s_["transforms"], None)
            rad = np.ascontiguousarray(s_["radiance"], np.float32)
            rc = L.b200rt_builder_add_light_sphere(self._b, s_["material"], s_["center"], s_["radius"], fp(xf), fp(rad))
            if rc < 0:
                _lib.check(rc, "add_light_sphere")
        for i in spec.instances:
            base = None if i.reset_transform else self._compose(spec.meshes[i.base_mesh].transforms, None)
            xf = self._compose(i.transforms, base)
            vel = np.ascontiguousarray(i.velocity, np.float32)
            rc = L.b200rt_builder_add_instance(self._b, i.base_mesh, i.material, fp(xf), fp(vel))
            if rc < 0:
                _lib.check(rc, "add_instance")
        for t in spec.triangles:
            xf = self._compose(t["transforms"], None)
            idx = (C.c_int32 * 3)(*[int(x) for x in t["idx"]])
            rc = L.b200rt_builder_add_triangle(self._b, t["material"], idx, fp(xf))
            if rc < 0:
                _lib.check(rc, "add_triangle")
        for s in spec.spheres:
            xf = self._compose(s["transforms"], None)
            vel = np.ascontiguousarray(s["velocity"], np.float32)
            rc = L.b200rt_builder_add_sphere(self._b, s["material"], s["center"], s["radius"], fp(xf), fp(vel))
            if rc < 0:
                _lib.check(rc, "add_sphere")
            if s.get("texture", -1) >= 0:
                _lib.check(L.b200rt_builder_set_sphere_texture(self._b, rc, s["texture"]), "set_sphere_texture")
        out = C.POINTER(_lib.SceneDesc)()
        _lib.check(L.b200rt_builder_finish(self._b, C.byref(out)), "builder_finish")
        self.desc_ptr = out
        self.desc = out.contents

    def _add_texture(self, t) -> int:
        tt = _lib.Texture()
        tt.kind = t["kind"]; tt.interp = t["interp"]; tt.decal = t["decal"]; tt.appearance = t["appearance"]
        tt.normalizer = t["normalizer"]; tt.scaling = t["scaling"]; tt.bump_multiplier = t["bump_multiplier"]
        tt.is_bump = int(bool(t["is_bump"]))
        tex = None
        if t["kind"] == _lib.TEX_IMAGE:
            tt.width = t["width"]; tt.height = t["height"]
            tex = np.ascontiguousarray(t["texels"], np.float32)
            self._keep.append(tex)
        else:
            tt.perm = (C.c_int32 * 16)(*[int(x) for x in t["perm"]])
        rc = self._L.b200rt_builder_add_texture(self._b, C.byref(tt), tex.ctypes.data if tex is not None else None)
        if rc < 0:
            _lib.check(rc, "add_texture")
        return rc

    def _compose(self, ops, base):
        """`M = T_k * M` in list order — HW3/src/Scene.cpp:636-657."""
        L = self._L
        fp = _lib.fptr
        m = np.zeros(16, np.float32)
        if base is None:
            L.b200rt_xf_identity(fp(m))
        else:
            m[:] = base
        t = np.zeros(16, np.float32)
        o = np.zeros(16, np.float32)
        for kind, idx in ops:
            if kind == "s":
                s = self.spec.scalings[idx]
                L.b200rt_xf_scaling(float(s[0]), float(s[1]), float(s[2]), fp(t))
            elif kind == "t":
                s = self.spec.translations[idx]
                L.b200rt_xf_translation(float(s[0]), float(s[1]), float(s[2]), fp(t))
            elif kind == "r":
                s = self.spec.rotations[idx]
                L.b200rt_xf_rotation(float(s[0]), float(s[1]), float(s[2]), float(s[3]), fp(t))
            else:
                continue  # the reference's switch ignores unknown letters
            L.b200rt_xf_mul(fp(t), fp(m), fp(o))
            m[:] = o
        return m

    # -- numpy views of the flattened description (tests / tools) ---------------
    def nodes(self):
        d = self.desc
        if d.n_nodes == 0:
            return np.zeros((0, 16), np.uint32)
        return np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint32)), shape=(d.n_nodes, 16))

    def tri_geom(self):
        d = self.desc
        return np.ctypeslib.as_array(d.tri_geom, shape=(d.n_tris, 12)) if d.n_tris else np.zeros((0, 12), np.float32)

    def move_instance(self, instance: int, xf, velocity=(0.0, 0.0, 0.0)) -> None:
        """New 4x4 matrix (row-major, 16 floats) and velocity of a top-level mesh instance; mesh m's base instance is
        instance m, the XML's MeshInstance k is instance len(meshes) + k.  Takes effect at refit()."""
        m = np.ascontiguousarray(xf, np.float32).reshape(16)
        v = np.ascontiguousarray(velocity, np.float32).reshape(3)
        _lib.check(self._L.b200rt_builder_set_instance_transform(self._b, int(instance), _lib.fptr(m), _lib.fptr(v)),
                   "b200rt_builder_set_instance_transform")

    def move_sphere(self, sphere: int, xf, velocity=(0.0, 0.0, 0.0)) -> None:
        m = np.ascontiguousarray(xf, np.float32).reshape(16)
        v = np.ascontiguousarray(velocity, np.float32).reshape(3)
        _lib.check(self._L.b200rt_builder_set_sphere_transform(self._b, int(sphere), _lib.fptr(m), _lib.fptr(v)),
                   "b200rt_builder_set_sphere_transform")

    def refit(self) -> None:
        """Re-runs the top-level half of the flattening (transforms, world boxes, Scene::bvh, ranks) after
        move_instance / move_sphere; mesh trees and triangle records are reused."""
        out = C.POINTER(_lib.SceneDesc)()
        _lib.check(self._L.b200rt_builder_refit(self._b, C.byref(out)), "b200rt_builder_refit")
        self.desc_ptr = out
        self.desc = out.contents

    def translation(self, x, y, z):
        m = np.zeros(16, np.float32)
        self._L.b200rt_xf_translation(float(x), float(y), float(z), _lib.fptr(m))
        return m

    def xf_mul(self, a, b_):
        o = np.zeros(16, np.float32)
        self._L.b200rt_xf_mul(_lib.fptr(np.ascontiguousarray(a, np.float32)), _lib.fptr(np.ascontiguousarray(b_, np.float32)),
                              _lib.fptr(o))
        return o

    def save(self, path: str) -> None:
        """Binary cache of the flattened description (include/b200rt.h: b200rt_scene_desc_save)."""
        _lib.check(self._L.b200rt_scene_desc_save(self.desc_ptr, os.fsencode(path)), "b200rt_scene_desc_save")

    def close(self):
        if getattr(self, "_b", None):
            self._L.b200rt_builder_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CachedHostScene:
    """A flattened description read back from a binary cache: the role of HostScene without parse or BVH build."""

    def __init__(self, path: str):
        self._L = _lib.load()
        out = C.POINTER(_lib.SceneDesc)()
        _lib.check(self._L.b200rt_scene_desc_load(os.fsencode(path), C.byref(out)), "b200rt_scene_desc_load")
        self.desc_ptr = out
        self.desc = out.contents
        d = self.desc
        cams = [d.cameras[i] for i in range(d.n_cameras)]
        # only what the device-facing methods of Scene read: semantics and the image size of each camera
        self.spec = SceneSpec(int(d.semantics), None, None, float(d.shadow_eps), int(d.max_depth),
                              [dict(width=int(c.width), height=int(c.height)) for c in cams],
                              [], [], [], [], [], [], None, [], [], [], [])

    def save(self, path: str) -> None:
        _lib.check(self._L.b200rt_scene_desc_save(self.desc_ptr, os.fsencode(path)), "b200rt_scene_desc_save")

    def close(self):
        if getattr(self, "desc_ptr", None):
            self._L.b200rt_scene_desc_free(self.desc_ptr)
            self.desc_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Reference-facing object: ``Scene(xml).render_image(camera_index)``."""

    @classmethod
    def from_cache(cls, path: str, device: int = 0, upload: bool = True) -> "Scene":
        """The same object from a file written by ``scene.host.save(path)``: no XML, no BVH build."""
        self = cls.__new__(cls)
        self.host = CachedHostScene(path)
        self.spec = self.host.spec
        self.semantics = self.spec.semantics
        self.device = device
        self._scene = None
        self._L = _lib.load()
        if upload:
            self.upload()
        return self

    def __init__(self, xml: str, semantics="hw3", device: int = 0, upload: bool = True, base_dir: str | None = None):
        self.spec = parse_xml(xml, semantics, base_dir=base_dir)
        self.host = HostScene(self.spec)
        self.semantics = self.spec.semantics
        self.device = device
        self._scene = None
        self._L = _lib.load()
        if upload:
            self.upload()

    @property
    def cameras(self):
        return self.spec.cameras

    def upload(self):
        if self._scene is not None:
            return
        h = C.c_void_p()
        _lib.check(self._L.b200rt_scene_create(self.host.desc_ptr, self.device, C.byref(h)), "b200rt_scene_create")
        self._scene = h

    def _opts(self, **kw):
        o = _lib.RenderOpts()
        o.row_stride = 1
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def refit(self) -> float:
        """After host.move_instance / host.move_sphere: refits the host description and copies its top-level half
        over the resident scene (b200rt_scene_refit).  Returns the device time of the copy in ms."""
        self.host.refit()
        self.upload()
        ms = C.c_float(0.0)
        _lib.check(self._L.b200rt_scene_refit(self._scene, self.host.desc_ptr, C.byref(ms)), "b200rt_scene_refit")
        return float(ms.value)

    def build_bvh_device(self) -> float:
        """Rebuilds the device-walked tree of every mesh ON the device (LBVH); returns the build time in ms."""
        self.upload()
        ms = C.c_float(0.0)
        _lib.check(self._L.b200rt_scene_build_bvh_device(self._scene, C.byref(ms)), "b200rt_scene_build_bvh_device")
        return float(ms.value)

    def image_size(self, camera_index=0):
        c = self.spec.cameras[camera_index]
        return c["width"], c["height"]

    def render_image(self, camera_index: int = 0, out: np.ndarray | None = None, **opts):
        """Returns (rgbw[H,W,4] float32, RenderStats).  rgbw = Pixel{color, weight}."""
        self.upload()
        w, h = self.image_size(camera_index)
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        st = _lib.RenderStats()
        o = self._opts(**opts)
        _lib.check(self._L.b200rt_render(self._scene, camera_index, C.byref(o), out.ctypes.data, C.byref(st)),
                   "b200rt_render")
        return out, st

    def render_image_async(self, camera_index: int, out: np.ndarray, **opts):
        """render_image whose host copy is only queued: `out` (page-locked memory, kept alive by the caller) is valid
        after render_wait(); the copy overlaps the next render of this scene.  Returns RenderStats."""
        self.upload()
        st = _lib.RenderStats()
        o = self._opts(**opts)
        _lib.check(self._L.b200rt_render_async(self._scene, camera_index, C.byref(o), out.ctypes.data, C.byref(st)),
                   "b200rt_render_async")
        return st

    def render_wait(self):
        _lib.check(self._L.b200rt_render_wait(self._scene), "b200rt_render_wait")

    def count_lit(self, enable: bool = True):
        """Parity hook: later renders count, per pixel and light, the shadow tests that came out lit."""
        self.upload()
        _lib.check(self._L.b200rt_debug_lit_counts(self._scene, 1 if enable else 0), "b200rt_debug_lit_counts")

    def lit_counts(self, camera_index: int = 0) -> np.ndarray:
        w, h = self.image_size(camera_index)
        out = np.zeros((h, w, max(1, self.host.desc.n_lights)), np.int32)
        _lib.check(self._L.b200rt_debug_get_lit_counts(self._scene, camera_index, out.ctypes.data),
                   "b200rt_debug_get_lit_counts")
        return out

    def render_device(self, camera_index: int = 0, **opts):
        self.upload()
        st = _lib.RenderStats()
        o = self._opts(**opts)
        p = C.c_void_p()
        _lib.check(self._L.b200rt_render_device(self._scene, camera_index, C.byref(o), C.byref(p), C.byref(st)),
                   "b200rt_render_device")
        return p.value, st

    def primary_ids(self, camera_index: int = 0, **opts):
        self.upload()
        w, h = self.image_size(camera_index)
        ids = np.zeros((h, w, 2), np.int32)
        t = np.zeros((h, w), np.float32)
        o = self._opts(**opts)
        _lib.check(self._L.b200rt_primary_ids(self._scene, camera_index, C.byref(o), ids.ctypes.data, t.ctypes.data),
                   "b200rt_primary_ids")
        return ids, t

    def trace_rays(self, rays: np.ndarray, exact: bool = False):
        self.upload()
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        ids = np.zeros((n, 2), np.int32)
        t = np.zeros(n, np.float32)
        st = _lib.RenderStats()
        _lib.check(self._L.b200rt_trace_rays(self._scene, rays.ctypes.data, n, int(exact), ids.ctypes.data,
                                             t.ctypes.data, C.byref(st)), "b200rt_trace_rays")
        return ids, t, st

    def tonemap(self, camera_index: int = 0, key: float | None = None, burn: float | None = None,
                saturation: float | None = None) -> np.ndarray:
        """Photographic TMO + sRGB of the frame the last render_image / render_device left in HBM
        (Camera::get_tmo()->apply_tmo + main's gamma loop, HW7/src/main.cpp:59-91).  uint8[H,W,3]."""
        tm = self.spec.cameras[camera_index].get("tonemap") or {}
        key = tm.get("key", 0.18) if key is None else key
        burn = tm.get("burn", 1.0) if burn is None else burn
        saturation = tm.get("saturation", 1.0) if saturation is None else saturation
        w, h = self.image_size(camera_index)
        out = np.zeros((h, w, 3), np.uint8)
        _lib.check(self._L.b200rt_tonemap_photographic(self._scene, camera_index, key, burn, saturation,
                                                       out.ctypes.data), "b200rt_tonemap_photographic")
        return out

    # ---- PPM: the calls PPM/src/main.cpp makes on Scene (PPM/include/Scene.h:48-55) --------------------
    def eye_trace_lines(self, camera_index: int = 0, seed: int = 0):
        """reset_hash_grid + eye_trace_lines + build_hash_grid.  Returns PpmStats."""
        self.upload()
        st = _lib.PpmStats()
        _lib.check(self._L.b200rt_ppm_eye_pass(self._scene, camera_index, seed, C.byref(st)), "b200rt_ppm_eye_pass")
        self._ppm_camera = camera_index
        return st

    def trace_n_photons(self, n: int, first_photon: int = 0, seed: int = 0):
        o = _lib.PpmOpts()
        o.photons = int(n); o.first_photon = int(first_photon); o.seed = int(seed)
        st = _lib.PpmStats()
        _lib.check(self._L.b200rt_ppm_trace_photons(self._scene, C.byref(o), C.byref(st)), "b200rt_ppm_trace_photons")
        return st

    def density_estimation(self, total_num_of_photons: float, out: np.ndarray | None = None):
        w, h = self.image_size(self._ppm_camera)
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        _lib.check(self._L.b200rt_ppm_density(self._scene, float(total_num_of_photons), out.ctypes.data), "b200rt_ppm_density")
        return out

    def ppm_export(self):
        """(device pointer of float4 {S.rgb, n} per hit point, count) — the additive state for the all-reduce."""
        p = C.c_void_p()
        n = C.c_int64()
        _lib.check(self._L.b200rt_ppm_export(self._scene, C.byref(p), C.byref(n)), "b200rt_ppm_export")
        return p.value, n.value

    def density_estimation_from(self, d_S_n: int, total_num_of_photons: float, out: np.ndarray | None = None):
        w, h = self.image_size(self._ppm_camera)
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        _lib.check(self._L.b200rt_ppm_density_from(self._scene, C.c_void_p(d_S_n), float(total_num_of_photons),
                                                   out.ctypes.data), "b200rt_ppm_density_from")
        return out

    def hit_points(self):
        n = C.c_int64()
        _lib.check(self._L.b200rt_ppm_get_hit_points(self._scene, None, 0, C.byref(n)), "b200rt_ppm_get_hit_points")
        out = np.zeros((n.value, 16), np.float32)
        if n.value:
            _lib.check(self._L.b200rt_ppm_get_hit_points(self._scene, out.ctypes.data, n.value, C.byref(n)),
                       "b200rt_ppm_get_hit_points")
        return out

    def render_ppm(self, camera_index: int = 0, photons: int | None = None, seed: int = 0, total_num_of_photons=None):
        """One PPM frame like PPM/src/main.cpp:33-102.  `total_num_of_photons` defaults to the number actually
        emitted (the reference passes count*count/threads*threads, SURVEY Appendix B.7 — pass that explicitly
        to reproduce its brightness)."""
        ppm = self.spec.ppm or dict(photon_count_per_iteration=8000, number_of_iterations=1000)
        if photons is None:
            photons = ppm["photon_count_per_iteration"] * ppm["number_of_iterations"]
        s0 = self.eye_trace_lines(camera_index, seed)
        s1 = self.trace_n_photons(photons, 0, seed)
        rgbw = self.density_estimation(photons if total_num_of_photons is None else total_num_of_photons)
        return rgbw, s0, s1

    # ---- multi-GPU, one process per GPU (include/b200rt.h "multi-GPU"): the handle is plain bytes ---------
    def frame_export(self, camera_index: int = 0) -> bytes:
        """Owner: the Pixel array of the camera as a handle other ranks attach to."""
        self.upload()
        f = _lib.SharedFrame()
        _lib.check(self._L.b200rt_frame_export(self._scene, camera_index, C.byref(f)), "b200rt_frame_export")
        return bytes(f)

    def frame_attach(self, handle: bytes):
        self.upload()
        f = _lib.SharedFrame.from_buffer_copy(handle)
        _lib.check(self._L.b200rt_frame_attach(self._scene, C.byref(f)), "b200rt_frame_attach")

    def frame_detach(self):
        _lib.check(self._L.b200rt_frame_detach(self._scene), "b200rt_frame_detach")

    def frame_begin(self):
        _lib.check(self._L.b200rt_frame_begin(self._scene), "b200rt_frame_begin")

    def multi_share(self, part: int, parts: int, camera_index: int = 0, **opts) -> dict:
        """RenderOpts of share `part` of `parts`, as b200rt_render_multi deals them (8-row bands round-robin, or
        batches of sample strata)."""
        o = self._opts(**opts)
        _lib.check(self._L.b200rt_multi_share(self._scene, camera_index, part, parts, C.byref(o)), "b200rt_multi_share")
        return {k: getattr(o, k) for k, _ in _lib.RenderOpts._fields_ if k != "pad"}

    def render_shared(self, camera_index: int = 0, **opts):
        st = _lib.RenderStats()
        o = self._opts(**opts)
        _lib.check(self._L.b200rt_render_shared(self._scene, camera_index, C.byref(o), C.byref(st)),
                   "b200rt_render_shared")
        return st

    def frame_collect_async(self, parts: int, out: np.ndarray):
        """frame_collect whose host copy is only queued (render_wait() waits for it)."""
        _lib.check(self._L.b200rt_frame_collect_async(self._scene, parts, out.ctypes.data), "b200rt_frame_collect_async")

    def frame_collect(self, parts: int, out: np.ndarray | None = None):
        """Owner: waits on the device for `parts` shares; copies the frame into `out` when given.  Returns the frame's
        device pointer."""
        p = C.c_void_p()
        _lib.check(self._L.b200rt_frame_collect(self._scene, parts, out.ctypes.data if out is not None else None,
                                                C.byref(p)), "b200rt_frame_collect")
        return p.value

    def ppm_set_share(self, part: int, parts: int):
        """Before eye_trace_lines: this scene traces share `part` of `parts` of every photon pass."""
        self.upload()
        _lib.check(self._L.b200rt_ppm_set_share(self._scene, part, parts), "b200rt_ppm_set_share")

    def ppm_counts_delta(self):
        """(device pointer of uint32 per hit point, count): photons this GPU accepted since the last sync — sum it
        over the ranks in place (all-reduce), then ppm_counts_apply()."""
        p = C.c_void_p()
        n = C.c_int64()
        _lib.check(self._L.b200rt_ppm_counts_delta(self._scene, C.byref(p), C.byref(n)), "b200rt_ppm_counts_delta")
        return p.value, n.value

    def ppm_counts_apply(self):
        _lib.check(self._L.b200rt_ppm_counts_apply(self._scene), "b200rt_ppm_counts_apply")

    def close(self):
        if self._scene is not None:
            self._L.b200rt_scene_destroy(self._scene)
            self._scene = None
        self.host.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiScene:
    """One host thread, one flattened description, one device scene per GPU: ``MultiScene(xml, "hw3", devices=[0, 1])
    .render_image(0)`` is the reference's thread fan-out over rows (HW7/src/main.cpp:35-46) with GPUs for threads
    (b200rt_render_multi), and ``render_ppm`` its photon fan-out (PPM/src/main.cpp:74-98)."""

    def __init__(self, xml: str, semantics="hw3", devices=None, base_dir: str | None = None):
        self._L = _lib.load()
        if devices is None:
            devices = list(range(self._L.b200rt_device_count()))
        if not devices:
            raise _lib.B200RTError("MultiScene: no CUDA device — the B200 core has no CPU fallback")
        self.parts = [Scene(xml, semantics, device=devices[0], base_dir=base_dir)]
        self.spec, self.host, self.semantics = self.parts[0].spec, self.parts[0].host, self.parts[0].semantics
        for d in devices[1:]:   # replicas share the flattened host description
            s = Scene.__new__(Scene)
            s.spec, s.host, s.semantics, s.device, s._scene, s._L = self.spec, self.host, self.semantics, d, None, self._L
            s.upload()
            self.parts.append(s)
        self.devices = list(devices)
        self._handles = (C.c_void_p * len(self.parts))(*[p._scene for p in self.parts])
        self._ppm_camera = 0

    def image_size(self, camera_index=0):
        return self.parts[0].image_size(camera_index)

    def render_image(self, camera_index: int = 0, out: np.ndarray | None = None, **opts):
        w, h = self.image_size(camera_index)
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        st = _lib.RenderStats()
        o = self.parts[0]._opts(**opts)
        _lib.check(self._L.b200rt_render_multi(self._handles, len(self.parts), camera_index, C.byref(o), out.ctypes.data,
                                               C.byref(st)), "b200rt_render_multi")
        return out, st

    def eye_trace_lines(self, camera_index: int = 0, seed: int = 0):
        st = _lib.PpmStats()
        _lib.check(self._L.b200rt_ppm_eye_pass_multi(self._handles, len(self.parts), camera_index, seed, C.byref(st)),
                   "b200rt_ppm_eye_pass_multi")
        self._ppm_camera = camera_index
        return st

    def trace_n_photons(self, n: int, first_photon: int = 0, seed: int = 0):
        o = _lib.PpmOpts()
        o.photons = int(n); o.first_photon = int(first_photon); o.seed = int(seed)
        st = _lib.PpmStats()
        _lib.check(self._L.b200rt_ppm_trace_photons_multi(self._handles, len(self.parts), C.byref(o), C.byref(st)),
                   "b200rt_ppm_trace_photons_multi")
        return st

    def density_estimation(self, total_num_of_photons: float, out: np.ndarray | None = None):
        w, h = self.image_size(self._ppm_camera)
        if out is None:
            out = np.zeros((h, w, 4), np.float32)
        _lib.check(self._L.b200rt_ppm_density_multi(self._handles, len(self.parts), float(total_num_of_photons),
                                                    out.ctypes.data), "b200rt_ppm_density_multi")
        return out

    def render_ppm(self, camera_index: int = 0, photons: int = 0, passes: int = 1, seed: int = 0):
        """`passes` passes of `photons` photons each, counts synchronised after every pass."""
        s0 = self.eye_trace_lines(camera_index, seed)
        s1 = None
        for k in range(passes):
            s1 = self.trace_n_photons(photons, k * photons, seed)
        return self.density_estimation(photons * passes), s0, s1

    def close(self):
        for p in self.parts[1:]:
            if p._scene is not None:
                p._L.b200rt_scene_destroy(p._scene)
                p._scene = None
        self.parts[0].close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
