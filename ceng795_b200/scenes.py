"""Seeded synthetic scene generators for BASELINE.json's configs.

The reference repository ships no scenes (`.gitignore:34-38` excludes them), so
both arms — the untouched reference binary and the CUDA core — render XML
written here, in each tree's own dialect.  Floats are printed with `%.9g`, which
round-trips binary32 exactly through either parser.
"""
from __future__ import annotations

import io

import numpy as np

SEED = 795


def _f(a):
    return " ".join("%.9g" % float(x) for x in np.asarray(a, dtype=np.float32).ravel())


def _rows(a, fmt):
    buf = io.StringIO()
    np.savetxt(buf, a, fmt=fmt)
    return buf.getvalue()


def _uv_sphere(center, radius, slices, stacks):
    """Closed UV sphere: 2*slices*(stacks-1) triangles, CCW seen from outside."""
    c = np.asarray(center, np.float32)
    verts = [c + np.array([0, radius, 0], np.float32)]
    for i in range(1, stacks):
        th = np.pi * i / stacks
        for j in range(slices):
            ph = 2 * np.pi * j / slices
            verts.append(c + radius * np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], np.float32))
    verts.append(c - np.array([0, radius, 0], np.float32))
    verts = np.asarray(verts, np.float32)
    faces = []
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)
    for j in range(slices):
        faces.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c2, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append((a, b, c2))
            faces.append((b, d, c2))
    last = len(verts) - 1
    for j in range(slices):
        faces.append((last, ring(stacks - 1, j), ring(stacks - 1, j + 1)))
    return verts, np.asarray(faces, np.int32)


def _material_xml(mid, amb, dif, spec, phong, mirror=None, transparency=None, refr=None, roughness=None):
    s = f'    <Material id="{mid}">\n'
    s += f"      <AmbientReflectance>{_f(amb)}</AmbientReflectance>\n"
    s += f"      <DiffuseReflectance>{_f(dif)}</DiffuseReflectance>\n"
    s += f"      <SpecularReflectance>{_f(spec)}</SpecularReflectance>\n"
    if mirror is not None:
        s += f"      <MirrorReflectance>{_f(mirror)}</MirrorReflectance>\n"
    if roughness is not None:
        s += f"      <Roughness>{_f([roughness])}</Roughness>\n"
    s += f"      <PhongExponent>{_f([phong])}</PhongExponent>\n"
    if transparency is not None:
        s += f"      <Transparency>{_f(transparency)}</Transparency>\n"
        s += f"      <RefractionIndex>{_f([refr])}</RefractionIndex>\n"
    s += "    </Material>\n"
    return s


def cfg1_hw1(width=800, height=800, n_spheres=64, sphere_mesh=(20, 26), depth=6, seed=SEED,
             image_name="cfg1.png") -> str:
    """Config 1 — HW1 dialect: spheres + 2 triangle meshes + stand-alone triangles,
    3 point lights, Blinn-Phong + mirror, 1 spp."""
    rng = np.random.default_rng(seed)
    verts = []
    nv = 0

    def push(v):
        nonlocal nv
        base = nv
        verts.append(np.asarray(v, np.float32))
        nv += len(v)
        return base

    mats = ""
    for m in range(6):
        dif = rng.uniform(0.2, 0.9, 3)
        spec = rng.uniform(0.1, 0.6, 3)
        mirror = np.full(3, 0.5) if m % 2 == 1 else None
        mats += _material_xml(m + 1, rng.uniform(0.05, 0.2, 3), dif, spec, float(rng.choice([8, 20, 50, 100])), mirror)
    # room: 10 stand-alone triangles (floor, back, left, right, ceiling), facing inwards
    R = 6.0
    room = np.array([[-R, -R, -R], [R, -R, -R], [R, -R, R], [-R, -R, R],
                     [-R, R, -R], [R, R, -R], [R, R, R], [-R, R, R]], np.float32)
    rb = push(room)
    quads = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (4, 5, 6, 7)]
    tri_xml = ""
    tid = 1
    for q in quads:
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            idx = [rb + t + 1 for t in tri]
            tri_xml += (f'    <Triangle id="{tid}">\n      <Material>{1 + (tid % 6)}</Material>\n'
                        f"      <Indices>{idx[0]} {idx[1]} {idx[2]}</Indices>\n    </Triangle>\n")
            tid += 1
    mesh_xml = ""
    for k, c in enumerate([(-2.5, -3.0, -1.0), (2.5, -2.0, 0.5)]):
        v, f = _uv_sphere(c, 1.6 + 0.3 * k, sphere_mesh[0], sphere_mesh[1])
        b = push(v)
        mesh_xml += (f'    <Mesh id="{k + 1}">\n      <Material>{2 + 3 * k}</Material>\n      <Faces>\n'
                     + _rows(f + b + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    sph_xml = ""
    centers = rng.uniform(-4, 4, (n_spheres, 3)).astype(np.float32)
    radii = rng.uniform(0.2, 0.6, n_spheres).astype(np.float32)
    cb = push(centers)
    for i in range(n_spheres):
        sph_xml += (f'    <Sphere id="{i + 1}">\n      <Material>{1 + int(rng.integers(0, 6))}</Material>\n'
                    f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([radii[i]])}</Radius>\n    </Sphere>\n")
    lights = ""
    for i in range(3):
        pos = np.array([rng.uniform(-4, 4), rng.uniform(3, 5.5), rng.uniform(-1, 5)])
        inten = rng.uniform(5e2, 2e3, 3)
        lights += (f'    <PointLight id="{i + 1}">\n      <Position>{_f(pos)}</Position>\n'
                   f"      <Intensity>{_f(inten)}</Intensity>\n    </PointLight>\n")
    allv = np.concatenate(verts, 0)
    aspect = width / height
    return ("<Scene>\n  <BackgroundColor>10 20 40</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 0.5 5.9</Position>\n      <Gaze>0 -0.05 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-aspect, aspect, -1, 1])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <ImageName>{image_name}</ImageName>\n"
            "    </Camera>\n  </Cameras>\n  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n" + lights +
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n  <VertexData>\n" + _rows(allv, "%.9g") +
            "  </VertexData>\n  <Objects>\n" + mesh_xml + tri_xml + sph_xml + "  </Objects>\n</Scene>\n")


def heightfield(n, seed=SEED, extent=4.0):
    """(n+1)^2 vertices, 2*n*n triangles: y = .25 sin3x cos3z + .02 U(-1,1)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-extent, extent, n + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = 0.25 * np.sin(3 * X) * np.cos(3 * Z) + 0.02 * rng.uniform(-1, 1, X.shape)
    v = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a = (j * (n + 1) + i).ravel()
    b = a + 1
    c = a + (n + 1)
    d = c + 1
    # CCW seen from +y
    f = np.empty((2 * n * n, 3), np.int64)
    f[0::2] = np.stack([a, c, b], -1)
    f[1::2] = np.stack([b, c, d], -1)
    return v, f.astype(np.int32)


def cfg2_hw3(n=724, width=1920, height=1080, n_instances=8, depth=2, seed=SEED, image_name="cfg2.png",
             smooth=False, cam_angle_deg=0.0, glass=False) -> str:
    """Config 2 — HW2-era features in the HW3 dialect (the HW2 tree lacks MeshInstance):
    a (2*n*n)-triangle heightfield (n=724 -> 1 048 352), `n_instances` transformed
    instances of it, 4 spheres, 2 point lights, mirror depth 2, 1 spp."""
    rng = np.random.default_rng(seed + 2)
    v, f = heightfield(n, seed)
    sph_c = np.array([[-2.0, 1.0, 1.5], [2.2, 0.9, 0.5], [0.0, 1.4, -1.5], [1.0, 0.7, 2.5]], np.float32)
    sph_r = [0.6, 0.5, 0.8, 0.35]
    allv = np.concatenate([v, sph_c], 0)
    cb = len(v)
    mats = (_material_xml(1, [0.3, 0.3, 0.3], [0.45, 0.55, 0.35], [0.2, 0.2, 0.2], 20)
            + _material_xml(2, [0.1, 0.1, 0.1], [0.15, 0.15, 0.2], [0.5, 0.5, 0.5], 80, mirror=[0.6, 0.6, 0.6])
            + _material_xml(3, [0.1, 0.1, 0.1], [0.3, 0.1, 0.1], [0.4, 0.4, 0.4], 50, mirror=[0.35, 0.3, 0.3])
            + _material_xml(4, [0.3, 0.3, 0.3], [0.7, 0.3, 0.2], [0.3, 0.3, 0.3], 10)
            + _material_xml(5, [0.3, 0.3, 0.3], [0.2, 0.4, 0.8], [0.6, 0.6, 0.6], 40)
            + (_material_xml(6, [0, 0, 0], [0.02, 0.02, 0.02], [0.2, 0.2, 0.2], 60, transparency=[0.85, 0.9, 0.95], refr=1.45)
               if glass else ""))
    xf = "  <Transformations>\n"
    inst = ""
    for k in range(n_instances):
        s = rng.uniform(0.08, 0.18)
        xf += f'    <Scaling id="{k + 1}">{_f([s, s * rng.uniform(1.0, 3.0), s])}</Scaling>\n'
        ang = np.linspace(0, 2 * np.pi, n_instances, endpoint=False)[k]
        xf += (f'    <Translation id="{k + 1}">{_f([2.6 * np.cos(ang), rng.uniform(1.0, 2.4), 2.6 * np.sin(ang) - 0.3])}'
               "</Translation>\n")
        axis = rng.normal(size=3)
        xf += f'    <Rotation id="{k + 1}">{_f([rng.uniform(0, 360), *axis])}</Rotation>\n'
        inst += (f'    <MeshInstance id="{k + 1}" baseMeshId="1" resetTransform="true">\n'
                 f"      <Material>{2 + (k % 4)}</Material>\n"
                 f"      <Transformations>s{k + 1} r{k + 1} t{k + 1}</Transformations>\n    </MeshInstance>\n")
    xf += "  </Transformations>\n"
    sph = ""
    for i in range(4):
        sph += (f'    <Sphere id="{i + 1}">\n      <Material>{[2, 6 if glass else 4, 3, 5][i]}</Material>\n'
                f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([sph_r[i]])}</Radius>\n    </Sphere>\n")
    aspect = width / height
    shading = ' shadingMode="smooth"' if smooth else ""
    ca, sa = np.cos(np.radians(cam_angle_deg)), np.sin(np.radians(cam_angle_deg))
    cam_pos = [5.6 * sa, 3.1, 5.6 * ca]       # orbit about +y (angle 0 = the canonical camera)
    cam_gaze = [-sa, -0.62, -ca]
    return ("<Scene>\n  <BackgroundColor>30 40 60</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            f"      <Position>{_f(cam_pos)}</Position>\n      <Gaze>{_f(cam_gaze)}</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.45 * aspect, 0.45 * aspect, -0.45, 0.45])}</NearPlane>\n"
            "      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>1</NumSamples>\n"
            f"      <ImageName>{image_name}</ImageName>\n    </Camera>\n  </Cameras>\n"
            "  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n"
            '    <PointLight id="1">\n      <Position>-3 6 4</Position>\n      <Intensity>5000 4800 4500</Intensity>\n    </PointLight>\n'
            '    <PointLight id="2">\n      <Position>4 5 -2</Position>\n      <Intensity>2500 3000 3500</Intensity>\n    </PointLight>\n'
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n" + xf + "  <VertexData>\n"
            + _rows(allv, "%.9g") + "  </VertexData>\n  <Objects>\n"
            + f'    <Mesh id="1"{shading}>\n      <Material>1</Material>\n      <Faces>\n' + _rows(f + 1, "%d")
            + "      </Faces>\n    </Mesh>\n" + inst + sph + "  </Objects>\n</Scene>\n")


# --------------------------------------------------------------------------- HW7
def write_test_texture(path, w=64, h=48, seed=SEED):
    """A seeded RGB image (checker x gradient x noise) for the image-texture paths."""
    from PIL import Image
    rng = np.random.default_rng(seed + 7)
    yy, xx = np.mgrid[0:h, 0:w]
    chk = (((xx // 8) + (yy // 6)) % 2).astype(np.float32)
    img = np.stack([60 + 150 * chk + 40 * xx / w, 200 - 120 * chk + 30 * yy / h, 80 + 100 * (xx + yy) / (w + h)], -1)
    img = np.clip(img + rng.uniform(-12, 12, img.shape), 0, 255).astype(np.uint8)
    Image.fromarray(img, "RGB").save(path)
    return path


def _material7(mid, amb, dif, spec, phong, mirror=None, transparency=None, refr=None, roughness=None, brdf=None,
               degamma=False):
    attrs = f' id="{mid}"' + (f' BRDF="{brdf}"' if brdf else "") + (' degamma="true"' if degamma else "")
    s = f"    <Material{attrs}>\n"
    s += f"      <AmbientReflectance>{_f(amb)}</AmbientReflectance>\n"
    s += f"      <DiffuseReflectance>{_f(dif)}</DiffuseReflectance>\n"
    s += f"      <SpecularReflectance>{_f(spec)}</SpecularReflectance>\n"
    if mirror is not None:
        s += f"      <MirrorReflectance>{_f(mirror)}</MirrorReflectance>\n"
    if roughness is not None:
        s += f"      <Roughness>{_f([roughness])}</Roughness>\n"
    s += f"      <PhongExponent>{_f([phong])}</PhongExponent>\n"
    if transparency is not None:
        s += f"      <Transparency>{_f(transparency)}</Transparency>\n"
        s += f"      <RefractionIndex>{_f([refr])}</RefractionIndex>\n"
    s += "    </Material>\n"
    return s


def hw7_scene(asset_dir, width=160, height=120, n=12, spp=1, depth=3, path_tracing=False, importance=True,
              stochastic_lights=False, perlin=False, seed=SEED, image_name="hw7.png", light_scale=1.0,
              mesh_slices=(10, 12), tonemap=False) -> str:
    """HW7 dialect: Cornell-style room (stand-alone triangles) + an image-textured, smooth-shaded
    heightfield mesh with texture coordinates + instanced UV-sphere meshes (one material per BRDF
    kind, one mirror, one glass) + textured spheres; point / spot / directional lights.
    `stochastic_lights` adds an AreaLight, a LightMesh and a LightSphere (random sampling: RMSE parity);
    without it every pixel is deterministic at 1 spp (byte parity against the HW7 binary).
    `perlin` adds Perlin textures (the reference shuffles their table from the clock: GPU-vs-oracle only).
    Writes the texture image into `asset_dir`; the XML refers to it by bare file name."""
    import os
    rng = np.random.default_rng(seed + 4)
    if not os.path.exists(os.path.join(asset_dir, "tex_checker.png")):
        write_test_texture(os.path.join(asset_dir, "tex_checker.png"), seed=SEED)
    verts, tcs = [], []
    nv = 0

    def push(v, uv=None):
        nonlocal nv
        base = nv
        v = np.asarray(v, np.float32)
        verts.append(v)
        tcs.append(np.zeros((len(v), 2), np.float32) if uv is None else np.asarray(uv, np.float32))
        nv += len(v)
        return base

    R = 5.0
    room = np.array([[-R, -1, -R], [R, -1, -R], [R, -1, R], [-R, -1, R],
                     [-R, 2 * R - 1, -R], [R, 2 * R - 1, -R], [R, 2 * R - 1, R], [-R, 2 * R - 1, R]], np.float32)
    rb = push(room)
    quads = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (4, 5, 6, 7)]
    tri_xml = ""
    tid = 1
    for qi, q in enumerate(quads):
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            idx = [rb + t + 1 for t in tri]
            tri_xml += (f'    <Triangle id="{tid}">\n      <Material>{1 + (qi % 3)}</Material>\n'
                        f"      <Indices>{idx[0]} {idx[1]} {idx[2]}</Indices>\n    </Triangle>\n")
            tid += 1
    # textured heightfield patch with uv = position in the patch (repeats 2x)
    hv, hf = heightfield(n, seed, extent=2.2)
    hv = hv * np.array([1, 0.8, 1], np.float32) + np.array([0, -0.55, 0.2], np.float32)
    huv = np.stack([(hv[:, 0] + 2.2) / 2.2, (hv[:, 2] + 2.0) / 2.2], -1)
    hb = push(hv, huv)
    mesh_xml = (f'    <Mesh id="1" shadingMode="smooth">\n      <Material>4</Material>\n      <Texture>1</Texture>\n'
                f'      <Faces vertexOffset="{hb}" textureOffset="{hb}">\n' + _rows(hf + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    # base UV-sphere mesh, instanced with the BRDF materials
    sv, sf = _uv_sphere((0, 0, 0), 1.0, mesh_slices[0], mesh_slices[1])
    sb = push(sv)
    mesh_xml += (f'    <Mesh id="2" shadingMode="smooth">\n      <Material>5</Material>\n'
                 f'      <Transformations>s1 t1</Transformations>\n'
                 f'      <Faces vertexOffset="{sb}">\n' + _rows(sf + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    xf = "  <Transformations>\n    <Scaling id=\"1\">0.55 0.55 0.55</Scaling>\n    <Translation id=\"1\">-2.4 0.3 -1.2</Translation>\n"
    inst = ""
    for k in range(6):
        s = 0.38 + 0.05 * k
        xf += f'    <Scaling id="{k + 2}">{_f([s, s * (1.0 + 0.15 * k), s])}</Scaling>\n'
        xf += f'    <Translation id="{k + 2}">{_f([-2.0 + 0.85 * k, 0.25 + 0.35 * (k % 3), -0.2 + 0.6 * (k % 2) - 0.25 * k])}</Translation>\n'
        xf += f'    <Rotation id="{k + 1}">{_f([37.0 * k + 11, *rng.normal(size=3)])}</Rotation>\n'
        inst += (f'    <MeshInstance id="{k + 1}" baseMeshId="2" resetTransform="true">\n'
                 f"      <Material>{6 + k}</Material>\n"
                 f"      <Transformations>s{k + 2} r{k + 1} t{k + 2}</Transformations>\n    </MeshInstance>\n")
    xf += "  </Transformations>\n"
    # spheres: textured (nearest/clamp replace_kd), glass, mirror, blend_kd bilinear
    sph_c = np.array([[1.9, 0.0, 1.6], [-1.0, 0.1, 2.0], [0.4, 1.9, -2.2], [2.8, 1.3, -0.5]], np.float32)
    sph_r = [0.7, 0.6, 0.9, 0.55]
    cb = push(sph_c)
    sph_tex = [2, None, None, 3]
    sph_mat = [4, 12, 13, 4]
    sph = ""
    for i in range(4):
        tx = f"      <Texture>{sph_tex[i]}</Texture>\n" if sph_tex[i] else ""
        sph += (f'    <Sphere id="{i + 1}">\n      <Material>{sph_mat[i]}</Material>\n{tx}'
                f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([sph_r[i]])}</Radius>\n    </Sphere>\n")
    if perlin:
        pc = push(np.array([[-2.9, 1.6, 0.9]], np.float32))
        sph += (f'    <Sphere id="5">\n      <Material>4</Material>\n      <Texture>4</Texture>\n'
                f"      <Center>{pc + 1}</Center>\n      <Radius>0.6</Radius>\n    </Sphere>\n")
    L = light_scale
    mats = (_material7(1, [0.25, 0.25, 0.25], [0.6, 0.6, 0.55], [0.05, 0.05, 0.05], 5)
            + _material7(2, [0.25, 0.2, 0.2], [0.65, 0.2, 0.2], [0.05, 0.05, 0.05], 5)
            + _material7(3, [0.2, 0.25, 0.2], [0.2, 0.6, 0.25], [0.05, 0.05, 0.05], 5)
            + _material7(4, [0.2, 0.2, 0.2], [0.5, 0.5, 0.5], [0.3, 0.3, 0.3], 30)
            + _material7(5, [0.2, 0.2, 0.25], [0.3, 0.3, 0.7], [0.4, 0.4, 0.4], 60, degamma=True)
            + "".join(_material7(6 + k, [0.15, 0.15, 0.15], rng.uniform(0.2, 0.7, 3), rng.uniform(0.2, 0.5, 3),
                                 float(10 + 15 * k), brdf=(k + 1) if k < 5 else None) for k in range(6))
            + _material7(12, [0, 0, 0], [0.02, 0.02, 0.02], [0.1, 0.1, 0.1], 80, transparency=[0.9, 0.95, 0.9], refr=1.5)
            + _material7(13, [0.05, 0.05, 0.05], [0.05, 0.05, 0.05], [0.3, 0.3, 0.3], 90, mirror=[0.8, 0.8, 0.8],
                         roughness=(0.12 if stochastic_lights else None)))
    brdfs = ('  <BRDFs>\n    <OriginalPhong id="1"><Exponent>20</Exponent></OriginalPhong>\n'
             '    <ModifiedPhong id="2" normalized="true"><Exponent>30</Exponent></ModifiedPhong>\n'
             '    <OriginalBlinnPhong id="3"><Exponent>40</Exponent></OriginalBlinnPhong>\n'
             '    <ModifiedBlinnPhong id="4" normalized="true"><Exponent>50</Exponent></ModifiedBlinnPhong>\n'
             '    <TorranceSparrow id="5"><Exponent>60</Exponent><RefractiveIndex>2.1</RefractiveIndex></TorranceSparrow>\n'
             "  </BRDFs>\n")
    textures = ('  <Textures>\n'
                '    <Texture id="1"><ImageName>tex_checker.png</ImageName><Interpolation>bilinear</Interpolation>'
                '<DecalMode>replace_kd</DecalMode><Appearance>repeat</Appearance></Texture>\n'
                '    <Texture id="2" degamma="true"><ImageName>tex_checker.png</ImageName><Interpolation>nearest</Interpolation>'
                '<DecalMode>replace_kd</DecalMode><Appearance>clamp</Appearance></Texture>\n'
                '    <Texture id="3" bumpmap="true" bumpmapMultiplier="0.02"><ImageName>tex_checker.png</ImageName>'
                '<Interpolation>bilinear</Interpolation><DecalMode>blend_kd</DecalMode><Appearance>clamp</Appearance>'
                '<Normalizer>255</Normalizer></Texture>\n'
                + ('    <Texture id="4" bumpmap="true" bumpmapMultiplier="0.6"><ImageName>perlin</ImageName>'
                   '<DecalMode>blend_kd</DecalMode><Appearance>patch</Appearance><ScalingFactor>3.5</ScalingFactor>'
                   '<Permutation>11 3 7 0 14 9 2 13 5 8 1 15 6 12 4 10</Permutation></Texture>\n' if perlin else "")
                + "  </Textures>\n")
    lights = ("  <Lights>\n    <AmbientLight>20 20 20</AmbientLight>\n"
              f'    <PointLight id="1"><Position>-2 7.5 3</Position><Intensity>{_f(np.array([2600, 2500, 2300]) * L)}</Intensity></PointLight>\n'
              + (f'    <AreaLight id="1"><Position>1 8.2 -1</Position><Intensity>{_f(np.array([1800, 1800, 1700]) * L)}</Intensity>'
                 '<EdgeVector1>1.2 0 0</EdgeVector1><EdgeVector2>0 0 1.2</EdgeVector2></AreaLight>\n' if stochastic_lights else "")
              + f'    <SpotLight id="1"><Position>3.5 6 3.5</Position><Direction>-0.5 -1 -0.45</Direction>'
              f'<Intensity>{_f(np.array([3000, 3200, 3400]) * L)}</Intensity><CoverageAngle>50</CoverageAngle><FalloffAngle>30</FalloffAngle></SpotLight>\n'
              f'    <DirectionalLight id="1"><Direction>0.3 -1 -0.6</Direction><Radiance>{_f(np.array([30, 28, 25]) * L)}</Radiance></DirectionalLight>\n'
              "  </Lights>\n")
    lobj = ""
    if stochastic_lights:
        lq = push(np.array([[-0.8, 8.7, -0.8], [0.8, 8.7, -0.8], [0.8, 8.7, 0.8], [-0.8, 8.7, 0.8]], np.float32))
        lobj += (f'    <LightMesh id="1">\n      <Material>1</Material>\n      <Radiance>{_f(np.array([900, 900, 850]) * L)}</Radiance>\n'
                 f'      <Faces vertexOffset="{lq}">\n1 2 3\n1 3 4\n      </Faces>\n    </LightMesh>\n')
        lc = push(np.array([[-3.2, 5.0, -2.5]], np.float32))
        lobj += (f'    <LightSphere id="1">\n      <Material>1</Material>\n      <Center>{lc + 1}</Center>\n      <Radius>0.45</Radius>\n'
                 f"      <Radiance>{_f(np.array([1200, 1000, 800]) * L)}</Radiance>\n    </LightSphere>\n")
    allv = np.concatenate(verts, 0)
    alluv = np.concatenate(tcs, 0)
    aspect = width / height
    integ = ""
    if path_tracing:
        integ = "  <Integrator>PathTracing</Integrator>\n" + ("  <IntegratorParams>ImportanceSampling</IntegratorParams>\n" if importance else "")
    tm = ("      <Tonemap><TMO>Photographic</TMO><TMOOptions>0.18 1</TMOOptions><Saturation>1.0</Saturation></Tonemap>\n"
          if tonemap else "")
    return ("<Scene>\n  <BackgroundColor>12 16 30</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n" + integ + "  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 2.6 4.8</Position>\n      <Gaze>0 -0.32 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.6 * aspect, 0.6 * aspect, -0.6, 0.6])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>{spp}</NumSamples>\n"
            f"      <ImageName>{image_name}</ImageName>\n" + tm + "    </Camera>\n  </Cameras>\n"
            + brdfs + lights + "  <Materials>\n" + mats + "  </Materials>\n" + xf + "  <VertexData>\n"
            + _rows(allv, "%.9g") + "  </VertexData>\n  <TexCoordData>\n" + _rows(alluv, "%.9g") + "  </TexCoordData>\n"
            + textures + "  <Objects>\n" + mesh_xml + inst + lobj + tri_xml + sph + "  </Objects>\n</Scene>\n")


def cfg4_hw7(asset_dir, width=1920, height=1080, spp=4096, depth=4, mesh_slices=(224, 225), tex_size=1024,
             seed=SEED, image_name="cfg4.png", radiance=15.0) -> str:
    """BASELINE config 4 as SURVEY.md 8(d) words it: Cornell-style box (5 walls, stand-alone triangles) + a LightMesh
    (2 triangles, radiance 15) + 1 LightSphere + 3 instances of a ~100 k-triangle smooth mesh carrying an image texture
    (tex_size^2, bilinear / repeat, replace_kd), a Perlin `patch` bump texture, one material per BRDF type (5), one
    mirror, one glass; PathTracing + ImportanceSampling, MaxRecursionDepth 4, Photographic TMO key .18 / 1 % burn,
    NumSamples `spp` (64 x 64 strata at 4096).  No point / area lights: the two light shapes are the only emitters.
    Writes the texture image into `asset_dir`; the XML refers to it by bare file name."""
    import os
    rng = np.random.default_rng(seed + 44)
    tex_name = f"tex_cfg4_{tex_size}.png"
    if not os.path.exists(os.path.join(asset_dir, tex_name)):
        write_test_texture(os.path.join(asset_dir, tex_name), w=tex_size, h=tex_size, seed=SEED)
    verts, tcs = [], []
    nv = 0

    def push(v, uv=None):
        nonlocal nv
        base = nv
        v = np.asarray(v, np.float32)
        verts.append(v)
        tcs.append(np.zeros((len(v), 2), np.float32) if uv is None else np.asarray(uv, np.float32))
        nv += len(v)
        return base

    R = 5.0
    top = 2 * R - 1
    room = np.array([[-R, -1, -R], [R, -1, -R], [R, -1, R], [-R, -1, R],
                     [-R, top, -R], [R, top, -R], [R, top, R], [-R, top, R]], np.float32)
    rb = push(room)
    # floor, back wall, left wall (red), right wall (green), ceiling — wound to face the inside (HW7 culls back faces)
    quads = [((0, 3, 2, 1), 1), ((0, 1, 5, 4), 1), ((0, 4, 7, 3), 2), ((1, 2, 6, 5), 3), ((4, 5, 6, 7), 1)]
    tri_xml = ""
    tid = 1
    for q, mat in quads:
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            idx = [rb + t + 1 for t in tri]
            tri_xml += (f'    <Triangle id="{tid}">\n      <Material>{mat}</Material>\n'
                        f"      <Indices>{idx[0]} {idx[1]} {idx[2]}</Indices>\n    </Triangle>\n")
            tid += 1
    # the ~100 k-triangle base mesh: a UV sphere with latitude / longitude texture coordinates (4 x 2 repeats)
    sv, sf = _uv_sphere((0, 0, 0), 1.0, mesh_slices[0], mesh_slices[1])
    suv = np.stack([4.0 * (np.arctan2(sv[:, 2], sv[:, 0]) + np.pi) / (2 * np.pi),
                    2.0 * np.arccos(np.clip(sv[:, 1], -1, 1)) / np.pi], -1)
    sb = push(sv, suv)
    mesh_xml = (f'    <Mesh id="1" shadingMode="smooth">\n      <Material>4</Material>\n      <Texture>1</Texture>\n'
                f'      <Transformations>s1 t1</Transformations>\n'
                f'      <Faces vertexOffset="{sb}" textureOffset="{sb}">\n' + _rows(sf + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    xf = ('  <Transformations>\n    <Scaling id="1">1.15 1.15 1.15</Scaling>\n    <Translation id="1">-2.6 0.15 -2.0</Translation>\n'
          '    <Scaling id="2">0.9 1.25 0.9</Scaling>\n    <Translation id="2">0.1 0.25 -2.6</Translation>\n'
          f'    <Rotation id="1">{_f([31.0, *rng.normal(size=3)])}</Rotation>\n'
          '    <Scaling id="3">1.0 0.8 1.0</Scaling>\n    <Translation id="3">2.7 -0.2 -1.6</Translation>\n'
          f'    <Rotation id="2">{_f([74.0, *rng.normal(size=3)])}</Rotation>\n  </Transformations>\n')
    inst = ('    <MeshInstance id="1" baseMeshId="1" resetTransform="true">\n      <Material>5</Material>\n'
            '      <Transformations>s2 r1 t2</Transformations>\n    </MeshInstance>\n'
            '    <MeshInstance id="2" baseMeshId="1" resetTransform="true">\n      <Material>6</Material>\n'
            '      <Transformations>s3 r2 t3</Transformations>\n    </MeshInstance>\n')
    # spheres: Perlin-bumped (BRDF 4), Torrance-Sparrow (BRDF 5), mirror, glass
    sph_c = np.array([[-1.5, -0.35, 1.0], [1.6, -0.3, 1.3], [-0.2, -0.1, 0.2], [0.9, -0.4, 2.4]], np.float32)
    sph_r = [0.65, 0.7, 0.9, 0.6]
    sph_mat = [7, 8, 9, 10]
    cb = push(sph_c)
    sph = ""
    for i in range(4):
        tx = "      <Texture>2</Texture>\n" if i == 0 else ""
        sph += (f'    <Sphere id="{i + 1}">\n      <Material>{sph_mat[i]}</Material>\n{tx}'
                f"      <Center>{cb + i + 1}</Center>\n      <Radius>{_f([sph_r[i]])}</Radius>\n    </Sphere>\n")
    mats = (_material7(1, [0, 0, 0], [0.73, 0.73, 0.7], [0, 0, 0], 1)
            + _material7(2, [0, 0, 0], [0.65, 0.08, 0.06], [0, 0, 0], 1)
            + _material7(3, [0, 0, 0], [0.12, 0.48, 0.1], [0, 0, 0], 1)
            + "".join(_material7(4 + k, [0, 0, 0], rng.uniform(0.25, 0.7, 3), rng.uniform(0.15, 0.4, 3),
                                 float(20 + 10 * k), brdf=k + 1) for k in range(5))
            + _material7(9, [0, 0, 0], [0.03, 0.03, 0.03], [0, 0, 0], 1, mirror=[0.9, 0.9, 0.9])
            + _material7(10, [0, 0, 0], [0.02, 0.02, 0.02], [0, 0, 0], 1, transparency=[0.95, 0.97, 0.95], refr=1.5)
            + _material7(11, [0, 0, 0], [0.5, 0.5, 0.5], [0, 0, 0], 1))
    brdfs = ('  <BRDFs>\n    <OriginalPhong id="1"><Exponent>20</Exponent></OriginalPhong>\n'
             '    <ModifiedPhong id="2" normalized="true"><Exponent>30</Exponent></ModifiedPhong>\n'
             '    <OriginalBlinnPhong id="3"><Exponent>40</Exponent></OriginalBlinnPhong>\n'
             '    <ModifiedBlinnPhong id="4" normalized="true"><Exponent>50</Exponent></ModifiedBlinnPhong>\n'
             '    <TorranceSparrow id="5"><Exponent>60</Exponent><RefractiveIndex>2.1</RefractiveIndex></TorranceSparrow>\n'
             "  </BRDFs>\n")
    textures = ('  <Textures>\n'
                f'    <Texture id="1"><ImageName>{tex_name}</ImageName><Interpolation>bilinear</Interpolation>'
                '<DecalMode>replace_kd</DecalMode><Appearance>repeat</Appearance></Texture>\n'
                '    <Texture id="2" bumpmap="true" bumpmapMultiplier="0.6"><ImageName>perlin</ImageName>'
                '<DecalMode>blend_kd</DecalMode><Appearance>patch</Appearance><ScalingFactor>3.5</ScalingFactor>'
                '<Permutation>11 3 7 0 14 9 2 13 5 8 1 15 6 12 4 10</Permutation></Texture>\n'
                "  </Textures>\n")
    lights = "  <Lights>\n    <AmbientLight>0 0 0</AmbientLight>\n  </Lights>\n"
    lq = push(np.array([[-1.2, top - 0.02, -1.8], [1.2, top - 0.02, -1.8], [1.2, top - 0.02, 0.6], [-1.2, top - 0.02, 0.6]], np.float32))
    lobj = (f'    <LightMesh id="1">\n      <Material>11</Material>\n      <Radiance>{_f([radiance] * 3)}</Radiance>\n'
            f'      <Faces vertexOffset="{lq}">\n1 2 3\n1 3 4\n      </Faces>\n    </LightMesh>\n')
    lc = push(np.array([[-3.4, 4.6, -3.0]], np.float32))
    lobj += (f'    <LightSphere id="1">\n      <Material>11</Material>\n      <Center>{lc + 1}</Center>\n      <Radius>0.5</Radius>\n'
             f"      <Radiance>{_f([radiance, radiance * 0.85, radiance * 0.7])}</Radiance>\n    </LightSphere>\n")
    allv = np.concatenate(verts, 0)
    alluv = np.concatenate(tcs, 0)
    aspect = width / height
    return ("<Scene>\n  <BackgroundColor>0 0 0</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Integrator>PathTracing</Integrator>\n"
            "  <IntegratorParams>ImportanceSampling</IntegratorParams>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 3.0 4.9</Position>\n      <Gaze>0 -0.3 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.6 * aspect, 0.6 * aspect, -0.6, 0.6])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>{spp}</NumSamples>\n"
            f"      <ImageName>{image_name}</ImageName>\n"
            "      <Tonemap><TMO>Photographic</TMO><TMOOptions>0.18 1</TMOOptions><Saturation>1.0</Saturation></Tonemap>\n"
            "    </Camera>\n  </Cameras>\n"
            + brdfs + lights + "  <Materials>\n" + mats + "  </Materials>\n" + xf + "  <VertexData>\n"
            + _rows(allv, "%.9g") + "  </VertexData>\n  <TexCoordData>\n" + _rows(alluv, "%.9g") + "  </TexCoordData>\n"
            + textures + "  <Objects>\n" + mesh_xml + inst + lobj + tri_xml + sph + "  </Objects>\n</Scene>\n")


# --------------------------------------------------------------------------- PPM
def cfg5_ppm(width=1920, height=1080, mesh_slices=(160, 158), photons_per_iteration=10_000_000, iterations=100,
             depth=8, spp=1, seed=SEED, image_name="cfg5.png") -> str:
    """Config 5 — PPM dialect: a closed diffuse box (a 12-triangle mesh; PPM has no stand-alone Triangle),
    one mirror sphere, one glass sphere (n = 1.5, T = .95) and one diffuse UV-sphere mesh of
    2*slices*(stacks-1) triangles (160 x 158 -> 50 240), one point light under the ceiling."""
    verts = []
    nv = 0

    def push(v):
        nonlocal nv
        base = nv
        verts.append(np.asarray(v, np.float32))
        nv += len(v)
        return base

    R = 4.0
    room = np.array([[-R, 0, -R], [R, 0, -R], [R, 0, R], [-R, 0, R],
                     [-R, 2 * R, -R], [R, 2 * R, -R], [R, 2 * R, R], [-R, 2 * R, R]], np.float32)
    rb = push(room)
    # faces wound so that their normals point INTO the room (culling is on in PPM)
    quads = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (4, 5, 6, 7), (3, 7, 6, 2)]
    groups = {1: [quads[0], quads[1], quads[4], quads[5]], 2: [quads[2]], 3: [quads[3]]}
    mesh_xml = ""
    mid = 1
    for mat, qs in groups.items():
        faces = []
        for q in qs:
            faces += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
        mesh_xml += (f'    <Mesh id="{mid}">\n      <Material>{mat}</Material>\n      <Faces vertexOffset="{rb}">\n'
                     + _rows(np.asarray(faces) + 1, "%d") + "      </Faces>\n    </Mesh>\n")
        mid += 1
    sv, sf = _uv_sphere((0, 0, 0), 1.0, mesh_slices[0], mesh_slices[1])
    sb = push(sv)
    mesh_xml += (f'    <Mesh id="{mid}" shadingMode="smooth">\n      <Material>4</Material>\n'
                 f'      <Transformations>s1 t1</Transformations>\n      <Faces vertexOffset="{sb}">\n'
                 + _rows(sf + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    cb = push(np.array([[-1.7, 1.0, -0.6], [1.6, 0.9, 1.2]], np.float32))
    sph = (f'    <Sphere id="1">\n      <Material>5</Material>\n      <Center>{cb + 1}</Center>\n      <Radius>1.0</Radius>\n    </Sphere>\n'
           f'    <Sphere id="2">\n      <Material>6</Material>\n      <Center>{cb + 2}</Center>\n      <Radius>0.9</Radius>\n    </Sphere>\n')
    mats = ('    <Material id="1"><DiffuseReflectance>0.75 0.75 0.75</DiffuseReflectance></Material>\n'
            '    <Material id="2"><DiffuseReflectance>0.75 0.25 0.25</DiffuseReflectance></Material>\n'
            '    <Material id="3"><DiffuseReflectance>0.25 0.25 0.75</DiffuseReflectance></Material>\n'
            '    <Material id="4"><DiffuseReflectance>0.6 0.7 0.3</DiffuseReflectance>'
            '<SpecularReflectance>0.2 0.2 0.2</SpecularReflectance><PhongExponent>20</PhongExponent></Material>\n'
            '    <Material id="5"><MirrorReflectance>0.95 0.95 0.95</MirrorReflectance></Material>\n'
            '    <Material id="6"><Transparency>0.95 0.95 0.95</Transparency><RefractionIndex>1.5</RefractionIndex></Material>\n')
    allv = np.concatenate(verts, 0)
    aspect = width / height
    return ("<Scene>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <PhotonCountPerIteration>{photons_per_iteration}</PhotonCountPerIteration>\n"
            f"  <NumberOfIterations>{iterations}</NumberOfIterations>\n  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n"
            "  <Cameras>\n    <Camera id=\"1\">\n      <Position>0 3.6 3.9</Position>\n      <Gaze>0 -0.25 -1</Gaze>\n"
            f"      <Up>0 1 0</Up>\n      <NearPlane>{_f([-0.62 * aspect, 0.62 * aspect, -0.62, 0.62])}</NearPlane>\n"
            f"      <NearDistance>1</NearDistance>\n      <ImageResolution>{width} {height}</ImageResolution>\n"
            f"      <NumSamples>{spp}</NumSamples>\n      <ImageName>{image_name}</ImageName>\n    </Camera>\n  </Cameras>\n"
            "  <Lights>\n    <PointLight id=\"1\"><Position>0 7.2 0</Position><Intensity>5 5 5</Intensity></PointLight>\n  </Lights>\n"
            "  <Materials>\n" + mats + "  </Materials>\n"
            "  <Transformations>\n    <Scaling id=\"1\">0.8 0.8 0.8</Scaling>\n    <Translation id=\"1\">0.1 0.8 -1.9</Translation>\n  </Transformations>\n"
            "  <VertexData>\n" + _rows(allv, "%.9g") + "  </VertexData>\n  <Objects>\n" + sph + mesh_xml + "  </Objects>\n</Scene>\n")


# --------------------------------------------------------------------------- cfg 3
def cfg3_hw3(n=362, width=1920, height=1080, spp=100, depth=3, seed=SEED, image_name="cfg3.png") -> str:
    """Config 3 — HW3 distribution ray tracing: the cfg-2 heightfield at 2*n*n triangles (n=362 -> 262 088),
    two AreaLights, a glossy-mirror floor slab (Roughness .15), one glass sphere, one motion-blurred sphere,
    depth of field (ApertureSize .1, focus at the scene centre), NumSamples strata."""
    v, f = heightfield(n, seed)
    floor = np.array([[-6, -0.6, -6], [6, -0.6, -6], [6, -0.6, 6], [-6, -0.6, 6]], np.float32)
    sph_c = np.array([[-1.6, 1.1, 1.4], [1.9, 0.9, 0.8], [0.2, 1.5, -1.3]], np.float32)
    allv = np.concatenate([v, floor, sph_c], 0)
    fb, cb = len(v), len(v) + 4
    mats = (_material_xml(1, [0.3, 0.3, 0.3], [0.45, 0.55, 0.35], [0.2, 0.2, 0.2], 20)
            + _material_xml(2, [0.1, 0.1, 0.1], [0.2, 0.2, 0.25], [0.4, 0.4, 0.4], 60, mirror=[0.55, 0.55, 0.55], roughness=0.15)
            + _material_xml(3, [0, 0, 0], [0.02, 0.02, 0.02], [0.2, 0.2, 0.2], 80, transparency=[0.9, 0.95, 0.9], refr=1.5)
            + _material_xml(4, [0.3, 0.3, 0.3], [0.7, 0.3, 0.2], [0.3, 0.3, 0.3], 10)
            + _material_xml(5, [0.3, 0.3, 0.3], [0.2, 0.4, 0.8], [0.6, 0.6, 0.6], 40))
    tri = (f'    <Triangle id="1"><Material>2</Material><Indices>{fb + 1} {fb + 4} {fb + 3}</Indices></Triangle>\n'
           f'    <Triangle id="2"><Material>2</Material><Indices>{fb + 1} {fb + 3} {fb + 2}</Indices></Triangle>\n')
    sph = (f'    <Sphere id="1"><Material>3</Material><Center>{cb + 1}</Center><Radius>0.7</Radius></Sphere>\n'
           f'    <Sphere id="2"><Material>4</Material><Center>{cb + 2}</Center><Radius>0.5</Radius>'
           '<MotionBlur>0.6 0.1 0</MotionBlur></Sphere>\n'
           f'    <Sphere id="3"><Material>5</Material><Center>{cb + 3}</Center><Radius>0.8</Radius></Sphere>\n')
    aspect = width / height
    return ("<Scene>\n  <BackgroundColor>30 40 60</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 3.1 5.6</Position>\n      <Gaze>0 -0.62 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.45 * aspect, 0.45 * aspect, -0.45, 0.45])}</NearPlane>\n"
            "      <NearDistance>1</NearDistance>\n      <FocusDistance>6.2</FocusDistance>\n      <ApertureSize>0.1</ApertureSize>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>{spp}</NumSamples>\n"
            f"      <ImageName>{image_name}</ImageName>\n    </Camera>\n  </Cameras>\n"
            "  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n"
            '    <AreaLight id="1"><Position>-3 6 3</Position><Intensity>5000 4800 4500</Intensity>'
            '<EdgeVector1>1.5 0 0</EdgeVector1><EdgeVector2>0 0 1.5</EdgeVector2></AreaLight>\n'
            '    <AreaLight id="2"><Position>3 5 -2</Position><Intensity>2500 3000 3500</Intensity>'
            '<EdgeVector1>1 0 0</EdgeVector1><EdgeVector2>0 0 1</EdgeVector2></AreaLight>\n'
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n  <VertexData>\n"
            + _rows(allv, "%.9g") + "  </VertexData>\n  <Objects>\n"
            + '    <Mesh id="1" shadingMode="smooth">\n      <Material>1</Material>\n      <Faces>\n' + _rows(f + 1, "%d")
            + "      </Faces>\n    </Mesh>\n" + tri + sph + "  </Objects>\n</Scene>\n")


def hw1_near_ties(width=96, height=72, layers=5, spacing=6e-7, seed=SEED, image_name="ties.png") -> str:
    """HW1 dialect, built to make the ORDER of primitives matter: HW1 keeps a hit iff `t < best - 1e-6`, so stacks of
    coplanar quads `spacing` apart (closer than 1e-6 in t), exact duplicates, quads listed back-to-front and
    front-to-back, and spheres tangent to the stacks give rays whose answer depends on the order of the list
    (rt_hw1_fast.cuh replays that order on the near-tie set; this scene is its test).  Every quad is a Mesh of its
    own: HW1 walks meshes first, in file order, and gives every stand-alone Triangle the LAST mesh's material
    (HW1/Scene.cpp:320; without any mesh it reads an uninitialised int, so there is always one here)."""
    rng = np.random.default_rng(seed + 21)
    verts = []
    nv = 0

    def push(v):
        nonlocal nv
        base = nv
        verts.append(np.asarray(v, np.float32))
        nv += len(v)
        return base

    mats = "".join(_material_xml(m + 1, rng.uniform(0.05, 0.2, 3), rng.uniform(0.2, 0.9, 3), rng.uniform(0.1, 0.6, 3),
                                 20.0, np.full(3, 0.4) if m == 2 else None) for m in range(6))
    mesh_xml = ""
    mid = 1

    def quad(z, x0, x1, y0, y1, mat):
        nonlocal mesh_xml, mid
        b = push([[x0, y0, z], [x1, y0, z], [x1, y1, z], [x0, y1, z]])
        mesh_xml += (f'    <Mesh id="{mid}">\n      <Material>{mat}</Material>\n      <Faces>\n'
                     f"{b + 1} {b + 2} {b + 3}\n{b + 1} {b + 3} {b + 4}\n      </Faces>\n    </Mesh>\n")
        mid += 1

    # stack A: back-to-front in the list (the farthest layer first), stack B: front-to-back, stack C: shuffled + duplicates
    for k in range(layers):
        quad(-2.0 - spacing * (layers - 1 - k), -2.6, -1.0, -1.0, 1.0, 1 + k % 6)
    for k in range(layers):
        quad(-2.0 - spacing * k, -0.8, 0.8, -1.0, 1.0, 1 + (k + 2) % 6)
    order = rng.permutation(layers)
    for k in order:
        quad(-2.0 - spacing * float(k), 1.0, 2.6, -1.0, 1.0, 1 + int(k) % 6)
        if k % 2 == 0:
            quad(-2.0 - spacing * float(k), 1.0, 2.6, -1.0, 1.0, 1 + (int(k) + 3) % 6)   # exact duplicate
    # far wall: two stand-alone triangles (they take the last mesh's material)
    wb = push([[-5, -4, -6.0], [5, -4, -6.0], [5, 4, -6.0], [-5, 4, -6.0]])
    tri_xml = "".join(f'    <Triangle id="{i + 1}">\n      <Material>4</Material>\n'
                      f"      <Indices>{wb + t[0] + 1} {wb + t[1] + 1} {wb + t[2] + 1}</Indices>\n    </Triangle>\n"
                      for i, t in enumerate(((0, 1, 2), (0, 2, 3))))
    # spheres tangent to the stacks' front plane (t of the sphere within 1e-6 of the plane's around the tangent point)
    cs = np.array([[-1.8, 0.4, -2.0 + 0.5], [0.0, -0.3, -2.0 + 0.35], [1.8, 0.2, -2.0 + 0.25 - 3e-7]], np.float32)
    rs = [0.5, 0.35, 0.25]
    cb = push(cs)
    sph_xml = "".join(f'    <Sphere id="{i + 1}">\n      <Material>{1 + i}</Material>\n      <Center>{cb + i + 1}</Center>\n'
                      f"      <Radius>{_f([rs[i]])}</Radius>\n    </Sphere>\n" for i in range(3))
    allv = np.concatenate(verts, 0)
    aspect = width / height
    return ("<Scene>\n  <BackgroundColor>10 20 40</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            "  <MaxRecursionDepth>2</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 0 3</Position>\n      <Gaze>0 0 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.75 * aspect, 0.75 * aspect, -0.75, 0.75])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <ImageName>{image_name}</ImageName>\n"
            "    </Camera>\n  </Cameras>\n  <Lights>\n    <AmbientLight>25 25 25</AmbientLight>\n"
            '    <PointLight id="1">\n      <Position>1 3 4</Position>\n      <Intensity>900 900 900</Intensity>\n    </PointLight>\n'
            "  </Lights>\n  <Materials>\n" + mats + "  </Materials>\n  <VertexData>\n" + _rows(allv, "%.9g")
            + "  </VertexData>\n  <Objects>\n" + mesh_xml + tri_xml + sph_xml + "  </Objects>\n</Scene>\n")


def fuzz_scene(seed, width=61, height=47, depth=3, spp=1, stochastic=False, camera_sphere_radius=2.5) -> str:
    """HW3-dialect scene for differential testing against the reference binaries (HW3 / HW4 / HW5 / HW6 / HW7 all load
    it): odd resolution and an axis-aligned camera (the centre ray is exactly (0,0,-1): the box test's skipped axes),
    meshes under random Scaling / Rotation / Translation chains including NEGATIVE (mirroring) scales, instances with
    and without resetTransform, transformed stand-alone triangles and spheres (non-uniform scale), zero-area and
    edge-on triangles, a sphere around the camera, smooth and flat shading, mirror and glass, two point lights.
    `stochastic` adds an area light, a glossy mirror, depth of field and a motion-blurred instance and sphere (for the
    draw-for-draw replay tests; use with spp > 1).

    `camera_sphere_radius`: the sphere around the camera.  At the default 2.5 it is a small room: most primary rays end
    on its inside (the `t2 < 0 -> t1` branch everywhere) and only what happens to lie within 2.5 of the camera shows.
    At 40 it is a backdrop behind everything, and the meshes, instances, triangles and spheres fill the frame — the
    variant the `open` tests use."""
    rng = np.random.default_rng(1000 + seed)
    verts = []
    nv = 0

    def push(v):
        nonlocal nv
        base = nv
        verts.append(np.asarray(v, np.float32))
        nv += len(v)
        return base

    mats = (_material_xml(1, [0.2, 0.2, 0.2], rng.uniform(0.2, 0.8, 3), rng.uniform(0.1, 0.5, 3), 12.0)
            + _material_xml(2, [0.1, 0.15, 0.1], rng.uniform(0.2, 0.8, 3), rng.uniform(0.1, 0.5, 3), 40.0,
                            mirror=[0.5, 0.5, 0.5], roughness=(0.2 if stochastic else None))
            + _material_xml(3, [0.05, 0.05, 0.05], [0.05, 0.05, 0.05], [0.2, 0.2, 0.2], 60.0,
                            transparency=[0.9, 0.85, 0.95], refr=1.45)
            + _material_xml(4, [0.25, 0.2, 0.15], rng.uniform(0.2, 0.8, 3), [0, 0, 0], 1.0)
            + _material_xml(5, [0.15, 0.2, 0.25], [0, 0, 0], rng.uniform(0.1, 0.5, 3), 25.0))
    n_xf = 6
    xf = "  <Transformations>\n"
    for k in range(n_xf):
        s = rng.uniform(0.4, 1.4, 3) * rng.choice([1, 1, -1], 3)      # mirroring scales
        xf += f'    <Scaling id="{k + 1}">{_f(s)}</Scaling>\n'
        xf += f'    <Translation id="{k + 1}">{_f(rng.uniform(-2.0, 2.0, 3) * [1, 0.6, 1] + [0, 0, -4])}</Translation>\n'
        xf += f'    <Rotation id="{k + 1}">{_f([rng.uniform(0, 360), *rng.normal(size=3)])}</Rotation>\n'
    xf += "  </Transformations>\n"

    def chain():
        ops = []
        for _ in range(int(rng.integers(1, 4))):
            ops.append(f"{rng.choice(['s', 'r', 't'])}{int(rng.integers(1, n_xf + 1))}")
        return " ".join(ops + [f"t{int(rng.integers(1, n_xf + 1))}"])

    mesh_xml = ""
    for m in range(3):
        v, f = _uv_sphere((0, 0, 0), float(rng.uniform(0.5, 0.9)), 7 + m, 8 + m)
        if m == 1:                                   # a zero-area triangle and an edge-on sliver in the mesh
            f = np.concatenate([f, [[0, 0, 1], [2, 3, 3]]], 0)
        b = push(v)
        mesh_xml += (f'    <Mesh id="{m + 1}"' + (' shadingMode="smooth"' if m != 1 else "") + ">\n"
                     f"      <Material>{1 + m}</Material>\n      <Transformations>{chain()}</Transformations>\n"
                     f'      <Faces vertexOffset="{b}">\n' + _rows(f + 1, "%d") + "      </Faces>\n    </Mesh>\n")
    inst = ""
    for k in range(4):
        mb = f"      <MotionBlur>{_f(rng.uniform(-0.4, 0.4, 3))}</MotionBlur>\n" if (stochastic and k == 1) else ""
        inst += (f'    <MeshInstance id="{k + 1}" baseMeshId="{1 + k % 3}" resetTransform="{"true" if k % 2 else "false"}">\n'
                 f"      <Material>{1 + int(rng.integers(0, 5))}</Material>\n{mb}"
                 f"      <Transformations>{chain()}</Transformations>\n    </MeshInstance>\n")
    # floor + back wall as stand-alone triangles, one of them transformed
    R = 6.0
    rb = push([[-R, -2.2, 2], [R, -2.2, 2], [R, -2.2, -10], [-R, -2.2, -10], [-R, 5, -9], [R, 5, -9], [R, -2.2, -9], [-R, -2.2, -9]])
    tri_xml = ""
    for i, t in enumerate(((0, 1, 2), (0, 2, 3), (7, 6, 5), (7, 5, 4))):
        tx = f"      <Transformations>r{1 + i}</Transformations>\n" if i == 3 else ""
        tri_xml += (f'    <Triangle id="{i + 1}">\n      <Material>{1 + (i % 2) * 3}</Material>\n{tx}'
                    f"      <Indices>{rb + t[0] + 1} {rb + t[1] + 1} {rb + t[2] + 1}</Indices>\n    </Triangle>\n")
    # spheres: plain, non-uniformly scaled, glass, and one AROUND the camera (t2 < 0 -> t1)
    cs = np.array([[-1.5, 0.2, -3.5], [0, 0, 0], [1.2, -0.4, -2.6], [0.0, 0.0, 1.0]], np.float32)
    cb = push(cs)
    sph = ""
    for i, (r, mat, tr) in enumerate(((0.7, 2, ""), (0.6, 1, chain()), (0.5, 3, ""), (float(camera_sphere_radius), 4, ""))):
        tx = f"      <Transformations>{tr}</Transformations>\n" if tr else ""
        mb = "      <MotionBlur>0.3 0.1 -0.2</MotionBlur>\n" if (stochastic and i == 0) else ""
        sph += (f'    <Sphere id="{i + 1}">\n      <Material>{mat}</Material>\n{tx}{mb}      <Center>{cb + i + 1}</Center>\n'
                f"      <Radius>{_f([r])}</Radius>\n    </Sphere>\n")
    allv = np.concatenate(verts, 0)
    aspect = width / height
    dof = "      <FocusDistance>4.5</FocusDistance>\n      <ApertureSize>0.08</ApertureSize>\n" if stochastic else ""
    area = ('    <AreaLight id="1">\n      <Position>-1 3.8 -3</Position>\n      <Intensity>1500 1400 1300</Intensity>\n'
            "      <EdgeVector1>1.5 0 0</EdgeVector1>\n      <EdgeVector2>0 0 1.5</EdgeVector2>\n    </AreaLight>\n") if stochastic else ""
    return ("<Scene>\n  <BackgroundColor>12 18 30</BackgroundColor>\n  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
            f"  <MaxRecursionDepth>{depth}</MaxRecursionDepth>\n  <Cameras>\n    <Camera id=\"1\">\n"
            "      <Position>0 0 1</Position>\n      <Gaze>0 0 -1</Gaze>\n      <Up>0 1 0</Up>\n"
            f"      <NearPlane>{_f([-0.8 * aspect, 0.8 * aspect, -0.8, 0.8])}</NearPlane>\n      <NearDistance>1</NearDistance>\n"
            f"      <ImageResolution>{width} {height}</ImageResolution>\n      <NumSamples>{spp}</NumSamples>\n" + dof
            + "      <ImageName>fuzz.png</ImageName>\n    </Camera>\n  </Cameras>\n  <Lights>\n"
            "    <AmbientLight>20 20 20</AmbientLight>\n"
            f'    <PointLight id="1">\n      <Position>{_f(rng.uniform(-3, 3, 3) * [1, 0.3, 0.5] + [0, 3.5, 0])}</Position>\n'
            "      <Intensity>2500 2400 2200</Intensity>\n    </PointLight>\n"
            f'    <PointLight id="2">\n      <Position>{_f(rng.uniform(-2, 2, 3) + [0, 1, 0.5])}</Position>\n'
            "      <Intensity>600 700 900</Intensity>\n    </PointLight>\n" + area + "  </Lights>\n  <Materials>\n" + mats
            + "  </Materials>\n" + xf + "  <VertexData>\n" + _rows(allv, "%.9g") + "  </VertexData>\n  <Objects>\n"
            + mesh_xml + inst + tri_xml + sph + "  </Objects>\n</Scene>\n")
